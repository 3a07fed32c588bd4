"""The plain-C oracle (oracle/fdtd_oracle.c) against golden vectors produced by the reference's
own functions (tests/drivers/gen_golden.cpp -> tests/golden/)."""
import pytest

from kernel_cases import HostMem, check_golden


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("kind", ["curl", "beta", "edhb", "lorentz", "dft", "cyl", "bfast", "gyro"])
def test_oracle_matches_reference_golden(kind, prec):
    worst = check_golden(kind, prec, HostMem)
    print("oracle vs reference golden: %s/%s worst rel err %.2e" % (kind, prec, worst))
