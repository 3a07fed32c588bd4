"""N > 1 on real GPUs: the cell sharded over 2 processes / 2 B200s (one process per GPU, launched
with torchrun-style environment), comm blocks moved device-to-device with NCCL, compared array by
array with the single-process reference run."""
import pytest

from meep_b200 import capi
from parity_util import TOL, compare, run_case, run_case_mp

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        return capi.load().mb200_device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("case,steps,chunks,world", [("c2_3d_pml", 40, 2, 2), ("3d_bloch", 40, 4, 2),
                                                     ("lorentz_3d", 30, 2, 2), ("c4_aniso_ring", 20, 4, 2)])
def test_two_gpu_sharded_run_matches_reference(case, steps, chunks, world):
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    ref = run_case("ref", "f64", case, steps, chunks)
    got = run_case_mp("b200", "f64", case, steps, chunks, world)
    rep = compare(got, ref, TOL["f64"])
    print(case, "worst group rel-L2 %.2e" % max(rep.values()))
