"""N > 1 on real GPUs: the cell sharded over 2 processes / 2 B200s (one process per GPU, launched
with torchrun-style environment), comm blocks packed straight into the neighbour GPU's HBM over
NVLink (CUDA IPC peer memory; MEEP_B200_P2P=0: NCCL send/recv), compared array by array with the
single-process reference run."""
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
                                                     ("lorentz_3d", 30, 2, 2), ("c4_aniso_ring", 20, 4, 2),
                                                     ("cyl_m1", 60, 4, 2), ("c3_au_sphere", 120, 8, 2),
                                                     ("3d_sync_magnetic", 40, 4, 2), ("dft_fields_3d", 40, 2, 2)])
@pytest.mark.parametrize("transport", ["peer", "nccl"])
def test_two_gpu_sharded_run_matches_reference(case, steps, chunks, world, transport):
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    if transport == "nccl" and case not in ("c2_3d_pml", "3d_bloch"):
        pytest.skip("NCCL fallback: two cases are enough")
    ref = run_case("ref", "f64", case, steps, chunks)
    got = run_case_mp("b200", "f64", case, steps, chunks, world,
                      env={"MEEP_B200_P2P": "1" if transport == "peer" else "0"})
    rep = compare(got, ref, TOL["f64"])
    print(case, "worst group rel-L2 %.2e" % max(rep.values()))


@pytest.mark.parametrize("case,steps,chunks", [("c2_3d_pml", 40, 2), ("3d_bloch", 40, 4), ("c3_au_sphere", 60, 8)])
def test_two_ranks_sharing_one_gpu_match_reference(case, steps, chunks):
    """The cross-process exchange on a ONE-GPU box (the driver's test box): two ranks, both on device 0,
    each packing its comm blocks into the other PROCESS's device memory (CUDA IPC works between
    processes on the same device exactly as between devices) and signalling through the same flag
    words.  Slower than two GPUs — the two contexts time-slice the device while one of them spins
    on a flag — but the code path, the tables and the values are those of the multi-GPU run."""
    if _ngpu() < 1:
        pytest.skip("needs a GPU")
    ref = run_case("ref", "f64", case, steps, chunks)
    got = run_case_mp("b200", "f64", case, steps, chunks, 2,
                      env={"MEEP_B200_P2P": "1", "MEEP_B200_DEVICE": "0", "MEEP_B200_PEER_TIMEOUT_S": "60"},
                      timeout=600)
    rep = compare(got, ref, TOL["f64"])
    print(case, "worst group rel-L2 %.2e" % max(rep.values()))
