"""Host-side engine logic (job construction from meep::fields_chunk, connection-table
translation, lazy allocation, mirror protocol, plan caching) exercised WITHOUT a GPU: the
drop-in library is linked against the test-only emulator of the C ABI (tests/emu), and the full
simulations are compared array by array with the unmodified reference build (oracle/_ref)."""
import numpy as np
import os
import subprocess

import pytest

from parity_util import TBUILD, TOL, compare, run_case

CASES = [  # (case, steps, num_chunks)
    ("c2_3d_pml", 30, 0),
    ("c2_3d_pml", 20, 3),
    ("c2_3d_pml_integrated", 20, 0),
    ("c2_3d_pml_complex", 12, 0),
    ("3d_metal", 30, 2),
    ("3d_bloch", 30, 0),
    ("3d_xperiodic_ypml", 30, 0),
    ("2d_bend_flux", 150, 0),
    ("2d_te_pml", 60, 2),
    ("1d_polariton", 100, 3),
    ("c3_au_sphere", 60, 0),
    ("lorentz_3d", 30, 2),
    ("c4_aniso_ring", 15, 0),
    ("offdiag_2d", 60, 0),
    ("cond_chi3_3d", 30, 0),
    ("dft_fields_3d", 30, 0),
    ("2d_beta", 60, 0),
    ("2d_beta_real", 60, 2),
    ("2d_mirror_sym", 60, 0),
    ("3d_rotate_sym", 40, 2),
    ("noisy_lorentz_3d", 30, 3),
    ("gyro_lorentz_3d", 30, 0),
    ("gyro_drude_3d", 30, 3),
    ("gyro_saturated_3d", 30, 0),
    ("lorentz_aniso_sigma", 30, 4),
    ("3d_phase_in", 45, 2),
    ("3d_phase_in_cond", 24, 0),
    ("3d_bloch_change", 45, 3),
    ("3d_midrun_changes", 48, 2),
    ("3d_tiled", 30, 0),
    ("3d_sync_magnetic", 30, 0),
    ("3d_flux_planes", 40, 2),
    ("3d_bfast", 40, 0),
    ("2d_bfast", 80, 3),
    ("cyl_m0", 60, 0),
    ("cyl_m1", 60, 2),
    ("cyl_m2", 60, 0),
    ("cyl_m3_nozero", 60, 0),
    ("cyl_m1_cond", 60, 3),
    ("cyl_m1_flux", 80, 0),
]


@pytest.mark.parametrize("case,steps,chunks", CASES)
def test_emulated_dropin_matches_reference_f64(case, steps, chunks):
    ref = run_case("ref", "f64", case, steps, chunks)
    got = run_case("emu", "f64", case, steps, chunks)
    assert any((v != 0).any() for v in ref.values())
    compare(got, ref, TOL["f64"])


@pytest.mark.parametrize("case,steps,chunks", [("c2_3d_pml", 20, 0), ("lorentz_3d", 20, 0),
                                               ("2d_bend_flux", 100, 0), ("3d_bloch", 20, 0),
                                               ("cyl_m1", 40, 0), ("cyl_m1_cond", 40, 0),
                                               ("gyro_lorentz_3d", 30, 0), ("gyro_saturated_3d", 30, 0),
                                               ("3d_bfast", 30, 0), ("lorentz_aniso_sigma", 30, 0),
                                               ("noisy_lorentz_3d", 20, 0), ("3d_sync_magnetic", 20, 0),
                                               ("cond_chi3_3d", 20, 0), ("c4_aniso_ring", 12, 0)])
def test_emulated_dropin_matches_reference_f32(case, steps, chunks):
    ref = run_case("ref", "f32", case, steps, chunks)
    got = run_case("emu", "f32", case, steps, chunks)
    compare(got, ref, TOL["f32"])


@pytest.mark.parametrize("case,steps", [("c2_3d_pml", 20), ("lorentz_3d", 15)])
def test_fused_and_unfused_paths_agree(case, steps):
    ref = run_case("ref", "f64", case, steps)
    a = run_case("emu", "f64", case, steps, env={"MEEP_B200_FUSE": "0"})
    b = run_case("emu", "f64", case, steps, env={"MEEP_B200_FUSE": "1"})
    compare(a, ref, TOL["f64"])
    compare(b, ref, TOL["f64"])


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("case,steps,chunks", [("c2_3d_pml", 20, 0), ("3d_metal", 20, 1), ("3d_metal", 20, 6),
                                               ("3d_bloch", 20, 0), ("c2_3d_pml_integrated", 20, 0),
                                               ("3d_sync_magnetic", 20, 4)])
def test_lean_plus_shell_decomposition_of_the_fast_path(case, steps, chunks, prec):
    """The default form of the fast path: every job is carried out as the lean march over its full box +
    two x-slab jobs + the list of (y, z) shell columns (csrc/fused.cuh: step3_shell).  The emulator
    walks exactly those pieces, thread by thread: a point missed or visited twice, or a lean march
    whose arithmetic differed from the masked one, shows up against the reference."""
    ref = run_case("ref", prec, case, steps, chunks)
    got = run_case("emu", prec, case, steps, chunks, env={"MEEP_B200_PLAIN_LEAN": "1"})
    compare(got, ref, TOL[prec])
    # and the decomposition is bit-identical to the masked march (MEEP_B200_PLAIN_LEAN=0) of the same build
    same = run_case("emu", prec, case, steps, chunks, env={"MEEP_B200_PLAIN_LEAN": "0"})
    for k in same:
        assert np.array_equal(np.asarray(got[k]), np.asarray(same[k])), k


def test_eager_mirror_mode():
    ref = run_case("ref", "f64", "3d_metal", 10)
    got = run_case("emu", "f64", "3d_metal", 10, env={"MEEP_B200_EAGER": "1"})
    compare(got, ref, TOL["f64"])


@pytest.mark.parametrize("name", ["known_results", "three_d", "one_dimensional", "physical", "integrate",
                                  "stress_tensor", "cylindrical", "symmetry", "bragg_transmission"])
def test_reference_test_program_passes_through_the_dropin(name):
    """the reference's own C++ test programs, compiled from /root/reference/tests unmodified and
    linked with the (emulated) drop-in in front of libmeep"""
    exe = os.path.join(TBUILD, "reftest_%s_emu_f64" % name)
    if not os.path.exists(exe):
        pytest.skip("not built")
    env = dict(os.environ, OMP_NUM_THREADS="2")
    r = subprocess.run([exe], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:]


# ---- N > 1: one process per (emulated) device, launched like torchrun launches ranks -------------
from parity_util import run_case_mp  # noqa: E402

MP_CASES = [  # (case, steps, num_chunks, world_size)
    ("c2_3d_pml", 20, 2, 2),
    ("3d_bloch", 20, 3, 2),
    ("lorentz_3d", 15, 4, 2),
    ("2d_bend_flux", 100, 4, 3),
    ("c4_aniso_ring", 12, 2, 2),
    ("cyl_m1", 40, 4, 2),
    ("c2_3d_pml", 20, 8, 4),
    ("2d_mirror_sym", 40, 4, 2),
    ("3d_rotate_sym", 30, 4, 3),
    ("cyl_m1_flux", 60, 4, 2),
    ("3d_bfast", 30, 4, 2),
    ("2d_beta", 40, 6, 3),
    ("dft_fields_3d", 30, 4, 2),
    ("c3_au_sphere", 15, 8, 4),
    ("cond_chi3_3d", 20, 4, 2),
    ("3d_sync_magnetic", 20, 4, 2),
    ("1d_polariton", 60, 4, 2),
    ("lorentz_aniso_sigma", 30, 4, 2),
    ("gyro_lorentz_3d", 30, 4, 2),
    ("3d_xperiodic_ypml", 20, 6, 3),
]


@pytest.mark.parametrize("case,steps,chunks,world", MP_CASES)
def test_sharded_multi_process_run_matches_single_process_reference(case, steps, chunks, world):
    """chunks distributed over `world` processes by the reference's own split_by_cost /
    is_mine() logic; rank, size and the small host reductions come from the MPI-free runtime
    (meep_b200/host/mympi_b200.cpp, TCP on 127.0.0.1), comm blocks are packed by halo jobs straight
    into the neighbour's arena (peer-memory protocol of DESIGN §5; the emulator maps the arenas as
    shared memory where the CUDA build uses CUDA IPC) and unpacked after the sequence-word
    handshake; every array of every rank must match the single-process reference run."""
    ref = run_case("ref", "f64", case, steps, chunks)
    got = run_case_mp("emu", "f64", case, steps, chunks, world)
    compare(got, ref, TOL["f64"])


@pytest.mark.parametrize("case,steps,chunks,world", [("c2_3d_pml", 20, 4, 4), ("2d_bend_flux", 60, 5, 3)])
def test_sharded_run_with_the_fallback_transport(case, steps, chunks, world):
    """MEEP_B200_P2P=0: comm blocks travel as separate transfers (NCCL on GPUs, the socket runtime
    under the emulator) between a pack and an unpack launch"""
    ref = run_case("ref", "f64", case, steps, chunks)
    got = run_case_mp("emu", "f64", case, steps, chunks, world, env={"MEEP_B200_P2P": "0"})
    compare(got, ref, TOL["f64"])


@pytest.mark.parametrize("world", [1, 2, 3, 5])
def test_process_runtime_reductions_and_broadcasts(world):
    """the reference's src/mympi.cpp entry points (sum_to_all, and_to_all, partial_sum_to_all,
    broadcast, ...) served by the MPI-free socket runtime, `world` cooperating processes"""
    import socket
    from parity_util import driver
    exe = driver("comm_driver", "emu", "f64")
    from parity_util import free_ports
    port, rt_port = free_ports(2)
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), MEEP_B200_PORT=str(rt_port))
        procs.append(subprocess.Popen([exe], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for r, p in enumerate(procs):
        out, _ = p.communicate(timeout=120)
        assert p.returncode == 0, "rank %d:\n%s" % (r, out[-2000:])


def test_ld_preload_over_a_program_linked_against_the_reference_only(tmp_path):
    """INTEGRATION.md route A: an existing binary that knows nothing about the drop-in (linked
    against the reference library alone) runs on the (emulated) device when the drop-in is
    preloaded, and produces the reference's arrays"""
    from parity_util import TBUILD, driver, read_dump
    exe = driver("sim_driver", "ref", "f64")
    pre = ":".join([os.path.join(TBUILD, "libmeep_b200_preload_emu_f64.so"), os.path.join(TBUILD, "libmeepb200_emu.so")])
    out = str(tmp_path / "pre.bin")
    env = dict(os.environ, LD_PRELOAD=pre, MEEP_B200_VERBOSE="1", OMP_NUM_THREADS="2")
    r = subprocess.run([exe, "3d_metal", "12", out, "2"], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "meep_b200: recorded" in r.stdout  # the engine, not the CPU loops, did the stepping
    ref = run_case("ref", "f64", "3d_metal", 12, 2)
    compare(read_dump(out), ref, TOL["f64"])


CONNECT_LAYOUTS = ["pml27", "pml64", "pml72", "pml27_complex", "metal5", "bloch4", "periodic_k0", "xperiodic_ypml",
                   "mirror2d", "rotate3d", "cyl3", "gyro3", "aniso_sigma4", "bend2d", "1d3"]


@pytest.mark.parametrize("layout", CONNECT_LAYOUTS)
def test_connection_tables_identical_to_the_reference(layout, tmp_path):
    """the analytic (run-based) connect_the_chunks / find_metals of meep_b200/host/connect.cpp build
    the SAME tables as the reference's per-point versions (src/boundaries.cpp:315-638): every entry
    of connections_in/out (as array id + offset, in order), connection_phases, comm_sizes and
    zeroes, on the SURVEY 8e layouts (27 / 64 / 72 chunks) and on periodic, Bloch, symmetric,
    cylindrical, metallic, 1-D/2-D and polarisation-carrying cells"""
    import numpy as np
    from parity_util import driver, read_dump
    outs = {}
    for arm in ("ref", "emu"):
        out = str(tmp_path / ("tables_%s.bin" % arm))
        r = subprocess.run([driver("connect_driver", arm, "f64"), layout, out], stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, text=True, timeout=600,
                           env=dict(os.environ, OMP_NUM_THREADS="2"))
        assert r.returncode == 0, r.stdout[-2000:]
        outs[arm] = read_dump(out)
    ref, got = outs["ref"], outs["emu"]
    # the reference's `comm_sizes[key] += 0` leaves zero-size entries behind (src/boundaries.cpp:445-446);
    # get_comm_size() reads a missing key as 0, so they carry no information
    for d in (ref, got):
        rows = d["comm_sizes"].reshape(-1, 5)
        d["comm_sizes"] = rows[rows[:, 4] > 0].ravel()
    # a key the reference left behind with an empty vector carries no information
    ref = {k: v for k, v in ref.items() if v.size or k == "comm_sizes"}
    got = {k: v for k, v in got.items() if v.size or k == "comm_sizes"}
    assert set(ref) == set(got), (sorted(set(ref) - set(got))[:5], sorted(set(got) - set(ref))[:5])
    assert sum(v.size for k, v in ref.items() if ".in." in k) > 0
    for k in ref:
        assert ref[k].shape == got[k].shape, k
        assert not (ref[k] < 0).any() or "phases" in k, "unresolved pointer in %s" % k
        assert np.array_equal(ref[k], got[k]), k


def test_phase_timers_are_credited_from_device_marks():
    """MEEP_B200_TIMERS=1 (or verbosity > 1): fields::time_spent_on reports per-phase DEVICE time taken
    from marks on the engine's stream (include/meep_b200.h: mb200_mark), not host launch latency"""
    d = run_case("emu", "f64", "c2_3d_pml", 12, env={"MEEP_B200_TIMERS": "1", "MB200_DUMP_TIMES": "1"})
    t = d["times"]
    stepping, bnd, ft, ub, uh, ud, ue = t[:7]
    wall = t[-1]
    assert ub > 0 and ud > 0 and bnd > 0
    # the per-phase sinks hold device time only: they add up to no more than the wall time of the loop
    # (Boundaries also holds the host time of connect_the_chunks, as in the reference)
    assert ub + uh + ud + ue <= wall * 1.05
    d0 = run_case("emu", "f64", "c2_3d_pml", 12, env={"MB200_DUMP_TIMES": "1"})
    assert d0["times"].shape == t.shape
