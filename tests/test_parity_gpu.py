"""End-to-end parity on the B200: the same C++ driver (tests/drivers/sim_driver.cpp, written
against the reference's public API) is run linked to the unmodified reference build (oracle
arm, host cores) and with libmeep_b200 in front of it (device arm); every field-like array,
DFT array and flux spectrum is compared with the north_star gates (rel-L2 <= 1e-12 double,
<= 1e-5 single).  Also runs the reference's own test programs through the drop-in."""
import os
import subprocess

import numpy as np
import pytest

from parity_util import TBUILD, TOL, compare, run_case

pytestmark = pytest.mark.gpu

CASES = [  # (case, steps, num_chunks)
    ("c2_3d_pml", 200, 0),            # BASELINE config 2 twin: "fields after 200 steps" 
    ("c2_3d_pml", 60, 5),
    ("c2_3d_pml_integrated", 60, 0),
    ("c2_3d_pml_complex", 40, 0),
    ("3d_metal", 60, 2),
    ("3d_bloch", 60, 0),
    ("3d_xperiodic_ypml", 60, 2),
    ("2d_bend_flux", 400, 0),          # BASELINE config 1 restated
    ("2d_bend_flux", 200, 4),
    ("2d_te_pml", 100, 2),
    ("1d_polariton", 200, 3),
    ("c3_au_sphere", 300, 0),          # BASELINE config 3 twin (stable: 10 nm pixels), 100-frequency flux box
    ("c3_au_sphere", 200, 8),
    ("lorentz_3d", 60, 3),
    ("c4_aniso_ring", 60, 0),          # BASELINE config 4 twin
    ("c4_aniso_ring", 30, 8),
    ("aniso_smooth", 30, 0),
    ("offdiag_2d", 100, 0),
    ("cond_chi3_3d", 60, 0),
    ("dft_fields_3d", 60, 2),
    ("2d_beta", 100, 0),
    ("2d_beta_real", 100, 3),
    ("2d_mirror_sym", 100, 2),
    ("3d_rotate_sym", 60, 0),
    ("noisy_lorentz_3d", 40, 0),
    ("gyro_lorentz_3d", 60, 0),
    ("gyro_drude_3d", 60, 3),
    ("gyro_saturated_3d", 60, 0),
    ("lorentz_aniso_sigma", 60, 0),
    ("3d_phase_in", 60, 0),
    ("3d_phase_in_cond", 60, 2),
    ("3d_bloch_change", 60, 2),
    ("3d_midrun_changes", 80, 0),
    ("3d_tiled", 40, 0),
    ("3d_sync_magnetic", 60, 2),
    ("3d_flux_planes", 60, 3),
    ("3d_bfast", 80, 0),
    ("2d_bfast", 150, 3),
    ("cyl_m0", 150, 0),
    ("cyl_m1", 150, 3),
    ("cyl_mneg1", 100, 0),
    ("cyl_m2", 150, 2),
    ("cyl_m3_nozero", 150, 0),
    ("cyl_m1_cond", 150, 0),
    ("cyl_m1_flux", 200, 2),
]


@pytest.mark.parametrize("case,steps,chunks", CASES)
def test_b200_matches_reference_f64(case, steps, chunks):
    ref = run_case("ref", "f64", case, steps, chunks)
    got = run_case("b200", "f64", case, steps, chunks)
    assert any((v != 0).any() for v in ref.values())
    rep = compare(got, ref, TOL["f64"])
    print(case, "worst group rel-L2 %.2e" % max(rep.values()))


@pytest.mark.parametrize("case,steps,chunks", [("c2_3d_pml", 200, 0), ("lorentz_3d", 60, 0),
                                               ("2d_bend_flux", 300, 0), ("3d_bloch", 60, 0),
                                               ("c4_aniso_ring", 40, 0), ("dft_fields_3d", 40, 0),
                                               ("cyl_m1_flux", 100, 0), ("gyro_lorentz_3d", 40, 0),
                                               ("3d_bfast", 40, 0), ("lorentz_aniso_sigma", 40, 0),
                                               ("cond_chi3_3d", 40, 0), ("c3_au_sphere", 200, 0),
                                               ("c3_au_sphere", 200, 8)])
def test_b200_matches_reference_f32(case, steps, chunks):
    ref = run_case("ref", "f32", case, steps, chunks)
    got = run_case("b200", "f32", case, steps, chunks)
    rep = compare(got, ref, TOL["f32"])
    print(case, "worst group rel-L2 %.2e" % max(rep.values()))


def test_unfused_path_matches_too():
    ref = run_case("ref", "f64", "c2_3d_pml", 40)
    got = run_case("b200", "f64", "c2_3d_pml", 40, env={"MEEP_B200_FUSE": "0"})
    compare(got, ref, TOL["f64"])


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("env", [{"MEEP_B200_PLAIN_LEAN": "0"}, {"MEEP_B200_SPLIT_PML": "3"},
                                 {"MEEP_B200_SPLIT_PML": "0"}, {"MEEP_B200_PML_PAIR": "1"}],
                         ids=["fast_path_masked_march", "pml_3_ctas_per_sm", "pml_three_components_per_thread", "pml_one_plane"])
def test_opt_in_kernel_forms_match_too(env, prec):
    """the kernel forms kept behind switches (A/B measurements in DESIGN.md section 4) stay correct"""
    ref = run_case("ref", prec, "c2_3d_pml", 60)
    got = run_case("b200", prec, "c2_3d_pml", 60, env=env)
    compare(got, ref, TOL[prec])


def test_chunk_count_invariance_on_device():
    """the reference's own invariant (tests/three_d.cpp): splitting into more chunks does not
    change the fields — checked through the point-probe path"""
    a = run_case("b200", "f64", "3d_metal", 50, 1)["probe.center"]
    b = run_case("b200", "f64", "3d_metal", 50, 6)["probe.center"]
    assert np.allclose(a, b, rtol=1e-9, atol=1e-14)


REFTESTS = ["known_results", "three_d", "two_dimensional", "one_dimensional", "physical", "integrate",
            "stress_tensor", "harmonics", "2D_convergence", "cylindrical", "flux", "symmetry", "near2far",
            "bragg_transmission"]


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("name", REFTESTS)
def test_reference_test_program_passes_on_the_device(name, prec):
    if prec == "f32" and name not in ("known_results", "three_d", "two_dimensional"):
        pytest.skip("the reference's CI runs only these in single precision here")
    exe = os.path.join(TBUILD, "reftest_%s_b200_%s" % (name, prec))
    if not os.path.exists(exe):
        pytest.skip("not built")
    env = dict(os.environ, OMP_NUM_THREADS="4")
    r = subprocess.run([exe], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:]


def test_ld_preload_over_a_program_linked_against_the_reference_only(tmp_path):
    """INTEGRATION.md route A on the real device: a binary linked against the reference library
    alone, with the drop-in preloaded"""
    import os
    import subprocess
    from parity_util import ROOT, driver, read_dump
    lib = os.path.join(ROOT, "meep_b200", "lib")
    exe = driver("sim_driver", "ref", "f64")
    out = str(tmp_path / "pre.bin")
    env = dict(os.environ, LD_PRELOAD=":".join([os.path.join(lib, "libmeep_b200_preload_f64.so"),
                                                os.path.join(lib, "libmeepb200.so")]),
               MEEP_B200_VERBOSE="1", OMP_NUM_THREADS="2")
    r = subprocess.run([exe, "c2_3d_pml", "40", out, "0"], env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "meep_b200: recorded" in r.stdout
    ref = run_case("ref", "f64", "c2_3d_pml", 40, 0)
    compare(read_dump(out), ref, TOL["f64"])
