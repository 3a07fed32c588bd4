"""The C-ABI library loads, exports every symbol include/meep_b200.h declares, its structs have
the layout the Python binding assumes, and it FAILS LOUDLY without a CUDA device (no fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from conftest import have_gpu
from meep_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "meep_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mb200_[a-zA-Z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    lib = capi.load()
    names = declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), "libmeepb200.so does not export %s" % n
    # and the Python binding declares every one of them
    assert set(capi.declare(lib)) == set(names)


def test_struct_layouts_match_the_header():
    """compile a tiny C program that prints sizeof() of every job struct and compare with ctypes"""
    names = ["mb200_box_t", "mb200_pml_t", "mb200_curl_job_t", "mb200_edhb_job_t", "mb200_lorentz_job_t",
             "mb200_fmp_job_t", "mb200_src_job_t", "mb200_halo_job_t", "mb200_zero_job_t", "mb200_dft_job_t",
             "mb200_flux_job_t", "mb200_step3_comp_t", "mb200_step3_job_t", "mb200_beta_job_t",
             "mb200_cylint_job_t", "mb200_cylr0_job_t", "mb200_xfer_t", "mb200_bfast_job_t", "mb200_halo_run_t", "mb200_average_job_t", "mb200_gyro_job_t", "mb200_noise_job_t"]
    types = [capi.Box, capi.Pml, capi.CurlJob, capi.EdhbJob, capi.LorentzJob, capi.FmpJob, capi.SrcJob,
             capi.HaloJob, capi.ZeroJob, capi.DftJob, capi.FluxJob, capi.Step3Comp, capi.Step3Job,
             capi.BetaJob, capi.CylIntJob, capi.CylR0Job, capi.Xfer, capi.BfastJob, capi.HaloRun, capi.AverageJob, capi.GyroJob, capi.NoiseJob]
    prog = '#include <stdio.h>\n#include "meep_b200.h"\nint main(void){%s return 0;}\n' % "".join(
        'printf("%%zu\\n", sizeof(%s));' % n for n in names)
    d = os.path.join(ROOT, "tests", "_build")
    os.makedirs(d, exist_ok=True)
    src, exe = os.path.join(d, "sizes.c"), os.path.join(d, "sizes")
    open(src, "w").write(prog)
    subprocess.check_call(["/usr/bin/gcc", "-I" + os.path.join(ROOT, "include"), src, "-o", exe])
    sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    assert sizes == [C.sizeof(t) for t in types]


def test_abi_version():
    assert capi.load().mb200_abi_version() == 2


@pytest.mark.skipif(have_gpu(), reason="this check is for machines without a GPU")
def test_fails_loudly_without_a_device():
    lib = capi.load()
    assert lib.mb200_device_count() == 0
    ctx = C.c_void_p()
    assert lib.mb200_init(0, C.byref(ctx)) != 0
    assert b"cuda" in lib.mb200_last_error().lower() or b"device" in lib.mb200_last_error().lower()
    with pytest.raises(capi.Error):
        capi.Context(0)


@pytest.mark.skipif(have_gpu(), reason="this check is for machines without a GPU")
def test_dropin_has_no_cpu_time_stepping_path():
    """fields::step() through the real drop-in aborts with a clear message when no GPU exists"""
    exe = os.path.join(ROOT, "tests", "_build", "sim_driver_b200_f64")
    r = subprocess.run([exe, "3d_metal", "2", "/tmp/mb200_nogpu.bin"], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True)
    assert r.returncode != 0
    assert "no CUDA device" in r.stdout


def test_product_libraries_do_not_link_the_oracle_or_the_emulator():
    for name in ["libmeepb200.so", "libmeep_b200_f64.so", "libmeep_b200_f32.so"]:
        out = subprocess.check_output(["readelf", "-d", os.path.join(capi.LIBDIR, name)], text=True)
        assert "liboracle" not in out and "emu" not in out, name
    syms = subprocess.check_output(["nm", "-D", "--defined-only", os.path.join(capi.LIBDIR, "libmeepb200.so")],
                                   text=True)
    assert "mb200_is_emulator" not in syms


def test_dropin_interposes_the_hot_path_symbols():
    """the replacement library defines exactly the hot-path symbols SURVEY §8b lists"""
    syms = subprocess.check_output(
        ["nm", "-D", "--defined-only", "-C", os.path.join(capi.LIBDIR, "libmeep_b200_f64.so")], text=True)
    for want in ["meep::fields::step()", "meep::fields::step_db(", "meep::fields_chunk::step_db(",
                 "meep::fields::update_eh(", "meep::fields_chunk::update_eh(", "meep::fields::update_pols(",
                 "meep::fields_chunk::update_pols(", "meep::fields::step_boundaries(",
                 "meep::fields::step_source(", "meep::fields::calc_sources(", "meep::fields::phase_material()",
                 "meep::fields::process_incoming_chunk_data(", "meep::fields::update_dfts()",
                 "meep::dft_chunk::update_dft(", "meep::dft_flux::flux()",
                 "meep::lorentzian_susceptibility::update_P(", "meep::lorentzian_susceptibility::subtract_P(",
                 "meep::step_curl(", "meep::step_update_EDHB(", "meep::step_curl_stride1(",
                 "meep::fields_chunk::needs_W_prev("]:
        assert want in syms, want


# ---- the link-level seam: every hot-path symbol must bind to the drop-in -----------------
HOT_PATH_PATTERNS = [
    r"^meep::fields::step\(\)$", r"^meep::fields::step_boundaries\(", r"^meep::fields::process_incoming_chunk_data\(",
    r"^meep::fields::step_source\(", r"^meep::fields::calc_sources\(", r"^meep::fields::phase_material\(\)$",
    r"^meep::fields_chunk::step_source\(", r"^meep::fields_chunk::calc_sources\(", r"^meep::fields_chunk::phase_material\(",
    r"^meep::fields::step_db\(", r"^meep::fields_chunk::step_db\(",
    r"^meep::step_curl(_stride1)?\(", r"^meep::step_update_EDHB(_stride1)?\(", r"^meep::step_beta(_stride1)?\(",
    r"^meep::step_bfast(_stride1)?\(",
    r"^meep::fields::update_eh\(", r"^meep::fields_chunk::update_eh\(", r"^meep::fields_chunk::needs_W_prev\(",
    r"^meep::fields::update_pols\(", r"^meep::fields_chunk::update_pols\(",
    r"^meep::lorentzian_susceptibility::update_P\(", r"^meep::lorentzian_susceptibility::subtract_P\(",
    r"^meep::dft_chunk::update_dft\(", r"^meep::fields_chunk::update_dfts\(", r"^meep::fields::update_dfts\(\)$",
    r"^meep::dft_flux::flux\(\)$",
    r"^meep::fields::connect_the_chunks\(\)$", r"^meep::fields::find_metals\(\)$",
]


def _defined_text_symbols(lib):
    """{demangled: mangled} of the functions a shared library defines"""
    import subprocess
    raw = subprocess.run(["nm", "-D", "--defined-only", lib], stdout=subprocess.PIPE, text=True, check=True).stdout
    dem = subprocess.run(["nm", "-D", "-C", "--defined-only", lib], stdout=subprocess.PIPE, text=True, check=True).stdout
    out = {}
    for a, b in zip(raw.splitlines(), dem.splitlines()):
        pa, pb = a.split(None, 2), b.split(None, 2)
        if len(pa) == 3 and pa[1] in "TW":
            out[pb[2]] = pa[2]
    return out


def _hot_path_symbols(dropin, hostlib):
    import re
    d, h = _defined_text_symbols(dropin), _defined_text_symbols(hostlib)
    hot = {}
    for pat in HOT_PATH_PATTERNS:
        names = [n for n in h if re.search(pat, n)]
        assert names, "the host library defines nothing matching %s" % pat
        for n in names:
            assert n in d, "hot-path symbol %s is NOT defined by the drop-in" % n
            hot[n] = d[n]
    return hot


@pytest.mark.parametrize("route", ["link_order", "ld_preload"])
def test_hot_path_symbols_bind_to_the_dropin(route):
    """dladdr of what the process actually resolves, for every symbol of the link-level seam"""
    import subprocess
    tb = os.path.join(ROOT, "tests", "_build")
    hostlib = os.path.join(ROOT, "third_party", "libmeep_host", "libmeep_host_f64.so")
    dropin = os.path.join(tb, "libmeep_b200_emu_f64.so")
    hot = _hot_path_symbols(dropin, hostlib)
    assert len(hot) >= 30
    env = dict(os.environ)
    if route == "link_order":
        exe = os.path.join(tb, "binding_driver_emu_f64")
    else:  # a program that only knows the reference library, drop-in preloaded
        exe = os.path.join(tb, "binding_driver_ref_f64")
        env["LD_PRELOAD"] = ":".join([os.path.join(tb, "libmeep_b200_preload_emu_f64.so"),
                                      os.path.join(tb, "libmeepb200_emu.so")])
    r = subprocess.run([exe], input="\n".join(hot.values()) + "\n", env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=120)
    assert r.returncode == 0, r.stdout[-2000:]
    got = dict(line.split() for line in r.stdout.splitlines() if line.startswith("_Z"))
    assert set(got) == set(hot.values())
    wrong = {s: lib for s, lib in got.items() if not lib.startswith("libmeep_b200")}
    assert not wrong, "hot-path symbols bound outside the drop-in: %s" % wrong


def test_ld_debug_shows_no_hot_path_binding_into_the_reference(tmp_path):
    """LD_DEBUG=bindings of a real (emulated-device) simulation: whenever any object binds one of the
    hot-path symbols, the target is the drop-in, never the host/reference library"""
    import re
    import subprocess
    tb = os.path.join(ROOT, "tests", "_build")
    hostlib = os.path.join(ROOT, "third_party", "libmeep_host", "libmeep_host_f64.so")
    hot = set(_hot_path_symbols(os.path.join(tb, "libmeep_b200_emu_f64.so"), hostlib).values())
    log = str(tmp_path / "ld")
    env = dict(os.environ, LD_DEBUG="bindings", LD_DEBUG_OUTPUT=log, OMP_NUM_THREADS="1")
    r = subprocess.run([os.path.join(tb, "sim_driver_emu_f64"), "3d_metal", "3", str(tmp_path / "o.bin"), "2"], env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:]
    seen, bad = 0, []
    for f in os.listdir(tmp_path):
        if not f.startswith("ld."):
            continue
        for line in open(os.path.join(tmp_path, f), errors="replace"):
            m = re.search(r"binding file (\S+) \[\d+\] to (\S+) \[\d+\]: normal symbol `([^']+)'", line)
            if m and m.group(3) in hot:
                seen += 1
                if "libmeep_b200" not in os.path.basename(m.group(2)):
                    bad.append((m.group(1), m.group(2), m.group(3)))
    assert seen > 0, "no hot-path binding was logged at all"
    assert not bad, bad[:5]
