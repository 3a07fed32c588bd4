"""Build everything in-tree.

Products (all git-ignored, all travel to the GPU box with gpurun):

  meep_b200/lib/libmeepb200.so          the C ABI (include/meep_b200.h): CUDA sm_100a, nvcc
  meep_b200/lib/libmeep_b200_<p>.so     host replacement TUs (meep_b200/host/*.cpp) compiled
                                        against the reference's unmodified meep.hpp; interposes
                                        the hot-path symbols of libmeep.  p = f64 | f32
  oracle/_ref/libmeep_ref_<p>.so        the unmodified reference as ORACLE / CPU baseline (tests, smoke, bench --impl reference)
  third_party/libmeep_host/libmeep_host_<p>.so   the same objects linked as the "installed Meep" the product sits on:
                                        the reference's non-hot-path host code (structure, fields set-up, readers)
  oracle/_build/liboracle.so            plain-C restatement of the inner loops (oracle/fdtd_oracle.c)
  tests/_build/*                        test-only emulator, parity drivers, golden generator

Steps that need the reference sources (/root/reference) are skipped when that tree is absent
(the GPU box): the prebuilt files are used.
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("MEEP_REFERENCE", "/root/reference")
GXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
GCC = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
PRECS = ("f64", "f32")

LIB = os.path.join(ROOT, "meep_b200", "lib")
CSRC = os.path.join(ROOT, "meep_b200", "csrc")
HOST = os.path.join(ROOT, "meep_b200", "host")
OREF = os.path.join(ROOT, "oracle", "_ref")
HOSTLIB = os.path.join(ROOT, "third_party", "libmeep_host")  # the "installed Meep" under the drop-in
OBUILD = os.path.join(ROOT, "oracle", "_build")
TBUILD = os.path.join(ROOT, "tests", "_build")

HOST_TUS = ["engine", "step", "step_db", "update_eh", "update_pols", "dft_hot", "hooks", "guards",
            "mympi_b200", "connect", "sync_magnetic", "hostmem", "materials_fill"]


def have_reference():
    return os.path.isdir(os.path.join(REF, "src"))


def _run(cmd, **kw):
    if os.environ.get("MEEP_B200_BUILD_VERBOSE"):
        print("+", " ".join(cmd), flush=True)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, **kw)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("build step failed: " + " ".join(cmd))
    return r.stdout


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(s) and os.path.getmtime(s) > t for s in sources)


def _glob(d, exts):
    out = []
    for base, _, files in os.walk(d):
        for f in files:
            if f.endswith(exts):
                out.append(os.path.join(base, f))
    return out


def build_cuda(force=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a -> libmeepb200.so (cross-compiles without a GPU)."""
    os.makedirs(LIB, exist_ok=True)
    out = os.path.join(LIB, "libmeepb200.so")
    srcs = _glob(CSRC, (".cu", ".cuh", ".h")) + [os.path.join(ROOT, "include", "meep_b200.h")]
    if force or _stale(out, srcs):
        _run([NVCC, "-ccbin", GXX, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
              "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
              os.path.join(CSRC, "capi.cu"), "-o", out])
    return out


def build_reference(precs=PRECS):
    """The unmodified reference stepping core -> oracle/_ref (recipe: oracle/ref_build/Makefile)."""
    outs = [os.path.join(OREF, "libmeep_ref_%s.so" % p) for p in precs] + \
        [os.path.join(HOSTLIB, "libmeep_host_%s.so" % p) for p in precs]
    if have_reference():
        _run(["make", "-C", os.path.join(ROOT, "oracle", "ref_build"), "-j8",
              "PRECS=" + " ".join(precs), "REF=" + REF])
    for o in outs:
        if not os.path.exists(o):
            raise RuntimeError("%s missing and the reference sources are not available" % o)
    return outs


def _ref_includes(prec):
    return ["-I" + os.path.join(OREF, "gen_" + prec), "-I" + os.path.join(ROOT, "oracle", "ref_build", "shim"),
            "-I" + os.path.join(REF, "src")]


def build_host(prec, backend="cuda"):
    """Host replacement TUs -> libmeep_b200_<prec>.so (backend 'emu' = test-only emulator)."""
    if backend == "cuda":
        out = os.path.join(LIB, "libmeep_b200_%s.so" % prec)
        link = ["-L" + LIB, "-lmeepb200", "-Wl,-rpath,$ORIGIN", "-Wl,-rpath,$ORIGIN/../../third_party/libmeep_host"]
    else:
        out = os.path.join(TBUILD, "libmeep_b200_emu_%s.so" % prec)
        link = ["-L" + TBUILD, "-lmeepb200_emu", "-Wl,-rpath,$ORIGIN", "-Wl,-rpath,$ORIGIN/../../third_party/libmeep_host"]
    if not have_reference():
        if not os.path.exists(out):
            raise RuntimeError("%s missing and the reference headers are not available" % out)
        return out
    objdir = os.path.join(ROOT, "build", "host_" + prec)
    os.makedirs(objdir, exist_ok=True)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    hdrs = _glob(HOST, (".hpp",)) + [os.path.join(ROOT, "include", "meep_b200.h")]
    objs = []
    for tu in HOST_TUS:
        src = os.path.join(HOST, tu + ".cpp")
        obj = os.path.join(objdir, tu + ".o")
        if _stale(obj, [src] + hdrs):
            _run([GXX, "-std=c++14", "-O2", "-fPIC", "-Wall", "-Wno-unused-variable"] + _ref_includes(prec) +
                 ["-I" + HOST, "-c", src, "-o", obj])
        objs.append(obj)
    if _stale(out, objs + [os.path.join(HOSTLIB, "libmeep_host_%s.so" % prec)]):
        _run([GXX, "-shared", "-fPIC", "-o", out] + objs + link +
             ["-L" + HOSTLIB, "-lmeep_host_" + prec, "-ldl"])
    # the same objects WITHOUT a dependency on any libmeep: the form to LD_PRELOAD over a program that
    # already links its own installed libmeep (INTEGRATION.md route A) — every non-hot-path symbol
    # then binds to that installation, and no second copy of the library enters the process
    pre = out.replace("libmeep_b200_", "libmeep_b200_preload_")
    if _stale(pre, objs):
        _run([GXX, "-shared", "-fPIC", "-o", pre] + objs + link + ["-ldl"])
    return out


def build_emulator():
    os.makedirs(TBUILD, exist_ok=True)
    out = os.path.join(TBUILD, "libmeepb200_emu.so")
    srcs = [os.path.join(ROOT, "tests", "emu", "capi_emu.cpp")] + _glob(CSRC, (".cuh", ".h")) + \
        [os.path.join(ROOT, "include", "meep_b200.h")]
    if _stale(out, srcs):
        _run([GXX, "-std=c++17", "-O2", "-fPIC", "-shared", "-Wall", "-Wno-unused-function",
              os.path.join(ROOT, "tests", "emu", "capi_emu.cpp"), "-o", out])
    return out


def build_oracle_c():
    """oracle/fdtd_oracle.c (plain C restatement of the inner loops) -> oracle/_build/liboracle.so"""
    os.makedirs(OBUILD, exist_ok=True)
    out = os.path.join(OBUILD, "liboracle.so")
    src = os.path.join(ROOT, "oracle", "fdtd_oracle.c")
    if not os.path.exists(src):
        return None
    if _stale(out, [src, os.path.join(ROOT, "oracle", "fdtd_oracle_impl.h")]):
        _run([GCC, "-std=c99", "-O2", "-ffp-contract=off", "-fPIC", "-shared", src, "-o", out, "-lm"])
    return out


def _meep_link(arm, prec):
    """which build of the reference a test program runs on: the ORACLE (arm 'ref': the unmodified
    reference does the time stepping) or the installed-Meep host library under the drop-in"""
    if arm == "ref":
        return ["-L" + OREF, "-lmeep_ref_" + prec, "-Wl,-rpath,$ORIGIN/../../oracle/_ref", "-ldl"]
    return ["-L" + HOSTLIB, "-lmeep_host_" + prec, "-Wl,-rpath,$ORIGIN/../../third_party/libmeep_host", "-ldl"]


def build_drivers(prec, arms=("ref", "b200", "emu")):
    """tests/drivers/*.cpp linked against the reference alone / the drop-in / the emulated drop-in."""
    os.makedirs(TBUILD, exist_ok=True)
    outs = {}
    drv_dir = os.path.join(ROOT, "tests", "drivers")
    for src in sorted(_glob(drv_dir, (".cpp",))):
        name = os.path.splitext(os.path.basename(src))[0]
        for arm in arms:
            if name == "gen_golden" and arm != "ref":
                continue
            out = os.path.join(TBUILD, "%s_%s_%s" % (name, arm, prec))
            outs[(name, arm)] = out
            if not have_reference():
                if not os.path.exists(out):
                    raise RuntimeError("%s missing and the reference headers are not available" % out)
                continue
            deps = [src, os.path.join(OREF, "libmeep_ref_%s.so" % prec), os.path.join(HOSTLIB, "libmeep_host_%s.so" % prec)]
            link = []
            if arm == "b200":
                deps.append(os.path.join(LIB, "libmeep_b200_%s.so" % prec))
                link = ["-L" + LIB, "-lmeep_b200_" + prec, "-lmeepb200", "-Wl,-rpath,$ORIGIN/../../meep_b200/lib"]
            elif arm == "emu":
                deps.append(os.path.join(TBUILD, "libmeep_b200_emu_%s.so" % prec))
                link = ["-L" + TBUILD, "-lmeep_b200_emu_" + prec, "-lmeepb200_emu", "-Wl,-rpath,$ORIGIN"]
            if _stale(out, deps):
                _run([GXX, "-std=c++14", "-O2", "-w", "-fopenmp"] + _ref_includes(prec) +
                     ["-I" + os.path.join(ROOT, "include"), "-I" + HOST, src, "-o", out, "-Wl,--no-as-needed"] + link +
                     _meep_link(arm, prec))
    return outs


# the reference's own C++ test programs (compiled from where they lie, never copied) that
# exercise the stepping path and need none of the absent libraries (libctl, HDF5, MPI, Harminv)
REFERENCE_TESTS = ["known_results", "three_d", "two_dimensional", "one_dimensional", "flux", "symmetry",
                   "harmonics", "pml", "physical", "integrate", "stress_tensor", "near2far",
                   "2D_convergence", "cylindrical", "bragg_transmission"]


def build_reference_tests(prec, arms=("ref", "b200", "emu"), names=None):
    os.makedirs(TBUILD, exist_ok=True)
    outs = {}
    for name in (names or REFERENCE_TESTS):
        src = os.path.join(REF, "tests", name + ".cpp")
        for arm in arms:
            out = os.path.join(TBUILD, "reftest_%s_%s_%s" % (name, arm, prec))
            outs[(name, arm)] = out
            if not have_reference():
                continue  # prebuilt (or absent: the test then skips)
            deps = [src, os.path.join(OREF, "libmeep_ref_%s.so" % prec), os.path.join(HOSTLIB, "libmeep_host_%s.so" % prec)]
            link = []
            if arm == "b200":
                deps.append(os.path.join(LIB, "libmeep_b200_%s.so" % prec))
                link = ["-L" + LIB, "-lmeep_b200_" + prec, "-lmeepb200", "-Wl,-rpath,$ORIGIN/../../meep_b200/lib"]
            elif arm == "emu":
                deps.append(os.path.join(TBUILD, "libmeep_b200_emu_%s.so" % prec))
                link = ["-L" + TBUILD, "-lmeep_b200_emu_" + prec, "-lmeepb200_emu", "-Wl,-rpath,$ORIGIN"]
            if _stale(out, deps):
                _run([GXX, "-std=c++11", "-O2", "-w", "-fopenmp"] + _ref_includes(prec) +
                     [src, "-o", out, "-Wl,--no-as-needed"] + link + _meep_link(arm, prec))
    return outs


def build_bench(prec):
    """bench/bench_driver.cpp -> in-process GPU arm (.so) and the CPU reference arm (executable)."""
    src = os.path.join(ROOT, "bench", "bench_driver.cpp")
    so = os.path.join(LIB, "libmeep_b200_bench_%s.so" % prec)
    exe = os.path.join(LIB, "bench_ref_%s" % prec)
    if not have_reference():
        for o in (so, exe):
            if not os.path.exists(o):
                raise RuntimeError("%s missing and the reference headers are not available" % o)
        return so, exe
    common = [GXX, "-std=c++14", "-O2", "-w", "-fopenmp"] + _ref_includes(prec)
    if _stale(so, [src, os.path.join(LIB, "libmeep_b200_%s.so" % prec), os.path.join(HOSTLIB, "libmeep_host_%s.so" % prec)]):
        _run(common + ["-fPIC", "-shared", src, "-o", so, "-Wl,--no-as-needed", "-L" + LIB,
                       "-lmeep_b200_" + prec, "-lmeepb200", "-L" + HOSTLIB, "-lmeep_host_" + prec,
                       "-Wl,-rpath,$ORIGIN", "-Wl,-rpath,$ORIGIN/../../third_party/libmeep_host", "-ldl"])
    if _stale(exe, [src, os.path.join(OREF, "libmeep_ref_%s.so" % prec)]):
        _run(common + ["-DMB200_BENCH_MAIN", src, "-o", exe, "-L" + OREF, "-lmeep_ref_" + prec,
                       "-Wl,-rpath,$ORIGIN/../../oracle/_ref", "-ldl"])
    return so, exe


def build_all(precs=PRECS, verbose=False):
    build_cuda()
    build_reference(precs)
    build_emulator()
    build_oracle_c()
    for p in precs:
        build_host(p, "cuda")
        build_host(p, "emu")
        build_drivers(p)
        build_reference_tests(p)
        build_bench(p)


if __name__ == "__main__":
    build_all()
    print("meep_b200: build complete")
