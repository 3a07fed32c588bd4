"""Property test of the lean + shell decomposition of the fast path on the CPU (tests/emu/lean_property.cpp):
random fast-path jobs — sizes, owned ranges, difference directions, metal planes, x-slabs, planes without the
epilogue, planes per CTA — carried out by the masked march and by lean march + slab jobs + shell columns must give
bit-identical arrays (every point exactly once, same arithmetic).  Host compile of the very headers the CUDA
build uses; no GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exe():
    out = os.path.join(ROOT, "tests", "_build", "lean_property")
    src = os.path.join(ROOT, "tests", "emu", "lean_property.cpp")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    deps = [src] + [os.path.join(ROOT, "meep_b200", "csrc", f) for f in ("fused.cuh", "kernels.cuh", "point_ops.h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-w", src, "-o", out], check=True)
    return out


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_lean_plus_shell_equals_the_masked_march_on_random_jobs(exe, seed):
    r = subprocess.run([exe, str(seed), "250"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:]
    words = r.stdout.split()
    assert words[0] == "OK" and int(words[2]) > 150, r.stdout  # (most random jobs do have a full box)
