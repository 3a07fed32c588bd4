"""Helpers shared by the parity tests: run a parity driver arm, parse its dump, compare arrays."""
import os
import struct
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TBUILD = os.path.join(ROOT, "tests", "_build")

# parity gates stated by BASELINE.json's north_star: relative L2 over whole arrays
TOL = {"f64": 1e-12, "f32": 1e-5}


def read_dump(path):
    """-> dict name -> np.ndarray (records: u32 name_len, name, u32 elem_size, u64 count, data)."""
    out = {}
    with open(path, "rb") as fh:
        data = fh.read()
    off = 0
    while off < len(data):
        (nl,) = struct.unpack_from("<I", data, off)
        off += 4
        name = data[off:off + nl].decode()
        off += nl
        es, cnt = struct.unpack_from("<IQ", data, off)
        off += 12
        dt = {8: np.float64, 4: np.float32}[es]
        out[name] = np.frombuffer(data, dtype=dt, count=cnt, offset=off).copy()
        off += es * cnt
    return out


def driver(name, arm, prec):
    return os.path.join(TBUILD, "%s_%s_%s" % (name, arm, prec))


def run_case(arm, prec, case, nsteps, num_chunks=0, env=None, name="sim_driver", timeout=900):
    if arm == "b200":  # dry-run of the GPU test code against the emulator on a GPU-less box
        arm = os.environ.get("MEEP_B200_TEST_ARM", arm)
    exe = driver(name, arm, prec)
    if not os.path.exists(exe):
        raise FileNotFoundError(exe)
    fd, path = tempfile.mkstemp(suffix=".bin", prefix="mb200_%s_%s_" % (case, arm))
    os.close(fd)
    e = dict(os.environ)
    e.setdefault("OMP_NUM_THREADS", "4")
    if env:
        e.update(env)
    try:
        r = subprocess.run([exe, case, str(nsteps), path, str(num_chunks)], env=e, timeout=timeout,
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("%s %s failed (rc=%d):\n%s" % (exe, case, r.returncode, r.stdout[-4000:]))
        return read_dump(path)
    finally:
        if os.path.exists(path):
            os.unlink(path)


def free_ports(n=2):
    """n distinct TCP ports on 127.0.0.1 that are free right now (all held open together, then
    released): MASTER_PORT for the launcher's contract and MEEP_B200_PORT for the process runtime —
    deriving the second from the first (MASTER_PORT + 37, the runtime's default) now and then hit a
    port somebody else was using"""
    import socket
    socks = []
    try:
        for _ in range(n):
            s = socket.socket()
            s.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
            s.bind(("127.0.0.1", 0))
            socks.append(s)
        return [s.getsockname()[1] for s in socks]
    finally:
        for s in socks:
            s.close()


def run_case_mp(arm, prec, case, nsteps, num_chunks, world=2, env=None, name="sim_driver", timeout=900):
    """the same driver as `world` cooperating processes (one per device), launched the way torchrun
    launches ranks: RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT in the environment.
    Returns the union of the per-rank dumps (each rank dumps the chunks it owns)."""
    import socket
    if arm == "b200":
        arm = os.environ.get("MEEP_B200_TEST_ARM", arm)
    exe = driver(name, arm, prec)
    fd, path = tempfile.mkstemp(suffix=".bin", prefix="mb200_%s_%s_mp_" % (case, arm))
    os.close(fd)
    port, rt_port = free_ports(2)
    procs = []
    for r in range(world):
        e = dict(os.environ)
        e.setdefault("OMP_NUM_THREADS", "2")
        e.update({"RANK": str(r), "WORLD_SIZE": str(world), "LOCAL_RANK": str(r),
                  "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "MEEP_B200_PORT": str(rt_port)})
        if env:
            e.update(env)
        procs.append(subprocess.Popen([exe, case, str(nsteps), path, str(num_chunks)], env=e,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    out = {}
    try:
        for r, p in enumerate(procs):
            so, _ = p.communicate(timeout=timeout)
            if p.returncode != 0:
                raise RuntimeError("rank %d of %s %s failed (rc=%d):\n%s" % (r, exe, case, p.returncode, so[-4000:]))
        for r in range(world):
            d = read_dump(path + ".rank%d" % r)
            for k, v in d.items():
                if k in out and k.startswith("chunk"):
                    raise RuntimeError("array %s dumped by two ranks" % k)
                out[k] = v
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
        for r in range(world):
            if os.path.exists(path + ".rank%d" % r):
                os.unlink(path + ".rank%d" % r)
        if os.path.exists(path):
            os.unlink(path)
    return out


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    nb = np.linalg.norm(b)
    d = np.linalg.norm(a - b)
    if nb == 0:
        return 0.0 if d == 0 else float("inf")
    return d / nb


def _group(name):
    """chunk3.f_u.bz.1 -> 'f_u.b' : one group per (array kind, field type), all chunks, all
    components, both cmp.  The north_star gate is a relative L2 over whole fields; a component
    that is zero by symmetry (e.g. Bz of an Ez dipole) only holds rounding noise of its
    siblings and is judged together with them, as SURVEY §7 / BASELINE.md prescribe."""
    parts = name.split(".")
    if not parts[0].startswith("chunk"):
        return name
    kind = parts[1]
    if kind.startswith("dft"):
        return "dft." + parts[2][0]
    if kind.startswith("P"):
        return kind.rstrip("0123456789") + "." + parts[2][0]
    return kind + "." + parts[2][0]


def _ftype(group):
    """field-type pool of a group: 'f_u.b' -> 'b' (f, f_u, f_cond of B share units and scale)."""
    return group.split(".")[-1] if "." in group else group


def _sumsq_scaled(x, scale):
    """sum((x/scale)^2): cannot overflow for finite x when scale >= max|x|"""
    if scale == 0:
        return 0.0
    y = x / scale
    return float(np.dot(y, y))


MAX_REF_MAGNITUDE = 1e100  # a reference run that grew beyond this is diverging: not a parity case


def compare(got, ref, tol, skip_prefix=()):
    """Gate: for every group g (see _group)

        ||got_g - ref_g||_2  <=  (tol + 64 * eps) * ||ref_g||_2

    i.e. the north_star relative-L2 tolerance on the group's OWN norm plus a rounding allowance of
    64 ulp of that same norm.  Only a group that holds NO signal in the reference — exactly zero,
    or at the level of the rounding noise of its field-type pool T over all array kinds,
    ||ref_g||_2 <= 64 * eps * ||ref_T||_2 (e.g. the f_u array of a B component that is zero by
    symmetry: 1e-16 of |B| in the reference itself) — is judged against that pool instead:
    ||got_g - ref_g||_2 <= 64 * eps * ||ref_T||_2.
    The reference itself must be a healthy run: all values finite and max|ref| < 1e100 (a diverging
    simulation would make any gate vacuous).  Norms are formed on values scaled by the group's
    largest magnitude so they cannot overflow.  Returns {group: rel err}."""
    assert set(got) == set(ref), "array sets differ: only-got=%s only-ref=%s" % (
        sorted(set(got) - set(ref))[:8], sorted(set(ref) - set(got))[:8])
    keys = [k for k in sorted(ref) if not any(k.startswith(p) for p in skip_prefix)]
    eps = None
    groups = {}
    for k in keys:
        if eps is None or ref[k].dtype == np.float32:
            eps = float(np.finfo(ref[k].dtype).eps) if k.startswith("chunk") else eps
        a, b = got[k].astype(np.float64), ref[k].astype(np.float64)
        assert a.shape == b.shape, k
        assert np.all(np.isfinite(b)), "non-finite values in the REFERENCE array %s" % k
        assert np.all(np.isfinite(a)), "non-finite values in %s" % k
        mb = float(np.abs(b).max()) if b.size else 0.0
        assert mb < MAX_REF_MAGNITUDE, "reference array %s has max|value| = %g: the run diverges" % (k, mb)
        groups.setdefault(_group(k), []).append((a, b, max(mb, float(np.abs(a).max()) if a.size else 0.0)))
    eps = eps or float(np.finfo(np.float64).eps)
    norms = {}
    for g, items in groups.items():
        scale = max(it[2] for it in items)
        num = sum(_sumsq_scaled(a - b, scale) for a, b, _ in items)
        den = sum(_sumsq_scaled(b, scale) for a, b, _ in items)
        norms[g] = (np.sqrt(num) * scale, np.sqrt(den) * scale)
    pool = {}
    for g, (dn, rn) in norms.items():
        pool[_ftype(g)] = float(np.hypot(pool.get(_ftype(g), 0.0), rn))
    report, bad = {}, []
    for g, (dn, rn) in norms.items():
        if rn > 64 * eps * pool[_ftype(g)]:
            allowed = (tol + 64 * eps) * rn
            report[g] = dn / rn
        else:
            allowed = 64 * eps * pool[_ftype(g)]
            report[g] = 0.0 if dn == 0 else (dn / pool[_ftype(g)] if pool[_ftype(g)] > 0 else float("inf"))
        if not dn <= allowed:
            bad.append((g, dn, rn, pool[_ftype(g)]))
    assert not bad, "parity gate failed (group, ||diff||, ||ref||, ||pool||): %s" % bad[:6]
    assert all(np.isfinite(v) for v in report.values()), report
    return report
