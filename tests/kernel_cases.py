"""Build C-ABI job structs from golden records / seeded random inputs and run them either through
the plain-C oracle (host pointers) or through the CUDA library (device pointers)."""
import ctypes as C
import glob
import os

import numpy as np

from meep_b200 import capi
from parity_util import ROOT, read_dump

GOLDEN = os.path.join(ROOT, "tests", "golden")
REAL = {"f64": np.float64, "f32": np.float32}
DT = {"f64": capi.F64, "f32": capi.F32}
# golden/oracle/CUDA differ only by FMA contraction and evaluation order
KTOL = {"f64": 2e-14, "f32": 2e-6}


class HostMem:
    """executor for oracle/_build/liboracle.so: pointers are numpy buffers"""

    def __init__(self, prec):
        self.prec = prec
        self.keep = []
        self.lib = C.CDLL(os.path.join(ROOT, "oracle", "_build", "liboracle.so"))

    def put(self, a):
        if a is None:
            return None
        b = np.array(a, copy=True)
        self.keep.append(b)
        return b.ctypes.data

    def get(self, ptr, like):
        for b in self.keep:
            if b.ctypes.data == ptr:
                return b.copy()
        raise KeyError(ptr)

    def run(self, kind, jobs, run_data=None):
        names = {capi.K_CURL: "oracle_step_curl", capi.K_EDHB: "oracle_step_update_EDHB",
                 capi.K_LORENTZ: "oracle_lorentzian_update_P", capi.K_FMP: "oracle_subtract_P",
                 capi.K_SOURCE: "oracle_step_source", capi.K_HALO: "oracle_step_boundaries",
                 capi.K_DFT: "oracle_update_dft", capi.K_FLUX: "oracle_dft_flux",
                 capi.K_BETA: "oracle_step_beta", capi.K_CYLINT: "oracle_cyl_rderiv_int",
                 capi.K_CYLR0: "oracle_cyl_origin", capi.K_ZERO: "oracle_zero_metal",
                 capi.K_BFAST: "oracle_step_bfast", capi.K_AVERAGE: "oracle_average_with_backup",
                 capi.K_GYRO: "oracle_gyrotropic_update_P", capi.K_NOISE: "oracle_add_noise"}
        fn = getattr(self.lib, names[kind] + "_" + self.prec)
        fn.restype = None
        for j in jobs:
            if run_data is not None:
                rd = np.ascontiguousarray(run_data)
                fn(C.byref(j), C.c_void_p(rd.ctypes.data))
            else:
                fn(C.byref(j))

    def close(self):
        self.keep = []


class DevMem:
    """executor for meep_b200/lib/libmeepb200.so on cuda:0"""

    def __init__(self, prec):
        self.prec = prec
        # MEEP_B200_TEST_CAPI lets the test CODE be dry-run against the emulator on a GPU-less box
        alt = os.environ.get("MEEP_B200_TEST_CAPI")
        self.ctx = capi.Context(0, lib=capi.load(alt) if alt else None)

    def put(self, a):
        if a is None:
            return None
        return self.ctx.upload(a)

    def get(self, ptr, like):
        return self.ctx.download(ptr, like)

    def run(self, kind, jobs, run_data=None):
        self.ctx.run_jobs(kind, DT[self.prec], jobs, run_data)

    def close(self):
        self.ctx.close()


def golden_files(prefix, prec):
    return sorted(glob.glob(os.path.join(GOLDEN, "%s*_%s.bin" % (prefix, prec))))


def mk_box(v):
    b = capi.Box()
    b.idx0 = int(v[0])
    for k in range(3):
        b.s[k] = int(v[1 + k])
        b.n[k] = int(v[4 + k])
    return b


def mk_pml(v, sig, kap, siginv):
    p = capi.Pml()
    if v is not None and v[0]:
        p.sig, p.kap, p.siginv = sig, kap, siginv
        p.k0 = int(v[1])
        for k in range(3):
            p.ks[k] = int(v[2 + k])
    return p


def rel_err(a, b):
    a = a.astype(np.float64)
    b = b.astype(np.float64)
    n = np.linalg.norm(b)
    return np.linalg.norm(a - b) / n if n > 0 else np.linalg.norm(a - b)


# ---- golden replays ------------------------------------------------------------------------------
def replay_curl(rec, mem):
    g = lambda k: rec.get("in." + k)
    P = {k: mem.put(g(k)) for k in ("f", "g1", "g2", "fu", "fcnd", "cnd", "cndinv", "sig", "kap", "siginv",
                                    "sigu", "kapu", "siginvu")}
    j = capi.CurlJob()
    j.box = mk_box(rec["box"])
    j.f, j.g1, j.g2 = P["f"], P["g1"], P["g2"]
    s1, s2, dtdx, dt = rec["scalars"]
    j.s1, j.s2, j.dtdx, j.dt = int(s1), int(s2), dtdx, dt
    j.pml = mk_pml(rec["pml"], P["sig"], P["kap"], P["siginv"])
    j.pmlu = mk_pml(rec["pmlu"], P["sigu"], P["kapu"], P["siginvu"])
    j.fu, j.cnd, j.cndinv, j.fcnd = P["fu"], P["cnd"], P["cndinv"], P["fcnd"]
    mem.run(capi.K_CURL, [j])
    return {k: mem.get(P[k], g(k)) for k in ("f", "fu", "fcnd") if g(k) is not None}


def replay_beta(rec, mem):
    g = lambda k: rec.get("in." + k)
    P = {k: mem.put(g(k)) for k in ("f", "g", "fu", "fcnd", "cndinv", "siginv", "siginvu")}
    j = capi.BetaJob()
    j.box = mk_box(rec["box"])
    j.f, j.g, j.betadt = P["f"], P["g"], rec["scalars"][0]
    j.pml = mk_pml(rec["pml"], None, None, P["siginv"])
    j.pmlu = mk_pml(rec["pmlu"], None, None, P["siginvu"])
    j.fu, j.cndinv, j.fcnd = P["fu"], P["cndinv"], P["fcnd"]
    mem.run(capi.K_BETA, [j])
    return {k: mem.get(P[k], g(k)) for k in ("f", "fu", "fcnd") if g(k) is not None}


def replay_bfast(rec, mem):
    g = lambda k: rec.get("in." + k)
    P = {k: mem.put(g(k)) for k in ("f", "g1", "g2", "F", "fu", "fcnd", "cnd", "cndinv", "siginv", "siginvu")}
    j = capi.BfastJob()
    j.box = mk_box(rec["box"])
    j.f, j.g1, j.g2, j.F = P["f"], P["g1"], P["g2"], P["F"]
    s1, s2, k1, k2 = rec["scalars"]
    j.s1, j.s2, j.k1, j.k2 = int(s1), int(s2), k1, k2
    j.pml = mk_pml(rec["pml"], None, None, P["siginv"])
    j.pmlu = mk_pml(rec["pmlu"], None, None, P["siginvu"])
    j.fu, j.cnd, j.cndinv, j.fcnd = P["fu"], P["cnd"], P["cndinv"], P["fcnd"]
    mem.run(capi.K_BFAST, [j])
    return {k: mem.get(P[k], g(k)) for k in ("f", "fu", "fcnd", "F") if g(k) is not None}


def replay_gyro(rec, mem):
    g = lambda k: rec.get("in." + k)
    names = ["p0", "p1", "p2", "pp0", "pp1", "pp2", "w0", "w1", "w2", "s"]
    P = {k: mem.put(g(k)) for k in names}
    j = capi.GyroJob()
    j.box = mk_box(rec["box"])
    for k in range(3):
        j.p[k], j.pp[k], j.w[k] = P["p%d" % k], P["pp%d" % k], P["w%d" % k]
    j.s = P["s"]
    sc = rec["scalars"]
    j.is_, j.is1, j.is2 = int(sc[0]), int(sc[1]), int(sc[2])
    for k in range(4):
        j.c[k] = sc[3 + k]
    j.model = int(sc[7])
    for a in range(3):
        for b in range(3):
            j.gt[a][b] = rec["gt"][3 * a + b]
            j.inv[a][b] = rec["inv"][3 * a + b]
    mem.run(capi.K_GYRO, [j])
    return {k: mem.get(P[k], g(k)) for k in names[:6]}


def replay_edhb(rec, mem):
    g = lambda k: rec.get("in." + k)
    P = {k: mem.put(g(k)) for k in ("f", "g", "g1", "g2", "u", "u1", "u2", "chi2", "chi3", "fw", "sigw", "kapw")}
    j = capi.EdhbJob()
    j.box = mk_box(rec["box"])
    j.f, j.g, j.g1, j.g2 = P["f"], P["g"], P["g1"], P["g2"]
    j.u, j.u1, j.u2 = P["u"], P["u1"], P["u2"]
    s, s1, s2 = rec["scalars"]
    j.s, j.s1, j.s2 = int(s), int(s1), int(s2)
    j.chi2, j.chi3, j.fw = P["chi2"], P["chi3"], P["fw"]
    j.pmlw = mk_pml(rec["pmlw"], P["sigw"], P["kapw"], None)
    # the C ABI expects the g1/g2 swap of step_generic.cpp:573 applied by the caller
    if (not j.g1 and j.g2) or (j.g1 and j.g2 and not j.u1 and j.u2):
        j.g1, j.g2 = j.g2, j.g1
        j.u1, j.u2 = j.u2, j.u1
        j.s1, j.s2 = j.s2, j.s1
    mem.run(capi.K_EDHB, [j])
    return {k: mem.get(P[k], g(k)) for k in ("f", "fw") if g(k) is not None}


def replay_lorentz(rec, mem):
    g = lambda k: rec.get("in." + k)
    P = {k: mem.put(g(k)) for k in ("p", "pp", "w", "s", "w1", "s1", "w2", "s2")}
    j = capi.LorentzJob()
    j.box = mk_box(rec["box"])
    j.p, j.pp, j.w, j.s = P["p"], P["pp"], P["w"], P["s"]
    sc = rec["scalars"]
    j.is_, j.is1, j.is2 = int(sc[0]), int(sc[1]), int(sc[2])
    if g("s1") is not None and g("w1") is not None:
        j.w1, j.s1 = P["w1"], P["s1"]
        if g("s2") is not None and g("w2") is not None:
            j.w2, j.s2 = P["w2"], P["s2"]
    j.gamma1inv, j.gamma1, j.omega0dtsqr, j.omega0dtsqr_denom = sc[3], sc[4], sc[5], sc[6]
    mem.run(capi.K_LORENTZ, [j])
    return {k: mem.get(P[k], g(k)) for k in ("p", "pp")}


def replay_dft(rec, mem):
    n = int(rec["nchunks"][0])
    jobs, ptrs, phases, slot = [], [], [], 0
    for k in range(n):
        p = "k%d." % k
        j = capi.DftJob()
        j.box = mk_box(rec[p + "box"])
        j.f_re = mem.put(rec[p + "in.f_re"])
        j.f_im = mem.put(rec.get(p + "in.f_im"))
        sc = rec[p + "scalars"]
        j.avg1, j.avg2, j.dV0, j.dV1 = int(sc[0]), int(sc[1]), sc[2], sc[3]
        j.use_weights, j.sqrt_weights, j.nomega = int(sc[4]), int(sc[5]), int(sc[6])
        w = rec[p + "weights"]
        for d in range(3):
            j.wgt_s0[d], j.wgt_s1[d], j.wgt_e0[d], j.wgt_e1[d] = w[4 * d], w[4 * d + 1], w[4 * d + 2], w[4 * d + 3]
        dptr = mem.put(rec[p + "in.dft"])
        j.dft = dptr
        j.phase_slot = slot
        slot += j.nomega
        phases.append(rec[p + "phase"])
        jobs.append(j)
        ptrs.append(dptr)
    mem.run(capi.K_DFT, jobs, np.concatenate(phases))
    return {"k%d.dft" % k: mem.get(ptrs[k], rec["k%d.in.dft" % k]) for k in range(n)}


# component numbers of src/meep/vec.hpp:31-52
(Er, Ep, Ez, Hr, Hp, Hz, Dr, Dp, Dz, Br, Bp, Bz) = (2, 3, 4, 7, 8, 9, 12, 13, 14, 17, 18, 19)


def replay_cyl(rec, mem):
    """one whole cylindrical fields::step_db(ft) (src/step_db.cpp:40-462, no PML / conductivity) out of
    C-ABI jobs: helper-array scan, curl jobs, i*m/r jobs, the r = 0 row, zeroed rows"""
    nr, nz, m, courant, dt, is_d = rec["dims"]
    nr, nz, is_d = int(nr), int(nz), bool(is_d)
    real = REAL[mem.prec]
    like, P = {}, {}
    for k, v in rec.items():
        if k.startswith("in.f."):
            c, cmp = (int(x) for x in k.split(".")[2:])
            like[(c, cmp)] = v
            P[(c, cmp)] = mem.put(v)
    n = len(next(iter(like.values())))
    # 1. helper arrays + curl jobs
    ints, curls = [], []
    k = 0
    while "job.curl.%d" % k in rec:
        v = rec["job.curl.%d" % k]
        k += 1
        cc, cmp, gp, gm, sp, sm, rderiv = (int(x) for x in v[:7])
        g1 = P[(gp, cmp)] if gp >= 0 else None
        g2 = P[(gm, cmp)] if gm >= 0 else None
        if rderiv:
            ij = capi.CylIntJob()
            ij.out, ij.fp, ij.nr, ij.sr, ij.ir0 = mem.put(np.zeros(n, real)), g1, nr, nz + 1, v[7]
            ints.append(ij)
            g1 = ij.out
        j = capi.CurlJob()
        j.box = mk_box(v[8:15])
        j.f, j.g1, j.g2, j.s1, j.s2 = P[(cc, cmp)], g1, g2, sp, sm
        j.dtdx, j.dt = float(real(courant)), float(real(dt))
        if not j.g1:
            j.g1, j.g2, j.s1, j.s2, j.dtdx = j.g2, j.g1, j.s2, j.s1, -j.dtdx
        if j.g1:
            curls.append(j)
    if ints:
        mem.run(capi.K_CYLINT, ints)
    mem.run(capi.K_CURL, curls)
    # 2. i*m/r terms
    mrs = []
    k = 0
    while "job.mr.%d" % k in rec:
        v = rec["job.mr.%d" % k]
        k += 1
        cc, cmp, cg = (int(x) for x in v[:3])
        j = capi.BetaJob()
        j.box = mk_box(v[5:12])
        j.f, j.g, j.betadt, j.cyl, j.r_is2 = P[(cc, cmp)], P[(cg, 1 - cmp)], v[3], 1, int(v[4])
        mrs.append(j)
    if mrs:
        mem.run(capi.K_BETA, mrs)
    # 3. r = 0 boundary conditions (src/step_db.cpp:282-462)
    r0, zero_rows = [], []  # zero_rows: (component, cmp, row)
    R = np.dtype(real).itemsize
    for cmp in (0, 1):
        def origin(cc, mode, fp, fm, fm_off, sd, c, mult):
            j = capi.CylR0Job()
            j.box = mk_box(rec["box.r0.%d" % cc])
            j.f, j.fp, j.sd, j.c, j.mult, j.dt, j.mode = P[(cc, cmp)], fp, sd, c, mult, dt, mode
            j.fm = (fm + fm_off * R) if fm else None
            r0.append(j)
        if m == 0 and is_d:
            origin(Dz, 0, P[(Hp, cmp)], None, 0, 0, courant * 4, 0)
            zero_rows.append((Dp, cmp, 0))
        elif m == 0:
            zero_rows.append((Br, cmp, 0))
        elif abs(m) == 1:
            if is_d:
                origin(Dp, 1, P[(Hr, cmp)], P[(Hz, cmp)], 0, +1, courant, 2)
                zero_rows.append((Dz, cmp, 0))
            else:
                origin(Br, 1, P[(Ep, cmp)], P[(Ez, 1 - cmp)], nz + 1, -1, -courant, (1 - 2 * cmp) * m)
        else:
            rows = [r for r in range(nr + 1) if r < abs(m)]  # zero_fields_near_cylorigin, origin_r = 0
            for r in rows:
                for c in ((Dr, Dp, Dz) if is_d else (Br, Bp, Bz)):
                    zero_rows.append((c, cmp, r))
    if r0:
        mem.run(capi.K_CYLR0, r0)
    zj = []
    for c, cmp, row in zero_rows:
        ptrs = np.array([P[(c, cmp)] + (row * (nz + 1) + k) * R for k in range(nz + 1)], np.uint64)
        j = capi.ZeroJob()
        j.ptrs, j.n = mem.put(ptrs), len(ptrs)
        zj.append(j)
    if zj:
        mem.run(capi.K_ZERO, zj)
    out = {}
    for key in rec:
        if key.startswith("out.f."):
            c, cmp = (int(x) for x in key.split(".")[2:])
            out["f.%d.%d" % (c, cmp)] = mem.get(P[(c, cmp)], like[(c, cmp)])
    return out


REPLAY = {"gyro": replay_gyro, "cyl": replay_cyl, "bfast": replay_bfast, "beta": replay_beta, "curl": replay_curl, "edhb": replay_edhb, "lorentz": replay_lorentz, "dft": replay_dft}


def check_golden(prefix, prec, mem_factory):
    files = golden_files(prefix + "_", prec)
    assert files, "no golden fixtures for %s/%s" % (prefix, prec)
    worst = 0.0
    for path in files:
        rec = read_dump(path)
        mem = mem_factory(prec)
        try:
            out = REPLAY[prefix](rec, mem)
        finally:
            mem.close()
        for k, v in out.items():
            want = rec[k if prefix == "dft" and False else ("out." + k if prefix != "dft" else k.replace(".dft", ".out.dft"))]
            e = rel_err(v, want)
            worst = max(worst, e)
            assert e <= KTOL[prec], "%s: %s rel err %.3e" % (os.path.basename(path), k, e)
    return worst
