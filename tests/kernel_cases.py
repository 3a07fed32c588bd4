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
                 capi.K_BETA: "oracle_step_beta"}
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


REPLAY = {"beta": replay_beta, "curl": replay_curl, "edhb": replay_edhb, "lorentz": replay_lorentz, "dft": replay_dft}


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
