"""Kernel-level parity on the GPU, through the C ABI (include/meep_b200.h):
  * every golden vector recorded from the reference's own functions,
  * seeded random cases at sizes that exercise every tile shape, against the plain-C oracle,
  * the fused step3 kernel against three step_curl + three step_update_EDHB oracle calls."""
import ctypes as C

import numpy as np
import pytest

from meep_b200 import capi
from kernel_cases import DT, KTOL, REAL, DevMem, HostMem, check_golden, mk_pml, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("kind", ["curl", "beta", "edhb", "lorentz", "dft", "cyl", "bfast", "gyro"])
def test_cuda_matches_reference_golden(kind, prec):
    worst = check_golden(kind, prec, DevMem)
    print("CUDA vs reference golden: %s/%s worst rel err %.2e" % (kind, prec, worst))


def _box(n, strides, idx0):
    b = capi.Box()
    b.idx0 = idx0
    for k in range(3):
        b.n[k] = n[k]
        b.s[k] = strides[k]
    return b


def _curl_case(rng, dims, variant, T):
    """a D-type update (low neighbours) on a (nx+1)(ny+1)(nz+1) array, owned box x:0..nx-1, y,z:1..n"""
    nx, ny, nz = dims
    sx, sy, sz = (ny + 1) * (nz + 1), nz + 1, 1
    ntot = (nx + 1) * (ny + 1) * (nz + 1)
    PML, FU, CND, G2 = bool(variant & 8), bool(variant & 4), bool(variant & 2), bool(variant & 1)
    A = lambda lo=-1.0, hi=1.0, n=ntot: rng.uniform(lo, hi, n).astype(T)
    arr = dict(f=A(), g1=A(), g2=A() if G2 else None, fu=A() if FU else None,
               fcnd=A() if (CND and PML) else None, cnd=A(0, 2) if CND else None,
               cndinv=A(0.5, 1) if CND else None,
               sig=A(0, .5, 2 * ny + 2), kap=A(1, 2, 2 * ny + 2), siginv=A(.3, 1, 2 * ny + 2),
               sigu=A(0, .5, 2 * nz + 2), kapu=A(1, 2, 2 * nz + 2), siginvu=A(.3, 1, 2 * nz + 2))
    box = ((nx, ny, nz), (sx, sy, sz), sy + sz)
    pml = (1.0 if PML else 0.0, 2, 0, 2, 0)   # k = 2 + 2*i2  (y direction, owned from index 1)
    pmlu = (1.0 if FU else 0.0, 2, 0, 0, 2)
    return arr, box, pml, pmlu, (-sy, -sz)


def _run_curl(mem, arr, box, pml, pmlu, s):
    P = {k: mem.put(v) for k, v in arr.items()}
    j = capi.CurlJob()
    j.box = _box(*box)
    j.f, j.g1, j.g2 = P["f"], P["g1"], P["g2"]
    j.s1, j.s2 = s
    j.dtdx, j.dt = 0.5, 0.05
    j.pml = mk_pml(pml, P["sig"], P["kap"], P["siginv"])
    j.pmlu = mk_pml(pmlu, P["sigu"], P["kapu"], P["siginvu"])
    j.fu, j.cnd, j.cndinv, j.fcnd = P["fu"], P["cnd"], P["cndinv"], P["fcnd"]
    mem.run(capi.K_CURL, [j])
    out = {k: mem.get(P[k], arr[k]) for k in ("f", "fu", "fcnd") if arr[k] is not None}
    mem.close()
    return out


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("dims", [(19, 7, 70), (9, 33, 30), (17, 40, 10), (3, 5, 130), (40, 3, 3)])
def test_curl_all_variants_all_tile_shapes(dims, prec):
    rng = np.random.default_rng(1234)
    for variant in range(16):
        arr, box, pml, pmlu, s = _curl_case(rng, dims, variant, REAL[prec])
        want = _run_curl(HostMem(prec), arr, box, pml, pmlu, s)
        got = _run_curl(DevMem(prec), arr, box, pml, pmlu, s)
        for k in want:
            assert rel_err(got[k], want[k]) <= KTOL[prec], (dims, variant, k)


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_fused_step3_equals_separate_passes(prec):
    """step3 (3 x step_curl + 3 x diagonal step_update_EDHB in one pass) vs the oracle's six calls"""
    T = REAL[prec]
    rng = np.random.default_rng(7)
    nx, ny, nz = 21, 13, 75
    st = ((ny + 1) * (nz + 1), nz + 1, 1)
    ntot = (nx + 1) * (ny + 1) * (nz + 1)
    A = lambda lo=-1.0, hi=1.0, n=ntot: rng.uniform(lo, hi, n).astype(T)
    H = [A(), A(), A()]
    D = [A(), A(), A()]
    E = [A(), A(), A()]
    U = [A(0.1, 1), None, A(0.1, 1)]
    FU = [A(), None, A()]
    FW = [None, A(), A()]
    nn = (nx, ny, nz)
    sig = [[A(0, .5, 2 * nn[d] + 2), A(1, 2, 2 * nn[d] + 2), A(.3, 1, 2 * nn[d] + 2)] for d in range(3)]
    # component d of a D-type field: Yee shift 1 along d -> owned 0..n-1 there, 1..n elsewhere
    lo = [[0 if k == d else 1 for k in range(3)] for d in range(3)]
    hi = [[nn[k] - 1 if k == d else nn[k] for k in range(3)] for d in range(3)]

    def run(mem, fused):
        pH, pD, pE = [mem.put(a) for a in H], [mem.put(a) for a in D], [mem.put(a) for a in E]
        pU, pFU, pFW = [mem.put(a) for a in U], [mem.put(a) for a in FU], [mem.put(a) for a in FW]
        pS = [[mem.put(a) for a in s3] for s3 in sig]
        comps = []
        for d in range(3):
            d1, d2 = (d + 1) % 3, (d + 2) % 3
            c = dict(lo=lo[d], hi=hi[d], f=pD[d], g1=pH[d2], g2=pH[d1], s1=-st[d1], s2=-st[d2],
                     dsig=d1, dsigu=d2 if FU[d] is not None else None, fu=pFU[d],
                     e=pE[d], u=pU[d], fw=pFW[d], dsigw=d if FW[d] is not None else None)
            comps.append(c)

        def pml_for(direction, base_lo, rebased):
            p = capi.Pml()
            if direction is None:
                return p
            p.sig, p.kap, p.siginv = pS[direction]
            p.k0 = 0 if rebased else 2 * base_lo[direction]
            for k in range(3):
                p.ks[k] = 2 if k == direction else 0
            return p

        if fused:
            J = capi.Step3Job()
            for k in range(3):
                J.n[k] = nn[k]
                J.stride[k] = st[k]
            J.dt = 0.05
            J.ix_lo, J.ix_hi = 0, nn[0]
            for d, c in enumerate(comps):
                Cc = J.c[d]
                for k in range(3):
                    Cc.lo[k], Cc.hi[k] = c["lo"][k], c["hi"][k]
                    Cc.metal_lo[k] = Cc.metal_hi[k] = -1
                Cc.f, Cc.g1, Cc.g2, Cc.s1, Cc.s2, Cc.dtdx = c["f"], c["g1"], c["g2"], c["s1"], c["s2"], 0.5
                Cc.pml = pml_for(c["dsig"], c["lo"], True)
                Cc.pmlu = pml_for(c["dsigu"], c["lo"], True)
                Cc.fu, Cc.e, Cc.u, Cc.fw = c["fu"], c["e"], c["u"], c["fw"]
                Cc.pmlw = pml_for(c["dsigw"], c["lo"], True)
            mem.run(capi.K_STEP3, [J])
        else:
            for d, c in enumerate(comps):
                n = [c["hi"][k] - c["lo"][k] + 1 for k in range(3)]
                idx0 = sum(c["lo"][k] * st[k] for k in range(3))
                j = capi.CurlJob()
                j.box = _box(n, st, idx0)
                j.f, j.g1, j.g2, j.s1, j.s2, j.dtdx, j.dt = c["f"], c["g1"], c["g2"], c["s1"], c["s2"], 0.5, 0.05
                j.pml = pml_for(c["dsig"], c["lo"], False)
                j.pmlu = pml_for(c["dsigu"], c["lo"], False)
                # loop-relative k0: the curl job indexes from the box start
                j.pml.k0 = 2 * c["lo"][c["dsig"]]
                if c["dsigu"] is not None:
                    j.pmlu.k0 = 2 * c["lo"][c["dsigu"]]
                j.fu = c["fu"]
                mem.run(capi.K_CURL, [j])
                e = capi.EdhbJob()
                e.box = _box(n, st, idx0)
                e.f, e.g, e.u, e.fw, e.s = c["e"], c["f"], c["u"], c["fw"], st[d]
                e.pmlw = pml_for(c["dsigw"], c["lo"], False)
                if c["dsigw"] is not None:
                    e.pmlw.k0 = 2 * c["lo"][c["dsigw"]]
                mem.run(capi.K_EDHB, [e])
        out = {}
        for d in range(3):
            out["D%d" % d] = mem.get(pD[d], D[d])
            out["E%d" % d] = mem.get(pE[d], E[d])
            if FU[d] is not None:
                out["FU%d" % d] = mem.get(pFU[d], FU[d])
            if FW[d] is not None:
                out["FW%d" % d] = mem.get(pFW[d], FW[d])
        mem.close()
        return out

    want = run(HostMem(prec), False)
    for fused in (False, True):
        got = run(DevMem(prec), fused)
        for k in want:
            assert rel_err(got[k], want[k]) <= KTOL[prec], (fused, k, rel_err(got[k], want[k]))


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_halo_zero_source_fmp_flux(prec):
    T = REAL[prec]
    R = np.dtype(T).itemsize
    rng = np.random.default_rng(11)
    n = 5000
    src_arr = rng.uniform(-1, 1, n).astype(T)
    dst_arr = rng.uniform(-1, 1, n).astype(T)
    nph, nneg, ncp = 301, 777, 1500
    perm = rng.permutation(n)
    si = perm[:2 * nph + nneg + ncp]
    di = rng.permutation(n)[:2 * nph + nneg + ncp]
    phase = rng.uniform(-1, 1, 2 * nph).astype(T)

    def halo(mem):
        ps, pd = mem.put(src_arr), mem.put(dst_arr)
        j = capi.HaloJob()
        j.src = mem.put((ps + si.astype(np.uint64) * R).astype(np.uint64))
        j.dst = mem.put((pd + di.astype(np.uint64) * R).astype(np.uint64))
        j.phase = mem.put(phase)
        j.n_phase, j.n_negate, j.n_copy = nph, nneg, ncp
        mem.run(capi.K_HALO, [j])
        out = mem.get(pd, dst_arr)
        mem.close()
        return out

    assert rel_err(halo(DevMem(prec)), halo(HostMem(prec))) <= KTOL[prec]

    # the same NEGATE || COPY transfers in run-length form (constant-stride runs, one warp per run)
    def halo_runs(mem):
        ps, pd = mem.put(src_arr), mem.put(dst_arr)
        runs = (capi.HaloRun * 7)()
        spec = [(0, 3, 10, 2, 300, 1), (1000, 1, 1200, 5, 129, 0), (2000, 7, 2500, 1, 33, 0),
                (2400, -2, 3000, 3, 64, 1), (3500, 1, 3600, 1, 1, 0), (3700, 2, 4000, -1, 130, 0),
                (4200, 1, 4300, 1, 500, 1)]
        total = 0
        for k, (s0, ds, d0, dd, cnt, neg) in enumerate(spec):
            runs[k].src0, runs[k].dst0 = ps + s0 * R, pd + d0 * R
            runs[k].dsrc, runs[k].ddst, runs[k].n, runs[k].negate = ds * R, dd * R, cnt, neg
            total += cnt
        j = capi.HaloJob()
        j.runs = mem.put(np.frombuffer(bytes(runs), dtype=np.uint8).copy())
        j.nrun = len(spec)
        j.n_negate = sum(c for (_, _, _, _, c, neg) in spec if neg)
        j.n_copy = total - j.n_negate
        mem.run(capi.K_HALO, [j])
        out = mem.get(pd, dst_arr)
        mem.close()
        return out

    want = dst_arr.copy()
    for (s0, ds, d0, dd, cnt, neg) in [(0, 3, 10, 2, 300, 1), (1000, 1, 1200, 5, 129, 0), (2000, 7, 2500, 1, 33, 0),
                                       (2400, -2, 3000, 3, 64, 1), (3500, 1, 3600, 1, 1, 0),
                                       (3700, 2, 4000, -1, 130, 0), (4200, 1, 4300, 1, 500, 1)]:
        for e in range(cnt):
            want[d0 + e * dd] = -src_arr[s0 + e * ds] if neg else src_arr[s0 + e * ds]
    assert np.array_equal(halo_runs(HostMem(prec)), want)
    assert np.array_equal(halo_runs(DevMem(prec)), want)

    # noise term of noisy_lorentzian_susceptibility: two jobs sharing one run-data buffer
    noise = rng.normal(0, 0.1, 6 * 7 * 9 + 5 * 4 * 3)

    def add_noise(mem):
        pa = mem.put(dst_arr)
        j1, j2 = capi.NoiseJob(), capi.NoiseJob()
        j1.box.idx0, j1.p, j1.slot = 11, pa, 0
        j2.box.idx0, j2.p, j2.slot = 2000, pa, 6 * 7 * 9
        for k, (nn, ss) in enumerate(zip((6, 7, 9), (100, 12, 1))):
            j1.box.n[k], j1.box.s[k] = nn, ss
        for k, (nn, ss) in enumerate(zip((5, 4, 3), (60, 13, 2))):
            j2.box.n[k], j2.box.s[k] = nn, ss
        mem.run(capi.K_NOISE, [j1, j2], noise)
        out = mem.get(pa, dst_arr)
        mem.close()
        return out

    assert np.array_equal(add_noise(DevMem(prec)), add_noise(HostMem(prec)))

    # average_with_backup
    def average(mem):
        pf = mem.put(dst_arr)
        j = capi.AverageJob()
        j.f, j.backup, j.n = pf, mem.put(src_arr), n
        mem.run(capi.K_AVERAGE, [j])
        out = mem.get(pf, dst_arr)
        mem.close()
        return out

    assert np.array_equal(average(DevMem(prec)), average(HostMem(prec)))

    # zero_metal
    mem = DevMem(prec)
    pd = mem.put(dst_arr)
    zi = rng.permutation(n)[:400]
    z = capi.ZeroJob()
    z.ptrs = mem.put((pd + zi.astype(np.uint64) * R).astype(np.uint64))
    z.n = len(zi)
    mem.run(capi.K_ZERO, [z])
    got = mem.get(pd, dst_arr)
    mem.close()
    want = dst_arr.copy()
    want[zi] = 0
    assert np.array_equal(got, want)

    # sources (both modes), with and without cndinv / imaginary part
    npts = 333
    idx = rng.permutation(n)[:npts].astype(np.int64)
    amp = rng.uniform(-1, 1, 2 * npts)
    scal = np.array([0.3, -0.7, 1.1, 0.2])
    fre, fim, cnd = rng.uniform(-1, 1, n).astype(T), rng.uniform(-1, 1, n).astype(T), rng.uniform(.5, 1, n).astype(T)

    def source(mem, mode, use_im, use_cnd):
        pr, pi = mem.put(fre), mem.put(fim)
        j = capi.SrcJob()
        j.f_re, j.f_im = pr, (pi if use_im else None)
        j.cndinv = mem.put(cnd) if use_cnd else None
        j.index, j.amp, j.npts, j.dt, j.scalar_slot, j.mode = mem.put(idx), mem.put(amp), npts, 0.05, 1, mode
        mem.run(capi.K_SOURCE, [j], scal)
        out = (mem.get(pr, fre), mem.get(pi, fim))
        mem.close()
        return out

    for mode in (0, 1):
        for use_im in (False, True):
            for use_cnd in (False, True):
                a, b = source(DevMem(prec), mode, use_im, use_cnd), source(HostMem(prec), mode, use_im, use_cnd)
                assert rel_err(a[0], b[0]) <= KTOL[prec] and rel_err(a[1], b[1]) <= KTOL[prec]

    # f_minus_p = D - sum P
    Ps = [rng.uniform(-1, 1, n).astype(T) for _ in range(6)]

    def fmp(mem):
        j = capi.FmpJob()
        pf = mem.put(dst_arr)
        j.fmp, j.d, j.np, j.ntot = pf, mem.put(src_arr), 6, n
        for k in range(6):
            j.p[k] = mem.put(Ps[k])
        mem.run(capi.K_FMP, [j])
        out = mem.get(pf, dst_arr)
        mem.close()
        return out

    assert rel_err(fmp(DevMem(prec)), fmp(HostMem(prec))) <= KTOL[prec]

    # dft_flux inner sum
    npt, nom = 777, 13
    e, h = rng.uniform(-1, 1, 2 * npt * nom).astype(T), rng.uniform(-1, 1, 2 * npt * nom).astype(T)

    def flux(mem):
        j = capi.FluxJob()
        out0 = np.zeros(nom)
        po = mem.put(out0)
        j.e, j.h, j.npts, j.nomega, j.out = mem.put(e), mem.put(h), npt, nom, po
        mem.run(capi.K_FLUX, [j])
        out = mem.get(po, out0)
        mem.close()
        return out

    assert rel_err(flux(DevMem(prec)), flux(HostMem(prec))) <= (1e-12 if prec == "f64" else 1e-5)

    # the device sum is a fixed two-stage tree (no atomics): bit-identical from run to run, also for
    # several jobs (dft_chunk pairs) accumulating into one spectrum and for more frequencies than a warp
    npt2, nom2 = 5000, 100
    e2 = rng.uniform(-1, 1, 2 * npt2 * nom2).astype(T)
    h2 = rng.uniform(-1, 1, 2 * npt2 * nom2).astype(T)

    def flux2(mem):
        out0 = np.zeros(nom2)
        po = mem.put(out0)
        jobs = []
        for lo, hi in ((0, 1234), (1234, 5000)):
            j = capi.FluxJob()
            j.e, j.h = mem.put(e2[2 * lo * nom2:2 * hi * nom2]), mem.put(h2[2 * lo * nom2:2 * hi * nom2])
            j.npts, j.nomega, j.out = hi - lo, nom2, po
            jobs.append(j)
        mem.run(capi.K_FLUX, jobs)
        out = mem.get(po, out0)
        mem.close()
        return out

    a, b = flux2(DevMem(prec)), flux2(DevMem(prec))
    assert np.array_equal(a, b)
    assert rel_err(a, flux2(HostMem(prec))) <= (1e-12 if prec == "f64" else 1e-5)


def test_size_independent_properties_at_scale():
    """at a size the oracle would take long for: linearity of the curl update and exactness of
    the copy/negate halo on 40M points (properties that hold bit-exactly in IEEE arithmetic)"""
    T = np.float64
    rng = np.random.default_rng(3)
    nx, ny, nz = 255, 255, 300
    st = ((ny + 1) * (nz + 1), nz + 1, 1)
    ntot = (nx + 1) * (ny + 1) * (nz + 1)
    mem = DevMem("f64")
    g1, g2 = rng.uniform(-1, 1, ntot), rng.uniform(-1, 1, ntot)
    f0 = np.zeros(ntot)
    j = capi.CurlJob()
    j.box = _box((nx, ny, nz), st, st[1] + st[2])
    pf = mem.put(f0)
    j.f, j.g1, j.g2, j.s1, j.s2, j.dtdx, j.dt = pf, mem.put(g1), mem.put(g2), -st[1], -st[2], 0.5, 0.05
    mem.run(capi.K_CURL, [j])
    once = mem.get(pf, f0)
    # closed form on the owned box (numpy, vectorised): f = -dtdx*(g1[i-sy]-g1[i] + g2[i]-g2[i-sz])
    G1, G2 = g1.reshape(nx + 1, ny + 1, nz + 1), g2.reshape(nx + 1, ny + 1, nz + 1)
    want = np.zeros_like(G1)
    want[:nx, 1:, 1:] = -0.5 * ((G1[:nx, :-1, 1:] - G1[:nx, 1:, 1:]) + G2[:nx, 1:, 1:] - G2[:nx, 1:, :-1])
    assert rel_err(once, want.ravel()) <= 1e-15
    assert np.array_equal(once.reshape(nx + 1, ny + 1, nz + 1)[nx], np.zeros((ny + 1, nz + 1)))  # not-owned plane
    # applying the same update again doubles it exactly (f starts at 0, update is f -= c)
    mem.run(capi.K_CURL, [j])
    twice = mem.get(pf, f0)
    assert np.array_equal(twice, 2 * once)
    mem.close()
