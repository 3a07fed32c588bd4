"""ctypes view of include/meep_b200.h (the C ABI of the B200 FDTD engine).

Used by the tests and bench.py to call the CUDA library directly.  Loading fails loudly if the
library is missing; creating a context fails loudly if no CUDA device is visible — there is no
CPU fallback anywhere in this package.
"""
import ctypes as C
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "meep_b200", "lib")

F64, F32 = 0, 1
(K_CURL, K_EDHB, K_LORENTZ, K_FMP, K_SOURCE, K_HALO, K_ZERO, K_DFT, K_FLUX, K_STEP3, K_BETA) = range(11)
K_EXCHANGE = 11
K_CYLINT = 12
K_CYLR0 = 13
K_STEP3_GENERAL = 14
K_BFAST = 15
K_AVERAGE = 16
K_GYRO = 17
K_NOISE = 18
NUM_KINDS = 19
MAX_P = 8


class Box(C.Structure):
    _fields_ = [("idx0", C.c_int64), ("s", C.c_int64 * 3), ("n", C.c_int32 * 3), ("reserved", C.c_int32)]


class Pml(C.Structure):
    _fields_ = [("sig", C.c_void_p), ("kap", C.c_void_p), ("siginv", C.c_void_p),
                ("k0", C.c_int32), ("ks", C.c_int32 * 3)]


class CurlJob(C.Structure):
    _fields_ = [("box", Box), ("f", C.c_void_p), ("g1", C.c_void_p), ("g2", C.c_void_p),
                ("s1", C.c_int64), ("s2", C.c_int64), ("dtdx", C.c_double), ("dt", C.c_double),
                ("pml", Pml), ("pmlu", Pml), ("fu", C.c_void_p), ("cnd", C.c_void_p),
                ("cndinv", C.c_void_p), ("fcnd", C.c_void_p)]


class EdhbJob(C.Structure):
    _fields_ = [("box", Box), ("f", C.c_void_p), ("g", C.c_void_p), ("g1", C.c_void_p), ("g2", C.c_void_p),
                ("u", C.c_void_p), ("u1", C.c_void_p), ("u2", C.c_void_p),
                ("s", C.c_int64), ("s1", C.c_int64), ("s2", C.c_int64),
                ("chi2", C.c_void_p), ("chi3", C.c_void_p), ("fw", C.c_void_p), ("pmlw", Pml)]


class LorentzJob(C.Structure):
    _fields_ = [("box", Box), ("p", C.c_void_p), ("pp", C.c_void_p), ("w", C.c_void_p), ("s", C.c_void_p),
                ("w1", C.c_void_p), ("s1", C.c_void_p), ("w2", C.c_void_p), ("s2", C.c_void_p),
                ("is_", C.c_int64), ("is1", C.c_int64), ("is2", C.c_int64),
                ("gamma1inv", C.c_double), ("gamma1", C.c_double), ("omega0dtsqr", C.c_double),
                ("omega0dtsqr_denom", C.c_double), ("pzero", C.c_void_p), ("szero", C.c_void_p),
                ("ntot", C.c_int64)]


class FmpJob(C.Structure):
    _fields_ = [("fmp", C.c_void_p), ("d", C.c_void_p), ("p", C.c_void_p * MAX_P),
                ("np", C.c_int32), ("reserved", C.c_int32), ("ntot", C.c_int64), ("pzero", C.c_void_p * MAX_P)]


class SrcJob(C.Structure):
    _fields_ = [("f_re", C.c_void_p), ("f_im", C.c_void_p), ("cndinv", C.c_void_p),
                ("index", C.c_void_p), ("amp", C.c_void_p), ("npts", C.c_int64), ("dt", C.c_double),
                ("scalar_slot", C.c_int32), ("mode", C.c_int32)]


class HaloRun(C.Structure):
    _fields_ = [("src0", C.c_uint64), ("dst0", C.c_uint64), ("dsrc", C.c_int64), ("ddst", C.c_int64),
                ("n", C.c_int32), ("negate", C.c_int32)]


class HaloJob(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("phase", C.c_void_p),
                ("n_phase", C.c_int64), ("n_negate", C.c_int64), ("n_copy", C.c_int64),
                ("dst_flag", C.c_void_p), ("runs", C.c_void_p), ("nrun", C.c_int64)]


class ZeroJob(C.Structure):
    _fields_ = [("ptrs", C.c_void_p), ("n", C.c_int64)]


class DftJob(C.Structure):
    _fields_ = [("box", Box), ("f_re", C.c_void_p), ("f_im", C.c_void_p),
                ("avg1", C.c_int64), ("avg2", C.c_int64),
                ("wgt_s0", C.c_double * 3), ("wgt_s1", C.c_double * 3),
                ("wgt_e0", C.c_double * 3), ("wgt_e1", C.c_double * 3),
                ("dV0", C.c_double), ("dV1", C.c_double),
                ("use_weights", C.c_int32), ("sqrt_weights", C.c_int32),
                ("dft", C.c_void_p), ("nomega", C.c_int32), ("phase_slot", C.c_int32)]


class FluxJob(C.Structure):
    _fields_ = [("e", C.c_void_p), ("h", C.c_void_p), ("npts", C.c_int64),
                ("nomega", C.c_int32), ("reserved", C.c_int32), ("out", C.c_void_p)]


class Step3Comp(C.Structure):
    _fields_ = [("lo", C.c_int32 * 3), ("hi", C.c_int32 * 3), ("f", C.c_void_p),
                ("g1", C.c_void_p), ("g2", C.c_void_p), ("s1", C.c_int64), ("s2", C.c_int64),
                ("dtdx", C.c_double), ("pml", Pml), ("pmlu", Pml), ("fu", C.c_void_p),
                ("cnd", C.c_void_p), ("cndinv", C.c_void_p), ("fcnd", C.c_void_p),
                ("e", C.c_void_p), ("u", C.c_void_p), ("fw", C.c_void_p), ("pmlw", Pml),
                ("metal_lo", C.c_int32 * 3), ("metal_hi", C.c_int32 * 3)]


class Step3Job(C.Structure):
    _fields_ = [("n", C.c_int32 * 3), ("reserved", C.c_int32), ("stride", C.c_int64 * 3),
                ("dt", C.c_double), ("ix_lo", C.c_int32), ("ix_hi", C.c_int32),
                ("noepi_lo", C.c_int32), ("noepi_n", C.c_int32), ("c", Step3Comp * 3)]


class BetaJob(C.Structure):
    _fields_ = [("box", Box), ("f", C.c_void_p), ("g", C.c_void_p), ("betadt", C.c_double),
                ("pml", Pml), ("pmlu", Pml), ("fu", C.c_void_p), ("cndinv", C.c_void_p), ("fcnd", C.c_void_p),
                ("cyl", C.c_int32), ("r_is2", C.c_int32)]


class BfastJob(C.Structure):
    _fields_ = [("box", Box), ("f", C.c_void_p), ("g1", C.c_void_p), ("g2", C.c_void_p), ("s1", C.c_int64),
                ("s2", C.c_int64), ("k1", C.c_double), ("k2", C.c_double), ("pml", Pml), ("pmlu", Pml),
                ("fu", C.c_void_p), ("cnd", C.c_void_p), ("cndinv", C.c_void_p), ("fcnd", C.c_void_p),
                ("F", C.c_void_p)]


class GyroJob(C.Structure):
    _fields_ = [("box", Box), ("p", C.c_void_p * 3), ("pp", C.c_void_p * 3), ("w", C.c_void_p * 3),
                ("s", C.c_void_p), ("is_", C.c_int64), ("is1", C.c_int64), ("is2", C.c_int64),
                ("c", C.c_double * 4), ("gt", (C.c_double * 3) * 3), ("inv", (C.c_double * 3) * 3),
                ("model", C.c_int32), ("reserved", C.c_int32)]


class NoiseJob(C.Structure):
    _fields_ = [("box", Box), ("p", C.c_void_p), ("slot", C.c_int64)]


class AverageJob(C.Structure):
    _fields_ = [("f", C.c_void_p), ("backup", C.c_void_p), ("n", C.c_int64)]


class CylIntJob(C.Structure):
    _fields_ = [("out", C.c_void_p), ("fp", C.c_void_p), ("nr", C.c_int64), ("sr", C.c_int64),
                ("ir0", C.c_double)]


class CylR0Job(C.Structure):
    _fields_ = [("box", Box), ("f", C.c_void_p), ("fu", C.c_void_p), ("fp", C.c_void_p), ("fm", C.c_void_p),
                ("sd", C.c_int64), ("c", C.c_double), ("mult", C.c_double), ("dt", C.c_double),
                ("mode", C.c_int32), ("reserved", C.c_int32), ("cnd", C.c_void_p), ("cndinv", C.c_void_p),
                ("fcnd", C.c_void_p), ("pml", Pml), ("pmlu", Pml)]


class Xfer(C.Structure):
    _fields_ = [("peer", C.c_int32), ("reserved", C.c_int32), ("buf", C.c_void_p), ("count", C.c_int64)]


JOB_TYPES = {K_CURL: CurlJob, K_EDHB: EdhbJob, K_LORENTZ: LorentzJob, K_FMP: FmpJob, K_SOURCE: SrcJob,
             K_HALO: HaloJob, K_ZERO: ZeroJob, K_DFT: DftJob, K_FLUX: FluxJob, K_STEP3: Step3Job, K_BETA: BetaJob,
             K_CYLINT: CylIntJob, K_CYLR0: CylR0Job, K_BFAST: BfastJob, K_AVERAGE: AverageJob, K_GYRO: GyroJob, K_NOISE: NoiseJob}


def declare(lib):
    """Attach argtypes/restypes of every entry point declared in include/meep_b200.h."""
    vp, i, i64, d, sz = C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_size_t
    P = C.POINTER
    sig = {
        "mb200_abi_version": (i, []),
        "mb200_last_error": (C.c_char_p, []),
        "mb200_device_count": (i, []),
        "mb200_init": (i, [i, P(vp)]),
        "mb200_destroy": (None, [vp]),
        "mb200_sync": (i, [vp]),
        "mb200_malloc": (i, [vp, sz, P(vp)]),
        "mb200_free": (i, [vp, vp]),
        "mb200_memset": (i, [vp, vp, i, sz]),
        "mb200_h2d": (i, [vp, vp, vp, sz]),
        "mb200_d2h": (i, [vp, vp, vp, sz]),
        "mb200_d2h_async": (i, [vp, vp, vp, sz]),
        "mb200_d2h_box": (i, [vp, vp, vp, sz, i64, i64, i64, i64, i64, i64, i64, i64]),
        "mb200_d2d": (i, [vp, vp, vp, sz]),
        "mb200_host_alloc": (i, [sz, P(vp)]),
        "mb200_host_free": (i, [vp]),
        "mb200_bytes_allocated": (sz, [vp]),
        "mb200_plan_create": (i, [vp, i, i, vp, i, P(vp)]),
        "mb200_plan_run": (i, [vp, vp, vp, sz]),
        "mb200_plan_destroy": (None, [vp, vp]),
        "mb200_plan_bytes": (d, [vp]),
        "mb200_plan_points": (d, [vp]),
        "mb200_step_curl": (i, [vp, i, vp, i]),
        "mb200_step_update_EDHB": (i, [vp, i, vp, i]),
        "mb200_lorentzian_update_P": (i, [vp, i, vp, i]),
        "mb200_subtract_P": (i, [vp, i, vp, i]),
        "mb200_step_source": (i, [vp, i, vp, i, vp, i]),
        "mb200_step_boundaries": (i, [vp, i, vp, i]),
        "mb200_zero_metal": (i, [vp, i, vp, i]),
        "mb200_update_dft": (i, [vp, i, vp, i, vp, i]),
        "mb200_dft_flux": (i, [vp, i, vp, i]),
        "mb200_step3": (i, [vp, i, vp, i]),
        "mb200_step_beta": (i, [vp, i, vp, i]),
        "mb200_step_bfast": (i, [vp, i, vp, i]),
        "mb200_average_with_backup": (i, [vp, i, vp, i]),
        "mb200_gyrotropic_update_P": (i, [vp, i, vp, i]),
        "mb200_add_noise": (i, [vp, i, vp, i, vp, C.c_int64]),
        "mb200_cyl_rderiv_int": (i, [vp, i, vp, i]),
        "mb200_cyl_origin": (i, [vp, i, vp, i]),
        "mb200_comm_unique_id": (i, [vp]),
        "mb200_comm_create": (i, [vp, i, i, vp, P(vp)]),
        "mb200_comm_destroy": (None, [vp]),
        "mb200_comm_exchange": (i, [vp, vp, i, vp, i, vp, i]),
        "mb200_ipc_export": (i, [vp, vp, vp]),
        "mb200_ipc_import": (i, [vp, vp, P(vp)]),
        "mb200_ipc_close": (i, [vp, vp]),
        "mb200_flag_signal": (i, [vp, vp, C.c_uint64]),
        "mb200_flag_wait": (i, [vp, vp, C.c_uint64]),
        "mb200_flag_signal_many": (i, [vp, vp, vp, i]),
        "mb200_flag_wait_many": (i, [vp, vp, vp, i]),
        "mb200_block_zero_flags": (i, [vp, i, vp, i64, vp]),
        "mb200_check_finite": (i, [vp, i, vp, i64, vp]),
        "mb200_timer_start": (i, [vp]),
        "mb200_timer_stop": (i, [vp, P(d)]),
        "mb200_profile_enable": (i, [vp, i]),
        "mb200_profile_reset": (i, [vp]),
        "mb200_profile_get": (i, [vp, i, P(i64), P(d), P(d)]),
        "mb200_mark": (i, [vp, i]),
        "mb200_marks_collect": (i, [vp, P(i), P(d), i, P(i)]),
        "mb200_launch_count": (i64, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    return sig


_lib = None


def load(path=None):
    """Load libmeepb200.so (the CUDA implementation).  Raises if it is not built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.path.join(LIBDIR, "libmeepb200.so")
    if not os.path.exists(p):
        raise RuntimeError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'`" % p)
    lib = C.CDLL(p, mode=C.RTLD_GLOBAL)
    declare(lib)
    if lib.mb200_abi_version() != 2:
        raise RuntimeError("libmeepb200 ABI version mismatch")
    if path is None:
        _lib = lib
    return lib


class Error(RuntimeError):
    pass


class Context:
    """Thin RAII wrapper of mb200_ctx for the tests / bench."""

    def __init__(self, device=0, lib=None):
        self.lib = lib or load()
        self.ctx = C.c_void_p()
        if self.lib.mb200_device_count() < 1:
            raise Error("meep_b200: no CUDA device visible (no CPU fallback exists)")
        self._ck(self.lib.mb200_init(device, C.byref(self.ctx)))
        self._bufs = []

    def _ck(self, rc):
        if rc:
            raise Error(self.lib.mb200_last_error().decode())

    def close(self):
        if self.ctx:
            for b in self._bufs:
                self.lib.mb200_free(self.ctx, b)
            self._bufs = []
            self.lib.mb200_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def malloc(self, nbytes):
        p = C.c_void_p()
        self._ck(self.lib.mb200_malloc(self.ctx, nbytes, C.byref(p)))
        self._bufs.append(p)
        return p.value

    def upload(self, arr):
        """numpy array -> device pointer (int)"""
        import numpy as np
        a = np.ascontiguousarray(arr)
        p = self.malloc(max(a.nbytes, 8))
        if a.nbytes:
            self._ck(self.lib.mb200_h2d(self.ctx, p, a.ctypes.data, a.nbytes))
        return p

    def download(self, ptr, like):
        import numpy as np
        out = np.empty_like(like)
        if out.nbytes:
            self._ck(self.lib.mb200_d2h(self.ctx, out.ctypes.data, ptr, out.nbytes))
        return out

    def sync(self):
        self._ck(self.lib.mb200_sync(self.ctx))

    def run_jobs(self, kind, dtype, jobs, run_data=None):
        """one-shot plan over a list of ctypes job structs"""
        T = JOB_TYPES[kind]
        arr = (T * len(jobs))(*jobs)
        plan = C.c_void_p()
        self._ck(self.lib.mb200_plan_create(self.ctx, kind, dtype, C.cast(arr, C.c_void_p), len(jobs), C.byref(plan)))
        try:
            if run_data is not None:
                import numpy as np
                rd = np.ascontiguousarray(run_data)
                self._ck(self.lib.mb200_plan_run(self.ctx, plan, rd.ctypes.data, rd.nbytes))
            else:
                self._ck(self.lib.mb200_plan_run(self.ctx, plan, None, 0))
            self.sync()
        finally:
            self.lib.mb200_plan_destroy(self.ctx, plan)
