#!/usr/bin/env python
"""bench.py — Yee cell-updates/s of meep::fields::step() on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # device arm (one process per GPU)
  python bench.py --impl reference --steps K --warmup W    # the reference's CPU path, host cores

Workload: BASELINE.json's metric is quoted "at 1/2/4/8 B200" on the 1024^3 dielectric+PML cell
(configs[4], the north-star case), which fits ONE B200 (102 GB of HBM, 25 GB of host RAM), so every
N — including N=1 — runs that cell: the lines of a scaling run form one strong-scaling series.
BASELINE.json configs[1] (the same geometry at 512^3, the reference's 1-GPU case) is measured in the
same N=1 run and reported under "configs1_512"; `--size 512` makes it the headline instead.  If the
box cannot hold 1024^3 the N=1 line falls back to 512^3 (and says so in config.workload).
3-D dielectric box, PML on all faces, non-dispersive, Gaussian dipole source, double precision, built
through the reference's public C++ API (meep::structure / meep::fields) by bench/bench_driver.cpp and
stepped with fields::step().  A "step" is one FDTD time step (one pass of the hot path over the grid).

  value   : cells * K / device time (CUDA events on the engine's stream), fields resident in HBM
  e2e     : the same K steps through the public API as a user's monitoring loop runs them:
            every step uploads that step's source amplitudes (H2D) and reads one field probe back
            (D2H, fields::get_field), timed on the host around the API calls
  roofline: dominant kernel (fused D/B update) algorithmic bytes / CUDA-event time / measured HBM peak
  cpu_baseline: the unmodified reference (oracle/_ref) on this box's host cores, bounded sample

N>1 (torchrun, one process per GPU): BASELINE.json configs[4], the north-star case — the 1024^3
dielectric+PML cell, STRONG scaling: the cell is sharded with the reference's own split_by_cost
(8 ranks: 2x2x2 leaves of 512^3, three cross-GPU faces each), every rank owns one leaf of the binary
partition plus its PML sub-chunks, and chunk boundaries that cross ranks are exchanged
device-to-device through peer memory once per sub-phase.  --scaling weak stacks n^3 cells per GPU
along x instead; --size overrides the edge.  rank/size and the few host-side reductions come from
the MPI-free runtime in meep_b200/host/mympi_b200.cpp (MPI is not installed).
Every line carries `probe`: field values at fixed points of the cell after the same number of steps,
read through fields::get_field — identical numbers for every N on the same cell (sharding changes
nothing in the arithmetic), so the lines of a scaling run check each other's VALUES, not only speed.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
LIB = os.path.join(ROOT, "meep_b200", "lib")


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []
        self.stop_flag = False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], stdout=subprocess.PIPE,
                                     stderr=subprocess.DEVNULL, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(s) > 3 + i and s[3 + i].lower().startswith("active")
                                                         for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": float(self.samples[0][1]) if self.samples[0][1].replace(".", "").isdigit() else None,
                "reasons": reasons, "samples": len(self.samples)}


def run_reference(args):
    """the reference's own CPU implementation (oracle/_ref) on the host cores, bounded sample"""
    exe = os.path.join(LIB, "bench_ref_%s" % args.prec)
    if not os.path.exists(exe):
        raise RuntimeError("%s missing: run __graft_entry__.build() where the reference sources exist" % exe)
    cores = os.cpu_count() or 1
    env = dict(os.environ, OMP_NUM_THREADS=str(cores))
    out = subprocess.run([exe, getattr(args, "workload", "c2"), str(args.cpu_n), str(args.warmup),
                          str(args.steps)], env=env,
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=3000)
    if out.returncode != 0:
        raise RuntimeError("reference arm failed: " + out.stderr[-2000:])
    r = json.loads(out.stdout.strip().splitlines()[-1])
    r["cores"] = cores
    return r


def dist_max(x, world, device):
    """max over ranks of a host scalar (device timings are reduced with this; bench contract)"""
    if world <= 1:
        return x
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def workload_text(workload, nx, ny, nz, world):
    """config.workload: the same text for the device arm and the reference arm"""
    return {
        "c2": "c2: 3D dielectric box %dx%dx%d, PML(1.0) on all faces, non-dispersive, Gaussian Ez dipole, "
              "res 10, Courant 0.5, real fields (BASELINE.json %s)"
              % (nx, ny, nz, "configs[1]" if (nx, ny, nz) == (512, 512, 512) and world == 1 else
                 ("configs[4]: 1024^3 strong-scaling series" if (nx, ny, nz) == (1024, 1024, 1024) else
                  "configs[1] geometry; configs[4] scaling form")),
        "c3": "c3: Drude + 5 Lorentz Au sphere %dx%dx%d, PML(1.0), 100-frequency DFT flux box "
              "(BASELINE.json configs[2], scaled to fit one GPU in double)" % (nx, ny, nz),
        "c4": "c4: anisotropic subpixel-smoothed Si ring %dx%dx%d, off-diagonal chi1inv, PML(1.0) "
              "(BASELINE.json configs[3], scaled)" % (nx, ny, nz)}[workload]


def scaling_label(args, n_dev):
    """strong: one fixed cell over all GPUs (a 1-GPU line is part of the strong series when it runs the
    series' 1024^3 cell); weak otherwise"""
    if n_dev > 1 and not args.replicas:
        return args.scaling
    return "strong" if (args.workload == "c2" and args.n == 1024 and args.scaling == "strong") else "weak"


def csrc_stamp():
    """identifies the kernel sources a committed ncu capture was taken on"""
    import hashlib
    h = hashlib.sha1()
    d = os.path.join(ROOT, "meep_b200", "csrc")
    for f in sorted(os.listdir(d)):
        with open(os.path.join(d, f), "rb") as fh:
            h.update(fh.read())
    with open(os.path.join(ROOT, "include", "meep_b200.h"), "rb") as fh:
        h.update(fh.read())
    return h.hexdigest()[:12]


def cpu_baseline_obj(r):
    return {"value": r["cells_per_s"], "unit": "cell-updates/s", "cores": r["cores"], "kind": "reference",
            "sample": "same workload at %d^3 (%d chunks), %d timed steps after %d warm-up, OpenMP on all host "
                      "cores (MPI is not installed: the reference's OpenMP path stands in for its MPI path)"
                      % (r["n"], r["num_chunks"], r["steps"], r["warmup"])}


def main():
    # stdout carries exactly ONE JSON line: anything libraries print there (e.g. NCCL's version
    # banner) is diverted to stderr for the duration of the run
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        os.write(real_stdout, (json.dumps(line) + "\n").encode())

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", "--n", dest="n", type=int, default=0,
                    help="cells per edge of the workload (use --size under torchrun: --n is ambiguous there); "
                         "default: 512 on one GPU (BASELINE.json configs[1]), 1024 on several (configs[4])")
    ap.add_argument("--cpu-n", type=int, default=192, help="edge of the bounded CPU sample")
    ap.add_argument("--cpu-steps", type=int, default=10)
    ap.add_argument("--prec", default="f64", choices=["f64", "f32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4"],
                    help="c2: dielectric+PML (headline); c3: Au Drude-Lorentz sphere + 100-frequency flux "
                         "box; c4: anisotropic Si ring (n x n x n/4)")
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="N>1: strong = one --size^3 cell over all GPUs (default, the north-star case); "
                         "weak = --size^3 per GPU stacked along x")
    ap.add_argument("--replicas", action="store_true",
                    help="N>1: independent replicas instead of one sharded problem")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    args.auto_n = args.n <= 0
    if args.n <= 0:
        if args.workload != "c2":
            args.n = {"c3": 320, "c4": 512}[args.workload]
        elif max(args.gpus, int(os.environ.get("WORLD_SIZE", "1"))) > 1 and not args.replicas:
            args.n = 1024 if args.scaling == "strong" else 512
        else:
            # one GPU: the same 1024^3 cell as the multi-GPU lines (the north-star case; it fits one B200:
            # 102 GB of HBM), so that the 1/2/4/8-GPU lines form ONE strong-scaling series; configs[1]
            # (512^3) is measured in the same run and reported under "configs1_512"
            args.n = int(os.environ.get("MEEP_B200_BENCH_N1", "1024"))

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        r = run_reference(args)
        # the device arm's config at this GPU count; each step is a bounded sample of it
        n_dev = max(args.gpus, 1)
        nx = args.n * n_dev if (n_dev > 1 and args.scaling == "weak" and not args.replicas) else args.n
        nz = args.n // 4 if args.workload == "c4" else args.n
        line = {"impl": "reference", "metric": "Yee cell-updates/s", "value": r["cells_per_s"],
                "unit": "cell-updates/s", "n_gpus": args.gpus, "steps": r["steps"], "warmup": r["warmup"],
                "ms_per_step": 1e3 * r["seconds"] / r["steps"], "higher_is_better": True,
                "scaling": scaling_label(args, n_dev),
                "vs_baseline": None, "dtype": args.prec, "data": "synthetic",
                "config": {"workload": workload_text(args.workload, nx, args.n, nz, n_dev), "n": args.n,
                           "cell": [nx, args.n, nz], "sample_n": r["n"], "sample_num_chunks": r["num_chunks"],
                           "parallelism": "host CPU, OpenMP on %d cores" % r["cores"]},
                "cpu_baseline": cpu_baseline_obj(r),
                "e2e": {"value": r["cells_per_s"], "unit": "cell-updates/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}}
        emit(line)
        return 0

    # ---- device arm ----------------------------------------------------------------------------
    import torch  # plumbing only: process group for the barrier / max-over-ranks
    import torch.distributed as dist
    from meep_b200 import capi

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    os.environ["MEEP_B200_DEVICE"] = str(local_rank)
    # host-side set-up (material fill, connection tables) uses the host cores this rank may claim;
    # the reference's initialize() falls back to ONE OpenMP thread when the variable is unset
    # (torchrun exports OMP_NUM_THREADS=1 to every rank unless the user set it: MEEP_B200_HOST_THREADS
    # is the explicit override here)
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    os.environ["OMP_NUM_THREADS"] = os.environ.get(
        "MEEP_B200_HOST_THREADS", str(max(1, (os.cpu_count() or 1) // max(local_world, 1))))
    if args.replicas:
        os.environ["MEEP_B200_WORLD_SIZE"] = "1"  # the C++ runtime sees a single process
        os.environ["MEEP_B200_RANK"] = "0"
    sharded = world > 1 and not args.replicas

    lib = capi.load()  # libmeepb200.so (CUDA kernels + C ABI); raises if missing
    if lib.mb200_device_count() < 1:
        raise RuntimeError("bench.py: no CUDA device visible; the device arm has no CPU fallback")
    C.CDLL(os.path.join(LIB, "libmeep_b200_%s.so" % args.prec), mode=C.RTLD_GLOBAL)
    drv = C.CDLL(os.path.join(LIB, "libmeep_b200_bench_%s.so" % args.prec), mode=C.RTLD_GLOBAL)
    drv.mb200_bench_create3d.restype = C.c_void_p
    drv.mb200_bench_create3d.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int]
    for fn in ("mb200_bench_cells", "mb200_bench_probe", "mb200_bench_field_bytes",
               "mb200_bench_algorithmic_bytes_per_step"):
        getattr(drv, fn).restype = C.c_double
        getattr(drv, fn).argtypes = [C.c_void_p]
    drv.mb200_bench_step.argtypes = [C.c_void_p, C.c_int]
    drv.mb200_bench_fields.restype = C.c_void_p
    drv.mb200_bench_fields.argtypes = [C.c_void_p]
    drv.mb200_bench_num_chunks.argtypes = [C.c_void_p]
    drv.mb200_bench_destroy.argtypes = [C.c_void_p]
    host = C.CDLL(os.path.join(LIB, "libmeep_b200_%s.so" % args.prec), mode=C.RTLD_GLOBAL)
    host.meep_b200_ctx.restype = C.c_void_p
    host.meep_b200_ctx.argtypes = [C.c_void_p]
    host.meep_b200_get_stats.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    host.meep_b200_sync_host.argtypes = [C.c_void_p]

    live = []  # simulations not yet destroyed (a failed measurement must not keep 100 GB of HBM)

    def measure(n, light=False):
        """one complete measurement of workload size n (light: skip the cold-start context run)"""
        t0 = time.time()
        nx = n * world if (sharded and args.scaling == "weak") else n
        nz = n // 4 if args.workload == "c4" else n
        h = drv.mb200_bench_create3d(args.workload.encode(), nx, n, nz, 0)
        if not h:
            raise RuntimeError("bench driver failed to build the workload")
        live.append(h)
        t_setup = time.time() - t0
        cells = drv.mb200_bench_cells(h)  # the whole (possibly sharded) cell
        fptr = drv.mb200_bench_fields(h)
        ranks_per_problem = world if sharded else 1

        def stats():
            a = (C.c_double * 8)()
            host.meep_b200_get_stats(fptr, a)
            return list(a)

        def barrier():
            if world > 1:
                dist.barrier()

        def max_over_ranks(x):
            return dist_max(x, world, "cuda")

        # warm-up: first step uploads all arrays, allocates PML aux fields, builds connection
        # tables (reference host code) and all launch plans
        t0 = time.time()
        if drv.mb200_bench_step(h, args.warmup):
            raise RuntimeError("warm-up failed")
        ctx = host.meep_b200_ctx(fptr)
        lib.mb200_sync(ctx)
        t_warm = time.time() - t0

        # ---- timed region 1: device-resident throughput (CUDA events on the engine's stream) --------
        sampler = ClockSampler(local_rank)
        sampler.start()
        s0 = stats()
        barrier()
        lib.mb200_sync(ctx)
        lib.mb200_timer_start(ctx)
        if drv.mb200_bench_step(h, args.steps):
            raise RuntimeError("timed steps failed")
        ms = C.c_double()
        lib.mb200_timer_stop(ctx, C.byref(ms))  # synchronises
        barrier()
        s1 = stats()
        sampler.stop_flag = True
        sampler.join(timeout=2)
        dev_ms = max_over_ranks(ms.value)
        launches = int(s1[3] - s0[3])

        # ---- timed region 2: end to end through the public API, host-visible result every step -------
        barrier()
        lib.mb200_sync(ctx)
        e0 = stats()
        t0 = time.perf_counter()
        acc = 0.0
        for _ in range(args.steps):
            drv.mb200_bench_step(h, 1)
            acc += drv.mb200_bench_probe(h)  # fields::get_field -> D2H of the probed values
        lib.mb200_sync(ctx)
        e2e_s = time.perf_counter() - t0
        barrier()
        e1 = stats()
        e2e_s = max_over_ranks(e2e_s)
        if acc != acc:
            raise RuntimeError("NaN in the probed field")
        # copies are made by the rank that owns the source / the probed point: report the busiest rank
        e2e_h2d = max_over_ranks((e1[1] - e0[1]) / args.steps)
        e2e_d2h = max_over_ranks((e1[2] - e0[2]) / args.steps)
        # value check across GPU counts: the same points of the same cell after the same number of steps
        nprobe = 8
        pv = (C.c_double * nprobe)()
        drv.mb200_bench_probes.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int]
        drv.mb200_bench_probes(h, pv, nprobe)
        drv.mb200_bench_time_step.argtypes = [C.c_void_p]
        drv.mb200_bench_probe_component.argtypes = [C.c_void_p]
        drv.mb200_bench_probe_component.restype = C.c_char_p
        anchor = {"c2": "cell centre", "c3": "source point", "c4": "source point"}[args.workload]
        probe = {"after_steps": int(drv.mb200_bench_time_step(h)),
                 "component": drv.mb200_bench_probe_component(h).decode(),
                 "points": anchor + " + (0.35,0.25,0.15) + s_k*(1,-1,1), s_k = +-0.1 k, k=0..7, via fields::get_field",
                 "values": [float(x) for x in pv]}

        # ---- timed region 3 (context for e2e): a COLD run of K steps — every field array starts on the
        # host (upload inside the timed region) and ends on the host (download inside it)
        cold = None
        if world == 1 and not light and drv.mb200_bench_field_bytes(h) < 20e9:
            host.meep_b200_mark_host_dirty.argtypes = [C.c_void_p]
            host.meep_b200_mark_host_dirty(fptr)  # arrays are downloaded; the next step re-uploads them
            c0 = stats()
            t0 = time.perf_counter()
            drv.mb200_bench_step(h, args.steps)
            host.meep_b200_sync_host(fptr)
            cold_s = time.perf_counter() - t0
            c1 = stats()
            cold = {"value": cells * args.steps / cold_s, "unit": "cell-updates/s", "seconds": cold_s,
                    "h2d_bytes": c1[1] - c0[1], "d2h_bytes": c1[2] - c0[2],
                    "what": "K steps starting and ending with all field arrays in (pageable) host memory"}

        # ---- profiling pass (separate, not part of any reported throughput): per-kind CUDA events ---
        lib.mb200_profile_reset(ctx)
        lib.mb200_profile_enable(ctx, 1)
        nprof = min(args.steps, 10)
        drv.mb200_bench_step(h, nprof)
        lib.mb200_profile_enable(ctx, 0)
        prof = {}
        kinds = ["curl", "edhb", "lorentz", "fmp", "source", "halo", "zero", "dft", "flux", "step3", "beta", "exchange",
                 "cylint", "cylr0", "step3_pml", "bfast", "average", "gyro", "noise"]
        for k, name in enumerate(kinds):
            n_, ms_, by_ = C.c_int64(), C.c_double(), C.c_double()
            lib.mb200_profile_get(ctx, k, C.byref(n_), C.byref(ms_), C.byref(by_))
            if n_.value:
                prof[name] = {"launches_per_step": n_.value / nprof, "ms_per_step": ms_.value / nprof,
                              "alg_bytes_per_step": by_.value / nprof}

        peaks, peak_src = measured_peaks()
        peak = float(peaks.get("hbm_gbs", 6650.0))
        alg_step = drv.mb200_bench_algorithmic_bytes_per_step(h)
        # kernels that skip known-zero polarisation blocks move less than the dense model: never credit
        # bytes that were not moved (the per-kernel figures count the blocks actually processed)
        kernel_sum = sum(v["alg_bytes_per_step"] for v in prof.values())
        kernel_sum_smaller = bool(prof) and kernel_sum < alg_step
        if kernel_sum_smaller:
            alg_step = kernel_sum
        dom = max(prof.items(), key=lambda kv: kv[1]["ms_per_step"]) if prof else (None, None)
        roofline = None
        if dom[0]:
            d = dom[1]
            per_launch_bytes = d["alg_bytes_per_step"] / d["launches_per_step"]
            per_launch_ms = d["ms_per_step"] / d["launches_per_step"]
            achieved = per_launch_bytes / (per_launch_ms * 1e-3) / 1e9
            roofline = {"bound": "hbm", "kernel": dom[0], "achieved": achieved, "peak": peak, "unit": "GB/s",
                        "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                        "alg_bytes_per_launch": per_launch_bytes, "ms_per_launch": per_launch_ms,
                        "share_of_step": d["ms_per_step"] / sum(v["ms_per_step"] for v in prof.values()),
                        "whole_step": {"alg_bytes_per_step": alg_step,
                                       "achieved": alg_step / (dev_ms / args.steps * 1e-3) / 1e9,
                                       "frac": alg_step / (dev_ms / args.steps * 1e-3) / 1e9 / peak,
                                       "bytes_per_cell": alg_step / (cells / ranks_per_problem),
                                       "model": "SURVEY 8d accounting on the actual chunk layout (every array element "
                                                "a half-step must read / write counted once; halo copies not counted)"
                                                + ("; polarisation arrays counted only over the blocks the kernels "
                                                   "do not skip (zero-block flags, measured)" if kernel_sum_smaller else "")},
                        "kernels": prof}

        # DRAM bytes per launch (= per plan run: the fast path of a half-step may be several kernels) of the
        # dominant kernel from a committed ncu capture (`--set full`, or a pass with just the dram byte
        # counters) of this very configuration (profiles/traffic_<workload>_<n>_<prec>.json).  The capture names the kernel
        # sources it was taken on: a capture of another build is reported as such, not as this build's.
        if roofline and world == 1:
            tf = os.path.join(ROOT, "profiles", "traffic_%s_%d_%s.json" % (args.workload, n, args.prec))
            try:
                with open(tf) as fh:
                    t = json.load(fh)
                if t.get("kernel") == dom[0]:
                    roofline["traffic"] = t["dram_bytes_per_launch"]
                    roofline["traffic_source"] = "%s (%s)" % (os.path.relpath(tf, ROOT), t.get(
                        "how", "ncu, dram__bytes_read.sum + dram__bytes_write.sum, mean over the launches of the kernel"))
                    roofline["traffic_build"] = t.get("csrc_stamp")
                    roofline["traffic_build_is_this_build"] = t.get("csrc_stamp") == csrc_stamp()
            except (OSError, KeyError, ValueError):
                pass
        if roofline:
            roofline["csrc_stamp"] = csrc_stamp()

        nproblems = world // ranks_per_problem
        if roofline and world > 1:
            roofline["note"] = "rank 0's share of the cell; kernels are identical on every rank"
        value = cells * args.steps * nproblems / (dev_ms * 1e-3)
        line = {"metric": "Yee cell-updates/s", "value": value, "unit": "cell-updates/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
                "higher_is_better": True, "scaling": scaling_label(args, world), "vs_baseline": None,
                "dtype": args.prec, "data": "synthetic",
                "config": {"workload": workload_text(args.workload, nx, n, nz, world),
                           "n": n, "cell": [nx, n, nz], "num_chunks": drv.mb200_bench_num_chunks(h),
                           "parallelism": ("sharded: split_by_cost over %d ranks, device-to-device halo exchange" % world)
                           if sharded else ("replicas only" if world > 1 else "single GPU"),
                           "l2_policy": "inputs larger than L2: %.1f GB of field arrays streamed per step vs 126 MB L2"
                                        % (drv.mb200_bench_field_bytes(h) / 1e9) + " (per rank)",
                           "setup_s": t_setup, "warmup_s": t_warm,
                           "host_max_rss_gb": __import__("resource").getrusage(
                               __import__("resource").RUSAGE_SELF).ru_maxrss / 1048576.0,
                           "host_threads": int(os.environ.get("OMP_NUM_THREADS", "1"))},
                "e2e": {"value": cells * args.steps * nproblems / e2e_s, "unit": "cell-updates/s",
                        "h2d_bytes_per_step": e2e_h2d,
                        "d2h_bytes_per_step": e2e_d2h,
                        "what": "K x (fields::step() + fields::get_field probe) through the meep C++ API; field "
                                "arrays stay resident in HBM between steps (state, like model weights)",
                        "cold_start": cold},
                "probe": probe,
                "gpu_launches": launches,
                "clocks": sampler.summary(),
                "roofline": roofline}
        drv.mb200_bench_destroy(h)
        live.remove(h)
        return line

    north_star_n1 = world == 1 and args.workload == "c2" and args.n == 1024 and args.auto_n
    try:
        line = measure(args.n)
    except Exception as e:  # the 1024^3 cell needs ~105 GB of HBM and ~40 GB of host RAM: if this box
        if not north_star_n1:  # cannot hold it, the 1-GPU line falls back to BASELINE.json configs[1]
            raise
        sys.stderr.write("bench.py: 1024^3 on one GPU failed (%s); falling back to 512^3\n" % e)
        for hh in list(live):
            drv.mb200_bench_destroy(hh)
            live.remove(hh)
        args.n = 512
        north_star_n1 = False
        line = measure(512)
    if north_star_n1:
        # BASELINE.json configs[1] (512^3, the reference's 1-GPU case) measured in the same run
        try:
            c1 = measure(512, light=True)
            line["configs1_512"] = {k: c1[k] for k in ("value", "unit", "ms_per_step", "e2e", "gpu_launches")}
            line["configs1_512"]["workload"] = c1["config"]["workload"]
            line["configs1_512"]["roofline"] = {k: c1["roofline"][k] for k in
                                                ("kernel", "achieved", "peak", "frac", "whole_step", "traffic")
                                                if k in c1["roofline"]}
        except Exception as e:
            line["configs1_512"] = {"failed": str(e)}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        a2 = argparse.Namespace(**vars(args))
        a2.steps, a2.warmup = args.cpu_steps, 3
        try:
            line["cpu_baseline"] = cpu_baseline_obj(run_reference(a2))
        except Exception as e:  # the baseline is reported, never required for the device number
            line["cpu_baseline"] = {"value": None, "unit": "cell-updates/s", "cores": os.cpu_count(),
                                    "kind": "reference", "sample": "failed: %s" % e}
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
