#!/usr/bin/env python
"""bench.py -- headline benchmark: CG time-to-solution on the 20k x 20k 5-point FP64 stencil
(BASELINE.json: "CG solve ms & SpMV HBM GB/s, 20k^2 5-pt FP64 stencil, 1/2/4/8 B200").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--grid n]

One "step" = one full CG solve (b = 1, x0 = 0, tol 1e-6, the reference CLI's defaults,
src/main/cg_solver.cu:50-51,124-128) through the reference-facing entry points of
libspmv_b200.so.  N > 1 runs one process per GPU (torchrun): torch.distributed is only plumbing
(rendezvous, IPC-handle all-gather, barrier, max over ranks); halos and scalar reductions move
through the library's own peer-memory kernels.  Strong scaling: the 20k^2 problem is split into
N row bands.

Printed JSON line (rank 0): see the keys below.  `value` = device-timed solve (inputs resident in
HBM, the reference's own timing scope, cg_solver.cu:494,640), max over ranks, mean over K steps.
`e2e` = the same solve through cg_solve_device / cg_solve_mgpu_partitioned with HOST (pinned)
buffers, wall clock, copies inside.  `roofline` = the fused STENCIL5 SpMV + p.Ap kernel, timed
with CUDA events on its stream inside the timed steps.
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
sys.path.insert(0, os.path.join(ROOT, "cuda-spmv-benchmark_b200", "python"))

METRIC = "cg_solve_ms_20k_stencil5_fp64"
TOL, MAX_ITERS = 1e-6, 1000


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--grid", type=int, default=20000)
    ap.add_argument("--weak-cap32", action="store_true", help="with --weak: keep the row count below 2^31")
    ap.add_argument("--weak", action="store_true",
                    help="weak scaling (BASELINE configs[4]): 20k x 20k rows PER GPU, square grid of side "
                         "20000*sqrt(N) rounded to a multiple of 2N (bands stay grid-row aligned); not the headline metric")
    ap.add_argument("--cpu-grid", type=int, default=0,
                    help="CPU arm: run (and report) this grid instead of --grid (0 = --grid, reduced only if host RAM is short)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-operators", action="store_true", help="skip the 10k x 10k operator comparison (configs[1])")
    ap.add_argument("--no-weak-leg", action="store_true",
                    help="N > 1 at the default grid: skip the short weak-scaling measurement (configs[4]) added to the line")
    ap.add_argument("--single-process", action="store_true",
                    help="N > 1 without torchrun: ONE process drives all N GPUs (one enqueue thread per GPU), the "
                         "north-star topology / the reference CLI's `cg_solver_mgpu_stencil`")
    ap.add_argument("--host-buffers", default="near", choices=["near", "torch"],
                    help="pinned host vectors: 'near' = library allocator, pages on the GPU's NUMA node; 'torch' = pin_memory()")
    ap.add_argument("--timers-every", type=int, default=10,
                    help="record the per-phase CUDA events (roofline.avg_launch_ms) on every K-th timed step only: an event "
                         "between two kernels keeps the second one from being scheduled under the tail of the first "
                         "(programmatic dependent launch); 1 = every step")
    return ap.parse_args()


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (the reference has no CPU compute path -- BASELINE.md section 3)
# ------------------------------------------------------------------------------------------------
KAT_20K = {"iterations": 14, "residual_norm": 1.08428e-2, "solution_sum": 3.9995055965e8, "solution_norm": 1.9997869532e4}
KAT_10K = {"iterations": 14, "solution_sum": 9.9975281007e7, "solution_norm": 9.9978695581e3}


def mem_available_bytes():
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable:"):
                return int(ln.split()[1]) * 1024
    except Exception:
        pass
    return None


class CpuCG:
    """The CPU arm: full CG to convergence (b = 1, x0 = 0, tol 1e-6) with the OpenMP oracle port of
    cg_solve_device on the CSR stencil operator, on the grid it is asked for.  Nothing is scaled or
    extrapolated: the grid that runs is the grid that is reported.  If the host cannot hold the
    matrix (112 B/row: CSR 64 + five vectors 40 + conversion slack) the grid is reduced and the
    reduced grid is what `config` names."""
    BYTES_PER_ROW = 112.0

    def __init__(self, grid, forced_grid=0):
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import numpy as np
        import orc
        self.np, self.orc = np, orc
        self.threads = orc.num_threads()
        self.requested = grid
        n = forced_grid or grid
        avail = mem_available_bytes()
        self.reduced = None
        if not forced_grid and avail is not None and self.BYTES_PER_ROW * n * n > 0.85 * avail:
            n = int((0.85 * avail / self.BYTES_PER_ROW) ** 0.5) // 1000 * 1000
            self.reduced = "host has %.0f GB available, %dx%d needs %.0f GB" % (avail / 1e9, grid, grid,
                                                                                 self.BYTES_PER_ROW * grid * grid / 1e9)
        self.n = max(n, 3)
        t0 = time.perf_counter()
        rp64, self.ci, self.va = orc.stencil5_csr_direct(self.n)
        self.rp = rp64.astype(np.int32)
        del rp64
        self.b = np.ones(self.n * self.n)
        self.setup_s = time.perf_counter() - t0
        self.last = None

    def solve(self):
        """one full solve -> ms"""
        x0 = self.np.zeros(self.n * self.n)
        t0 = time.perf_counter()
        _, res, _ = self.orc.cg_device(self.rp, self.ci, self.va, self.n, 1, self.b, x0, MAX_ITERS, TOL, hist=0, inplace=True)
        ms = (time.perf_counter() - t0) * 1e3
        self.last = res
        if not res["converged"]:
            raise SystemExit("CPU arm: CG did not converge")
        return ms

    def workload(self):
        return "cg_%dx%d_stencil5_b1_x0_tol1e-6" % (self.n, self.n)

    def sample(self, solves, ms):
        s = ("full CG to convergence (oracle port of cg_solve_device, stencil5-csr operator, OpenMP %d threads) on the %dx%d grid "
             "(%d rows, %d iterations, %d timed solve(s), %.0f ms each; matrix build %.1f s untimed); measured, not scaled"
             % (self.threads, self.n, self.n, self.n * self.n, self.last["iterations"], solves, ms, self.setup_s))
        if self.n != self.requested:
            s += "; REDUCED from %dx%d: %s" % (self.requested, self.requested, self.reduced or "--cpu-grid")
        return s


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # torchrun exports OMP_NUM_THREADS=1 to every rank unless the caller set it; the CPU arm runs on
    # rank 0 alone and is meant to use all host cores
    if os.environ.get("OMP_NUM_THREADS") == "1" and ("TORCHELASTIC_RUN_ID" in os.environ or "LOCAL_RANK" in os.environ):
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    cpu = CpuCG(args.grid, args.cpu_grid)
    for _ in range(args.warmup):
        cpu.solve()
    vals = [cpu.solve() for _ in range(args.steps)]
    val = sum(vals) / len(vals)
    n = cpu.n
    line = {"metric": METRIC, "value": val, "unit": "ms", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": val, "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": cpu.workload(), "grid": n, "rows": n * n, "nnz": 5 * n * n - 4 * n, "tol": TOL,
                       "iterations": cpu.last["iterations"], "operator": "stencil5-csr (oracle port)",
                       "requested_grid": args.grid, "host_threads": cpu.threads},
            "cg": {k: cpu.last[k] for k in ("iterations", "residual_norm", "solution_sum", "solution_norm")},
            "cpu_baseline": {"value": val, "unit": "ms", "cores": cpu.threads, "kind": "port",
                             "sample": cpu.sample(len(vals), val)},
            "e2e": {"value": val, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# configs[1]: 10k x 10k single-GPU SpMV, STENCIL5 vs generic CSR vs generic ELLPACK (extra keys)
# ------------------------------------------------------------------------------------------------
def operator_table(L, B, torch, peak, n=10000, reps=10, warm=5):
    """Kernel-only times (CUDA events on the launching stream, median of `reps` after `warm`
    warm-ups, the reference's own protocol -- src/main/main.cu:136-137, benchmark_stats.cu:39-89)
    through the C ABI, x = 1.  GB/s from the ALGORITHMIC bytes of SURVEY.md section 8(d)."""
    N, nnz = n * n, 5 * n * n - 4 * n
    s = torch.cuda.current_stream().cuda_stream
    dp = lambda t: C.c_void_p(t.data_ptr())

    def timed(fn):
        for _ in range(warm):
            fn()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2]

    def row(ms, nbytes, y):
        gbs = nbytes / (ms * 1e-3) / 1e9
        return {"ms": round(ms, 4), "gb_s": round(gbs, 1), "frac_of_peak": round(gbs / peak, 3),
                "bytes": nbytes, "checksum_ok": float(y.sum().item()) == N + 4 * n}

    out = {"grid": n, "rows": N, "nnz": nnz, "x": "ones", "protocol": "%d warm-up + median of %d" % (warm, reps)}
    rp = torch.empty(N + 1, dtype=torch.int32, device="cuda")
    ci = torch.empty(nnz + 2, dtype=torch.int32, device="cuda")
    va = torch.zeros(nnz + 2, dtype=torch.float64, device="cuda")
    B.check(L.b200_gen_stencil5_csr(n, 0, N, 5.0, -1.0, dp(rp), dp(ci), dp(va), s), "gen csr")
    x = torch.ones(N, dtype=torch.float64, device="cuda")
    y = torch.empty(N, dtype=torch.float64, device="cuda")
    ms = timed(lambda: B.check(L.b200_spmv_stencil5_csr(dp(rp), dp(ci), dp(va), dp(x), dp(y), N, n, s), "stencil5"))
    out["stencil5-csr"] = row(ms, 8.0 * nnz + 16.0 * N, y)
    plan = B.CsrPlan()
    B.check(L.b200_csr_plan_build(dp(rp), N, nnz, C.byref(plan), s), "plan")
    y.fill_(float("nan"))
    ms = timed(lambda: B.check(L.b200_spmv_csr(C.byref(plan), dp(rp), dp(ci), dp(va), dp(x), dp(y), N, 1.0, 0.0, s), "csr"))
    out["csr"] = row(ms, 12.0 * nnz + 4.0 * (N + 1) + 16.0 * N, y)
    del rp, ci, va
    idx = torch.empty(5 * N + 2, dtype=torch.int32, device="cuda")
    val = torch.empty(5 * N + 2, dtype=torch.float64, device="cuda")
    B.check(L.b200_gen_stencil5_ellpack(n, 0, N, 5.0, -1.0, dp(idx), dp(val), s), "gen ell")
    y.fill_(float("nan"))
    ms = timed(lambda: B.check(L.b200_spmv_ellpack(dp(idx), dp(val), dp(x), dp(y), N, 5, 1.0, 0.0, s), "ellpack"))
    out["ellpack"] = row(ms, 76.0 * N, y)
    y.fill_(float("nan"))
    ms = timed(lambda: B.check(L.b200_spmv_stencil5_ellpack(dp(val), dp(idx), dp(x), dp(y), N, 5, 1.0, 0.0, n, s), "st-ell"))
    out["stencil5-ellpack"] = row(ms, 56.0 * N, y)
    return out


def weak_leg(L, B, torch, dist, world, rank, local_rank, barrier):
    """configs[4] inside the strong-scaling line: 20000^2 rows PER GPU (reference recipe
    scripts/benchmarking/benchmark_weak_scaling.sh:15-21, square grids n0*sqrt(P)), 1 warm-up + 3 timed solves
    through the same public entry point.  The exchange blocks are re-created for the larger grid.
    Every step that can fail on one rank is followed by a MAX all-reduce of the failure flag, so that all ranks
    leave together (a rank that raised alone would leave the others in a collective)."""
    import math
    import mgpu_bootstrap

    def agree(failed, what):
        t = torch.tensor([1.0 if failed else 0.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if float(t[0]) != 0.0:
            raise RuntimeError("weak-scaling leg: %s failed on some rank" % what)

    step = 2 * world
    n = int(round(20000 * math.sqrt(world) / step)) * step
    N = n * n
    L.b200_mgpu_finalize()
    handle = (C.c_ubyte * mgpu_bootstrap.HANDLE_BYTES)()
    agree(L.b200_mgpu_init_rank(rank, world, local_rank, n, handle) != 0, "b200_mgpu_init_rank")
    raw = mgpu_bootstrap.all_gather_handles(dist, bytes(handle), "cuda")
    agree(L.b200_mgpu_connect((C.c_ubyte * len(raw)).from_buffer_copy(raw)) != 0, "b200_mgpu_connect")
    nl, off = mgpu_bootstrap.partition(N, world, rank)
    mat = b_host = x_host = None
    try:
        mat = B.HostMatrix.synthetic_stencil(n)
        b_host = torch.full((nl,), 1.0, dtype=torch.float64).pin_memory()
        x_host = torch.zeros(nl, dtype=torch.float64).pin_memory()
        ok = True
    except Exception:
        ok = False
    agree(not ok, "host buffer allocation")
    stats = B.CGStatsMultiGPU()
    cfg = B.cg_config(MAX_ITERS, TOL, 0, 0)
    times, kats = [], []
    for step_i in range(4):
        x_host.zero_()
        barrier()
        rc = L.cg_solve_mgpu_partitioned(None, mat.ptr(), b_host.data_ptr() - off * 8, x_host.data_ptr() - off * 8, cfg,
                                         C.byref(stats))
        torch.cuda.synchronize()
        t = torch.tensor([stats.time_total_ms, 1.0 if (rc != 0 or not stats.converged) else 0.0], dtype=torch.float64,
                         device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if float(t[1]) != 0.0:
            raise RuntimeError("weak-scaling leg: a solve failed on some rank (rc=%d here)" % rc)
        if step_i > 0:
            times.append(float(t[0]))
            kats.append((stats.iterations, stats.residual_norm, stats.solution_sum))
    agree(len(set(kats)) != 1, "bit-reproducibility of the solves")
    ms = sum(times) / len(times)
    rows_rank0 = mgpu_bootstrap.partition(N, world, 0)[0]
    return {"config": "20000^2 rows per GPU (BASELINE.json configs[4])", "grid": n, "rows": N, "rows_per_gpu": rows_rank0,
            "iterations": stats.iterations, "ms": round(ms, 4), "ms_steps": [round(v, 3) for v in times],
            "ms_per_iteration": round(ms / stats.iterations, 4), "warmup": 1, "steps": 3,
            "solution_sum": stats.solution_sum, "residual_norm": stats.residual_norm,
            "note": "device-timed, max over ranks; compare ms_per_iteration with the one-GPU 20000^2 line (value / 14): "
                    "equal per-GPU work, weak-scaling efficiency = that ratio"}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import numpy as np
    import torch
    import spmv_b200 as B

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device visible; the B200 path has no CPU fallback")
    single = args.single_process and args.gpus > 1
    world = args.gpus if single else int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and not single:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py: --gpus %d needs torchrun (one process per GPU) or --single-process" % args.gpus)
    torch.cuda.set_device(local_rank)
    L = B.load()
    if args.weak:
        # the reference's weak-scaling recipe (scripts/benchmarking/benchmark_weak_scaling.sh:15-21: constant
        # unknowns per GPU, square grids n0*sqrt(P)).  8 GPUs = 56576^2 = 3.2e9 unknowns: beyond the 32-bit
        # MatrixData.rows of the reference API; the synthetic-stencil path goes by grid_size (64-bit row ids,
        # columns modulo 2^32, < 2^31 non-zeros per band).  --weak-cap32 keeps N below 2^31 instead.
        import math
        step = 2 * world
        args.grid = int(round(20000 * math.sqrt(world) / step)) * step
        if args.weak_cap32:
            args.grid = min(args.grid, 46340 // step * step)
    n = args.grid
    N = n * n
    dist = None
    saved_stdout = None
    if single:
        devs = (C.c_int * world)(*range(world))
        if L.b200_mgpu_init_single_process(world, devs, n) != 0:
            raise SystemExit("b200_mgpu_init_single_process failed")
    elif world > 1:
        # NCCL prints its version banner straight to fd 1 when the communicator comes up: keep stdout for
        # the one JSON line, send everything else to stderr until then
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        import mgpu_bootstrap
        mgpu_bootstrap.connect(L, dist, rank, world, local_rank, n)
        dist.barrier()

    def barrier():
        if single:
            for d in range(world):
                torch.cuda.synchronize(d)
            return
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    mat = B.HostMatrix.synthetic_stencil(n)
    import mgpu_bootstrap
    nl, off = (N, 0) if single else mgpu_bootstrap.partition(N, world, rank)
    # pinned host buffers of the local slice; the solver only touches [off, off+nl) of b and x
    host_node = None
    near_ptrs = []

    def host_vector(fill):
        """pinned host vector of this rank's rows: the library's NUMA-aware allocator (host/host_alloc.cpp) or torch"""
        nonlocal host_node
        if args.host_buffers == "near" and not single:
            p, node = C.c_void_p(), C.c_int(-1)
            if L.b200_host_alloc_near(local_rank, 8 * nl, C.byref(p), C.byref(node)) == 0:
                host_node = node.value
                near_ptrs.append(p)
                t = torch.frombuffer((C.c_double * nl).from_address(p.value), dtype=torch.float64)
                t.fill_(fill)
                return t
        return torch.full((nl,), fill, dtype=torch.float64).pin_memory()

    b_host = host_vector(1.0)
    x_host = host_vector(0.0)
    b_ptr = b_host.data_ptr() - off * 8
    x_ptr = x_host.data_ptr() - off * 8
    cfg_timed = B.cg_config(MAX_ITERS, TOL, 0, 1)  # detailed timers: event records only, no extra syncs
    cfg_plain = B.cg_config(MAX_ITERS, TOL, 0, 0)

    if world == 1:
        op = L.get_operator(b"stencil5-csr")
        if op.contents.init(mat.ptr()) != 0:
            raise SystemExit("operator init failed")
        stats = B.CGStats()

        def solve(cfg):
            return L.cg_solve_device(op, mat.ptr(), b_ptr, x_ptr, cfg, C.byref(stats))
    else:
        stats = B.CGStatsMultiGPU()

        def solve(cfg):
            return L.cg_solve_mgpu_partitioned(None, mat.ptr(), b_ptr, x_ptr, cfg, C.byref(stats))

    # The timed steps upload BOTH host vectors every step, as the e2e contract asks (every input of the step is copied
    # inside the timed region).  The library's default would notice that x0 is all zeros and clear the device vector
    # instead (host scan overlapped with the upload of b): that path is measured separately below (e2e.zero_guess_detection).
    L.b200_cg_set_skip_zero_x0(0)
    for _ in range(args.warmup):
        x_host.zero_()
        barrier()
        if solve(cfg_timed) != 0:
            raise SystemExit("warm-up solve failed")
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = L.b200_launch_count()
    dev_ms, wall_ms, k1_ms, k1_cnt, iters = [], [], 0.0, 0, None
    ph = (C.c_double * 9)()
    pc = (C.c_int * 9)()
    tl = (C.c_double * 8)()
    tc = (C.c_int * 8)()
    gp = (C.c_double * 8)()
    gap_ms, steps_without_events = [0.0] * 8, 0
    barrier()
    t_block0 = time.perf_counter()
    phase_sum = [0.0] * 9
    tail_ms, tail_n, steps_with_events = [0.0] * 8, [0] * 8, 0
    kats = []
    for step in range(args.steps):
        x_host.zero_()  # initial guess x0 = 0 (not part of the solve)
        barrier()
        with_events = (step % max(args.timers_every, 1) == 0)
        t0 = time.perf_counter()
        rc = solve(cfg_timed if with_events else cfg_plain)
        if single:
            for d in range(world):
                torch.cuda.synchronize(d)
        else:
            torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        if rc != 0:
            raise SystemExit("solve failed rc=%d" % rc)
        d = stats.time_total_ms
        if dist is not None:  # slowest rank defines the step
            t = torch.tensor([d, wall], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            d, wall = float(t[0]), float(t[1])
        dev_ms.append(d)
        wall_ms.append(wall)
        if with_events:
            L.b200_last_phase_times(ph, pc)
            k1_ms += ph[1]
            k1_cnt += pc[1]
            steps_with_events += 1
            for t in range(9):
                phase_sum[t] += ph[t]
        L.b200_last_tail_times(tl, tc)
        L.b200_last_gap_times(gp)
        for t in range(8):
            tail_ms[t] += tl[t]
            tail_n[t] += tc[t]
            if not with_events:
                gap_ms[t] += gp[t]
        steps_without_events += 0 if with_events else 1
        iters = stats.iterations
        h2d_bytes = L.b200_last_h2d_bytes() // (world if single else 1)  # per rank, counted by the library
        if not stats.converged:
            raise SystemExit("CG did not converge")
        kats.append((stats.iterations, stats.residual_norm, stats.solution_sum, stats.solution_norm))
    barrier()
    block_ms = (time.perf_counter() - t_block0) * 1e3
    launches = L.b200_launch_count() - launches0
    # the library's default behaviour for a zero initial guess, 3 extra solves (not part of `value` / `e2e.value`)
    L.b200_cg_set_skip_zero_x0(1)
    zg_wall, zg_bytes = [], 0
    for _ in range(3):
        x_host.zero_()
        barrier()
        t0 = time.perf_counter()
        rc = solve(cfg_plain)
        if single:
            for d_ in range(world):
                torch.cuda.synchronize(d_)
        else:
            torch.cuda.synchronize()
        w_ = (time.perf_counter() - t0) * 1e3
        if rc != 0 or not stats.converged:
            raise SystemExit("solve failed rc=%d" % rc)
        if dist is not None:
            t = torch.tensor([w_], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            w_ = float(t[0])
        zg_wall.append(w_)
        zg_bytes = L.b200_last_h2d_bytes() // (world if single else 1)
        if (stats.iterations, stats.residual_norm, stats.solution_sum, stats.solution_norm) != kats[0]:
            raise SystemExit("bench.py: the zero-guess path changed the result")
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    phase_sum = [v / max(steps_with_events, 1) for v in phase_sum]
    # parity gate inside the bench: the timed solves must reproduce the known answers of this grid
    #  on any number of GPUs, bit-identically from step to step
    kat = {20000: KAT_20K, 10000: KAT_10K}.get(n) if not args.weak else None
    if len(set(kats)) != 1:
        raise SystemExit("bench.py: timed solves are not bit-reproducible: %r" % sorted(set(kats)))
    if kat is not None:
        import math
        it_k, res_k, sum_k, norm_k = kats[0]
        bad = [k for k, got, tol in (("iterations", it_k, 0), ("solution_sum", sum_k, 1e-10), ("solution_norm", norm_k, 1e-10),
                                     ("residual_norm", res_k, 1e-5))
               if k in kat and not (got == kat[k] if tol == 0 else math.isclose(got, kat[k], rel_tol=tol))]
        if bad:
            raise SystemExit("bench.py: known-answer mismatch on %s: got %r, expected %r" % (bad, kats[0], kat))

    ms = sum(dev_ms) / len(dev_ms)
    e2e_ms = sum(wall_ms) / len(wall_ms)
    peak, peak_src = measured_peak()
    # per-RANK band (the roofline / byte models below are per GPU); a single process holds all the bands
    rows_local, off_rank = mgpu_bootstrap.partition(N, world, 0) if single else (nl, off)
    nnz_local = L.b200_stencil5_nnz_before(off_rank + rows_local, n) - L.b200_stencil5_nnz_before(off_rank, n)
    # Dominant kernel: the STENCIL5 SpMV.  Deferred-x schedule (default): the first launch of a solve is
    # the plain SpMV + p.Ap (values + p + Ap = 8 nnz + 16 N bytes), every later one is the fused
    # direction-update SpMV (values + r + p_old + x in, p_new + x + Ap out = 8 nnz + 48 N bytes);
    # B200_CG_SCHEDULE=classic runs the plain one every iteration.
    deferred_x = os.environ.get("B200_CG_SCHEDULE", "") != "classic"
    sweep_family = L.b200_cg_get_kernel() == 1
    xdepth = L.b200_cg_set_xdepth(0) if (deferred_x and sweep_family) else 1  # 0 = query; ring kernels: depth 1
    kname = "stencil5_sweep_kernel" if sweep_family else "stencil5_kernel"
    dot_bytes = 8.0 * nnz_local + 16.0 * rows_local

    def fused_bytes_nx(nx):
        """values + r + p_old in, p_new + Ap out; a launch that retires nx pending x updates also reads and
        writes x once and reads the nx - 1 older directions"""
        return 8.0 * nnz_local + 32.0 * rows_local + ((16.0 + 8.0 * (nx - 1)) * rows_local if nx > 0 else 0.0)

    # launches of one solve: iteration 0 is the plain SpMV + p.Ap, iteration it >= 1 the fused one, which
    # retires the last `xdepth` x updates when it % xdepth == 0
    nx_of = [None] + [(xdepth if it % xdepth == 0 else 0) for it in range(1, iters or 1)]
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    measured = {}
    if os.path.exists(tp) and world == 1:  # the ncu captures are of one-GPU launches: not a measurement of a band
        try:
            measured = json.load(open(tp))
        except Exception:
            measured = {}
    if deferred_x and iters and iters > 1:
        per_launch = [dot_bytes] + [fused_bytes_nx(nx) for nx in nx_of[1:]]
        k1_bytes = sum(per_launch) / iters  # mean over the launches of a solve
        mix = {nx: nx_of[1:].count(nx) for nx in sorted(set(nx_of[1:]))}
        k1_name = ("%s<ST_FUSED*>: p = r + beta p, SpMV, p.Ap, and x += alpha p for the last %d iterations in every "
                   "%s launch; per solve of %d iterations: 1 x ST_DOT, %s"
                   % (kname, xdepth, "launch" if xdepth == 1 else "%d-th" % xdepth, iters,
                      ", ".join("%d x %s" % (c, "ST_FUSED_X%d" % nx if nx != 1 else "ST_FUSED") for nx, c in mix.items())))
        fam = "sweep" if sweep_family else "ring"
        keys = ["stencil5_%s_dot_dram_bytes_per_launch_%d" % (fam, n)] + \
               ["stencil5_%s_fused_x%d_dram_bytes_per_launch_%d" % (fam, nx, n) for nx in nx_of[1:]]
        traffic = (sum(measured[k] for k in keys) / iters) if all(k in measured for k in keys) else None
        traffic_alg = k1_bytes
        pending = iters - xdepth * ((iters - 1) // xdepth)
        iter_bytes = (96.0 + 8.0 + 8.0 / xdepth) * rows_local  # K1F 72 + K2r 24 + x (16 + 8 (d - 1)) / d
        solve_bytes = 72.0 * rows_local + sum(per_launch) + 24.0 * iters * rows_local + (16.0 + 8.0 * pending) * rows_local
        schedule = ("deferred-x, depth %d (2 launches, %.1f B/row per iteration; %s kernels)" % (xdepth, iter_bytes / rows_local, fam))
    else:
        k1_bytes = dot_bytes
        k1_name = "%s<ST_DOT> (SpMV + p.Ap)" % kname
        k = "stencil5_%s_dot_dram_bytes_per_launch_%d" % ("sweep" if sweep_family else "ring", n)
        traffic = measured.get(k)
        traffic_alg = dot_bytes
        iter_bytes = 128.0 * rows_local          # K1 56 + K2 48 + K3 24
        solve_bytes = (72.0 + 128.0 * iters) * rows_local
        schedule = "classic (3 launches, 128 B/row per iteration)"
    k1_avg_ms = k1_ms / max(k1_cnt, 1)
    achieved = k1_bytes / (k1_avg_ms * 1e-3) / 1e9 if k1_cnt else None
    line = {
        "metric": METRIC if not args.weak else "cg_solve_ms_weak_20k_x_20k_rows_per_gpu_stencil5_fp64",
        "value": ms, "unit": "ms", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "ms_steps": [round(v, 3) for v in dev_ms],
        "higher_is_better": False, "scaling": "weak" if args.weak else "strong", "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic", "impl": "b200",
        "config": {"workload": "cg_%dx%d_stencil5_b1_x0_tol1e-6" % (n, n), "grid": n, "rows": N,
                   "nnz": 5 * N - 4 * n, "tol": TOL, "iterations": iters, "operator": "stencil5-csr",
                   "partition": "row bands x%d" % world,
                   "rows_per_gpu": rows_local,
                   "cache": "vectors (3.2 GB each) and matrix (16 GB values) exceed the 126 MB L2; no flush needed"},
        "e2e": {"value": e2e_ms, "unit": "ms", "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": 8 * nl,
                "h2d_note": ("b uploaded; x0 is all zeros: the library scans the caller's vector (host threads, while b is on its way) "
                             "and clears the device vector instead of uploading it (B200_SKIP_ZERO_X0=0 uploads it)")
                if h2d_bytes < 16 * nl else "b and x0 uploaded every step (zero-guess detection switched off for these steps)",
                "zero_guess_detection": {"value": round(sum(zg_wall[1:]) / max(len(zg_wall) - 1, 1), 3), "unit": "ms",
                                         "h2d_bytes_per_step": int(zg_bytes), "steps": len(zg_wall) - 1,
                                         "note": "library default: the caller's x0 is scanned on host threads while b is uploaded; all zeros -> "
                                                 "cudaMemset instead of the upload.  Same iterates, checked.  Not the headline: `value` above "
                                                 "copies every input"},
                "api": "cg_solve_device" if world == 1 else "cg_solve_mgpu_partitioned",
                "pcie_gbs_per_rank": round((h2d_bytes + 8.0 * nl) / max((e2e_ms - ms) * 1e-3, 1e-9) / 1e9, 2),
                "host_buffers": ("pinned, first-touched on NUMA node %s (GPU's own node: %s)"
                                 % (host_node, L.b200_host_node_of_device(local_rank))) if near_ptrs else "pinned (torch pin_memory)",
                "block_wall_ms_per_step": block_ms / args.steps},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": k1_name, "achieved": achieved,
                     "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None,
                     "frac_of_nominal_7700GBs": (achieved / 7700.0) if achieved else None,
                     "peak_note": "the measured peak is a device copy (50 % writes); this launch mix writes 24 % of its bytes, "
                                  "so frac can exceed 1 (tools/sweep.py: plain SpMV, 14 % writes, reaches 7.2 TB/s)",
                     "traffic": traffic,
                     "bytes_per_launch": k1_bytes, "traffic_kernel_algorithmic_bytes": traffic_alg,
                     "avg_launch_ms": k1_avg_ms, "launches_timed": k1_cnt},
        "spmv": {"ms": k1_avg_ms, "gb_s": achieved, "note": "per-GPU fused SpMV launch inside CG (see roofline.kernel)"},
        "cg": {"iterations": iters, "residual_norm": stats.residual_norm, "solution_sum": stats.solution_sum,
               "solution_norm": stats.solution_norm,
               "schedule": schedule,
               "topology": ("one process, %d GPUs, one enqueue thread per GPU" % world) if single
               else ("one process per GPU (torchrun), peer memory via CUDA IPC" if world > 1 else "one GPU"),
               "iter_bytes_model": iter_bytes, "solve_bytes_model": solve_bytes,
               "solve_gb_s": solve_bytes / (ms * 1e-3) / 1e9},
        # per-phase CUDA-event times of the steps that recorded them (every --timers-every-th step); the final
        # sums, scalar recurrences and rank exchanges run in the TAIL of the producing kernels (no reduce launches)
        "phases_ms_per_step": dict(zip(["_", "spmv(K1/K1F incl. p.Ap tail)", "reduce_pAp(stand-alone)",
                                        "update_r(K2r incl. r.r tail)" if deferred_x else "update_xr(K2 incl. r.r tail)",
                                        "reduce_rr(stand-alone)", "finish_x" if deferred_x else "update_p(K3)",
                                        "halo_dir(setup)" if deferred_x else "halo_push", "residual_init", "reduce_rr0"],
                                       [round(v, 4) for v in phase_sum])),
        "phase_event_steps": steps_with_events,
        # device-measured (globaltimer) duration of the reduction tails: fixed-order final sum + LL exchange over
        # NVLink peer memory (includes waiting for the slowest rank) + scalar update, per exchange, rank 0
        "tails_us_per_exchange": {nm: (round(1e3 * tail_ms[k] / tail_n[k], 3) if tail_n[k] else None)
                                  for k, nm in ((0, "r0.r0"), (1, "p.Ap"), (2, "r.r"), (3, "checksum"))},
        "tails_ms_per_step": round(sum(tail_ms[k] for k in (0, 1, 2)) / args.steps, 4),
        # device-clock time between consecutive tails in the steps WITHOUT events (programmatic launches overlap
        # there): spmv = halo direction + K1 / K1F + its reduce kernel's group sums, update_r = K2r / K2
        "phases_device_clock_ms_per_step": {"steps": steps_without_events,
                                            "spmv": round(gap_ms[1] / max(steps_without_events, 1), 4),
                                            "update_r": round(gap_ms[2] / max(steps_without_events, 1), 4)},
        "clocks": clocks,
    }
    if world > 1:
        # the halo exchange on its own (it rides inside K2r during a solve): isolated push of one grid row to each
        # neighbour, against NVLink 5's 900 GB/s per direction -- latency-bound by construction (160 KB per edge)
        us, nb = C.c_double(), C.c_longlong()
        if L.b200_mgpu_halo_probe(50, C.byref(us), C.byref(nb)) == 0 and us.value > 0:
            line["halo"] = {"bytes_per_iter_per_direction": int(nb.value), "neighbours": 2,
                            "us": round(us.value, 3), "gb_s_per_direction": round(nb.value / us.value / 1e3, 2),
                            "frac_of_900GBs": round(nb.value / (us.value * 1e-6) / 900e9, 5),
                            "how": "50 back-to-back b200_halo_push launches (peer stores + arrival word), CUDA events, slowest local rank; "
                                   "inside a solve the same stores are issued by K2r and overlap its stream"}
    if world > 1 and not single:
        # what the platform gives every rank when ALL ranks copy at once (the e2e solve does exactly that: every
        # rank uploads b and x0, solves, downloads x): the floor of `e2e` on this box
        dev_buf = torch.empty(nl, dtype=torch.float64, device="cuda")

        def copy_gbs(fn):
            best = None
            for _ in range(2):
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                t_ms = e0.elapsed_time(e1)
                best = t_ms if best is None else min(best, t_ms)
            t = torch.tensor([best], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)  # slowest rank
            return 8.0 * nl / (float(t[0]) * 1e-3) / 1e9
        h2d = copy_gbs(lambda: dev_buf.copy_(b_host, non_blocking=True))
        d2h = copy_gbs(lambda: x_host.copy_(dev_buf, non_blocking=True))
        del dev_buf
        line["e2e"]["pcie_probe"] = {
            "h2d_gbs_per_rank_all_ranks_copying": round(h2d, 2), "d2h_gbs_per_rank_all_ranks_copying": round(d2h, 2),
            "floor_ms": round(ms + 16.0 * nl / h2d / 1e6 + 8.0 * nl / d2h / 1e6, 2),
            "note": "slowest rank, pinned buffers of this run; floor = device solve + this rank's bytes at these rates. "
                    "tools/pcie_probe.py: one GPU alone reaches ~55 GB/s either way on the same box"}
    ref_gpu = os.path.join(ROOT, "tests", "golden", "ref_gpu_10k.json")
    if rank == 0 and os.path.exists(ref_gpu):
        try:  # the reference's OWN kernels (recompiled for sm_100) on a B200 at 10k x 10k, recorded by oracle/run_ref_gpu.sh
            c = json.load(open(ref_gpu))["cases"][0]
            line["extra"] = {"reference_gpu_b200": {
                "grid": c["n"], "how": "oracle/_ref/{spmv_bench,cg_solver} (reference sources, -arch=sm_100) run by oracle/run_ref_gpu.sh on a B200; "
                                       "fixture tests/golden/ref_gpu_10k.json; not the same box as this line",
                "spmv_stencil5_csr_ms": c["spmv"]["stencil5-csr"]["execution_time_ms"],
                "spmv_cusparse_csr_ms": c["spmv"]["cusparse-csr"]["execution_time_ms"],
                "cg_stencil5_csr_ms_per_iteration": round(c["cg"]["stencil5-csr"]["ms_per_iteration"], 4),
                "cg_cusparse_csr_ms_per_iteration": round(c["cg"]["cusparse-csr"]["ms_per_iteration"], 4),
                "cg_note": "the reference CLI measures warm restarts (x is not reset): compare per iteration"}}
        except Exception:
            pass
    if world == 1 and not args.no_operators and not args.weak:
        op.contents.free()
        torch.cuda.empty_cache()
        line["operators_10k"] = operator_table(L, B, torch, peak)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        del b_host, x_host  # 6.4 GB of pinned host memory the CPU arm can use
        for p_ in near_ptrs:
            L.b200_host_free(p_)
        near_ptrs.clear()
        cpu = CpuCG(n, args.cpu_grid)
        v = cpu.solve()
        line["cpu_baseline"] = {"value": v, "unit": "ms", "cores": cpu.threads, "kind": "port", "grid": cpu.n,
                                "workload": cpu.workload(), "sample": cpu.sample(1, v)}
    if world > 1 and not single and not args.weak and not args.no_weak_leg and n == 20000:
        del b_host, x_host
        for p_ in near_ptrs:
            L.b200_host_free(p_)
        near_ptrs.clear()
        try:
            line["weak_scaling"] = weak_leg(L, B, torch, dist, world, rank, local_rank, barrier)
        except Exception as e:  # the leg is an extra: never lose the strong-scaling line over it
            line["weak_scaling"] = {"error": str(e)[:300]}
    if saved_stdout is not None:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        L.b200_mgpu_finalize()
        dist.destroy_process_group()
    elif single:
        L.b200_mgpu_finalize()
    return 0


if __name__ == "__main__":
    a = parse()
    sys.exit(run_reference(a) if a.impl == "reference" else run_b200(a))
