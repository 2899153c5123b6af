#!/usr/bin/env python
"""csr_sweep.py -- generic CSR / ELLPACK SpMV on NON-stencil inputs (one B200; run under gpurun).

  python tools/csr_sweep.py [--rows 10000000] [--out gpurun_out/csr_sweep.json] [--variants 0,6]

Matrices are built on the device (torch): rows of a fixed length L in {5, 12, 20, 40, 200} with uniformly
random distinct-ish sorted columns, and the row-length PROFILE of the reference's `unbalanced_rows`
fixture (tests/helpers/matrix_fixtures.cpp:296-370: 10 % long rows, 40 % rows of 3..7, 50 % rows of
1..3; the long rows are capped at --long entries so that the matrix fits at 1e7 rows).
For every case: kernel time (CUDA events, median of 10 after 5 warm-ups, the reference's protocol),
GB/s against the byte model 12 nnz + 4 (N+1) + 16 N (x counted once: with 1e7 columns the 80 MB
vector lives in the 126 MB L2), and a parity check against a float64 torch reference of the same
product (gather + index_add_) at 1e-12 relative L2."""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-spmv-benchmark_b200", "python"))


def build(torch, lens, cols, seed, band=0):
    """CSR with the given row lengths; columns uniform in [0, cols) -- or, band > 0, within +-band of the
    row (the locality of a discretised PDE) -- sorted inside a row"""
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    rp = torch.zeros(lens.numel() + 1, dtype=torch.int64, device="cuda")
    torch.cumsum(lens, 0, out=rp[1:])
    nnz = int(rp[-1])
    row_of = torch.repeat_interleave(torch.arange(lens.numel(), device="cuda"), lens)
    if band > 0:
        ci = (row_of + torch.randint(-band, band + 1, (nnz,), device="cuda", generator=g, dtype=torch.int64)).clamp_(0, cols - 1)
    else:
        ci = torch.randint(0, cols, (nnz,), device="cuda", generator=g, dtype=torch.int64)
    key = row_of * cols + ci
    key, _ = torch.sort(key)
    ci = (key % cols).to(torch.int32)
    del key, row_of
    va = torch.rand(nnz, device="cuda", generator=g, dtype=torch.float64) * 2 - 1
    return rp.to(torch.int32), ci, va, nnz


def reference(torch, lens, ci, va, x):
    """float64 reference of the same product: per-entry products added into their rows"""
    prod = va * x[ci.long()]
    row_of = torch.repeat_interleave(torch.arange(lens.numel(), device="cuda"), lens)
    return torch.zeros(lens.numel(), dtype=torch.float64, device="cuda").index_add_(0, row_of, prod)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--long", type=int, default=200)
    ap.add_argument("--band", type=int, default=0, help="columns within +-band of the row instead of uniform over all columns")
    ap.add_argument("--variants", default="0,6")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "csr_sweep.json"))
    a = ap.parse_args()
    import torch
    import spmv_b200 as B
    L = B.load()
    peak = 6551.0
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    s = torch.cuda.current_stream().cuda_stream
    dp = lambda t: C.c_void_p(t.data_ptr())
    N = a.rows
    res = {"rows": N, "peak_gbs": peak, "columns": ("within +-%d of the row" % a.band) if a.band else "uniform over all columns",
           "cases": []}

    def timed(fn, warm=5, reps=10):
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

    cases = [("fixed_%d" % k, torch.full((N if k < 100 else N // 4,), k, dtype=torch.int64, device="cuda"))
             for k in (5, 12, 20, 40, 200)]
    i = torch.arange(N, device="cuda")
    lens = torch.where(i < N // 10, torch.full_like(i, a.long), torch.where(i < N // 2, 3 + i % 5, 1 + i % 3))
    cases.append(("unbalanced_rows(10%% x %d, 40%% x 3..7, 50%% x 1..3)" % a.long, lens))
    for name, lens in cases:
        if int(lens.sum()) > 2_100_000_000:
            continue
        N = lens.numel()
        x = torch.rand(N, dtype=torch.float64, device="cuda")
        rp, ci, va, nnz = build(torch, lens, N, 7, a.band)
        # pad the arrays by two entries (the bulk copies move 16-byte granules)
        ci = torch.cat([ci, torch.zeros(2, dtype=torch.int32, device="cuda")])
        va = torch.cat([va, torch.zeros(2, dtype=torch.float64, device="cuda")])
        y = torch.empty(N, dtype=torch.float64, device="cuda")
        plan = B.CsrPlan()
        B.check(L.b200_csr_plan_build(dp(rp), N, nnz, C.byref(plan), s), "plan")
        picked = plan.variant
        yr = reference(torch, lens, ci[:nnz], va[:nnz], x)
        nbytes = 12.0 * nnz + 4.0 * (N + 1) + 16.0 * N
        for v in sorted({picked} | {int(t) for t in a.variants.split(",")}):
            plan.variant = v
            y.fill_(float("nan"))
            ms = timed(lambda: B.check(L.b200_spmv_csr(C.byref(plan), dp(rp), dp(ci), dp(va), dp(x), dp(y), N, 1.0, 0.0, s), "csr"))
            err = float(torch.linalg.norm(y - yr) / torch.linalg.norm(yr))
            r = {"case": name, "rows": N, "nnz": nnz, "variant": v, "picked_by_histogram": v == picked, "ms": round(ms, 4),
                 "gb_s": round(nbytes / ms / 1e6, 1), "frac_of_peak": round(nbytes / ms / 1e6 / peak, 3), "rel_l2_err": err,
                 "ok": err < 1e-12}
            res["cases"].append(r)
            print(json.dumps(r), flush=True)
        del rp, ci, va, y, yr
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
