#!/usr/bin/env python
"""Kernel micro-benchmarks on one GPU through the C ABI (development tool, not the bench).

  python tools/sweep.py --grid 10000 [--what stencil,cg,csr,ell] [--out gpurun_out/sweep.json]

Every kernel is timed with CUDA events on the launching stream after warm-up; working sets are
far larger than the 126 MB L2.  Achieved GB/s uses ALGORITHMIC bytes:
STENCIL5 8*nnz + 16*N (fused K1F: 8*nnz + 48*N), CSR 12*nnz + 4(N+1) + 16*N, ELLPACK 76*N, K2 48*N, K3 24*N.
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-spmv-benchmark_b200", "python"))
import torch  # noqa: E402

import spmv_b200 as B  # noqa: E402


def dptr(t):
    return C.c_void_p(t.data_ptr())


def timeit(fn, reps=7, warm=3, batch=4):
    """median / best time of one launch; `batch` back-to-back launches per event pair so that the
    launch latency of an idle GPU is not charged to the kernel"""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(batch):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / batch)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=10000)
    ap.add_argument("--what", default="stencil,cg,csr,ell")
    ap.add_argument("--variants", default="0,3,9,12,13,20,21,22")
    ap.add_argument("--rows", default="8,16,32,64,128")
    ap.add_argument("--csr-variants", default="0")
    ap.add_argument("--out", default="gpurun_out/sweep.json")
    a = ap.parse_args()
    what = a.what.split(",")
    L = B.load()
    torch.cuda.set_device(0)
    s = torch.cuda.current_stream().cuda_stream
    n = a.grid
    N = n * n
    nnz = 5 * N - 4 * n
    peak = 6551.0
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    res = {"grid": n, "peak_gbs": peak, "results": []}

    def rec(name, ms, best, nbytes, **kw):
        gbs = nbytes / (ms * 1e-3) / 1e9
        r = dict(kernel=name, ms=round(ms, 4), best_ms=round(best, 4), gbs=round(gbs, 1), frac=round(gbs / peak, 3), **kw)
        res["results"].append(r)
        print(json.dumps(r), flush=True)

    rp = torch.empty(N + 1, dtype=torch.int32, device="cuda")
    ci = torch.empty(nnz + 2, dtype=torch.int32, device="cuda")
    va = torch.zeros(nnz + 2, dtype=torch.float64, device="cuda")
    B.check(L.b200_gen_stencil5_csr(n, 0, N, 5.0, -1.0, dptr(rp), dptr(ci), dptr(va), s), "gen")
    x = torch.ones(N, dtype=torch.float64, device="cuda")
    y = torch.empty(N, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()

    # reference point: plain device copy of the same byte volume class (what MEASURED_PEAKS measures)
    big = torch.empty(1 << 29, dtype=torch.float64, device="cuda")  # 4 GiB
    big2 = torch.empty_like(big)
    ms, best = timeit(lambda: big2.copy_(big))
    rec("torch_copy_4GiB", ms, best, 2.0 * big.numel() * 8)
    del big, big2

    def reduce_ctx(capacity, phases=3):
        """b200_reduce_ctx over fresh device buffers (fused reduction tail inside the producing kernel)"""
        groups = (capacity + 255) // 256
        t = {"sc": torch.zeros(64, dtype=torch.float64, device="cuda"),
             "pa": torch.empty(capacity, dtype=torch.float64, device="cuda"),
             "pb": torch.empty(capacity, dtype=torch.float64, device="cuda"),
             "gs": torch.empty(2 * groups, dtype=torch.float64, device="cuda"),
             "tk": torch.zeros(groups + 1, dtype=torch.int32, device="cuda"),
             "st": torch.zeros(4, dtype=torch.float64, device="cuda")}
        t["sc"][0] = 1.0   # rr_old
        t["sc"][3] = 0.5   # alpha
        t["sc"][4] = 0.25  # beta
        t["sc"][5] = 1.0   # b_norm
        c = B.ReduceCtx(t["sc"].data_ptr(), None, t["pa"].data_ptr(), t["pb"].data_ptr(), t["gs"].data_ptr(),
                        t["tk"].data_ptr(), capacity, t["st"].data_ptr(), t["st"].data_ptr() + 16, 0, 1, None, 0.0, phases)
        c._keep = t
        return c

    if "stencil" in what:
        st_bytes = 8.0 * nnz + 16.0 * N
        ctx = reduce_ctx(1 << 20)
        if "fused" in what:
            fr, fp, fx = (torch.ones(N, dtype=torch.float64, device="cuda") for _ in range(3))
        for v in [int(t) for t in a.variants.split(",")]:
            info = L.b200_stencil5_variant_info(v)
            if info is None:
                continue
            for R in [int(t) for t in a.rows.split(",")]:
                band = B.Band(rp.data_ptr(), ci.data_ptr(), va.data_ptr(), nnz + 2, 0, N, n, 0, None, None, None,
                              None, 0, R, v)
                ms, best = timeit(lambda: B.check(L.b200_stencil5_spmv(C.byref(band), dptr(x), dptr(y), s), "st"))
                ok = float(y.sum().item()) == N + 4 * n
                rec("stencil5_plain", ms, best, st_bytes, variant=v, rows_per_item=R, ok=ok)
                if v >= 20:  # sequential-sweep kernels: plain product only, no rows_per_item
                    break
                ms, best = timeit(lambda: B.check(L.b200_cg_spmv_dot(C.byref(band), dptr(x), dptr(y), C.byref(ctx), s), "dot"))
                rec("stencil5_dot+tail", ms, best, st_bytes, variant=v, rows_per_item=R)
                if "fused" in what:
                    ms, best = timeit(lambda: B.check(L.b200_cg_spmv_fused(C.byref(band), dptr(x), dptr(fr), dptr(fp),
                                                                           dptr(fx), dptr(y), C.byref(ctx), s), "fused"))
                    rec("stencil5_fused(K1F)+tail", ms, best, 8.0 * nnz + 48.0 * N, variant=v, rows_per_item=R)
        if "cgsweep" in what:  # fused CG passes, ring against sequential sweep, by number of x updates retired per launch
            band = B.Band(rp.data_ptr(), ci.data_ptr(), va.data_ptr(), nnz + 2, 0, N, n, 0, None, None, None,
                          None, 0, 4, 0)
            fr, fp, fx = (torch.ones(N, dtype=torch.float64, device="cuda") for _ in range(3))
            older_t = [torch.ones(N, dtype=torch.float64, device="cuda") for _ in range(3)]
            older = (C.c_void_p * 3)(*[t.data_ptr() for t in older_t])
            for fam, name in ((0, "ring"), (1, "sweep")):
                L.b200_cg_set_kernel(fam)
                ms, best = timeit(lambda: B.check(L.b200_cg_spmv_dot(C.byref(band), dptr(x), dptr(y), C.byref(ctx), s), "dot"))
                rec(name + "_dot+tail", ms, best, st_bytes)
                for nx in (0, 1, 2, 3, 4):
                    if fam == 0 and nx != 1:
                        continue  # the ring family only exists with one x update per launch (the library would run the sweep kernel)
                    ms, best = timeit(lambda: B.check(L.b200_cg_spmv_fused_nx(C.byref(band), dptr(x), older, nx, dptr(fr), dptr(fp),
                                                                              dptr(fx), dptr(y), C.byref(ctx), s), "fused"))
                    nb = 8.0 * nnz + 32.0 * N + (16.0 * N + 8.0 * N * (nx - 1) if nx > 0 else 0.0)
                    rec("%s_fused_x%d(K1F)+tail" % (name, nx), ms, best, nb)
            L.b200_cg_set_kernel(0)
        del ctx

    if "cg" in what:
        ctx = reduce_ctx(1 << 16)
        sc = ctx._keep["sc"]
        p, Ap, xx, r = (torch.ones(N, dtype=torch.float64, device="cuda") for _ in range(4))

        def keep_going():  # the timed kernels are no-ops once `converged` is set
            sc[7] = 0.0  # converged / iterations
        keep_going()
        ms, best = timeit(lambda: B.check(L.b200_cg_update_xr(N, dptr(p), dptr(Ap), dptr(xx), dptr(r), C.byref(ctx), s), "k2"))
        rec("cg_update_xr(K2)+tail", ms, best, 48.0 * N)
        keep_going()
        ms, best = timeit(lambda: B.check(L.b200_cg_update_p(N, dptr(sc), dptr(r), dptr(p), s), "k3"))
        rec("cg_update_p(K3)", ms, best, 24.0 * N)
        keep_going()
        ms, best = timeit(lambda: B.check(L.b200_cg_update_r(N, dptr(Ap), dptr(r), None, C.byref(ctx), s), "k2r"))
        rec("cg_update_r(K2r)+tail", ms, best, 24.0 * N)
        ms, best = timeit(lambda: B.check(L.b200_cg_reduce(C.byref(ctx), 3, 1184, 0, 3, s), "red"))
        rec("cg_reduce(1184 partials, stand-alone)", ms, best, 1184 * 8.0)
        del p, Ap, xx, r, ctx

    cvars = [int(t) for t in a.csr_variants.split(",")]
    if "csr" in what:
        plan = B.CsrPlan()
        B.check(L.b200_csr_plan_build(dptr(rp), N, nnz, C.byref(plan), s), "plan")
        for cv in cvars:
            plan.variant = cv
            y.fill_(float("nan"))
            ms, best = timeit(lambda: B.check(L.b200_spmv_csr(C.byref(plan), dptr(rp), dptr(ci), dptr(va), dptr(x),
                                                              dptr(y), N, 1.0, 0.0, s), "csr"))
            ok = float(y.sum().item()) == N + 4 * n
            rec("csr_generic", ms, best, 12.0 * nnz + 4.0 * (N + 1) + 16.0 * N, variant=cv, ok=ok)

    if "ell" in what:
        del ci, va, rp
        idx = torch.empty(5 * N + 2, dtype=torch.int32, device="cuda")
        val = torch.empty(5 * N + 2, dtype=torch.float64, device="cuda")
        B.check(L.b200_gen_stencil5_ellpack(n, 0, N, 5.0, -1.0, dptr(idx), dptr(val), s), "gen ell")
        for cv in cvars:
            L.b200_csr_set_default_variant(cv)
            y.fill_(float("nan"))
            ms, best = timeit(lambda: B.check(L.b200_spmv_ellpack(dptr(idx), dptr(val), dptr(x), dptr(y), N, 5, 1.0, 0.0, s), "ell"))
            ok = float(y.sum().item()) == N + 4 * n
            rec("ellpack_generic", ms, best, 76.0 * N, variant=cv, ok=ok)
        L.b200_csr_set_default_variant(0)
        ms, best = timeit(lambda: B.check(L.b200_spmv_stencil5_ellpack(dptr(val), dptr(idx), dptr(x), dptr(y), N, 5, 1.0,
                                                                       0.0, n, s), "st-ell"))
        ok = float(y.sum().item()) == N + 4 * n
        rec("stencil5_ellpack", ms, best, 56.0 * N, ok=ok)

    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
