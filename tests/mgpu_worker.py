#!/usr/bin/env python
"""torchrun worker of tests/test_gpu_baseline_sizes.py::test_cg_mgpu_one_process_per_gpu_small_grids.

One process per GPU, bootstrapped like bench.py (CUDA-IPC handles all-gathered over
torch.distributed, then no NCCL call on the data path).  Many back-to-back solves on small grids:
an iteration lasts a few microseconds, so each rank's host lags the device by a different number of
iterations when convergence fires and enqueues a different number of no-op launches.  Every solve
must still return rc 0, the oracle's iteration count and solution, and device-side checksums that
describe the whole vector (the checksum exchange runs right after the loop)."""
import ctypes as C
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-spmv-benchmark_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    import torch
    import torch.distributed as dist
    import mgpu_bootstrap
    import orc
    import spmv_b200 as B

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)  # NCCL banner -> stderr
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = B.load()
    mgpu_bootstrap.connect(L, dist, rank, world, local, 2048)
    grids = [n for n in (64, 81, 130, 257, 512, 1000) if (n * n) // world >= n]
    fails = []
    for sched in (1, 0):
        L.b200_cg_set_schedule(sched)
        for n in grids:
            N = n * n
            nl, off = mgpu_bootstrap.partition(N, world, rank)
            hm = B.HostMatrix.synthetic_stencil(n)
            rp64, ci, va = orc.stencil5_csr_direct(n)
            b_full = np.ones(N)
            xo, ro, _ = orc.cg_device(rp64.astype(np.int32), ci, va, n, 1, b_full, np.zeros(N))
            for rep in range(6):
                b_loc, x_loc = np.ones(nl), np.zeros(nl)
                st = B.CGStatsMultiGPU()
                rc = L.cg_solve_mgpu_partitioned(None, hm.ptr(), b_loc.ctypes.data - off * 8, x_loc.ctypes.data - off * 8,
                                                 B.cg_config(), C.byref(st))
                ok = (rc == 0 and st.iterations == ro["iterations"] and st.converged == 1
                      and math.isclose(st.residual_norm, ro["residual_norm"], rel_tol=1e-10)
                      and np.linalg.norm(x_loc - xo[off:off + nl]) <= 1e-10 * np.linalg.norm(xo)
                      and math.isclose(st.solution_sum, ro["solution_sum"], rel_tol=1e-9)
                      and math.isclose(st.solution_norm, ro["solution_norm"], rel_tol=1e-9))
                if not ok:
                    fails.append("rank %d sched %d n %d rep %d rc %d it %d/%d sum %r/%r" % (
                        rank, sched, n, rep, rc, st.iterations, ro["iterations"], st.solution_sum, ro["solution_sum"]))
                    break
    L.b200_cg_set_schedule(1)
    t = torch.tensor([len(fails)], dtype=torch.int64, device="cuda")
    dist.all_reduce(t)
    for f in fails:
        print("FAIL", f, file=sys.stderr, flush=True)
    dist.barrier()
    L.b200_mgpu_finalize()
    dist.destroy_process_group()
    sys.stdout.flush()
    os.dup2(saved, 1)
    if rank == 0 and int(t[0]) == 0:
        print("MGPU_WORKER_OK world=%d grids=%s" % (world, grids), flush=True)
    return 1 if int(t[0]) else 0


if __name__ == "__main__":
    sys.exit(main())
