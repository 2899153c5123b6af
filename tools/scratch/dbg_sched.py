import sys, os, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "cuda-spmv-benchmark_b200", "python"))
import torch
import spmv_b200 as B
L = B.load()
n, P = int(sys.argv[1]), int(sys.argv[2])
N = n * n
devs = (C.c_int * P)(*([0] * P))
hm = B.HostMatrix.synthetic_stencil(n)
rng = np.random.default_rng(n + P)
b = rng.standard_normal(N)
out = []
for sched in (0, 1):
    L.b200_cg_set_schedule(sched)
    assert L.b200_mgpu_init_single_process(P, devs, n) == 0
    x = np.full(N, 0.25)
    st = B.CGStatsMultiGPU()
    rc = L.cg_solve_mgpu_partitioned(None, hm.ptr(), b.ctypes.data, x.ctypes.data, B.cg_config(max_iters=int(sys.argv[3]) if len(sys.argv) > 3 else 1000), C.byref(st))
    out.append((x, st.iterations, st.residual_norm))
    L.b200_mgpu_finalize()
d = np.abs(out[0][0] - out[1][0])
idx = np.nonzero(d)[0]
print("iters", out[0][1], out[1][1], "res", out[0][2], out[1][2], "ndiff", len(idx), "max", d.max())
q = N // P
for i in idx[:40]:
    print(i, "grid", divmod(int(i), n), "rank", min(int(i) // q, P - 1), "local", int(i) - min(int(i) // q, P - 1) * q, out[0][0][i], out[1][0][i])
