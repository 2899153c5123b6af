"""GPU parity gates at the BASELINE.json sizes (configs[1] 10k x 10k SpMV through the four operators,
configs[2] 20k x 20k CG) and on real multi-GPU hardware (configs[3]); the randomised stress cases of
tests/stress_parity.py with a bounded budget.

Oracle side: the OpenMP restatement in oracle/ (CSR product in the reference's k order, STENCIL5 in
the reference's W,C,E,N,S fma order); KATs: BASELINE.md section 2 (independent numpy/scipy statement
of the reference recurrence) plus the oracle's own 20k x 20k residual (bench.py --impl reference,
run in the build container: 14 iterations, ||r|| = 0.010842825287918112)."""
import ctypes as C
import math
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# BASELINE.md section 2 (11 significant digits printed there)
KAT = {
    10000: dict(iterations=14, rel_res=8.356e-7, solution_sum=9.9975281007e7, solution_norm=9.9978695581e3),
    20000: dict(iterations=14, rel_res=5.421e-7, solution_sum=3.9995055965e8, solution_norm=1.9997869532e4,
                residual_norm=0.010842825287918112),
}


def x_patterns(N):
    """the reference test helpers' vectors (tests/helpers/cuda_test_utils.cpp:89-146)"""
    inc = (np.arange(N, dtype=np.int64) % 97).astype(np.float64)
    rnd = np.random.default_rng(42).uniform(-1.0, 1.0, N)
    return {"incremental_mod_97": inc, "random_uniform_42": rnd}


@pytest.fixture(scope="module")
def oracle_10k(orc):
    n = 10000
    rp64, ci, va = orc.stencil5_csr_direct(n)
    return n, rp64.astype(np.int32), ci, va


@pytest.mark.parametrize("opname", [b"stencil5-csr", b"cusparse-csr", b"ellpack", b"stencil5-ellpack"])
def test_spmv_10k_operators_bit_exact(B, orc, torch_cuda, oracle_10k, opname):
    """configs[1]: 10k x 10k (1e8 rows) through get_operator(...)->run_timed with HOST vectors, against
    the oracle.  STENCIL5 operators: bit-exact with the reference's interior fma order; generic CSR /
    ELLPACK (5 entries per row => lane-per-row path): bit-exact with the sequential-k CSR product.
    x = 1: every y is an exact small integer, sum(y) = N + 4n."""
    L = B.load()
    n, rp, ci, va = oracle_10k
    N = n * n
    hm = B.HostMatrix.synthetic_stencil(n)
    op = L.get_operator(opname).contents
    assert op.init(hm.ptr()) == 0
    try:
        stencil = opname.startswith(b"stencil5")
        ms = C.c_double()
        y = np.empty(N)
        ones = np.ones(N)
        assert op.run_timed(ones.ctypes.data, y.ctypes.data, C.byref(ms)) == 0
        assert float(y.sum()) == N + 4 * n and y.min() == 1.0 and y.max() == 3.0
        assert 0 < ms.value < 50
        for name, x in x_patterns(N).items():
            y.fill(np.nan)
            assert op.run_timed(x.ctypes.data, y.ctypes.data, C.byref(ms)) == 0
            yo = orc.stencil5_spmv(rp, ci, va, x, n) if stencil else orc.csr_spmv(rp, ci, va, x)
            assert np.array_equal(y, yo), (opname, name, float(np.abs(y - yo).max()))
    finally:
        op.free()


@pytest.mark.parametrize("n", [10000, 20000])
def test_cg_baseline_sizes_kat(B, torch_cuda, n):
    """configs[2]: full CG (b = 1, x0 = 0, tol 1e-6) through cg_solve_device on the BASELINE grids:
    14 iterations, residual / checksums equal to the known answers, twice (second solve re-uses the
    workspace) with bit-identical results."""
    L = B.load()
    N = n * n
    hm = B.HostMatrix.synthetic_stencil(n)
    op = L.get_operator(b"stencil5-csr")
    assert op.contents.init(hm.ptr()) == 0
    try:
        b = np.ones(N)
        k = KAT[n]
        res = []
        for _ in range(2):
            x = np.zeros(N)
            st = B.CGStats()
            rc = L.cg_solve_device(op, hm.ptr(), b.ctypes.data, x.ctypes.data, B.cg_config(), C.byref(st))
            assert rc == 0
            assert st.iterations == k["iterations"] and st.converged == 1
            assert math.isclose(st.solution_sum, k["solution_sum"], rel_tol=1e-10)
            assert math.isclose(st.solution_norm, k["solution_norm"], rel_tol=1e-10)
            rel = st.residual_norm / math.sqrt(N)  # ||r0|| = ||b|| = sqrt(N)
            assert math.isclose(rel, k["rel_res"], rel_tol=2e-4)  # 4 printed digits
            if "residual_norm" in k:
                assert math.isclose(st.residual_norm, k["residual_norm"], rel_tol=1e-10)
            # the device-side checksums describe the x that came back
            assert math.isclose(st.solution_sum, float(x.sum()), rel_tol=1e-12)
            # (pairwise sum: BLAS nrm2 behind np.linalg.norm drifts by 1e-12 over 1e8 elements)
            assert math.isclose(st.solution_norm, math.sqrt(float(np.sum(x * x))), rel_tol=1e-12)
            res.append((st.residual_norm, st.solution_sum, st.solution_norm))
        assert res[0] == res[1]
    finally:
        op.contents.free()


# ------------------------------------------------------------------------------------------------
# real multi-GPU hardware (skipped on a one-GPU box; the driver's multi-GPU tier and gpurun --gpus N run them)
# ------------------------------------------------------------------------------------------------
def n_gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("n", [512, 2000])
def test_cg_mgpu_single_process_real_devices(B, orc, torch_cuda, n):
    """one process driving P distinct GPUs (the north-star topology): NVLink peer stores, system-scope
    flag waits and the LL exchange inside the producing kernels -- against the oracle (<= 1e-10) and
    against the one-GPU solve; the classic schedule must agree bit for bit with the deferred-x one."""
    P = min(n_gpus(), 8)
    if P < 2:
        pytest.skip("needs >= 2 GPUs")
    L = B.load()
    N = n * n
    hm = B.HostMatrix.synthetic_stencil(n)
    b = np.ones(N)
    rp64, ci, va = orc.stencil5_csr_direct(n)
    xo, ro, _ = orc.cg_device(rp64.astype(np.int32), ci, va, n, 1, b, np.zeros(N))
    op = L.get_operator(b"stencil5-csr")
    assert op.contents.init(hm.ptr()) == 0
    x1 = np.zeros(N)
    s1 = B.CGStats()
    assert L.cg_solve_device(op, hm.ptr(), b.ctypes.data, x1.ctypes.data, B.cg_config(), C.byref(s1)) == 0
    op.contents.free()
    worlds = sorted({2, P} | ({4} if P >= 4 else set()))
    for world in worlds:
        outs = []
        for sched in (1, 0):
            L.b200_cg_set_schedule(sched)
            devs = (C.c_int * world)(*range(world))
            assert L.b200_mgpu_init_single_process(world, devs, n) == 0
            try:
                for rep in range(3):  # repeated solves: the device-side sequence numbers keep advancing
                    x = np.zeros(N)
                    st = B.CGStatsMultiGPU()
                    rc = L.cg_solve_mgpu_partitioned(None, hm.ptr(), b.ctypes.data, x.ctypes.data, B.cg_config(), C.byref(st))
                    assert rc == 0, (world, sched, rep)
                    assert st.iterations == ro["iterations"] == s1.iterations and st.converged == 1
                    assert math.isclose(st.residual_norm, ro["residual_norm"], rel_tol=1e-10)
                    assert np.linalg.norm(x - xo) / np.linalg.norm(xo) < 1e-10
                    assert np.linalg.norm(x - x1) / np.linalg.norm(x1) < 1e-12
                    assert math.isclose(st.solution_sum, float(x.sum()), rel_tol=1e-12)
                    outs.append((x, st.residual_norm))
            finally:
                L.b200_mgpu_finalize()
                L.b200_cg_set_schedule(1)
        for x, r in outs[1:]:  # bit-reproducible run to run and across the two schedules
            assert r == outs[0][1] and np.array_equal(x, outs[0][0])


@pytest.mark.parametrize("world", [2, 8])
def test_cg_mgpu_one_process_per_gpu_small_grids(world):
    """one process per GPU under torchrun (what bench.py --gpus N runs), on grids so small that an
    iteration is shorter than a host wake-up: ranks enqueue different numbers of no-op iterations behind
    the convergence point.  The exchange sequence numbers live on the device, so every solve still
    agrees with the oracle; see tests/mgpu_worker.py."""
    if n_gpus() < world:
        pytest.skip("needs >= %d GPUs" % world)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", str(29600 + world),
                        os.path.join(ROOT, "tests", "mgpu_worker.py")],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MGPU_WORKER_OK" in r.stdout


# ------------------------------------------------------------------------------------------------
# randomised stress cases (tests/stress_parity.py) with a bounded budget
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", [0, 1])
def test_stress_parity_bounded(B, orc, torch_cuda, seed):
    import stress_parity as sp
    L = B.load()
    rng = np.random.default_rng(seed)
    cases = [sp.stencil_case, sp.cg_case, sp.csr_case]
    fails = []
    for i in range(45):
        msg = cases[i % 3](L, rng)
        if msg:
            fails.append(msg)
    assert not fails, fails
