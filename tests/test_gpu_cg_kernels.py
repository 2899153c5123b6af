"""GPU parity of BOTH kernel families of the fused CG passes -- sequential sweep (csrc/stencil5_direct.cuh,
the default, `b200_cg_set_kernel(1)` / B200_CG_KERNEL=sweep) and bulk-copy ring (csrc/stencil5.cuh,
`b200_cg_set_kernel(0)` / B200_CG_KERNEL=ring) -- and of every x retirement depth: same bar as
tests/test_gpu_cg.py (iteration count of the oracle, x within 1e-10), and the CG schedules stay bit-identical
to each other, on one GPU and over virtual ranks (halo tiles wait for the neighbour's arrival word)."""
import ctypes as C

import numpy as np
import pytest

from test_gpu_cg import oracle_solve, solve_device

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[0, 1], ids=["ring", "sweep"])
def sweep_kernel(B, request):
    L = B.load()
    family0 = L.b200_cg_get_kernel()
    L.b200_cg_set_kernel(request.param)
    yield L
    L.b200_cg_set_kernel(family0)
    L.b200_cg_set_schedule(1)


@pytest.mark.parametrize("n", [3, 5, 33, 64, 81, 130, 257, 700, 1031])
def test_cg_sweep_kernel_matches_oracle_and_schedules_agree(B, orc, torch_cuda, sweep_kernel, n):
    L = sweep_kernel
    N = n * n
    hm = B.HostMatrix.synthetic_stencil(n)
    rng = np.random.default_rng(n)
    b, x0 = rng.standard_normal(N), rng.standard_normal(N)
    xo, ro, _ = oracle_solve(orc, n, 1, b, x0)
    out = []
    for sched in (0, 1):
        L.b200_cg_set_schedule(sched)
        x, st, op = solve_device(B, b"stencil5-csr", hm, b, x0)
        op.contents.free()
        assert st["iterations"] == ro["iterations"] and st["converged"] == 1
        assert np.linalg.norm(x - xo) / np.linalg.norm(xo) < 1e-10
        out.append((x, st))
    assert np.array_equal(out[0][0], out[1][0])
    assert out[0][1]["residual_norm"] == out[1][1]["residual_norm"]


@pytest.mark.parametrize("n,P", [(81, 2), (64, 8), (130, 3), (512, 4), (1031, 3)])
def test_cg_sweep_kernel_virtual_ranks(B, orc, torch_cuda, sweep_kernel, n, P):
    L = sweep_kernel
    N = n * n
    devs = (C.c_int * P)(*([0] * P))
    hm = B.HostMatrix.synthetic_stencil(n)
    rng = np.random.default_rng(n + P)
    b = rng.standard_normal(N)
    xo, ro, _ = oracle_solve(orc, n, 1, b, np.full(N, 0.25))
    out = []
    try:
        for sched in (0, 1):
            L.b200_cg_set_schedule(sched)
            assert L.b200_mgpu_init_single_process(P, devs, n) == 0
            x = np.full(N, 0.25)
            st = B.CGStatsMultiGPU()
            rc = L.cg_solve_mgpu_partitioned(None, hm.ptr(), b.ctypes.data, x.ctypes.data, B.cg_config(), C.byref(st))
            assert rc == 0, L.b200_last_error()
            L.b200_mgpu_finalize()
            assert st.iterations == ro["iterations"] and st.converged == 1
            assert np.linalg.norm(x - xo) / np.linalg.norm(xo) < 1e-10
            out.append((x, st.residual_norm))
    finally:
        L.b200_mgpu_finalize()
    assert np.array_equal(out[0][0], out[1][0]) and out[0][1] == out[1][1]


# ---- x retirement depth (b200_cg_set_xdepth): the SpMV launch of every depth-th iteration retires the last
# `depth` x updates at once (oldest first) -- every iterate must stay bit-identical to the classic schedule,
# whatever iteration the solve stops at (1 .. depth updates are then pending for cg_finish_x)
@pytest.fixture
def restore_cg_defaults(B):
    L = B.load()
    depth0 = L.b200_cg_set_xdepth(0)  # 0 = query only
    family0 = L.b200_cg_get_kernel()
    yield L
    L.b200_cg_set_kernel(family0)
    L.b200_cg_set_schedule(1)
    L.b200_cg_set_xdepth(depth0)


@pytest.mark.parametrize("n", [3, 64, 130, 700])
@pytest.mark.parametrize("kernel", [0, 1])
def test_cg_xdepth_bit_identical_to_classic(B, torch_cuda, restore_cg_defaults, n, kernel):
    L = restore_cg_defaults
    N = n * n
    hm = B.HostMatrix.synthetic_stencil(n)
    rng = np.random.default_rng(n)
    b, x0 = rng.standard_normal(N), rng.standard_normal(N)
    L.b200_cg_set_kernel(kernel)
    for tol, max_iters in [(1e-6, m) for m in (1000, 1, 2, 3, 4, 5, 6, 7, 8, 9)] + [(1e-4, 1000), (1e-2, 1000), (1e-1, 1000)]:
        L.b200_cg_set_schedule(0)
        xc, sc, op = solve_device(B, b"stencil5-csr", hm, b, x0, tol=tol, max_iters=max_iters)
        op.contents.free()
        L.b200_cg_set_schedule(1)
        for depth in (1, 2, 3, 4):
            L.b200_cg_set_xdepth(depth)
            xd, sd, op = solve_device(B, b"stencil5-csr", hm, b, x0, tol=tol, max_iters=max_iters)
            op.contents.free()
            assert sd["iterations"] == sc["iterations"] and sd["converged"] == sc["converged"], (tol, max_iters, depth)
            assert sd["residual_norm"] == sc["residual_norm"], (tol, max_iters, depth)
            assert np.array_equal(xc, xd), (tol, max_iters, depth)


@pytest.mark.parametrize("n,P", [(81, 2), (64, 8), (130, 3), (512, 4)])
@pytest.mark.parametrize("kernel", [0, 1])
def test_cg_xdepth_virtual_ranks_bit_identical_to_classic(B, torch_cuda, restore_cg_defaults, n, P, kernel):
    L = restore_cg_defaults
    N = n * n
    devs = (C.c_int * P)(*([0] * P))
    hm = B.HostMatrix.synthetic_stencil(n)
    rng = np.random.default_rng(n + P)
    b = rng.standard_normal(N)
    L.b200_cg_set_kernel(kernel)

    def solve(max_iters):
        assert L.b200_mgpu_init_single_process(P, devs, n) == 0
        x = np.full(N, 0.25)
        st = B.CGStatsMultiGPU()
        rc = L.cg_solve_mgpu_partitioned(None, hm.ptr(), b.ctypes.data, x.ctypes.data, B.cg_config(max_iters), C.byref(st))
        L.b200_mgpu_finalize()
        assert rc == 0, L.b200_last_error()
        return x, st.iterations, st.residual_norm

    try:
        for max_iters in (1000, 3, 4, 6):
            L.b200_cg_set_schedule(0)
            ref = solve(max_iters)
            L.b200_cg_set_schedule(1)
            for depth in (1, 2, 3, 4):
                L.b200_cg_set_xdepth(depth)
                got = solve(max_iters)
                assert got[1] == ref[1] and got[2] == ref[2], (max_iters, depth)
                assert np.array_equal(got[0], ref[0]), (max_iters, depth)
    finally:
        L.b200_mgpu_finalize()


@pytest.mark.parametrize("opname", [b"cusparse-csr", b"ellpack"])
@pytest.mark.parametrize("n", [3, 81, 300])
def test_cg_k3x_depth_bit_identical_to_classic_generic_operators(B, torch_cuda, restore_cg_defaults, n, opname):
    """operators without a fused SpMV: the p update (K3x) writes a fresh direction buffer and retires the x updates
    of the last `depth` iterations every depth-th launch; solves stopped by max_iters after 1..9 iterations leave
    0..depth-1 updates pending (the K3x launch of the last iteration has run), converged ones 1..depth."""
    L = restore_cg_defaults
    N = n * n
    hm = B.HostMatrix.synthetic_stencil(n)
    rng = np.random.default_rng(n)
    b, x0 = rng.standard_normal(N), rng.standard_normal(N)
    for tol in (1e-6, 1e-3, 1e-1):  # converge after different iteration counts
        for max_iters in (1000, 1, 2, 3, 4, 5, 6, 7, 8, 9):
            if tol != 1e-6 and max_iters != 1000:
                continue
            L.b200_cg_set_schedule(0)
            xc, sc, op = solve_device(B, opname, hm, b, x0, tol=tol, max_iters=max_iters)
            op.contents.free()
            L.b200_cg_set_schedule(1)
            for depth in (1, 2, 3, 4):
                L.b200_cg_set_xdepth(depth)
                xd, sd, op = solve_device(B, opname, hm, b, x0, tol=tol, max_iters=max_iters)
                op.contents.free()
                assert sd["iterations"] == sc["iterations"] and sd["converged"] == sc["converged"], (tol, max_iters, depth)
                assert sd["residual_norm"] == sc["residual_norm"], (tol, max_iters, depth)
                assert np.array_equal(xc, xd), (tol, max_iters, depth)
