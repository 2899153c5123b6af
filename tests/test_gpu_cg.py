"""GPU parity: CG through the reference-facing entry points (cg_solve_device, cg_solve,
cg_solve_mgpu_partitioned, the bench wrappers) vs the CPU oracle.

Bar (BASELINE.json north_star): same iteration count as the reference recurrence; final residual
within 1e-10 relative of the oracle's; solution within 1e-10 relative L2 (the dot products are
summed in a different -- fixed -- order than the reference's block tree, so the trajectories agree
to rounding, not bit for bit); results reproducible bit for bit from run to run."""
import ctypes as C
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def solve_device(B, opname, hm, b, x0, tol=1e-6, max_iters=1000, timers=0, entry="cg_solve_device"):
    L = B.load()
    op = L.get_operator(opname)
    assert op.contents.init(hm.ptr()) == 0
    x = np.array(x0, dtype=np.float64, copy=True)
    st = B.CGStats()
    rc = getattr(L, entry)(op, hm.ptr(), b.ctypes.data, x.ctypes.data, B.cg_config(max_iters, tol, 0, timers), C.byref(st))
    assert rc == 0
    return x, B.stats_dict(st), op


def oracle_solve(orc, n, op, b, x0, center=5.0, tol=1e-6, max_iters=1000):
    rp64, ci, va = orc.stencil5_csr_direct(n, center, -1.0)
    return orc.cg_device(rp64.astype(np.int32), ci, va, n, op, b, x0, max_iters, tol)


@pytest.mark.parametrize("n,iters", [(3, 3), (81, 18), (512, 17), (1000, None)])
@pytest.mark.parametrize("opname", [b"stencil5-csr", b"cusparse-csr", b"ellpack", b"stencil5-ellpack"])
def test_cg_solve_device_matches_oracle(B, orc, torch_cuda, n, iters, opname):
    N = n * n
    hm = B.HostMatrix.synthetic_stencil(n)
    b, x0 = np.ones(N), np.zeros(N)
    x, st, op = solve_device(B, opname, hm, b, x0)
    xo, ro, _ = oracle_solve(orc, n, 1 if opname.startswith(b"stencil5") else 0, b, x0)
    if iters is not None:
        assert ro["iterations"] == iters
    assert st["iterations"] == ro["iterations"] and st["converged"] == 1 == ro["converged"]
    # "final residual within 1e-10 relative": relative to the initial residual (a residual that has
    # dropped to rounding level, as on the 3x3 grid, has no meaningful digits of its own)
    assert abs(st["residual_norm"] - ro["residual_norm"]) <= 1e-10 * max(ro["residual_norm"], ro["b_norm"] * 1e-6)
    assert np.linalg.norm(x - xo) / np.linalg.norm(xo) < 1e-10
    # x agrees to 1e-10 relative, so its sum / norm agree to the same order; the device-side
    # checksum itself must match a host sum of the returned x to rounding
    assert math.isclose(st["solution_sum"], ro["solution_sum"], rel_tol=1e-9)
    assert math.isclose(st["solution_norm"], ro["solution_norm"], rel_tol=1e-9)
    assert math.isclose(st["solution_sum"], float(x.sum()), rel_tol=1e-12)
    assert math.isclose(st["solution_norm"], float(np.linalg.norm(x)), rel_tol=1e-12)
    assert st["time_total_ms"] > 0
    op.contents.free()


def test_cg_from_mtx_file_bundled(B, orc, torch_cuda, tmp_path):
    """config[0]: bundled 81x81 (centre -4) loaded from the .mtx, 40 iterations"""
    p = str(tmp_path / "example81x81.mtx")
    orc.write_mtx_stencil5(81, p, "-4.0", "-1.0")
    hm = B.HostMatrix.from_mtx(p)
    b, x0 = np.ones(6561), np.zeros(6561)
    for opname in (b"stencil5-csr", b"cusparse-csr"):
        x, st, op = solve_device(B, opname, hm, b, x0)
        xo, ro, _ = oracle_solve(orc, 81, 1, b, x0, center=-4.0)
        assert st["iterations"] == 40 == ro["iterations"] and st["converged"] == 1
        assert math.isclose(st["residual_norm"], ro["residual_norm"], rel_tol=1e-10)
        assert math.isclose(st["solution_sum"], -826.0838884, rel_tol=1e-9)
        assert np.linalg.norm(x - xo) / np.linalg.norm(xo) < 1e-10
        op.contents.free()


def test_cg_nonzero_guess_random_rhs_and_host_entry(B, orc, torch_cuda):
    n = 200
    N = n * n
    rng = np.random.default_rng(42)
    b, x0 = rng.standard_normal(N), rng.standard_normal(N)
    hm = B.HostMatrix.synthetic_stencil(n)
    x, st, op = solve_device(B, b"stencil5-csr", hm, b, x0, tol=1e-9, entry="cg_solve")
    xo, ro, _ = oracle_solve(orc, n, 1, b, x0, tol=1e-9)
    assert st["iterations"] == ro["iterations"] and st["converged"] == 1
    assert np.linalg.norm(x - xo) / np.linalg.norm(xo) < 1e-10
    op.contents.free()


def test_cg_max_iters_and_timers(B, orc, torch_cuda):
    n = 300
    N = n * n
    hm = B.HostMatrix.synthetic_stencil(n)
    b, x0 = np.ones(N), np.zeros(N)
    x, st, op = solve_device(B, b"stencil5-csr", hm, b, x0, tol=1e-14, max_iters=5, timers=1)
    xo, ro, hist = oracle_solve(orc, n, 1, b, x0, tol=1e-14, max_iters=5)
    assert st["iterations"] == 5 == ro["iterations"] and st["converged"] == 0 == ro["converged"]
    assert math.isclose(st["residual_norm"], ro["residual_norm"], rel_tol=1e-10)
    assert np.linalg.norm(x - xo) / np.linalg.norm(xo) < 1e-12
    assert st["time_spmv_ms"] > 0 and st["time_blas1_ms"] > 0 and st["time_reductions_ms"] > 0
    assert st["time_spmv_ms"] + st["time_blas1_ms"] + st["time_reductions_ms"] <= st["time_total_ms"] * 1.05
    op.contents.free()


def test_cg_bitwise_reproducible(B, torch_cuda):
    n = 700
    N = n * n
    hm = B.HostMatrix.synthetic_stencil(n)
    b, x0 = np.ones(N), np.zeros(N)
    x1, st1, op = solve_device(B, b"stencil5-csr", hm, b, x0)
    x2, st2, op = solve_device(B, b"stencil5-csr", hm, b, x0)
    assert np.array_equal(x1, x2) and st1["residual_norm"] == st2["residual_norm"]
    assert st1["iterations"] == st2["iterations"]
    op.contents.free()


def test_cg_bench_wrapper(B, torch_cuda):
    L = B.load()
    n = 256
    N = n * n
    hm = B.HostMatrix.synthetic_stencil(n)
    op = L.get_operator(b"stencil5-csr")
    assert op.contents.init(hm.ptr()) == 0
    b, x = np.ones(N), np.zeros(N)
    bs, cs = B.BenchmarkStats(), B.CGStats()
    rc = L.cg_benchmark_with_stats_device(op, hm.ptr(), b.ctypes.data, x.ctypes.data, B.cg_config(), 5, C.byref(bs), C.byref(cs))
    assert rc == 0 and 3 <= bs.valid_runs <= 5 and bs.min_ms <= bs.median_ms <= bs.max_ms
    assert cs.converged == 1 and cs.iterations > 0
    op.contents.free()


@pytest.mark.parametrize("n,P", [(81, 2), (81, 4), (64, 8), (130, 3), (512, 4)])
def test_cg_mgpu_partitioned_virtual_ranks(B, orc, torch_cuda, n, P):
    """P row bands as P virtual ranks on one GPU: exercises partition, local CSR slices, peer halo
    push, flag waits and the LL scalar exchange of the multi-GPU path without needing P GPUs."""
    L = B.load()
    N = n * n
    devs = (C.c_int * P)(*([0] * P))
    assert L.b200_mgpu_init_single_process(P, devs, n) == 0
    try:
        assert L.b200_mgpu_world() == P
        hm = B.HostMatrix.synthetic_stencil(n)
        b, x = np.ones(N), np.zeros(N)
        st = B.CGStatsMultiGPU()
        rc = L.cg_solve_mgpu_partitioned(None, hm.ptr(), b.ctypes.data, x.ctypes.data, B.cg_config(timers=1), C.byref(st))
        assert rc == 0
        xo, ro, _ = oracle_solve(orc, n, 1, b, np.zeros(N))
        rp64, ci, va = orc.stencil5_csr_direct(n)
        xm, rm = orc.cg_mgpu(rp64.astype(np.int32), ci, va, n, P, b, np.zeros(N))
        assert st.iterations == ro["iterations"] == rm["iterations"] and st.converged == 1
        assert math.isclose(st.residual_norm, ro["residual_norm"], rel_tol=1e-10)
        assert np.linalg.norm(x - xo) / np.linalg.norm(xo) < 1e-10
        assert math.isclose(st.solution_sum, float(x.sum()), rel_tol=1e-12)
        assert st.time_total_ms > 0 and st.time_allgather_ms > 0
        # a second solve re-uses the workspace and the epochs keep advancing
        x2 = np.zeros(N)
        st2 = B.CGStatsMultiGPU()
        assert L.cg_solve_mgpu_partitioned(None, hm.ptr(), b.ctypes.data, x2.ctypes.data, B.cg_config(), C.byref(st2)) == 0
        assert np.array_equal(x, x2) and st2.iterations == st.iterations
    finally:
        L.b200_mgpu_finalize()


def test_cg_mgpu_from_mtx_entries(B, orc, torch_cuda, tmp_path):
    """band slices cut from the host CSR (not generated): bundled matrix on 3 virtual ranks"""
    L = B.load()
    p = str(tmp_path / "example81x81.mtx")
    orc.write_mtx_stencil5(81, p, "-4.0", "-1.0")
    hm = B.HostMatrix.from_mtx(p)
    devs = (C.c_int * 3)(0, 0, 0)
    assert L.b200_mgpu_init_single_process(3, devs, 81) == 0
    try:
        b, x = np.ones(6561), np.zeros(6561)
        st = B.CGStatsMultiGPU()
        assert L.cg_solve_mgpu_partitioned(None, hm.ptr(), b.ctypes.data, x.ctypes.data, B.cg_config(), C.byref(st)) == 0
        assert st.iterations == 40 and st.converged == 1
        assert math.isclose(st.solution_sum, -826.0838884, rel_tol=1e-9)
    finally:
        L.b200_mgpu_finalize()


def test_halo_mgpu_operator(B, orc, torch_cuda):
    """"stencil5-halo-mgpu" (declared-only in the reference): host vectors in/out over row bands"""
    import os
    L = B.load()
    n = 96
    N = n * n
    os.environ["B200_GPUS"] = "3"
    try:
        hm = B.HostMatrix.synthetic_stencil(n)
        op = L.get_operator(b"stencil5-halo-mgpu").contents
        assert not op.run_device
        assert op.init(hm.ptr()) == 0
        rng = np.random.default_rng(9)
        x = rng.standard_normal(N)
        y = np.full(N, np.nan)
        ms = C.c_double()
        assert op.run_timed(x.ctypes.data, y.ctypes.data, C.byref(ms)) == 0
        rp64, ci, va = orc.stencil5_csr_direct(n)
        assert np.array_equal(y, orc.stencil5_spmv(rp64.astype(np.int32), ci, va, x, n))
        op.free()
    finally:
        del os.environ["B200_GPUS"]


@pytest.mark.parametrize("opname", [b"cusparse-csr", b"ellpack"])
@pytest.mark.parametrize("n", [3, 81, 300])
@pytest.mark.parametrize("max_iters", [1000, 1, 4])
def test_cg_k3x_schedule_bit_identical_to_classic_generic_operators(B, torch_cuda, n, max_iters, opname):
    """operators without a fused SpMV: K1 / K2r / K3x (x retired inside the p update) must reproduce the
    classic K1 / K2 / K3 grouping bit for bit, also when max_iters cuts the solve (no pending x update then)
    and when the convergence test does (one pending update, applied by cg_finish_x)"""
    L = B.load()
    N = n * n
    hm = B.HostMatrix.synthetic_stencil(n)
    rng = np.random.default_rng(n)
    b, x0 = rng.standard_normal(N), rng.standard_normal(N)
    out = []
    try:
        for sched in (0, 1):
            L.b200_cg_set_schedule(sched)
            x, st, op = solve_device(B, opname, hm, b, x0, max_iters=max_iters)
            out.append((x, st))
            op.contents.free()
    finally:
        L.b200_cg_set_schedule(1)
    (xc, sc), (xd, sd) = out
    assert sc["iterations"] == sd["iterations"] and sc["converged"] == sd["converged"]
    assert sc["residual_norm"] == sd["residual_norm"]
    assert np.array_equal(xc, xd)


@pytest.mark.parametrize("n", [3, 64, 81, 130, 700])
@pytest.mark.parametrize("max_iters", [1000, 1, 2, 5])
def test_cg_deferred_x_schedule_bit_identical_to_classic(B, torch_cuda, n, max_iters):
    """The 4-launch deferred-x schedule (p formed inside the SpMV, x retired one iteration late,
    r.r summed in K2's order) must reproduce the classic 5-launch schedule bit for bit: iterates,
    residual, iteration count -- also when the solve is cut off by max_iters (pending x update)."""
    L = B.load()
    N = n * n
    hm = B.HostMatrix.synthetic_stencil(n)
    rng = np.random.default_rng(n)
    b, x0 = rng.standard_normal(N), rng.standard_normal(N)
    out = []
    try:
        for sched in (0, 1):
            L.b200_cg_set_schedule(sched)
            x, st, op = solve_device(B, b"stencil5-csr", hm, b, x0, max_iters=max_iters)
            out.append((x, st))
            op.contents.free()
    finally:
        L.b200_cg_set_schedule(1)
    (xc, sc), (xd, sd) = out
    assert sc["iterations"] == sd["iterations"] and sc["converged"] == sd["converged"]
    assert sc["residual_norm"] == sd["residual_norm"]
    assert np.array_equal(xc, xd)


@pytest.mark.parametrize("n,P", [(81, 2), (64, 8), (130, 3), (512, 4)])
def test_cg_mgpu_deferred_x_schedule_bit_identical_to_classic(B, torch_cuda, n, P):
    """same, over P virtual ranks: r-edge push fused into K2r, halo copies of the direction kept by
    cg_halo_dir (p_halo = r_halo + beta p_halo_old)"""
    L = B.load()
    N = n * n
    devs = (C.c_int * P)(*([0] * P))
    hm = B.HostMatrix.synthetic_stencil(n)
    rng = np.random.default_rng(n + P)
    b = rng.standard_normal(N)
    out = []
    try:
        for sched in (0, 1):
            L.b200_cg_set_schedule(sched)
            assert L.b200_mgpu_init_single_process(P, devs, n) == 0
            x = rng.standard_normal(N) if False else np.full(N, 0.25)
            st = B.CGStatsMultiGPU()
            rc = L.cg_solve_mgpu_partitioned(None, hm.ptr(), b.ctypes.data, x.ctypes.data, B.cg_config(timers=1), C.byref(st))
            assert rc == 0
            out.append((x, st.iterations, st.residual_norm, st.converged))
            L.b200_mgpu_finalize()
    finally:
        L.b200_cg_set_schedule(1)
        L.b200_mgpu_finalize()
    (xc, ic, rc_, cc), (xd, id_, rd, cd) = out
    assert ic == id_ and cc == cd == 1 and rc_ == rd
    assert np.array_equal(xc, xd)


def variable_diagonal_stencil(orc, n, seed):
    """5-point stencil entries (generator order) with a strongly varying diagonal: SPD (diagonally
    dominant), and a case where Jacobi preconditioning actually pays"""
    rng = np.random.default_rng(seed)
    ent = orc.stencil5_entries(n, 5.0, -1.0)
    diag = ent["row"] == ent["col"]
    ent["value"][diag] = 4.0 + rng.uniform(0.0, 60.0, int(diag.sum()))
    return ent


@pytest.mark.parametrize("opname", [b"stencil5-csr", b"cusparse-csr", b"ellpack", b"stencil5-ellpack"])
@pytest.mark.parametrize("n", [40, 257])
def test_pcg_jacobi_matches_oracle_and_beats_cg(B, orc, torch_cuda, opname, n):
    """pcg_solve_device (not in the reference, parity = the oracle's restatement of the
    textbook recurrence over the reference's CG conventions): same iteration count as the oracle,
    solution within 1e-10, fewer iterations than plain CG on a variable-diagonal matrix"""
    L = B.load()
    N = n * n
    ent = variable_diagonal_stencil(orc, n, n)
    hm = B.HostMatrix.from_entries(N, N, ent, grid_size=n)
    orp, oci, ova = orc.build_csr(N, N, ent)
    rng = np.random.default_rng(1)
    b, x0 = rng.standard_normal(N), np.zeros(N)
    xp, sp, op = solve_device(B, opname, hm, b, x0, entry="pcg_solve_device")
    xc, sc, op = solve_device(B, opname, hm, b, x0, entry="cg_solve_device")
    op.contents.free()
    xo, ro = orc.pcg_device(orp, oci, ova, n, 1 if opname.startswith(b"stencil5") else 0, b, x0)
    assert sp["converged"] == 1 == ro["converged"] and sp["iterations"] == ro["iterations"]
    assert np.linalg.norm(xp - xo) / np.linalg.norm(xo) < 1e-10
    assert abs(sp["residual_norm"] - ro["residual_norm"]) <= 1e-9 * ro["b_norm"]
    assert sp["iterations"] < sc["iterations"]
    # both solve the same system
    assert np.linalg.norm(xp - xc) / np.linalg.norm(xc) < 1e-5


@pytest.mark.parametrize("n,P", [(40, 2), (130, 3), (257, 4)])
def test_pcg_mgpu_virtual_ranks_matches_oracle(B, orc, torch_cuda, n, P):
    """pcg_solve_mgpu_partitioned over P row bands (virtual ranks on one GPU): bands cut from the caller's
    entries, 1 / diag(A) per band, the edges of the preconditioned direction pushed by K3p -- same
    iteration count as the oracle, solution within 1e-10, and equal to the one-GPU PCG.  Two different
    matrices of the SAME shape back to back: the second solve must not see the first one's values."""
    L = B.load()
    N = n * n
    devs = (C.c_int * P)(*([0] * P))
    rng = np.random.default_rng(n + P)
    b = rng.standard_normal(N)
    for seed in (n, n + 1):
        ent = variable_diagonal_stencil(orc, n, seed)
        hm = B.HostMatrix.from_entries(N, N, ent, grid_size=n)
        orp, oci, ova = orc.build_csr(N, N, ent)
        xo, ro = orc.pcg_device(orp, oci, ova, n, 1, b, np.zeros(N))
        x1, s1, op = solve_device(B, b"stencil5-csr", hm, b, np.zeros(N), entry="pcg_solve_device")
        op.contents.free()
        assert L.b200_mgpu_init_single_process(P, devs, n) == 0
        try:
            x = np.zeros(N)
            st = B.CGStatsMultiGPU()
            rc = L.pcg_solve_mgpu_partitioned(None, hm.ptr(), b.ctypes.data, x.ctypes.data, B.cg_config(), C.byref(st))
            assert rc == 0
            assert st.converged == 1 and st.iterations == ro["iterations"] == s1["iterations"]
            assert np.linalg.norm(x - xo) / np.linalg.norm(xo) < 1e-10
            assert np.linalg.norm(x - x1) / np.linalg.norm(x1) < 1e-12
            assert abs(st.residual_norm - ro["residual_norm"]) <= 1e-9 * ro["b_norm"]
        finally:
            L.b200_mgpu_finalize()


@pytest.mark.parametrize("opname", [b"stencil5-csr", b"cusparse-csr", b"ellpack", b"stencil5-ellpack"])
@pytest.mark.parametrize("n", [40, 257])
def test_pcg_block_jacobi_matches_oracle(B, orc, torch_cuda, opname, n):
    """block-Jacobi with line blocks (one tridiagonal block per grid row; SURVEY 8f-3, parity = the oracle's
    restatement): same iteration count as the oracle, solution within 1e-10, not more iterations than Jacobi"""
    L = B.load()
    N = n * n
    ent = variable_diagonal_stencil(orc, n, n)
    hm = B.HostMatrix.from_entries(N, N, ent, grid_size=n)
    orp, oci, ova = orc.build_csr(N, N, ent)
    rng = np.random.default_rng(2)
    b, x0 = rng.standard_normal(N), np.zeros(N)
    xj, sj, op = solve_device(B, opname, hm, b, x0, entry="pcg_solve_device")
    assert L.b200_pcg_set_preconditioner(2) == 0
    try:
        xb, sb, op = solve_device(B, opname, hm, b, x0, entry="pcg_solve_device")
    finally:
        L.b200_pcg_set_preconditioner(1)
    op.contents.free()
    xo, ro = orc.pcg_block_device(orp, oci, ova, n, 1 if opname.startswith(b"stencil5") else 0, 1, b, x0)
    assert sb["converged"] == 1 == ro["converged"] and sb["iterations"] == ro["iterations"]
    assert np.linalg.norm(xb - xo) / np.linalg.norm(xo) < 1e-10
    assert abs(sb["residual_norm"] - ro["residual_norm"]) <= 1e-9 * ro["b_norm"]
    assert sb["iterations"] <= sj["iterations"]
    assert np.linalg.norm(xb - xj) / np.linalg.norm(xj) < 1e-5


@pytest.mark.parametrize("n,P", [(40, 2), (130, 3), (81, 4)])
def test_pcg_block_jacobi_mgpu_virtual_ranks(B, orc, torch_cuda, n, P):
    """the same over P row bands: blocks are clipped at the band boundaries (81 x 81 over 4 ranks cuts grid rows
    in the middle), the oracle uses the same blocks; edges of the new direction travel like in Jacobi PCG"""
    L = B.load()
    N = n * n
    devs = (C.c_int * P)(*([0] * P))
    ent = variable_diagonal_stencil(orc, n, n + P)
    hm = B.HostMatrix.from_entries(N, N, ent, grid_size=n)
    orp, oci, ova = orc.build_csr(N, N, ent)
    b = np.random.default_rng(P).standard_normal(N)
    xo, ro = orc.pcg_block_device(orp, oci, ova, n, 1, P, b, np.zeros(N))
    assert L.b200_pcg_set_preconditioner(2) == 0
    assert L.b200_mgpu_init_single_process(P, devs, n) == 0
    try:
        x = np.zeros(N)
        st = B.CGStatsMultiGPU()
        rc = L.pcg_solve_mgpu_partitioned(None, hm.ptr(), b.ctypes.data, x.ctypes.data, B.cg_config(), C.byref(st))
        assert rc == 0
        assert st.converged == 1 and st.iterations == ro["iterations"]
        assert np.linalg.norm(x - xo) / np.linalg.norm(xo) < 1e-10
    finally:
        L.b200_mgpu_finalize()
        L.b200_pcg_set_preconditioner(1)


def test_pcg_constant_diagonal_equals_cg_iterations(B, orc, torch_cuda):
    """on the 5 / -1 stencil D = 5 I: PCG is CG in exact arithmetic -- same iteration count"""
    n = 300
    N = n * n
    hm = B.HostMatrix.synthetic_stencil(n)
    b, x0 = np.ones(N), np.zeros(N)
    xp, sp, op = solve_device(B, b"stencil5-csr", hm, b, x0, entry="pcg_solve_device")
    xc, sc, op = solve_device(B, b"stencil5-csr", hm, b, x0, entry="cg_solve_device")
    op.contents.free()
    assert sp["iterations"] == sc["iterations"] and sp["converged"] == 1
    assert np.linalg.norm(xp - xc) / np.linalg.norm(xc) < 1e-10


def test_pcg_rejects_missing_diagonal(B, orc, torch_cuda):
    n = 6
    N = n * n
    ent = orc.stencil5_entries(n, 5.0, -1.0)
    ent = ent[~((ent["row"] == 7) & (ent["col"] == 7))]
    hm = B.HostMatrix.from_entries(N, N, ent, grid_size=-1)
    L = B.load()
    op = L.get_operator(b"cusparse-csr")
    assert op.contents.init(hm.ptr()) == 0
    x, st = np.zeros(N), B.CGStats()
    assert L.pcg_solve_device(op, hm.ptr(), np.ones(N).ctypes.data, x.ctypes.data, B.cg_config(), C.byref(st)) != 0
    op.contents.free()


@pytest.mark.parametrize("n", [130, 258, 274, 300])
def test_cg_operator_reinit_between_solves(B, orc, torch_cuda, n):
    """regression: a workspace re-used after op.free() + op.init() must pick up the operator's NEW
    device arrays (all three may move); stale row_ptr / col_idx pointers showed up as wrong boundary
    rows in one solve out of a few.  Six solves with re-initialisation in between, both schedules,
    each checked against the oracle after 2 iterations."""
    L = B.load()
    N = n * n
    hm = B.HostMatrix.synthetic_stencil(n)
    rng = np.random.default_rng(n)
    b, x0 = rng.standard_normal(N), rng.standard_normal(N)
    xo, ro, _ = oracle_solve(orc, n, 1, b, x0, max_iters=2)
    try:
        for k in range(6):
            L.b200_cg_set_schedule(k & 1)
            x, st, op = solve_device(B, b"stencil5-csr", hm, b, x0, max_iters=2)
            op.contents.free()
            assert st["iterations"] == 2
            assert abs(st["residual_norm"] - ro["residual_norm"]) <= 1e-10 * ro["residual_norm"], (k, n)
            assert np.linalg.norm(x - xo) / np.linalg.norm(xo) < 1e-12
    finally:
        L.b200_cg_set_schedule(1)


def test_cg_zero_initial_guess_is_not_uploaded(B, orc, torch_cuda):
    """x0 all (+0.0): the engine scans the caller's vector (host threads, while b is uploaded) and clears the
    device vector instead of uploading it -- same iterates bit for bit; anything else (a single non-zero at the
    very end, -0.0) is uploaded as before.  b200_last_h2d_bytes counts what was copied."""
    L = B.load()
    n = 300
    N = n * n
    hm = B.HostMatrix.synthetic_stencil(n)
    rng = np.random.default_rng(7)
    b = rng.standard_normal(N)
    old = L.b200_cg_set_skip_zero_x0(1)
    try:
        x_on, st_on, op = solve_device(B, b"stencil5-csr", hm, b, np.zeros(N))
        assert L.b200_last_h2d_bytes() == 8 * N
        op.contents.free()
        L.b200_cg_set_skip_zero_x0(0)
        x_off, st_off, op = solve_device(B, b"stencil5-csr", hm, b, np.zeros(N))
        assert L.b200_last_h2d_bytes() == 16 * N
        op.contents.free()
        assert st_on["iterations"] == st_off["iterations"] and st_on["residual_norm"] == st_off["residual_norm"]
        assert np.array_equal(x_on, x_off)
        L.b200_cg_set_skip_zero_x0(1)
        for x0 in (np.concatenate([np.zeros(N - 1), [1e-300]]), np.full(N, -0.0), np.concatenate([[3.0], np.zeros(N - 1)])):
            x, st, op = solve_device(B, b"stencil5-csr", hm, b, x0)
            assert L.b200_last_h2d_bytes() == 16 * N
            op.contents.free()
            xo, ro, _ = oracle_solve(orc, n, 1, b, x0)
            assert st["iterations"] == ro["iterations"]
            assert np.linalg.norm(x - xo) / np.linalg.norm(xo) < 1e-10
    finally:
        L.b200_cg_set_skip_zero_x0(old)
