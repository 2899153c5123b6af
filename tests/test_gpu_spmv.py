"""GPU parity: SpMV operators through the C ABI vs the CPU oracle on the same seeded inputs.

Bar: structure (generated CSR / ELLPACK / COO) bit-exact; STENCIL5, CSR (stream path) and ELLPACK
outputs BIT-EXACT against the oracle (same fma order as the reference's PTX); the CSR warp-per-row
path for long rows within 1e-12 relative L2 (different summation order, north_star tolerance)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def dptr(t):
    return C.c_void_p(t.data_ptr())


def vec_patterns(N, rng):
    """reference tests/helpers/cuda_test_utils.cpp:89-146 patterns (+ ours: incremental mod 97)"""
    return {
        "ones": np.ones(N),
        "incremental_mod97": (np.arange(N) % 97 + 1).astype(np.float64),
        "alternating": np.where(np.arange(N) % 2 == 0, 1.0, -1.0),
        "uniform": rng.random(N),
        "normal": rng.standard_normal(N),
    }


def device_stencil_csr(B, torch, n, off=0, nl=None, center=5.0, nb=-1.0):
    L = B.load()
    N = n * n
    nl = N - off if nl is None else nl
    lnnz = L.b200_stencil5_nnz_before(off + nl, n) - L.b200_stencil5_nnz_before(off, n)
    rp = torch.empty(nl + 1, dtype=torch.int32, device="cuda")
    ci = torch.empty(lnnz + 2, dtype=torch.int32, device="cuda")
    va = torch.zeros(lnnz + 2, dtype=torch.float64, device="cuda")
    B.check(L.b200_gen_stencil5_csr(n, off, nl, center, nb, dptr(rp), dptr(ci), dptr(va), None), "gen csr")
    torch.cuda.synchronize()
    return rp, ci, va, lnnz


@pytest.mark.parametrize("n", [1, 2, 3, 5, 33, 81])
def test_device_generation_bit_exact(B, orc, torch_cuda, n):
    torch = torch_cuda
    L = B.load()
    rp, ci, va, nnz = device_stencil_csr(B, torch, n)
    orp, oci, ova = orc.stencil5_csr_direct(n)
    assert np.array_equal(rp.cpu().numpy(), orp) and np.array_equal(ci.cpu().numpy()[:nnz], oci)
    assert np.array_equal(va.cpu().numpy()[:nnz], ova)
    # COO entries in generator emission order
    ent = torch.empty(nnz * 16, dtype=torch.uint8, device="cuda")
    B.check(L.b200_gen_stencil5_entries(n, 0, n * n, 5.0, -1.0, dptr(ent), None), "gen entries")
    assert ent.cpu().numpy().tobytes() == orc.stencil5_entries(n).tobytes()
    # ELLPACK
    idx = torch.empty(5 * n * n, dtype=torch.int32, device="cuda")
    val = torch.empty(5 * n * n, dtype=torch.float64, device="cuda")
    B.check(L.b200_gen_stencil5_ellpack(n, 0, n * n, 5.0, -1.0, dptr(idx), dptr(val), None), "gen ell")
    w, oidx, oval = orc.build_ellpack(orp.astype(np.int32), oci, ova, n * n, n * n)
    if w == 5:
        assert np.array_equal(idx.cpu().numpy(), oidx) and np.array_equal(val.cpu().numpy(), oval)


@pytest.mark.parametrize("n,P", [(81, 2), (81, 4), (9, 2), (16, 3)])
def test_device_generation_band_slices(B, orc, torch_cuda, n, P):
    torch = torch_cuda
    orp64, oci, ova = orc.stencil5_csr_direct(n)
    for g in range(P):
        nl, off = orc.partition(n * n, P, g)
        rp, ci, va, lnnz = device_stencil_csr(B, torch, n, off, nl)
        lrp, lci, lva = orc.local_csr_slice(orp64.astype(np.int32), oci, ova, off, nl)
        assert np.array_equal(rp.cpu().numpy(), lrp) and np.array_equal(ci.cpu().numpy()[:lnnz], lci)
        assert np.array_equal(va.cpu().numpy()[:lnnz], lva)


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 31, 33, 34, 64, 65, 66, 67, 81, 130, 257, 600])
@pytest.mark.parametrize("variant", [0, 3, 9, 12, 13, 20, 21, 22])
def test_stencil5_csr_bit_exact(B, orc, torch_cuda, n, variant):
    torch = torch_cuda
    L = B.load()
    N = n * n
    rng = np.random.default_rng(42)
    rp, ci, va, nnz = device_stencil_csr(B, torch, n)
    orp64, oci, ova = orc.stencil5_csr_direct(n)
    band = B.Band(rp.data_ptr(), ci.data_ptr(), va.data_ptr(), nnz + 2, 0, N, n, 0, None, None, None, None, 0,
                  7 if n > 40 else 0, variant)
    for name, xh in vec_patterns(N, rng).items():
        x = torch.from_numpy(xh).cuda()
        y = torch.full((N,), float("nan"), dtype=torch.float64, device="cuda")
        B.check(L.b200_stencil5_spmv(C.byref(band), dptr(x), dptr(y), None), "stencil5")
        torch.cuda.synchronize()
        yo = orc.stencil5_spmv(orp64.astype(np.int32), oci, ova, xh, n)
        assert np.array_equal(y.cpu().numpy(), yo), (n, variant, name)
    # x = 1 closed form: sum = N + 4n exactly
    x = torch.ones(N, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    B.check(L.b200_spmv_stencil5_csr(dptr(rp), dptr(ci), dptr(va), dptr(x), dptr(y), N, n, None), "wrapper")
    assert float(y.sum().item()) == N + 4 * n


def test_stencil5_nonstandard_values_bundled(B, orc, torch_cuda, tmp_path):
    """config[0]: bundled example81x81.mtx (centre -4) through the operator table, host vectors."""
    torch = torch_cuda
    L = B.load()
    p = str(tmp_path / "example81x81.mtx")
    orc.write_mtx_stencil5(81, p, "-4.0", "-1.0")
    hm = B.HostMatrix.from_mtx(p)
    rows, cols, nnz, grid, ent = orc.load_mtx(p)
    orp, oci, ova = orc.build_csr(rows, cols, ent)
    x = np.ones(rows)
    results = {}
    for name in (b"stencil5-csr", b"cusparse-csr", b"ellpack", b"stencil5-ellpack", b"stencil5", b"csr"):
        op = L.get_operator(name).contents
        assert op.init(hm.ptr()) == 0, name
        y = np.full(rows, np.nan)
        ms = C.c_double(-1.0)
        assert op.run_timed(x.ctypes.data, y.ctypes.data, C.byref(ms)) == 0
        assert ms.value > 0
        results[name] = y.copy()
        op.free()
    yo = orc.csr_spmv(orp, oci, ova, x)
    assert yo.sum() == -52164.0
    for name, y in results.items():
        assert np.array_equal(y, yo), name  # reference test: CSR == STENCIL5 element-wise <= 1e-12; here exact


@pytest.mark.parametrize("n,P", [(81, 2), (81, 4), (64, 8), (9, 2), (16, 3), (130, 3)])
def test_stencil5_halo_bands_bit_exact(B, orc, torch_cuda, n, P):
    """band SpMV with halos (the reference halo kernel's argument list) == oracle band == full product"""
    torch = torch_cuda
    L = B.load()
    N = n * n
    rng = np.random.default_rng(11)
    xh = rng.standard_normal(N)
    orp64, oci, ova = orc.stencil5_csr_direct(n)
    y_full = orc.stencil5_spmv(orp64.astype(np.int32), oci, ova, xh, n)
    for g in range(P):
        nl, off = orc.partition(N, P, g)
        rp, ci, va, lnnz = device_stencil_csr(B, torch, n, off, nl)
        xl = torch.from_numpy(xh[off:off + nl].copy()).cuda()
        hp = torch.from_numpy(xh[off - n:off].copy()).cuda() if g > 0 else None
        hn = torch.from_numpy(xh[off + nl:off + nl + n].copy()).cuda() if g < P - 1 else None
        y = torch.full((nl,), float("nan"), dtype=torch.float64, device="cuda")
        B.check(L.b200_spmv_stencil5_halo(dptr(rp), dptr(ci), dptr(va), dptr(xl), dptr(hp) if hp is not None else None,
                                          dptr(hn) if hn is not None else None, dptr(y), nl, off, N, n, None), "halo")
        torch.cuda.synchronize()
        assert np.array_equal(y.cpu().numpy(), y_full[off:off + nl]), (n, P, g)


@pytest.mark.parametrize("where", ["last", "cross_2_31", "cross_2_32_minus"])
def test_stencil5_band_beyond_2_31_rows(B, torch_cuda, where):
    """Weak scaling at 8 x 400M rows puts bands beyond row 2^31 (the reference's int ids overflow
    there): 64-bit row offsets, global column ids stored modulo 2^32 and read back as unsigned by
    the boundary rows.  A thin band of a 56576^2 grid (3.2e9 rows) is generated on the device and
    multiplied with halos; the expected product comes from the stencil formula in numpy."""
    torch = torch_cuda
    L = B.load()
    n = 56576
    N = n * n
    rows_band = 3 * n + 1234  # not grid-row aligned on purpose
    off = {"last": N - rows_band, "cross_2_31": (2 ** 31 // n) * n - n - 17, "cross_2_32_minus": N - 2 * rows_band - 5}[where]
    nl = rows_band
    rp, ci, va, lnnz = device_stencil_csr(B, torch, n, off, nl)
    rng = np.random.default_rng(5)
    xl_h, hp_h, hn_h = rng.standard_normal(nl), rng.standard_normal(n), rng.standard_normal(n)
    last = off + nl == N
    xl, hp = torch.from_numpy(xl_h).cuda(), torch.from_numpy(hp_h).cuda()
    hn = None if last else torch.from_numpy(hn_h).cuda()
    y = torch.full((nl,), float("nan"), dtype=torch.float64, device="cuda")
    band = B.Band(rp.data_ptr(), ci.data_ptr(), va.data_ptr(), lnnz + 2, off, nl, n, 0, hp.data_ptr(),
                  hn.data_ptr() if hn is not None else None, None, None, 0, 0, 0)
    B.check(L.b200_stencil5_spmv(C.byref(band), dptr(xl), dptr(y), None), "band spmv")
    torch.cuda.synchronize()
    # expected: x extended by the halos, y = 5 x[r] - x[r-1] - x[r+1] - x[r-n] - x[r+n] inside the grid,
    # accumulated in the reference order (W, C, E, N, S for interior rows; k order N,W,C,E,S at the boundary)
    ext = np.concatenate([hp_h, xl_h, hn_h if not last else np.zeros(n)])
    r = off + np.arange(nl, dtype=np.int64)
    i, j = r // n, r % n
    c = ext[n:n + nl]
    W = np.where(j > 0, ext[n - 1:n - 1 + nl], 0.0)
    E = np.where(j < n - 1, ext[n + 1:n + 1 + nl], 0.0)
    Nn = np.where(i > 0, ext[0:nl], 0.0)
    S = np.where(i < n - 1, ext[2 * n:2 * n + nl], 0.0)
    interior = (i > 0) & (i < n - 1) & (j > 0) & (j < n - 1)
    # every product with -1 is exact, so only the order of the additions matters (no fma difference)
    y_int = ((((5.0 * c) - W) - E) - Nn) - S
    y_bnd = np.zeros(nl)
    for cond, v in ((i > 0, Nn), (j > 0, W)):
        y_bnd = np.where(cond, y_bnd - v, y_bnd)
    y_bnd = y_bnd + 5.0 * c
    for cond, v in ((j < n - 1, E), (i < n - 1, S)):
        y_bnd = np.where(cond, y_bnd - v, y_bnd)
    yo = np.where(interior, y_int, y_bnd)
    yd = y.cpu().numpy()
    assert np.abs(yd - yo).max() <= 4e-15 * np.abs(yo).max()
    assert np.array_equal(yd[interior], yo[interior])


def random_csr(rng, rows, cols, lens):
    ent = []
    for r, ln in enumerate(lens):
        cs = rng.choice(cols, size=min(ln, cols), replace=False)
        for c in cs:
            ent.append((r, int(c), rng.uniform(-1, 1)))
    order = rng.permutation(len(ent))
    a = np.zeros(len(ent), dtype=[("row", np.int32), ("col", np.int32), ("value", np.float64)])
    for k, o in enumerate(order):
        a[k] = ent[o]
    return a


@pytest.mark.parametrize("variant", [0, 6, 100])
@pytest.mark.parametrize("kind", ["uniform5", "unbalanced", "empty_rows", "long_rows", "single", "mixed_tail"])
def test_generic_csr_and_ellpack(B, orc, torch_cuda, kind, variant):
    """row-length variety of tests/helpers/matrix_fixtures.cpp:296-370 (random_sparse, unbalanced_rows),
    through every tuning variant of the warp-ring kernel (and the legacy warp-stream kernel, 100)"""
    torch = torch_cuda
    L = B.load()
    if variant not in (0, 100) and L.b200_csr_variant_info(variant) is None:
        pytest.skip("no such variant")
    L.b200_csr_set_default_variant(variant)
    try:
        _generic_csr_and_ellpack(B, orc, torch, L, kind, variant)
    finally:
        L.b200_csr_set_default_variant(0)


def _generic_csr_and_ellpack(B, orc, torch, L, kind, variant):
    rng = np.random.default_rng(42)
    rows = {"uniform5": 3000, "unbalanced": 2000, "empty_rows": 1500, "long_rows": 300, "single": 1,
            "mixed_tail": 1237}[kind]
    cols = rows if kind != "long_rows" else 6000
    if kind == "uniform5":
        lens = np.full(rows, 5)
    elif kind == "unbalanced":
        lens = np.where(rng.random(rows) < 0.05, rng.integers(50, 400, rows), rng.integers(1, 6, rows))
    elif kind == "empty_rows":
        lens = np.where(rng.random(rows) < 0.5, 0, rng.integers(1, 9, rows))
    elif kind == "long_rows":
        lens = rng.integers(3000, 5000, rows)
    elif kind == "mixed_tail":  # short rows, a run of long rows, empty rows at the very end
        lens = rng.integers(0, 12, rows)
        lens[400:470] = rng.integers(200, 900, 70)
        lens[-40:] = 0
    else:
        lens = np.array([1])
    ent = random_csr(rng, rows, cols, lens)
    orp, oci, ova = orc.build_csr(rows, cols, ent.astype(orc.ENTRY_DTYPE))
    xh = rng.standard_normal(cols)
    yo = orc.csr_spmv(orp, oci, ova, xh)
    rp, ci, va = (torch.from_numpy(a).cuda() for a in (orp, oci, ova))
    if len(ova) == 0:
        ci = torch.zeros(1, dtype=torch.int32, device="cuda")
        va = torch.zeros(1, dtype=torch.float64, device="cuda")
    x = torch.from_numpy(xh).cuda()
    y = torch.full((rows,), float("nan"), dtype=torch.float64, device="cuda")
    plan = B.CsrPlan()
    B.check(L.b200_csr_plan_build(dptr(rp), rows, len(ova), C.byref(plan), None), "plan")
    assert sum(plan.hist) == rows and plan.max_row_len == lens.max()
    plan.variant = variant
    B.check(L.b200_spmv_csr(C.byref(plan), dptr(rp), dptr(ci), dptr(va), dptr(x), dptr(y), rows, 1.0, 0.0, None), "csr")
    torch.cuda.synchronize()
    yd = y.cpu().numpy()
    # groups with long rows go warp-per-row: different summation order, BASELINE tolerance 1e-12
    loose = kind in ("unbalanced", "long_rows", "mixed_tail")
    if loose:
        assert np.linalg.norm(yd - yo) / np.linalg.norm(yo) < 1e-12
    else:
        assert np.array_equal(yd, yo)
    # alpha / beta
    y2 = torch.from_numpy(np.arange(rows, dtype=np.float64)).cuda()
    B.check(L.b200_spmv_csr(C.byref(plan), dptr(rp), dptr(ci), dptr(va), dptr(x), dptr(y2), rows, 2.0, -0.5, None), "csr ab")
    ref2 = 2.0 * yo - 0.5 * np.arange(rows)
    assert np.allclose(y2.cpu().numpy(), ref2, rtol=1e-12, atol=1e-12)
    # ELLPACK of the same matrix (skip the 5000-wide one: MAX_WIDTH 1000)
    w, oidx, oval = orc.build_ellpack(orp, oci, ova, rows, cols)
    if 1 <= w <= 1000:
        idx, val = torch.from_numpy(oidx).cuda(), torch.from_numpy(oval).cuda()
        y3 = torch.full((rows,), float("nan"), dtype=torch.float64, device="cuda")
        B.check(L.b200_spmv_ellpack(dptr(idx), dptr(val), dptr(x), dptr(y3), rows, w, 1.0, 0.0, None), "ell")
        torch.cuda.synchronize()
        y3h = y3.cpu().numpy()
        if loose:  # wide ELLPACK rows are walked warp-per-row as well
            assert np.linalg.norm(y3h - yo) / np.linalg.norm(yo) < 1e-12
        else:
            assert np.array_equal(y3h, yo)
    else:
        assert L.b200_spmv_ellpack(dptr(rp), dptr(va), dptr(x), dptr(y), rows, 1001, 1.0, 0.0, None) == 1


@pytest.mark.parametrize("variant", [0, 6])
@pytest.mark.parametrize("shift", [0, 1, 3])
def test_generic_csr_many_groups_per_warp(B, orc, torch_cuda, variant, shift):
    """1200 x 1200 stencil through the GENERIC kernels: every persistent warp walks many 32-row groups
    and several row_ptr chunks; `shift` mis-aligns all three arrays against the 16-byte bulk-copy
    granularity (head / tail patches).  Bit-exact against the sequential-k oracle."""
    torch = torch_cuda
    L = B.load()
    if L.b200_csr_variant_info(variant) is None:
        pytest.skip("no such variant")
    n = 1200
    N = n * n
    orp64, oci, ova = orc.stencil5_csr_direct(n)
    orp = orp64.astype(np.int32)
    rng = np.random.default_rng(7)
    xh = rng.standard_normal(N)
    yo = orc.csr_spmv(orp, oci, ova, xh)

    def shifted(a):
        t = torch.empty(len(a) + shift, dtype=torch.from_numpy(a[:1]).dtype, device="cuda")
        t[shift:] = torch.from_numpy(a).cuda()
        return t[shift:]
    rp, ci, va = shifted(orp), shifted(oci), shifted(ova)
    x = torch.from_numpy(xh).cuda()
    y = torch.full((N,), float("nan"), dtype=torch.float64, device="cuda")
    plan = B.CsrPlan()
    B.check(L.b200_csr_plan_build(dptr(rp), N, len(ova), C.byref(plan), None), "plan")
    plan.variant = variant
    B.check(L.b200_spmv_csr(C.byref(plan), dptr(rp), dptr(ci), dptr(va), dptr(x), dptr(y), N, 1.0, 0.0, None), "csr")
    torch.cuda.synchronize()
    assert np.array_equal(y.cpu().numpy(), yo)
    w, oidx, oval = orc.build_ellpack(orp, oci, ova, N, N)
    assert w == 5
    idx, val = shifted(oidx), shifted(oval)
    y.fill_(float("nan"))
    L.b200_csr_set_default_variant(variant)
    try:
        B.check(L.b200_spmv_ellpack(dptr(idx), dptr(val), dptr(x), dptr(y), N, w, 1.0, 0.0, None), "ell")
        torch.cuda.synchronize()
    finally:
        L.b200_csr_set_default_variant(0)
    assert np.array_equal(y.cpu().numpy(), yo)


@pytest.mark.parametrize("rowlen", [5, 12, 20, 40])
def test_csr_plan_picks_ring_by_row_length_histogram(B, orc, torch_cuda, rowlen):
    """the plan's histogram picks the ring: median row <= 8 -> small ring (variant 0), 9..32 -> large
    ring (variant 6) so that such rows stay lane-per-row = bit-exact; longer -> warp-per-row (1e-12)"""
    torch = torch_cuda
    L = B.load()
    rng = np.random.default_rng(rowlen)
    rows = cols = 5000
    lens = np.full(rows, rowlen)
    ent = random_csr(rng, rows, cols, lens)
    orp, oci, ova = orc.build_csr(rows, cols, ent.astype(orc.ENTRY_DTYPE))
    xh = rng.standard_normal(cols)
    yo = orc.csr_spmv(orp, oci, ova, xh)
    rp, ci, va = (torch.from_numpy(a).cuda() for a in (orp, oci, ova))
    x = torch.from_numpy(xh).cuda()
    y = torch.full((rows,), float("nan"), dtype=torch.float64, device="cuda")
    plan = B.CsrPlan()
    B.check(L.b200_csr_plan_build(dptr(rp), rows, len(ova), C.byref(plan), None), "plan")
    assert plan.variant == (6 if 8 < rowlen <= 32 else 0)
    B.check(L.b200_spmv_csr(C.byref(plan), dptr(rp), dptr(ci), dptr(va), dptr(x), dptr(y), rows, 1.0, 0.0, None), "csr")
    torch.cuda.synchronize()
    yd = y.cpu().numpy()
    if rowlen <= 24:
        assert np.array_equal(yd, yo)
    else:
        assert np.linalg.norm(yd - yo) / np.linalg.norm(yo) < 1e-12
    w, oidx, oval = orc.build_ellpack(orp, oci, ova, rows, cols)
    idx, val = torch.from_numpy(oidx).cuda(), torch.from_numpy(oval).cuda()
    y.fill_(float("nan"))
    B.check(L.b200_spmv_ellpack(dptr(idx), dptr(val), dptr(x), dptr(y), rows, w, 1.0, 0.0, None), "ell")
    torch.cuda.synchronize()
    if rowlen <= 24:
        assert np.array_equal(y.cpu().numpy(), yo)
    else:
        assert np.linalg.norm(y.cpu().numpy() - yo) / np.linalg.norm(yo) < 1e-12


@pytest.mark.parametrize("fmt", ["csr", "ell"])
@pytest.mark.parametrize("n", [7, 130, 900])
def test_generic_spmv_with_fused_dot(B, orc, torch_cuda, fmt, n):
    """b200_spmv_csr_dot / b200_spmv_ellpack_dot (CG: Ap = A p and the p.Ap partials in one launch):
    y bit-identical to the plain launch, partials summed in item order == x.y to rounding, and the
    launch is a no-op once the scalars say converged"""
    torch = torch_cuda
    L = B.load()
    N = n * n
    orp64, oci, ova = orc.stencil5_csr_direct(n)
    orp = orp64.astype(np.int32)
    rng = np.random.default_rng(n)
    xh = rng.standard_normal(N)
    yo = orc.csr_spmv(orp, oci, ova, xh)
    x = torch.from_numpy(xh).cuda()
    y = torch.full((N,), float("nan"), dtype=torch.float64, device="cuda")
    cap = L.b200_csr_dot_partials_capacity(N)
    partials = torch.full((cap,), float("nan"), dtype=torch.float64, device="cuda")
    sc = torch.zeros(L.b200_cg_scalars_bytes() // 8 + 1, dtype=torch.float64, device="cuda")
    npart = C.c_int(0)
    if fmt == "csr":
        rp, ci, va = (torch.from_numpy(a).cuda() for a in (orp, oci, ova))
        plan = B.CsrPlan()
        B.check(L.b200_csr_plan_build(dptr(rp), N, len(ova), C.byref(plan), None), "plan")

        def launch(out):
            return L.b200_spmv_csr_dot(C.byref(plan), dptr(rp), dptr(ci), dptr(va), dptr(x), dptr(out), N, dptr(partials),
                                       cap, C.byref(npart), dptr(sc), None)
    else:
        w, oidx, oval = orc.build_ellpack(orp, oci, ova, N, N)
        idx, val = torch.from_numpy(oidx).cuda(), torch.from_numpy(oval).cuda()

        def launch(out):
            return L.b200_spmv_ellpack_dot(dptr(idx), dptr(val), dptr(x), dptr(out), N, w, dptr(partials), cap,
                                           C.byref(npart), dptr(sc), None)
    B.check(launch(y), "spmv+dot")
    torch.cuda.synchronize()
    assert np.array_equal(y.cpu().numpy(), yo)
    assert 1 <= npart.value <= cap
    got = float(partials[:npart.value].cpu().numpy().sum())
    ref = float(np.dot(xh, yo))
    assert abs(got - ref) <= 1e-12 * float(np.dot(np.abs(xh), np.abs(yo)))
    # too small a buffer is refused, not overrun
    assert L.b200_spmv_csr_dot is not None
    # converged flag set -> no-op
    conv_off = None
    st = np.zeros(sc.numel(), dtype=np.float64)
    raw = st.view(np.int32)
    raw[2 * 7] = 1  # CGScalars.converged: 7 doubles, then the ints (csrc/cg_kernels.cuh)
    sc.copy_(torch.from_numpy(st))
    y2 = torch.full((N,), -7.0, dtype=torch.float64, device="cuda")
    B.check(launch(y2), "spmv+dot converged")
    torch.cuda.synchronize()
    assert float(y2.min().item()) == -7.0 and float(y2.max().item()) == -7.0


def test_stencil5_ellpack_kernel_signature(B, orc, torch_cuda):
    """include/spmv_stencil.h:40-42 argument list incl. alpha / beta"""
    torch = torch_cuda
    L = B.load()
    n = 70
    N = n * n
    rng = np.random.default_rng(1)
    idx = torch.empty(5 * N, dtype=torch.int32, device="cuda")
    val = torch.empty(5 * N + 2, dtype=torch.float64, device="cuda")
    B.check(L.b200_gen_stencil5_ellpack(n, 0, N, 5.0, -1.0, dptr(idx), dptr(val), None), "gen ell")
    xh = rng.random(N)
    x = torch.from_numpy(xh).cuda()
    orp64, oci, ova = orc.stencil5_csr_direct(n)
    yo = orc.stencil5_spmv(orp64.astype(np.int32), oci, ova, xh, n)
    y = torch.full((N,), float("nan"), dtype=torch.float64, device="cuda")
    B.check(L.b200_spmv_stencil5_ellpack(dptr(val), dptr(idx), dptr(x), dptr(y), N, 5, 1.0, 0.0, n, None), "st-ell")
    assert np.array_equal(y.cpu().numpy(), yo)
    y0 = rng.random(N)
    y = torch.from_numpy(y0.copy()).cuda()
    B.check(L.b200_spmv_stencil5_ellpack(dptr(val), dptr(idx), dptr(x), dptr(y), N, 5, 0.5, 2.0, n, None), "st-ell ab")
    assert np.allclose(y.cpu().numpy(), 0.5 * yo + 2.0 * y0, rtol=1e-13, atol=1e-13)


def test_error_paths(B, torch_cuda):
    torch = torch_cuda
    L = B.load()
    x = torch.ones(16, dtype=torch.float64, device="cuda")
    band = B.Band()
    assert L.b200_stencil5_spmv(C.byref(band), dptr(x), dptr(x), None) == 1  # B200_EINVAL
    assert b"stencil5" in L.b200_last_error()
    assert L.b200_spmv_stencil5_csr(dptr(x), dptr(x), dptr(x), dptr(x), dptr(x), 17, 4, None) == 1
    op = L.get_operator(b"stencil5-csr").contents
    md = B.MatrixData(10, 10, 10, -1, None)
    assert op.init(C.byref(md)) != 0  # no grid size: not a stencil matrix


@pytest.mark.parametrize("n", [2000])
def test_stencil5_large_properties(B, torch_cuda, n):
    """full-size style checks with size-independent properties: x=1 closed form, linearity"""
    torch = torch_cuda
    L = B.load()
    N = n * n
    rp, ci, va, nnz = device_stencil_csr(B, torch, n)
    band = B.Band(rp.data_ptr(), ci.data_ptr(), va.data_ptr(), nnz + 2, 0, N, n, 0, None, None, None, None, 0, 0, 0)
    ones = torch.ones(N, dtype=torch.float64, device="cuda")
    y = torch.empty_like(ones)
    B.check(L.b200_stencil5_spmv(C.byref(band), dptr(ones), dptr(y), None), "big")
    assert float(y.sum().item()) == N + 4 * n
    yv = y.view(n, n)
    assert float(yv[1:-1, 1:-1].min()) == 1.0 == float(yv[1:-1, 1:-1].max())
    g = torch.Generator(device="cuda").manual_seed(42)
    a = torch.randint(-8, 9, (N,), generator=g, device="cuda").double()
    b = torch.randint(-8, 9, (N,), generator=g, device="cuda").double()
    ya, yb, yab = torch.empty_like(a), torch.empty_like(a), torch.empty_like(a)
    ab = a + 2.0 * b
    for src, dst in ((a, ya), (b, yb), (ab, yab)):
        B.check(L.b200_stencil5_spmv(C.byref(band), dptr(src), dptr(dst), None), "lin")
    assert torch.equal(yab, ya + 2.0 * yb)  # small integers: exact in f64
