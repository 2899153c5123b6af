"""Guard-band tests: every device array a kernel touches sits between poisoned guard regions.

compute-sanitizer is closed on the GPU pool this repo is measured on ("runs under it have left GPUs
needing a reset", see profiles/sanitizer_r02.md), so out-of-bounds accesses are hunted the way the pool
asks for: small cases, guards of our own, comparison with the CPU oracle.
  * a write outside an output array lands in its guard -> the guard is no longer intact;
  * a read outside an input array that reaches the result picks up the guard's NaN (float arrays) or a
    wild index (int arrays, guard value far outside any array) -> the result differs from the oracle.
The band sizes are chosen ragged on purpose (bands that start and end in the middle of a grid row, odd nnz
counts so that the 16-byte bulk copies meet their manual tails, row counts that are not multiples of 32)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GUARD = 1024  # elements on either side


class Guarded:
    """device array of `n` elements inside [guard | payload | guard]; float guards are NaN, int guards 0x7f7f7f7f"""

    def __init__(self, torch, data=None, n=None, dtype=None, align_shift=0):
        self.torch = torch
        if data is not None:
            data = np.ascontiguousarray(data)
            n, dtype = len(data), {np.dtype("float64"): torch.float64, np.dtype("int32"): torch.int32}[data.dtype]
        self.n, self.dtype = n, dtype
        self.fill = float("nan") if dtype == torch.float64 else 0x7F7F7F7F
        # align_shift (elements) keeps the payload 16-byte aligned (bulk copies) when it is even
        self.off = GUARD + align_shift
        self.buf = torch.full((self.off + n + GUARD,), self.fill, dtype=dtype, device="cuda")
        if data is not None:
            self.buf[self.off:self.off + n] = torch.from_numpy(data).cuda()

    def ptr(self):
        return C.c_void_p(self.buf.data_ptr() + self.off * self.buf.element_size())

    def payload(self):
        return self.buf[self.off:self.off + self.n].cpu().numpy()

    def guards_intact(self):
        lo, hi = self.buf[:self.off], self.buf[self.off + self.n:]
        if self.dtype == self.torch.float64:
            return bool(self.torch.isnan(lo).all() and self.torch.isnan(hi).all())
        return bool((lo == self.fill).all() and (hi == self.fill).all())


@pytest.mark.parametrize("variant", [0, 3, 9, 12, 13, 20, 21, 22])
@pytest.mark.parametrize("n,P", [(37, 3), (64, 5), (130, 4)])
def test_stencil5_bands_between_guards(B, orc, torch_cuda, n, P, variant):
    torch = torch_cuda
    L = B.load()
    N = n * n
    rng = np.random.default_rng(n * 100 + P)
    xh = rng.standard_normal(N)
    rp64, oci, ova = orc.stencil5_csr_direct(n)
    orp = rp64.astype(np.int32)
    y_full = orc.stencil5_spmv(orp, oci, ova, xh, n)
    # ragged bands: cut points in the middle of grid rows
    cuts = sorted({0, N} | {int(c) for c in rng.integers(n + 1, N - n - 1, P - 1)})
    for g in range(len(cuts) - 1):
        off, nl = cuts[g], cuts[g + 1] - cuts[g]
        if nl < n:
            continue
        lrp, lci, lva = orc.local_csr_slice(orp, oci, ova, off, nl)
        lnnz = len(lva)
        vals = Guarded(torch, np.concatenate([lva, np.zeros(2)]))  # values_len = lnnz + 2 (bulk-copy granule)
        cols = Guarded(torch, np.concatenate([lci, np.zeros(2, dtype=np.int32)]))
        rowp = Guarded(torch, lrp)
        x = Guarded(torch, xh[off:off + nl])
        hp = Guarded(torch, xh[off - n:off]) if off > 0 else None
        hn = Guarded(torch, xh[off + nl:off + nl + n]) if off + nl < N else None
        y = Guarded(torch, n=nl, dtype=torch.float64)
        band = B.Band(rowp.ptr(), cols.ptr(), vals.ptr(), lnnz + 2, off, nl, n, 0, hp.ptr() if hp else None,
                      hn.ptr() if hn else None, None, None, 0, 3, variant)
        B.check(L.b200_stencil5_spmv(C.byref(band), x.ptr(), y.ptr(), None), "spmv")
        torch.cuda.synchronize()
        assert np.array_equal(y.payload(), y_full[off:off + nl]), (n, g, variant)
        for a in (vals, cols, rowp, x, y) + tuple(t for t in (hp, hn) if t):
            assert a.guards_intact()


@pytest.mark.parametrize("variant", [0, 6, 100])
@pytest.mark.parametrize("kind", ["short", "medium", "long", "mixed"])
def test_generic_csr_between_guards(B, orc, torch_cuda, kind, variant):
    """lane-per-row, sub-warp and warp-per-row groups; 997 rows (not a multiple of 32), odd nnz"""
    torch = torch_cuda
    L = B.load()
    rng = np.random.default_rng(len(kind) + variant)
    rows, cols = 997, 1301
    lens = {"short": rng.integers(0, 8, rows), "medium": rng.integers(9, 60, rows), "long": rng.integers(70, 400, rows),
            "mixed": np.where(rng.random(rows) < 0.1, rng.integers(100, 900, rows), rng.integers(0, 30, rows))}[kind]
    rp = np.zeros(rows + 1, dtype=np.int32)
    np.cumsum(lens, out=rp[1:])
    nnz = int(rp[-1])
    ci = np.concatenate([np.sort(rng.choice(cols, size=int(k), replace=False)) for k in lens]).astype(np.int32)
    va = rng.uniform(-1, 1, nnz)
    xh = rng.standard_normal(cols)
    yo = orc.csr_spmv(rp, ci, va, xh)
    d_rp, d_ci, d_va, d_x = Guarded(torch, rp), Guarded(torch, ci), Guarded(torch, va), Guarded(torch, xh)
    d_y = Guarded(torch, n=rows, dtype=torch.float64)
    plan = B.CsrPlan()
    B.check(L.b200_csr_plan_build(d_rp.ptr(), rows, nnz, C.byref(plan), None), "plan")
    plan.variant = variant
    B.check(L.b200_spmv_csr(C.byref(plan), d_rp.ptr(), d_ci.ptr(), d_va.ptr(), d_x.ptr(), d_y.ptr(), rows, 1.0, 0.0, None), "csr")
    torch.cuda.synchronize()
    yd = d_y.payload()
    assert np.all(np.isfinite(yd))
    assert np.linalg.norm(yd - yo) <= 1e-12 * np.linalg.norm(yo)
    for a in (d_rp, d_ci, d_va, d_x, d_y):
        assert a.guards_intact()


@pytest.mark.parametrize("width", [1, 5, 8, 17, 40])
def test_generic_ellpack_between_guards(B, orc, torch_cuda, width):
    torch = torch_cuda
    L = B.load()
    rng = np.random.default_rng(width)
    rows = cols = 1013
    lens = rng.integers(0, width + 1, rows)
    lens[rng.integers(0, rows)] = width
    idx = np.full((rows, width), -1, dtype=np.int32)
    val = np.zeros((rows, width))
    for r in range(rows):
        k = int(lens[r])
        idx[r, :k] = np.sort(rng.choice(cols, size=k, replace=False))
        val[r, :k] = rng.uniform(-1, 1, k)
    xh = rng.standard_normal(cols)
    yo = orc.ell_spmv(width, idx.ravel(), val.ravel(), xh, rows, cols)
    d_i, d_v, d_x = Guarded(torch, idx.ravel()), Guarded(torch, val.ravel()), Guarded(torch, xh)
    d_y = Guarded(torch, n=rows, dtype=torch.float64)
    B.check(L.b200_spmv_ellpack(d_i.ptr(), d_v.ptr(), d_x.ptr(), d_y.ptr(), rows, width, 1.0, 0.0, None), "ell")
    torch.cuda.synchronize()
    yd = d_y.payload()
    assert np.all(np.isfinite(yd))
    assert np.linalg.norm(yd - yo) <= 1e-12 * max(np.linalg.norm(yo), 1e-300)
    for a in (d_i, d_v, d_x, d_y):
        assert a.guards_intact()
