"""ctypes bindings for the CPU oracle (oracle/liboracle.so) and, when present, the reference's
own host code compiled into oracle/_ref/libref_host.so.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_HOST_PATH = os.path.join(HERE, "_ref", "libref_host.so")


class Entry(C.Structure):
    _fields_ = [("row", C.c_int), ("col", C.c_int), ("value", C.c_double)]


ENTRY_DTYPE = np.dtype([("row", np.int32), ("col", np.int32), ("value", np.float64)], align=True)
assert ENTRY_DTYPE.itemsize == C.sizeof(Entry) == 16


class Matrix(C.Structure):
    _fields_ = [("rows", C.c_int), ("cols", C.c_int), ("nnz", C.c_int), ("grid_size", C.c_int),
                ("entries", C.POINTER(Entry))]


class CSR(C.Structure):
    _fields_ = [("nb_rows", C.c_int), ("nb_cols", C.c_int), ("nb_nonzeros", C.c_int),
                ("row_ptr", C.POINTER(C.c_int)), ("col_indices", C.POINTER(C.c_int)),
                ("values", C.POINTER(C.c_double))]


class ELL(C.Structure):
    _fields_ = [("nb_rows", C.c_int), ("nb_cols", C.c_int), ("ell_width", C.c_int),
                ("grid_size", C.c_int), ("indices", C.POINTER(C.c_int)), ("nb_nonzeros", C.c_int),
                ("values", C.POINTER(C.c_double))]


class CGResult(C.Structure):
    _fields_ = [("iterations", C.c_int), ("converged", C.c_int), ("residual_norm", C.c_double),
                ("b_norm", C.c_double), ("solution_sum", C.c_double), ("solution_norm", C.c_double)]


class BenchStats(C.Structure):
    _fields_ = [("median_ms", C.c_double), ("mean_ms", C.c_double), ("std_dev_ms", C.c_double),
                ("min_ms", C.c_double), ("max_ms", C.c_double), ("valid_runs", C.c_int),
                ("outliers_removed", C.c_int)]


def build():
    """(Re)build liboracle.so -- and oracle/_ref when the reference tree is mounted."""
    subprocess.run(["make", "-C", HERE, "-s"], check=True)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        L = _lib
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.orc_stencil5_nnz.restype = C.c_longlong
        L.orc_stencil5_nnz.argtypes = [C.c_int]
        L.orc_stencil5_entries.argtypes = [C.c_int, C.c_double, C.c_double, C.c_void_p]
        L.orc_write_mtx_stencil5.argtypes = [C.c_int, C.c_char_p, C.c_char_p, C.c_char_p]
        L.orc_load_mtx.argtypes = [C.c_char_p, C.POINTER(Matrix)]
        L.orc_build_csr.argtypes = [C.POINTER(Matrix), C.POINTER(CSR)]
        L.orc_stencil5_csr_direct.argtypes = [C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_build_ellpack.argtypes = [C.POINTER(CSR), C.POINTER(ELL), ip]
        L.orc_interior_csr_offset.restype = C.c_int64
        L.orc_interior_csr_offset.argtypes = [C.c_int64, C.c_int]
        L.orc_csr_spmv.argtypes = [C.c_void_p] * 5 + [C.c_int]
        L.orc_stencil5_spmv.argtypes = [C.c_void_p] * 5 + [C.c_int, C.c_int]
        L.orc_ell_spmv.argtypes = [C.POINTER(ELL), C.c_void_p, C.c_void_p]
        L.orc_halo_spmv.argtypes = [C.c_void_p] * 7 + [C.c_int, C.c_int64, C.c_int64, C.c_int]
        L.orc_dot_blocktree.restype = C.c_double
        L.orc_dot_blocktree.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.orc_dot_sequential.restype = C.c_double
        L.orc_dot_sequential.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.orc_cg_device.argtypes = [C.POINTER(CSR), C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                    C.c_double, C.POINTER(CGResult), C.c_void_p, C.c_int]
        L.orc_cg_mgpu.argtypes = [C.POINTER(CSR), C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                  C.c_double, C.POINTER(CGResult)]
        L.orc_partition.argtypes = [C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.orc_local_csr_slice.restype = C.c_int64
        L.orc_local_csr_slice.argtypes = [C.POINTER(CSR), C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_halo_ranges.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_int] + [C.POINTER(C.c_int64)] * 4
        L.orc_bench_stats_from_times.argtypes = [C.c_void_p, C.c_int, C.POINTER(BenchStats)]
        L.orc_num_threads.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


# ---------------------------------------------------------------- structure
def stencil5_nnz(n):
    return int(lib().orc_stencil5_nnz(n))


def stencil5_entries(n, center=5.0, neighbour=-1.0):
    """COO entries in the generator's emission order, as a structured array (row, col, value)."""
    out = np.zeros(stencil5_nnz(n), dtype=ENTRY_DTYPE)
    lib().orc_stencil5_entries(n, center, neighbour, _p(out))
    return out


def write_mtx_stencil5(n, path, center_txt="5.0", nb_txt="-1.0"):
    rc = lib().orc_write_mtx_stencil5(n, path.encode(), center_txt.encode(), nb_txt.encode())
    if rc:
        raise OSError("orc_write_mtx_stencil5 failed: %d" % rc)


def load_mtx(path):
    """-> (rows, cols, nnz, grid_size, entries structured array)"""
    m = Matrix()
    rc = lib().orc_load_mtx(path.encode(), C.byref(m))
    if rc:
        raise OSError("orc_load_mtx failed: %d" % rc)
    ent = np.ctypeslib.as_array(C.cast(m.entries, C.POINTER(C.c_byte)), shape=(m.nnz * 16,)).copy().view(ENTRY_DTYPE)
    C.CDLL(None).free(m.entries)
    return m.rows, m.cols, m.nnz, m.grid_size, ent


def build_csr(rows, cols, entries):
    """COO (structured array) -> (row_ptr int32[rows+1], col int32[nnz], val f64[nnz])"""
    entries = np.ascontiguousarray(entries, dtype=ENTRY_DTYPE)
    m = Matrix(rows, cols, len(entries), -1, C.cast(_p(entries), C.POINTER(Entry)))
    c = CSR()
    rc = lib().orc_build_csr(C.byref(m), C.byref(c))
    if rc:
        raise MemoryError("orc_build_csr")
    nnz = len(entries)
    rp = np.ctypeslib.as_array(c.row_ptr, shape=(rows + 1,)).copy()
    ci = np.ctypeslib.as_array(c.col_indices, shape=(max(nnz, 1),)).copy()[:nnz]
    va = np.ctypeslib.as_array(c.values, shape=(max(nnz, 1),)).copy()[:nnz]
    lib().orc_free_csr(C.byref(c))
    return rp, ci, va


def stencil5_csr_direct(n, center=5.0, neighbour=-1.0):
    """closed-form stencil CSR -> (row_ptr int64[N+1], col int32, val f64)"""
    N, nnz = n * n, stencil5_nnz(n)
    rp = np.zeros(N + 1, dtype=np.int64)
    ci = np.zeros(nnz, dtype=np.int32)
    va = np.zeros(nnz, dtype=np.float64)
    lib().orc_stencil5_csr_direct(n, center, neighbour, _p(rp), _p(ci), _p(va))
    return rp, ci, va


def _csr_struct(rp, ci, va, rows, cols):
    rp = np.ascontiguousarray(rp, dtype=np.int32)
    ci = np.ascontiguousarray(ci, dtype=np.int32)
    va = np.ascontiguousarray(va, dtype=np.float64)
    c = CSR(rows, cols, len(va), C.cast(_p(rp), C.POINTER(C.c_int)), C.cast(_p(ci), C.POINTER(C.c_int)),
            C.cast(_p(va), C.POINTER(C.c_double)))
    c._keep = (rp, ci, va)
    return c


def build_ellpack(rp, ci, va, rows, cols):
    """-> (width, indices int32[rows*width], values f64[rows*width]) row-major, pad = (-1, 0.0)"""
    c = _csr_struct(rp, ci, va, rows, cols)
    e = ELL()
    w = C.c_int(0)
    rc = lib().orc_build_ellpack(C.byref(c), C.byref(e), C.byref(w))
    if rc:
        raise MemoryError("orc_build_ellpack")
    tot = rows * max(w.value, 1)
    idx = np.ctypeslib.as_array(e.indices, shape=(tot,)).copy()[: rows * w.value]
    val = np.ctypeslib.as_array(e.values, shape=(tot,)).copy()[: rows * w.value]
    lib().orc_free_ell(C.byref(e))
    return w.value, idx, val


def interior_csr_offset(row, grid):
    return int(lib().orc_interior_csr_offset(row, grid))


# ---------------------------------------------------------------- SpMV
def csr_spmv(rp, ci, va, x):
    rp = np.ascontiguousarray(rp, dtype=np.int32); ci = np.ascontiguousarray(ci, dtype=np.int32)
    va = np.ascontiguousarray(va, dtype=np.float64); x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.zeros(len(rp) - 1, dtype=np.float64)
    lib().orc_csr_spmv(_p(rp), _p(ci), _p(va), _p(x), _p(y), len(rp) - 1)
    return y


def stencil5_spmv(rp, ci, va, x, grid):
    rp = np.ascontiguousarray(rp, dtype=np.int32); ci = np.ascontiguousarray(ci, dtype=np.int32)
    va = np.ascontiguousarray(va, dtype=np.float64); x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.zeros(len(rp) - 1, dtype=np.float64)
    lib().orc_stencil5_spmv(_p(rp), _p(ci), _p(va), _p(x), _p(y), len(rp) - 1, grid)
    return y


def ell_spmv(width, idx, val, x, rows, cols):
    idx = np.ascontiguousarray(idx, dtype=np.int32); val = np.ascontiguousarray(val, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    e = ELL(rows, cols, width, -1, C.cast(_p(idx), C.POINTER(C.c_int)), 0, C.cast(_p(val), C.POINTER(C.c_double)))
    y = np.zeros(rows, dtype=np.float64)
    lib().orc_ell_spmv(C.byref(e), _p(x), _p(y))
    return y


def halo_spmv(rp_local, col_global, val, x_local, halo_prev, halo_next, row_offset, N, grid):
    rp_local = np.ascontiguousarray(rp_local, dtype=np.int32)
    col_global = np.ascontiguousarray(col_global, dtype=np.int32)
    val = np.ascontiguousarray(val, dtype=np.float64)
    x_local = np.ascontiguousarray(x_local, dtype=np.float64)
    hp = None if halo_prev is None else np.ascontiguousarray(halo_prev, dtype=np.float64)
    hn = None if halo_next is None else np.ascontiguousarray(halo_next, dtype=np.float64)
    nl = len(rp_local) - 1
    y = np.zeros(nl, dtype=np.float64)
    lib().orc_halo_spmv(_p(rp_local), _p(col_global), _p(val), _p(x_local), _p(hp), _p(hn), _p(y), nl,
                        row_offset, N, grid)
    return y


# ---------------------------------------------------------------- reductions / CG
def dot_blocktree(x, y):
    x = np.ascontiguousarray(x, dtype=np.float64); y = np.ascontiguousarray(y, dtype=np.float64)
    return float(lib().orc_dot_blocktree(len(x), _p(x), _p(y)))


def dot_sequential(x, y):
    x = np.ascontiguousarray(x, dtype=np.float64); y = np.ascontiguousarray(y, dtype=np.float64)
    return float(lib().orc_dot_sequential(len(x), _p(x), _p(y)))


def cg_device(rp, ci, va, grid, op, b, x0, max_iters=1000, tol=1e-6, hist=64, inplace=False):
    """Restated cg_solve_device.  op: 0 generic CSR, 1 stencil5.  -> (x, result dict, rel history)
    inplace: x0 (contiguous float64) is overwritten with the solution instead of being copied."""
    rows = len(rp) - 1
    c = _csr_struct(rp, ci, va, rows, rows)
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = x0 if inplace else np.array(x0, dtype=np.float64, copy=True)
    assert x.dtype == np.float64 and x.flags["C_CONTIGUOUS"]
    res = CGResult()
    h = np.zeros(max(hist, 1), dtype=np.float64)
    rc = lib().orc_cg_device(C.byref(c), grid, op, _p(b), _p(x), max_iters, tol, C.byref(res), _p(h), hist)
    if rc:
        raise RuntimeError("orc_cg_device rc=%d" % rc)
    d = {f: getattr(res, f) for f, _ in CGResult._fields_}
    return x, d, h[: min(hist, res.iterations)]


def pcg_device(rp, ci, va, grid, op, b, x0, max_iters=1000, tol=1e-6):
    """Jacobi-preconditioned CG (oracle only convention, see oracle.c).  -> (x, result dict)"""
    rows = len(rp) - 1
    c = _csr_struct(rp, ci, va, rows, rows)
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.array(x0, dtype=np.float64, copy=True)
    res = CGResult()
    f = lib().orc_pcg_device
    f.argtypes = [C.POINTER(CSR), C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.POINTER(CGResult)]
    rc = f(C.byref(c), grid, op, _p(b), _p(x), max_iters, tol, C.byref(res))
    if rc:
        raise RuntimeError("orc_pcg_device rc=%d" % rc)
    return x, {k: getattr(res, k) for k, _ in CGResult._fields_}


def pcg_block_device(rp, ci, va, grid, op, P, b, x0, max_iters=1000, tol=1e-6):
    """block-Jacobi PCG, line blocks clipped to a P-band partition (oracle-only convention).  -> (x, result dict)"""
    rows = len(rp) - 1
    c = _csr_struct(rp, ci, va, rows, rows)
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.array(x0, dtype=np.float64, copy=True)
    res = CGResult()
    f = lib().orc_pcg_block_device
    f.argtypes = [C.POINTER(CSR), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.POINTER(CGResult)]
    rc = f(C.byref(c), grid, op, P, _p(b), _p(x), max_iters, tol, C.byref(res))
    if rc:
        raise RuntimeError("orc_pcg_block_device rc=%d" % rc)
    return x, {k: getattr(res, k) for k, _ in CGResult._fields_}


def cg_mgpu(rp, ci, va, grid, P, b, x0, max_iters=1000, tol=1e-6):
    rows = len(rp) - 1
    c = _csr_struct(rp, ci, va, rows, rows)
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.array(x0, dtype=np.float64, copy=True)
    res = CGResult()
    rc = lib().orc_cg_mgpu(C.byref(c), grid, P, _p(b), _p(x), max_iters, tol, C.byref(res))
    if rc:
        raise RuntimeError("orc_cg_mgpu rc=%d" % rc)
    return x, {f: getattr(res, f) for f, _ in CGResult._fields_}


# ---------------------------------------------------------------- partition
def partition(N, P, g):
    nl, off = C.c_int64(), C.c_int64()
    lib().orc_partition(N, P, g, C.byref(nl), C.byref(off))
    return nl.value, off.value


def local_csr_slice(rp, ci, va, row_offset, n_local):
    rows = len(rp) - 1
    c = _csr_struct(rp, ci, va, rows, rows)
    lnnz = int(rp[row_offset + n_local] - rp[row_offset])
    rpo = np.zeros(n_local + 1, dtype=np.int32)
    cio = np.zeros(max(lnnz, 1), dtype=np.int32)
    vao = np.zeros(max(lnnz, 1), dtype=np.float64)
    got = lib().orc_local_csr_slice(C.byref(c), row_offset, n_local, _p(rpo), _p(cio), _p(vao))
    assert got == lnnz
    return rpo, cio[:lnnz], vao[:lnnz]


def halo_ranges(n_local, grid, g, P):
    v = [C.c_int64() for _ in range(4)]
    lib().orc_halo_ranges(n_local, grid, g, P, *[C.byref(t) for t in v])
    return tuple(t.value for t in v)


def bench_stats(times):
    t = np.ascontiguousarray(times, dtype=np.float64)
    st = BenchStats()
    rc = lib().orc_bench_stats_from_times(_p(t), len(t), C.byref(st))
    return rc, {f: getattr(st, f) for f, _ in BenchStats._fields_}


def num_threads():
    return int(lib().orc_num_threads())


# ---------------------------------------------------------------- the reference's own host code
class RefHost:
    """oracle/_ref/libref_host.so: write_matrix_market_stencil5 / load_matrix_market (C linkage,
    include/io.h) and build_csr_struct (C++ linkage, include/spmv_csr.h:47) + global csr_mat."""

    def __init__(self):
        if not os.path.exists(REF_HOST_PATH):
            raise FileNotFoundError(REF_HOST_PATH)
        # DEEPBIND: the reference library must bind ITS OWN csr_mat / load_matrix_market / ... even if
        # the product library (same symbol names by design) is loaded in this process
        self.L = C.CDLL(REF_HOST_PATH, mode=os.RTLD_LOCAL | os.RTLD_DEEPBIND)
        self.L.write_matrix_market_stencil5.argtypes = [C.c_int, C.c_char_p]
        self.L.load_matrix_market.argtypes = [C.c_char_p, C.POINTER(Matrix)]
        self.build = getattr(self.L, "_Z16build_csr_structP10MatrixData")
        self.build.argtypes = [C.POINTER(Matrix)]
        self.csr = CSR.in_dll(self.L, "csr_mat")

    def write_stencil(self, n, path):
        return self.L.write_matrix_market_stencil5(n, path.encode())

    def load(self, path):
        m = Matrix()
        self.L.load_matrix_market(path.encode(), C.byref(m))
        ent = np.ctypeslib.as_array(C.cast(m.entries, C.POINTER(C.c_byte)), shape=(m.nnz * 16,)).copy().view(ENTRY_DTYPE)
        return m, ent

    def build_csr(self, m):
        # defeat the (rows, nnz) re-use guard (spmv_cusparse_csr.cu:64-69) between matrices
        self.csr.row_ptr = None
        rc = self.build(C.byref(m))
        assert rc == 0
        rows, nnz = m.rows, m.nnz
        rp = np.ctypeslib.as_array(self.csr.row_ptr, shape=(rows + 1,)).copy()
        ci = np.ctypeslib.as_array(self.csr.col_indices, shape=(nnz,)).copy()
        va = np.ctypeslib.as_array(self.csr.values, shape=(nnz,)).copy()
        return rp, ci, va


def ref_host_available():
    return os.path.exists(REF_HOST_PATH)
