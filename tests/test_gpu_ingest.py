"""GPU parity: device-side Matrix Market parsing and COO->CSR against the oracle's
restatement of the reference reader (src/io/io.cu:109-171) and builder
(src/spmv/spmv_cusparse_csr.cu:85-157).  Bar: bit-exact (indices, values, order)."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def dptr(t):
    return C.c_void_p(t.data_ptr())


def device_csr(B, torch, ent, rows):
    L = B.load()
    nnz = len(ent)
    d_ent = torch.from_numpy(np.frombuffer(np.ascontiguousarray(ent).tobytes(), dtype=np.uint8).copy()).cuda() \
        if nnz else torch.zeros(16, dtype=torch.uint8, device="cuda")
    rp = torch.full((rows + 1,), -7, dtype=torch.int32, device="cuda")
    ci = torch.full((max(nnz, 1),), -7, dtype=torch.int32, device="cuda")
    va = torch.zeros(max(nnz, 1), dtype=torch.float64, device="cuda")
    B.check(L.b200_coo_to_csr(dptr(d_ent), nnz, rows, 0, dptr(rp), dptr(ci), dptr(va), None), "coo_to_csr")
    return rp.cpu().numpy(), ci.cpu().numpy()[:nnz], va.cpu().numpy()[:nnz]


@pytest.mark.parametrize("kind", ["stencil", "shuffled_dups", "long_rows", "empty_rows", "single_row", "empty"])
def test_coo_to_csr_bit_exact(B, orc, torch_cuda, kind):
    rng = np.random.default_rng(42)
    if kind == "stencil":
        rows = 33 * 33
        ent = orc.stencil5_entries(33)
    elif kind == "shuffled_dups":
        rows, nnz = 500, 9000  # many duplicate (row, col) pairs, random file order
        ent = np.zeros(nnz, dtype=orc.ENTRY_DTYPE)
        ent["row"], ent["col"] = rng.integers(0, rows, nnz), rng.integers(0, 40, nnz)
        ent["value"] = rng.standard_normal(nnz)
    elif kind == "long_rows":
        rows, nnz = 70, 30000  # rows of ~430 entries: the warp rank-sort path
        ent = np.zeros(nnz, dtype=orc.ENTRY_DTYPE)
        ent["row"], ent["col"] = rng.integers(0, rows, nnz), rng.integers(0, 300, nnz)
        ent["value"] = rng.standard_normal(nnz)
    elif kind == "empty_rows":
        rows, nnz = 4000, 3000
        ent = np.zeros(nnz, dtype=orc.ENTRY_DTYPE)
        ent["row"], ent["col"] = rng.integers(0, rows, nnz) // 7 * 7, rng.integers(0, rows, nnz)
        ent["value"] = rng.standard_normal(nnz)
    elif kind == "single_row":
        rows, nnz = 1, 100
        ent = np.zeros(nnz, dtype=orc.ENTRY_DTYPE)
        ent["col"], ent["value"] = rng.integers(0, 50, nnz), rng.standard_normal(nnz)
    else:
        rows = 10
        ent = np.zeros(0, dtype=orc.ENTRY_DTYPE)
    rp, ci, va = device_csr(B, torch_cuda, ent, rows)
    orp, oci, ova = orc.build_csr(rows, max(rows, 300), ent)
    assert np.array_equal(rp, orp) and np.array_equal(ci, oci) and np.array_equal(va, ova)


def test_coo_to_csr_rejects_bad_rows(B, orc, torch_cuda):
    ent = np.zeros(3, dtype=orc.ENTRY_DTYPE)
    ent["row"] = [0, 5, 1]
    L = B.load()
    d_ent = torch_cuda.from_numpy(np.frombuffer(ent.tobytes(), dtype=np.uint8).copy()).cuda()
    rp = torch_cuda.zeros(4, dtype=torch_cuda.int32, device="cuda")
    ci = torch_cuda.zeros(3, dtype=torch_cuda.int32, device="cuda")
    va = torch_cuda.zeros(3, dtype=torch_cuda.float64, device="cuda")
    assert L.b200_coo_to_csr(dptr(d_ent), 3, 3, 3, dptr(rp), dptr(ci), dptr(va), None) == 1
    assert b"row index" in L.b200_last_error()
    ent["row"] = [0, 2, 1]
    ent["col"] = [0, 3, 1]  # column 3 of a 3-column matrix
    d_ent = torch_cuda.from_numpy(np.frombuffer(ent.tobytes(), dtype=np.uint8).copy()).cuda()
    assert L.b200_coo_to_csr(dptr(d_ent), 3, 3, 3, dptr(rp), dptr(ci), dptr(va), None) == 1
    assert b"column index" in L.b200_last_error()


def load_device(B, torch, path):
    L = B.load()
    md = B.MatrixData()
    d_ent = C.c_void_p()
    rc = L.b200_load_matrix_market_device(path.encode(), C.byref(md), C.byref(d_ent))
    if rc != 0:
        return rc, md, None
    n = md.nnz
    host = np.zeros(max(n, 1), dtype=B.ENTRY_DTYPE)
    if n:
        assert L.b200_copy_to_host(host.ctypes.data, d_ent, 16 * n) == 0
    return 0, md, (host[:n], d_ent)


def write_mtx(path, rows, cols, triples, fmt="%.17g", header_extra="", symmetric=False, sep="\n"):
    with open(path, "w") as f:
        f.write("%%%%MatrixMarket matrix coordinate real %s\n" % ("symmetric" if symmetric else "general"))
        f.write(header_extra)
        f.write("%d %d %d\n" % (rows, cols, len(triples)))
        for r, c, v in triples:
            f.write(("%d %d " + fmt + sep) % (r + 1, c + 1, v))


@pytest.mark.parametrize("fmt", ["%.17g", "%.6e", "%g", "%.1f", "%.20e"])
def test_mtx_parse_matches_reader(B, orc, torch_cuda, tmp_path, fmt):
    """random values in several notations; %.17g / %.20e exceed the exact fast path -> strtod fix-up"""
    rng = np.random.default_rng(7)
    rows, nnz = 300, 5000
    tr = [(int(rng.integers(0, rows)), int(rng.integers(0, rows)), float(rng.standard_normal() * 10.0 ** int(rng.integers(-8, 8))))
          for _ in range(nnz)]
    p = str(tmp_path / "m.mtx")
    write_mtx(p, rows, rows, tr, fmt=fmt, header_extra="% a comment\n% STENCIL_GRID_SIZE 17\n")
    rc, md, got = load_device(B, torch_cuda, p)
    assert rc == 0
    orows, ocols, onnz, ogrid, oent = orc.load_mtx(p)
    assert (md.rows, md.cols, md.nnz, md.grid_size) == (orows, ocols, onnz, ogrid) == (rows, rows, nnz, 17)
    assert got[0].tobytes() == oent.tobytes()
    B.load().b200_free_device(got[1])


def test_mtx_parse_generator_file_and_operator(B, orc, torch_cuda, tmp_path):
    """generator file -> GPU parse -> GPU COO->CSR -> operator, no host Entry[] / CSR at any point"""
    L = B.load()
    n = 120
    p = str(tmp_path / "s.mtx")
    orc.write_mtx_stencil5(n, p)
    rc, md, got = load_device(B, torch_cuda, p)
    assert rc == 0 and md.grid_size == n and md.nnz == 5 * n * n - 4 * n
    assert got[0].tobytes() == orc.stencil5_entries(n).tobytes()
    rng = np.random.default_rng(3)
    x = rng.standard_normal(n * n)
    orp64, oci, ova = orc.stencil5_csr_direct(n)
    expect = {b"stencil5-csr": orc.stencil5_spmv(orp64.astype(np.int32), oci, ova, x, n),  # C,W,E,N,S order
              b"cusparse-csr": orc.csr_spmv(orp64.astype(np.int32), oci, ova, x)}          # k order
    for name in (b"stencil5-csr", b"cusparse-csr"):
        op = L.get_operator(name)
        assert L.b200_operator_init_device_coo(op, C.byref(md), got[1]) == 0
        y = np.full(n * n, np.nan)
        ms = C.c_double()
        assert op.contents.run_timed(x.ctypes.data, y.ctypes.data, C.byref(ms)) == 0
        assert np.array_equal(y, expect[name]), name
        op.contents.free()
    L.b200_free_device(got[1])


def test_mtx_parse_fallbacks_and_errors(B, orc, torch_cuda, tmp_path):
    L = B.load()
    # two entries per line: token-based like fscanf -> host reader path, same result
    p = str(tmp_path / "two.mtx")
    open(p, "w").write("%%MatrixMarket matrix coordinate real general\n3 3 4\n1 1 2.5 2 2 -1\n3 1 7\n1 3 1e-3\n")
    rc, md, got = load_device(B, torch_cuda, p)
    orows, ocols, onnz, ogrid, oent = orc.load_mtx(p)
    assert rc == 0 and md.nnz == 4 and got[0].tobytes() == oent.tobytes()
    L.b200_free_device(got[1])
    # blank lines, CRLF, trailing blanks
    p = str(tmp_path / "blank.mtx")
    open(p, "w").write("%%MatrixMarket matrix coordinate real general\n2 2 3\n\n1 1 5.0 \r\n  2 2 -1.0\n\n2 1 3\n\n")
    rc, md, got = load_device(B, torch_cuda, p)
    orows, ocols, onnz, ogrid, oent = orc.load_mtx(p)
    assert rc == 0 and got[0].tobytes() == oent.tobytes()
    L.b200_free_device(got[1])
    # symmetric: expanded (mirror after each off-diagonal entry)
    p = str(tmp_path / "sym.mtx")
    write_mtx(p, 3, 3, [(0, 0, 2.0), (1, 0, -1.0), (2, 1, -1.5)], symmetric=True)
    rc, md, got = load_device(B, torch_cuda, p)
    assert rc == 0 and md.nnz == 5
    assert list(zip(got[0]["row"], got[0]["col"])) == [(0, 0), (1, 0), (0, 1), (2, 1), (1, 2)]
    L.b200_free_device(got[1])
    # truncated and missing files are errors
    p = str(tmp_path / "short.mtx")
    open(p, "w").write("%%MatrixMarket matrix coordinate real general\n2 2 3\n1 1 1.0\n")
    assert load_device(B, torch_cuda, p)[0] != 0
    assert load_device(B, torch_cuda, str(tmp_path / "nope.mtx"))[0] != 0


def test_cli_device_ingest_flag(B, orc, torch_cuda, tmp_path):
    """`spmv_bench` / `cg_solver --device-ingest`: the .mtx is parsed and turned into CSR on the GPU (no host
    Entry[] / CSR); answers equal the host-reader path -- bundled 81 x 81 matrix (centre -4): Sum(y) = -52164,
    CG 40 iterations, Sum(x) = -826.0838884."""
    import os
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    bindir = os.path.join(root, "cuda-spmv-benchmark_b200", "bin")
    mtx = str(tmp_path / "example81x81.mtx")
    orc.write_mtx_stencil5(81, mtx, "-4.0", "-1.0")
    outs = {}
    for flag in ([], ["--device-ingest"]):
        r = subprocess.run([os.path.join(bindir, "spmv_bench"), mtx, "--mode=stencil5-csr,cusparse-csr"] + flag,
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        sums = [float(v) for v in re.findall(r"Sum\(y\):\s+([-+0-9.eE]+)", r.stdout)]
        assert sums == [-52164.0, -52164.0]
        r = subprocess.run([os.path.join(bindir, "cg_solver"), mtx, "--mode=stencil5-csr"] + flag,
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        assert "Converged: YES in 40 iterations" in r.stdout
        outs[bool(flag)] = re.search(r"Sum\(x\):\s+([-+0-9.eE]+)", r.stdout).group(1)
        if flag:
            assert "parsed on the device" in r.stdout
    assert outs[False] == outs[True] and abs(float(outs[True]) + 826.0838884) < 1e-6
    r = subprocess.run([os.path.join(bindir, "spmv_bench"), mtx, "--mode=ellpack", "--device-ingest"], capture_output=True, text=True)
    assert r.returncode != 0 and "supports the cusparse-csr and stencil5-csr" in r.stderr
