"""CPU: the product's host layer (libspmv_b200.so) -- ABI exports, Matrix Market I/O, COO->CSR,
CSR->ELLPACK, bench statistics, exporters -- against the oracle and the golden fixtures.
No compute entry point is called here (no GPU in this tier)."""
import ctypes as C
import hashlib
import json
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

META = json.load(open(os.path.join(GOLDEN, "structure_meta.json")))


def test_library_exports_every_declared_symbol(B):
    L = B.load()
    missing = [s for s in B.C_SYMBOLS if not hasattr(L, s)]
    missing += [k for k, v in B.CXX_SYMBOLS.items() if not hasattr(L, v)]
    assert not missing, missing
    # every function declared in include/b200_kernels.h is bound
    hdr = open(os.path.join(ROOT, "include", "b200_kernels.h")).read()
    declared = set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", hdr))
    assert declared and declared <= set(B.C_SYMBOLS), declared - set(B.C_SYMBOLS)
    assert L.b200_version().startswith(b"b200-spmv-cg")


def test_struct_layouts_match_reference_headers(B):
    assert C.sizeof(B.Entry) == 16 and C.sizeof(B.MatrixData) == 24
    assert C.sizeof(B.CGConfig) == 24 and C.sizeof(B.CGStats) == 72 and C.sizeof(B.BenchmarkStats) == 48
    assert C.sizeof(B.SpmvOperator) == 40 and C.sizeof(B.CGStatsMultiGPU) == 144
    L = B.load()
    assert L.b200_cg_scalars_bytes() >= 64 and L.b200_xchg_bytes() % 8 == 0


@pytest.mark.parametrize("n", [2, 3, 4, 5, 7, 16])
def test_writer_reader_builder_vs_golden(B, orc, n, tmp_path):
    L = B.load()
    p = str(tmp_path / "s.mtx")
    assert L.write_matrix_market_stencil5(n, p.encode()) == 0
    assert hashlib.sha256(open(p, "rb").read()).hexdigest() == META["files"]["stencil_%d" % n]
    assert L.read_matrix_type(p.encode()) == 1
    hm = B.HostMatrix.from_mtx(p)
    assert (hm.md.rows, hm.md.cols, hm.md.nnz, hm.md.grid_size) == (n * n, n * n, 5 * n * n - 4 * n, n)
    ent = hm.entries_array()
    assert ent.tobytes() == orc.stencil5_entries(n).tobytes()
    L.csr_mat  # noqa: B018
    B.csr_mat().row_ptr = None  # force a rebuild (re-use guard keys on rows/nnz only)
    assert L.build_csr_struct(hm.ptr()) == 0
    rp, ci, va = B.host_csr_arrays()
    g = np.load(os.path.join(GOLDEN, "structure.npz"))
    assert np.array_equal(rp, g["n%d_row_ptr" % n]) and np.array_equal(ci, g["n%d_col" % n])
    assert np.array_equal(va, g["n%d_val" % n])


def test_builder_random_duplicates_and_ellpack(B, orc, tmp_path):
    L = B.load()
    rng = np.random.default_rng(5)
    rows, nnz = 50, 700
    ent = np.zeros(nnz, dtype=B.ENTRY_DTYPE)
    ent["row"], ent["col"] = rng.integers(0, rows, nnz), rng.integers(0, rows, nnz)
    ent["value"] = rng.uniform(-1, 1, nnz)
    hm = B.HostMatrix.from_entries(rows, rows, ent)
    B.csr_mat().row_ptr = None
    assert L.build_csr_struct(hm.ptr()) == 0
    rp, ci, va = B.host_csr_arrays()
    orp, oci, ova = orc.build_csr(rows, rows, ent)
    assert np.array_equal(rp, orp) and np.array_equal(ci, oci) and np.array_equal(va, ova)
    B.ellpack_matrix().indices = None
    assert L.ensure_ellpack_structure_built(hm.ptr()) == 0
    e = B.ellpack_matrix()
    w, oidx, oval = orc.build_ellpack(orp, oci, ova, rows, rows)
    assert e.ell_width == w and e.nb_rows == rows and e.nb_nonzeros == nnz
    idx = np.ctypeslib.as_array(e.indices, shape=(rows * w,))
    val = np.ctypeslib.as_array(e.values, shape=(rows * w,))
    assert np.array_equal(idx, oidx) and np.array_equal(val, oval)


def test_synthetic_matrix_host_csr_equals_oracle(B, orc):
    L = B.load()
    hs = B.HostMatrix.synthetic_stencil(12)
    assert hs.md.rows == 144 and hs.md.nnz == 5 * 144 - 48 and not hs.md.entries
    B.csr_mat().row_ptr = None
    assert L.build_csr_struct(hs.ptr()) == 0
    rp, ci, va = B.host_csr_arrays()
    rp64, oci, ova = orc.stencil5_csr_direct(12)
    assert np.array_equal(rp, rp64) and np.array_equal(ci, oci) and np.array_equal(va, ova)
    for r in (0, 1, 11, 12, 13, 77, 143, 144):
        assert L.b200_stencil5_nnz_before(r, 12) == rp64[r]


def test_symmetric_reader_expands(B, tmp_path):
    p = str(tmp_path / "sym.mtx")
    open(p, "w").write("%%MatrixMarket matrix coordinate real symmetric\n3 3 4\n1 1 2.0\n2 1 -1.0\n2 2 2.0\n3 2 -1.5\n")
    assert B.load().read_matrix_type(p.encode()) == 2
    hm = B.HostMatrix.from_mtx(p)
    e = hm.entries_array()
    assert hm.md.nnz == 6
    assert list(zip(e["row"], e["col"], e["value"])) == [(0, 0, 2.0), (1, 0, -1.0), (0, 1, -1.0), (1, 1, 2.0),
                                                          (2, 1, -1.5), (1, 2, -1.5)]


def test_reader_errors_are_reported(B, tmp_path):
    L = B.load()
    md = B.MatrixData()
    assert L.load_matrix_market(str(tmp_path / "missing.mtx").encode(), C.byref(md)) != 0
    p = str(tmp_path / "short.mtx")
    open(p, "w").write("%%MatrixMarket matrix coordinate real general\n2 2 3\n1 1 1.0\n")
    assert L.load_matrix_market(p.encode(), C.byref(md)) != 0 and not md.entries
    assert L.get_operator(b"does-not-exist") in (None,) or not L.get_operator(b"does-not-exist")
    for name in (b"cusparse-csr", b"csr", b"stencil5-csr", b"stencil5", b"ellpack", b"stencil5-ellpack",
                 b"stencil5-halo-mgpu"):
        assert L.get_operator(name)


def test_benchmark_with_stats_rule(B, orc):
    L = B.load()
    times = [10.0, 10.2, 9.9, 10.1, 30.0, 10.0, 9.8, 10.3, 10.1, 10.0]
    it = iter(times)

    @B.RUN_TIMED_FN
    def fake_run(x, y, ms):
        ms[0] = next(it)
        return 0

    st = B.BenchmarkStats()
    assert L.benchmark_with_stats(fake_run, None, None, len(times), C.byref(st)) == 0
    rc, ost = orc.bench_stats(times)
    got = B.stats_dict(st)
    for k, v in ost.items():
        assert got[k] == pytest.approx(v, rel=1e-15), k
    it = iter([1.0, 2.0])
    assert L.benchmark_with_stats(fake_run, None, None, 2, C.byref(st)) == -1


def test_cg_json_export_schema(B, tmp_path):
    """Key set / nesting / order of export_cg_json (reference cg_metrics.cu:20-81)."""
    L = B.load()
    hs = B.HostMatrix.synthetic_stencil(4)
    bs = B.BenchmarkStats(1.5, 1.6, 0.1, 1.4, 1.9, 9, 1)
    cs = B.CGStats(14, 1e-3, 1.5, 0.7, 0.5, 0.1, 1, 3.0, 2.0)
    p = str(tmp_path / "cg.json")
    L.export_cg_json(p.encode(), b"stencil5-csr", hs.ptr(), C.byref(bs), C.byref(cs))
    d = json.load(open(p))
    assert list(d.keys()) == ["timestamp", "solver", "mode", "matrix", "convergence", "timing", "statistics",
                              "performance", "validation"]
    assert d["solver"] == "CG" and d["convergence"] == {"converged": True, "iterations": 14, "residual_norm": 1e-3}
    assert list(d["timing"].keys()) == ["median_ms", "mean_ms", "min_ms", "max_ms", "std_dev_ms", "spmv_ms",
                                        "blas1_ms", "reductions_ms"]
    assert d["matrix"] == {"rows": 16, "cols": 16, "nnz": 64, "grid_size": 4}
    ms = B.CGStatsMultiGPU()
    ms.iterations, ms.time_spmv_ms, ms.converged = 14, 0.5, 1
    p2 = str(tmp_path / "mg.json")
    L.export_cg_mgpu_json(p2.encode(), b"partitioned-halo", hs.ptr(), C.byref(bs), C.byref(ms), 8)
    d2 = json.load(open(p2))
    assert d2["num_gpus"] == 8 and d2["solver"] == "CG Multi-GPU" and "allgather_ms" in d2["timing"]
    p3 = str(tmp_path / "cg.csv")
    L.export_cg_csv(p3.encode(), b"stencil5-csr", hs.ptr(), C.byref(bs), C.byref(cs), True)
    lines = open(p3).read().splitlines()
    assert lines[0].startswith("mode,rows,cols,nnz,grid_size,converged,iterations") and lines[1].startswith("stencil5-csr,16,16,64,4,1,14,")


def test_mpirun_shim_maps_ranks_to_gpus():
    """reference scripts call `mpirun -np P ./bin/cg_solver_mgpu_stencil <mtx> --json=F`
    (scripts/benchmarking/benchmark_weak_scaling.sh, benchmark_problem_sizes.sh): the shim must turn
    that into one process with --gpus=P and drop the MPI-only options"""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    shim = os.path.join(root, "cuda-spmv-benchmark_b200", "scripts", "mpirun")
    env = dict(os.environ, B200_MPIRUN_DRYRUN="1")

    def run(*args):
        return subprocess.run([shim, *args], env=env, capture_output=True, text=True)
    r = run("-np", "8", "--allow-run-as-root", "--bind-to", "none", "--mca", "btl", "self,vader",
            "./bin/cg_solver_mgpu_stencil", "matrix/20000", "--json=out.json", "--timers")
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip() == "./bin/cg_solver_mgpu_stencil matrix/20000 --json=out.json --timers --gpus=8"
    r = run("-n", "2", "./bin/cg_solver_mgpu_stencil", "m.mtx")
    assert r.stdout.strip() == "./bin/cg_solver_mgpu_stencil m.mtx --gpus=2"
    r = run("-np", "1", "./bin/cg_solver", "m.mtx", "--mode=stencil5-csr")
    assert r.stdout.strip() == "./bin/cg_solver m.mtx --mode=stencil5-csr"
    assert run("-np", "4").returncode == 2


def test_bench_reference_arm_json_contract():
    """`bench.py --impl reference` (CPU arm, runs anywhere): one JSON line with the contract's keys"""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-grid", "200"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["unit"] == "ms" and line["higher_is_better"] is False
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    # the arm reports the grid it actually ran -- never a scaled number under the 20k label
    assert line["value"] > 0 and line["config"]["workload"].startswith("cg_200x200_")
    assert line["config"]["grid"] == 200 and line["config"]["rows"] == 40000
    assert line["config"]["requested_grid"] == 20000
    assert "REDUCED from 20000x20000" in line["cpu_baseline"]["sample"]
    assert "scaled" not in line["cpu_baseline"]["sample"].replace("not scaled", "")
    # the oracle's own KAT for this grid (tests/test_oracle_golden.py pins the oracle itself)
    assert line["cg"]["iterations"] == line["config"]["iterations"] > 0


def test_synthetic_matrix_beyond_32_bit_rows(B):
    """the synthetic stencil is defined by grid_size; the reference's 32-bit rows / cols / nnz fields
    saturate instead of wrapping (weak scaling: 56576^2 = 3.2e9 rows over 8 GPUs), and the closed-form
    non-zero prefix stays exact in 64 bits"""
    L = B.load()
    m = L.b200_synthetic_stencil(56576)
    assert m.grid_size == 56576 and m.rows == 2**31 - 1 and m.cols == 2**31 - 1 and m.nnz == 2**31 - 1
    assert not m.entries
    m = L.b200_synthetic_stencil(20000)
    assert m.rows == 400_000_000 and m.nnz == 1_999_920_000
    n = 56576
    N = n * n
    assert L.b200_stencil5_nnz_before(N, n) == 5 * N - 4 * n
    assert L.b200_stencil5_nnz_before(n, n) == 4 * n - 2  # first grid row: two corners of 3, n - 2 rows of 4
    # per-band non-zeros of the 8-GPU weak-scaling case stay below 2^31 (32-bit local offsets)
    for r in range(8):
        lo, hi = r * (N // 8), (r + 1) * (N // 8)
        assert 0 < L.b200_stencil5_nnz_before(hi, n) - L.b200_stencil5_nnz_before(lo, n) < 2**31


def test_csr_cache_is_keyed_on_the_matrix_not_only_its_shape(B, orc):
    """the reference re-uses its global csr_mat whenever (rows, nnz) match (spmv_cusparse_csr.cu:64-69);
    here a second matrix of the same shape must get its own values, while the SAME matrix is not rebuilt.
    Entries outside the matrix are rejected instead of corrupting the heap."""
    import numpy as np
    L = B.load()
    n = 12
    N = n * n
    e1 = orc.stencil5_entries(n, 5.0, -1.0)
    e2 = e1.copy()
    e2["value"] = np.where(e2["row"] == e2["col"], 7.5, -2.0)
    h1 = B.HostMatrix.from_entries(N, N, e1, grid_size=n)
    h2 = B.HostMatrix.from_entries(N, N, e2, grid_size=n)
    assert L.build_csr_struct(h1.ptr()) == 0
    _, _, va1 = B.host_csr_arrays()
    ptr1 = C.cast(B.csr_mat().values, C.c_void_p).value
    assert L.build_csr_struct(h1.ptr()) == 0  # same matrix: re-used, not rebuilt
    assert C.cast(B.csr_mat().values, C.c_void_p).value == ptr1
    assert L.build_csr_struct(h2.ptr()) == 0  # same shape, other values: rebuilt
    rp2, ci2, va2 = B.host_csr_arrays()
    orp, oci, ova = orc.build_csr(N, N, e2)
    assert np.array_equal(va2, ova) and np.array_equal(ci2, oci) and not np.array_equal(va1, va2)
    bad = e1.copy()
    bad["col"][5] = N + 3
    hb = B.HostMatrix.from_entries(N, N, bad, grid_size=n)
    assert L.build_csr_struct(hb.ptr()) != 0


def test_host_zero_guess_scan(B):
    """the host-side scan behind the zero-initial-guess path (host/cg_engine.cpp: host_all_zero): a wrong 'all
    zero' would silently drop the caller's initial guess, so every position class is probed -- the serial head,
    the first / last element of every thread's range, block boundaries, the very last element -- and -0.0 and
    denormals must count as non-zero (bit patterns are compared: the cleared device vector must be identical)."""
    import numpy as np
    L = B.load()
    rng = np.random.default_rng(5)
    for n in (0, 1, 2, 4095, 4096, 4097, 131072 + 4096, 131072 + 4097, 1_000_003, 3_500_000):
        for threads in (1, 3, 16):
            x = np.zeros(max(n, 1))
            assert L.b200_host_all_zero(x.ctypes.data, n, threads) == 1, (n, threads)
            if n == 0:
                continue
            probes = {0, n - 1, n // 2, min(n - 1, 4095), min(n - 1, 4096), min(n - 1, 4096 + 131072 - 1), min(n - 1, 4096 + 131072)}
            probes |= {int(v) for v in rng.integers(0, n, size=6)}
            for nt in (threads,):
                per = (max(n - 4096, 0) + nt - 1) // nt if n > 4096 else 0
                for t in range(nt):
                    if per:
                        probes |= {min(n - 1, 4096 + t * per), min(n - 1, 4096 + (t + 1) * per - 1)}
            for i in sorted(probes):
                for v in (1.0, -0.0, 5e-324, float("nan")):
                    x[i] = v
                    assert L.b200_host_all_zero(x.ctypes.data, n, threads) == 0, (n, threads, i, v)
                    x[i] = 0.0
            assert L.b200_host_all_zero(x.ctypes.data, n, threads) == 1
