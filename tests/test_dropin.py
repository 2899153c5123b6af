"""The drop-in claim as a regression test.

CPU tier (runs where /root/reference is mounted, i.e. the build container): the reference's
UNMODIFIED main programs -- src/main/main.cu, src/main/cg_solver.cu and the MPI program
src/main/cg_solver_mgpu_stencil.cu -- compile against this repo's include/ and link with
libspmv_b200.so (oracle/Makefile target `dropin`); struct layouts are compared with the
reference's OWN headers by compiling the same probe against both include trees.

GPU tier: the binaries built that way (oracle/_ref/dropin, shipped to the GPU box like every other
prebuilt artefact) run on a small matrix and print the answers the oracle computes."""
import json
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
PKG = os.path.join(ROOT, "cuda-spmv-benchmark_b200")
DROPIN = os.path.join(ROOT, "oracle", "_ref", "dropin")
have_ref = os.path.exists(os.path.join(REF, "src", "main", "main.cu"))

PROBE = r"""
#include <stddef.h>
#include <stdio.h>
#include "spmv.h"
#include "io.h"
#include "spmv_csr.h"
#include "spmv_ellpack.h"
#include "benchmark_stats.h"
#include "solvers/cg_solver.h"
#include "solvers/cg_solver_mgpu.h"
#define S(T) printf("\"sizeof(" #T ")\": %zu,\n", sizeof(T))
#define O(T, f) printf("\"offsetof(" #T "," #f ")\": %zu,\n", offsetof(T, f))
int main() {
    printf("{\n");
    S(Entry); O(Entry, row); O(Entry, col); O(Entry, value);
    S(MatrixData); O(MatrixData, rows); O(MatrixData, cols); O(MatrixData, nnz); O(MatrixData, grid_size); O(MatrixData, entries);
    S(CSRMatrix); O(CSRMatrix, nb_rows); O(CSRMatrix, nb_nonzeros); O(CSRMatrix, row_ptr); O(CSRMatrix, col_indices); O(CSRMatrix, values);
    S(ELLPACKMatrix); O(ELLPACKMatrix, ell_width); O(ELLPACKMatrix, indices); O(ELLPACKMatrix, values);
    S(SpmvOperator); O(SpmvOperator, name); O(SpmvOperator, init); O(SpmvOperator, run_timed); O(SpmvOperator, run_device); O(SpmvOperator, free);
    S(BenchmarkStats); O(BenchmarkStats, median_ms); O(BenchmarkStats, valid_runs); O(BenchmarkStats, outliers_removed);
    S(CGConfig); O(CGConfig, max_iters); O(CGConfig, tolerance); O(CGConfig, verbose); O(CGConfig, enable_detailed_timers);
    S(CGStats); O(CGStats, iterations); O(CGStats, residual_norm); O(CGStats, time_total_ms); O(CGStats, converged); O(CGStats, solution_sum); O(CGStats, solution_norm);
    S(CGConfigMultiGPU); S(CGStatsMultiGPU); O(CGStatsMultiGPU, time_allreduce_ms); O(CGStatsMultiGPU, time_allgather_ms);
    O(CGStatsMultiGPU, converged); O(CGStatsMultiGPU, time_initial_r_ms); O(CGStatsMultiGPU, solution_sum); O(CGStatsMultiGPU, solution_norm);
    printf("\"end\": 0}\n");
    return 0;
}
"""


def _layout(inc_root, tmp_path, tag):
    src = tmp_path / ("probe_%s.cu" % tag)
    exe = tmp_path / ("probe_%s" % tag)
    src.write_text(PROBE)
    subprocess.run(["nvcc", "-w", "-std=c++17", "-I" + inc_root, "-I" + os.path.join(inc_root, "solvers"), "-o", str(exe),
                    str(src)], check=True, capture_output=True, timeout=600)
    return json.loads(subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout)


@pytest.mark.skipif(not have_ref, reason="reference tree not mounted")
def test_struct_layouts_equal_the_reference_headers(B, tmp_path):
    """sizes and field offsets of every struct on the boundary, taken from the reference's own headers
    and from include/ by the same probe -- and the ctypes mirrors the tests use agree with both"""
    import ctypes as C
    ref = _layout(os.path.join(REF, "include"), tmp_path, "ref")
    ours = _layout(os.path.join(ROOT, "include"), tmp_path, "ours")
    assert ref == ours
    for name, ct in (("Entry", B.Entry), ("MatrixData", B.MatrixData), ("SpmvOperator", B.SpmvOperator),
                     ("BenchmarkStats", B.BenchmarkStats), ("CGConfig", B.CGConfig), ("CGStats", B.CGStats),
                     ("CGStatsMultiGPU", B.CGStatsMultiGPU), ("CSRMatrix", B.CSRMatrix), ("ELLPACKMatrix", B.ELLPACKMatrix)):
        assert C.sizeof(ct) == ref["sizeof(%s)" % name], name


@pytest.mark.skipif(not have_ref, reason="reference tree not mounted")
def test_reference_mains_compile_and_link_against_this_library():
    """src/main/{main,cg_solver,cg_solver_mgpu_stencil}.cu, unmodified, against include/ + -lspmv_b200
    (the MPI program against the single-process stand-in compat/mpi.h)"""
    subprocess.run(["make", "-C", PKG, "-s", "lib"], check=True, timeout=1800)
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-B", "dropin"], capture_output=True, text=True, timeout=1800)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    for exe in ("spmv_bench", "cg_solver", "cg_solver_mgpu_stencil"):
        p = os.path.join(DROPIN, exe)
        assert os.path.exists(p)
        needed = subprocess.run(["readelf", "-d", p], capture_output=True, text=True).stdout
        assert "libspmv_b200.so" in needed and "cusparse" not in needed and "cublas" not in needed and "libmpi" not in needed


def _num(pattern, text):
    m = re.search(pattern, text)
    assert m, (pattern, text[-1500:])
    return float(m.group(1))


@pytest.mark.gpu
def test_reference_mains_run_on_this_library(B, orc, torch_cuda, tmp_path):
    """the reference's own CLIs (unmodified mains) on top of libspmv_b200.so: 512 x 512 stencil .mtx,
    checksums and iteration counts as the oracle computes them; JSON written by the reference-facing
    exporters of this library"""
    for exe in ("spmv_bench", "cg_solver", "cg_solver_mgpu_stencil"):
        if not os.path.exists(os.path.join(DROPIN, exe)):
            pytest.skip("oracle/_ref/dropin not built (reference tree was not present at build time)")
    n = 512
    N = n * n
    mtx = str(tmp_path / "stencil_512x512.mtx")
    assert B.load().write_matrix_market_stencil5(n, mtx.encode()) == 0
    rp64, ci, va = orc.stencil5_csr_direct(n)
    rp = rp64.astype(np.int32)
    y = orc.stencil5_spmv(rp, ci, va, np.ones(N), n)
    xo, ro, _ = orc.cg_device(rp, ci, va, n, 1, np.ones(N), np.zeros(N))
    env = dict(os.environ, B200_GPUS="1")

    out = subprocess.run([os.path.join(DROPIN, "spmv_bench"), mtx, "--mode=stencil5-csr,cusparse-csr",
                          "--json=" + str(tmp_path / "spmv.json")], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    sums = [float(v) for v in re.findall(r"Sum\(y\):\s+([-+0-9.eE]+)", out.stdout)]
    assert sums == [float(y.sum())] * 2 == [float(N + 4 * n)] * 2
    assert json.load(open(tmp_path / "spmv_stencil5-csr.json"))

    # the single-GPU main keeps x between its warm-up solves and the measured ones (reference quirk,
    # src/main/cg_solver.cu:155-173): only the solution is checked, not the (warm restart) iteration count
    out = subprocess.run([os.path.join(DROPIN, "cg_solver"), mtx, "--mode=stencil5-csr"], capture_output=True, text=True,
                         timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "Converged: YES" in out.stdout
    assert abs(_num(r"Sum\(x\):\s+([-+0-9.eE]+)", out.stdout) - ro["solution_sum"]) <= 1e-5 * abs(ro["solution_sum"])

    # the MPI main resets x before every solve: iteration count and checksums are the oracle's
    js = tmp_path / "mgpu.json"
    out = subprocess.run([os.path.join(PKG, "scripts", "mpirun"), "-np", "1", "--allow-run-as-root",
                          os.path.join(DROPIN, "cg_solver_mgpu_stencil"), mtx, "--json=" + str(js)],
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "Converged: YES in %d iterations" % ro["iterations"] in out.stdout
    assert abs(_num(r"Sum\(x\):\s+([-+0-9.eE]+)", out.stdout) - ro["solution_sum"]) <= 1e-9 * abs(ro["solution_sum"])
    assert abs(_num(r"Norm2\(x\):\s+([-+0-9.eE]+)", out.stdout) - ro["solution_norm"]) <= 1e-9 * ro["solution_norm"]
    assert json.load(open(js))
