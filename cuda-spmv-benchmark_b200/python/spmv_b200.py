"""ctypes binding of libspmv_b200.so -- the stub a maintainer of a Python caller would write.

It binds exactly the symbols declared in include/b200/api.h (the reference's host API) and
include/b200_kernels.h (the thin kernel C ABI).  No compute happens in Python and there is no
fallback: if the shared library is missing, `load()` raises; if there is no sm_100 GPU, every
launcher returns B200_ENODEV / a CUDA error, which `check()` turns into an exception.
"""
import ctypes as C
import os
import subprocess

import numpy as np

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("B200_LIB_PATH") or os.path.join(PKG_DIR, "libspmv_b200.so")  # B200_LIB_PATH: A/B builds of the same sources
REPO_ROOT = os.path.dirname(PKG_DIR)


# ------------------------------------------------------------------ structs (include/b200/types.h)
class Entry(C.Structure):
    _fields_ = [("row", C.c_int), ("col", C.c_int), ("value", C.c_double)]


ENTRY_DTYPE = np.dtype([("row", np.int32), ("col", np.int32), ("value", np.float64)], align=True)


class MatrixData(C.Structure):
    _fields_ = [("rows", C.c_int), ("cols", C.c_int), ("nnz", C.c_int), ("grid_size", C.c_int),
                ("entries", C.POINTER(Entry))]


class CSRMatrix(C.Structure):
    _fields_ = [("nb_rows", C.c_int), ("nb_cols", C.c_int), ("nb_nonzeros", C.c_int),
                ("row_ptr", C.POINTER(C.c_int)), ("col_indices", C.POINTER(C.c_int)),
                ("values", C.POINTER(C.c_double))]


class ELLPACKMatrix(C.Structure):
    _fields_ = [("nb_rows", C.c_int), ("nb_cols", C.c_int), ("ell_width", C.c_int), ("grid_size", C.c_int),
                ("indices", C.POINTER(C.c_int)), ("nb_nonzeros", C.c_int), ("values", C.POINTER(C.c_double))]


INIT_FN = C.CFUNCTYPE(C.c_int, C.POINTER(MatrixData))
RUN_TIMED_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_double))
RUN_DEVICE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p)
FREE_FN = C.CFUNCTYPE(None)


class SpmvOperator(C.Structure):
    _fields_ = [("name", C.c_char_p), ("init", INIT_FN), ("run_timed", RUN_TIMED_FN),
                ("run_device", RUN_DEVICE_FN), ("free", FREE_FN)]


class CGConfig(C.Structure):
    _fields_ = [("max_iters", C.c_int), ("tolerance", C.c_double), ("verbose", C.c_int),
                ("enable_detailed_timers", C.c_int)]


class CGStats(C.Structure):
    _fields_ = [("iterations", C.c_int), ("residual_norm", C.c_double), ("time_total_ms", C.c_double),
                ("time_spmv_ms", C.c_double), ("time_blas1_ms", C.c_double), ("time_reductions_ms", C.c_double),
                ("converged", C.c_int), ("solution_sum", C.c_double), ("solution_norm", C.c_double)]


class CGStatsMultiGPU(C.Structure):
    _fields_ = [("iterations", C.c_int), ("residual_norm", C.c_double), ("time_total_ms", C.c_double),
                ("time_spmv_ms", C.c_double), ("time_blas1_ms", C.c_double), ("time_reductions_ms", C.c_double),
                ("time_allreduce_ms", C.c_double), ("time_allgather_ms", C.c_double), ("converged", C.c_int),
                ("time_dot_rs_initial_ms", C.c_double), ("time_dot_pAp_ms", C.c_double),
                ("time_dot_rs_new_ms", C.c_double), ("time_axpy_update_x_ms", C.c_double),
                ("time_axpy_update_r_ms", C.c_double), ("time_axpby_update_p_ms", C.c_double),
                ("time_initial_r_ms", C.c_double), ("solution_sum", C.c_double), ("solution_norm", C.c_double)]


class BenchmarkStats(C.Structure):
    _fields_ = [("median_ms", C.c_double), ("mean_ms", C.c_double), ("std_dev_ms", C.c_double),
                ("min_ms", C.c_double), ("max_ms", C.c_double), ("valid_runs", C.c_int),
                ("outliers_removed", C.c_int)]


class Band(C.Structure):  # b200_band, include/b200_kernels.h
    _fields_ = [("d_row_ptr", C.c_void_p), ("d_col_idx", C.c_void_p), ("d_values", C.c_void_p),
                ("values_len", C.c_longlong), ("row_offset", C.c_longlong), ("n_local", C.c_longlong),
                ("grid_size", C.c_int), ("layout", C.c_int), ("d_halo_prev", C.c_void_p),
                ("d_halo_next", C.c_void_p), ("d_flag_prev", C.c_void_p), ("d_flag_next", C.c_void_p),
                ("epoch", C.c_uint32), ("rows_per_item", C.c_int), ("variant", C.c_int), ("d_epoch_ptr", C.c_void_p)]


class ReduceCtx(C.Structure):  # b200_reduce_ctx
    _fields_ = [("d_scalars", C.c_void_p), ("h_status_mapped", C.c_void_p), ("d_partials", C.c_void_p),
                ("d_partials_b", C.c_void_p), ("d_group_sums", C.c_void_p), ("d_tickets", C.c_void_p),
                ("capacity", C.c_longlong), ("d_stash", C.c_void_p), ("d_out", C.c_void_p), ("rank", C.c_int),
                ("world", C.c_int), ("d_peer_xchg", C.c_void_p), ("tol", C.c_double), ("phases", C.c_int)]


class HaloPushArgs(C.Structure):  # b200_halo_push_args
    _fields_ = [("d_dst_prev", C.c_void_p), ("d_dst_next", C.c_void_p), ("d_flag_prev", C.c_void_p),
                ("d_flag_next", C.c_void_p), ("d_my_xchg", C.c_void_p), ("halo", C.c_int)]


class TailTimes(C.Structure):  # b200_cg_tail_times
    _fields_ = [("ns", C.c_ulonglong * 8), ("count", C.c_uint * 8), ("gap_ns", C.c_ulonglong * 8), ("error", C.c_int)]


class CsrPlan(C.Structure):
    _fields_ = [("rows_per_block", C.c_int), ("window", C.c_int), ("vector_threshold", C.c_int),
                ("variant", C.c_int), ("hist", C.c_ulonglong * 33),
                ("max_row_len", C.c_ulonglong), ("mean_row_len", C.c_double)]


# every symbol the two public headers declare: (name, mangled-or-None)
C_SYMBOLS = [
    # include/b200_kernels.h
    "b200_version", "b200_last_error", "b200_launch_count", "b200_stencil5_spmv", "b200_spmv_stencil5_csr",
    "b200_spmv_stencil5_halo", "b200_spmv_stencil5_ellpack", "b200_stencil5_num_partials",
    "b200_stencil5_variant_info", "b200_stencil5_set_plain_variant", "b200_cg_set_kernel", "b200_cg_get_kernel", "b200_csr_variant_info", "b200_csr_set_default_variant", "b200_csr_plan_build", "b200_spmv_csr", "b200_spmv_ellpack",
    "b200_cg_scalars_bytes", "b200_cg_status_bytes", "b200_xchg_bytes", "b200_cg_max_partials",
    "b200_cg_residual_init", "b200_cg_spmv_dot", "b200_cg_update_xr", "b200_cg_update_p", "b200_cg_reduce",
    "b200_cg_update_p_push", "b200_cg_spmv_fused", "b200_cg_update_r", "b200_cg_halo_dir",
    "b200_cg_finish_x", "b200_cg_spmv_fused_nx", "b200_cg_finish_x_depth", "b200_cg_set_xdepth", "b200_cg_update_px", "b200_cg_update_px_nx", "b200_cg_set_schedule", "b200_cg_set_pdl", "b200_cg_read_tail_times",
    "b200_csr_dot_partials_capacity",
    "b200_spmv_csr_dot", "b200_spmv_ellpack_dot", "b200_pcg_diag_inv", "b200_pcg_init", "b200_pcg_update_xr",
    "b200_pcg_update_p", "b200_bj_factor", "b200_bj_solve", "b200_pcg_update_xr_stored_z",
    "b200_dot_partials", "b200_residual_init_generic", "b200_checksum", "b200_halo_push",
    "b200_xchg_flag_prev_offset", "b200_xchg_flag_next_offset", "b200_xchg_halo_seq_offset", "b200_stencil5_nnz_before",
    "b200_gen_stencil5_csr", "b200_gen_stencil5_ellpack", "b200_gen_stencil5_entries", "b200_fill",
    "b200_coo_to_csr", "b200_patch_entry_values", "b200_parse_mtx_entries",
    # include/b200/api.h, extern "C" part
    "csr_mat", "ellpack_matrix", "build_ellpack_from_csr_local", "ensure_ellpack_structure_built", "get_operator",
    "calculate_spmv_metrics", "get_gpu_properties", "print_benchmark_metrics", "print_metrics_json",
    "print_metrics_csv", "read_matrix_type", "read_matrix_general", "read_matrix_symtogen", "load_matrix_market",
    "convert_csr_to_ellpack", "write_matrix_market_stencil5", "benchmark_with_stats",
    "cg_benchmark_with_stats_device", "cg_benchmark_with_stats_mgpu_partitioned", "export_cg_json",
    "export_cg_mgpu_json", "export_cg_csv",
    # extensions (host/host_common.h)
    "b200_operator_band", "b200_mgpu_init_single_process", "b200_mgpu_init_rank", "b200_mgpu_connect",
    "b200_mgpu_world", "b200_mgpu_rank", "b200_mgpu_finalize", "b200_synthetic_stencil", "b200_set_tuning",
    "b200_get_tuning", "b200_last_phase_times", "b200_last_h2d_bytes", "b200_cg_set_skip_zero_x0", "b200_host_all_zero", "b200_last_tail_times", "b200_last_gap_times", "b200_pcg_set_preconditioner", "b200_mgpu_halo_probe", "b200_load_matrix_market_device", "b200_operator_init_device_coo",
    "b200_operator_device_csr", "b200_free_device", "b200_copy_to_host", "b200_host_node_of_device",
    "b200_host_alloc_near", "b200_host_free",
]
# C++-linkage symbols of the reference API (Itanium mangling)
CXX_SYMBOLS = {
    "build_csr_struct": "_Z16build_csr_structP10MatrixData",
    "build_ellpack_from_csr_struct": "_Z29build_ellpack_from_csr_structPK9CSRMatrixP13ELLPACKMatrixPi",
    "cg_solve": "_Z8cg_solveP12SpmvOperatorP10MatrixDataPKdPd8CGConfigP7CGStats",
    "cg_solve_device": "_Z15cg_solve_deviceP12SpmvOperatorP10MatrixDataPKdPd8CGConfigP7CGStats",
    "pcg_solve_device": "_Z16pcg_solve_deviceP12SpmvOperatorP10MatrixDataPKdPd8CGConfigP7CGStats",
    "cg_solve_mgpu": "_Z13cg_solve_mgpuP12SpmvOperatorP10MatrixDataPKdPd16CGConfigMultiGPUP15CGStatsMultiGPU",
    "cg_solve_mgpu_partitioned":
        "_Z25cg_solve_mgpu_partitionedP12SpmvOperatorP10MatrixDataPKdPd16CGConfigMultiGPUP15CGStatsMultiGPU",
    "pcg_solve_mgpu_partitioned":
        "_Z26pcg_solve_mgpu_partitionedP12SpmvOperatorP10MatrixDataPKdPd16CGConfigMultiGPUP15CGStatsMultiGPU",
    "SPMV_CSR": "SPMV_CSR", "SPMV_STENCIL5_CSR": "SPMV_STENCIL5_CSR",
    "SPMV_STENCIL_HALO_MGPU": "SPMV_STENCIL_HALO_MGPU", "SPMV_ELLPACK": "SPMV_ELLPACK",
    "SPMV_STENCIL5_ELLPACK": "SPMV_STENCIL5_ELLPACK",
}

_lib = None


def build():
    subprocess.run(["make", "-C", PKG_DIR, "-s", "lib", "-j8"], check=True)


def load():
    """Load libspmv_b200.so and declare prototypes.  Raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(
            "%s not built (run `make -C %s lib`); there is no CPU fallback" % (LIB_PATH, PKG_DIR))
    L = C.CDLL(LIB_PATH)  # RTLD_LOCAL: must not interpose on other loaded libraries
    vp, ll, i32, dbl = C.c_void_p, C.c_longlong, C.c_int, C.c_double
    L.b200_version.restype = C.c_char_p
    L.b200_last_error.restype = C.c_char_p
    L.b200_launch_count.restype = C.c_ulonglong
    L.b200_stencil5_spmv.argtypes = [C.POINTER(Band), vp, vp, vp]
    L.b200_spmv_stencil5_csr.argtypes = [vp, vp, vp, vp, vp, i32, i32, vp]
    L.b200_spmv_stencil5_halo.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, ll, ll, i32, vp]
    L.b200_spmv_stencil5_ellpack.argtypes = [vp, vp, vp, vp, i32, i32, dbl, dbl, i32, vp]
    L.b200_stencil5_num_partials.argtypes = [C.POINTER(Band)]
    L.b200_stencil5_variant_info.restype = C.c_char_p
    L.b200_stencil5_variant_info.argtypes = [i32]
    L.b200_cg_set_kernel.restype = None
    L.b200_cg_set_kernel.argtypes = [i32]
    L.b200_stencil5_set_plain_variant.restype = None
    L.b200_stencil5_set_plain_variant.argtypes = [i32]
    L.b200_csr_variant_info.restype = C.c_char_p
    L.b200_csr_variant_info.argtypes = [i32]
    L.b200_csr_set_default_variant.restype = None
    L.b200_csr_set_default_variant.argtypes = [i32]
    L.b200_csr_plan_build.argtypes = [vp, ll, ll, C.POINTER(CsrPlan), vp]
    L.b200_spmv_csr.argtypes = [C.POINTER(CsrPlan), vp, vp, vp, vp, vp, ll, dbl, dbl, vp]
    L.b200_spmv_ellpack.argtypes = [vp, vp, vp, vp, ll, i32, dbl, dbl, vp]
    L.b200_csr_dot_partials_capacity.restype = ll
    L.b200_csr_dot_partials_capacity.argtypes = [ll]
    L.b200_spmv_csr_dot.argtypes = [C.POINTER(CsrPlan), vp, vp, vp, vp, vp, ll, vp, ll, C.POINTER(i32), vp, vp]
    L.b200_spmv_ellpack_dot.argtypes = [vp, vp, vp, vp, ll, i32, vp, ll, C.POINTER(i32), vp, vp]
    for f in ("b200_cg_scalars_bytes", "b200_cg_status_bytes", "b200_xchg_bytes", "b200_xchg_flag_prev_offset",
              "b200_xchg_flag_next_offset", "b200_xchg_halo_seq_offset"):
        getattr(L, f).restype = C.c_size_t
    ctx, push = C.POINTER(ReduceCtx), C.POINTER(HaloPushArgs)
    L.b200_cg_max_partials.argtypes = [C.POINTER(Band)]
    L.b200_cg_set_pdl.restype = None
    L.b200_cg_set_pdl.argtypes = [i32]
    L.b200_cg_residual_init.argtypes = [C.POINTER(Band), vp, vp, vp, vp, ctx, vp]
    L.b200_cg_spmv_dot.argtypes = [C.POINTER(Band), vp, vp, ctx, vp]
    L.b200_cg_update_xr.argtypes = [ll, vp, vp, vp, vp, ctx, vp]
    L.b200_cg_update_p.argtypes = [ll, vp, vp, vp, vp]
    L.b200_cg_reduce.argtypes = [ctx, i32, i32, i32, i32, vp]
    L.b200_cg_update_p_push.argtypes = [ll, vp, vp, vp, push, vp]
    L.b200_cg_spmv_fused.argtypes = [C.POINTER(Band), vp, vp, vp, vp, vp, ctx, vp]
    L.b200_cg_update_r.argtypes = [ll, vp, vp, push, ctx, vp]
    L.b200_cg_halo_dir.argtypes = [vp, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, i32, vp]
    L.b200_cg_finish_x.argtypes = [ll, vp, vp, vp, vp, i32, vp]
    L.b200_cg_spmv_fused_nx.argtypes = [C.POINTER(Band), vp, C.POINTER(vp), i32, vp, vp, vp, vp, C.POINTER(ReduceCtx), vp]
    L.b200_cg_finish_x_depth.argtypes = [ll, vp, C.POINTER(vp), i32, i32, i32, vp, vp]
    L.b200_cg_set_xdepth.argtypes = [i32]
    L.b200_cg_update_px_nx.argtypes = [ll, vp, vp, vp, C.POINTER(vp), i32, vp, vp, vp]
    L.b200_cg_update_px.argtypes = [ll, vp, vp, vp, vp, vp]
    L.b200_cg_set_schedule.restype = None
    L.b200_cg_set_schedule.argtypes = [i32]
    L.b200_cg_read_tail_times.argtypes = [vp, C.POINTER(TailTimes), vp]
    L.b200_dot_partials.argtypes = [ll, vp, vp, i32, ctx, vp]
    L.b200_residual_init_generic.argtypes = [ll, vp, vp, vp, vp, ctx, vp]
    L.b200_checksum.argtypes = [ll, vp, ctx, vp]
    L.b200_halo_push.argtypes = [vp, ll, push, vp, vp]
    L.b200_pcg_diag_inv.argtypes = [vp, vp, vp, ll, ll, i32, vp, vp, vp]
    L.b200_pcg_init.argtypes = [ll, vp, vp, vp, vp, ctx, vp]
    L.b200_bj_factor.argtypes = [vp, vp, vp, ll, ll, i32, i32, vp, vp, vp, vp, vp]
    L.b200_bj_solve.argtypes = [ll, ll, i32, vp, vp, vp, vp, vp, vp, vp]
    L.b200_pcg_update_xr_stored_z.argtypes = [ll, vp, vp, vp, vp, ctx, vp]
    L.b200_pcg_update_xr.argtypes = [ll, vp, vp, vp, vp, vp, ctx, vp]
    L.b200_pcg_update_p.argtypes = [ll, vp, vp, vp, vp, push, vp]
    L.b200_stencil5_nnz_before.restype = ll
    L.b200_stencil5_nnz_before.argtypes = [ll, ll]
    L.b200_gen_stencil5_csr.argtypes = [i32, ll, ll, dbl, dbl, vp, vp, vp, vp]
    L.b200_gen_stencil5_ellpack.argtypes = [i32, ll, ll, dbl, dbl, vp, vp, vp]
    L.b200_gen_stencil5_entries.argtypes = [i32, ll, ll, dbl, dbl, vp, vp]
    L.b200_fill.argtypes = [vp, ll, dbl, vp]
    L.b200_coo_to_csr.argtypes = [vp, ll, i32, i32, vp, vp, vp, vp]
    L.b200_patch_entry_values.argtypes = [vp, vp, i32, vp]
    L.b200_parse_mtx_entries.argtypes = [vp, ll, ll, vp, C.POINTER(ll), C.POINTER(i32), vp, i32, C.POINTER(i32), vp]
    # host API
    L.get_operator.restype = C.POINTER(SpmvOperator)
    L.get_operator.argtypes = [C.c_char_p]
    L.load_matrix_market.argtypes = [C.c_char_p, C.POINTER(MatrixData)]
    L.read_matrix_type.argtypes = [C.c_char_p]
    L.write_matrix_market_stencil5.argtypes = [i32, C.c_char_p]
    L.ensure_ellpack_structure_built.argtypes = [C.POINTER(MatrixData)]
    L.build_ellpack_from_csr_local.argtypes = [C.POINTER(CSRMatrix)]
    L.benchmark_with_stats.argtypes = [RUN_TIMED_FN, vp, vp, i32, C.POINTER(BenchmarkStats)]
    L.cg_benchmark_with_stats_device.argtypes = [C.POINTER(SpmvOperator), C.POINTER(MatrixData), vp, vp, CGConfig,
                                                 i32, C.POINTER(BenchmarkStats), C.POINTER(CGStats)]
    L.cg_benchmark_with_stats_mgpu_partitioned.argtypes = [C.POINTER(SpmvOperator), C.POINTER(MatrixData), vp, vp,
                                                           CGConfig, i32, C.POINTER(BenchmarkStats),
                                                           C.POINTER(CGStatsMultiGPU)]
    L.export_cg_json.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(MatrixData), C.POINTER(BenchmarkStats),
                                 C.POINTER(CGStats)]
    L.export_cg_mgpu_json.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(MatrixData), C.POINTER(BenchmarkStats),
                                      C.POINTER(CGStatsMultiGPU), i32]
    L.export_cg_csv.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(MatrixData), C.POINTER(BenchmarkStats),
                                C.POINTER(CGStats), C.c_bool]
    L.b200_operator_band.argtypes = [C.POINTER(SpmvOperator), C.POINTER(Band)]
    L.b200_mgpu_init_single_process.argtypes = [i32, C.POINTER(i32), i32]
    L.b200_mgpu_init_rank.argtypes = [i32, i32, i32, i32, vp]
    L.b200_mgpu_connect.argtypes = [vp]
    L.b200_mgpu_finalize.restype = None
    L.b200_synthetic_stencil.restype = MatrixData
    L.b200_synthetic_stencil.argtypes = [i32]
    L.b200_set_tuning.argtypes = [i32, i32]
    L.b200_last_phase_times.argtypes = [C.POINTER(dbl), C.POINTER(i32)]
    L.b200_last_h2d_bytes.restype = ll
    L.b200_last_h2d_bytes.argtypes = []
    L.b200_cg_set_skip_zero_x0.argtypes = [i32]
    L.b200_host_all_zero.argtypes = [vp, ll, i32]
    L.b200_last_tail_times.argtypes = [C.POINTER(dbl), C.POINTER(i32)]
    L.b200_last_gap_times.argtypes = [C.POINTER(dbl)]
    L.b200_pcg_set_preconditioner.argtypes = [i32]
    L.b200_mgpu_halo_probe.argtypes = [i32, C.POINTER(dbl), C.POINTER(ll)]
    L.b200_load_matrix_market_device.argtypes = [C.c_char_p, C.POINTER(MatrixData), C.POINTER(vp)]
    L.b200_operator_init_device_coo.argtypes = [C.POINTER(SpmvOperator), C.POINTER(MatrixData), vp]
    L.b200_operator_device_csr.argtypes = [C.POINTER(SpmvOperator), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(ll)]
    L.b200_free_device.argtypes = [vp]
    L.b200_free_device.restype = None
    L.b200_copy_to_host.argtypes = [vp, vp, C.c_size_t]
    L.b200_host_node_of_device.argtypes = [i32]
    L.b200_host_alloc_near.argtypes = [i32, C.c_size_t, C.POINTER(vp), C.POINTER(i32)]
    L.b200_host_free.argtypes = [vp]
    L.b200_host_free.restype = None
    # C++-linkage entry points
    L.build_csr_struct = getattr(L, CXX_SYMBOLS["build_csr_struct"])
    L.build_csr_struct.argtypes = [C.POINTER(MatrixData)]
    L.build_ellpack_from_csr_struct = getattr(L, CXX_SYMBOLS["build_ellpack_from_csr_struct"])
    L.build_ellpack_from_csr_struct.argtypes = [C.POINTER(CSRMatrix), C.POINTER(ELLPACKMatrix), C.POINTER(i32)]
    for nm in ("cg_solve", "cg_solve_device", "pcg_solve_device"):
        f = getattr(L, CXX_SYMBOLS[nm])
        f.argtypes = [C.POINTER(SpmvOperator), C.POINTER(MatrixData), vp, vp, CGConfig, C.POINTER(CGStats)]
        setattr(L, nm, f)
    for nm in ("cg_solve_mgpu", "cg_solve_mgpu_partitioned", "pcg_solve_mgpu_partitioned"):
        f = getattr(L, CXX_SYMBOLS[nm])
        f.argtypes = [C.POINTER(SpmvOperator), C.POINTER(MatrixData), vp, vp, CGConfig, C.POINTER(CGStatsMultiGPU)]
        setattr(L, nm, f)
    _lib = L
    return L


class B200Error(RuntimeError):
    pass


def check(rc, what=""):
    if rc != 0:
        raise B200Error("%s failed (rc=%d): %s" % (what, rc, load().b200_last_error().decode()))


def csr_mat():
    return CSRMatrix.in_dll(load(), "csr_mat")


def ellpack_matrix():
    return ELLPACKMatrix.in_dll(load(), "ellpack_matrix")


# ------------------------------------------------------------------ MatrixData helpers
class HostMatrix:
    """Owns a MatrixData for the lifetime of the Python object."""

    def __init__(self, md, keep=None, owned_by_c=False):
        self.md = md
        self._keep = keep
        self._owned_by_c = owned_by_c

    @classmethod
    def from_mtx(cls, path):
        md = MatrixData()
        rc = load().load_matrix_market(path.encode(), C.byref(md))
        if rc != 0:
            raise B200Error("load_matrix_market(%s) rc=%d" % (path, rc))
        return cls(md, owned_by_c=True)

    @classmethod
    def from_entries(cls, rows, cols, entries, grid_size=-1):
        ent = np.ascontiguousarray(entries, dtype=ENTRY_DTYPE)
        md = MatrixData(rows, cols, len(ent), grid_size, C.cast(ent.ctypes.data, C.POINTER(Entry)))
        return cls(md, keep=ent)

    @classmethod
    def synthetic_stencil(cls, n):
        return cls(load().b200_synthetic_stencil(n))

    def entries_array(self):
        n = self.md.nnz
        buf = (C.c_byte * (16 * n)).from_address(C.addressof(self.md.entries.contents))
        return np.frombuffer(buf, dtype=ENTRY_DTYPE).copy()

    def ptr(self):
        return C.byref(self.md)

    def __del__(self):
        try:
            if self._owned_by_c and self.md.entries:
                C.CDLL(None).free(self.md.entries)
                self.md.entries = None
        except Exception:
            pass


def host_csr_arrays():
    """Copy of the library's global host CSR (csr_mat) as numpy arrays."""
    c = csr_mat()
    rp = np.ctypeslib.as_array(c.row_ptr, shape=(c.nb_rows + 1,)).copy()
    ci = np.ctypeslib.as_array(c.col_indices, shape=(max(c.nb_nonzeros, 1),)).copy()[: c.nb_nonzeros]
    va = np.ctypeslib.as_array(c.values, shape=(max(c.nb_nonzeros, 1),)).copy()[: c.nb_nonzeros]
    return rp, ci, va


def cg_config(max_iters=1000, tol=1e-6, verbose=0, timers=0):
    return CGConfig(max_iters, tol, verbose, timers)


def stats_dict(s):
    return {f: getattr(s, f) for f, _ in s._fields_}
