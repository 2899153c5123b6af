/*
 * b200_kernels.h -- the thin C ABI in front of the hand-written sm_100a kernels.
 *
 * Plain pointers and sizes only (no CUDA, torch or C++ types): this is what a cgo / JNI /
 * ctypes / plain-C binding of the reference would bind.  Every pointer prefixed d_ is DEVICE
 * memory on the current CUDA device; launches are asynchronous on `stream` (NULL = default
 * stream) and return 0 on success or a B200_E* code (b200_last_error() has the text).
 * Host code (cuda-spmv-benchmark_b200/host, the CLIs, tests) includes only this header, never a
 * kernel header.  Each entry point names the reference kernel / call it stands in for.
 */
#ifndef B200_KERNELS_H
#define B200_KERNELS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* b200_stream; /* a cudaStream_t */

enum {
    B200_OK = 0,
    B200_EINVAL = 1,   /* bad argument (NULL pointer, misaligned values, grid mismatch) */
    B200_ECUDA = 2,    /* CUDA runtime error, see b200_last_error() */
    B200_ENOMEM = 3,
    B200_ETIMEOUT = 4, /* peer flag wait timed out */
    B200_ENODEV = 5    /* no sm_100 device / kernels unavailable: the library never falls back to the CPU */
};

const char* b200_version(void);
const char* b200_last_error(void);
/* number of kernels launched by this library in this process (for bench.py's gpu_launches) */
unsigned long long b200_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Band descriptor: a contiguous block of matrix rows [row_offset, row_offset + n_local) of the
 * grid_size x grid_size 5-point stencil, with its slice of the matrix arrays.
 *   layout 0 (CSR-direct): d_values / d_col_idx are the slice [row_ptr[row_offset],
 *     row_ptr[row_offset+n_local]) of the full CSR arrays, d_row_ptr is rebased to 0 and has
 *     n_local+1 entries, col ids stay GLOBAL -- exactly the reference's local partition
 *     (src/solvers/cg_solver_mgpu_partitioned.cu:306-329).  Single GPU: row_offset = 0,
 *     n_local = N, i.e. the plain arrays of src/spmv/spmv_stencil_csr_direct.cu:201-213.
 *   layout 1 (ELLPACK width 5, row-major, padding index -1): d_row_ptr = NULL.
 * d_values must be 16-byte aligned; values_len = number of readable doubles behind it.
 * Halo pointers follow stencil5_csr_partitioned_halo_kernel
 * (src/spmv/spmv_stencil_partitioned_halo_kernel.cu:17-21): grid_size doubles each, NULL when
 * there is no neighbour.  d_flag_prev/next (optional) are 32-bit arrival epochs in THIS GPU's
 * memory, release-stored by the neighbours' b200_halo_push; rows that read a halo wait until
 * flag >= epoch, all other rows run immediately (transfer overlaps the interior).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    const int* d_row_ptr;
    const int* d_col_idx;
    const double* d_values;
    long long values_len;
    long long row_offset;
    long long n_local;
    int grid_size;
    int layout;
    const double* d_halo_prev;
    const double* d_halo_next;
    const uint32_t* d_flag_prev;
    const uint32_t* d_flag_next;
    uint32_t epoch;
    int rows_per_item; /* tuning: grid rows marched per warp item, 0 = default */
    int variant;       /* tuning: kernel instantiation, 0 = default (see b200_stencil5_variant_info) */
} b200_band;

/* y = A x on the band.  Stands in for the launches at spmv_stencil_csr_direct.cu:230-240,267-271
 * (single GPU) and cg_solver_mgpu_partitioned.cu:467-469,550-552 (band + halos). */
int b200_stencil5_spmv(const b200_band* band, const double* d_x, double* d_y, b200_stream stream);

/* Convenience wrappers with the reference kernels' own argument lists. */
int b200_spmv_stencil5_csr(const int* d_row_ptr, const int* d_col_idx, const double* d_values,
                           const double* d_x, double* d_y, int N, int grid_size, b200_stream stream);
int b200_spmv_stencil5_halo(const int* d_row_ptr, const int* d_col_idx, const double* d_values,
                            const double* d_x_local, const double* d_x_halo_prev,
                            const double* d_x_halo_next, double* d_y, int n_local, long long row_offset,
                            long long N, int grid_size, b200_stream stream);
/* The ELLPACK stencil kernel the reference declares but never defines
 * (include/spmv_stencil.h:40-42): y = alpha * A x + beta * y. */
int b200_spmv_stencil5_ellpack(const double* d_values, const int* d_col_indices, const double* d_x,
                               double* d_y, int num_rows, int width, double alpha, double beta,
                               int grid_size, b200_stream stream);

/* number of per-CTA partial sums the fused kernels below write for this band */
int b200_stencil5_num_partials(const b200_band* band);
/* human-readable description of tuning variant v (NULL past the last one) */
const char* b200_stencil5_variant_info(int v);

/* ---- generic CSR / ELLPACK (no cuSPARSE) ------------------------------------------------ */
typedef struct {
    int rows_per_block;       /* rows per warp item */
    int window;               /* bulk-copy window, in non-zeros */
    int vector_threshold;     /* longest row of a 32-row group above which it runs warp-per-row */
    int variant;              /* tuning variant (0 = default), see b200_csr_variant_info */
    unsigned long long hist[33]; /* rows with length in (2^(b-1), 2^b] */
    unsigned long long max_row_len;
    double mean_row_len;
} b200_csr_plan;

/* human-readable description of CSR tuning variant v (NULL past the last one) */
const char* b200_csr_variant_info(int v);
/* variant used when a plan says 0, and by b200_spmv_ellpack (tuning knob for tools/sweep.py) */
void b200_csr_set_default_variant(int v);
/* takes the row-length histogram on the device and picks the block shape */
int b200_csr_plan_build(const int* d_row_ptr, long long n_rows, long long nnz, b200_csr_plan* plan,
                        b200_stream stream);
/* y = alpha*A x + beta*y.  Stands in for cusparseSpMV (spmv_cusparse_csr.cu:246,281) and the
 * scalar csr_spmv_kernel (cg_solver_mgpu_partitioned.cu:40-56). */
int b200_spmv_csr(const b200_csr_plan* plan, const int* d_row_ptr, const int* d_col_idx,
                  const double* d_values, const double* d_x, double* d_y, long long n_rows,
                  double alpha, double beta, b200_stream stream);
int b200_spmv_ellpack(const int* d_indices, const double* d_values, const double* d_x, double* d_y,
                      long long n_rows, int width, double alpha, double beta, b200_stream stream);
/* y = A x fused with the partial sums of x.y (CG: p.Ap of a square operator; replaces the separate
 * dot_kernel pass, cg_solver.cu:550-556).  One partial per warp item, summed in item order by
 * b200_cg_reduce; d_partials needs b200_csr_dot_partials_capacity(n_rows) slots.  d_scalars (optional):
 * the launch is a no-op once the solve has converged.  Needs 16-byte aligned col_idx / values. */
long long b200_csr_dot_partials_capacity(long long n_rows);
int b200_spmv_csr_dot(const b200_csr_plan* plan, const int* d_row_ptr, const int* d_col_idx,
                      const double* d_values, const double* d_x, double* d_y, long long n_rows,
                      double* d_partials, long long partials_capacity, int* n_partials_out,
                      const void* d_scalars, b200_stream stream);
int b200_spmv_ellpack_dot(const int* d_indices, const double* d_values, const double* d_x, double* d_y,
                          long long n_rows, int width, double* d_partials, long long partials_capacity,
                          int* n_partials_out, const void* d_scalars, b200_stream stream);

/* ---- fused CG steps ---------------------------------------------------------------------
 * d_scalars: b200_cg_scalars_bytes() bytes of device memory (zero it before a solve).
 * d_partials: scratch for per-CTA partial sums, at least b200_cg_max_partials(band) doubles.
 * Replaces the 11-launch iteration of cg_solve_device (src/solvers/cg_solver.cu:538-638).      */
size_t b200_cg_scalars_bytes(void);
size_t b200_cg_status_bytes(void);
size_t b200_xchg_bytes(void);
int b200_cg_max_partials(const b200_band* band);

/* setup: r = b - A x ; p = r ; partials <- r.r          (cg_solver.cu:498-517) */
int b200_cg_residual_init(const b200_band* band, const double* d_x, const double* d_b, double* d_r,
                          double* d_p, double* d_partials, void* d_scalars, b200_stream stream);
/* K1: Ap = A p ; partials <- p.Ap                         (cg_solver.cu:541-551) */
int b200_cg_spmv_dot(const b200_band* band, const double* d_p, double* d_Ap, double* d_partials,
                     const void* d_scalars, b200_stream stream);
/* K2: x += alpha p ; r -= alpha Ap ; partials <- r.r      (cg_solver.cu:564-585) */
int b200_cg_update_xr(long long n, const void* d_scalars, const double* d_p, const double* d_Ap,
                      double* d_x, double* d_r, double* d_partials, int* n_partials_out,
                      b200_stream stream);
/* K3: p = r + beta p                                      (cg_solver.cu:628) */
int b200_cg_update_p(long long n, const void* d_scalars, const double* d_r, double* d_p,
                     b200_stream stream);

/* R: fixed-order sum of partials (+ rank exchange) and the scalar recurrences.
 * which: 0 = r0.r0 (sets rr_old, b_norm), 1 = p.Ap (alpha), 2 = r.r (convergence test, beta,
 * iteration count), 3 = plain sum into d_out.
 * phases: 1 = local sum + push to peers, 2 = wait for peers + recurrences, 3 = both in one launch.
 * Multi-rank: d_peer_xchg[world] are every rank's exchange areas as mapped in this process,
 * epoch must be the same on all ranks for the same reduction and differ between reductions.
 * Replaces dot_kernel/final_sum_kernel/scalar_divide_kernel/check_convergence_kernel and, on
 * the multi-GPU path, cublasDdot + MPI_Allreduce (cg_solver_mgpu_partitioned.cu:145-154,531). */
int b200_cg_reduce(const double* d_partials, int n_partials, int which, int phases, double tol,
                   void* d_scalars, void* h_status_mapped, double* d_out, int rank, int world,
                   uint32_t epoch, void* const* d_peer_xchg, double* d_stash, b200_stream stream);

/* generic helpers for operators without a fused entry point */
int b200_dot_partials(long long n, const void* d_scalars, const double* d_x, const double* d_y,
                      double* d_partials, int* n_partials_out, b200_stream stream);
int b200_residual_init_generic(long long n, const double* d_b, const double* d_Ap, double* d_r,
                               double* d_p, double* d_partials, int* n_partials_out, b200_stream stream);
/* sum(x) and sum(x^2) partials, n_partials_out each            (cg_solver.cu:658-665) */
int b200_checksum_partials(long long n, const double* d_x, double* d_psum, double* d_psq,
                           int* n_partials_out, b200_stream stream);

/* Halo push over NVLink peer memory.  Replaces exchange_halo_mpi
 * (cg_solver_mgpu_partitioned.cu:173-231).  dst pointers / flags live in the NEIGHBOURS' memory
 * (peer-mapped); d_my_xchg is this rank's exchange area. */
int b200_halo_push(const double* d_v_local, long long n_local, int halo, double* d_dst_prev,
                   double* d_dst_next, uint32_t* d_flag_prev, uint32_t* d_flag_next, uint32_t epoch,
                   void* d_my_xchg, const void* d_scalars, b200_stream stream);
/* K3 with the halo push fused in: p = r + beta p, the first / last `halo` elements are also stored
 * into the neighbours' landing buffers and the arrival epoch is published by the last CTA. */
int b200_cg_update_p_push(long long n, const void* d_scalars, const double* d_r, double* d_p, int halo,
                          double* d_dst_prev, double* d_dst_next, uint32_t* d_flag_prev,
                          uint32_t* d_flag_next, uint32_t epoch, void* d_my_xchg, b200_stream stream);

/* ---- "deferred x" CG schedule: 4 launches and 112 B/row per iteration ----------------------
 * Same recurrences as cg_solve_device (src/solvers/cg_solver.cu:538-638), regrouped:
 *   b200_cg_spmv_fused : p_new = r + beta p_old (update_p_kernel :91-96), x += alpha_prev p_old
 *                        (axpy_kernel_device :59-66, one iteration late), Ap = A p_new and the p.Ap
 *                        partials, in ONE pass of the STENCIL5 kernel.  band halos hold the NEW p.
 *   b200_cg_update_r   : r -= alpha Ap (axpy_sub_kernel_device :70-78) and the r.r partials, summed
 *                        in the order of b200_cg_update_xr.  The _push form also stores the first /
 *                        last `halo` elements of r into the neighbours' landing buffers.
 *   b200_cg_halo_dir   : halo copies of p follow the same recurrence from the pushed r edges.
 *   b200_cg_finish_x   : the x update still pending after the last iteration.
 * Every iterate is bit-identical to the 5-launch schedule.                                      */
int b200_cg_spmv_fused(const b200_band* band, const double* d_p_old, const double* d_r, double* d_p_new,
                       double* d_x, double* d_Ap, double* d_partials, const void* d_scalars,
                       b200_stream stream);
int b200_cg_update_r(long long n, const void* d_scalars, const double* d_Ap, double* d_r,
                     double* d_partials, int* n_partials_out, b200_stream stream);
int b200_cg_update_r_push(long long n, const void* d_scalars, const double* d_Ap, double* d_r,
                          double* d_partials, int* n_partials_out, int halo, double* d_dst_prev,
                          double* d_dst_next, uint32_t* d_flag_prev, uint32_t* d_flag_next,
                          uint32_t epoch, void* d_my_xchg, b200_stream stream);
int b200_cg_halo_dir(const double* d_r_prev, const double* d_r_next, const double* d_pold_prev,
                     const double* d_pold_next, double* d_pnew_prev, double* d_pnew_next, int halo,
                     const uint32_t* d_flag_prev, const uint32_t* d_flag_next, uint32_t epoch,
                     void* d_scalars, int beta_zero, b200_stream stream);
/* b200_cg_reduce(which = 2) and b200_cg_halo_dir in one launch: once beta is known the reduce CTA
 * advances the halo copies of the direction (multi-GPU, deferred-x schedule) */
int b200_cg_reduce_rr_dir(const double* d_partials, int n_partials, int phases, double tol, void* d_scalars,
                          void* h_status_mapped, int rank, int world, uint32_t epoch,
                          void* const* d_peer_xchg, double* d_stash, const double* d_r_prev,
                          const double* d_r_next, const double* d_pold_prev, const double* d_pold_next,
                          double* d_pnew_prev, double* d_pnew_next, int halo, const uint32_t* d_flag_prev,
                          const uint32_t* d_flag_next, uint32_t halo_epoch, b200_stream stream);
int b200_cg_finish_x(long long n, const void* d_scalars, const double* d_p0, const double* d_p1,
                     double* d_x, b200_stream stream);
/* host engine: 1 = deferred-x schedule (default for the STENCIL5 path), 0 = classic 5-launch schedule;
 * the environment variable B200_CG_SCHEDULE=classic selects 0 at start-up */
void b200_cg_set_schedule(int deferred_x);
/* ---- Jacobi-preconditioned CG (single GPU) --------------------------------------------------
 * Not in the reference (its cg_solver.h:6-7 / README name preconditioning as the next step): the
 * textbook recurrence over the conventions of cg_solve_device (cg_solver.cu:498-638).  z = D^-1 r is
 * never stored.  rho = r.z lives in the scalar block where plain CG keeps r.r, so b200_cg_spmv_dot /
 * b200_cg_reduce(which = 1) serve unchanged; b200_cg_reduce(which = 4) stores rho_0.               */
int b200_pcg_diag_inv(const int* d_row_ptr, const int* d_col_idx, const double* d_values, long long n_local,
                      long long row_offset, int ell_width, double* d_dinv, int* d_err, b200_stream stream);
int b200_pcg_init(long long n, const double* d_r, const double* d_dinv, double* d_p, double* d_partials,
                  int* n_partials_out, b200_stream stream);
int b200_pcg_update_xr(long long n, const void* d_scalars, const double* d_p, const double* d_Ap,
                       const double* d_dinv, double* d_x, double* d_r, double* d_partials_rr,
                       double* d_partials_rz, int* n_partials_out, b200_stream stream);
int b200_pcg_update_p(long long n, const void* d_scalars, const double* d_r, const double* d_dinv, double* d_p,
                      b200_stream stream);
int b200_pcg_reduce(const double* d_partials_rr, const double* d_partials_rz, int n_partials, double tol,
                    void* d_scalars, void* h_status_mapped, b200_stream stream);
/* offsets of the halo flags inside an exchange area */
size_t b200_xchg_flag_prev_offset(void);
size_t b200_xchg_flag_next_offset(void);

/* ---- device-side matrix construction (bit-identical to generator -> reader -> CSR build) ---- */
long long b200_stencil5_nnz_before(long long row, long long grid_size);
int b200_gen_stencil5_csr(int grid_size, long long row_offset, long long n_local, double center,
                          double neighbour, int* d_row_ptr, int* d_col_idx, double* d_values,
                          b200_stream stream);
int b200_gen_stencil5_ellpack(int grid_size, long long row_offset, long long n_local, double center,
                              double neighbour, int* d_indices, double* d_values, b200_stream stream);
/* d_entries: array of {int row; int col; double value;} in the generator's emission order */
int b200_gen_stencil5_entries(int grid_size, long long row_offset, long long n_local, double center,
                              double neighbour, void* d_entries, b200_stream stream);
int b200_fill(double* d_p, long long n, double value, b200_stream stream);

/* ---- device-side ingest (arbitrary matrices) -----------------------------------------------
 * COO -> CSR on the device, bit-identical to build_csr_struct (src/spmv/spmv_cusparse_csr.cu:85-157):
 * rows ascending, columns ascending inside a row, equal columns in input order.  d_entries is an
 * array of {int row; int col; double value;} (0-based).  d_row_ptr: rows+1 ints. */
int b200_coo_to_csr(const void* d_entries, long long nnz, int rows, int* d_row_ptr, int* d_col_idx,
                    double* d_values, b200_stream stream);
/* Parses the ENTRY lines of a Matrix Market file ("i j value", 1-based) held in device memory into
 * d_entries (0-based), like the fscanf loop of src/io/io.cu:153-166.  n_lines_out = non-blank lines
 * found; literals outside the exact fast path are listed as (entry index, byte offset of the value
 * token) pairs in h_inexact_pairs for the caller to redo with strtod; malformed_out != 0 when a
 * line does not hold exactly three tokens (caller should use the host reader). */
int b200_parse_mtx_entries(const void* d_text, long long n_bytes, long long max_entries, void* d_entries,
                           long long* n_lines_out, int* n_inexact_out, long long* h_inexact_pairs,
                           int inexact_cap, int* malformed_out, b200_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* B200_KERNELS_H */
