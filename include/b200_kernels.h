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
 * flag >= epoch (or *d_epoch_ptr: this rank's own halo sequence counter, which equals the
 * neighbours' because every rank pushes in the same phases), all other rows run immediately
 * (transfer overlaps the interior).
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
    const uint32_t* d_epoch_ptr; /* if set, the number to wait for is read from here (device) instead of `epoch` */
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

/* kernel behind b200_stencil5_spmv when a band asks for the default: 0 = bulk-copy ring (csrc/stencil5.cuh),
 * 20 / 21 / 22 = sequential sweep with 1 / 2 / 4 rows per thread (csrc/stencil5_direct.cuh) */
void b200_stencil5_set_plain_variant(int v);
/* kernel family of the fused CG passes: 0 = bulk-copy ring, 1 = sequential sweep (default; env
 * B200_CG_KERNEL=ring|sweep).  The ring family retires x once per launch only (x depth 1). */
void b200_cg_set_kernel(int sweep);
int b200_cg_get_kernel(void);
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
 * Replaces the 11-launch iteration of cg_solve_device (src/solvers/cg_solver.cu:538-638).
 *
 * Reduction context (one per rank).  Every kernel below that produces a dot product writes one
 * partial sum per CTA.  The BLAS-1 kernels (a few hundred long-lived CTAs) end with a *tail*: the
 * last CTA to finish (ticket counters) adds the partials in a fixed two-level order, exchanges the
 * total with the other ranks over peer memory and applies the scalar recurrence.  The STENCIL5
 * kernels (O(1e5) short CTAs, where a ticket per CTA would cost 3 %) are followed by a small reduce
 * kernel that does the same in the same order, launched programmatically behind them (no launch gap).
 * Either way: no host round trip, identical bits.  Replaces
 * dot_kernel / final_sum_kernel / scalar_divide_kernel / check_convergence_kernel
 * (cg_solver.cu:110-132,384-431) and, multi-GPU, cublasDdot + MPI_Allreduce
 * (cg_solver_mgpu_partitioned.cu:145-154,531,583,645).
 *   d_scalars      b200_cg_scalars_bytes() bytes, zeroed before a solve
 *   h_status_mapped  optional pinned+mapped b200_cg_status_bytes() block the host may poll
 *   d_partials(_b) capacity doubles each; d_group_sums 2*ceil(capacity/256) doubles;
 *   d_tickets      1+ceil(capacity/256) 32-bit counters, zeroed ONCE (self-resetting)
 *   d_stash        2 doubles; d_out 2 doubles (which = 3)
 *   d_peer_xchg    world device pointers to every rank's exchange area (b200_xchg_bytes() bytes,
 *                  zeroed once), as mapped in this process; NULL when world == 1
 *   phases         3: whole reduction inside the producing kernel.  1: local total + stores to the
 *                  peers only -- b200_cg_reduce(phases = 2) must follow once every rank's producer
 *                  has been enqueued (several ranks on one stream).  0: partials only,
 *                  b200_cg_reduce(phases = 3) does the rest.
 * Exchange sequence numbers (scalar exchanges, halo pushes) are counted ON THE DEVICE inside the
 * exchange areas: ranks stay in lockstep however many no-op iterations each host enqueues. */
typedef struct {
    void* d_scalars;
    void* h_status_mapped;
    double* d_partials;
    double* d_partials_b;
    double* d_group_sums;
    uint32_t* d_tickets;
    long long capacity;
    double* d_stash;
    double* d_out;
    int rank, world;
    void* const* d_peer_xchg;
    double tol;
    int phases;
} b200_reduce_ctx;

/* which: 0 = r0.r0 (sets rr_old, b_norm), 1 = p.Ap (alpha), 2 = r.r (convergence test, beta,
 * iteration count), 3 = plain sum(s) into d_out, 4 = PCG rho_0, 5 = PCG r.r + r.z */
enum { B200_RED_RR0 = 0, B200_RED_PAP = 1, B200_RED_RR = 2, B200_RED_SUM = 3, B200_RED_RZ0 = 4, B200_RED_PCG = 5,
       B200_RED_RRC = 6 /* r.r: convergence test only */, B200_RED_RZ = 7 /* r.z: beta = rho_new / rho */ };

size_t b200_cg_scalars_bytes(void);
size_t b200_cg_status_bytes(void);
size_t b200_xchg_bytes(void);
/* partial sums the STENCIL5 kernels write for this band (= their grid size, halo CTAs included) */
int b200_cg_max_partials(const b200_band* band);
/* programmatic dependent launch in the iteration loop (the next kernel is scheduled while the tail of
 * the previous one runs; every such kernel starts with griddepcontrol.wait).  mode bit 0: the small
 * kernels (reduce, halo direction, finish_x), bit 1: the STENCIL5 kernels; default 3.  The persistent
 * BLAS-1 kernels launch the normal way (a dependent launch stacks their CTAs unevenly over the SMs) unless
 * bit 2 is set (A/B switch: gains 2 us per iteration at 12.5 M rows, loses 1.3 % at 400 M).
 * env B200_PDL=<0..7> */
void b200_cg_set_pdl(int mode);

/* Halo addressing of a band inside a multi-GPU solve: where neighbours write, what to wait for. */
typedef struct {
    double* d_dst_prev;      /* rank-1's landing buffer for MY first `halo` elements (peer memory) */
    double* d_dst_next;      /* rank+1's landing buffer for my last `halo` elements */
    uint32_t* d_flag_prev;   /* rank-1's arrival word for them */
    uint32_t* d_flag_next;
    void* d_my_xchg;         /* this rank's exchange area */
    int halo;                /* elements per edge = grid_size */
} b200_halo_push_args;

/* setup: r = b - A x ; p = r ; r.r -> rr_old, b_norm           (cg_solver.cu:498-528) */
int b200_cg_residual_init(const b200_band* band, const double* d_x, const double* d_b, double* d_r,
                          double* d_p, const b200_reduce_ctx* ctx, b200_stream stream);
/* K1: Ap = A p ; p.Ap -> alpha                                  (cg_solver.cu:541-560) */
int b200_cg_spmv_dot(const b200_band* band, const double* d_p, double* d_Ap, const b200_reduce_ctx* ctx,
                     b200_stream stream);
/* K2: x += alpha p ; r -= alpha Ap ; r.r -> convergence, beta   (cg_solver.cu:564-637) */
int b200_cg_update_xr(long long n, const double* d_p, const double* d_Ap, double* d_x, double* d_r,
                      const b200_reduce_ctx* ctx, b200_stream stream);
/* K3: p = r + beta p                                            (cg_solver.cu:628) */
int b200_cg_update_p(long long n, const void* d_scalars, const double* d_r, double* d_p,
                     b200_stream stream);

/* R: the same fixed-order sum + exchange + recurrence as a launch of its own (operators behind
 * run_device; the wait half when several ranks share one stream; exchanges without data). */
int b200_cg_reduce(const b200_reduce_ctx* ctx, int which, int n_partials, int two_sums, int phases,
                   b200_stream stream);

/* generic helpers for operators without a fused entry point */
int b200_dot_partials(long long n, const double* d_x, const double* d_y, int which, const b200_reduce_ctx* ctx,
                      b200_stream stream);
int b200_residual_init_generic(long long n, const double* d_b, const double* d_Ap, double* d_r,
                               double* d_p, const b200_reduce_ctx* ctx, b200_stream stream);
/* sum(x) -> d_out[0], sum(x^2) -> d_out[1], over all ranks        (cg_solver.cu:658-665) */
int b200_checksum(long long n, const double* d_x, const b200_reduce_ctx* ctx, b200_stream stream);

/* Halo push over NVLink peer memory.  Replaces exchange_halo_mpi
 * (cg_solver_mgpu_partitioned.cu:173-231): the first / last `halo` elements of d_v_local are
 * stored into the neighbours' landing buffers, the last CTA bumps this rank's halo sequence number
 * and release-stores it into the neighbours' arrival words. */
int b200_halo_push(const double* d_v_local, long long n_local, const b200_halo_push_args* h,
                   const void* d_scalars, b200_stream stream);
/* K3 with the halo push fused in */
int b200_cg_update_p_push(long long n, const void* d_scalars, const double* d_r, double* d_p,
                          const b200_halo_push_args* h, b200_stream stream);

/* ---- "deferred x" CG schedule: 112 B/row per iteration -----------------------------------
 * Same recurrences as cg_solve_device (src/solvers/cg_solver.cu:538-638), regrouped:
 *   b200_cg_spmv_fused : p_new = r + beta p_old (update_p_kernel :91-96), x += alpha_prev p_old
 *                        (axpy_kernel_device :59-66, one iteration late), Ap = A p_new, p.Ap -> alpha,
 *                        in ONE pass of the STENCIL5 kernel.  band halos hold the NEW p.
 *   b200_cg_update_r   : r -= alpha Ap (axpy_sub_kernel_device :70-78), r.r -> convergence, beta,
 *                        summed in the order of b200_cg_update_xr.  push != NULL: the first / last
 *                        `halo` elements of r also go into the neighbours' landing buffers.
 *   b200_cg_halo_dir   : halo copies of p follow the same recurrence from the pushed r edges,
 *                        p_halo = fma(beta, p_halo_old, r_halo) (beta_zero: first direction, p0 = r0).
 *   b200_cg_finish_x   : the x update still pending after the last iteration.
 * Every iterate is bit-identical to the classic K1 / K2 / K3 schedule.                          */
int b200_cg_spmv_fused(const b200_band* band, const double* d_p_old, const double* d_r, double* d_p_new,
                       double* d_x, double* d_Ap, const b200_reduce_ctx* ctx, b200_stream stream);
/* The same pass with the x stream amortised over `depth` iterations (depth <= 4, depth + 1 direction buffers,
 * direction j in buffer j % (depth + 1)): nx = 0 leaves x alone, nx = m retires the m pending updates
 *   x = fma(alpha_{it-1}, p_{it-1}, ... fma(alpha_{it-m}, p_{it-m}, x))       (oldest first: bit-identical to
 * m separate axpy_kernel_device launches, cg_solver.cu:59-66) in ONE read-modify-write of x;
 * d_p_older[k] = p_{it-2-k}, k < nx - 1.  alpha history: kept by the p.Ap tail in the scalar block.
 * b200_cg_finish_x_depth retires what is still pending after the last iteration. */
int b200_cg_spmv_fused_nx(const b200_band* band, const double* d_p_old, const double* const* d_p_older, int nx,
                          const double* d_r, double* d_p_new, double* d_x, double* d_Ap,
                          const b200_reduce_ctx* ctx, b200_stream stream);
int b200_cg_finish_x_depth(long long n, const void* d_scalars, const double* const* d_pbuf, int nbuf, int depth,
                           int only_if_converged, double* d_x, b200_stream stream);
/* x retirement depth of the deferred-x schedule (1..4, env B200_CG_XDEPTH); returns the previous value */
int b200_cg_set_xdepth(int depth);
int b200_cg_update_r(long long n, const double* d_Ap, double* d_r, const b200_halo_push_args* push,
                     const b200_reduce_ctx* ctx, b200_stream stream);
int b200_cg_halo_dir(const double* d_r_prev, const double* d_r_next, const double* d_pold_prev,
                     const double* d_pold_next, double* d_pnew_prev, double* d_pnew_next, int halo,
                     const uint32_t* d_flag_prev, const uint32_t* d_flag_next, const void* d_my_xchg,
                     void* d_scalars, int beta_zero, b200_stream stream);
int b200_cg_finish_x(long long n, const void* d_scalars, const double* d_p0, const double* d_p1,
                     double* d_x, int only_if_converged, b200_stream stream);
/* K3x, for operators without a fused SpMV: p = r + beta p and x += alpha p_old in one pass, so that K2
 * shrinks to b200_cg_update_r (64 instead of 72 B/row of BLAS-1 traffic per iteration; update_p_kernel
 * :91-96 + axpy_kernel_device :59-66).  b200_cg_finish_x(only_if_converged = 1) closes the solve. */
int b200_cg_update_px(long long n, const void* d_scalars, const double* d_r, double* d_p, double* d_x,
                      b200_stream stream);
/* K3x with the x stream amortised over `depth` iterations, like b200_cg_spmv_fused_nx: p_new = r + beta p_old into
 * a fresh direction buffer; nx = m > 0 also retires the m pending x updates (p_old and d_p_older[k] = the
 * directions before it), oldest first.  b200_cg_finish_x_depth(only_if_converged = 1) closes the solve. */
int b200_cg_update_px_nx(long long n, const void* d_scalars, const double* d_r, const double* d_p_old,
                         const double* const* d_p_older, int nx, double* d_p_new, double* d_x, b200_stream stream);
/* host engine: 1 = deferred-x schedule (default for the STENCIL5 path), 0 = classic 3-launch schedule;
 * the environment variable B200_CG_SCHEDULE=classic selects 0 at start-up */
void b200_cg_set_schedule(int deferred_x);
/* device-measured time spent in the reduction tails of a solve, per `which` */
typedef struct {
    unsigned long long ns[8];      /* time inside the tails */
    unsigned int count[8];
    unsigned long long gap_ns[8];  /* time between the end of the previous tail and the begin of this one: the
                                      kernel(s) that produced the partial sums (no events needed) */
    int error;
} b200_cg_tail_times;
int b200_cg_read_tail_times(const void* d_scalars, b200_cg_tail_times* h_out, b200_stream stream);

/* ---- Jacobi-preconditioned CG ----------------------------------------------------------------
 * Not in the reference (its cg_solver.h:6-7 / README name preconditioning as the next step): the
 * textbook recurrence over the conventions of cg_solve_device (cg_solver.cu:498-638).  z = D^-1 r is
 * never stored.  rho = r.z lives in the scalar block where plain CG keeps r.r, so b200_cg_spmv_dot
 * serves unchanged.  Multi-GPU: K3p pushes the edges of the new p like b200_cg_update_p_push. */
int b200_pcg_diag_inv(const int* d_row_ptr, const int* d_col_idx, const double* d_values, long long n_local,
                      long long row_offset, int ell_width, double* d_dinv, int* d_err, b200_stream stream);
/* p0 = z0 ; r0.z0 -> rho_0, with z0 = D^-1 r0 (d_dinv) or an already solved z0 (d_dinv NULL, d_z: block-Jacobi) */
int b200_pcg_init(long long n, const double* d_r, const double* d_dinv, const double* d_z, double* d_p,
                  const b200_reduce_ctx* ctx, b200_stream stream);
/* K2p: x += alpha p ; r -= alpha Ap ; r.r -> convergence ; r.z -> beta, rho */
int b200_pcg_update_xr(long long n, const double* d_p, const double* d_Ap, const double* d_dinv, double* d_x,
                       double* d_r, const b200_reduce_ctx* ctx, b200_stream stream);
/* Block-Jacobi with line blocks: M = the W / C / E part of A inside every grid row (one tridiagonal block per
 * grid row, clipped to the band); z = M^-1 r by the Thomas algorithm, one thread per block.  b200_bj_factor
 * (once per solve) stores the forward elimination of the matrix, b200_bj_solve applies it.  The iteration is
 * K1 -> b200_pcg_update_xr_stored_z (r.r: convergence only) -> b200_bj_solve -> b200_dot_partials(r, z,
 * B200_RED_RZ) -> b200_pcg_update_p(d_dinv = NULL, d_r = z). */
int b200_bj_factor(const int* d_row_ptr, const int* d_col_idx, const double* d_values, long long n_local,
                   long long row_offset, int grid_size, int ell_width, double* d_m, double* d_invd, double* d_c,
                   int* d_err, b200_stream stream);
int b200_bj_solve(long long n_local, long long row_offset, int grid_size, const void* d_scalars, const double* d_m,
                  const double* d_invd, const double* d_c, const double* d_r, double* d_z, b200_stream stream);
int b200_pcg_update_xr_stored_z(long long n, const double* d_p, const double* d_Ap, double* d_x, double* d_r,
                                const b200_reduce_ctx* ctx, b200_stream stream);
/* K3p: p = D^-1 r + beta p (push != NULL: + halo push); d_dinv NULL: d_r already holds z */
int b200_pcg_update_p(long long n, const void* d_scalars, const double* d_r, const double* d_dinv, double* d_p,
                      const b200_halo_push_args* push, b200_stream stream);
/* offsets of the halo arrival words / the halo sequence counter inside an exchange area */
size_t b200_xchg_flag_prev_offset(void);
size_t b200_xchg_flag_next_offset(void);
size_t b200_xchg_halo_seq_offset(void);

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
 * array of {int row; int col; double value;} (0-based).  d_row_ptr: rows+1 ints.  Entries with a row
 * outside [0, rows) or a column outside [0, cols) (cols <= 0: only negative columns) are rejected. */
int b200_coo_to_csr(const void* d_entries, long long nnz, int rows, int cols, int* d_row_ptr, int* d_col_idx,
                    double* d_values, b200_stream stream);
/* entries[pair[2k]].value = bits pair[2k+1], k < n: one staged upload + one launch (values the host
 * re-read with strtod after b200_parse_mtx_entries) */
int b200_patch_entry_values(void* d_entries, const long long* h_pairs, int n, b200_stream stream);
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
