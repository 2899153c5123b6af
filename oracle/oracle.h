/*
 * oracle.h -- CPU restatement of the reference's SpMV + CG hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this library, and there only as the
 * checker or as the timed CPU baseline.  The product path (the CUDA library
 * under cuda-spmv-benchmark_b200/) never links or calls it.
 *
 * Every function cites the reference file:line (relative to the upstream
 * repo root) whose behaviour it restates.  Parity pinning:
 *   - structure functions (generator, reader, COO->CSR) are pinned against the
 *     reference's own host code compiled into oracle/_ref/libref_host.so
 *     (tests/test_oracle_vs_ref.py) and against tests/golden/ fixtures made
 *     from it (tests/golden/make_golden.py);
 *   - numeric functions (SpMV, CG) are pinned against the reference's CLI
 *     binaries run on a B200 (tests/golden/ref_gpu_*.json, made by
 *     oracle/run_ref_gpu.sh) and the KATs the reference documents.
 *   - ELLPACK and the cuSPARSE / cuBLAS boundaries have no reference
 *     implementation or golden vector: "parity unpinned" (defined here as
 *     y_ELL == y_CSR with sequential-k accumulation).
 */
#ifndef ORACLE_H
#define ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* same layout as the reference's Entry (include/io.h:43-47) */
typedef struct {
    int row;
    int col;
    double value;
} orc_entry;

/* same field order as the reference's MatrixData (include/io.h:53-59) */
typedef struct {
    int rows;
    int cols;
    int nnz;
    int grid_size;
    orc_entry* entries;
} orc_matrix;

typedef struct {
    int nb_rows, nb_cols, nb_nonzeros;
    int* row_ptr;
    int* col_indices;
    double* values;
} orc_csr;

typedef struct {
    int nb_rows, nb_cols, ell_width, grid_size;
    int* indices;
    int nb_nonzeros;
    double* values;
} orc_ell;

typedef struct {
    int iterations;
    int converged;
    double residual_norm;  /* sqrt(rr) of the last checked iteration */
    double b_norm;         /* ||r0|| (the reference's "b_norm") */
    double solution_sum;
    double solution_norm;
} orc_cg_result;

/* ---- structure ---- */
long long orc_stencil5_nnz(int n);
/* entries in the generator's emission order; out must hold orc_stencil5_nnz(n) */
void orc_stencil5_entries(int n, double center, double neighbour, orc_entry* out);
/* byte-for-byte the reference's file; center_txt / nb_txt are the literal tokens ("5.0", "-1.0") */
int orc_write_mtx_stencil5(int n, const char* filename, const char* center_txt, const char* nb_txt);
/* 0 on success; mat->entries malloc'd */
int orc_load_mtx(const char* filename, orc_matrix* mat);
int orc_build_csr(const orc_matrix* mat, orc_csr* out);
void orc_free_csr(orc_csr* c);
/* closed-form stencil CSR (what generator -> COO->CSR yields), arrays preallocated */
void orc_stencil5_csr_direct(int n, double center, double neighbour, int64_t* row_ptr64,
                             int* col_idx, double* values);
int orc_build_ellpack(const orc_csr* csr, orc_ell* out, int* max_width);
void orc_free_ell(orc_ell* e);
int64_t orc_interior_csr_offset(int64_t row, int grid_size);

/* ---- SpMV ---- */
void orc_csr_spmv(const int* row_ptr, const int* col, const double* val, const double* x,
                  double* y, int n_rows);
void orc_stencil5_spmv(const int* row_ptr, const int* col, const double* val, const double* x,
                       double* y, int n_rows, int grid);
void orc_ell_spmv(const orc_ell* e, const double* x, double* y);
void orc_halo_spmv(const int* row_ptr_local, const int* col_global, const double* val,
                   const double* x_local, const double* x_halo_prev, const double* x_halo_next,
                   double* y, int n_local, int64_t row_offset, int64_t N, int grid);

/* ---- reductions / CG ---- */
double orc_dot_blocktree(int n, const double* x, const double* y);
double orc_dot_sequential(int n, const double* x, const double* y);
/* op: 0 = generic CSR (sequential-k), 1 = stencil5 csr-direct order */
int orc_cg_device(const orc_csr* A, int grid, int op, const double* b, double* x, int max_iters,
                  double tol, orc_cg_result* res, double* rel_hist, int rel_hist_cap);
/* Jacobi-preconditioned CG (not in the reference: parity unpinned), same conventions as orc_cg_device */
int orc_pcg_device(const orc_csr* A, int grid, int op, const double* b, double* x, int max_iters,
                   double tol, orc_cg_result* res);
/* block-Jacobi (one tridiagonal block per grid row, clipped to the bands of a P-rank partition) */
int orc_pcg_block_device(const orc_csr* A, int grid, int op, int P, const double* b, double* x, int max_iters,
                         double tol, orc_cg_result* res);
int orc_cg_mgpu(const orc_csr* A, int grid, int P, const double* b, double* x, int max_iters,
                double tol, orc_cg_result* res);

/* ---- partition ---- */
void orc_partition(int64_t N, int P, int g, int64_t* n_local, int64_t* row_offset);
/* local CSR slice: row_ptr rebased to 0 (n_local+1 entries), col global; returns local nnz */
int64_t orc_local_csr_slice(const orc_csr* A, int64_t row_offset, int64_t n_local, int* row_ptr_out,
                            int* col_out, double* val_out);
/* halo send ranges in local indices: [prev_lo,prev_hi) goes to rank g-1, [next_lo,next_hi) to g+1 */
void orc_halo_ranges(int64_t n_local, int grid, int g, int P, int64_t* prev_lo, int64_t* prev_hi,
                     int64_t* next_lo, int64_t* next_hi);

/* ---- bench statistics ---- */
typedef struct {
    double median_ms, mean_ms, std_dev_ms, min_ms, max_ms;
    int valid_runs, outliers_removed;
} orc_bench_stats;
int orc_bench_stats_from_times(const double* times, int n, orc_bench_stats* st);

int orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
