/*
 * b200/api.h -- the reference's host-facing API, as exported by libspmv_b200.so.
 *
 * Same names, argument meaning, linkage and error behaviour as the reference headers:
 *   get_operator / SPMV_* objects           reference include/spmv.h:137-150, src/spmv/spmv.cu:11-23
 *   build_csr_struct, csr_mat               reference include/spmv_csr.h:47, include/spmv.h:34
 *   ELLPACK builders, ellpack_matrix        reference include/spmv_ellpack.h:50-51, include/spmv.h:37-39
 *   io functions                            reference include/io.h:75-134
 *   cg_solve, cg_solve_device               reference include/solvers/cg_solver.h:59-77
 *   cg_solve_mgpu_partitioned, cg_solve_mgpu reference include/solvers/cg_solver_mgpu_partitioned.h:50-51,
 *                                            include/solvers/cg_solver_mgpu.h:88-89
 *   benchmark_with_stats & co               reference include/benchmark_stats.h:23-29, benchmark_stats_mgpu.h:12-15
 *   metrics / exporters                     reference include/spmv.h:159-183, include/solvers/cg_metrics.h:23-37
 * Linkage follows the reference exactly (extern "C" where its headers say so, C++ linkage for the
 * operator objects, build_csr_struct, build_ellpack_from_csr_struct and the cg_solve* family).
 */
#ifndef B200_API_H
#define B200_API_H

#include "types.h"

#ifdef __cplusplus
#define B200_C_BEGIN extern "C" {
#define B200_C_END }
#else
#define B200_C_BEGIN
#define B200_C_END
#include <stdbool.h>
#endif

/* ---- error macros kept for source compatibility (reference include/spmv.h:46-64) ---- */
#define CUDA_CHECK(call)                                                                      \
    {                                                                                         \
        cudaError_t b200_err_ = (call);                                                       \
        if (b200_err_ != cudaSuccess) {                                                       \
            fprintf(stderr, "CUDA error: %s, line %d\n", cudaGetErrorString(b200_err_), __LINE__); \
            exit(EXIT_FAILURE);                                                               \
        }                                                                                     \
    }

/* ---- operators ---- */
extern SpmvOperator SPMV_CSR;               /* "cusparse-csr" (name kept; hand-written CSR, no cuSPARSE) */
extern SpmvOperator SPMV_STENCIL5_CSR;      /* "stencil5-csr" */
extern SpmvOperator SPMV_STENCIL_HALO_MGPU; /* "stencil5-halo-mgpu": band operator over all visible GPUs */
extern SpmvOperator SPMV_ELLPACK;           /* "ellpack" (declared-only in the reference) */
extern SpmvOperator SPMV_STENCIL5_ELLPACK;  /* "stencil5-ellpack": include/spmv_stencil.h:40-42 */

int build_csr_struct(struct MatrixData* mat);
int build_ellpack_from_csr_struct(const struct CSRMatrix* csr_matrix, ELLPACKMatrix* ellpack_matrix,
                                  int* max_width);

B200_C_BEGIN
extern CSRMatrix csr_mat;
extern ELLPACKMatrix ellpack_matrix;
int build_ellpack_from_csr_local(CSRMatrix* csr_matrix);
int ensure_ellpack_structure_built(MatrixData* mat);

SpmvOperator* get_operator(const char* mode);
void calculate_spmv_metrics(double execution_time_ms, const MatrixData* mat, const char* operator_name,
                            BenchmarkMetrics* metrics);
int get_gpu_properties(BenchmarkMetrics* metrics);
void print_benchmark_metrics(const BenchmarkMetrics* metrics, FILE* output_file);
void print_metrics_json(const BenchmarkMetrics* metrics, FILE* output_file);
void print_metrics_csv(const BenchmarkMetrics* metrics, FILE* output_file);

/* ---- io ---- */
int read_matrix_type(const char* filename);
void read_matrix_general(MatrixData* mat, const char* filename, int* rows, int* cols, int* nnz,
                         int** csr_rowptr, int** csr_colind, double** csr_val);
void read_matrix_symtogen(MatrixData* mat, const char* filename, int* rows, int* cols, int* nnz,
                          int** csr_rowptr, int** csr_colind, double** csr_val, int* nnz_general);
int load_matrix_market(const char* filename, MatrixData* mat);
void convert_csr_to_ellpack(const struct CSRMatrix* csr_matrix, struct ELLPACKMatrix* ellpack_matrix,
                            int* max_width);
int write_matrix_market_stencil5(int n, const char* filename);

/* ---- bench ---- */
int benchmark_with_stats(int (*run_func)(const double*, double*, double*), const double* x, double* y,
                         int num_runs, BenchmarkStats* stats);
int cg_benchmark_with_stats_device(SpmvOperator* spmv_op, MatrixData* mat, double* b, double* x,
                                   CGConfig config, int num_runs, BenchmarkStats* bench_stats,
                                   CGStats* final_stats);
int cg_benchmark_with_stats_mgpu_partitioned(SpmvOperator* spmv_op, MatrixData* mat, double* b, double* x,
                                             CGConfigMultiGPU config, int num_runs,
                                             BenchmarkStats* bench_stats, CGStatsMultiGPU* final_stats);
void export_cg_json(const char* filename, const char* mode, const MatrixData* mat,
                    const BenchmarkStats* bench_stats, const CGStats* cg_stats);
void export_cg_mgpu_json(const char* filename, const char* mode, const MatrixData* mat,
                         const BenchmarkStats* bench_stats, const CGStatsMultiGPU* cg_stats, int num_gpus);
void export_cg_csv(const char* filename, const char* mode, const MatrixData* mat,
                   const BenchmarkStats* bench_stats, const CGStats* cg_stats, bool write_header);
B200_C_END

/* ---- solvers (C++ linkage in the reference) ---- */
int cg_solve(SpmvOperator* spmv_op, MatrixData* mat, const double* b, double* x, CGConfig config,
             CGStats* stats);
int cg_solve_device(SpmvOperator* spmv_op, MatrixData* mat, const double* b, double* x, CGConfig config,
                    CGStats* stats);
/* extension (the reference lists preconditioning as future work, cg_solver.h:6-7):
 * Jacobi-preconditioned CG, M = diag(A); arguments, conventions and statistics of cg_solve_device */
int pcg_solve_device(SpmvOperator* spmv_op, MatrixData* mat, const double* b, double* x, CGConfig config,
                     CGStats* stats);
int cg_solve_mgpu(SpmvOperator* spmv_op, MatrixData* mat, const double* b, double* x,
                  CGConfigMultiGPU config, CGStatsMultiGPU* stats);
int cg_solve_mgpu_partitioned(SpmvOperator* spmv_op, MatrixData* mat, const double* b, double* x,
                              CGConfigMultiGPU config, CGStatsMultiGPU* stats);
/* extension: Jacobi-preconditioned CG over the same row-band partition (arguments of
 * cg_solve_mgpu_partitioned; the edges of the preconditioned direction travel like p does there) */
int pcg_solve_mgpu_partitioned(SpmvOperator* spmv_op, MatrixData* mat, const double* b, double* x,
                               CGConfigMultiGPU config, CGStatsMultiGPU* stats);

#endif /* B200_API_H */
