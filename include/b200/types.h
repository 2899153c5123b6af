/*
 * b200/types.h -- data structures of the drop-in boundary.
 *
 * Binary-compatible with the reference's public structs so that a caller compiled against the
 * reference headers links against libspmv_b200.so unchanged:
 *   Entry, MatrixData      <- reference include/io.h:43-59
 *   CSRMatrix              <- reference include/spmv_csr.h:28-35
 *   ELLPACKMatrix          <- reference include/spmv_ellpack.h:28-36
 *   SpmvOperator           <- reference include/spmv.h:125-134
 *   BenchmarkMetrics       <- reference include/spmv.h:76-112
 *   CGConfig, CGStats      <- reference include/solvers/cg_solver.h:21-43
 *   CGConfigMultiGPU, CGStatsMultiGPU <- reference include/solvers/cg_solver_mgpu.h:38-71
 *   BenchmarkStats         <- reference include/benchmark_stats.h:12-20
 * Field order and types are the contract; everything else here is new.
 */
#ifndef B200_TYPES_H
#define B200_TYPES_H

#include <stdio.h>
#include <stdlib.h>

#define MAX_LINE_LENGTH 1024
#define MAX_WIDTH 1000

typedef struct {
    int row;      /* 0-based after load */
    int col;      /* 0-based after load */
    double value;
} Entry;

/* COO container produced by load_matrix_market().
 * Extension (B200): entries == NULL together with grid_size > 0 denotes the synthetic
 * 5-point stencil (centre 5.0, neighbours -1.0) of that grid size; operators and solvers
 * then generate their device-side structures directly on the GPU, bit-identical to what
 * generator -> reader -> build_csr_struct would have produced. */
typedef struct MatrixData {
    int rows;
    int cols;
    int nnz;
    int grid_size; /* n of the n x n stencil grid, -1 if unknown */
    Entry* entries;
} MatrixData;

struct CSRMatrix {
    int nb_rows;
    int nb_cols;
    int nb_nonzeros;
    int* row_ptr;
    int* col_indices;
    double* values;
};

struct ELLPACKMatrix {
    int nb_rows;
    int nb_cols;
    int ell_width;
    int grid_size;
    int* indices; /* row-major: slot k of row r at r*ell_width + k; padding slot = -1 */
    int nb_nonzeros;
    double* values; /* same layout; padding slot = 0.0 */
};

#ifdef __cplusplus
typedef struct CSRMatrix CSRMatrix;
typedef struct ELLPACKMatrix ELLPACKMatrix;
#else
typedef struct CSRMatrix CSRMatrix;
typedef struct ELLPACKMatrix ELLPACKMatrix;
#endif

typedef struct {
    double execution_time_ms;
    double gflops;
    double bandwidth_gb_s;
    int matrix_rows;
    int matrix_cols;
    int matrix_nnz;
    int grid_size;
    double sparsity_ratio;
    const char* operator_name;
    double sum_y;
    double norm2_y;
    struct {
        char name[128];
        int memory_mb;
        char compute_capability[16];
        int multiprocessor_count;
        int max_threads_per_block;
        int memory_clock_khz;
        int graphics_clock_mhz;
        int cuda_runtime_version;
        int cuda_driver_version;
        int cusparse_version; /* always 0: this build links no cuSPARSE */
        int current_temp_c;
        int max_temp_c;
        int power_draw_w;
        int power_limit_w;
        char persistence_mode[16];
        char cpu_model[128];
        int system_ram_gb;
        char pcie_generation[16];
        int pcie_link_width;
    } gpu_info;
} BenchmarkMetrics;

typedef struct {
    const char* name;
    int (*init)(MatrixData* mat);
    /* host pointers; does H2D + kernel + D2H, reports the event-timed kernel only */
    int (*run_timed)(const double* x, double* y, double* kernel_time_ms);
    /* device pointers; asynchronous on the default stream */
    int (*run_device)(const double* d_x, double* d_y);
    void (*free)();
} SpmvOperator;

typedef struct {
    int max_iters;
    double tolerance;
    int verbose;
    int enable_detailed_timers;
} CGConfig;

typedef struct {
    int iterations;
    double residual_norm;
    double time_total_ms;
    double time_spmv_ms;
    double time_blas1_ms;
    double time_reductions_ms;
    int converged;
    double solution_sum;
    double solution_norm;
} CGStats;

typedef struct {
    int max_iters;
    double tolerance;
    int verbose;
    int enable_detailed_timers;
} CGConfigMultiGPU;

typedef struct {
    int iterations;
    double residual_norm;
    double time_total_ms;
    double time_spmv_ms;
    double time_blas1_ms;
    double time_reductions_ms;
    double time_allreduce_ms;
    double time_allgather_ms; /* halo exchange */
    int converged;
    double time_dot_rs_initial_ms;
    double time_dot_pAp_ms;
    double time_dot_rs_new_ms;
    double time_axpy_update_x_ms;
    double time_axpy_update_r_ms;
    double time_axpby_update_p_ms;
    double time_initial_r_ms;
    double solution_sum;
    double solution_norm;
} CGStatsMultiGPU;

typedef struct {
    double median_ms;
    double mean_ms;
    double std_dev_ms;
    double min_ms;
    double max_ms;
    int valid_runs;
    int outliers_removed;
} BenchmarkStats;

#endif /* B200_TYPES_H */
