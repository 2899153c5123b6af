// host_common.h -- shared declarations of the host layer (C++ above the kernel C ABI).
// The host layer mirrors the reference's operator / solver / io / bench interface
// (include/b200/api.h) and talks to the GPU only through include/b200_kernels.h and the CUDA
// runtime API (allocation, copies, streams, events).  It contains no arithmetic on vectors or
// matrices: there is no CPU fallback anywhere in the product path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/b200/api.h"
#include "../../include/b200_kernels.h"

// CUDA failure inside the library: report and return an error code to the caller.  (The
// reference's CUDA_CHECK calls exit(); a shared library that is also loaded from Python must not
// take the process down, so the CLIs exit on a non-zero return instead.)
#define B200_CUDA(call)                                                                          \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess) {                                                                 \
            fprintf(stderr, "[b200] CUDA error: %s (%s:%d)\n", cudaGetErrorString(e_), __FILE__, \
                    __LINE__);                                                                   \
            return 2;                                                                            \
        }                                                                                        \
    } while (0)

#define B200_K(call)                                                                             \
    do {                                                                                         \
        int rc_ = (call);                                                                        \
        if (rc_ != 0) {                                                                          \
            fprintf(stderr, "[b200] kernel ABI error %d: %s (%s:%d)\n", rc_, b200_last_error(),  \
                    __FILE__, __LINE__);                                                         \
            return rc_;                                                                          \
        }                                                                                        \
    } while (0)

namespace b200host {

// Is `mat` the synthetic stencil of the B200 extension (entries == NULL, grid_size > 0)?
inline bool is_synthetic(const MatrixData* m) { return m && m->entries == nullptr && m->grid_size > 0; }
// Row count as a 64-bit number.  MatrixData.rows is a 32-bit int (reference include/io.h:50-56); the
// synthetic stencil is defined by its grid size alone, so it may exceed 2^31 rows (multi-GPU only:
// a band still holds fewer than 2^31 non-zeros) -- rows / cols / nnz are then saturated and ignored.
inline long long rows64(const MatrixData* m) {
    return is_synthetic(m) ? (long long)m->grid_size * m->grid_size : (long long)m->rows;
}

// Device-resident matrix slice of one operator (or one band of the multi-GPU solver).
struct DeviceBand {
    int* d_row_ptr = nullptr;
    int* d_col_idx = nullptr;
    double* d_values = nullptr;
    long long values_len = 0;
    long long row_offset = 0;
    long long n_local = 0;
    long long nnz_local = 0;
    int grid = -1;
    int layout = 0;
    void release();
    void describe(b200_band* out) const;
};

// Builds the device arrays of rows [off, off+nl): from the host CSR in csr_mat (uploaded slice,
// row_ptr rebased, global columns) or generated on the device for the synthetic stencil.
int upload_band_csr(const MatrixData* mat, long long off, long long nl, DeviceBand* out, cudaStream_t s);
int upload_band_ell(const MatrixData* mat, long long off, long long nl, DeviceBand* out, cudaStream_t s);

// operator side-table: the band behind a stencil-type operator (for the fused CG path)
// generic CSR / ELLPACK operators of this library: y = A x with the x.y partials in the same launch
// (0 = done, -1 = not available for this operator: use run_device + b200_dot_partials)
int operator_spmv_dot(const SpmvOperator* op, const double* d_x, double* d_y, double* d_partials, long long cap, int* np,
                      const void* scalars);
const DeviceBand* operator_band(const SpmvOperator* op);
const DeviceBand* operator_matrix(const SpmvOperator* op, int* ell_width);

// verbose printing switch shared by the solvers
extern int g_quiet;

}  // namespace b200host

extern "C" {
// ---- extensions exported next to the reference API (declared here, documented in INTEGRATION.md)
int b200_operator_band(const SpmvOperator* op, b200_band* out);
int b200_mgpu_init_single_process(int world, const int* devices, int max_grid);
int b200_mgpu_init_rank(int rank, int world, int device, int max_grid, void* handle_out64);
int b200_mgpu_connect(const void* handles);
int b200_mgpu_world(void);
int b200_mgpu_rank(void);
void b200_mgpu_finalize(void);
MatrixData b200_synthetic_stencil(int grid_size);
int b200_set_tuning(int variant, int rows_per_item);
void b200_get_tuning(int* variant, int* rows_per_item);
int b200_last_phase_times(double* ms9, int* count9);
long long b200_last_h2d_bytes(void);
int b200_cg_set_skip_zero_x0(int on);
int b200_host_all_zero(const double* p, long long n, int threads);
int b200_pcg_set_preconditioner(int kind);
int b200_last_tail_times(double* ms8, int* count8);
int b200_last_gap_times(double* ms8);
int b200_mgpu_halo_probe(int reps, double* us_per_push, long long* bytes_per_direction);

// device-resident ingest: .mtx -> COO on the GPU -> operator (no host Entry[] / CSR)
int b200_load_matrix_market_device(const char* filename, MatrixData* meta, void** d_entries_out);
int b200_operator_init_device_coo(SpmvOperator* op, const MatrixData* meta, const void* d_entries);
int b200_operator_device_csr(SpmvOperator* op, const int** d_row_ptr, const int** d_col_idx, const double** d_values,
                             long long* nnz);
void b200_free_device(void* d_ptr);
int b200_copy_to_host(void* h_dst, const void* d_src, size_t bytes);
// pinned host buffers on the GPU's own NUMA node (host/host_alloc.cpp)
int b200_host_node_of_device(int device);
int b200_host_alloc_near(int device, size_t bytes, void** out, int* node_out);
void b200_host_free(void* p);
}
