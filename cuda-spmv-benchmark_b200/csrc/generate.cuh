// generate.cuh -- device-side construction of the stencil matrix, bit-identical to the
// reference's host pipeline generator -> .mtx -> reader -> build_csr_struct
// (src/io/io.cu:322-399,109-171; src/spmv/spmv_cusparse_csr.cu:85-157), without the 45 GB text
// file and the 32 GB Entry[] it needs at n = 20000.  Each kernel fills the slice that belongs to
// the local row band [row_offset, row_offset + n_local).
#pragma once
#include "common.cuh"
#include "stencil_layout.h"

namespace b200 {

struct EntryPOD {
    int row;
    int col;
    double value;
};

// CSR: row_ptr rebased to the band (row_ptr[0] = 0), col_idx GLOBAL, sorted N,W,C,E,S
__global__ void gen_stencil5_csr_kernel(int n, long long row_offset, long long n_local, double center,
                                        double neighbour, int* __restrict__ row_ptr, int* __restrict__ col_idx,
                                        double* __restrict__ values) {
    const long long base = stencil5_nnz_before(row_offset, n);
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t <= n_local;
         t += (long long)gridDim.x * blockDim.x) {
        const long long r = row_offset + t;
        long long k = stencil5_nnz_before(r, n) - base;
        row_ptr[t] = (int)k;
        if (t == n_local) break;
        const long long i = r / n, j = r % n;
        if (i > 0) { col_idx[k] = (int)(r - n); values[k] = neighbour; k++; }
        if (j > 0) { col_idx[k] = (int)(r - 1); values[k] = neighbour; k++; }
        col_idx[k] = (int)r; values[k] = center; k++;
        if (j < n - 1) { col_idx[k] = (int)(r + 1); values[k] = neighbour; k++; }
        if (i < n - 1) { col_idx[k] = (int)(r + n); values[k] = neighbour; k++; }
    }
}

// ELLPACK width 5, row-major, padding (-1, 0.0); slot order = CSR order
__global__ void gen_stencil5_ell_kernel(int n, long long row_offset, long long n_local, double center,
                                        double neighbour, int* __restrict__ indices, double* __restrict__ values) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n_local;
         t += (long long)gridDim.x * blockDim.x) {
        const long long r = row_offset + t, i = r / n, j = r % n;
        int k = 0;
        int* ix = indices + t * 5;
        double* vx = values + t * 5;
        if (i > 0) { ix[k] = (int)(r - n); vx[k] = neighbour; k++; }
        if (j > 0) { ix[k] = (int)(r - 1); vx[k] = neighbour; k++; }
        ix[k] = (int)r; vx[k] = center; k++;
        if (j < n - 1) { ix[k] = (int)(r + 1); vx[k] = neighbour; k++; }
        if (i < n - 1) { ix[k] = (int)(r + n); vx[k] = neighbour; k++; }
        for (; k < 5; k++) { ix[k] = -1; vx[k] = 0.0; }
    }
}

// COO entries in the generator's emission order (Center, Left, Right, Top, Bottom), 0-based
__global__ void gen_stencil5_entries_kernel(int n, long long row_offset, long long n_local, double center,
                                            double neighbour, EntryPOD* __restrict__ out) {
    const long long base = stencil5_nnz_before(row_offset, n);
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n_local;
         t += (long long)gridDim.x * blockDim.x) {
        const long long r = row_offset + t, i = r / n, j = r % n;
        long long k = stencil5_nnz_before(r, n) - base;
        out[k].row = (int)r; out[k].col = (int)r; out[k].value = center; k++;
        if (j > 0) { out[k].row = (int)r; out[k].col = (int)(r - 1); out[k].value = neighbour; k++; }
        if (j < n - 1) { out[k].row = (int)r; out[k].col = (int)(r + 1); out[k].value = neighbour; k++; }
        if (i > 0) { out[k].row = (int)r; out[k].col = (int)(r - n); out[k].value = neighbour; k++; }
        if (i < n - 1) { out[k].row = (int)r; out[k].col = (int)(r + n); out[k].value = neighbour; k++; }
    }
}

__global__ void fill_kernel(double* __restrict__ p, long long n, double v) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n;
         t += (long long)gridDim.x * blockDim.x)
        p[t] = v;
}

}  // namespace b200
