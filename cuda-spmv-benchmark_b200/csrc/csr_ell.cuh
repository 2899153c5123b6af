// csr_ell.cuh -- generic CSR and ELLPACK SpMV for sm_100a (no cuSPARSE).
//
// Replaces cusparseSpMV (src/spmv/spmv_cusparse_csr.cu:246,281) behind the "cusparse-csr"
// operator name, and creates the ELLPACK operator the reference only declares
// (include/spmv_ellpack.h:28-51).  Semantics = the reference's scalar CSR kernel
// (src/solvers/cg_solver_mgpu_partitioned.cu:40-56): sum_k fma(v[k], x[col[k]], sum), k ascending.
//
// Scheme (picked per matrix from a row-length histogram taken on the device at plan time):
//   * STREAM blocks: a CTA owns ROWS consecutive rows whose non-zeros fit its shared-memory
//     window.  Phase 1 streams col_idx / values of the whole block with fully coalesced loads
//     (independent of row boundaries) and gathers x[col]; phase 2 is one thread per row adding
//     its products from shared memory in k order -- bit-identical to the scalar reference order,
//     while HBM only ever sees contiguous 128-byte bursts (the scalar kernel issues 12-byte
//     strided requests per lane).
//   * VECTOR rows: if a block does not fit (long rows), its rows are processed warp-per-row with
//     lanes striding over the row and a fixed-order butterfly sum (order differs from the scalar
//     reference => equal only to rounding, documented tolerance 1e-12 relative).
// ELLPACK (row-major, padding index -1) reuses the stream path with row_ptr[r] = r * width.
#pragma once
#include "common.cuh"

namespace b200 {

struct CsrArgs {
    const int* row_ptr;  // NULL => ELLPACK addressing with `ell_width`
    const int* col_idx;
    const double* values;
    const double* x;
    double* y;
    long long n_rows;
    int ell_width;
    int rows_per_block;
    int window;  // shared-memory capacity in non-zeros
    double alpha, beta;
};

// histogram of row lengths: bin b counts rows with length in (2^(b-1), 2^b], bin 0 = empty/1
__global__ void row_length_histogram_kernel(const int* __restrict__ row_ptr, long long n_rows,
                                            unsigned long long* __restrict__ bins /*[33]*/,
                                            unsigned long long* __restrict__ max_len) {
    __shared__ unsigned int sb[33];
    __shared__ unsigned int smax;
    if (threadIdx.x < 33) sb[threadIdx.x] = 0;
    if (threadIdx.x == 0) smax = 0;
    __syncthreads();
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows;
         r += (long long)gridDim.x * blockDim.x) {
        const unsigned int len = (unsigned int)(row_ptr[r + 1] - row_ptr[r]);
        const int b = len <= 1 ? 0 : 32 - __clz(len - 1);
        atomicAdd(&sb[b], 1u);
        atomicMax(&smax, len);
    }
    __syncthreads();
    if (threadIdx.x < 33 && sb[threadIdx.x]) atomicAdd(&bins[threadIdx.x], (unsigned long long)sb[threadIdx.x]);
    if (threadIdx.x == 0) atomicMax(max_len, (unsigned long long)smax);
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS) csr_adaptive_kernel(const CsrArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* sv = reinterpret_cast<double*>(smem_raw);  // values
    double* sx = sv + a.window;                         // gathered x (NaN-safe skip flag via col<0)
    const long long r0 = (long long)blockIdx.x * a.rows_per_block;
    if (r0 >= a.n_rows) return;
    const long long r1 = min(r0 + (long long)a.rows_per_block, a.n_rows);
    const bool ell = (a.row_ptr == nullptr);
    const long long k0 = ell ? r0 * a.ell_width : (long long)a.row_ptr[r0];
    const long long k1 = ell ? r1 * a.ell_width : (long long)a.row_ptr[r1];
    const long long cnt = k1 - k0;

    if (cnt <= a.window) {
        // ---- STREAM: coalesced sweep over the block's non-zeros
        for (long long k = threadIdx.x; k < cnt; k += THREADS) {
            const int c = __ldcs(a.col_idx + k0 + k);
            const double v = __ldcs(a.values + k0 + k);
            sv[k] = (c >= 0) ? v : 0.0;
            sx[k] = (c >= 0) ? __ldg(a.x + c) : 0.0;  // padding: 0*0, exact no-op under fma
        }
        __syncthreads();
        for (long long r = r0 + threadIdx.x; r < r1; r += THREADS) {
            const long long s = (ell ? r * a.ell_width : (long long)a.row_ptr[r]) - k0;
            const long long e = (ell ? (r + 1) * a.ell_width : (long long)a.row_ptr[r + 1]) - k0;
            double sum = 0.0;
            for (long long k = s; k < e; k++) sum = fma(sv[k], sx[k], sum);
            if (a.beta == 0.0) a.y[r] = a.alpha * sum;
            else a.y[r] = fma(a.alpha, sum, a.beta * a.y[r]);
        }
    } else {
        // ---- VECTOR: warp per row
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (long long r = r0 + warp; r < r1; r += THREADS / 32) {
            const long long s = ell ? r * a.ell_width : (long long)a.row_ptr[r];
            const long long e = ell ? (r + 1) * a.ell_width : (long long)a.row_ptr[r + 1];
            double sum = 0.0;
            for (long long k = s + lane; k < e; k += 32) {
                const int c = a.col_idx[k];
                if (c >= 0) sum = fma(a.values[k], __ldg(a.x + c), sum);
            }
            sum = warp_sum(sum);
            if (lane == 0) {
                if (a.beta == 0.0) a.y[r] = a.alpha * sum;
                else a.y[r] = fma(a.alpha, sum, a.beta * a.y[r]);
            }
        }
    }
}

}  // namespace b200
