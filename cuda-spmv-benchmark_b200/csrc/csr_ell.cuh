// csr_ell.cuh -- generic CSR and ELLPACK SpMV for sm_100a (no cuSPARSE).
//
// Replaces cusparseSpMV (src/spmv/spmv_cusparse_csr.cu:246,281) behind the "cusparse-csr"
// operator name, and creates the ELLPACK operator the reference only declares
// (include/spmv_ellpack.h:28-51).  Semantics = the reference's scalar CSR kernel
// (src/solvers/cg_solver_mgpu_partitioned.cu:40-56): sum_k fma(v[k], x[col[k]], sum), k ascending.
//
// Scheme ("warp-stream", chosen per 32-row group from the row lengths; the per-matrix row-length
// histogram taken at plan time sets the long-row threshold):
//   * A WARP owns 32 consecutive rows.  Their non-zeros are one contiguous range of col_idx /
//     values, which the warp sweeps in windows of 256 entries with perfectly coalesced loads
//     (lane k, k+32, ... -- independent of where rows begin), gathers x[col] for all of them at
//     once (8 independent loads per lane in flight) and parks value and x side by side in its
//     private shared-memory slice.  Then each lane adds the products of ITS row that fall into the
//     window, in k order, carrying the running sum across windows.  The result is bit-identical to
//     the scalar reference order, there is no block-wide barrier (only __syncwarp), and HBM only
//     sees full-line bursts -- the scalar kernel issues strided 12-byte requests per lane.
//   * Groups whose rows are long (more than `vector_threshold` entries per row on average) are
//     processed warp-per-row instead: lanes stride over the row, fixed-order butterfly sum (order
//     differs from the scalar reference => equal to rounding only, documented tolerance 1e-12).
// ELLPACK (row-major, padding index -1) is the same kernel with row_ptr[r] = r * width.
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
    int vector_threshold;  // mean entries per row of a 32-row group above which it goes warp-per-row
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

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) csr_warp_stream_kernel(const CsrArgs a) {
    constexpr int WIN = 256;  // entries per window = 8 per lane
    __shared__ double sv[WARPS][WIN];
    __shared__ double sx[WARPS][WIN];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long r0 = ((long long)blockIdx.x * WARPS + warp) * 32;
    if (r0 >= a.n_rows) return;
    const long long r = r0 + lane;
    const bool live = r < a.n_rows;
    const bool ell = (a.row_ptr == nullptr);
    long long s = 0, e = 0;
    if (live) {
        s = ell ? r * a.ell_width : (long long)__ldg(a.row_ptr + r);
        e = ell ? (r + 1) * a.ell_width : (long long)__ldg(a.row_ptr + r + 1);
    }
    const long long k_begin = __shfl_sync(B200_FULL, s, 0);
    long long k_end = live ? e : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const long long t = __shfl_xor_sync(B200_FULL, k_end, o);
        k_end = t > k_end ? t : k_end;
    }
    const long long rows_here = min(32LL, a.n_rows - r0);
    double sum = 0.0;

    if (k_end - k_begin <= (long long)a.vector_threshold * rows_here) {
        // ---- stream: windows of 256 entries, lane-per-row accumulation in k order
        double* mv = sv[warp];
        double* mx = sx[warp];
        for (long long w = k_begin; w < k_end; w += WIN) {
            int c[8];
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const long long k = w + u * 32 + lane;
                c[u] = -1;
                v[u] = 0.0;
                if (k < k_end) {
                    c[u] = __ldcs(a.col_idx + k);
                    v[u] = __ldcs(a.values + k);
                }
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                // padding / out-of-window: 0 * 0, an exact no-op under fma
                const double xv = (c[u] >= 0) ? __ldg(a.x + c[u]) : 0.0;
                mv[u * 32 + lane] = (c[u] >= 0) ? v[u] : 0.0;
                mx[u * 32 + lane] = xv;
            }
            __syncwarp();
            const long long lo = s > w ? s : w;
            const long long hi = e < w + WIN ? e : w + WIN;
            for (long long k = lo; k < hi; k++) sum = fma(mv[k - w], mx[k - w], sum);
            __syncwarp();
        }
        if (live) {
            if (a.beta == 0.0) a.y[r] = a.alpha * sum;
            else a.y[r] = fma(a.alpha, sum, a.beta * a.y[r]);
        }
    } else {
        // ---- vector: the warp walks its rows one by one
        for (int q = 0; q < rows_here; q++) {
            const long long qs = __shfl_sync(B200_FULL, s, q), qe = __shfl_sync(B200_FULL, e, q);
            double part = 0.0;
            for (long long k = qs + lane; k < qe; k += 32) {
                const int cc = a.col_idx[k];
                if (cc >= 0) part = fma(a.values[k], __ldg(a.x + cc), part);
            }
            part = warp_sum(part);
            if (lane == q) sum = part;
        }
        if (live) {
            if (a.beta == 0.0) a.y[r] = a.alpha * sum;
            else a.y[r] = fma(a.alpha, sum, a.beta * a.y[r]);
        }
    }
}

}  // namespace b200
