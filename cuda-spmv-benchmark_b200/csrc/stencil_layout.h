// stencil_layout.h -- closed-form layout of the 5-point stencil matrix, shared by host and
// device code.  All quantities are 64-bit: n = 20000 puts nnz at 93 % of INT_MAX and the weak
// scaling configs exceed it (the reference's int arithmetic, spmv_stencil_csr_direct.cu:50-67,
// overflows beyond n = 20724).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define B200_HD __host__ __device__ __forceinline__
#else
#define B200_HD inline
#endif

namespace b200 {

// total non-zeros of the n x n stencil (reference: counting loop, src/io/io.cu:327-340)
B200_HD long long stencil5_nnz(long long n) { return 5 * n * n - 4 * n; }

// non-zeros of matrix row r = i*n + j
B200_HD int stencil5_row_nnz(long long i, long long j, long long n) {
    return 1 + (i > 0) + (i < n - 1) + (j > 0) + (j < n - 1);
}

// number of non-zeros in all rows < r  (== row_ptr[r] of the full CSR matrix, and == the index of
// the first COO entry the generator emits for point r, src/io/io.cu:362-392)
B200_HD long long stencil5_nnz_before(long long r, long long n) {
    if (n == 1) return r > 0 ? 1 : 0;
    const long long i = r / n, j = r % n;
    long long k = 0;
    if (i > 0) k = (4 * n - 2) + (i - 1) * (5 * n - 2);  // grid row 0, then i-1 middle rows
    if (i >= n) return stencil5_nnz(n);
    const long long per = 5 - (i == 0) - (i == n - 1);
    k += j * per - (j > 0 ? 1 : 0);  // column 0 has one neighbour less
    return k;
}

// element index of the first stored value (the north coefficient) of interior point (i,j) in the
// FULL CSR values array; equals the reference's calculate_interior_csr_offset
B200_HD long long stencil5_interior_offset(long long i, long long j, long long n) {
    return (4 * n - 2) + (i - 1) * (5 * n - 2) + 4 + (j - 1) * 5;
}

}  // namespace b200
