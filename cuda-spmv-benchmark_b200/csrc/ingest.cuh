// ingest.cuh -- device-side Matrix Market entry parsing and COO -> CSR, bit-identical to the
// reference's host pipeline (reader: src/io/io.cu:153-166, "%d %d %le" + 1->0-based shift;
// builder: src/spmv/spmv_cusparse_csr.cu:85-157, count / prefix-sum / scatter in file order /
// per-row stable sort by column).  SURVEY.md section 8(f) item 1.
//
// Parsing: the text is cut into fixed chunks; pass 1 counts the lines that start in each chunk,
// a scan turns the counts into entry indices, pass 2 parses every line in place.  Doubles take
// Clinger's exact fast path (<= 19 significant digits folded into a 64-bit mantissa that fits
// 2^53, |decimal exponent| <= 22: one correctly-rounded multiply or divide, identical to strtod);
// anything else is flagged and re-read by the host with strtod (I/O corner case, not compute).
//
// COO -> CSR: histogram of rows (atomics), exclusive scan, scatter tagged with the entry's file
// position, then a per-row sort by (column, file position) -- which reproduces the reference's
// "scatter in file order, stable insertion sort by column" for any input order and duplicates.
#pragma once
#include "common.cuh"
#include "generate.cuh"

namespace b200 {

// ---------------------------------------------------------------- exclusive scan (3 kernels)
constexpr int SCAN_TILE = 2048;  // 256 threads x 8

template <typename T>
__global__ void __launch_bounds__(256) scan_tiles_kernel(const T* __restrict__ in, long long* __restrict__ out,
                                                         long long n, long long* __restrict__ tile_sums) {
    __shared__ long long warp_tot[8];
    const long long base = (long long)blockIdx.x * SCAN_TILE + threadIdx.x * 8;
    long long v[8], run = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        v[i] = (base + i < n) ? (long long)in[base + i] : 0;
        run += v[i];
    }
    // inclusive scan of the per-thread totals over the CTA
    long long incl = run;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long t = __shfl_up_sync(B200_FULL, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[w] = incl;
    __syncthreads();
    long long woff = 0;
    for (int i = 0; i < w; i++) woff += warp_tot[i];
    long long excl = woff + incl - run;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        if (base + i < n) out[base + i] = excl;
        excl += v[i];
    }
    if (threadIdx.x == 255) tile_sums[blockIdx.x] = woff + incl;
}

__global__ void __launch_bounds__(1024) scan_tile_sums_kernel(long long* __restrict__ tile_sums, long long n_tiles,
                                                              long long* __restrict__ total) {
    __shared__ long long warp_tot[32];
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (long long base = 0; base < n_tiles; base += 1024) {
        const long long i = base + threadIdx.x;
        const long long v = i < n_tiles ? tile_sums[i] : 0;
        long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long t = __shfl_up_sync(B200_FULL, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_tot[w] = incl;
        __syncthreads();
        long long woff = 0;
        for (int k = 0; k < w; k++) woff += warp_tot[k];
        const long long c = carry;
        if (i < n_tiles) tile_sums[i] = c + woff + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = c + woff + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = carry;
}

__global__ void __launch_bounds__(256) scan_add_offsets_kernel(long long* __restrict__ out, long long n,
                                                               const long long* __restrict__ tile_sums) {
    const long long off = tile_sums[blockIdx.x];
    const long long base = (long long)blockIdx.x * SCAN_TILE + threadIdx.x * 8;
#pragma unroll
    for (int i = 0; i < 8; i++)
        if (base + i < n) out[base + i] += off;
}

// ---------------------------------------------------------------- COO -> CSR
// cols > 0: column ids are range-checked as well (a malformed file must not make the x gather of the
// SpMV kernels read out of bounds); bad: 1 = row, 2 = column outside the matrix
__global__ void coo_count_rows_kernel(const EntryPOD* __restrict__ e, long long nnz, int rows, int cols,
                                      int* __restrict__ counts, int* __restrict__ bad) {
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += (long long)gridDim.x * blockDim.x) {
        const int r = e[k].row, c = e[k].col;
        if (r < 0 || r >= rows) { atomicExch(bad, 1); continue; }
        if (c < 0 || (cols > 0 && c >= cols)) { atomicMax(bad, 2); }
        atomicAdd(&counts[r], 1);
    }
}

// values re-read on the host (literals outside the exact fast path of the device parser): one staged
// upload of (entry index, value bits) pairs, one launch
__global__ void patch_entry_values_kernel(EntryPOD* __restrict__ e, const long long* __restrict__ pairs, int n) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x)
        e[pairs[2 * k]].value = __longlong_as_double(pairs[2 * k + 1]);
}

__global__ void coo_finish_row_ptr_kernel(const long long* __restrict__ scan, long long total, int rows,
                                          int* __restrict__ row_ptr, int* __restrict__ cursor) {
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r <= rows; r += (long long)gridDim.x * blockDim.x) {
        const int v = (r < rows) ? (int)scan[r] : (int)total;
        row_ptr[r] = v;
        if (r < rows) cursor[r] = v;
    }
}

__global__ void coo_scatter_kernel(const EntryPOD* __restrict__ e, long long nnz, int rows, int* __restrict__ cursor,
                                   int* __restrict__ col, double* __restrict__ val, int* __restrict__ pos) {
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += (long long)gridDim.x * blockDim.x) {
        const int r = e[k].row;
        if (r < 0 || r >= rows) continue;
        const int d = atomicAdd(&cursor[r], 1);
        col[d] = e[k].col;
        val[d] = e[k].value;
        pos[d] = (int)k;  // file position: tie-break that restores the reference's scatter order
    }
}

__device__ __forceinline__ bool key_less(int ca, int pa, int cb, int pb) { return ca < cb || (ca == cb && pa < pb); }

// rows up to SHORT entries: one thread, insertion sort on (col, pos)
// longer rows: one warp, rank sort through the scratch arrays (O(len^2), rows that long are rare)
template <int SHORT>
__global__ void __launch_bounds__(256) csr_sort_rows_kernel(const int* __restrict__ row_ptr, int rows, int* __restrict__ col,
                                                            double* __restrict__ val, int* __restrict__ pos,
                                                            int* __restrict__ col_tmp, double* __restrict__ val_tmp) {
    const int lane = threadIdx.x & 31;
    const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long r0 = warp_global * 32;
    if (r0 >= rows) return;
    const long long r = r0 + lane;
    int s = 0, e = 0;
    if (r < rows) { s = row_ptr[r]; e = row_ptr[r + 1]; }
    if (e - s <= SHORT) {
        for (int a = s + 1; a < e; a++) {
            const int c = col[a], p = pos[a];
            const double v = val[a];
            int b = a;
            while (b > s && key_less(c, p, col[b - 1], pos[b - 1])) {
                col[b] = col[b - 1]; val[b] = val[b - 1]; pos[b] = pos[b - 1];
                b--;
            }
            col[b] = c; val[b] = v; pos[b] = p;
        }
    }
    // long rows of this 32-row group, cooperatively
    const unsigned long_mask = __ballot_sync(B200_FULL, (e - s) > SHORT);
    for (int q = 0; q < 32; q++) {
        if (!((long_mask >> q) & 1u)) continue;
        const int qs = __shfl_sync(B200_FULL, s, q), qe = __shfl_sync(B200_FULL, e, q);
        for (int a = qs + lane; a < qe; a += 32) {
            const int c = col[a], p = pos[a];
            int rank = 0;
            for (int b = qs; b < qe; b++) rank += key_less(col[b], pos[b], c, p) ? 1 : 0;
            col_tmp[qs + rank] = c;
            val_tmp[qs + rank] = val[a];
        }
        __syncwarp();
        for (int a = qs + lane; a < qe; a += 32) { col[a] = col_tmp[a]; val[a] = val_tmp[a]; }
        __syncwarp();
    }
}

// ---------------------------------------------------------------- Matrix Market entry text
constexpr int PARSE_CHUNK = 2048;

__device__ __forceinline__ bool is_space(unsigned char c) {
    return c == ' ' || c == '\n' || c == '\t' || c == '\r' || c == '\v' || c == '\f';
}

// lines that START inside chunk j (a line starts at byte 0 or right after a '\n'); empty / blank
// lines do not count (fscanf skips white space between tokens)
__device__ __forceinline__ bool line_is_blank(const unsigned char* t, long long i, long long n) {
    while (i < n && t[i] != '\n') {
        if (!is_space(t[i])) return false;
        i++;
    }
    return true;
}

__global__ void mtx_count_lines_kernel(const unsigned char* __restrict__ t, long long n, int* __restrict__ counts) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long lo = j * PARSE_CHUNK;
    if (lo >= n) return;
    const long long hi = min(lo + (long long)PARSE_CHUNK, n);
    int c = 0;
    for (long long i = lo; i < hi; i++) {
        const bool starts = (i == 0) || (t[i - 1] == '\n');
        if (starts && !line_is_blank(t, i, n)) c++;
    }
    counts[j] = c;
}

__device__ __constant__ double kPow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                             1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

__device__ __forceinline__ long long parse_int_tok(const unsigned char* t, long long& i, long long n, bool& ok) {
    while (i < n && is_space(t[i])) i++;
    bool neg = false;
    if (i < n && (t[i] == '-' || t[i] == '+')) { neg = t[i] == '-'; i++; }
    long long v = 0;
    int digits = 0;
    while (i < n && t[i] >= '0' && t[i] <= '9') { v = v * 10 + (t[i] - '0'); i++; digits++; }
    ok = digits > 0;
    return neg ? -v : v;
}

// returns true when the literal was converted exactly; false => host must redo it with strtod
__device__ __forceinline__ bool parse_double_tok(const unsigned char* t, long long& i, long long n, double& out, bool& ok) {
    while (i < n && is_space(t[i])) i++;
    bool neg = false;
    if (i < n && (t[i] == '-' || t[i] == '+')) { neg = t[i] == '-'; i++; }
    unsigned long long m = 0;
    int sig = 0, dropped = 0, frac = 0, digits = 0;
    bool seen_dot = false, inexact = false;
    while (i < n) {
        const unsigned char c = t[i];
        if (c >= '0' && c <= '9') {
            digits++;
            if (sig < 19) {
                m = m * 10 + (c - '0');
                if (m != 0) sig++;
                if (seen_dot) frac++;
            } else {
                if (c != '0') inexact = true;
                if (!seen_dot) dropped++;
            }
            i++;
        } else if (c == '.' && !seen_dot) {
            seen_dot = true;
            i++;
        } else {
            break;
        }
    }
    ok = digits > 0;
    int e10 = 0;
    if (i < n && (t[i] == 'e' || t[i] == 'E')) {
        long long k = i + 1;
        bool eneg = false;
        if (k < n && (t[k] == '-' || t[k] == '+')) { eneg = t[k] == '-'; k++; }
        int ed = 0, ev = 0;
        while (k < n && t[k] >= '0' && t[k] <= '9') { if (ev < 100000) ev = ev * 10 + (t[k] - '0'); k++; ed++; }
        if (ed > 0) { e10 = eneg ? -ev : ev; i = k; }
    }
    // anything glued to the number (inf, nan, hex floats, ...) goes to the host
    if (i < n && !is_space(t[i])) { while (i < n && !is_space(t[i])) i++; inexact = true; }
    const int exp10 = e10 - frac + dropped;
    double v;
    bool exact = !inexact && m <= (1ull << 53);
    if (m == 0) { v = 0.0; exact = !inexact; }
    else if (exact && exp10 >= 0 && exp10 <= 22) v = (double)m * kPow10[exp10];
    else if (exact && exp10 < 0 && exp10 >= -22) v = (double)m / kPow10[-exp10];
    else { v = 0.0; exact = false; }
    out = neg ? -v : v;
    return exact;
}

__global__ void mtx_parse_lines_kernel(const unsigned char* __restrict__ t, long long n,
                                       const long long* __restrict__ first_line, long long max_entries,
                                       EntryPOD* __restrict__ out, int* __restrict__ n_inexact,
                                       long long* __restrict__ inexact_list /* [entry, byte offset] pairs */,
                                       int inexact_cap, int* __restrict__ bad) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long lo = j * PARSE_CHUNK;
    if (lo >= n) return;
    const long long hi = min(lo + (long long)PARSE_CHUNK, n);
    long long k = first_line[j];
    for (long long i = lo; i < hi; i++) {
        const bool starts = (i == 0) || (t[i - 1] == '\n');
        if (!starts || line_is_blank(t, i, n)) continue;
        if (k < max_entries) {
            long long p = i;
            bool ok1, ok2, ok3;
            const long long r = parse_int_tok(t, p, n, ok1);
            const long long c = parse_int_tok(t, p, n, ok2);
            const long long vpos = p;
            double v;
            const bool exact = parse_double_tok(t, p, n, v, ok3);
            // the reference reads white-space separated tokens, not lines: anything but exactly three
            // tokens on a line sends the whole file to the host reader
            while (p < n && t[p] != '\n') { if (!is_space(t[p])) ok3 = false; p++; }
            if (!(ok1 && ok2 && ok3)) atomicExch(bad, 1);
            out[k].row = (int)r - 1;  // Matrix Market is 1-based
            out[k].col = (int)c - 1;
            out[k].value = v;
            if (!exact) {
                const int slot = atomicAdd(n_inexact, 1);
                if (slot < inexact_cap) { inexact_list[2 * slot] = k; inexact_list[2 * slot + 1] = vpos; }
            }
        }
        k++;
    }
}

}  // namespace b200
