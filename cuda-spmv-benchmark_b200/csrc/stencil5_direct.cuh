// stencil5_direct.cuh -- plain y = A x for the 5-point stencil, "sequential sweep" form.
//
// Why a second SpMV kernel next to stencil5.cuh: measured on the same B200, the reference's
// one-thread-per-row kernel (src/spmv/spmv_stencil_csr_direct.cu:76-123), recompiled for sm_100, runs
// the 10k x 10k product in 0.794 ms = 7.06 TB/s of algorithmic traffic, the bulk-copy ring of
// stencil5.cuh in 0.833 ms = 6.72 TB/s (tests/golden/ref_gpu_10k.json, profiles/).  The ring tiles the
// grid into 8-row x 128-column items: each of the ~2400 resident warps streams its own 5 KB pieces,
// 400 KB apart from row to row.  A sweep in which consecutive threads own consecutive rows makes the
// whole GPU read ONE contiguous window of `values` that moves forward through the array -- the access
// order HBM likes best -- and the north / south x rows come out of L2 (written ~n rows earlier) instead
// of registers.  For the un-fused product that wins; the fused CG kernels (7 streams, x-row reuse in
// registers, no L2 re-reads of r and p_old) stay on the ring.
//
// This kernel keeps the arithmetic of the reference bit for bit (t = vC*xC; fma(vW,xW,t); fma(vE,xE,t);
// fma(vN,xN,t); fma(vS,xS,t); boundary rows: fma chain over the CSR row from 0.0) and differs in shape:
// ROWS consecutive rows per thread (their 5*ROWS coefficients are one contiguous span: 128-bit loads
// where the span is 16-byte aligned), x_C / x_W / x_E of the thread's rows from ROWS + 2 loads instead of
// 3*ROWS, streaming stores for y, band + halo addressing as in stencil5.cuh.
#pragma once
#include "common.cuh"
#include "stencil5.cuh"

namespace b200 {

template <int ROWS, bool CG_LOADS>
__global__ void __launch_bounds__(256) stencil5_direct_kernel(const Stencil5Args a) {
    // 32-bit index arithmetic wherever a band allows it (a band has < 2^31 / 5 rows): one thread per row leaves
    // ~9 issue slots per byte-time, a 64-bit division alone would eat most of them
    const unsigned int t = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int nloc = (unsigned int)a.n_local, n = (unsigned int)a.n;
    const unsigned int lr0 = t * ROWS;  // first local row of this thread
    if (lr0 >= nloc) return;
    // grid coordinates of local row 0 (64-bit once per thread, cheap: constant operands), then 32-bit
    const unsigned int i_off = (unsigned int)(a.row_offset / a.n), j_off = (unsigned int)(a.row_offset - (long long)i_off * a.n);
    const unsigned int q = (j_off + lr0) / n;
    const unsigned int i = i_off + q, j0 = j_off + lr0 - q * n;
    // fast path: all ROWS rows are interior points of the same grid row and lie inside the band
    const bool fast = (i >= 1) && (i + 2 <= n) && (j0 >= 1) && (j0 + ROWS + 1 <= n) && (lr0 + ROWS <= nloc);
    if (fast) {
        const double* v = a.values + (a.base0 + (long long)i * a.row_stride + 5 * (long long)j0);
        double c[5 * ROWS];
        // read-only path WITH L1 allocation: a thread's 40-byte run shares its 32-byte sectors with its neighbours'
        // runs, and the five loads of a warp cover the same 1280 contiguous bytes -- streaming loads (ld.cs)
        // would pull every sector from L2 several times
#pragma unroll
        for (int k = 0; k < 5 * ROWS; k++) c[k] = __ldg(v + k);
        // x(i, j0-1 .. j0+ROWS): one more than the rows on either side
        double xc[ROWS + 2], xn[ROWS], xs[ROWS];
        if (lr0 >= n && lr0 + ROWS + n <= nloc) {  // every neighbour is a local element: plain loads
            const double* xp = a.x + lr0;
#pragma unroll
            for (int k = 0; k < ROWS + 2; k++) xc[k] = __ldg(xp - 1 + k);
#pragma unroll
            for (int k = 0; k < ROWS; k++) {
                xn[k] = __ldg(xp + k - (long long)n);
                xs[k] = __ldg(xp + k + n);
            }
        } else {  // first / last grid rows of a band: halo addressing
#pragma unroll
            for (int k = 0; k < ROWS + 2; k++) xc[k] = x_at<ST_PLAIN, CG_LOADS>(a, (long long)lr0 - 1 + k);
#pragma unroll
            for (int k = 0; k < ROWS; k++) {
                xn[k] = x_at<ST_PLAIN, CG_LOADS>(a, (long long)lr0 + k - n);
                xs[k] = x_at<ST_PLAIN, CG_LOADS>(a, (long long)lr0 + k + n);
            }
        }
#pragma unroll
        for (int k = 0; k < ROWS; k++) {
            const double* cc = c + 5 * k;  // N, W, C, E, S
            double s = cc[2] * xc[k + 1];
            s = fma(cc[1], xc[k], s);
            s = fma(cc[3], xc[k + 2], s);
            s = fma(cc[0], xn[k], s);
            s = fma(cc[4], xs[k], s);
            a.y[lr0 + k] = s;
        }
        return;
    }
    // boundary rows of the grid (and the ragged end of a band): CSR walk, reference order
#pragma unroll 1
    for (int k = 0; k < ROWS; k++) {
        const long long lr = (long long)lr0 + k;
        if (lr >= a.n_local) break;
        const long long r = a.row_offset + lr;
        const long long gi = r / a.n, gj = r - gi * a.n;
        if (gi >= 1 && gi <= a.n - 2 && gj >= 1 && gj <= a.n - 2) {  // interior row next to a boundary one
            const double* v = a.values + (a.base0 + gi * a.row_stride + 5 * gj);
            double s = v[2] * x_at<ST_PLAIN, CG_LOADS>(a, lr);
            s = fma(v[1], x_at<ST_PLAIN, CG_LOADS>(a, lr - 1), s);
            s = fma(v[3], x_at<ST_PLAIN, CG_LOADS>(a, lr + 1), s);
            s = fma(v[0], x_at<ST_PLAIN, CG_LOADS>(a, lr - a.n), s);
            s = fma(v[4], x_at<ST_PLAIN, CG_LOADS>(a, lr + a.n), s);
            a.y[lr] = s;
        } else {
            (void)boundary_row<ST_PLAIN, CG_LOADS>(a, r, XAlphas<ST_PLAIN>(), 0.0);
        }
    }
}


// ---- the same sweep for the fused CG passes (ST_DOT, ST_RESID, ST_FUSED) -------------------------------------
// One thread per row, SWEEP_TILES consecutive 256-row tiles per CTA (one partial sum per CTA), CTAs in row
// order: the whole GPU streams one contiguous window of `values`, r, p_old, x that moves forward through the
// arrays.  The north / south neighbours of r and p_old come out of L2 (they were streamed n rows earlier),
// west / east out of L1: 96 B/row of L2 -> SM traffic against 72 B/row for the ring, in exchange for the
// DRAM access order (measured on the plain product: 7.2 against 7.0 TB/s).
// Band mode: tiles that read a halo wait for the neighbour's arrival word and are mapped to the LAST CTAs of
// the grid (the mapping is keyed on the halo pointers, not on the flags, so that the order of the partial
// sums does not depend on whether a launch has to wait).
// A/B switches (tools: build a second library with -D..., load it with B200_LIB_PATH): streaming stores for the
// result vectors, evict-first loads for the coefficient stream
#ifndef B200_SWEEP_STCS
#define B200_SWEEP_STCS 0
#endif
#ifndef B200_SWEEP_LDCS
#define B200_SWEEP_LDCS 0
#endif
__device__ __forceinline__ void sweep_store(double* p, double v) {
#if B200_SWEEP_STCS
    __stcs(p, v);
#else
    *p = v;
#endif
}
__device__ __forceinline__ double sweep_coef(const double* p) {
#if B200_SWEEP_LDCS
    return __ldcs(p);
#else
    return __ldg(p);
#endif
}
// west / east neighbours of the fast path from the neighbouring lanes' centre elements (shuffle) instead of two
// more (L1-hit) loads per vector: 11 instead of 15 load instructions per row in the fused modes.  Correct, and
// slower (20k x 20k CG 92.95 -> 98.9 ms): the shuffles wait for the centre loads, and everything behind them in
// program order waits with them -- the independent L1-hit loads are the cheaper way.  Kept as an A/B switch.
#ifndef B200_SWEEP_SHFL
#define B200_SWEEP_SHFL 0
#endif
#ifndef B200_SWEEP_TILES
#define B200_SWEEP_TILES 2
#endif
constexpr int SWEEP_TILES = B200_SWEEP_TILES;

template <int MODE, bool CG_LOADS>
__device__ __forceinline__ double sweep_row_generic(const Stencil5Args& a, long long lr, const XAlphas<MODE>& xa, double beta) {
    // interior grid point whose neighbours are not all plain local elements: halo / band-edge addressing
    const long long r = a.row_offset + lr;
    const long long gi = r / a.n, gj = r - gi * a.n;
    const double* v = a.values + (a.base0 + gi * a.row_stride + 5 * gj);
    double po = 0.0;
    const double xc = x_at<MODE, CG_LOADS>(a, lr, beta, &po);
    double t = v[2] * xc;
    t = fma(v[1], x_at<MODE, CG_LOADS>(a, lr - 1, beta), t);
    t = fma(v[3], x_at<MODE, CG_LOADS>(a, lr + 1, beta), t);
    t = fma(v[0], x_at<MODE, CG_LOADS>(a, lr - a.n, beta), t);
    t = fma(v[4], x_at<MODE, CG_LOADS>(a, lr + a.n, beta), t);
    if (MODE == ST_RESID) {
        const double rv = a.b[lr] - t;
        a.y[lr] = rv;
        a.y2[lr] = rv;
        return rv * rv;
    }
    a.y[lr] = t;
    if (st_fused(MODE)) {
        a.y2[lr] = xc;
        if (st_nx(MODE) > 0) a.xs[lr] = retire_x<MODE>(a, xa, lr, a.xs[lr], po);
    }
    return (MODE == ST_PLAIN) ? 0.0 : xc * t;
}

template <int MODE, bool CG_LOADS>
__global__ void __launch_bounds__(256) stencil5_sweep_kernel(const Stencil5Args a) {
    __shared__ double warp_part[8];
    griddep_wait();
    // the scalars of this launch are loaded together, in front of the branch on the first of them: one
    // global-memory latency per CTA instead of two (a CTA lives for a few microseconds only)
    const int conv = (a.converged != nullptr) ? *a.converged : 0;
    const double beta = st_fused(MODE) ? a.ab[1] : 0.0;
    XAlphas<MODE> xa;
    xa.load(a);
    if (conv != 0) return;
    const unsigned int nloc = (unsigned int)a.n_local, n = (unsigned int)a.n;
    constexpr unsigned int ROWS_PER_CTA = 256u * SWEEP_TILES;
    // logical CTA: the CTAs that cover the first grid row of a band with a halo run last
    unsigned int cta = blockIdx.x;
    const bool halos = (a.halo_prev != nullptr || a.halo_next != nullptr);
    if (halos && gridDim.x > 2) {
        const unsigned int head = min((n + ROWS_PER_CTA - 1) / ROWS_PER_CTA, gridDim.x - 1);  // CTAs touching halo_prev
        cta = (cta + head) % gridDim.x;
    }
    const unsigned int row_lo = cta * ROWS_PER_CTA;
    if (a.flag_prev != nullptr || a.flag_next != nullptr) {
        const unsigned int row_hi = min(row_lo + ROWS_PER_CTA, nloc);
        const bool need_prev = a.flag_prev != nullptr && row_lo < n;
        const bool need_next = a.flag_next != nullptr && row_hi + n > nloc;
        if (need_prev || need_next) {
            if (threadIdx.x == 0) {
                const uint32_t want = (a.epoch_ptr != nullptr) ? __ldcg(a.epoch_ptr) : a.epoch;
                if (need_prev) wait_flag(a.flag_prev, want, a.error_word);
                if (need_next) wait_flag(a.flag_next, want, a.error_word);
            }
            __syncthreads();
        }
    }
    const unsigned int i_off = (unsigned int)(a.row_offset / a.n), j_off = (unsigned int)(a.row_offset - (long long)i_off * a.n);
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < SWEEP_TILES; k++) {
        const unsigned int lr = row_lo + k * 256u + threadIdx.x;
#if B200_SWEEP_SHFL
        // centre element of every valid row first (all lanes take part in the shuffles: no early exit)
        const bool valid = lr < nloc;
        double poC = 0.0, xC = 0.0;
        if (valid) {
            if (st_fused(MODE)) {
                poC = __ldg(a.x + lr);  // p_old
                xC = fma(beta, poC, __ldg(a.r + lr));
            } else {
                xC = __ldg(a.x + lr);
            }
        }
        double xW = __shfl_up_sync(0xffffffffu, xC, 1), xE = __shfl_down_sync(0xffffffffu, xC, 1);
        if (!valid) continue;
#else
        if (lr >= nloc) break;
#endif
        const unsigned int q = (j_off + lr) / n;
        const unsigned int i = i_off + q, j = j_off + lr - q * n;
        const bool interior = (i >= 1) && (i + 2 <= n) && (j >= 1) && (j + 2 <= n);
        if (interior && lr >= n && lr + n < nloc) {
            // ---------------------------------------------------------------- every neighbour is a local element
            const double* v = a.values + (a.base0 + (long long)i * a.row_stride + 5 * (long long)j);
            const double vN = sweep_coef(v), vW = sweep_coef(v + 1), vC = sweep_coef(v + 2), vE = sweep_coef(v + 3), vS = sweep_coef(v + 4);
#if B200_SWEEP_SHFL
            double xN, xS;
            const unsigned int lane = threadIdx.x & 31u;
            if (st_fused(MODE)) {
                const double* pr = a.r + lr;
                const double* pp = a.x + lr;
                if (lane == 0) xW = fma(beta, __ldg(pp - 1), __ldg(pr - 1));
                if (lane == 31) xE = fma(beta, __ldg(pp + 1), __ldg(pr + 1));
                xN = fma(beta, __ldg(pp - (long long)n), __ldg(pr - (long long)n));
                xS = fma(beta, __ldg(pp + n), __ldg(pr + n));
            } else {
                const double* px = a.x + lr;
                if (lane == 0) xW = __ldg(px - 1);
                if (lane == 31) xE = __ldg(px + 1);
                xN = __ldg(px - (long long)n); xS = __ldg(px + n);
            }
#else
            double xW, xC, xE, xN, xS, poC = 0.0;
            if (st_fused(MODE)) {
                const double* pr = a.r + lr;
                const double* pp = a.x + lr;  // p_old
                poC = __ldg(pp);
                xC = fma(beta, poC, __ldg(pr));
                xW = fma(beta, __ldg(pp - 1), __ldg(pr - 1));
                xE = fma(beta, __ldg(pp + 1), __ldg(pr + 1));
                xN = fma(beta, __ldg(pp - (long long)n), __ldg(pr - (long long)n));
                xS = fma(beta, __ldg(pp + n), __ldg(pr + n));
            } else {
                const double* px = a.x + lr;
                xC = __ldg(px); xW = __ldg(px - 1); xE = __ldg(px + 1);
                xN = __ldg(px - (long long)n); xS = __ldg(px + n);
            }
#endif
            double t = vC * xC;
            t = fma(vW, xW, t);
            t = fma(vE, xE, t);
            t = fma(vN, xN, t);
            t = fma(vS, xS, t);
            if (MODE == ST_RESID) {
                const double rv = __ldg(a.b + lr) - t;
                sweep_store(a.y + lr, rv);
                sweep_store(a.y2 + lr, rv);
                acc = fma(rv, rv, acc);
            } else {
                sweep_store(a.y + lr, t);
                if (st_fused(MODE)) {
                    sweep_store(a.y2 + lr, xC);
                    if (st_nx(MODE) > 0) sweep_store(a.xs + lr, retire_x<MODE>(a, xa, lr, a.xs[lr], poC));
                }
                if (MODE != ST_PLAIN) acc = fma(xC, t, acc);
            }
        } else if (interior) {
            acc += sweep_row_generic<MODE, CG_LOADS>(a, lr, xa, beta);
        } else {
            acc += boundary_row<MODE, CG_LOADS>(a, a.row_offset + lr, xa, beta);
        }
    }
    griddep_launch();
    if (MODE != ST_PLAIN) {
        acc = warp_sum(acc);
        if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < 8; w++) t += warp_part[w];
            a.partials[blockIdx.x] = t;
        }
    }
}

}  // namespace b200
