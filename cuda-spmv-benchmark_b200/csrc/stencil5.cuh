// stencil5.cuh -- STENCIL5 SpMV for sm_100a (interior fast path + CSR boundary pass).
//
// What it replaces in the reference: stencil5_csr_direct_kernel
// (src/spmv/spmv_stencil_csr_direct.cu:76-123) and stencil5_csr_partitioned_halo_kernel
// (src/spmv/spmv_stencil_partitioned_halo_kernel.cu:17-98) -- one thread per row, scalar loads.
//
// B200 design (HBM-bound: 40 B values + 8 B x + 8 B y per row, 5 DFMA):
//   * The interior of the grid is tiled into items of ROWS grid rows x (32*COLS) columns; one
//     WARP owns one item and marches down its rows.
//   * The 5-doubles-per-row `values` stream of an item row is ONE contiguous span of the CSR
//     values array (closed-form offset, no row_ptr / col_idx reads).  Lane 0 moves it with a
//     1-D bulk async copy (cp.async.bulk, TMA engine, SASS UBLKCP) into a warp-private
//     STAGES-deep shared-memory ring guarded by mbarriers: no registers are tied up, several KB
//     are in flight per warp, HBM sees full-line bursts.  Lanes then read their 5 coefficients
//     with conflict-free LDS.64 (stride 40 B).
//   * x is register-rotated down the column (north <- centre <- south): each x element is loaded
//     once per item (coalesced 256 B per warp), west/east come from neighbouring lanes by shuffle.
//   * Boundary rows of the grid (4n-4 rows) walk the CSR arrays in extra CTAs at the end of the
//     same launch.
//   * Optional fusion (template MODE): p.Ap partial sums (CG), or r = b - A x, p = r, r.r, or the
//     whole direction update (ST_FUSED): p = r + beta p_old is formed in registers while the rows
//     are loaded, written once, x += alpha p_old is retired on the way, then Ap = A p and p.Ap.
//   * Band mode (multi-GPU): x is addressed as local / halo_prev / halo_next exactly like the
//     reference halo kernel; items that touch a halo spin on the neighbour's arrival flag, and are
//     scheduled last so the halo transfer overlaps the interior work.
// Arithmetic order is the reference's (read from its PTX): t = vC*xC; fma(vW,xW,t); fma(vE,xE,t);
// fma(vN,xN,t); fma(vS,xS,t); boundary rows: fma chain over k from 0.0.
#pragma once
#include "cg_kernels.cuh"
#include "common.cuh"

namespace b200 {

// ST_FUSED retires ONE pending x update per launch (x += alpha_{k-1} p_{k-1}).  The x stream (8 B read + 8 B
// written per row) can be amortised: ST_FUSED_X0 leaves x alone, ST_FUSED_X<m> retires the m pending updates
// of the last m iterations in one read-modify-write of x, oldest first (same fma chain as m separate
// updates: bit-identical), reading m-1 older directions next to the p_old it needs anyway.
enum { ST_PLAIN = 0, ST_DOT = 1, ST_RESID = 2, ST_FUSED = 3, ST_FUSED_X0 = 4, ST_FUSED_X2 = 5, ST_FUSED_X3 = 6, ST_FUSED_X4 = 7 };
__host__ __device__ constexpr bool st_fused(int mode) { return mode >= ST_FUSED; }
__host__ __device__ constexpr int st_nx(int mode) { return mode == ST_FUSED ? 1 : (mode >= ST_FUSED_X2 ? mode - 3 : 0); }
constexpr int ST_MAX_NX = 4;

struct Stencil5Args {
    const int* row_ptr;    // local, rebased to 0 (boundary rows only)
    const int* col_idx;    // GLOBAL column ids (boundary rows only)
    const double* values;  // local slice of the CSR / ELLPACK values
    long long values_len;  // number of readable doubles behind `values`
    const double* x;       // local part: x[0] is global row `row_offset`
    const double* halo_prev;  // global rows [row_offset-n, row_offset) or NULL
    const double* halo_next;  // global rows [row_offset+n_local, +n) or NULL
    double* y;                // PLAIN/DOT: A x ; RESID: r = b - A x
    double* y2;               // RESID: second copy (p = r)
    const double* b;          // RESID
    double* partials;         // DOT/RESID/FUSED: one double per CTA (gridDim.x entries)
    long long base0;          // interior element (i,j) lives at base0 + i*row_stride + 5*j
    long long row_stride;
    long long row_offset;
    long long n_local;
    int n;        // grid side
    int i_first;  // first / last interior grid row intersecting the band
    int i_last;
    int rows_per_item;
    int n_strips;
    int n_chunks;
    int ctas_per_chunk;
    int n_interior_ctas;
    int n_boundary_rows;  // 4n-4 (1 for n == 1)
    const uint32_t* flag_prev;  // halo arrival flags (this GPU's memory), NULL = no wait
    const uint32_t* flag_next;
    uint32_t epoch;
    const int* converged;  // optional: kernel is a no-op once *converged != 0
    int* error_word;       // optional: set to 1 on a flag-wait timeout
    // ST_FUSED: x = p_old (local part; the halos hold the NEW p), y = Ap, y2 = p_new
    const double* r;       // residual (local)
    double* xs;            // solution vector: xs += alpha * p_old
    const double* ab;      // device scalars: ab[0] = alpha (of the previous iteration), ab[1] = beta
    const uint32_t* epoch_ptr;  // if set: the halo sequence number to wait for is read from here
    // ST_FUSED_X<m>: xp[k] = direction p_{it-2-k} (k < m-1), alpha_hist[j & 7] = alpha of iteration j,
    // *iter_ptr = it (iterations completed so far)
    const double* xp[ST_MAX_NX - 1];
    const double* alpha_hist;
    const int* iter_ptr;
};

// alphas of the pending x updates: al[k] belongs to p_{it-1-k}
template <int MODE>
struct XAlphas {
    double al[st_nx(MODE) > 0 ? st_nx(MODE) : 1];
    __device__ __forceinline__ void load(const Stencil5Args& a) {
        constexpr int NX = st_nx(MODE);
        if (NX >= 1) al[0] = a.ab[0];
        if (NX >= 2) {
            const int it = *a.iter_ptr;
#pragma unroll
            for (int k = 1; k < NX; k++) al[k] = a.alpha_hist[(it - 1 - k) & 7];
        }
    }
};
// x after the pending updates, oldest first; xv = x[lr], pold = p_{it-1}[lr]
template <int MODE>
__device__ __forceinline__ double retire_x(const Stencil5Args& a, const XAlphas<MODE>& xa, long long lr, double xv, double pold) {
    constexpr int NX = st_nx(MODE);
#pragma unroll
    for (int k = NX - 1; k >= 1; k--) xv = fma(xa.al[k], __ldg(a.xp[k - 1] + lr), xv);
    return fma(xa.al[0], pold, xv);
}

__device__ __forceinline__ void wait_flag(const uint32_t* flag, uint32_t epoch, int* error_word) {
    const uint64_t t0 = globaltimer_ns();
    while ((int32_t)(ld_acquire_sys(flag) - epoch) < 0) {
        // 8 s (or an earlier timeout anywhere in this solve): report instead of hanging the GPU
        if ((error_word && *(volatile int*)error_word) || globaltimer_ns() - t0 > 8000000000ull) {
            if (error_word) *error_word = 1;
            break;
        }
        __nanosleep(64);
    }
}

// Vector element `idx` (local index; outside [0, n_local) = halo).  ST_FUSED: the local part is
// formed on the fly, p = fma(beta, p_old, r) (reference update_p_kernel, cg_solver.cu:91-96), the
// halos already hold the new p.  `po` receives p_old (0 outside the band).
template <int MODE, bool CG_LOADS>
__device__ __forceinline__ double x_at(const Stencil5Args& a, long long idx, double beta = 0.0, double* po = nullptr) {
    const double* p;
    if (st_fused(MODE) && po) *po = 0.0;
    if (idx >= 0 && idx < a.n_local) {
        if (st_fused(MODE)) {
            const double pold = __ldg(a.x + idx);
            if (po) *po = pold;
            return fma(beta, pold, __ldg(a.r + idx));
        }
        p = a.x + idx;
    } else if (idx < 0) {
        if (a.halo_prev == nullptr || idx < -(long long)a.n) return 0.0;
        p = a.halo_prev + (idx + a.n);
    } else {
        if (a.halo_next == nullptr || idx >= a.n_local + a.n) return 0.0;
        p = a.halo_next + (idx - a.n_local);
    }
    return CG_LOADS ? __ldcg(p) : __ldg(p);
}

// One boundary row: CSR walk, reference order (spmv_stencil_csr_direct.cu:113-119).
template <int MODE, bool CG_LOADS>
__device__ __forceinline__ double boundary_row(const Stencil5Args& a, long long r, const XAlphas<MODE>& xa, double beta) {
    const long long lr = r - a.row_offset;
    long long s, e;
    if (a.row_ptr != nullptr) { s = a.row_ptr[lr]; e = a.row_ptr[lr + 1]; }
    else { s = lr * 5; e = s + 5; }  // ELLPACK width 5, padding index -1
    double sum = 0.0, xc = 0.0;
    for (long long k = s; k < e; k++) {
        // CSR column ids are read as unsigned: grids beyond 46340^2 rows (multi-GPU weak scaling) store
        // global columns modulo 2^32; only the ELLPACK layout has padding (-1)
        const long long c = a.row_ptr != nullptr ? (long long)(unsigned int)a.col_idx[k] : (long long)a.col_idx[k];
        if (c < 0) continue;
        const double xv = x_at<MODE, CG_LOADS>(a, c - a.row_offset, beta);
        if (c == r) xc = xv;
        sum = fma(a.values[k], xv, sum);
    }
    if (st_fused(MODE)) {  // this thread owns row r: retire x, publish the new p
        const double pold = a.x[lr];
        xc = fma(beta, pold, a.r[lr]);
        a.y2[lr] = xc;
        if (st_nx(MODE) > 0) a.xs[lr] = retire_x<MODE>(a, xa, lr, a.xs[lr], pold);
        a.y[lr] = sum;
        return xc * sum;
    }
    if (MODE == ST_RESID) {
        const double rv = a.b[lr] - sum;
        a.y[lr] = rv;
        a.y2[lr] = rv;
        return rv * rv;
    }
    a.y[lr] = sum;
    return (MODE == ST_DOT) ? xc * sum : 0.0;
}

// registers: the fused mode sits at exactly 128 per thread with 4 columns per lane (4 CTAs of 4 warps per
// SM); anything that adds live state to the row loop (a ticket per CTA, a second halo stream) drops it to
// 3 CTAs per SM and costs ~10 % -- reductions and halo copies of the direction live in their own launches
template <int MODE, int COLS, int WARPS, int STAGES, bool CG_LOADS>
__global__ void __launch_bounds__(WARPS * 32) stencil5_kernel(const Stencil5Args a) {
    constexpr int W = 32 * COLS;                 // strip width in columns
    constexpr int STAGE_DOUBLES = 5 * W + 2;     // +2: 16-byte alignment slack at both ends
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double warp_part[WARPS];

    if (MODE != ST_PLAIN) griddep_wait();  // programmatic dependent launch: nothing of the previous kernel is read above
    if (a.converged != nullptr && *a.converged != 0) return;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double acc = 0.0;
    const double beta = st_fused(MODE) ? a.ab[1] : 0.0;
    XAlphas<MODE> xa;
    xa.load(a);
    // the halo sequence number to wait for is only read where a wait happens (no live register elsewhere)
    auto wanted = [&]() -> uint32_t { return (a.epoch_ptr != nullptr) ? __ldcg(a.epoch_ptr) : a.epoch; };

    if ((int)blockIdx.x >= a.n_interior_ctas) {
        // ------------------------------------------------------------ boundary pass (CSR walk)
        if (a.flag_prev != nullptr || a.flag_next != nullptr) {
            if (threadIdx.x == 0) {
                if (a.flag_prev) wait_flag(a.flag_prev, wanted(), a.error_word);
                if (a.flag_next) wait_flag(a.flag_next, wanted(), a.error_word);
            }
            __syncthreads();
        }
        const int n = a.n;
        const long long t = (long long)(blockIdx.x - a.n_interior_ctas) * blockDim.x + threadIdx.x;
        if (t < a.n_boundary_rows) {
            long long r;
            if (t < n) r = t;                                              // grid row 0
            else if (t < 2LL * n) r = (long long)(n - 1) * n + (t - n);    // grid row n-1
            else if (t < 2LL * n + (n - 2)) r = (t - 2LL * n + 1) * n;     // column 0
            else r = (t - 2LL * n - (n - 2) + 1) * n + (n - 1);            // column n-1
            if (r >= a.row_offset && r < a.row_offset + a.n_local) acc = boundary_row<MODE, CG_LOADS>(a, r, xa, beta);
        }
    } else {
        // ------------------------------------------------------------ interior fast path
        double* ring = reinterpret_cast<double*>(smem_raw) + (size_t)warp * STAGES * STAGE_DOUBLES;
        uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)WARPS * STAGES * STAGE_DOUBLES * 8) +
                         warp * STAGES;
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < STAGES; s++) mbar_init(&bars[s], 1);
            mbar_fence_init();
        }
        __syncwarp();

        const int n = a.n;
        int chunk = blockIdx.x / a.ctas_per_chunk;
        const int strip = (blockIdx.x % a.ctas_per_chunk) * WARPS + warp;
        const bool halo_mode = (a.flag_prev != nullptr || a.flag_next != nullptr);
        if ((halo_mode || a.halo_prev != nullptr || a.halo_next != nullptr) && a.n_chunks >= 3) {
            // chunks that read a halo (first / last of the band) run last: transfer overlaps the rest
            // (keyed on the halo pointers as well, so that the CTA -> item map, and with it the order
            // of the partial sums, does not depend on whether the launch has to wait for flags)
            chunk = (chunk < a.n_chunks - 2) ? chunk + 1 : (chunk == a.n_chunks - 2 ? 0 : a.n_chunks - 1);
        }
        const int i0 = a.i_first + chunk * a.rows_per_item;
        const int i1 = min(i0 + a.rows_per_item, a.i_last + 1);
        const int j0 = 1 + strip * W;
        const int count = min(W, (n - 1) - j0);  // interior columns j0 .. j0+count-1 (<= n-2)

        if (strip < a.n_strips && i0 < i1 && count > 0) {
            const long long off = a.row_offset, nl = a.n_local;
            if (halo_mode) {
                // does this item read x outside the local band?
                const bool need_prev = ((long long)(i0 - 1) * n + j0 - 1 < off);
                const bool need_next = ((long long)i1 * n + j0 + count + 1 > off + nl);
                if (lane == 0) {
                    if (need_prev && a.flag_prev) wait_flag(a.flag_prev, wanted(), a.error_word);
                    if (need_next && a.flag_next) wait_flag(a.flag_next, wanted(), a.error_word);
                }
                __syncwarp();
            }
            const uint64_t policy = l2_policy_evict_first();

            // --- values ring producer (lane 0); returns nothing, geometry recomputed by consumers
            auto row_span = [&](int i, long long& e_lo, long long& e_hi) {
                // in-band interior columns of grid row i inside this strip -> element range
                long long jlo = j0, jhi = j0 + count;
                const long long rb = (long long)i * n;
                if (off - rb > jlo) jlo = off - rb;
                if (off + nl - rb < jhi) jhi = off + nl - rb;
                if (jlo >= jhi) { e_lo = 0; e_hi = 0; return; }
                const long long base = a.base0 + (long long)i * a.row_stride;
                e_lo = base + 5 * jlo;
                e_hi = base + 5 * jhi;
            };
            auto issue = [&](int i) {
                long long e_lo, e_hi;
                row_span(i, e_lo, e_hi);
                if (e_hi <= e_lo) return;
                const int slot = (i - i0) % STAGES;
                double* dst = ring + slot * STAGE_DOUBLES;
                const long long a_al = e_lo & ~1LL;
                long long b_al = (e_hi + 1) & ~1LL;
                if (b_al > a.values_len) b_al -= 2;
                if (b_al > a_al) {
                    if (lane == 0) {
                        const uint32_t bytes = (uint32_t)(b_al - a_al) * 8u;
                        // order the ring slot's earlier generic-proxy reads before the async-proxy write
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        mbar_arrive_expect_tx(&bars[slot], bytes);
                        bulk_g2s(dst, a.values + a_al, bytes, &bars[slot], policy);
                    }
                }
                // tail the 16-byte granularity could not cover (only at the very end of a slice)
                const long long m_lo = (b_al > e_lo) ? b_al : e_lo;
                for (long long e = m_lo + lane; e < e_hi; e += 32) dst[e - a_al] = a.values[e];
            };

            // --- prologue
#pragma unroll
            for (int s = 0; s < STAGES - 1; s++)
                if (i0 + s < i1) issue(i0 + s);

            // x is register-rotated down the column with a prefetch distance of two grid rows: the
            // load for row i+2 is issued a full iteration before its first use, so the row loop never
            // stalls on global-memory latency (only on the values ring).
            double xN[COLS], xC[COLS], xS[COLS], xF[COLS] = {};
            double eC = 0.0, eS = 0.0, eF = 0.0;  // lane 0: x(i, j0-1)   lane 31: x(i, j0+W)
            // ST_FUSED: xF / eF hold the raw r of the row in flight (or the halo's new p), poF / peF its
            // p_old; the row becomes p = fma(beta, p_old, r) one turn later, when the loads have
            // landed.  poS / poC carry p_old down to the turn that retires x += alpha p_old.
            double poF[COLS] = {}, poS[COLS] = {}, poC[COLS] = {};
            double peF = 0.0;
            const long long col_base = (long long)j0 + lane - off;  // + i*n + 32c -> local index
            auto raw_at = [&](long long idx, double& po) -> double {  // (r or halo p, p_old): no arithmetic
                po = 0.0;
                if (idx >= 0 && idx < nl) {
                    po = __ldg(a.x + idx);
                    return __ldg(a.r + idx);
                }
                return x_at<ST_PLAIN, CG_LOADS>(a, idx);  // the halos already hold the NEW p
            };
            auto load_row = [&](int ii, double (&xr)[COLS], double& er, double (&po)[COLS], double& pe) {
                const long long rb = (long long)ii * n;
#pragma unroll
                for (int c = 0; c < COLS; c++) {
                    const int j = j0 + lane + 32 * c;
                    po[c] = 0.0;
                    if (st_fused(MODE)) xr[c] = (j <= n - 1) ? raw_at(rb + col_base + 32 * c, po[c]) : 0.0;
                    else xr[c] = (j <= n - 1) ? x_at<MODE, CG_LOADS>(a, rb + col_base + 32 * c) : 0.0;
                }
                er = 0.0;
                pe = 0.0;
                if (st_fused(MODE)) {
                    if (lane == 0) er = raw_at(rb + j0 - 1 - off, pe);
                    if (lane == 31 && j0 + W <= n - 1) er = raw_at(rb + j0 + W - off, pe);
                } else {
                    if (lane == 0) er = x_at<MODE, CG_LOADS>(a, rb + j0 - 1 - off);
                    if (lane == 31 && j0 + W <= n - 1) er = x_at<MODE, CG_LOADS>(a, rb + j0 + W - off);
                }
            };
            auto finish_row = [&](double (&xr)[COLS], double& er, const double (&po)[COLS], double pe) {
                if (st_fused(MODE)) {  // raw -> p = fma(beta, p_old, r); halo / padding: fma(beta, 0, v) = v
#pragma unroll
                    for (int c = 0; c < COLS; c++) xr[c] = fma(beta, po[c], xr[c]);
                    er = fma(beta, pe, er);
                }
            };
            {
                double dummy, dpe, dpo[COLS];
                load_row(i0 - 1, xN, dummy, dpo, dpe);
                finish_row(xN, dummy, dpo, dpe);
            }
            {
                double dpe;
                load_row(i0, xC, eC, poC, dpe);
                finish_row(xC, eC, poC, dpe);
            }
            load_row(i0 + 1, xS, eS, poS, peF);
            if (!st_fused(MODE)) {
                // nothing: xS is final
            } else {
                // keep row i0+1 raw in the F registers: it is finished at the top of the first turn
#pragma unroll
                for (int c = 0; c < COLS; c++) { xF[c] = xS[c]; poF[c] = poS[c]; }
                eF = eS;
            }

            uint32_t phase_bits = 0;
            for (int i = i0; i < i1; i++) {
                // refill the slot freed by the previous row (all lanes are past its LDS: syncwarp)
                __syncwarp();
                if (i + STAGES - 1 < i1) issue(i + STAGES - 1);

                double xsv[COLS];
                if (st_fused(MODE)) {
                    // row i+1 was loaded one turn ago: finish it, then reuse the F registers for row i+2
#pragma unroll
                    for (int c = 0; c < COLS; c++) { xS[c] = xF[c]; poS[c] = poF[c]; }
                    eS = eF;
                    finish_row(xS, eS, poS, peF);
                    if (i + 2 <= i1) load_row(i + 2, xF, eF, poF, peF);
#pragma unroll
                    for (int c = 0; c < COLS; c++) {
                        const long long r = (long long)i * n + j0 + lane + 32 * c;
                        if (st_nx(MODE) > 0)
                            xsv[c] = (j0 + lane + 32 * c <= n - 2 && r >= off && r < off + nl) ? a.xs[r - off] : 0.0;
                    }
                } else {
                    double dpo[COLS], dpe;
                    if (i + 2 <= i1) load_row(i + 2, xF, eF, dpo, dpe);
                }
                double bv[COLS];
                if (MODE == ST_RESID) {
#pragma unroll
                    for (int c = 0; c < COLS; c++) {
                        const long long r = (long long)i * n + j0 + lane + 32 * c;
                        bv[c] = (j0 + lane + 32 * c <= n - 2 && r >= off && r < off + nl) ? a.b[r - off] : 0.0;
                    }
                }

                long long e_lo, e_hi;
                row_span(i, e_lo, e_hi);
                if (e_hi > e_lo) {
                    const int slot = (i - i0) % STAGES;
                    const long long a_al = e_lo & ~1LL;
                    long long b_al = (e_hi + 1) & ~1LL;
                    if (b_al > a.values_len) b_al -= 2;
                    if (b_al > a_al) {
                        mbar_wait(&bars[slot], (phase_bits >> slot) & 1u);
                        phase_bits ^= (1u << slot);
                    }
                    __syncwarp();  // manual tail stores (if any) visible to all lanes
                    const double* v = ring + slot * STAGE_DOUBLES;
                    const long long ebase = a.base0 + (long long)i * a.row_stride - a_al;  // + 5*j
                    const long long gr = (long long)i * n;
#pragma unroll
                    for (int c = 0; c < COLS; c++) {
                        const int j = j0 + lane + 32 * c;
                        // west / east neighbours: adjacent lane, wrapping into the next 32-col block
                        double xw = __shfl_up_sync(B200_FULL, xC[c], 1);
                        double xe = __shfl_down_sync(B200_FULL, xC[c], 1);
                        if (c > 0) {
                            const double wrap = __shfl_sync(B200_FULL, xC[c - 1], 31);
                            if (lane == 0) xw = wrap;
                        } else if (lane == 0) {
                            xw = eC;
                        }
                        if (c < COLS - 1) {
                            const double wrap = __shfl_sync(B200_FULL, xC[c + 1], 0);
                            if (lane == 31) xe = wrap;
                        } else if (lane == 31) {
                            xe = eC;
                        }
                        const long long r = gr + j;
                        if (j <= n - 2 && r >= off && r < off + nl) {
                            const double* vv = v + (ebase + 5LL * j);
                            double t = vv[2] * xC[c];
                            t = fma(vv[1], xw, t);
                            t = fma(vv[3], xe, t);
                            t = fma(vv[0], xN[c], t);
                            t = fma(vv[4], xS[c], t);
                            const long long lr = r - off;
                            if (MODE == ST_RESID) {
                                const double rv = bv[c] - t;
                                a.y[lr] = rv;
                                a.y2[lr] = rv;
                                acc = fma(rv, rv, acc);
                            } else {
                                a.y[lr] = t;
                                if (MODE == ST_DOT || st_fused(MODE)) acc = fma(xC[c], t, acc);
                                if (st_fused(MODE)) {
                                    a.y2[lr] = xC[c];  // the new p, written once
                                    // x += alpha p_old (+ the older pending updates), deferred
                                    if (st_nx(MODE) > 0) a.xs[lr] = retire_x<MODE>(a, xa, lr, xsv[c], poC[c]);
                                }
                            }
                        }
                    }
                }
                if (st_fused(MODE)) {
#pragma unroll
                    for (int c = 0; c < COLS; c++) { xN[c] = xC[c]; xC[c] = xS[c]; poC[c] = poS[c]; }
                    eC = eS;
                } else {
#pragma unroll
                    for (int c = 0; c < COLS; c++) { xN[c] = xC[c]; xC[c] = xS[c]; xS[c] = xF[c]; }
                    eC = eS;
                    eS = eF;
                }
            }
        }
    }

    if (MODE != ST_PLAIN) griddep_launch();
    if (MODE != ST_PLAIN) {
        acc = warp_sum(acc);
        if (lane == 0) warp_part[warp] = acc;
        __syncthreads();
        // One partial per CTA, fixed owner.  The grid has O(1e5) short-lived CTAs: a ticket per CTA (fence +
        // atomic round trip, ~1 us of a ~30 us CTA) cost 3 % of the kernel, so the fixed-order final sum runs
        // in cg_reduce_kernel instead, launched programmatically behind this grid (no launch gap).
        if (threadIdx.x == 0) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < WARPS; w++) t += warp_part[w];
            a.partials[blockIdx.x] = t;
        }
    }
}

template <int COLS, int WARPS, int STAGES>
constexpr size_t stencil5_smem_bytes() {
    return (size_t)WARPS * STAGES * (5 * 32 * COLS + 2) * 8 + (size_t)WARPS * STAGES * 8;
}

}  // namespace b200
