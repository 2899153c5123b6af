// csr_ell.cuh -- generic CSR and ELLPACK SpMV for sm_100a (no cuSPARSE).
//
// Replaces cusparseSpMV (src/spmv/spmv_cusparse_csr.cu:246,281) behind the "cusparse-csr"
// operator name, and creates the ELLPACK operator the reference only declares
// (include/spmv_ellpack.h:28-51).  Semantics = the reference's scalar CSR kernel
// (src/solvers/cg_solver_mgpu_partitioned.cu:40-56): sum_k fma(v[k], x[col[k]], sum), k ascending.
//
// Two kernels live here:
//   * csr_ring_kernel (below, the production path): warp-private bulk-copy ring, lane-per-row with
//     x gathered one group ahead for short rows, warp-per-row streaming for long rows; the scheme
//     is picked per 32-row group, the long-row threshold comes from the plan's row-length histogram.
//   * csr_warp_stream_kernel ("warp-stream", register staged): the first-generation kernel, used
//     when col_idx / values are not 16-byte aligned and kept as variant 100 for A/B measurements.
//     A WARP owns 32 consecutive rows, sweeps their contiguous non-zero range in windows of 256
//     entries with coalesced loads, gathers x[col] for all of them at once, parks value and x side
//     by side in shared memory, then each lane adds the products of ITS row in k order (bit-exact);
//     groups with long rows go warp-per-row (butterfly sum, tolerance 1e-12).
// ELLPACK (row-major, padding index -1) is the same code with row_ptr[r] = r * width.
#pragma once
#include <type_traits>

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
    int vector_threshold;  // ring kernel: longest row of a 32-row group above which it goes warp-per-row
                           // (warp-stream kernel: mean entries per row of the group)
    double alpha, beta;
    // optional fusion for CG (ring kernel only): partials[item] = sum over the item's rows of x[row] * y[row]
    // (the p.Ap dot product of a square operator); `converged` (optional) turns the launch into a no-op
    double* dot_partials;
    long long dot_capacity;  // slots behind dot_partials (checked by the launcher)
    const int* converged;
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


// ================================================================================================
// csr_ring_kernel -- the production CSR / ELLPACK kernel ("warp ring").
//
// A WARP owns one item = `groups_per_warp` consecutive 32-row groups, i.e. one contiguous range
// [K0, K1) of col_idx / values.  Items are handed out in launch order, so the resident warps form
// a wavefront over the matrix and banded x accesses of neighbouring items hit in L2.
//   * Lane 0 streams the range in fixed windows of WIN entries into a warp-private circular
//     shared-memory buffer of STAGES windows with 1-D bulk async copies (TMA engine, SASS UBLKCP;
//     mbarrier complete_tx, L2 evict-first): independent of where rows begin, no registers tied
//     up, up to STAGES-1 windows in flight per warp.  Row extents are coalesced 4-byte cp.async
//     loads into a small shared-memory queue, issued three pipeline turns before they are needed.
//   * Short rows ("lane per row", groups with at most RING - 2 WIN entries and rows no longer than
//     `vector_threshold` <= 32): every lane owns one row of the group.  One group AHEAD of the
//     arithmetic it reads its first 8 column ids from shared memory and launches the x gathers
//     into registers, so the gather latency hides behind the previous group's work; then
//     sum = fma(v[k], x[col[k]], sum) for k ascending -- bit-identical to the scalar reference
//     (entries past the 8th of a row are gathered in place).
//   * Long rows ("warp per row"): lanes stride over the row, lane l takes the entries with
//     (k - row_start) % 32 == l in k order, butterfly sum at the row end (rounding-level
//     difference to the sequential order, documented tolerance 1e-12).  Rows may be longer than
//     the ring: they stream through it window by window.
// Only __syncwarp is used; warps are fully independent.  col_idx and values must be 16-byte aligned
// (the launcher routes anything else to csr_warp_stream_kernel above).  DOT = true additionally
// writes one partial of x.y per item (CG: p.Ap), see CsrArgs::dot_partials.
// ================================================================================================
constexpr int kCsrPrefetch = 8;  // x values per row gathered one group ahead
constexpr int kCsrExtentSlots = 4;  // row extents of 3 groups in flight (cp.async) + 1 being read
constexpr int kCsrMirror = 32;   // ring[0, 32) is mirrored behind the ring end: a lane-per-row row
                                 // (at most 32 entries) never wraps, its shared addresses are base + j

// lane-per-row step 1: column ids of the lane's first N entries -> x gathers in flight.  The loads are
// unconditional (slots past the row end read x[0]) so that they compile to straight-line code;
// FULL = every row of the group has exactly N entries (no per-lane predicates at all).
// ELLPACK padding (column -1, value 0.0 as written by the builders): x is replaced by 0, and
// fma(0.0, 0.0, sum) leaves the sum untouched.
template <int N, bool ELL, bool FULL>
__device__ __forceinline__ void lpr_gather(const int* __restrict__ pc, int len, const double* __restrict__ x,
                                           double (&xq)[kCsrPrefetch]) {
#pragma unroll
    for (int j = 0; j < N; j++) {
        int c = pc[j];  // always inside the ring (+ mirror); meaningful only for j < len
        if (!FULL) c = (j < len) ? c : 0;
        if (ELL) {
            const double xv = __ldg(x + (unsigned)max(c, 0));
            xq[j] = c >= 0 ? xv : 0.0;
        } else {
            xq[j] = __ldg(x + (unsigned)c);
        }
    }
}
// lane-per-row step 2: the k-ordered fma chain over the first N entries
template <int N, bool FULL>
__device__ __forceinline__ double lpr_chain(const double* __restrict__ pv, int len, const double (&xq)[kCsrPrefetch]) {
    double sum = 0.0;
#pragma unroll
    for (int j = 0; j < N; j++)
        if (FULL || j < len) sum = fma(pv[j], xq[j], sum);
    return sum;
}
#define B200_CSR_DISPATCH_N(n, CALL)                                                              \
    switch (n) {                                                                                  \
        case 0: break;                                                                            \
        case 1: { constexpr int N_ = 1; CALL; } break;                                            \
        case 2: { constexpr int N_ = 2; CALL; } break;                                            \
        case 3: { constexpr int N_ = 3; CALL; } break;                                            \
        case 4: { constexpr int N_ = 4; CALL; } break;                                            \
        case 5: { constexpr int N_ = 5; CALL; } break;                                            \
        case 6: { constexpr int N_ = 6; CALL; } break;                                            \
        case 7: { constexpr int N_ = 7; CALL; } break;                                            \
        default: { constexpr int N_ = 8; CALL; } break;                                           \
    }

template <int STAGES, int WIN>
constexpr bool csr_ring_ell_is_lpr(int width) {
    return 32 * width <= STAGES * WIN - 2 * WIN && width <= kCsrMirror;
}

template <int STAGES, int WIN>
__host__ __device__ constexpr size_t csr_ring_warp_bytes() {
    return (size_t)(STAGES * WIN + kCsrMirror) * 12  // values + column ring (+ wrap mirror)
           + (size_t)kCsrExtentSlots * 32 * 4        // row extents queue (cp.async landing slots)
           + (size_t)STAGES * 8;                     // mbarriers
}

__device__ __forceinline__ void mbar_expect_tx_a(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s_a(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}

struct CsrGroup {
    int s, len;            // per lane: local start of its row, row length
    int gs, nnz, maxlen;   // uniform: local start of the group, its entries, longest row
    bool lpr;              // uniform: lane-per-row (else warp-per-row)
    bool full;             // uniform: all 32 rows have exactly maxlen entries
};

// MODE: 0 = CSR (every group picks lane-per-row or warp-per-row), 1 = ELLPACK narrow enough for
// lane-per-row throughout, 2 = wide ELLPACK, warp-per-row throughout (the host picks with
// csr_ring_ell_is_lpr).
template <int WARPS, int STAGES, int WIN, int MODE, int MINB, bool DOT = false>
__global__ void __launch_bounds__(WARPS * 32, MINB) csr_ring_kernel(const CsrArgs a, const int groups_per_warp) {
    constexpr bool ELL = MODE != 0;
    constexpr int RING = STAGES * WIN, M = RING - 1, J = kCsrPrefetch;
    constexpr int LIMIT = RING - 2 * WIN;  // a lane-per-row group must fit the ring beside one window in flight
    static_assert((RING & M) == 0 && (STAGES & (STAGES - 1)) == 0, "ring sizes must be powers of two");
    static_assert(LIMIT >= 32, "ring too small");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(B200_FULL, threadIdx.x >> 5, 0);  // tells the compiler it is warp-uniform
    if (a.converged != nullptr && *a.converged != 0) return;

    const long long Ra = ((long long)blockIdx.x * WARPS + warp) * groups_per_warp * 32;
    if (Ra >= a.n_rows) return;
    const int nrows = (int)min((long long)groups_per_warp * 32, a.n_rows - Ra);
    const int ngroups = (nrows + 31) >> 5;

    unsigned char* wbase = smem_raw + (size_t)warp * csr_ring_warp_bytes<STAGES, WIN>();
    double* sval = reinterpret_cast<double*>(wbase);
    int* scol = reinterpret_cast<int*>(sval + RING + kCsrMirror);
    int* sext = scol + RING + kCsrMirror;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sext + kCsrExtentSlots * 32);
    const uint32_t sval_a = smem_u32(sval), scol_a = smem_u32(scol), bar_a = smem_u32(bars);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < STAGES; i++) mbar_init(bars + i, 1);
        mbar_fence_init();
    }
    __syncwarp();
    const uint64_t policy = l2_policy_evict_first();
    const double* __restrict__ xp = a.x;

    long long K0, K1, total;
    if (ELL) {
        K0 = Ra * a.ell_width;
        K1 = (Ra + nrows) * a.ell_width;
        total = a.n_rows * a.ell_width;
    } else {
        long long t = 0;
        if (lane == 0) t = a.row_ptr[Ra];
        if (lane == 1) t = a.row_ptr[Ra + nrows];
        if (lane == 2) t = a.row_ptr[a.n_rows];
        K0 = __shfl_sync(B200_FULL, t, 0);
        K1 = __shfl_sync(B200_FULL, t, 1);
        total = __shfl_sync(B200_FULL, t, 2);
    }
    // item-local entry index = absolute index - wk0 (wk0 16-byte aligned for both arrays)
    const long long wk0 = K0 & ~3LL;
    const int k0l = (int)(K0 - wk0), k1l = (int)(K1 - wk0);
    const long long t4 = (total & ~3LL) - wk0;
    const int tl4 = t4 > 0x7fffff00LL ? 0x7fffff00 : (int)t4;   // bulk-copyable entries
    const int nwin = (k1l > k0l) ? (k1l + WIN - 1) / WIN : 0;
    const int lim4 = min(tl4, (k1l + 3) & ~3);  // bulk copies stop at the item's end (16-byte granule)
    const int* colp = a.col_idx + wk0;
    const double* valp = a.values + wk0;
    const int wk0i = (int)wk0;  // CSR: absolute entry indices are ints

    // ---------------------------------------------------------------- window ring
    int issued = 0, landed = 0, landed_end = 0, released_end = WIN;  // landed_end = landed * WIN
    auto issue_window = [&](int j) {
        const int cnt = min(WIN, lim4 - j * WIN);
        if (lane == 0 && cnt > 0) {
            // WAR on the slot: every lane's shared loads from it were consumed before the __syncwarp
            // in release(); reads need no proxy fence against the async-proxy write that follows
            const int slot = j & (STAGES - 1);
            const uint32_t bar = bar_a + slot * 8;
            mbar_expect_tx_a(bar, cnt * 12);
            bulk_g2s_a(sval_a + slot * WIN * 8, valp + (size_t)j * WIN, cnt * 8, bar, policy);
            bulk_g2s_a(scol_a + slot * WIN * 4, colp + (size_t)j * WIN, cnt * 4, bar, policy);
        }
    };
    // make entries [0, kend) available in the ring
    auto ensure = [&](int kend) {
        while (landed_end < kend) {
            const int j = landed;
            if (lim4 > j * WIN) mbar_wait_a(bar_a + (j & (STAGES - 1)) * 8, ((unsigned)j / STAGES) & 1u);
            if ((j + 1) * WIN > lim4 && k1l > lim4) {  // array tail outside the 16-byte granules (matrix end only)
                const int lo = max(lim4, j * WIN), hi = min(k1l, (j + 1) * WIN);
                for (int k = lo + lane; k < hi; k += 32) {
                    sval[k & M] = valp[k];
                    scol[k & M] = colp[k];
                }
                __syncwarp();
            }
            if ((j & (STAGES - 1)) == 0) {  // refresh the wrap mirror
                sval[RING + lane] = sval[lane];
                scol[RING + lane] = scol[lane];
                __syncwarp();
            }
            landed++;
            landed_end += WIN;
        }
    };
    // entries below k are dead: recycle their windows
    auto release = [&](int k) {
        while (released_end <= k) {
            __syncwarp();
            released_end += WIN;
            if (issued < nwin) issue_window(issued++);
        }
    };
#pragma unroll
    for (int i = 0; i < STAGES; i++)
        if (issued < nwin) issue_window(issued++);

    // ---------------------------------------------------------------- groups
    // Row extents: lane l of group gi needs row_ptr[gi*32 + l + 1] (its row end); the row start is the
    // neighbour lane's end.  The load is issued one pipeline turn before the values are looked at.
    const int* __restrict__ rowp = ELL ? nullptr : a.row_ptr + Ra;
    // The loads are 4-byte cp.async straight into a small shared-memory queue (one slot of 32 ints per
    // group, 3 groups in flight): no register waits on them until the group is actually built.
    int carry = k0l;  // local end of the previous group = start of the next one
    const uint32_t sext_a = smem_u32(sext) + lane * 4;
    auto fetch_extents = [&](int gi) {
        if (!ELL) {
            // (a 4-byte cp.async does not take an L2 cache hint: illegal instruction on sm_100a)
            const int* src = rowp + min(gi * 32 + lane + 1, nrows);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sext_a + (gi & (kCsrExtentSlots - 1)) * 128),
                         "l"(src)
                         : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
    };
    auto make_group = [&](int gi, CsrGroup& G) {  // gi >= ngroups: empty sentinel group
        int e;
        if (ELL) e = k0l + min(gi * 32 + lane + 1, nrows) * a.ell_width;
        else {
            asm volatile("cp.async.wait_group 2;" ::: "memory");  // all but the two youngest fetches landed
            e = sext[(gi & (kCsrExtentSlots - 1)) * 32 + lane] - wk0i;
        }
        const int up = __shfl_up_sync(B200_FULL, e, 1);
        G.s = lane == 0 ? carry : up;
        G.len = e - G.s;
        G.gs = carry;
        carry = __shfl_sync(B200_FULL, e, 31);
        G.nnz = carry - G.gs;
        if (ELL) {  // uniform rows: nothing to decide per group
            G.maxlen = a.ell_width;
            G.lpr = MODE == 1;
            G.full = (gi + 1) * 32 <= nrows;
        } else {
            G.maxlen = __reduce_max_sync(B200_FULL, G.len);
            G.lpr = G.nnz <= LIMIT && G.maxlen <= min(a.vector_threshold, kCsrMirror);
            G.full = G.nnz == 32 * G.maxlen;
        }
    };
    // lane-per-row, step 1 (one group ahead): column ids -> x gathers in flight
    auto prefetch = [&](const CsrGroup& G, double (&xq)[J]) {
        const int* pc = scol + (G.s & M);
        if (G.full) {
            B200_CSR_DISPATCH_N(G.maxlen, (lpr_gather<N_, ELL, true>(pc, G.len, xp, xq)));
        } else {
            B200_CSR_DISPATCH_N(G.maxlen, (lpr_gather<N_, ELL, false>(pc, G.len, xp, xq)));
        }
    };
    // lane-per-row, step 2: the k-ordered fma chain of this lane's row
    auto process_lpr = [&](const CsrGroup& G, const double (&xq)[J]) {
        double sum = 0.0;
        const double* pv = sval + (G.s & M);
        if (G.full) {
            B200_CSR_DISPATCH_N(G.maxlen, (sum = lpr_chain<N_, true>(pv, G.len, xq)));
        } else {
            B200_CSR_DISPATCH_N(G.maxlen, (sum = lpr_chain<N_, false>(pv, G.len, xq)));
        }
        if (G.maxlen > J) {  // rows longer than the prefetch depth
            const int* pc = scol + (G.s & M);
            for (int j = J; j < G.maxlen; j++) {
                if (j < G.len) {
                    const int c = pc[j];
                    if (!ELL || c >= 0) sum = fma(pv[j], __ldg(xp + c), sum);
                }
            }
        }
        return sum;
    };
    // sub-warp rows: S lanes per row, 32 / S rows of the group at a time -- for groups whose longest row has
    // at most 2 * S... 4 * S entries, where a whole warp per row would leave most lanes idle (9..64-entry rows
    // that do not fit the lane-per-row ring).  Lane j of a row's S lanes takes the entries j, j + S, ... in
    // k order, a butterfly over the S lanes closes the row.
    auto process_sub = [&](const CsrGroup& G, int rows_here, auto s_tag) {
        constexpr int S = decltype(s_tag)::value, R = 32 / S;
        double sum = 0.0;
        const int sub = lane / S, j0 = lane % S;
        for (int q0 = 0; q0 < rows_here; q0 += R) {
            const int qlast = min(q0 + R, rows_here) - 1;
            const int first = __shfl_sync(B200_FULL, G.s, q0);
            const int last_end = __shfl_sync(B200_FULL, G.s + G.len, qlast);
            release(first);
            ensure(last_end);
            const int r = min(q0 + sub, qlast);  // lanes past the last row redo it (their result is not used)
            const int rs = __shfl_sync(B200_FULL, G.s, r);
            const int re = rs + __shfl_sync(B200_FULL, G.len, r);
            double part = 0.0;
            for (int kk = rs + j0; kk < re; kk += S) {
                const int c = scol[kk & M];
                if (!ELL || c >= 0) part = fma(sval[kk & M], __ldg(xp + c), part);
            }
#pragma unroll
            for (int o = S / 2; o > 0; o >>= 1) part += __shfl_xor_sync(B200_FULL, part, o);
            const double mine = __shfl_sync(B200_FULL, part, ((lane - q0) & (R - 1)) * S);
            if (lane >= q0 && lane <= qlast) sum = mine;
        }
        return sum;
    };
    // warp-per-row: the rows of the group one after the other, streamed through the ring
    auto process_vec = [&](const CsrGroup& G, int rows_here) {
        if (G.maxlen <= 32) return process_sub(G, rows_here, std::integral_constant<int, 8>());
        if (G.maxlen <= 64) return process_sub(G, rows_here, std::integral_constant<int, 16>());
        double sum = 0.0;
        for (int q = 0; q < rows_here; q++) {
            const int qs = __shfl_sync(B200_FULL, G.s, q);
            const int qe = qs + __shfl_sync(B200_FULL, G.len, q);
            double part = 0.0;
            int k = qs;
            while (k < qe) {
                const int ce = min(qe, (k / WIN + 1) * WIN);  // stay inside one window
                release(k);
                ensure(ce);
                for (int kk = k + ((lane - (k - qs)) & 31); kk < ce; kk += 32) {
                    const int c = scol[kk & M];
                    if (!ELL || c >= 0) part = fma(sval[kk & M], __ldg(xp + c), part);
                }
                k = ce;
            }
            part = warp_sum(part);
            if (lane == q) sum = part;
        }
        return sum;
    };

    // Software pipeline over the groups (one copy of the code, one set of x registers):
    //   chain(cur) -> y -> recycle windows below nxt -> gather(nxt) into the registers just freed
    //   -> row extents of the group after nxt.
    // The x gathers of `nxt` are in flight while everything behind them runs; behind the last group
    // come empty sentinel groups, so there is no "has next" case.
    double xq[J];
    CsrGroup cur, nxt;
    double* yp = a.y + Ra + lane;
    const double alpha = a.alpha, beta = a.beta;
    constexpr bool dot = DOT;  // compiled separately: the plain SpMV pays nothing for the fusion
    const double* xrow = xp + Ra + lane;  // dot fusion: x at the lane's own row
    double dacc = 0.0;
    cur.lpr = true; cur.full = false; cur.s = cur.len = cur.gs = cur.nnz = cur.maxlen = 0;
    nxt = cur;
    fetch_extents(0);
    fetch_extents(1);
    fetch_extents(2);
    for (int gi = -2; gi < ngroups; gi++) {  // two warm-up turns fill the pipeline through the same code
        if (gi >= 0) {
            double sum;
            if (MODE == 1) sum = process_lpr(cur, xq);
            else if (MODE == 2) sum = process_vec(cur, min(32, nrows - gi * 32));
            else sum = cur.lpr ? process_lpr(cur, xq) : process_vec(cur, min(32, nrows - gi * 32));
            if (gi * 32 + lane < nrows) {
                const double yv = (beta == 0.0) ? alpha * sum : fma(alpha, sum, beta * *yp);
                __stcs(yp, yv);
                if (dot) dacc = fma(__ldg(xrow), yv, dacc);
            }
            yp += 32;
            xrow += 32;
        }
        if (gi >= -1) {
            release(nxt.gs);
            if (MODE == 1 || (MODE == 0 && nxt.lpr)) {
                ensure(nxt.gs + nxt.nnz);
                prefetch(nxt, xq);
            }
        }
        cur = nxt;
        make_group(gi + 2, nxt);
        fetch_extents(gi + 5);
    }
    if (dot) {  // fixed order: lanes (butterfly) -> one slot per item -> final pass in item order
        dacc = warp_sum(dacc);
        if (lane == 0) a.dot_partials[(long long)blockIdx.x * WARPS + warp] = dacc;
    }
}

}  // namespace b200
