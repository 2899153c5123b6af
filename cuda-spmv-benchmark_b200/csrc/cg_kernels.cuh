// cg_kernels.cuh -- fused CG building blocks for sm_100a.
//
// Reference sequence per iteration (src/solvers/cg_solver.cu:538-638): SpMV, dot_kernel +
// final_sum_kernel, scalar_divide, axpy, axpy_sub, dot + final_sum, check_convergence, blocking
// 4-byte D2H, scalar_divide, update_p, 8-byte D2D  -- 11 launches, 152 B/row, one host sync.
// Here: K1 (SpMV + p.Ap partials, stencil5.cuh) -> R (fixed-order final sum, alpha) ->
// K2 (x += alpha p, r -= alpha Ap, r.r partials) -> R (final sum, convergence, beta) ->
// K3 (p = r + beta p)  -- 5 launches, 128 B/row, no host sync (scalars stay on the device, the
// host polls a pinned status word a few iterations behind).
//
// Determinism: every partial sum has a fixed owner (CTA id), every CTA sums in a fixed order,
// the final pass adds partials in index order and ranks in rank order, so iteration counts are
// reproducible run to run and identical on every GPU of a multi-GPU solve.
#pragma once
#include "common.cuh"

namespace b200 {

#define B200_MAX_RANKS 16

// device-resident CG scalars (one per rank)
struct CGScalars {
    double rr_old;
    double rr_new;
    double pAp;
    double alpha;
    double beta;
    double b_norm;    // sqrt(r0.r0): the reference's "b_norm" (cg_solver.cu:527-528)
    double residual;  // sqrt(rr_new) of the last checked iteration
    int converged;
    int iterations;   // completed iterations (counts the converging one, cg_solver.cu:619)
    int error;        // a peer-flag wait timed out: later waits return at once
    int pad_;
    // device-measured duration of the reduction tails (final sum + rank exchange + scalar update),
    // indexed by RED_*: what used to be separate cg_reduce launches
    unsigned long long tail_ns[8];
    unsigned int tail_cnt[8];
    // ... and of what ran between the end of the previous tail and the begin of this one (the producing
    // kernel(s) of this reduction): per-phase times that cost no events and no synchronisation
    unsigned long long gap_ns[8];
    unsigned long long last_tail_end;
    // alpha of iteration j at [j & 7]: the STENCIL5 kernel retires several pending x updates at once
    // (ST_FUSED_X<m>, stencil5.cuh) and cg_finish_x_kernel the ones left at the end
    double alpha_hist[8];
};

// host-visible mirror (pinned, mapped), written by the reduction tail after every r.r
struct CGStatus {
    volatile int iterations;  // (iterations, converged): one aligned 8-byte word, stored at once; the host polls it
    volatile int converged;
    volatile double residual;
    volatile double b_norm;
    volatile int error;  // flag-wait timeout in a peer exchange
};

// exchange area of one rank, in that rank's device memory, written by its peers over NVLink
struct XchgArea {
    uint64_t slots[2][B200_MAX_RANKS][4];  // LL words: (seq << 32 | 32 data bits) x 2 per double, 2 doubles
    uint32_t halo_flag_prev;               // sequence number of the last halo written by rank-1
    uint32_t halo_flag_next;               // sequence number of the last halo written by rank+1
    uint32_t push_count[2];                // local: CTAs done with the halo push (prev, next)
    // Exchange sequence numbers live ON THE DEVICE: every rank executes the same sequence of
    // exchanges (kernels turn into no-ops on all ranks at the same iteration once the solve has
    // converged), so the counters stay in lockstep no matter how many no-op iterations each host
    // enqueues behind the convergence point.
    uint32_t red_seq;                      // scalar exchanges started by this rank
    uint32_t halo_seq;                     // halo pushes published by this rank
    uint32_t pad[26];
};

// arguments of the kernels that push halo elements into the neighbours' landing buffers
struct HaloPushArgs {
    const double* v_local;
    long long n_local;
    int halo;
    double* dst_prev;  // rank-1's halo_next landing buffer (NULL if rank == 0)
    double* dst_next;  // rank+1's halo_prev landing buffer (NULL if last rank)
    uint32_t* flag_prev;  // rank-1's halo_flag_next
    uint32_t* flag_next;  // rank+1's halo_flag_prev
    XchgArea* my_xchg;    // push_count, halo_seq
    const CGScalars* sc;
};

// halo copies of the search direction (deferred-x schedule): see cg_halo_dir_kernel
struct HaloDirArgs {
    const double* r_prev;  // landing buffers written by the neighbours (NULL at the ends)
    const double* r_next;
    const double* pold_prev;
    const double* pold_next;
    double* pnew_prev;
    double* pnew_next;
    int halo;
    const uint32_t* flag_prev;
    const uint32_t* flag_next;
    const uint32_t* seq_ptr;  // my_xchg->halo_seq: the push every neighbour must have published
    CGScalars* sc;
    int beta_zero;
};

// 4, 5: Jacobi PCG (z = D^-1 r formed on the fly); 6, 7: PCG with a stored z (block-Jacobi): the r.r tail only
// tests convergence, beta comes out of the r.z tail that follows the preconditioner solve
enum { RED_RR0 = 0, RED_PAP = 1, RED_RR = 2, RED_SUM = 3, RED_RZ0 = 4, RED_PCG = 5, RED_RRC = 6, RED_RZ = 7 };
enum { RED_PUSH = 1, RED_COMBINE = 2 };
#define B200_RED_GROUP 256

// Reduction context of a launch: where the per-CTA partial sums go and what happens to their total.
// Fused form ("tail"): every CTA takes a ticket after writing its partial; the last CTA of each group
// of 256 adds the group in a fixed order, the last group closer adds the group sums in a fixed order,
// exchanges the total with the other ranks and applies the scalar recurrence `which` -- no separate
// reduce launch.  The order of every addition is independent of which CTA happens to do it, so the
// result is bit-reproducible and identical to cg_reduce_kernel over the same partials.
struct TailArgs {
    int which;    // RED_*; < 0: this launch only writes its partials (cg_reduce_kernel follows)
    int phases;   // RED_PUSH: local total + LL stores to the peers; RED_COMBINE: wait + recurrence
    int two;      // a second sum travels along (PCG r.z, checksum sum of squares)
    double tol;
    CGScalars* sc;
    CGStatus* status;  // may be NULL
    double* out;       // RED_SUM: out[0] (and out[1]) receive the total(s)
    double* stash;     // [2] local totals between a PUSH-only tail and the COMBINE launch
    double* partials;
    double* partials_b;
    double* gsum;      // [2][cap_groups]
    uint32_t* tickets; // [1 + cap_groups], zero between launches (self-resetting)
    int cap_groups;
    int rank, world;
    XchgArea* my_xchg;
    XchgArea* peer_xchg[B200_MAX_RANKS];
    // halo publication by the final CTA (kernels that push the edges of a vector to the neighbours)
    int publish;
    uint32_t* flag_prev;  // rank-1's halo_flag_next
    uint32_t* flag_next;  // rank+1's halo_flag_prev
};

// programmatic dependent launch: a kernel launched with the stream-serialisation attribute may be
// scheduled while its predecessor drains; everything it reads must come after griddep_wait().
// NOTE: the scalar block is therefore never passed as `const ... __restrict__`: that qualifier turns its
// reads into invariant loads (LDG.CONSTANT), which the compiler is free to schedule ABOVE the wait -- it did,
// in cg_finish_x_kernel: the iteration count of the previous kernel's tail was read before that tail had run
// (tests/test_sass_pdl.py scans the SASS of every kernel for global loads in front of the wait).
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// fixed-order sum of partials [first, first + 256) (clipped to n): lane l adds l, l+32, ... in order,
// then the butterfly.  Warp-collective; every lane returns the sum.
__device__ __forceinline__ double group_sum_fixed(const double* p, long long first, long long n) {
    const int lane = threadIdx.x & 31;
    double v[B200_RED_GROUP / 32];
#pragma unroll
    for (int j = 0; j < B200_RED_GROUP / 32; j++) {
        const long long i = first + lane + 32 * j;
        v[j] = (i < n) ? __ldcg(p + i) : 0.0;
    }
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < B200_RED_GROUP / 32; j++) s += v[j];
    return warp_sum(s);
}

// fixed-order sum of n group sums: lane l adds l, l+32, ... in order, then the butterfly
__device__ __forceinline__ double top_sum_fixed(const double* g, int n) {
    const int lane = threadIdx.x & 31;
    double s = 0.0;
    for (int i = lane; i < n; i += 32) s += __ldcg(g + i);
    return warp_sum(s);
}

// Ticketed two-level reduction at the end of a producing kernel.  Called by ALL threads of every CTA
// (blockDim.x >= 32) with the CTA's partial(s) valid in thread 0.  Returns true, in all threads of
// exactly one CTA, once every CTA of the grid has contributed; tot[0..1] (shared) then hold the totals.
__device__ __forceinline__ bool reduce_tickets(const TailArgs& t, double acc, double acc_b, double* tot, int* flags) {
    const unsigned nb = gridDim.x;
    const unsigned g = blockIdx.x / B200_RED_GROUP, ng = (nb + B200_RED_GROUP - 1) / B200_RED_GROUP;
    if (threadIdx.x == 0) {
        t.partials[blockIdx.x] = acc;
        if (t.two) t.partials_b[blockIdx.x] = acc_b;
        __threadfence();
        const unsigned gsize = min((unsigned)B200_RED_GROUP, nb - g * B200_RED_GROUP);
        flags[0] = (atomicAdd(&t.tickets[1 + g], 1u) + 1u == gsize);
    }
    __syncthreads();
    if (!flags[0]) return false;
    if (threadIdx.x < 32) {  // this CTA closes group g
        __threadfence();
        double s = group_sum_fixed(t.partials, (long long)g * B200_RED_GROUP, nb);
        double sb = t.two ? group_sum_fixed(t.partials_b, (long long)g * B200_RED_GROUP, nb) : 0.0;
        int last = 0;
        if (threadIdx.x == 0) {
            t.gsum[g] = s;
            if (t.two) t.gsum[t.cap_groups + g] = sb;
            t.tickets[1 + g] = 0;
            __threadfence();
            last = (atomicAdd(&t.tickets[0], 1u) + 1u == ng);
        }
        last = __shfl_sync(B200_FULL, last, 0);
        if (last) {  // ... and the grid
            __threadfence();
            s = top_sum_fixed(t.gsum, (int)ng);
            sb = t.two ? top_sum_fixed(t.gsum + t.cap_groups, (int)ng) : 0.0;
            if (threadIdx.x == 0) { tot[0] = s; tot[1] = sb; t.tickets[0] = 0; }
        }
        if (threadIdx.x == 0) flags[1] = last;
    }
    __syncthreads();
    return flags[1] != 0;
}

// What the CTA holding the grand total does with it: publish the halo sequence number, exchange the
// total with the other ranks (LL protocol: data and sequence number travel in the same 8-byte store,
// so no fence; every rank adds the P slots in rank order => identical result everywhere), then the
// scalar recurrences the reference spreads over scalar_divide_kernel / check_convergence_kernel / a
// D2D copy (cg_solver.cu:414-431,560,594-637) and, multi-GPU, cublasDdot + MPI_Allreduce
// (cg_solver_mgpu_partitioned.cu:145-154).  Called by all threads of one CTA (>= 32 threads).
// sh: >= 2 * B200_MAX_RANKS + 4 doubles of shared memory.
__device__ __forceinline__ void cg_tail(const TailArgs& a, double total, double total_b, double* sh) {
    const int t = threadIdx.x;
    const uint64_t t_begin = globaltimer_ns();
    CGScalars* sc = a.sc;
    uint32_t* sh_seq = reinterpret_cast<uint32_t*>(sh + 2 * B200_MAX_RANKS);
    if (a.world > 1) {
        if (t == 0) {
            uint32_t seq = a.my_xchg->red_seq;
            if (a.phases & RED_PUSH) { seq += 1u; a.my_xchg->red_seq = seq; }
            *sh_seq = seq;
        }
        __syncthreads();
    }
    const uint32_t seq = (a.world > 1) ? *sh_seq : 0u;
    if (a.phases & RED_PUSH) {
        if (t == 0 && !(a.phases & RED_COMBINE)) { a.stash[0] = total; a.stash[1] = total_b; }
        // the scalar goes out FIRST: the peers' tails are waiting for it, while the halo sequence number
        // below is only looked at by their next kernel
        if (a.world > 1 && t < a.world) {
            const uint64_t tag = (uint64_t)seq << 32;
            uint64_t* dst = a.peer_xchg[t]->slots[seq & 1][a.rank];
            const uint64_t bits = (uint64_t)__double_as_longlong(total);
            st_volatile_u64(dst, tag | (bits & 0xffffffffull));
            st_volatile_u64(dst + 1, tag | (bits >> 32));
            if (a.two) {
                const uint64_t bits_b = (uint64_t)__double_as_longlong(total_b);
                st_volatile_u64(dst + 2, tag | (bits_b & 0xffffffffull));
                st_volatile_u64(dst + 3, tag | (bits_b >> 32));
            }
        }
    }
    if (a.publish && t == 32 % blockDim.x) {
        // every CTA fenced its peer stores (system scope) before it took its ticket, and this CTA saw all the
        // tickets: ONE system fence orders those stores before both arrival words.  Done by a lane outside
        // the warp that exchanges the scalars, so it overlaps the wait below.
        const uint32_t hs = a.my_xchg->halo_seq + 1u;
        a.my_xchg->halo_seq = hs;
        __threadfence_system();
        if (a.flag_prev != nullptr) st_relaxed_sys(a.flag_prev, hs);
        if (a.flag_next != nullptr) st_relaxed_sys(a.flag_next, hs);
    }
    if (!(a.phases & RED_COMBINE)) return;
    if (a.world > 1) {
        if (t < a.world) {
            const uint64_t* src = a.my_xchg->slots[seq & 1][t];
            const uint64_t t0 = globaltimer_ns();
            const bool dead = (sc != nullptr && sc->error != 0);
            const int nw = a.two ? 4 : 2;
            uint64_t w[4] = {0, 0, 0, 0};
            for (;;) {
                bool ok = true;
                for (int k = 0; k < nw; k++) {
                    w[k] = ld_volatile_u64(src + k);
                    ok = ok && ((uint32_t)(w[k] >> 32) == seq);
                }
                if (ok) break;
                if (dead || globaltimer_ns() - t0 > 8000000000ull) {  // 8 s: report, never hang
                    if (sc) sc->error = 1;
                    if (a.status) a.status->error = 1;
                    break;
                }
            }
            sh[t] = __longlong_as_double((long long)((w[1] << 32) | (w[0] & 0xffffffffull)));
            sh[B200_MAX_RANKS + t] = __longlong_as_double((long long)((w[3] << 32) | (w[2] & 0xffffffffull)));
        }
        __syncthreads();
        if (t == 0) {
            total = 0.0; total_b = 0.0;
            for (int r = 0; r < a.world; r++) { total += sh[r]; total_b += sh[B200_MAX_RANKS + r]; }  // rank order
        }
    }
    if (t != 0) return;
    if (a.which == RED_SUM) {
        if (a.out) { a.out[0] = total; if (a.two) a.out[1] = total_b; }
    } else if (a.which == RED_RR0) {
        sc->rr_old = total;
        sc->b_norm = sqrt(total);
        sc->residual = sc->b_norm;
        if (a.status) { a.status->b_norm = sc->b_norm; a.status->residual = sc->b_norm; }
    } else if (a.which == RED_PAP) {
        sc->pAp = total;
        sc->alpha = sc->rr_old / total;  // PCG keeps rho = r.z in rr_old
        sc->alpha_hist[sc->iterations & 7] = sc->alpha;
    } else if (a.which == RED_RZ0) {
        sc->rr_old = total;  // rho_0 = r0.z0 (b_norm was set by RED_RR0)
    } else if (a.which == RED_RZ) {
        sc->beta = total / sc->rr_old;  // rho_new / rho
        sc->rr_old = total;
    } else {  // RED_RR, RED_PCG (total = r.r, total_b = r.z), RED_RRC (convergence test only)
        sc->rr_new = total;
        const double res = sqrt(total);
        sc->residual = res;
        const int it = sc->iterations + 1;
        sc->iterations = it;
        const int conv = (res / sc->b_norm < a.tol) ? 1 : 0;
        if (conv) {
            sc->converged = 1;
        } else if (a.which == RED_PCG) {
            sc->beta = total_b / sc->rr_old;
            sc->rr_old = total_b;
        } else if (a.which == RED_RR) {
            sc->beta = total / sc->rr_old;
            sc->rr_old = total;
        }
        if (a.status) {
            // (iterations, converged) travel in ONE 8-byte store: the host polls that word and needs no
            // ordering against the other fields (it reads those after a stream synchronisation), so no
            // system fence sits on the critical path between K2 and the next SpMV
            a.status->residual = res;
            if (sc->error) a.status->error = 1;
            *reinterpret_cast<volatile unsigned long long*>(&a.status->iterations) =
                (unsigned long long)(unsigned int)it | ((unsigned long long)(unsigned int)conv << 32);
        }
    }
    if (sc != nullptr && a.which >= 0 && a.which < 8) {
        const uint64_t now = globaltimer_ns();
        if (sc->last_tail_end != 0 && t_begin > sc->last_tail_end) sc->gap_ns[a.which] += t_begin - sc->last_tail_end;
        sc->tail_ns[a.which] += now - t_begin;
        sc->tail_cnt[a.which] += 1u;
        sc->last_tail_end = now;
    }
}

// end of a producing kernel: tickets, and the tail in the CTA that ends up with the total
#define B200_TAIL_SHARED                                   \
    __shared__ double tail_sh_[2 * B200_MAX_RANKS + 4];    \
    __shared__ double tail_tot_[2];                        \
    __shared__ int tail_flags_[2]
#define B200_TAIL_FINISH(t, acc, acc_b)                                                   \
    do {                                                                                  \
        if ((t).tickets != nullptr && reduce_tickets((t), (acc), (acc_b), tail_tot_, tail_flags_)) \
            cg_tail((t), tail_tot_[0], tail_tot_[1], tail_sh_);                           \
    } while (0)

// Stand-alone form: one warp-sized CTA per group of 256 partials adds its group, the last one to finish
// adds the group sums and runs the tail -- the same two-level order as reduce_tickets.  Launched
// programmatically behind the producing kernel (griddepcontrol: no launch gap).  Used behind the
// STENCIL5 kernels (O(1e5) short CTAs: a ticket per producer CTA would cost more than this launch),
// behind operators without a fused tail (run_device), for the COMBINE half when several ranks share
// one device and one stream (a wait inside the producer could never be satisfied there), and for
// exchanges without partials (barrier).  gridDim.x = max(1, ceil(n_partials / 256)).
struct ReduceArgs {
    TailArgs t;
    int n_partials;
};

__global__ void __launch_bounds__(32) cg_reduce_kernel(const ReduceArgs a) {
    __shared__ double sh[2 * B200_MAX_RANKS + 4];
    griddep_wait();
    const TailArgs& t = a.t;
    if (t.which != RED_SUM && t.sc->converged) return;
    griddep_launch();
    const int lane = threadIdx.x;
    double total = 0.0, total_b = 0.0;
    if (t.phases & RED_PUSH) {
        const int n = a.n_partials, ng = (n + B200_RED_GROUP - 1) / B200_RED_GROUP;
        if (ng > 0) {
            const int g = blockIdx.x;
            const double s = group_sum_fixed(t.partials, (long long)g * B200_RED_GROUP, n);
            const double sb = t.two ? group_sum_fixed(t.partials_b, (long long)g * B200_RED_GROUP, n) : 0.0;
            int last = 0;
            if (lane == 0) {
                t.gsum[g] = s;
                if (t.two) t.gsum[t.cap_groups + g] = sb;
                __threadfence();
                last = (atomicAdd(&t.tickets[0], 1u) + 1u == (unsigned)ng);
            }
            last = __shfl_sync(B200_FULL, last, 0);
            if (!last) return;
            __threadfence();
            total = top_sum_fixed(t.gsum, ng);
            total_b = t.two ? top_sum_fixed(t.gsum + t.cap_groups, ng) : 0.0;
            if (lane == 0) t.tickets[0] = 0;
        }
    } else {
        total = t.stash[0];
        total_b = t.stash[1];
    }
    cg_tail(t, total, total_b, sh);
}

// K2: x += alpha p ; r -= alpha Ap ; partials[cta] = sum r^2 over the CTA's tiles.
// reference: axpy_kernel_device + axpy_sub_kernel_device + dot_kernel (cg_solver.cu:59-78,110-132)
// contracted the same way (fma(alpha,p,x), fma(-alpha,Ap,r)).
template <int VEC>
__global__ void __launch_bounds__(256) cg_update_xr_kernel(long long n, const CGScalars* sc,
                                                           const double* __restrict__ p,
                                                           const double* __restrict__ Ap, double* __restrict__ x,
                                                           double* __restrict__ r, const TailArgs tail) {
    __shared__ double scratch[8];
    B200_TAIL_SHARED;
    griddep_wait();
    if (sc->converged) return;
    const double alpha = sc->alpha, nalpha = -alpha;
    constexpr int UNROLL = 4;
    const long long tile = 256LL * VEC * UNROLL;
    double acc = 0.0;
    for (long long base = (long long)blockIdx.x * tile; base < n; base += (long long)gridDim.x * tile) {
        if (VEC == 2) {
            double2 pv[UNROLL], av[UNROLL], xv[UNROLL], rv[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                const long long i = base + ((long long)u * 256 + threadIdx.x) * 2;
                if (i + 1 < n) {
                    pv[u] = __ldcs(reinterpret_cast<const double2*>(p + i));
                    av[u] = __ldcs(reinterpret_cast<const double2*>(Ap + i));
                    xv[u] = __ldcs(reinterpret_cast<const double2*>(x + i));
                    rv[u] = __ldcs(reinterpret_cast<const double2*>(r + i));
                } else if (i < n) {
                    pv[u] = make_double2(p[i], 0.0); av[u] = make_double2(Ap[i], 0.0);
                    xv[u] = make_double2(x[i], 0.0); rv[u] = make_double2(r[i], 0.0);
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                const long long i = base + ((long long)u * 256 + threadIdx.x) * 2;
                if (i < n) {
                    xv[u].x = fma(alpha, pv[u].x, xv[u].x);
                    rv[u].x = fma(nalpha, av[u].x, rv[u].x);
                    acc = fma(rv[u].x, rv[u].x, acc);
                    if (i + 1 < n) {
                        xv[u].y = fma(alpha, pv[u].y, xv[u].y);
                        rv[u].y = fma(nalpha, av[u].y, rv[u].y);
                        acc = fma(rv[u].y, rv[u].y, acc);
                        *reinterpret_cast<double2*>(x + i) = xv[u];
                        *reinterpret_cast<double2*>(r + i) = rv[u];
                    } else {
                        x[i] = xv[u].x;
                        r[i] = rv[u].x;
                    }
                }
            }
        } else {
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                const long long i = base + (long long)u * 256 + threadIdx.x;
                if (i < n) {
                    const double xn = fma(alpha, p[i], x[i]);
                    const double rn = fma(nalpha, Ap[i], r[i]);
                    x[i] = xn;
                    r[i] = rn;
                    acc = fma(rn, rn, acc);
                }
            }
        }
    }
    griddep_launch();
    acc = block_sum(acc, scratch);
    B200_TAIL_FINISH(tail, acc, 0.0);
}

// ---- "deferred x" schedule (4 launches, 112 B/row per iteration instead of 5 launches, 128 B/row) -------
// The stencil kernel in ST_FUSED mode forms p = r + beta p_old while it loads the rows, retires
// x += alpha_prev p_old on the way and writes the new p once (stencil5.cuh).  What is left of K2 is
//   K2r: r -= alpha Ap ; partials r.r     (24 B/row; the p and x streams of K2 are gone)
// with the SAME tiling and summation order as cg_update_xr_kernel, so r.r -- and with it alpha, beta
// and every iterate -- is bit-identical to the classic schedule.  Multi-GPU (PUSH): the CTAs that
// produce the first / last `halo` elements of r store them straight into the neighbours' landing
// buffers over NVLink; the last CTA publishes the arrival epoch (as cg_update_p_push_kernel does for p).
template <int VEC, bool PUSH>
__global__ void __launch_bounds__(256) cg_update_r_kernel(long long n, const CGScalars* sc,
                                                          const double* __restrict__ Ap, double* __restrict__ r,
                                                          const HaloPushArgs h, const TailArgs tail) {
    __shared__ double scratch[8];
    B200_TAIL_SHARED;
    griddep_wait();
    if (sc->converged) return;
    const double nalpha = -sc->alpha;
    constexpr int UNROLL = 4;
    const long long tile = 256LL * VEC * UNROLL;
    const long long next_lo = n - h.halo;
    bool pushed = false;
    auto push = [&](long long i, double v) {
        if (PUSH) {
            if (h.dst_prev != nullptr && i < h.halo) { h.dst_prev[i] = v; pushed = true; }
            if (h.dst_next != nullptr && i >= next_lo) { h.dst_next[i - next_lo] = v; pushed = true; }
        }
    };
    double acc = 0.0;
    for (long long base = (long long)blockIdx.x * tile; base < n; base += (long long)gridDim.x * tile) {
        if (VEC == 2) {
            double2 av[UNROLL], rv[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                const long long i = base + ((long long)u * 256 + threadIdx.x) * 2;
                if (i + 1 < n) {
                    av[u] = __ldcs(reinterpret_cast<const double2*>(Ap + i));
                    rv[u] = __ldcs(reinterpret_cast<const double2*>(r + i));
                } else if (i < n) {
                    av[u] = make_double2(Ap[i], 0.0);
                    rv[u] = make_double2(r[i], 0.0);
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                const long long i = base + ((long long)u * 256 + threadIdx.x) * 2;
                if (i < n) {
                    rv[u].x = fma(nalpha, av[u].x, rv[u].x);
                    acc = fma(rv[u].x, rv[u].x, acc);
                    push(i, rv[u].x);
                    if (i + 1 < n) {
                        rv[u].y = fma(nalpha, av[u].y, rv[u].y);
                        acc = fma(rv[u].y, rv[u].y, acc);
                        push(i + 1, rv[u].y);
                        *reinterpret_cast<double2*>(r + i) = rv[u];
                    } else {
                        r[i] = rv[u].x;
                    }
                }
            }
        } else {
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                const long long i = base + (long long)u * 256 + threadIdx.x;
                if (i < n) {
                    const double rn = fma(nalpha, Ap[i], r[i]);
                    r[i] = rn;
                    acc = fma(rn, rn, acc);
                    push(i, rn);
                }
            }
        }
    }
    griddep_launch();
    // the edge stores must be visible system-wide before this CTA's ticket: the CTA that ends up with
    // the total publishes the halo sequence number to the neighbours (cg_tail).  Only the threads that
    // stored to a peer pay for the fence (a handful of CTAs at the two ends of the band).
    if (PUSH && pushed) __threadfence_system();
    acc = block_sum(acc, scratch);  // contains a __syncthreads
    B200_TAIL_FINISH(tail, acc, 0.0);
}

// Halo part of the direction update (multi-GPU, deferred-x schedule): the neighbours pushed the edges
// of their NEW r; the halo copies of p follow the same recurrence as the local part,
// p_halo = fma(beta, p_halo_old, r_halo), kept in two local ping-pong buffers.  beta_zero: first
// direction (p0 = r0).  One small CTA group; waits (bounded) for the arrival epochs first.

__global__ void __launch_bounds__(256) cg_halo_dir_kernel(const HaloDirArgs a) {
    griddep_wait();
    if (a.sc->converged) return;
    if (threadIdx.x == 0) {
        const uint64_t t0 = globaltimer_ns();
        const uint32_t epoch = __ldcg(a.seq_ptr);
        const uint32_t* fl[2] = {a.r_prev ? a.flag_prev : nullptr, a.r_next ? a.flag_next : nullptr};
        for (int d = 0; d < 2; d++) {
            if (!fl[d]) continue;
            while ((int32_t)(ld_acquire_sys(fl[d]) - epoch) < 0) {
                if (a.sc->error || globaltimer_ns() - t0 > 8000000000ull) { a.sc->error = 1; break; }
                __nanosleep(64);
            }
        }
    }
    __syncthreads();
    const double beta = a.beta_zero ? 0.0 : a.sc->beta;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < a.halo; i += gridDim.x * 256) {
        if (a.r_prev) a.pnew_prev[i] = a.beta_zero ? __ldcg(a.r_prev + i) : fma(beta, a.pold_prev[i], __ldcg(a.r_prev + i));
        if (a.r_next) a.pnew_next[i] = a.beta_zero ? __ldcg(a.r_next + i) : fma(beta, a.pold_next[i], __ldcg(a.r_next + i));
    }
}

// After the last iteration of the deferred-x schedule x still lacks the updates that no SpMV launch has
// retired.  Direction j lives in pbuf[j % nbuf] (nbuf = depth + 1); the launches of iterations depth, 2 depth,
// ... each retired the `depth` updates before them, so with c completed iterations (known on the device)
// the pending ones are j = depth * floor((c - 1) / depth) .. c - 1, applied oldest first.
// only_if_converged: the K3x schedule (below) retires x inside the p update of the same iteration (or of every
// depth-th iteration), so the update of the LAST iteration is pending only when the convergence test stopped
// the loop in front of that kernel.
struct FinishXArgs {
    const double* pbuf[5];
    int nbuf;
    int depth;
    int only_if_converged;
};
__global__ void __launch_bounds__(256) cg_finish_x_kernel(long long n, const CGScalars* sc,
                                                          const FinishXArgs a, double* __restrict__ x) {
    griddep_wait();
    const int c = sc->iterations;
    if (c <= 0) return;
    // K3x schedule, solve cut off by max_iters: the K3x launch of the last iteration did run (and retired, if
    // it was its turn), so the retired prefix is depth * floor(c / depth) -- nothing pending at depth 1
    const int first = (a.only_if_converged && !sc->converged) ? a.depth * (c / a.depth) : a.depth * ((c - 1) / a.depth);
    const int cnt = c - first;  // 0 .. depth
    if (cnt <= 0) return;
    double al[4] = {0.0, 0.0, 0.0, 0.0};
    const double* p[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int k = 0; k < 4; k++)
        if (k < cnt) {
            const int j = first + k;
            al[k] = (j == c - 1) ? sc->alpha : sc->alpha_hist[j & 7];
            p[k] = a.pbuf[j % a.nbuf];
        }
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        double xv = x[i];
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (k < cnt) xv = fma(al[k], p[k][i], xv);
        x[i] = xv;
    }
}

// K3: p = r + beta p   (reference update_p_kernel, cg_solver.cu:91-96: fma(beta,p,r))
template <int VEC>
__global__ void __launch_bounds__(256) cg_update_p_kernel(long long n, const CGScalars* sc,
                                                          const double* __restrict__ r, double* __restrict__ p) {
    griddep_wait();
    if (sc->converged) return;
    const double beta = sc->beta;
    constexpr int UNROLL = 4;
    const long long tile = 256LL * VEC * UNROLL;
    for (long long base = (long long)blockIdx.x * tile; base < n; base += (long long)gridDim.x * tile) {
        if (VEC == 2) {
            double2 pv[UNROLL], rv[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                const long long i = base + ((long long)u * 256 + threadIdx.x) * 2;
                if (i + 1 < n) {
                    pv[u] = __ldcs(reinterpret_cast<const double2*>(p + i));
                    rv[u] = __ldcs(reinterpret_cast<const double2*>(r + i));
                } else if (i < n) {
                    pv[u] = make_double2(p[i], 0.0);
                    rv[u] = make_double2(r[i], 0.0);
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                const long long i = base + ((long long)u * 256 + threadIdx.x) * 2;
                if (i + 1 < n) {
                    pv[u].x = fma(beta, pv[u].x, rv[u].x);
                    pv[u].y = fma(beta, pv[u].y, rv[u].y);
                    *reinterpret_cast<double2*>(p + i) = pv[u];
                } else if (i < n) {
                    p[i] = fma(beta, pv[u].x, rv[u].x);
                }
            }
        } else {
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                const long long i = base + (long long)u * 256 + threadIdx.x;
                if (i < n) p[i] = fma(beta, p[i], r[i]);
            }
        }
    }
}

// K3x: p = r + beta p AND x += alpha p_old in one pass (operators without a fused SpMV: K2 shrinks to K2r,
// 24 B/row, and this kernel moves 40 B/row instead of K3's 24 -- 64 instead of 72 B/row for the BLAS-1
// part of an iteration).  Same fma's as K2 / K3 (cg_solver.cu:59-66,91-96), so the iterates are
// bit-identical to the classic schedule.  alpha is still the alpha of this iteration: the next p.Ap tail
// has not run yet.
template <int VEC>
__global__ void __launch_bounds__(256) cg_update_px_kernel(long long n, const CGScalars* sc,
                                                           const double* __restrict__ r, double* __restrict__ p,
                                                           double* __restrict__ x) {
    griddep_wait();
    if (sc->converged) return;
    const double alpha = sc->alpha, beta = sc->beta;
    constexpr int UNROLL = 4;
    const long long tile = 256LL * VEC * UNROLL;
    for (long long base = (long long)blockIdx.x * tile; base < n; base += (long long)gridDim.x * tile) {
        if (VEC == 2) {
            double2 pv[UNROLL], rv[UNROLL], xv[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                const long long i = base + ((long long)u * 256 + threadIdx.x) * 2;
                if (i + 1 < n) {
                    pv[u] = __ldcs(reinterpret_cast<const double2*>(p + i));
                    rv[u] = __ldcs(reinterpret_cast<const double2*>(r + i));
                    xv[u] = __ldcs(reinterpret_cast<const double2*>(x + i));
                } else if (i < n) {
                    pv[u] = make_double2(p[i], 0.0);
                    rv[u] = make_double2(r[i], 0.0);
                    xv[u] = make_double2(x[i], 0.0);
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                const long long i = base + ((long long)u * 256 + threadIdx.x) * 2;
                if (i + 1 < n) {
                    xv[u].x = fma(alpha, pv[u].x, xv[u].x);
                    xv[u].y = fma(alpha, pv[u].y, xv[u].y);
                    pv[u].x = fma(beta, pv[u].x, rv[u].x);
                    pv[u].y = fma(beta, pv[u].y, rv[u].y);
                    *reinterpret_cast<double2*>(x + i) = xv[u];
                    *reinterpret_cast<double2*>(p + i) = pv[u];
                } else if (i < n) {
                    x[i] = fma(alpha, pv[u].x, xv[u].x);
                    p[i] = fma(beta, pv[u].x, rv[u].x);
                }
            }
        } else {
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                const long long i = base + (long long)u * 256 + threadIdx.x;
                if (i < n) {
                    const double po = p[i];
                    x[i] = fma(alpha, po, x[i]);
                    p[i] = fma(beta, po, r[i]);
                }
            }
        }
    }
}

// K3x with the x stream amortised over `depth` iterations (the generic-operator twin of ST_FUSED_X<m>,
// stencil5.cuh): p_new = r + beta p_old goes to a FRESH direction buffer (direction j lives in buffer
// j mod (depth + 1)), and the launch that forms direction j with j % depth == 0 (NX = depth) retires the updates
// of directions j - depth .. j - 1 in one read-modify-write of x, oldest first -- the same fma chain as `depth`
// separate launches of K3x.  24 + (16 + 8 (d - 1)) / d B/row instead of 40: 34 at d = 4.
// alpha of direction j - 1 is this iteration's alpha; the older ones come from the history the p.Ap tail keeps
// (sc->iterations == j here: the r.r tail of this iteration has run).
struct UpdatePxArgs {
    const double* older[3];  // p_{j-2}, p_{j-3}, p_{j-4}
};
template <int NX>
__global__ void __launch_bounds__(256) cg_update_px_depth_kernel(long long n, const CGScalars* sc,
                                                                 const double* __restrict__ r,
                                                                 const double* __restrict__ p_old,
                                                                 double* __restrict__ p_new, double* __restrict__ x,
                                                                 const UpdatePxArgs a) {
    griddep_wait();
    if (sc->converged) return;
    const double beta = sc->beta;
    double al[NX > 0 ? NX : 1];
    if (NX >= 1) al[0] = sc->alpha;
    if (NX >= 2) {
        const int j = sc->iterations;
#pragma unroll
        for (int k = 1; k < NX; k++) al[k] = sc->alpha_hist[(j - 1 - k) & 7];
    }
    constexpr int UNROLL = 4;
    const long long tile = 256LL * UNROLL;
    for (long long base = (long long)blockIdx.x * tile; base < n; base += (long long)gridDim.x * tile) {
        double pv[UNROLL], rv[UNROLL], xv[UNROLL], ov[UNROLL][NX > 1 ? NX - 1 : 1];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const long long i = base + (long long)u * 256 + threadIdx.x;
            if (i < n) {
                pv[u] = __ldcs(p_old + i);
                rv[u] = __ldcs(r + i);
                if (NX >= 1) xv[u] = __ldcs(x + i);
#pragma unroll
                for (int k = 1; k < NX; k++) ov[u][k - 1] = __ldcs(a.older[k - 1] + i);
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const long long i = base + (long long)u * 256 + threadIdx.x;
            if (i < n) {
                if (NX >= 1) {
                    double t = xv[u];
#pragma unroll
                    for (int k = NX - 1; k >= 1; k--) t = fma(al[k], ov[u][k - 1], t);
                    x[i] = fma(al[0], pv[u], t);
                }
                p_new[i] = fma(beta, pv[u], rv[u]);
            }
        }
    }
}

// ---- Jacobi-preconditioned CG (not in the reference: its README lists preconditioning as the next step) ----
// z = D^-1 r is never stored: K2p forms it for the r.z partials, K3p forms it again for p = z + beta p.
// dinv[i] = 1 / A(i,i) from the CSR / ELLPACK arrays (row_ptr == NULL: ELLPACK of width `width`);
// *err is set when a row has no (or a zero) diagonal entry.
__global__ void __launch_bounds__(256) pcg_diag_inv_kernel(long long n_local, long long row_offset,
                                                           const int* __restrict__ row_ptr, int width,
                                                           const int* __restrict__ col_idx,
                                                           const double* __restrict__ values,
                                                           double* __restrict__ dinv, int* __restrict__ err) {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n_local; i += (long long)gridDim.x * 256) {
        const long long s = row_ptr ? row_ptr[i] : i * width, e = row_ptr ? row_ptr[i + 1] : (i + 1) * width;
        const unsigned int want = (unsigned int)(row_offset + i);  // columns are stored modulo 2^32
        double d = 0.0;
        for (long long k = s; k < e; k++)
            if ((unsigned int)col_idx[k] == want && (row_ptr || col_idx[k] >= 0)) d = values[k];
        if (d == 0.0) { *err = 1; dinv[i] = 0.0; }
        else dinv[i] = 1.0 / d;
    }
}

// after the residual init (r = b - A x0): p0 = z0 = dinv r0, partials of rho_0 = r0.z0
// dinv == NULL: z = zin (already solved, block-Jacobi)
__global__ void __launch_bounds__(256) pcg_init_kernel(long long n, const double* __restrict__ r,
                                                       const double* __restrict__ dinv, const double* __restrict__ zin,
                                                       double* __restrict__ p, const TailArgs tail) {
    __shared__ double scratch[8];
    B200_TAIL_SHARED;
    double acc = 0.0;
    const long long tile = 1024;
    for (long long base = (long long)blockIdx.x * tile; base < n; base += (long long)gridDim.x * tile) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const long long i = base + u * 256 + threadIdx.x;
            if (i < n) {
                const double rv = r[i], z = dinv ? dinv[i] * rv : zin[i];
                p[i] = z;
                acc = fma(rv, z, acc);
            }
        }
    }
    acc = block_sum(acc, scratch);
    B200_TAIL_FINISH(tail, acc, 0.0);
}

// K2p: x += alpha p ; r -= alpha Ap ; partials of r.r (convergence) and r.z (rho), z = dinv r
__global__ void __launch_bounds__(256) pcg_update_xr_kernel(long long n, const CGScalars* sc,
                                                            const double* __restrict__ p,
                                                            const double* __restrict__ Ap,
                                                            const double* __restrict__ dinv, double* __restrict__ x,
                                                            double* __restrict__ r, const TailArgs tail) {
    __shared__ double scratch[8];
    B200_TAIL_SHARED;
    if (sc->converged) return;
    const double alpha = sc->alpha, nalpha = -alpha;
    double arr = 0.0, arz = 0.0;
    const long long tile = 1024;
    for (long long base = (long long)blockIdx.x * tile; base < n; base += (long long)gridDim.x * tile) {
        double pv[4], av[4], xv[4], rv[4], dv[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const long long i = base + u * 256 + threadIdx.x;
            if (i < n) { pv[u] = __ldcs(p + i); av[u] = __ldcs(Ap + i); xv[u] = __ldcs(x + i); rv[u] = __ldcs(r + i); dv[u] = __ldcs(dinv + i); }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const long long i = base + u * 256 + threadIdx.x;
            if (i < n) {
                x[i] = fma(alpha, pv[u], xv[u]);
                const double rn = fma(nalpha, av[u], rv[u]);
                r[i] = rn;
                arr = fma(rn, rn, arr);
                arz = fma(rn, dv[u] * rn, arz);
            }
        }
    }
    arr = block_sum(arr, scratch);
    __syncthreads();
    arz = block_sum(arz, scratch);
    B200_TAIL_FINISH(tail, arr, arz);
}

// K3p: p = z + beta p, z = dinv r.  Multi-GPU (h.my_xchg != NULL): the first / last `halo` elements of
// the new p also go into the neighbours' landing buffers, the last CTA publishes the sequence number
// (same protocol as cg_update_p_push_kernel).
__global__ void __launch_bounds__(256) pcg_update_p_kernel(long long n, const CGScalars* sc,
                                                           const double* __restrict__ r,
                                                           const double* __restrict__ dinv, double* __restrict__ p,
                                                           const HaloPushArgs h) {
    if (sc->converged) return;
    const double beta = sc->beta;
    const long long tile = 1024;
    const bool push = h.my_xchg != nullptr;
    const long long next_lo = n - h.halo;
    for (long long base = (long long)blockIdx.x * tile; base < n; base += (long long)gridDim.x * tile) {
        double pv[4], rv[4], dv[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const long long i = base + u * 256 + threadIdx.x;
            if (i < n) { pv[u] = __ldcs(p + i); rv[u] = __ldcs(r + i); dv[u] = dinv ? __ldcs(dinv + i) : 1.0; }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const long long i = base + u * 256 + threadIdx.x;
            if (i < n) {
                // dinv == NULL: `r` already holds z = M^-1 r (block-Jacobi)
                const double v = fma(beta, pv[u], dinv ? dv[u] * rv[u] : rv[u]);
                p[i] = v;
                if (push) {
                    if (h.dst_prev != nullptr && i < h.halo) h.dst_prev[i] = v;
                    if (h.dst_next != nullptr && i >= next_lo) h.dst_next[i - next_lo] = v;
                }
            }
        }
    }
    if (!push) return;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t done = atomicAdd(&h.my_xchg->push_count[0], 1u) + 1u;
        if (done == gridDim.x) {
            h.my_xchg->push_count[0] = 0;
            const uint32_t seq = h.my_xchg->halo_seq + 1u;
            h.my_xchg->halo_seq = seq;
            __threadfence_system();
            if (h.dst_prev != nullptr) st_release_sys(h.flag_prev, seq);
            if (h.dst_next != nullptr) st_release_sys(h.flag_next, seq);
        }
    }
}

// ---- block-Jacobi preconditioner with line blocks ------------------------------------------------------
// M = the tridiagonal part of A inside every grid row (W, C, E of the 5-point stencil), i.e. one n x n
// tridiagonal block per grid row, clipped to the rank's band.  z = M^-1 r is one Thomas solve per block.
// One thread per block, marching along the row: neighbouring threads work on neighbouring grid rows, their
// accesses are n elements apart, and the four consecutive elements of a 32-byte sector are consumed by the
// same thread in its next steps (L1 hit).  Set-up (untimed, once per solve): the forward elimination of the
// matrix, m_j = a_j / d'_{j-1}, d'_j = d_j - m_j c_{j-1}, stored as m[], 1 / d'[] and c[].
struct LineBlocks {
    long long row_offset, n_local;
    int n;              // grid side
    long long first_grid_row;  // grid row of block 0
    int n_blocks;
};
__device__ __forceinline__ void line_block_range(const LineBlocks& lb, int b, long long* lo, long long* hi) {
    const long long gr = lb.first_grid_row + b;
    long long s = gr * lb.n, e = s + lb.n;
    if (s < lb.row_offset) s = lb.row_offset;
    if (e > lb.row_offset + lb.n_local) e = lb.row_offset + lb.n_local;
    *lo = s - lb.row_offset;  // local rows [lo, hi)
    *hi = e - lb.row_offset;
}

// coefficient of column `want` in local row lr (CSR: row_ptr != NULL; ELLPACK of `width` otherwise); 0 if absent
__device__ __forceinline__ double row_coeff(const int* row_ptr, const int* col_idx, const double* values, int width,
                                            long long lr, long long want) {
    const long long s = row_ptr ? row_ptr[lr] : lr * width, e = row_ptr ? row_ptr[lr + 1] : (lr + 1) * width;
    double v = 0.0;
    for (long long k = s; k < e; k++)
        if ((unsigned int)col_idx[k] == (unsigned int)want && (row_ptr || col_idx[k] >= 0)) v = values[k];
    return v;
}

__global__ void __launch_bounds__(128) bj_factor_kernel(const LineBlocks lb, const int* __restrict__ row_ptr, int width,
                                                        const int* __restrict__ col_idx, const double* __restrict__ values,
                                                        double* __restrict__ m, double* __restrict__ invd,
                                                        double* __restrict__ c, int* __restrict__ err) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= lb.n_blocks) return;
    long long lo, hi;
    line_block_range(lb, b, &lo, &hi);
    double dprev = 1.0, cprev = 0.0;
    for (long long lr = lo; lr < hi; lr++) {
        const long long r = lb.row_offset + lr;
        const double d = row_coeff(row_ptr, col_idx, values, width, lr, r);
        const double a = lr > lo ? row_coeff(row_ptr, col_idx, values, width, lr, r - 1) : 0.0;
        const double cc = lr + 1 < hi ? row_coeff(row_ptr, col_idx, values, width, lr, r + 1) : 0.0;
        const double mj = lr > lo ? a / dprev : 0.0;
        const double dj = fma(-mj, cprev, d);
        if (dj == 0.0 || d == 0.0) *err = 1;
        m[lr] = mj;
        invd[lr] = 1.0 / dj;
        c[lr] = cc;
        dprev = dj;
        cprev = cc;
    }
}

// z = M^-1 r: forward y_j = r_j - m_j y_{j-1}, backward z_j = (y_j - c_j z_{j+1}) / d'_j  (y kept in z)
__global__ void __launch_bounds__(128) bj_solve_kernel(const LineBlocks lb, const CGScalars* sc,
                                                       const double* __restrict__ m, const double* __restrict__ invd,
                                                       const double* __restrict__ c, const double* __restrict__ r,
                                                       double* __restrict__ z) {
    if (sc != nullptr && sc->converged) return;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= lb.n_blocks) return;
    long long lo, hi;
    line_block_range(lb, b, &lo, &hi);
    double y = 0.0;
    for (long long lr = lo; lr < hi; lr++) {
        y = fma(-m[lr], y, r[lr]);
        z[lr] = y;
    }
    double zn = 0.0;
    for (long long lr = hi - 1; lr >= lo; lr--) {
        zn = fma(-c[lr], zn, z[lr]) * invd[lr];
        z[lr] = zn;
    }
}

// Generic pieces for operators without a fused entry point (any SpmvOperator with run_device):
// partial dot, r = b - Ap with p = r and r.r
__global__ void __launch_bounds__(256) dot_partials_kernel(long long n, const CGScalars* sc,
                                                           const double* __restrict__ x, const double* __restrict__ y,
                                                           const TailArgs tail) {
    __shared__ double scratch[8];
    B200_TAIL_SHARED;
    if (sc != nullptr && sc->converged) return;
    double acc = 0.0;
    const long long tile = 1024;
    for (long long base = (long long)blockIdx.x * tile; base < n; base += (long long)gridDim.x * tile) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const long long i = base + u * 256 + threadIdx.x;
            if (i < n) acc = fma(x[i], y[i], acc);
        }
    }
    acc = block_sum(acc, scratch);
    B200_TAIL_FINISH(tail, acc, 0.0);
}

__global__ void __launch_bounds__(256) residual_init_kernel(long long n, const double* __restrict__ b,
                                                            const double* __restrict__ Ap, double* __restrict__ r,
                                                            double* __restrict__ p, const TailArgs tail) {
    __shared__ double scratch[8];
    B200_TAIL_SHARED;
    double acc = 0.0;
    const long long tile = 1024;
    for (long long base = (long long)blockIdx.x * tile; base < n; base += (long long)gridDim.x * tile) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const long long i = base + u * 256 + threadIdx.x;
            if (i < n) {
                const double rv = b[i] - Ap[i];
                r[i] = rv;
                p[i] = rv;
                acc = fma(rv, rv, acc);
            }
        }
    }
    acc = block_sum(acc, scratch);
    B200_TAIL_FINISH(tail, acc, 0.0);
}

// checksum partials: sum x and sum x^2 (fixed order); replaces the host loops at
// cg_solver.cu:658-665 for vectors that never leave the device at 20k x 20k
__global__ void __launch_bounds__(256) checksum_partials_kernel(long long n, const double* __restrict__ x,
                                                                const TailArgs tail) {
    __shared__ double scratch[8];
    B200_TAIL_SHARED;
    double s = 0.0, q = 0.0;
    const long long tile = 1024;
    for (long long base = (long long)blockIdx.x * tile; base < n; base += (long long)gridDim.x * tile) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const long long i = base + u * 256 + threadIdx.x;
            if (i < n) { const double v = x[i]; s += v; q = fma(v, v, q); }
        }
    }
    s = block_sum(s, scratch);
    __syncthreads();
    q = block_sum(q, scratch);
    B200_TAIL_FINISH(tail, s, q);
}

// Halo push: the first / last `halo` elements of the local vector go straight into the
// neighbours' landing buffers over NVLink (peer stores), followed by a release-store of the epoch
// into the neighbour's flag.  Replaces exchange_halo_mpi (cg_solver_mgpu_partitioned.cu:173-231:
// D2H, MPI_Isend/Irecv, H2D, two stream syncs).  gridDim.x = 2 * ctas_per_dir.

__global__ void __launch_bounds__(256) halo_push_kernel(const HaloPushArgs a) {
    if (a.sc != nullptr && a.sc->converged) return;
    const int per_dir = gridDim.x / 2;
    const int dir = blockIdx.x / per_dir, blk = blockIdx.x % per_dir;
    double* dst = dir == 0 ? a.dst_prev : a.dst_next;
    if (dst != nullptr) {
        const double* src = dir == 0 ? a.v_local : a.v_local + (a.n_local - a.halo);
        for (int i = blk * 256 + threadIdx.x; i < a.halo; i += per_dir * 256) dst[i] = src[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t done = atomicAdd(&a.my_xchg->push_count[0], 1u) + 1u;
        if (done == gridDim.x) {  // last CTA: every rank publishes, also the ones at the ends of the chain
            a.my_xchg->push_count[0] = 0;
            const uint32_t seq = a.my_xchg->halo_seq + 1u;
            a.my_xchg->halo_seq = seq;
            __threadfence_system();
            if (a.dst_prev != nullptr) st_release_sys(a.flag_prev, seq);
            if (a.dst_next != nullptr) st_release_sys(a.flag_next, seq);
        }
    }
}

// K3 with the halo push fused in (multi-GPU): p = r + beta p, and the CTAs that produce the first /
// last `halo` elements store them straight into the neighbours' landing buffers over NVLink.  The
// last CTA to finish (device-scope counter after a system fence) release-stores the arrival epoch
// into both neighbours' flags.  One launch less on the critical path than K3 + halo_push_kernel.
__global__ void __launch_bounds__(256) cg_update_p_push_kernel(long long n, const CGScalars* sc,
                                                               const double* __restrict__ r, double* __restrict__ p,
                                                               const HaloPushArgs h) {
    griddep_wait();
    if (sc->converged) return;
    const double beta = sc->beta;
    const long long tile = 1024;
    const long long next_lo = n - h.halo;
    for (long long base = (long long)blockIdx.x * tile; base < n; base += (long long)gridDim.x * tile) {
        double pv[4], rv[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const long long i = base + u * 256 + threadIdx.x;
            if (i < n) { pv[u] = __ldcs(p + i); rv[u] = __ldcs(r + i); }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const long long i = base + u * 256 + threadIdx.x;
            if (i < n) {
                const double v = fma(beta, pv[u], rv[u]);
                p[i] = v;
                if (h.dst_prev != nullptr && i < h.halo) h.dst_prev[i] = v;
                if (h.dst_next != nullptr && i >= next_lo) h.dst_next[i - next_lo] = v;
            }
        }
    }
    griddep_launch();
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t done = atomicAdd(&h.my_xchg->push_count[0], 1u) + 1u;
        if (done == gridDim.x) {
            h.my_xchg->push_count[0] = 0;
            const uint32_t seq = h.my_xchg->halo_seq + 1u;
            h.my_xchg->halo_seq = seq;
            __threadfence_system();
            if (h.dst_prev != nullptr) st_release_sys(h.flag_prev, seq);
            if (h.dst_next != nullptr) st_release_sys(h.flag_next, seq);
        }
    }
}

}  // namespace b200
