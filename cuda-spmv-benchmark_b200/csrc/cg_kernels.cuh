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
};

// host-visible mirror (pinned, mapped), written by the reduce kernel after every r.r
struct CGStatus {
    volatile int iterations;
    volatile int converged;
    volatile double residual;
    volatile double b_norm;
    volatile int error;  // flag-wait timeout in a peer exchange
};

// exchange area of one rank, in that rank's device memory, written by its peers over NVLink
struct XchgArea {
    uint64_t slots[2][B200_MAX_RANKS][2];  // LL words: (epoch << 32 | 32 data bits) x 2 per double
    uint32_t halo_flag_prev;               // epoch of the last halo written by rank-1
    uint32_t halo_flag_next;               // epoch of the last halo written by rank+1
    uint32_t push_count[2];                // local: CTAs done with the halo push (prev, next)
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
    uint32_t epoch;
    uint32_t* push_count;  // my_xchg->push_count
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
    uint32_t epoch;
    CGScalars* sc;
    int beta_zero;
};

enum { RED_RR0 = 0, RED_PAP = 1, RED_RR = 2, RED_SUM = 3, RED_RZ0 = 4, RED_PCG = 5 };  // 4, 5: Jacobi PCG
enum { RED_PUSH = 1, RED_COMBINE = 2 };

struct ReduceArgs {
    const double* partials;
    const double* partials_b;  // RED_PCG: the r.z partials (partials = r.r), same count; single rank only
    int n_partials;
    int which;   // RED_*
    int phases;  // RED_PUSH | RED_COMBINE
    double tol;
    CGScalars* sc;
    CGStatus* status;  // may be NULL
    double* out;       // RED_SUM: result
    // multi-rank
    int rank, world;
    uint32_t epoch;
    XchgArea* my_xchg;
    XchgArea* peer_xchg[B200_MAX_RANKS];
    double* stash;  // local: my partial sum between PUSH and COMBINE launches
    // RED_RR, multi-GPU deferred-x schedule: once beta is known the same CTA advances the halo copies
    // of the direction, p_halo = fma(beta, p_halo_old, r_halo) (one launch less per iteration)
    int with_halo_dir;
    HaloDirArgs hd;
};

// One CTA of 1024 threads.  Fixed-order final sum of the per-CTA partials, optional rank
// exchange (LL protocol: data and epoch travel in the same 8-byte store, so no fence), then the
// scalar recurrences that the reference spreads over scalar_divide_kernel /
// check_convergence_kernel / a D2D copy (cg_solver.cu:414-431,560,594-637).
__global__ void __launch_bounds__(1024) cg_reduce_kernel(const ReduceArgs a) {
    __shared__ double scratch[32];
    __shared__ double rank_sum[B200_MAX_RANKS];
    __shared__ double local_total;
    CGScalars* sc = a.sc;
    if (a.which != RED_SUM && sc->converged) return;

    const int t = threadIdx.x;
    __shared__ double local_total_b;
    if (a.phases & RED_PUSH) {
        double v = 0.0;
        for (int i = t; i < a.n_partials; i += 1024) v += a.partials[i];
        v = block_sum(v, scratch);
        if (t == 0) { local_total = v; if (a.stash) *a.stash = v; }
        __syncthreads();
        if (a.which == RED_PCG) {
            double w = 0.0;
            for (int i = t; i < a.n_partials; i += 1024) w += a.partials_b[i];
            w = block_sum(w, scratch);
            if (t == 0) local_total_b = w;
            __syncthreads();
        }
        if (a.world > 1 && t < a.world) {
            const uint64_t bits = (uint64_t)__double_as_longlong(local_total);
            const uint64_t tag = (uint64_t)a.epoch << 32;
            uint64_t* dst = a.peer_xchg[t]->slots[a.epoch & 1][a.rank];
            st_volatile_u64(dst, tag | (bits & 0xffffffffull));
            st_volatile_u64(dst + 1, tag | (bits >> 32));
        }
    } else if (t == 0) {
        local_total = *a.stash;
    }
    if (!(a.phases & RED_COMBINE)) return;
    __syncthreads();

    double total = local_total;
    if (a.world > 1) {
        if (t < a.world) {
            const uint64_t* src = a.my_xchg->slots[a.epoch & 1][t];
            const uint64_t t0 = globaltimer_ns();
            const bool dead = (sc != nullptr && sc->error != 0);
            uint64_t w0, w1;
            for (;;) {
                w0 = ld_volatile_u64(src);
                w1 = ld_volatile_u64(src + 1);
                if ((uint32_t)(w0 >> 32) == a.epoch && (uint32_t)(w1 >> 32) == a.epoch) break;
                if (dead || globaltimer_ns() - t0 > 8000000000ull) {  // 8 s: report, never hang
                    if (sc) sc->error = 1;
                    if (a.status) a.status->error = 1;
                    break;
                }
            }
            rank_sum[t] = __longlong_as_double((long long)((w1 << 32) | (w0 & 0xffffffffull)));
        }
        __syncthreads();
        if (t == 0) {
            total = 0.0;
            for (int r = 0; r < a.world; r++) total += rank_sum[r];  // rank order: same on every GPU
        }
    }
    __shared__ int s_go_halo;
    if (t == 0) {
        s_go_halo = 0;
        if (a.which == RED_SUM) {
            *a.out = total;
        } else if (a.which == RED_RR0) {
            sc->rr_old = total;
            sc->b_norm = sqrt(total);
            sc->residual = sc->b_norm;
            if (a.status) { a.status->b_norm = sc->b_norm; a.status->residual = sc->b_norm; }
        } else if (a.which == RED_PAP) {
            sc->pAp = total;
            sc->alpha = sc->rr_old / total;  // PCG keeps rho = r.z in rr_old
        } else if (a.which == RED_RZ0) {
            sc->rr_old = total;  // rho_0 = r0.z0 (b_norm was set by RED_RR0)
        } else if (a.which == RED_PCG) {
            sc->rr_new = total;
            const double res = sqrt(total);
            sc->residual = res;
            const int it = sc->iterations + 1;
            sc->iterations = it;
            const int conv = (res / sc->b_norm < a.tol) ? 1 : 0;
            if (conv) {
                sc->converged = 1;
            } else {
                sc->beta = local_total_b / sc->rr_old;
                sc->rr_old = local_total_b;
            }
            if (a.status) {
                a.status->residual = res;
                a.status->iterations = it;
                a.status->converged = conv;
            }
        } else {  // RED_RR
            sc->rr_new = total;
            const double res = sqrt(total);
            sc->residual = res;
            const int it = sc->iterations + 1;
            sc->iterations = it;
            const int conv = (res / sc->b_norm < a.tol) ? 1 : 0;
            if (conv) {
                sc->converged = 1;
            } else {
                sc->beta = total / sc->rr_old;
                sc->rr_old = total;
                s_go_halo = a.with_halo_dir;
            }
            if (a.status) {
                // the host reads these only after synchronising on an event recorded behind this
                // kernel (cg_engine.cpp), so no system-scope fence is needed here
                a.status->residual = res;
                a.status->iterations = it;
                a.status->converged = conv;
            }
        }
    }
    if (!a.with_halo_dir) return;
    __syncthreads();
    if (!s_go_halo) return;
    // ---- halo copies of the new direction (the neighbours' r edges arrived before their LL words)
    const HaloDirArgs& h = a.hd;
    if (t == 0) {
        const uint64_t t0 = globaltimer_ns();
        const uint32_t* fl[2] = {h.r_prev ? h.flag_prev : nullptr, h.r_next ? h.flag_next : nullptr};
        for (int d = 0; d < 2; d++) {
            if (!fl[d]) continue;
            while ((int32_t)(ld_acquire_sys(fl[d]) - h.epoch) < 0) {
                if (sc->error || globaltimer_ns() - t0 > 8000000000ull) { sc->error = 1; break; }
                __nanosleep(64);
            }
        }
    }
    __syncthreads();
    const double beta = sc->beta;
    for (int i = t; i < h.halo; i += 1024) {
        if (h.r_prev) h.pnew_prev[i] = fma(beta, h.pold_prev[i], __ldcg(h.r_prev + i));
        if (h.r_next) h.pnew_next[i] = fma(beta, h.pold_next[i], __ldcg(h.r_next + i));
    }
}

// K2: x += alpha p ; r -= alpha Ap ; partials[cta] = sum r^2 over the CTA's tiles.
// reference: axpy_kernel_device + axpy_sub_kernel_device + dot_kernel (cg_solver.cu:59-78,110-132)
// contracted the same way (fma(alpha,p,x), fma(-alpha,Ap,r)).
template <int VEC>
__global__ void __launch_bounds__(256) cg_update_xr_kernel(long long n, const CGScalars* __restrict__ sc,
                                                           const double* __restrict__ p,
                                                           const double* __restrict__ Ap, double* __restrict__ x,
                                                           double* __restrict__ r, double* __restrict__ partials) {
    __shared__ double scratch[8];
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
    acc = block_sum(acc, scratch);
    if (threadIdx.x == 0) partials[blockIdx.x] = acc;
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
__global__ void __launch_bounds__(256) cg_update_r_kernel(long long n, const CGScalars* __restrict__ sc,
                                                          const double* __restrict__ Ap, double* __restrict__ r,
                                                          double* __restrict__ partials, const HaloPushArgs h) {
    __shared__ double scratch[8];
    if (sc->converged) return;
    const double nalpha = -sc->alpha;
    constexpr int UNROLL = 4;
    const long long tile = 256LL * VEC * UNROLL;
    const long long next_lo = n - h.halo;
    auto push = [&](long long i, double v) {
        if (PUSH) {
            if (h.dst_prev != nullptr && i < h.halo) h.dst_prev[i] = v;
            if (h.dst_next != nullptr && i >= next_lo) h.dst_next[i - next_lo] = v;
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
    if (PUSH) __threadfence_system();
    acc = block_sum(acc, scratch);  // contains a __syncthreads
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = acc;
        if (PUSH) {
            const uint32_t done = atomicAdd(&h.push_count[0], 1u) + 1u;
            if (done == gridDim.x) {
                h.push_count[0] = 0;
                __threadfence_system();
                if (h.dst_prev != nullptr) st_release_sys(h.flag_prev, h.epoch);
                if (h.dst_next != nullptr) st_release_sys(h.flag_next, h.epoch);
            }
        }
    }
}

// Halo part of the direction update (multi-GPU, deferred-x schedule): the neighbours pushed the edges
// of their NEW r; the halo copies of p follow the same recurrence as the local part,
// p_halo = fma(beta, p_halo_old, r_halo), kept in two local ping-pong buffers.  beta_zero: first
// direction (p0 = r0).  One small CTA group; waits (bounded) for the arrival epochs first.

__global__ void __launch_bounds__(256) cg_halo_dir_kernel(const HaloDirArgs a) {
    if (a.sc->converged) return;
    if (threadIdx.x == 0) {
        const uint64_t t0 = globaltimer_ns();
        const uint32_t* fl[2] = {a.r_prev ? a.flag_prev : nullptr, a.r_next ? a.flag_next : nullptr};
        for (int d = 0; d < 2; d++) {
            if (!fl[d]) continue;
            while ((int32_t)(ld_acquire_sys(fl[d]) - a.epoch) < 0) {
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

// After the last iteration of the deferred-x schedule x still lacks alpha_last * p_last.
// p_last lives in p0 or p1 depending on the parity of the completed iterations (known on the device).
__global__ void __launch_bounds__(256) cg_finish_x_kernel(long long n, const CGScalars* __restrict__ sc,
                                                          const double* __restrict__ p0,
                                                          const double* __restrict__ p1, double* __restrict__ x) {
    const int it = sc->iterations;
    if (it <= 0) return;
    const double alpha = sc->alpha;
    const double* __restrict__ p = ((it - 1) & 1) ? p1 : p0;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256)
        x[i] = fma(alpha, p[i], x[i]);
}

// K3: p = r + beta p   (reference update_p_kernel, cg_solver.cu:91-96: fma(beta,p,r))
template <int VEC>
__global__ void __launch_bounds__(256) cg_update_p_kernel(long long n, const CGScalars* __restrict__ sc,
                                                          const double* __restrict__ r, double* __restrict__ p) {
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
__global__ void __launch_bounds__(256) pcg_init_kernel(long long n, const double* __restrict__ r,
                                                       const double* __restrict__ dinv, double* __restrict__ p,
                                                       double* __restrict__ partials) {
    __shared__ double scratch[8];
    double acc = 0.0;
    const long long tile = 1024;
    for (long long base = (long long)blockIdx.x * tile; base < n; base += (long long)gridDim.x * tile) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const long long i = base + u * 256 + threadIdx.x;
            if (i < n) {
                const double rv = r[i], z = dinv[i] * rv;
                p[i] = z;
                acc = fma(rv, z, acc);
            }
        }
    }
    acc = block_sum(acc, scratch);
    if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}

// K2p: x += alpha p ; r -= alpha Ap ; partials of r.r (convergence) and r.z (rho), z = dinv r
__global__ void __launch_bounds__(256) pcg_update_xr_kernel(long long n, const CGScalars* __restrict__ sc,
                                                            const double* __restrict__ p,
                                                            const double* __restrict__ Ap,
                                                            const double* __restrict__ dinv, double* __restrict__ x,
                                                            double* __restrict__ r, double* __restrict__ partials_rr,
                                                            double* __restrict__ partials_rz) {
    __shared__ double scratch[8];
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
    if (threadIdx.x == 0) { partials_rr[blockIdx.x] = arr; partials_rz[blockIdx.x] = arz; }
}

// K3p: p = z + beta p, z = dinv r
__global__ void __launch_bounds__(256) pcg_update_p_kernel(long long n, const CGScalars* __restrict__ sc,
                                                           const double* __restrict__ r,
                                                           const double* __restrict__ dinv, double* __restrict__ p) {
    if (sc->converged) return;
    const double beta = sc->beta;
    const long long tile = 1024;
    for (long long base = (long long)blockIdx.x * tile; base < n; base += (long long)gridDim.x * tile) {
        double pv[4], rv[4], dv[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const long long i = base + u * 256 + threadIdx.x;
            if (i < n) { pv[u] = __ldcs(p + i); rv[u] = __ldcs(r + i); dv[u] = __ldcs(dinv + i); }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const long long i = base + u * 256 + threadIdx.x;
            if (i < n) p[i] = fma(beta, pv[u], dv[u] * rv[u]);
        }
    }
}

// Generic pieces for operators without a fused entry point (any SpmvOperator with run_device):
// partial dot, r = b - Ap with p = r and r.r
__global__ void __launch_bounds__(256) dot_partials_kernel(long long n, const CGScalars* __restrict__ sc,
                                                           const double* __restrict__ x, const double* __restrict__ y,
                                                           double* __restrict__ partials) {
    __shared__ double scratch[8];
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
    if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}

__global__ void __launch_bounds__(256) residual_init_kernel(long long n, const double* __restrict__ b,
                                                            const double* __restrict__ Ap, double* __restrict__ r,
                                                            double* __restrict__ p, double* __restrict__ partials) {
    __shared__ double scratch[8];
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
    if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}

// checksum partials: sum x and sum x^2 (fixed order); replaces the host loops at
// cg_solver.cu:658-665 for vectors that never leave the device at 20k x 20k
__global__ void __launch_bounds__(256) checksum_partials_kernel(long long n, const double* __restrict__ x,
                                                                double* __restrict__ psum, double* __restrict__ psq) {
    __shared__ double scratch[8];
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
    if (threadIdx.x == 0) { psum[blockIdx.x] = s; psq[blockIdx.x] = q; }
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
    if (dst == nullptr) return;
    const double* src = dir == 0 ? a.v_local : a.v_local + (a.n_local - a.halo);
    for (int i = blk * 256 + threadIdx.x; i < a.halo; i += per_dir * 256) dst[i] = src[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t done = atomicAdd(&a.push_count[dir], 1u) + 1u;
        if (done == (uint32_t)per_dir) {
            a.push_count[dir] = 0;
            __threadfence_system();
            st_release_sys(dir == 0 ? a.flag_prev : a.flag_next, a.epoch);
        }
    }
}

// K3 with the halo push fused in (multi-GPU): p = r + beta p, and the CTAs that produce the first /
// last `halo` elements store them straight into the neighbours' landing buffers over NVLink.  The
// last CTA to finish (device-scope counter after a system fence) release-stores the arrival epoch
// into both neighbours' flags.  One launch less on the critical path than K3 + halo_push_kernel.
__global__ void __launch_bounds__(256) cg_update_p_push_kernel(long long n, const CGScalars* __restrict__ sc,
                                                               const double* __restrict__ r, double* __restrict__ p,
                                                               const HaloPushArgs h) {
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
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t done = atomicAdd(&h.push_count[0], 1u) + 1u;
        if (done == gridDim.x) {
            h.push_count[0] = 0;
            __threadfence_system();
            if (h.dst_prev != nullptr) st_release_sys(h.flag_prev, h.epoch);
            if (h.dst_next != nullptr) st_release_sys(h.flag_next, h.epoch);
        }
    }
}

}  // namespace b200
