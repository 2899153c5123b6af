// cg_engine.cpp -- Conjugate Gradient behind cg_solve / cg_solve_device
// (reference src/solvers/cg_solver.cu:154-706) and cg_solve_mgpu_partitioned
// (reference src/solvers/cg_solver_mgpu_partitioned.cu:236-908), built on the fused kernels.
//
// One engine serves both: a solve runs over `world` row bands ("ranks").  This process drives
// the ranks it owns -- all of them (single process over several GPUs, or several virtual ranks on
// one GPU for tests) or exactly one (one process per GPU under torchrun, peers mapped through
// CUDA IPC).  The reference's MPI + pinned-host staging is replaced by peer-memory stores over
// NVLink with epoch flags: the halo push writes straight into the neighbour's landing buffer,
// the scalar all-reduce is an 8-byte-store exchange summed in rank order inside the reduce
// kernel (include/b200_kernels.h: b200_halo_push, b200_cg_reduce).
//
// Launch order is phase-major (all local ranks do phase k before any does phase k+1), so a kernel
// that waits on a peer flag is always enqueued after the kernel that sets it: with virtual ranks on
// one stream the waits are satisfied on arrival, with real GPUs they overlap.
//
// The host never blocks inside the iteration loop: the convergence test lives in the reduce
// kernel, later kernels turn into no-ops once it fires, and the host polls a pinned status word
// kLag iterations behind (the reference does a blocking 4-byte D2H every iteration,
// cg_solver.cu:598-599).
#include <math.h>

#include <algorithm>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "host_common.h"

using namespace b200host;

namespace {

constexpr int kMaxRanks = 16;
constexpr int kLag = 3;

// NVTX ranges under the reference's names (cg_solver_mgpu_partitioned.cu:540-717) so that existing
// nsys views keep working.  They bracket the host-side ENQUEUE of a phase (the host never waits).
struct Nvtx {
    explicit Nvtx(const char* name) { nvtxRangePushA(name); }
    ~Nvtx() { nvtxRangePop(); }
};

struct HostStatus {  // mirrors b200::CGStatus (csrc/cg_kernels.cuh)
    volatile int iterations;
    volatile int converged;
    volatile double residual;
    volatile double b_norm;
    volatile int error;
};

// ------------------------------------------------------------------------------------------------
// exchange context (who are the ranks, where are their landing buffers)
// ------------------------------------------------------------------------------------------------
struct Mgpu {
    bool inited = false;
    int world = 1;
    int nlocal = 1;
    int local_rank[kMaxRanks];
    int local_dev[kMaxRanks];
    void* xchg[kMaxRanks];   // every rank's exchange block, as mapped in this process
    bool opened[kMaxRanks];  // mapped through cudaIpcOpenMemHandle
    size_t halo_cap = 0;     // doubles per landing buffer
    size_t area_bytes = 0;   // offset of the first landing buffer inside a block
    uint32_t halo_epoch = 0, red_epoch = 0;
    bool single_device = true;  // all local ranks on one device: split reduce launches
} g;

// block = exchange area | landing_prev | landing_next (written by the neighbours) | 4 local buffers:
// halo copies of the search direction, [parity][prev / next] (deferred-x schedule)
size_t block_bytes() { return g.area_bytes + 6 * g.halo_cap * sizeof(double); }
double* landing_prev(int r) { return reinterpret_cast<double*>(static_cast<char*>(g.xchg[r]) + g.area_bytes); }
double* landing_next(int r) { return landing_prev(r) + g.halo_cap; }
double* halo_dir(int r, int parity, int next) { return landing_prev(r) + (2 + 2 * (parity & 1) + (next ? 1 : 0)) * g.halo_cap; }
uint32_t* flag_prev(int r) { return reinterpret_cast<uint32_t*>(static_cast<char*>(g.xchg[r]) + b200_xchg_flag_prev_offset()); }
uint32_t* flag_next(int r) { return reinterpret_cast<uint32_t*>(static_cast<char*>(g.xchg[r]) + b200_xchg_flag_next_offset()); }

void mgpu_reset() {
    if (!g.inited) return;
    for (int r = 0; r < g.world; r++) {
        if (!g.xchg[r]) continue;
        if (g.opened[r]) cudaIpcCloseMemHandle(g.xchg[r]);
    }
    for (int l = 0; l < g.nlocal; l++) {
        cudaSetDevice(g.local_dev[l]);
        cudaFree(g.xchg[g.local_rank[l]]);
    }
    g = Mgpu();
}

int alloc_block(int dev, void** out) {
    B200_CUDA(cudaSetDevice(dev));
    B200_CUDA(cudaMalloc(out, block_bytes()));
    B200_CUDA(cudaMemset(*out, 0, block_bytes()));
    return 0;
}

}  // namespace

namespace { extern int g_schedule; }
extern "C" void b200_cg_set_schedule(int deferred_x) { g_schedule = deferred_x ? 1 : 0; }
extern "C" int b200_mgpu_world(void) { return g.inited ? g.world : 1; }
extern "C" int b200_mgpu_rank(void) { return (g.inited && g.nlocal == 1) ? g.local_rank[0] : 0; }
extern "C" void b200_mgpu_finalize(void) { mgpu_reset(); }

// all ranks driven by this process; devices[r] may repeat (virtual ranks on one GPU)
extern "C" int b200_mgpu_init_single_process(int world, const int* devices, int max_grid) {
    if (world < 1 || world > kMaxRanks || max_grid < 1) return 1;
    mgpu_reset();
    g.world = world; g.nlocal = world;
    g.halo_cap = (size_t)max_grid;
    g.area_bytes = (b200_xchg_bytes() + 255) & ~(size_t)255;
    g.single_device = true;
    for (int r = 0; r < world; r++) {
        g.local_rank[r] = r;
        g.local_dev[r] = devices ? devices[r] : r;
        if (g.local_dev[r] != g.local_dev[0]) g.single_device = false;
        g.opened[r] = false;
        if (alloc_block(g.local_dev[r], &g.xchg[r])) return 2;
    }
    for (int a = 0; a < world; a++)  // peer access between distinct devices (NVLink / NVSwitch)
        for (int b = 0; b < world; b++) {
            if (g.local_dev[a] == g.local_dev[b]) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, g.local_dev[a], g.local_dev[b]);
            if (!can) { fprintf(stderr, "[b200] GPU %d cannot access GPU %d\n", g.local_dev[a], g.local_dev[b]); return 3; }
            cudaSetDevice(g.local_dev[a]);
            cudaError_t e = cudaDeviceEnablePeerAccess(g.local_dev[b], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return 3;
            cudaGetLastError();
        }
    cudaSetDevice(g.local_dev[0]);
    g.inited = true;
    return 0;
}

// one rank per process: allocate my block and hand out its IPC handle (64 bytes) ...
extern "C" int b200_mgpu_init_rank(int rank, int world, int device, int max_grid, void* handle_out64) {
    if (world < 1 || world > kMaxRanks || rank < 0 || rank >= world || max_grid < 1 || !handle_out64) return 1;
    mgpu_reset();
    g.world = world; g.nlocal = 1;
    g.local_rank[0] = rank; g.local_dev[0] = device;
    g.halo_cap = (size_t)max_grid;
    g.area_bytes = (b200_xchg_bytes() + 255) & ~(size_t)255;
    for (int r = 0; r < world; r++) { g.xchg[r] = nullptr; g.opened[r] = false; }
    if (alloc_block(device, &g.xchg[rank])) return 2;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    B200_CUDA(cudaIpcGetMemHandle(&h, g.xchg[rank]));
    memcpy(handle_out64, &h, 64);
    g.single_device = true;
    return 0;
}

// ... and map everybody else's once the caller has all-gathered the handles (world x 64 bytes)
extern "C" int b200_mgpu_connect(const void* handles) {
    if (g.nlocal != 1 || !handles) return 1;
    B200_CUDA(cudaSetDevice(g.local_dev[0]));
    for (int r = 0; r < g.world; r++) {
        if (r == g.local_rank[0]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char*>(handles) + 64 * r, 64);
        B200_CUDA(cudaIpcOpenMemHandle(&g.xchg[r], h, cudaIpcMemLazyEnablePeerAccess));
        g.opened[r] = true;
    }
    g.inited = true;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// per-rank solver workspace (cached between solves of the same shape)
// ------------------------------------------------------------------------------------------------
namespace {

struct RankWs {
    int rank = 0, dev = 0;
    long long off = 0, nl = 0;
    cudaStream_t st = nullptr;
    bool own_stream = false;
    DeviceBand band;
    bool own_band = false;
    double *x = nullptr, *r = nullptr, *p = nullptr, *p2 = nullptr, *Ap = nullptr, *b = nullptr;
    double *partials = nullptr, *partials2 = nullptr, *stash = nullptr, *sums = nullptr;
    double* dinv = nullptr;  // Jacobi PCG: 1 / diag(A), allocated on first use
    int* dinv_err = nullptr;
    void* scalars = nullptr;
    HostStatus* status = nullptr;  // pinned + mapped
    void* status_dev = nullptr;
    void** peer_table_host = nullptr;
    int max_partials = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<cudaEvent_t> iter_ev;
    std::vector<cudaEvent_t> phase_ev;  // detailed timers

    void free_vectors() {
        cudaFree(x); cudaFree(r); cudaFree(p); cudaFree(p2); cudaFree(Ap); cudaFree(b);
        cudaFree(partials); cudaFree(partials2); cudaFree(stash); cudaFree(sums); cudaFree(scalars);
        cudaFree(dinv); cudaFree(dinv_err);
        dinv = nullptr; dinv_err = nullptr;
        x = r = p = p2 = Ap = b = partials = partials2 = stash = sums = nullptr;
        scalars = nullptr;
        if (status) cudaFreeHost((void*)status);
        status = nullptr;
        for (auto e : iter_ev) cudaEventDestroy(e);
        for (auto e : phase_ev) cudaEventDestroy(e);
        iter_ev.clear(); phase_ev.clear();
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        ev0 = ev1 = nullptr;
        if (own_stream && st) cudaStreamDestroy(st);
        st = nullptr;
        if (own_band) band.release();
        nl = 0;
    }
};

struct Workspace {
    std::vector<RankWs> ranks;
    long long N = 0;
    int grid = 0, world = 0;
    const void* matrix_key = nullptr;  // entries pointer or operator pointer
    int matrix_nnz = 0;
    bool synthetic = false;
    void release() {
        for (auto& w : ranks) { cudaSetDevice(w.dev); w.free_vectors(); }
        ranks.clear();
        N = 0;
    }
} ws;

int alloc_rank_vectors(RankWs& w, int max_partials) {
    B200_CUDA(cudaSetDevice(w.dev));
    const size_t vb = (size_t)w.nl * sizeof(double);
    B200_CUDA(cudaMalloc(&w.x, vb));
    B200_CUDA(cudaMalloc(&w.r, vb));
    B200_CUDA(cudaMalloc(&w.p, vb));
    B200_CUDA(cudaMalloc(&w.p2, vb));  // second direction buffer (deferred-x schedule)
    B200_CUDA(cudaMalloc(&w.Ap, vb));
    B200_CUDA(cudaMalloc(&w.b, vb));
    w.max_partials = max_partials;
    B200_CUDA(cudaMalloc(&w.partials, (size_t)max_partials * sizeof(double)));
    B200_CUDA(cudaMalloc(&w.partials2, (size_t)max_partials * sizeof(double)));
    B200_CUDA(cudaMalloc(&w.stash, 4 * sizeof(double)));
    B200_CUDA(cudaMalloc(&w.sums, 4 * sizeof(double)));
    B200_CUDA(cudaMalloc(&w.scalars, b200_cg_scalars_bytes()));
    B200_CUDA(cudaHostAlloc((void**)&w.status, sizeof(HostStatus), cudaHostAllocMapped | cudaHostAllocPortable));
    B200_CUDA(cudaHostGetDevicePointer(&w.status_dev, (void*)w.status, 0));
    B200_CUDA(cudaEventCreate(&w.ev0));
    B200_CUDA(cudaEventCreate(&w.ev1));
    return 0;
}

cudaEvent_t iter_event(RankWs& w, int k) {
    while ((int)w.iter_ev.size() <= k) {
        cudaEvent_t e;
        cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        w.iter_ev.push_back(e);
    }
    return w.iter_ev[k];
}

struct PhaseTimer {  // detailed timers on the first local rank only (enable_detailed_timers)
    RankWs* w = nullptr;
    bool on = false;
    size_t used = 0;
    std::vector<int> tags;
    void mark(int tag) {
        if (!on) return;
        if (used == w->phase_ev.size()) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            w->phase_ev.push_back(e);
        }
        cudaEventRecord(w->phase_ev[used++], w->st);
        tags.push_back(tag);
    }
};
enum { T_BEGIN = 0, T_SPMV, T_RED_PAP, T_XR, T_RED_RR, T_P, T_HALO, T_INIT_R, T_RED_RR0, T_N };

struct SolveOut {
    int iterations = 0, converged = 0;
    double residual = 0, b_norm = 0, total_ms = 0, sum = 0, norm = 0;
    double phase_ms[T_N] = {0};
    int phase_cnt[T_N] = {0};
};
SolveOut g_last;  // per-phase event times of the most recent solve (enable_detailed_timers)

// band descriptor of rank w with its halo wiring for this epoch
void wire_band(const RankWs& w, b200_band* b, bool halos, uint32_t epoch) {
    w.band.describe(b);
    if (halos && g.world > 1) {
        if (w.rank > 0) { b->d_halo_prev = landing_prev(w.rank); b->d_flag_prev = flag_prev(w.rank); }
        if (w.rank < g.world - 1) { b->d_halo_next = landing_next(w.rank); b->d_flag_next = flag_next(w.rank); }
        b->epoch = epoch;
    }
}

// band descriptor whose halos are the local direction copies of the given parity (no flags: the
// halo-direction kernel has already waited for the neighbours)
void wire_band_dir(const RankWs& w, b200_band* b, int parity) {
    w.band.describe(b);
    if (g.world > 1) {
        if (w.rank > 0) b->d_halo_prev = halo_dir(w.rank, parity, 0);
        if (w.rank < g.world - 1) b->d_halo_next = halo_dir(w.rank, parity, 1);
    }
}

// deferred-x schedule (4 launches, 112 B/row per iteration) unless B200_CG_SCHEDULE=classic or
// b200_cg_set_schedule(0)
int g_schedule = -1;
bool schedule_deferred_x() {
    if (g_schedule < 0) {
        const char* e = getenv("B200_CG_SCHEDULE");
        g_schedule = (e && strcmp(e, "classic") == 0) ? 0 : 1;
    }
    return g_schedule == 1;
}

int push_halo(const RankWs& w, const double* v, uint32_t epoch, const void* scalars) {
    const int n = ws.grid;
    double* dprev = w.rank > 0 ? landing_next(w.rank - 1) : nullptr;  // I am the "next" neighbour of rank-1
    double* dnext = w.rank < g.world - 1 ? landing_prev(w.rank + 1) : nullptr;
    uint32_t* fprev = w.rank > 0 ? flag_next(w.rank - 1) : nullptr;
    uint32_t* fnext = w.rank < g.world - 1 ? flag_prev(w.rank + 1) : nullptr;
    return b200_halo_push(v, w.nl, n, dprev, dnext, fprev, fnext, epoch, g.xchg[w.rank], scalars, w.st);
}

}  // namespace

namespace {

struct Engine {
    bool fused;           // band kernels available (stencil operators / mgpu); else op->run_device
    bool pcg = false;     // Jacobi-preconditioned CG (single GPU, classic launch grouping)
    SpmvOperator* op;     // generic path
    int n_partials_spmv[kMaxRanks];

    // dir_it >= 0 (multi-GPU deferred-x schedule, which == 2): the reduce CTA also advances the halo
    // copies of the direction from parity dir_it to dir_it + 1
    int for_ranks_reduce(int which, double tol, const int* n_partials, bool second_buf, uint32_t epoch, int dir_it = -1) {
        const int phases_list_fused[1] = {3};
        const int phases_list_split[2] = {1, 2};
        const bool split = (g.world > 1 && g.single_device && ws.ranks.size() > 1);
        const int* pl = split ? phases_list_split : phases_list_fused;
        const int np = split ? 2 : 1;
        for (int pi = 0; pi < np; pi++) {
            for (size_t l = 0; l < ws.ranks.size(); l++) {
                RankWs& w = ws.ranks[l];
                B200_CUDA(cudaSetDevice(w.dev));
                if (dir_it >= 0 && which == 2 && g.world > 1) {
                    B200_K(b200_cg_reduce_rr_dir(second_buf ? w.partials2 : w.partials, n_partials[l], pl[pi], tol, w.scalars,
                                                 w.status_dev, w.rank, g.world, epoch, g.xchg, w.stash,
                                                 w.rank > 0 ? landing_prev(w.rank) : nullptr,
                                                 w.rank < g.world - 1 ? landing_next(w.rank) : nullptr,
                                                 halo_dir(w.rank, dir_it, 0), halo_dir(w.rank, dir_it, 1),
                                                 halo_dir(w.rank, dir_it + 1, 0), halo_dir(w.rank, dir_it + 1, 1), ws.grid,
                                                 flag_prev(w.rank), flag_next(w.rank), g.halo_epoch, w.st));
                    continue;
                }
                B200_K(b200_cg_reduce(second_buf ? w.partials2 : w.partials, n_partials[l], which, pl[pi], tol,
                                      w.scalars, which == 3 ? nullptr : w.status_dev, w.sums, w.rank, g.world, epoch,
                                      g.world > 1 ? g.xchg : nullptr, w.stash, w.st));
            }
        }
        return 0;
    }

    int solve(const double* b_host, double* x_host, int max_iters, double tol, int verbose, int timers,
              const char* tag, SolveOut* out) {
        const size_t L = ws.ranks.size();
        const bool multi = g.world > 1;
        // ---- untimed: upload b and the initial guess (reference cg_solver.cu:473-474)
        for (auto& w : ws.ranks) {
            B200_CUDA(cudaSetDevice(w.dev));
            B200_CUDA(cudaMemsetAsync(w.scalars, 0, b200_cg_scalars_bytes(), w.st));
            memset((void*)w.status, 0, sizeof(HostStatus));
            B200_CUDA(cudaMemcpyAsync(w.b, b_host + w.off, (size_t)w.nl * sizeof(double), cudaMemcpyHostToDevice, w.st));
            B200_CUDA(cudaMemcpyAsync(w.x, x_host + w.off, (size_t)w.nl * sizeof(double), cudaMemcpyHostToDevice, w.st));
        }
        if (pcg) {
            // untimed set-up like the uploads: dinv = 1 / diag(A), rebuilt for every solve (the matrix
            // behind an operator may have changed while the shape stayed the same)
            if (multi) { fprintf(stderr, "[ERROR] Jacobi PCG is single-GPU\n"); return 1; }
            RankWs& w = ws.ranks[0];
            int ell_width = 0;
            const DeviceBand* m = fused ? &w.band : operator_matrix(op, &ell_width);
            if (!m) { fprintf(stderr, "[ERROR] Jacobi PCG needs one of this library's operators (diagonal access)\n"); return 1; }
            if (m->layout == 1 && ell_width == 0) ell_width = 5;  // stencil ELLPACK band
            if (!w.dinv) {
                B200_CUDA(cudaMalloc(&w.dinv, (size_t)w.nl * sizeof(double)));
                B200_CUDA(cudaMalloc(&w.dinv_err, sizeof(int)));
            }
            B200_CUDA(cudaMemsetAsync(w.dinv_err, 0, sizeof(int), w.st));
            B200_K(b200_pcg_diag_inv(m->layout == 0 ? m->d_row_ptr : nullptr, m->d_col_idx, m->d_values, w.nl, w.off, ell_width,
                                     w.dinv, w.dinv_err, w.st));
            int bad = 0;
            B200_CUDA(cudaMemcpyAsync(&bad, w.dinv_err, sizeof(int), cudaMemcpyDeviceToHost, w.st));
            B200_CUDA(cudaStreamSynchronize(w.st));
            if (bad) { fprintf(stderr, "[ERROR] Jacobi PCG: a row has no (or a zero) diagonal entry\n"); return 1; }
        }
        for (auto& w : ws.ranks) { B200_CUDA(cudaSetDevice(w.dev)); B200_CUDA(cudaStreamSynchronize(w.st)); }
        if (multi) {
            // Align the ranks before the clock starts (the reference does MPI_Barrier right before its
            // start event, cg_solver_mgpu_partitioned.cu:405-413): an empty rank-exchange reduction is
            // a device-side barrier over peer memory.  Without it a rank whose upload finished early
            // would count its neighbours' PCIe time as solver time.
            int zero[kMaxRanks] = {0};
            if (for_ranks_reduce(3, tol, zero, false, ++g.red_epoch)) return 1;
            for (auto& w : ws.ranks) { B200_CUDA(cudaSetDevice(w.dev)); B200_CUDA(cudaStreamSynchronize(w.st)); }
        }

        PhaseTimer pt;
        pt.w = &ws.ranks[0];
        pt.on = timers != 0;
        for (auto& w : ws.ranks) { B200_CUDA(cudaSetDevice(w.dev)); B200_CUDA(cudaEventRecord(w.ev0, w.st)); }
        B200_CUDA(cudaSetDevice(ws.ranks[0].dev));
        pt.mark(T_BEGIN);

        int np[kMaxRanks];
        // ---- setup: r = b - A x0, p = r, rr_old = r.r, b_norm = sqrt(rr_old)   (cg_solver.cu:498-528)
        if (multi) {
            const uint32_t e = ++g.halo_epoch;
            for (auto& w : ws.ranks) { B200_CUDA(cudaSetDevice(w.dev)); B200_K(push_halo(w, w.x, e, nullptr)); }
        }
        for (size_t l = 0; l < L; l++) {
            RankWs& w = ws.ranks[l];
            B200_CUDA(cudaSetDevice(w.dev));
            if (fused) {
                b200_band band;
                wire_band(w, &band, true, g.halo_epoch);
                B200_K(b200_cg_residual_init(&band, w.x, w.b, w.r, w.p, w.partials, w.scalars, w.st));
                np[l] = n_partials_spmv[l];
            } else {
                if (op->run_device(w.x, w.Ap) != 0) return 1;
                B200_K(b200_residual_init_generic(w.nl, w.b, w.Ap, w.r, w.p, w.partials, &np[l], w.st));
            }
        }
        B200_CUDA(cudaSetDevice(ws.ranks[0].dev));
        pt.mark(T_INIT_R);
        if (for_ranks_reduce(0, tol, np, false, ++g.red_epoch)) return 1;
        B200_CUDA(cudaSetDevice(ws.ranks[0].dev));
        pt.mark(T_RED_RR0);
        if (pcg) {  // p0 = z0 = D^-1 r0, rho_0 = r0.z0
            RankWs& w = ws.ranks[0];
            B200_K(b200_pcg_init(w.nl, w.r, w.dinv, w.p, w.partials, &np[0], w.st));
            if (for_ranks_reduce(4, tol, np, false, ++g.red_epoch)) return 1;
        }
        if (multi) {
            const uint32_t e = ++g.halo_epoch;
            for (auto& w : ws.ranks) { B200_CUDA(cudaSetDevice(w.dev)); B200_K(push_halo(w, w.p, e, nullptr)); }
            B200_CUDA(cudaSetDevice(ws.ranks[0].dev));
            pt.mark(T_HALO);
        }
        if (verbose >= 1) {
            RankWs& w0 = ws.ranks[0];
            B200_CUDA(cudaSetDevice(w0.dev));
            B200_CUDA(cudaStreamSynchronize(w0.st));
            printf("[%s] Initial residual: %e\n", tag, (double)w0.status->b_norm);
        }

        // ---- iterations (cg_solver.cu:538-638)
        const int lag = verbose >= 2 ? 0 : kLag;
        const bool dx = fused && schedule_deferred_x() && !pcg;
        if (multi && dx) {
            // first direction: p0 = r0, its edges are in the landing buffers (pushed above)
            for (auto& w : ws.ranks) {
                B200_CUDA(cudaSetDevice(w.dev));
                B200_K(b200_cg_halo_dir(w.rank > 0 ? landing_prev(w.rank) : nullptr,
                                        w.rank < g.world - 1 ? landing_next(w.rank) : nullptr, nullptr, nullptr,
                                        halo_dir(w.rank, 0, 0), halo_dir(w.rank, 0, 1), ws.grid, flag_prev(w.rank),
                                        flag_next(w.rank), g.halo_epoch, w.scalars, 1, w.st));
            }
            B200_CUDA(cudaSetDevice(ws.ranks[0].dev));
            pt.mark(T_HALO);
        }
        const size_t loop_mark0 = pt.used;            // first phase mark of the iteration loop
        // classic: K1, R, K2, R, K3 (halo push fused into K3); deferred x: K1F, R, K2r, R (+ halo direction)
        const size_t marks_per_iter = dx ? 4 : 5;
        int launched = 0;
        bool done = false;
        Nvtx range_solver("CG_Solver");
        for (int it = 0; it < max_iters && !done; it++) {
            Nvtx range_iter("CG_Iteration");
            nvtxRangePushA("SpMV");
            for (size_t l = 0; l < L; l++) {  // K1: Ap = A p, partials p.Ap
                RankWs& w = ws.ranks[l];
                B200_CUDA(cudaSetDevice(w.dev));
                if (dx) {
                    // direction k lives in p (k even) or p2 (k odd)
                    double* pcur = (it & 1) ? w.p2 : w.p;
                    double* pold = (it & 1) ? w.p : w.p2;
                    b200_band band;
                    wire_band_dir(w, &band, it);
                    if (it == 0) B200_K(b200_cg_spmv_dot(&band, pcur, w.Ap, w.partials, w.scalars, w.st));
                    else B200_K(b200_cg_spmv_fused(&band, pold, w.r, pcur, w.x, w.Ap, w.partials, w.scalars, w.st));
                    np[l] = n_partials_spmv[l];
                } else if (fused) {
                    b200_band band;
                    wire_band(w, &band, true, g.halo_epoch);
                    B200_K(b200_cg_spmv_dot(&band, w.p, w.Ap, w.partials, w.scalars, w.st));
                    np[l] = n_partials_spmv[l];
                } else if (operator_spmv_dot(op, w.p, w.Ap, w.partials, w.max_partials, &np[l], w.scalars) != 0) {
                    // foreign operator (or unaligned arrays): SpMV through the vtable, then the dot pass
                    if (op->run_device(w.p, w.Ap) != 0) return 1;
                    B200_K(b200_dot_partials(w.nl, w.scalars, w.Ap, w.p, w.partials, &np[l], w.st));
                }
            }
            B200_CUDA(cudaSetDevice(ws.ranks[0].dev));
            pt.mark(T_SPMV);
            nvtxRangePop();
            nvtxRangePushA("Dot_Product");
            if (for_ranks_reduce(1, tol, np, false, ++g.red_epoch)) return 1;  // alpha
            nvtxRangePop();
            B200_CUDA(cudaSetDevice(ws.ranks[0].dev));
            pt.mark(T_RED_PAP);
            nvtxRangePushA("BLAS_AXPY");
            if (dx) {
                // K2r: r -= alpha Ap, partials r.r; multi-GPU: the edges of the new r go straight into
                // the neighbours' landing buffers, the last CTA publishes the arrival epoch
                const uint32_t e = multi ? ++g.halo_epoch : 0;
                for (size_t l = 0; l < L; l++) {
                    RankWs& w = ws.ranks[l];
                    B200_CUDA(cudaSetDevice(w.dev));
                    if (multi) {
                        double* dprev = w.rank > 0 ? landing_next(w.rank - 1) : nullptr;
                        double* dnext = w.rank < g.world - 1 ? landing_prev(w.rank + 1) : nullptr;
                        uint32_t* fprev = w.rank > 0 ? flag_next(w.rank - 1) : nullptr;
                        uint32_t* fnext = w.rank < g.world - 1 ? flag_prev(w.rank + 1) : nullptr;
                        B200_K(b200_cg_update_r_push(w.nl, w.scalars, w.Ap, w.r, w.partials2, &np[l], ws.grid, dprev, dnext,
                                                     fprev, fnext, e, g.xchg[w.rank], w.st));
                    } else {
                        B200_K(b200_cg_update_r(w.nl, w.scalars, w.Ap, w.r, w.partials2, &np[l], w.st));
                    }
                }
            } else if (pcg) {  // K2p: + partials r.z with z = D^-1 r
                RankWs& w = ws.ranks[0];
                B200_K(b200_pcg_update_xr(w.nl, w.scalars, w.p, w.Ap, w.dinv, w.x, w.r, w.partials2, w.partials, &np[0], w.st));
            } else {
                for (size_t l = 0; l < L; l++) {  // K2: x += alpha p, r -= alpha Ap, partials r.r
                    RankWs& w = ws.ranks[l];
                    B200_CUDA(cudaSetDevice(w.dev));
                    B200_K(b200_cg_update_xr(w.nl, w.scalars, w.p, w.Ap, w.x, w.r, w.partials2, &np[l], w.st));
                }
            }
            B200_CUDA(cudaSetDevice(ws.ranks[0].dev));
            pt.mark(T_XR);
            nvtxRangePop();
            nvtxRangePushA("Dot_Product");
            // convergence, beta; multi-GPU deferred-x: + halo copies of the next direction,
            // p_halo = r_halo + beta p_halo_old (the r edges were pushed by K2r above)
            if (pcg) {
                RankWs& w = ws.ranks[0];
                B200_K(b200_pcg_reduce(w.partials2, w.partials, np[0], tol, w.scalars, w.status_dev, w.st));
            } else if (for_ranks_reduce(2, tol, np, true, ++g.red_epoch, (dx && multi) ? it : -1)) return 1;
            nvtxRangePop();
            B200_CUDA(cudaSetDevice(ws.ranks[0].dev));
            pt.mark(T_RED_RR);
            if (!dx) {
                Nvtx range_p(multi ? "BLAS_AXPBY+Halo_Exchange" : "BLAS_AXPBY");
                if (pcg) {  // K3p: p = D^-1 r + beta p
                    RankWs& w = ws.ranks[0];
                    B200_K(b200_pcg_update_p(w.nl, w.scalars, w.r, w.dinv, w.p, w.st));
                } else if (!multi) {
                    for (auto& w : ws.ranks) {  // K3: p = r + beta p
                        B200_CUDA(cudaSetDevice(w.dev));
                        B200_K(b200_cg_update_p(w.nl, w.scalars, w.r, w.p, w.st));
                    }
                } else {
                    // K3 + halo push in one launch: the edge elements of the new p go straight into the
                    // neighbours' landing buffers, the last CTA publishes the arrival epoch
                    const uint32_t e = ++g.halo_epoch;
                    for (auto& w : ws.ranks) {
                        B200_CUDA(cudaSetDevice(w.dev));
                        double* dprev = w.rank > 0 ? landing_next(w.rank - 1) : nullptr;
                        double* dnext = w.rank < g.world - 1 ? landing_prev(w.rank + 1) : nullptr;
                        uint32_t* fprev = w.rank > 0 ? flag_next(w.rank - 1) : nullptr;
                        uint32_t* fnext = w.rank < g.world - 1 ? flag_prev(w.rank + 1) : nullptr;
                        B200_K(b200_cg_update_p_push(w.nl, w.scalars, w.r, w.p, ws.grid, dprev, dnext, fprev, fnext, e,
                                                     g.xchg[w.rank], w.st));
                    }
                }
                B200_CUDA(cudaSetDevice(ws.ranks[0].dev));
                pt.mark(T_P);
            }
            RankWs& w0 = ws.ranks[0];
            B200_CUDA(cudaSetDevice(w0.dev));
            B200_CUDA(cudaEventRecord(iter_event(w0, it), w0.st));
            launched = it + 1;
            if (it >= lag) {  // look at the iteration finished `lag` launches ago
                B200_CUDA(cudaEventSynchronize(w0.iter_ev[it - lag]));
                if (verbose >= 2)
                    printf("[%s] Iter %3d: residual = %e (rel = %e)\n", tag, (int)w0.status->iterations,
                           (double)w0.status->residual, (double)w0.status->residual / (double)w0.status->b_norm);
                if (w0.status->converged || w0.status->error) done = true;
            }
        }
        const size_t loop_mark1 = pt.used;
        if (dx) {
            // the x update of the last completed iteration is still pending: x += alpha p_last
            for (auto& w : ws.ranks) {
                B200_CUDA(cudaSetDevice(w.dev));
                B200_K(b200_cg_finish_x(w.nl, w.scalars, w.p, w.p2, w.x, w.st));
            }
            B200_CUDA(cudaSetDevice(ws.ranks[0].dev));
            pt.mark(T_P);
        }
        (void)launched;
        for (auto& w : ws.ranks) { B200_CUDA(cudaSetDevice(w.dev)); B200_CUDA(cudaEventRecord(w.ev1, w.st)); }
        double total = 0.0;
        for (auto& w : ws.ranks) {
            B200_CUDA(cudaSetDevice(w.dev));
            B200_CUDA(cudaEventSynchronize(w.ev1));
            float ms = 0.f;
            B200_CUDA(cudaEventElapsedTime(&ms, w.ev0, w.ev1));
            if (ms > total) total = ms;  // slowest rank, like the reference's MPI_Reduce(MAX) (:758-763)
        }
        RankWs& w0 = ws.ranks[0];
        if (w0.status->error) {
            fprintf(stderr, "[b200] peer exchange timed out\n");
            return B200_ETIMEOUT;
        }
        out->total_ms = total;
        out->iterations = w0.status->iterations;
        out->residual = w0.status->residual;
        out->b_norm = w0.status->b_norm;
        out->converged = (out->residual / out->b_norm < tol) ? 1 : 0;  // recomputed on the host (cg_solver.cu:656)
        if (pt.on) {
            B200_CUDA(cudaSetDevice(w0.dev));
            for (size_t k = 1; k < pt.used; k++) {
                // launches enqueued after convergence (the host polls kLag iterations behind) are
                // no-ops on the device: keep them out of the per-phase times and launch counts
                if (k >= loop_mark0 && k < loop_mark1 && (k - loop_mark0) / marks_per_iter >= (size_t)out->iterations) continue;
                float ms = 0.f;
                cudaEventElapsedTime(&ms, w0.phase_ev[k - 1], w0.phase_ev[k]);
                out->phase_ms[pt.tags[k]] += ms;
                out->phase_cnt[pt.tags[k]]++;
            }
        }

        // ---- solution back to the host + checksums (cg_solver.cu:645,658-665), sums on the device
        for (auto& w : ws.ranks) {
            B200_CUDA(cudaSetDevice(w.dev));
            B200_CUDA(cudaMemcpyAsync(x_host + w.off, w.x, (size_t)w.nl * sizeof(double), cudaMemcpyDeviceToHost, w.st));
        }
        int npc[kMaxRanks];
        for (size_t l = 0; l < L; l++) {
            RankWs& w = ws.ranks[l];
            B200_CUDA(cudaSetDevice(w.dev));
            B200_K(b200_checksum_partials(w.nl, w.x, w.partials, w.partials2, &npc[l], w.st));
        }
        double sums[2] = {0, 0};
        for (int which_buf = 0; which_buf < 2; which_buf++) {
            if (for_ranks_reduce(3, tol, npc, which_buf == 1, ++g.red_epoch)) return 1;
            B200_CUDA(cudaSetDevice(w0.dev));
            B200_CUDA(cudaMemcpyAsync(&sums[which_buf], w0.sums, sizeof(double), cudaMemcpyDeviceToHost, w0.st));
            B200_CUDA(cudaStreamSynchronize(w0.st));
        }
        for (auto& w : ws.ranks) { B200_CUDA(cudaSetDevice(w.dev)); B200_CUDA(cudaStreamSynchronize(w.st)); }
        out->sum = sums[0];
        out->norm = sqrt(sums[1]);
        B200_CUDA(cudaSetDevice(ws.ranks[0].dev));
        return 0;
    }
};

// partition rule of the reference (cg_solver_mgpu_partitioned.cu:262-268)
void partition(long long N, int P, int r, long long* nl, long long* off) {
    long long q = N / P;
    *off = (long long)r * q;
    *nl = (r == P - 1) ? N - *off : q;
}

// (re)build the workspace for this matrix / world; returns 0 when ws is ready
int prepare_workspace(MatrixData* mat, SpmvOperator* op, bool fused_from_op, Engine* eng) {
    const long long N = rows64(mat);
    // Workspace re-use between solves.  The vectors only depend on the shape; a band owned by the
    // workspace is only trusted again when its content is fully determined by the shape (synthetic
    // stencil) -- bands cut from caller-provided entries are rebuilt on every call, like the
    // reference, which uploads its local CSR per solve (cg_solver_mgpu_partitioned.cu:303-343).
    const void* key = (fused_from_op || !eng->fused) ? (const void*)op : (const void*)nullptr;
    bool same = ws.N == N && ws.grid == mat->grid_size && ws.world == g.world && ws.matrix_key == key &&
                ws.matrix_nnz == mat->nnz && !ws.ranks.empty();
    // (a band borrowed from an operator is re-read below on every call: the operator may have been
    // freed and re-initialised between two solves, and any of its three arrays may have moved)
    if (same && !fused_from_op && eng->fused) same = is_synthetic(mat) && ws.synthetic;
    if (!same) {
        ws.release();
        ws.N = N; ws.grid = mat->grid_size; ws.world = g.world; ws.matrix_key = key; ws.matrix_nnz = mat->nnz;
        ws.synthetic = is_synthetic(mat);
        const int L = g.inited ? g.nlocal : 1;
        ws.ranks.resize(L);
        for (int l = 0; l < L; l++) {
            RankWs& w = ws.ranks[l];
            w.rank = g.inited ? g.local_rank[l] : 0;
            if (g.inited) w.dev = g.local_dev[l];
            else cudaGetDevice(&w.dev);
            partition(N, g.world, w.rank, &w.nl, &w.off);
            B200_CUDA(cudaSetDevice(w.dev));
            if (eng->fused || fused_from_op) {
                if (g.single_device && l > 0) { w.st = ws.ranks[0].st; w.own_stream = false; }
                else { B200_CUDA(cudaStreamCreateWithFlags(&w.st, cudaStreamNonBlocking)); w.own_stream = true; }
            } else {
                w.st = nullptr;  // operator launches on the default stream
            }
            if (fused_from_op) {
                w.band = *operator_band(op);
                w.own_band = false;
            } else if (eng->fused) {
                if (!is_synthetic(mat) && build_csr_struct(mat) != EXIT_SUCCESS) return 1;
                if (upload_band_csr(mat, w.off, w.nl, &w.band, w.st)) return 1;
                w.own_band = true;
            }
            int maxp = 148 * 8;
            if (!(eng->fused || fused_from_op)) {
                const long long k = b200_csr_dot_partials_capacity(w.nl);  // generic operators: fused SpMV + dot
                if (k > maxp) maxp = (int)k;
            }
            if (eng->fused || fused_from_op) {
                b200_band band;
                w.band.describe(&band);
                const int k = b200_cg_max_partials(&band);
                if (k < 0) { fprintf(stderr, "[b200] %s\n", b200_last_error()); return 1; }
                maxp = k;
            }
            if (alloc_rank_vectors(w, maxp)) return 1;
        }
    }
    for (size_t l = 0; l < ws.ranks.size(); l++) {
        if (fused_from_op) {
            const DeviceBand* ob = operator_band(op);
            if (!ob || ob->n_local != ws.ranks[l].nl) return 1;
            ws.ranks[l].band = *ob;  // current device arrays of the operator (never owned here)
            ws.ranks[l].own_band = false;
        }
        if (eng->fused || fused_from_op) {
            b200_band band;
            ws.ranks[l].band.describe(&band);
            eng->n_partials_spmv[l] = b200_stencil5_num_partials(&band);
        }
    }
    return 0;
}

void fill_stats(const SolveOut& o, CGStats* s) {
    s->iterations = o.iterations;
    s->residual_norm = o.residual;
    s->time_total_ms = o.total_ms;
    // fused kernels: SpMV time includes the p.Ap partial sums, BLAS-1 time the r.r partial sums;
    // "reductions" are the two final-sum launches per iteration
    s->time_spmv_ms = o.phase_ms[T_SPMV] + o.phase_ms[T_INIT_R];
    s->time_blas1_ms = o.phase_ms[T_XR] + o.phase_ms[T_P];
    s->time_reductions_ms = o.phase_ms[T_RED_PAP] + o.phase_ms[T_RED_RR] + o.phase_ms[T_RED_RR0];
    s->converged = o.converged;
    s->solution_sum = o.sum;
    s->solution_norm = o.norm;
}

void print_summary(const char* tag, const CGStats* s) {
    printf("[%s] Converged: %s\n", tag, s->converged ? "YES" : "NO");
    printf("[%s] Iterations: %d\n", tag, s->iterations);
    printf("[%s] Final residual: %e\n", tag, s->residual_norm);
    printf("[%s] Time breakdown:\n", tag);
    const double t = s->time_total_ms > 0 ? s->time_total_ms : 1.0;
    printf("     Total:      %.3f ms\n", s->time_total_ms);
    printf("     SpMV:       %.3f ms (%.1f%%)\n", s->time_spmv_ms, 100.0 * s->time_spmv_ms / t);
    printf("     BLAS1:      %.3f ms (%.1f%%)\n", s->time_blas1_ms, 100.0 * s->time_blas1_ms / t);
    printf("     Reductions: %.3f ms (%.1f%%)\n", s->time_reductions_ms, 100.0 * s->time_reductions_ms / t);
}

int solve_single(SpmvOperator* op, MatrixData* mat, const double* b, double* x, CGConfig cfg, CGStats* stats,
                 const char* tag, bool pcg = false) {
    if (!op || !mat || !b || !x || !stats) return 1;
    if (!op->run_device) {
        fprintf(stderr, "[ERROR] Operator '%s' does not support device-native interface\n", op->name);
        return 1;
    }
    if (g.inited && g.world > 1) {
        fprintf(stderr, "[ERROR] %s: a multi-GPU world is active; use cg_solve_mgpu_partitioned\n", tag);
        return 1;
    }
    Engine eng;
    eng.op = op;
    eng.fused = false;
    const bool fused_from_op = operator_band(op) != nullptr;
    if (prepare_workspace(mat, op, fused_from_op, &eng)) return 1;
    eng.fused = fused_from_op;
    eng.pcg = pcg;
    SolveOut o;
    int rc = eng.solve(b, x, cfg.max_iters, cfg.tolerance, cfg.verbose, cfg.enable_detailed_timers, tag, &o);
    if (rc) return rc;
    g_last = o;
    fill_stats(o, stats);
    if (cfg.verbose >= 1) print_summary(tag, stats);
    return 0;
}

}  // namespace

int cg_solve_device(SpmvOperator* spmv_op, MatrixData* mat, const double* b, double* x, CGConfig config,
                    CGStats* stats) {
    return solve_single(spmv_op, mat, b, x, config, stats, "CG-DEVICE");
}

// Jacobi-preconditioned CG, M = diag(A).  Not in the reference (cg_solver.h:6-7 names preconditioning
// as future work); same argument list, conventions and statistics as cg_solve_device.
int pcg_solve_device(SpmvOperator* spmv_op, MatrixData* mat, const double* b, double* x, CGConfig config,
                     CGStats* stats) {
    return solve_single(spmv_op, mat, b, x, config, stats, "PCG-DEVICE", true);
}

// Host-interface variant.  The reference runs the same recurrence but round-trips every SpMV
// through host memory (cg_solver.cu:243-250); results are identical, so operators with a device
// entry point take the device-resident path here as well.
int cg_solve(SpmvOperator* spmv_op, MatrixData* mat, const double* b, double* x, CGConfig config, CGStats* stats) {
    return solve_single(spmv_op, mat, b, x, config, stats, "CG");
}

int cg_solve_mgpu_partitioned(SpmvOperator* spmv_op, MatrixData* mat, const double* b, double* x,
                              CGConfigMultiGPU config, CGStatsMultiGPU* stats) {
    (void)spmv_op;  // unused in the reference too (callers pass NULL)
    if (!mat || !b || !x || !stats) return 1;
    if (!g.inited) {  // default world: every visible GPU (override with B200_GPUS)
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
            fprintf(stderr, "[b200] no CUDA device\n");
            return B200_ENODEV;
        }
        const char* env = getenv("B200_GPUS");
        int world = env ? atoi(env) : ndev;
        if (world < 1) world = 1;
        int devs[kMaxRanks];
        for (int r = 0; r < world && r < kMaxRanks; r++) devs[r] = r % ndev;
        if (b200_mgpu_init_single_process(world, devs, mat->grid_size > 0 ? mat->grid_size : 1)) return 1;
    }
    const long long n = mat->grid_size;
    if (n < 1 || n * n != rows64(mat)) {
        fprintf(stderr, "[ERROR] cg_solve_mgpu_partitioned needs a stencil matrix with STENCIL_GRID_SIZE\n");
        return 1;
    }
    if ((size_t)n > g.halo_cap) {
        fprintf(stderr, "[ERROR] grid %lld exceeds the halo capacity %zu of the active multi-GPU world\n", n, g.halo_cap);
        return 1;
    }
    if (g.world > 1 && rows64(mat) / g.world < n) {
        fprintf(stderr, "[ERROR] band of %lld rows is smaller than one grid row (%lld)\n", rows64(mat) / g.world, n);
        return 1;
    }
    {
        // a band is addressed with 32-bit local offsets: fewer than 2^31 non-zeros per rank; global
        // column ids are stored modulo 2^32
        const long long band_rows = rows64(mat) / g.world + rows64(mat) % g.world;
        if (5 * band_rows >= 2147483647LL || rows64(mat) >= 4294967295LL) {
            fprintf(stderr, "[ERROR] %lld rows over %d rank(s): a band must stay below 2^31 non-zeros\n", rows64(mat), g.world);
            return 1;
        }
    }
    Engine eng;
    eng.op = nullptr;
    eng.fused = true;
    if (prepare_workspace(mat, nullptr, false, &eng)) return 1;
    SolveOut o;
    int rc = eng.solve(b, x, config.max_iters, config.tolerance, config.verbose, config.enable_detailed_timers,
                       "CG-MGPU", &o);
    if (rc) return rc;
    g_last = o;
    memset(stats, 0, sizeof *stats);
    stats->iterations = o.iterations;
    stats->residual_norm = o.residual;
    stats->time_total_ms = o.total_ms;
    stats->converged = o.converged;
    stats->solution_sum = o.sum;
    stats->solution_norm = o.norm;
    stats->time_spmv_ms = o.phase_ms[T_SPMV];
    stats->time_blas1_ms = o.phase_ms[T_XR] + o.phase_ms[T_P];
    stats->time_reductions_ms = o.phase_ms[T_RED_PAP] + o.phase_ms[T_RED_RR] + o.phase_ms[T_RED_RR0];
    stats->time_allreduce_ms = 0.0;  // the rank exchange is inside the reduce kernels
    stats->time_allgather_ms = o.phase_ms[T_HALO];
    auto avg = [&](int t) { return o.phase_cnt[t] ? o.phase_ms[t] / o.phase_cnt[t] : 0.0; };
    stats->time_dot_rs_initial_ms = o.phase_ms[T_RED_RR0];
    stats->time_dot_pAp_ms = avg(T_RED_PAP);
    stats->time_dot_rs_new_ms = avg(T_RED_RR);
    stats->time_axpy_update_x_ms = avg(T_XR);  // x and r are updated by one fused kernel
    stats->time_axpy_update_r_ms = 0.0;
    stats->time_axpby_update_p_ms = avg(T_P);
    stats->time_initial_r_ms = o.phase_ms[T_INIT_R];
    if (config.verbose >= 1 && b200_mgpu_rank() == 0) {
        printf("[CG-MGPU] GPUs: %d  Converged: %s  Iterations: %d  Final residual: %e\n", g.world,
               stats->converged ? "YES" : "NO", stats->iterations, stats->residual_norm);
        printf("[CG-MGPU] Total: %.3f ms\n", stats->time_total_ms);
    }
    return 0;
}

// Per-phase event times (ms) and launch counts of the most recent solve that ran with
// enable_detailed_timers: [0] unused, [1] K1 SpMV+p.Ap, [2] reduce p.Ap, [3] K2 x/r update + r.r,
// [4] reduce r.r, [5] K3 p update, [6] halo push, [7] residual init, [8] reduce r0.r0.
extern "C" int b200_last_phase_times(double* ms9, int* count9) {
    for (int t = 0; t < T_N; t++) {
        if (ms9) ms9[t] = g_last.phase_ms[t];
        if (count9) count9[t] = g_last.phase_cnt[t];
    }
    return T_N;
}

// Declared but never defined in the reference (cg_solver_mgpu.h:88-89, "full replication").
// The partitioned solver supersedes it; same signature, same result.
int cg_solve_mgpu(SpmvOperator* spmv_op, MatrixData* mat, const double* b, double* x, CGConfigMultiGPU config,
                  CGStatsMultiGPU* stats) {
    return cg_solve_mgpu_partitioned(spmv_op, mat, b, x, config, stats);
}

// ------------------------------------------------------------------------------------------------
// "stencil5-halo-mgpu": SpMV over row bands on all visible GPUs (declared in the reference,
// include/spmv.h:139, never defined).  Host vectors in, host vectors out; halos are cut from the
// host vector, so no device exchange is needed.  kernel_time_ms = slowest GPU.
// ------------------------------------------------------------------------------------------------
namespace {
struct HaloOp {
    struct Part {
        int dev = 0;
        long long off = 0, nl = 0;
        DeviceBand band;
        double *x = nullptr, *hp = nullptr, *hn = nullptr, *y = nullptr;
        cudaStream_t st = nullptr;
        cudaEvent_t e0 = nullptr, e1 = nullptr;
    };
    std::vector<Part> parts;
    int n = 0;
    long long N = 0;
    void reset() {
        for (auto& p : parts) {
            cudaSetDevice(p.dev);
            p.band.release();
            cudaFree(p.x); cudaFree(p.hp); cudaFree(p.hn); cudaFree(p.y);
            if (p.st) cudaStreamDestroy(p.st);
            if (p.e0) cudaEventDestroy(p.e0);
            if (p.e1) cudaEventDestroy(p.e1);
        }
        parts.clear();
    }
    int init(MatrixData* mat) {
        reset();
        n = mat->grid_size; N = mat->rows;
        if (n < 1 || (long long)n * n != N) return EXIT_FAILURE;
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return EXIT_FAILURE;
        const char* env = getenv("B200_GPUS");
        int P = env ? atoi(env) : ndev;
        if (P < 1) P = 1;
        while (P > 1 && N / P < n) P--;
        if (!is_synthetic(mat) && build_csr_struct(mat) != EXIT_SUCCESS) return EXIT_FAILURE;
        parts.resize(P);
        for (int r = 0; r < P; r++) {
            Part& p = parts[r];
            p.dev = r % ndev;
            partition(N, P, r, &p.nl, &p.off);
            B200_CUDA(cudaSetDevice(p.dev));
            B200_CUDA(cudaStreamCreateWithFlags(&p.st, cudaStreamNonBlocking));
            if (upload_band_csr(mat, p.off, p.nl, &p.band, p.st)) return EXIT_FAILURE;
            B200_CUDA(cudaMalloc(&p.x, (size_t)p.nl * sizeof(double)));
            B200_CUDA(cudaMalloc(&p.y, (size_t)p.nl * sizeof(double)));
            if (r > 0) B200_CUDA(cudaMalloc(&p.hp, (size_t)n * sizeof(double)));
            if (r < P - 1) B200_CUDA(cudaMalloc(&p.hn, (size_t)n * sizeof(double)));
            B200_CUDA(cudaEventCreate(&p.e0));
            B200_CUDA(cudaEventCreate(&p.e1));
        }
        cudaSetDevice(parts[0].dev);
        return EXIT_SUCCESS;
    }
    int run_timed(const double* x, double* y, double* ms) {
        if (parts.empty()) return EXIT_FAILURE;
        for (auto& p : parts) {
            B200_CUDA(cudaSetDevice(p.dev));
            B200_CUDA(cudaMemcpyAsync(p.x, x + p.off, (size_t)p.nl * sizeof(double), cudaMemcpyHostToDevice, p.st));
            if (p.hp) B200_CUDA(cudaMemcpyAsync(p.hp, x + p.off - n, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, p.st));
            if (p.hn) B200_CUDA(cudaMemcpyAsync(p.hn, x + p.off + p.nl, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, p.st));
            B200_CUDA(cudaEventRecord(p.e0, p.st));
            b200_band b;
            p.band.describe(&b);
            b.d_halo_prev = p.hp; b.d_halo_next = p.hn;
            B200_K(b200_stencil5_spmv(&b, p.x, p.y, p.st));
            B200_CUDA(cudaEventRecord(p.e1, p.st));
            B200_CUDA(cudaMemcpyAsync(y + p.off, p.y, (size_t)p.nl * sizeof(double), cudaMemcpyDeviceToHost, p.st));
        }
        double worst = 0.0;
        for (auto& p : parts) {
            B200_CUDA(cudaSetDevice(p.dev));
            B200_CUDA(cudaStreamSynchronize(p.st));
            float t = 0.f;
            B200_CUDA(cudaEventElapsedTime(&t, p.e0, p.e1));
            if (t > worst) worst = t;
        }
        cudaSetDevice(parts[0].dev);
        if (ms) *ms = worst;
        return EXIT_SUCCESS;
    }
} g_halo_op;

int halo_init(MatrixData* m) { return g_halo_op.init(m); }
int halo_run_timed(const double* x, double* y, double* ms) { return g_halo_op.run_timed(x, y, ms); }
void halo_free() { g_halo_op.reset(); }
}  // namespace

SpmvOperator SPMV_STENCIL_HALO_MGPU = {"stencil5-halo-mgpu", halo_init, halo_run_timed, nullptr, halo_free};
