// b200_abi.cu -- extern "C" launchers: the only translation unit that includes kernel headers.
// See include/b200_kernels.h for the contract of every entry point.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <utility>

#include "../../include/b200_kernels.h"
#include "cg_kernels.cuh"
#include "csr_ell.cuh"
#include "generate.cuh"
#include "ingest.cuh"
#include "stencil5.cuh"
#include "stencil5_direct.cuh"
#include "stencil_layout.h"

#ifndef B200_PLAIN_VARIANT_DEFAULT
#define B200_PLAIN_VARIANT_DEFAULT 20
#endif
#ifndef B200_CG_KERNEL_DEFAULT
#define B200_CG_KERNEL_DEFAULT 1
#endif
#ifndef B200_PDL_DEFAULT
#define B200_PDL_DEFAULT 3
#endif

using namespace b200;

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

static int fail(int code, const char* fmt, const char* detail = "") {
    snprintf(g_err, sizeof g_err, fmt, detail);
    return code;
}

static int check_launch(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        cudaGetLastError();
        snprintf(g_err, sizeof g_err, "%s: %s", what, cudaGetErrorString(e));
        return (e == cudaErrorNoKernelImageForDevice || e == cudaErrorNoDevice ||
                e == cudaErrorInsufficientDriver)
                   ? B200_ENODEV
                   : B200_ECUDA;
    }
    return B200_OK;
}

extern "C" const char* b200_version(void) { return "b200-spmv-cg 0.1 (sm_100a)"; }
extern "C" const char* b200_last_error(void) { return g_err; }
extern "C" unsigned long long b200_launch_count(void) { return g_launches.load(); }

namespace {
// programmatic dependent launch, bit 0: the small kernels of the loop (reduce, halo direction, finish_x),
// bit 1: the STENCIL5 kernels, bit 2 (off by default): the BLAS-1 kernels K2 / K2r / K3 as well -- with the sweep
// kernels in front of them that gains 2 us per iteration at 12.5 M rows (3.538 -> 3.505 ms) and still loses 1.3 %
// at 400 M rows (93.05 -> 94.3 ms).  B200_PDL=<0..7>, b200_cg_set_pdl().
std::atomic<int> g_pdl{-1};
int pdl_mode() {
    int v = g_pdl.load(std::memory_order_relaxed);
    if (v < 0) {
        const char* e = getenv("B200_PDL");
        v = (e && e[0] >= '0' && e[0] <= '7') ? e[0] - '0' : B200_PDL_DEFAULT;
        g_pdl.store(v, std::memory_order_relaxed);
    }
    return v;
}

// Launch with programmatic stream serialisation: the kernel may be scheduled while the previous
// kernel of the stream drains; it calls griddep_wait() before it reads anything.
template <int BIT = 1, typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl_mode() & BIT) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// STENCIL5
// ------------------------------------------------------------------------------------------------
namespace {

struct Variant {
    int cols, warps, stages;
    const char* info;
};
// Tuning variants that are compiled in.  Round 1 swept fourteen shapes (profiles/, gpurun_out/sweep_*): only
// the default is used in production; the others that survive are the ones a measurement still compares
// against (3-stage ring, 8 warps, 2 warps).  Pruned ids fall back to the default.  Keep in sync with the
// dispatch switch below.
const Variant kVariants[] = {
    {4, 4, 2, "v0 (default): 128-col strips (4 cols/lane), 4 warps/CTA, 2-stage ring"},
    {0, 0, 0, nullptr},
    {0, 0, 0, nullptr},
    {4, 4, 3, "v3: 128-col strips (4 cols/lane), 4 warps/CTA, 3-stage ring"},
    {0, 0, 0, nullptr},
    {0, 0, 0, nullptr},
    {0, 0, 0, nullptr},
    {0, 0, 0, nullptr},
    {0, 0, 0, nullptr},
    {2, 2, 3, "v9: 64-col strips (2 cols/lane), 2 warps/CTA, 3-stage ring (round-1 default until the K1F sweep)"},
    {0, 0, 0, nullptr},
    {0, 0, 0, nullptr},
    {4, 8, 2, "v12: 128-col strips, 8 warps/CTA, 2-stage ring"},
    {4, 2, 2, "v13: 128-col strips, 2 warps/CTA, 2-stage ring"},
};
const int kNumVariants = (int)(sizeof(kVariants) / sizeof(kVariants[0]));
const int kDefaultRowsPerItem = 4;  // round-2 sweep (profiles/sweep_r02.md): 4 rows beat 8 by 1-1.5 % on the plain and dot kernels, equal on K1F
constexpr int kMaxDevices = 64;
int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = 0;
    return dev;
}

struct Geometry {
    Stencil5Args a;
    TailArgs tail;  // what follows the launch: fixed-order sum of the partials + exchange + recurrence (cg_reduce_kernel)
    bool sweep = false;  // fused CG passes in sequential-sweep form (csrc/stencil5_direct.cuh) instead of the ring
    int grid;
    int threads;
    size_t smem;
};

int build_geometry(const b200_band* b, const double* x, Geometry* g) {
    if (!b || !x) return fail(B200_EINVAL, "stencil5: NULL band or vector");
    if (b->grid_size < 1 || b->n_local < 0 || b->row_offset < 0) return fail(B200_EINVAL, "stencil5: bad geometry");
    const long long n = b->grid_size, N = n * n;
    if (b->row_offset + b->n_local > N) return fail(B200_EINVAL, "stencil5: band exceeds grid_size^2 rows");
    if (!b->d_values || !b->d_col_idx) return fail(B200_EINVAL, "stencil5: NULL matrix arrays");
    if (b->layout == 0 && !b->d_row_ptr) return fail(B200_EINVAL, "stencil5: CSR layout needs row_ptr");
    if (((uintptr_t)b->d_values & 15) != 0) return fail(B200_EINVAL, "stencil5: values must be 16-byte aligned");
    const int v = (b->variant >= 0 && b->variant < kNumVariants && kVariants[b->variant].info) ? b->variant : 0;
    const Variant& V = kVariants[v];
    Stencil5Args& a = g->a;
    memset(&a, 0, sizeof a);
    memset(&g->tail, 0, sizeof g->tail);
    g->tail.which = -1;
    a.row_ptr = b->layout == 0 ? b->d_row_ptr : nullptr;
    a.col_idx = b->d_col_idx;
    a.values = b->d_values;
    a.values_len = b->values_len;
    a.x = x;
    a.halo_prev = b->d_halo_prev;
    a.halo_next = b->d_halo_next;
    a.row_offset = b->row_offset;
    a.n_local = b->n_local;
    a.n = (int)n;
    if (b->layout == 0) {
        // element(i,j) = nnz_before(full CSR) - slice start = (4n-2) + (i-1)(5n-2) + 4 + 5(j-1) - base
        a.base0 = -n - 1 - stencil5_nnz_before(b->row_offset, n);
        a.row_stride = 5 * n - 2;
    } else {
        a.base0 = -5 * b->row_offset;
        a.row_stride = 5 * n;
    }
    const long long first = b->row_offset, last = b->row_offset + b->n_local - 1;
    long long i_first = first / n, i_last = b->n_local > 0 ? last / n : -1;
    if (i_first < 1) i_first = 1;
    if (i_last > n - 2) i_last = n - 2;
    const int W = 32 * V.cols;
    int R = b->rows_per_item > 0 ? b->rows_per_item : kDefaultRowsPerItem;
    a.rows_per_item = R;
    a.i_first = (int)i_first;
    a.i_last = (int)i_last;
    if (i_last >= i_first && n > 2) {
        a.n_strips = (int)((n - 2 + W - 1) / W);
        a.n_chunks = (int)((i_last - i_first + 1 + R - 1) / R);
        a.ctas_per_chunk = (a.n_strips + V.warps - 1) / V.warps;
        a.n_interior_ctas = a.n_chunks * a.ctas_per_chunk;
    } else {
        a.n_strips = a.n_chunks = a.n_interior_ctas = 0;
        a.ctas_per_chunk = 1;
    }
    a.n_boundary_rows = (n == 1) ? 1 : (int)(4 * n - 4);
    a.flag_prev = b->d_flag_prev;
    a.flag_next = b->d_flag_next;
    a.epoch = b->epoch;
    a.epoch_ptr = b->d_epoch_ptr;
    g->threads = V.warps * 32;
    const int nb = (a.n_boundary_rows + g->threads - 1) / g->threads;
    g->grid = a.n_interior_ctas + nb;
    g->smem = (size_t)V.warps * V.stages * (5 * W + 2) * 8 + (size_t)V.warps * V.stages * 8;
    return B200_OK;
}

template <int MODE, int COLS, int WARPS, int STAGES, bool CG>
int launch_one(const Geometry& g, cudaStream_t s) {
    auto k = stencil5_kernel<MODE, COLS, WARPS, STAGES, CG>;
    // per instantiation AND per device: the shared-memory opt-in is a per-device function attribute
    // (one process may drive several GPUs, cg_solver_mgpu_stencil --gpus=P)
    static std::atomic<bool> attr_set[kMaxDevices];
    const int dev = current_device();
    if (!attr_set[dev].load(std::memory_order_acquire)) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) {
            cudaGetLastError();
            snprintf(g_err, sizeof g_err, "stencil5: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return (e == cudaErrorNoKernelImageForDevice || e == cudaErrorInvalidDeviceFunction) ? B200_ENODEV
                                                                                                 : B200_ECUDA;
        }
        attr_set[dev].store(true, std::memory_order_release);
    }
    if (MODE == ST_PLAIN) k<<<g.grid, g.threads, g.smem, s>>>(g.a);
    else launch_pdl<2>(k, g.grid, g.threads, g.smem, s, g.a);
    return check_launch("stencil5_kernel");
}

template <int MODE, bool CG>
int launch_variant(int v, const Geometry& g, cudaStream_t s) {
    switch (v) {
        case 3: return launch_one<MODE, 4, 4, 3, CG>(g, s);
        case 9: return launch_one<MODE, 2, 2, 3, CG>(g, s);
        case 12: return launch_one<MODE, 4, 8, 2, CG>(g, s);
        case 13: return launch_one<MODE, 4, 2, 2, CG>(g, s);
        default: return launch_one<MODE, 4, 4, 2, CG>(g, s);
    }
}

// kernel family of the fused CG passes: 0 = bulk-copy ring (csrc/stencil5.cuh), 1 = sequential sweep
// (csrc/stencil5_direct.cuh, the default: equal or faster in every fused mode, and it has the registers for
// the deeper x retirement modes).  B200_CG_KERNEL=ring|sweep, b200_cg_set_kernel().  All three passes
// (residual init, SpMV + p.Ap, fused direction update) switch together: the CG schedules stay bit-identical.
std::atomic<int> g_cg_kernel{-1};
int cg_kernel_choice() {
    int v = g_cg_kernel.load(std::memory_order_relaxed);
    if (v < 0) {
        const char* e = getenv("B200_CG_KERNEL");
        v = e ? (strcmp(e, "ring") == 0 ? 0 : 1) : B200_CG_KERNEL_DEFAULT;
        g_cg_kernel.store(v, std::memory_order_relaxed);
    }
    return v;
}
int sweep_grid(long long n_local) {
    const long long rows_per_cta = 256LL * SWEEP_TILES;
    return (int)((n_local + rows_per_cta - 1) / rows_per_cta);
}
// called by the CG launchers right after build_geometry: the partial count follows the kernel family
void choose_cg_kernel(Geometry& g, bool force_sweep = false) {
    if ((force_sweep || cg_kernel_choice() == 1) && g.a.n_local > 0) {
        g.sweep = true;
        g.grid = sweep_grid(g.a.n_local);
    }
}

template <int MODE>
int launch_sweep(const b200_band* b, Geometry& g, cudaStream_t s) {
    const bool cg = (b->d_halo_prev != nullptr || b->d_halo_next != nullptr);
    if (cg) launch_pdl<2>(stencil5_sweep_kernel<MODE, true>, g.grid, 256, 0, s, g.a);
    else launch_pdl<2>(stencil5_sweep_kernel<MODE, false>, g.grid, 256, 0, s, g.a);
    return check_launch("stencil5_sweep_kernel");
}

template <int MODE>
int launch_stencil(const b200_band* b, Geometry& g, cudaStream_t s) {
    if (g.grid == 0) return B200_OK;
    if (MODE != ST_PLAIN && g.sweep) return launch_sweep<MODE>(b, g, s);
    // the ring kernel sits at its register limit with one x update per launch (128: 4 CTAs per SM); with a
    // second direction stream it spills (measured 7.24 against 5.80 ms at 20k x 20k), so the deeper x
    // retirement modes exist in sweep form only
    if constexpr (st_fused(MODE) && st_nx(MODE) != 1) return fail(B200_EINVAL, "stencil5: this mode runs on the sweep kernel only");
    else {
    const int v = (b->variant >= 0 && b->variant < kNumVariants && kVariants[b->variant].info) ? b->variant : 0;
    // peer-written halos must be read through L2 (ld.global.cg); single-GPU uses the read-only path
    const bool cg = (b->d_halo_prev != nullptr || b->d_halo_next != nullptr);
    return cg ? launch_variant<MODE, true>(v, g, s) : launch_variant<MODE, false>(v, g, s);
    }
}

}  // namespace

extern "C" const char* b200_stencil5_variant_info(int v) {
    if (v == 20) return "v20: sequential sweep, 1 row per thread (plain SpMV only)";
    if (v == 21) return "v21: sequential sweep, 2 rows per thread (plain SpMV only)";
    if (v == 22) return "v22: sequential sweep, 4 rows per thread (plain SpMV only)";
    return (v >= 0 && v < kNumVariants) ? kVariants[v].info : nullptr;
}

extern "C" int b200_stencil5_num_partials(const b200_band* band) {
    Geometry g;
    static const double dummy = 0;
    if (build_geometry(band, &dummy, &g) != B200_OK) return -1;
    const int sw = sweep_grid(band->n_local);  // either kernel family may run: room for the larger grid
    return g.grid > sw ? g.grid : sw;
}
extern "C" void b200_cg_set_kernel(int sweep) { g_cg_kernel.store(sweep ? 1 : 0, std::memory_order_relaxed); }
extern "C" int b200_cg_get_kernel(void) { return cg_kernel_choice(); }

namespace {
// plain product, sequential-sweep form (csrc/stencil5_direct.cuh): variants 20 / 21 / 22 = 1 / 2 / 4 rows per thread
template <int ROWS>
int launch_direct(const b200_band* b, const Geometry& g, cudaStream_t s) {
    const long long threads = (g.a.n_local + ROWS - 1) / ROWS;
    const long long blocks = (threads + 255) / 256;
    if (blocks == 0) return B200_OK;
    if (blocks > 2147483647LL) return fail(B200_EINVAL, "stencil5: too many blocks");
    const bool cg = (b->d_halo_prev != nullptr || b->d_halo_next != nullptr);
    if (cg) stencil5_direct_kernel<ROWS, true><<<(unsigned)blocks, 256, 0, s>>>(g.a);
    else stencil5_direct_kernel<ROWS, false><<<(unsigned)blocks, 256, 0, s>>>(g.a);
    return check_launch("stencil5_direct_kernel");
}
std::atomic<int> g_plain_variant{-1};
}  // namespace

// variant used by b200_stencil5_spmv when the band asks for the default (0): 0 = bulk-copy ring, 20..22 = sweep
extern "C" void b200_stencil5_set_plain_variant(int v) { g_plain_variant.store(v, std::memory_order_relaxed); }

extern "C" int b200_stencil5_spmv(const b200_band* band, const double* d_x, double* d_y, b200_stream stream) {
    Geometry g;
    int rc = build_geometry(band, d_x, &g);
    if (rc) return rc;
    if (!d_y) return fail(B200_EINVAL, "stencil5: NULL y");
    g.a.y = d_y;
    int v = band->variant;
    if (v == 0) {
        v = g_plain_variant.load(std::memory_order_relaxed);
        if (v < 0) {  // first use: B200_PLAIN_VARIANT overrides the built-in default
            const char* e = getenv("B200_PLAIN_VARIANT");
            v = e ? atoi(e) : B200_PLAIN_VARIANT_DEFAULT;
            g_plain_variant.store(v, std::memory_order_relaxed);
        }
    }
    // the sweep form has no flag waits: bands whose halos are still in flight stay on the ring kernel
    if (v >= 20 && v <= 22 && !band->d_flag_prev && !band->d_flag_next) {
        if (v == 20) return launch_direct<1>(band, g, (cudaStream_t)stream);
        if (v == 21) return launch_direct<2>(band, g, (cudaStream_t)stream);
        return launch_direct<4>(band, g, (cudaStream_t)stream);
    }
    return launch_stencil<ST_PLAIN>(band, g, (cudaStream_t)stream);
}

extern "C" int b200_spmv_stencil5_csr(const int* d_row_ptr, const int* d_col_idx, const double* d_values,
                                      const double* d_x, double* d_y, int N, int grid_size, b200_stream stream) {
    if ((long long)grid_size * grid_size != N) return fail(B200_EINVAL, "stencil5-csr: N != grid_size^2");
    b200_band b;
    memset(&b, 0, sizeof b);
    b.d_row_ptr = d_row_ptr; b.d_col_idx = d_col_idx; b.d_values = d_values;
    b.values_len = stencil5_nnz(grid_size);
    b.row_offset = 0; b.n_local = N; b.grid_size = grid_size; b.layout = 0;
    return b200_stencil5_spmv(&b, d_x, d_y, stream);
}

extern "C" int b200_spmv_stencil5_halo(const int* d_row_ptr, const int* d_col_idx, const double* d_values,
                                       const double* d_x_local, const double* d_x_halo_prev,
                                       const double* d_x_halo_next, double* d_y, int n_local, long long row_offset,
                                       long long N, int grid_size, b200_stream stream) {
    if ((long long)grid_size * grid_size != N) return fail(B200_EINVAL, "stencil5-halo: N != grid_size^2");
    b200_band b;
    memset(&b, 0, sizeof b);
    b.d_row_ptr = d_row_ptr; b.d_col_idx = d_col_idx; b.d_values = d_values;
    b.values_len = stencil5_nnz_before(row_offset + n_local, grid_size) - stencil5_nnz_before(row_offset, grid_size);
    b.row_offset = row_offset; b.n_local = n_local; b.grid_size = grid_size; b.layout = 0;
    b.d_halo_prev = d_x_halo_prev; b.d_halo_next = d_x_halo_next;
    return b200_stencil5_spmv(&b, d_x_local, d_y, stream);
}

// ------------------------------------------------------------------------------------------------
// generic CSR / ELLPACK
// ------------------------------------------------------------------------------------------------
namespace {
constexpr int kCsrWarps = 8;  // legacy warp-stream kernel (variant 100), kept for A/B measurements

struct CsrVariant {
    int warps, stages, win;
    const char* info;
};
// c0 and c6 are what b200_csr_plan_build picks; the other ring shapes of the round-1 sweep
// (profiles/csr_ring_r01.md) are pruned.  Keep in sync with the dispatch switch in launch_csr_variant.
const CsrVariant kCsrVariants[] = {
    {8, 4, 128, "c0 (default): 8 warps/CTA, ring of 4 x 128-entry windows, 4 CTAs/SM"},
    {0, 0, 0, nullptr},
    {0, 0, 0, nullptr},
    {0, 0, 0, nullptr},
    {0, 0, 0, nullptr},
    {0, 0, 0, nullptr},
    {8, 8, 128, "c6: 8 warps/CTA, ring of 8 x 128-entry windows, 2 CTAs/SM"},
};
const int kNumCsrVariants = (int)(sizeof(kCsrVariants) / sizeof(kCsrVariants[0]));

std::atomic<int> g_csr_default_variant{0};
thread_local long long g_csr_last_items = 0;  // items (= dot partials) of the last ring launch on this thread
int csr_default_variant() { return g_csr_default_variant.load(std::memory_order_relaxed); }
constexpr int kCsrGroupsPerItem = 32;  // 1024 rows per warp item

template <int WARPS, int STAGES, int WIN, int MODE, int MINB, bool DOT = false>
int launch_csr_ring_mode(const CsrArgs& a, int gpw_override, cudaStream_t s) {
    auto k = csr_ring_kernel<WARPS, STAGES, WIN, MODE, MINB, DOT>;
    const size_t smem = (size_t)WARPS * csr_ring_warp_bytes<STAGES, WIN>();
    // per instantiation and per device: CTAs that fit the whole GPU at once
    static std::atomic<int> resident_tab[kMaxDevices];
    const int dev_id = current_device();
    int resident_ctas = resident_tab[dev_id].load(std::memory_order_acquire);
    if (resident_ctas == 0) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int occ = 0, dev = 0, n_sm = 0;
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, WARPS * 32, smem);
        if (e == cudaSuccess) e = cudaGetDevice(&dev);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess || occ < 1 || n_sm < 1) {
            cudaGetLastError();
            snprintf(g_err, sizeof g_err, "csr: kernel setup: %s", cudaGetErrorString(e));
            return (e == cudaErrorNoKernelImageForDevice || e == cudaErrorInvalidDeviceFunction ||
                    e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver)
                       ? B200_ENODEV
                       : B200_ECUDA;
        }
        resident_ctas = occ * n_sm;
        resident_tab[dev_id].store(resident_ctas, std::memory_order_release);
    }
    // one item (run of 32-row groups) per warp, handed out in launch order; small matrices get
    // shorter items so that every SM has work
    const long long groups = (a.n_rows + 31) / 32;
    const long long resident_warps = (long long)resident_ctas * WARPS;
    long long gpw = gpw_override > 0 ? gpw_override : kCsrGroupsPerItem;
    if (gpw_override <= 0 && groups < 4 * resident_warps * gpw) {
        gpw = groups / (4 * resident_warps);
        if (gpw < 1) gpw = 1;
    }
    const long long warps = (groups + gpw - 1) / gpw;
    const long long blocks = (warps + WARPS - 1) / WARPS;
    if (blocks > 2147483647LL) return fail(B200_EINVAL, "csr: too many row blocks");
    g_csr_last_items = warps;  // one dot partial per item
    if (a.dot_partials && warps > a.dot_capacity) return fail(B200_EINVAL, "csr: dot partials buffer too small");
    k<<<(unsigned)blocks, WARPS * 32, smem, s>>>(a, (int)gpw);
    return check_launch("csr_ring_kernel");
}

template <int WARPS, int STAGES, int WIN, bool ELL, int MINB>
int launch_csr_ring(const CsrArgs& a, int gpw, cudaStream_t s) {
    if (!ELL) return launch_csr_ring_mode<WARPS, STAGES, WIN, 0, MINB>(a, gpw, s);
    return csr_ring_ell_is_lpr<STAGES, WIN>(a.ell_width) ? launch_csr_ring_mode<WARPS, STAGES, WIN, 1, MINB>(a, gpw, s)
                                                          : launch_csr_ring_mode<WARPS, STAGES, WIN, 2, MINB>(a, gpw, s);
}

template <bool ELL>
int launch_csr_variant(int v, int gpw, const CsrArgs& a, cudaStream_t s) {
    switch (v) {
        case 6: return launch_csr_ring<8, 8, 128, ELL, 2>(a, gpw, s);
        default: return launch_csr_ring<8, 4, 128, ELL, 4>(a, gpw, s);
    }
}

int launch_csr_legacy(const CsrArgs& a, cudaStream_t s) {
    const long long rows_per_cta = 32LL * kCsrWarps;
    const long long blocks = (a.n_rows + rows_per_cta - 1) / rows_per_cta;
    if (blocks > 2147483647LL) return fail(B200_EINVAL, "csr: too many row blocks");
    csr_warp_stream_kernel<kCsrWarps><<<(unsigned)blocks, kCsrWarps * 32, 0, s>>>(a);
    return check_launch("csr_warp_stream_kernel");
}

// variant = tuning variant + 1000 * (groups per item override); 100 = legacy warp-stream kernel
int launch_csr(const CsrArgs& a, int variant, cudaStream_t s) {
    if (a.n_rows == 0) return B200_OK;
    if (a.n_rows < 0) return fail(B200_EINVAL, "csr: negative row count");
    if (variant <= 0) variant = csr_default_variant();
    const int gpw = variant / 1000;
    variant %= 1000;
    // the bulk copies need 16-byte aligned arrays; anything else takes the register-staged kernel
    const bool aligned = (((uintptr_t)a.col_idx | (uintptr_t)a.values) & 15) == 0;
    if (variant == 100 || !aligned) {
        if (a.dot_partials) return fail(B200_EINVAL, "csr: the fused dot needs 16-byte aligned col_idx / values");
        return launch_csr_legacy(a, s);
    }
    if (variant >= kNumCsrVariants || !kCsrVariants[variant].info) variant = 0;
    if (a.dot_partials) {  // fused x.y partials: compiled for the default variant only
        if (a.row_ptr) return launch_csr_ring_mode<8, 4, 128, 0, 4, true>(a, gpw, s);
        return csr_ring_ell_is_lpr<4, 128>(a.ell_width) ? launch_csr_ring_mode<8, 4, 128, 1, 4, true>(a, gpw, s)
                                                        : launch_csr_ring_mode<8, 4, 128, 2, 4, true>(a, gpw, s);
    }
    return a.row_ptr ? launch_csr_variant<false>(variant, gpw, a, s) : launch_csr_variant<true>(variant, gpw, a, s);
}
}  // namespace

extern "C" void b200_csr_set_default_variant(int v) {
    g_csr_default_variant.store(v > 0 ? v : 0, std::memory_order_relaxed);
}

extern "C" const char* b200_csr_variant_info(int v) {
    return (v >= 0 && v < kNumCsrVariants) ? kCsrVariants[v].info : nullptr;
}

extern "C" int b200_csr_plan_build(const int* d_row_ptr, long long n_rows, long long nnz, b200_csr_plan* plan,
                                   b200_stream stream) {
    if (!plan || (!d_row_ptr && n_rows > 0)) return fail(B200_EINVAL, "csr plan: NULL argument");
    memset(plan, 0, sizeof *plan);
    cudaStream_t s = (cudaStream_t)stream;
    unsigned long long* d_bins = nullptr;
    if (cudaMalloc(&d_bins, 34 * sizeof(unsigned long long)) != cudaSuccess) return fail(B200_ENOMEM, "csr plan: cudaMalloc");
    cudaMemsetAsync(d_bins, 0, 34 * sizeof(unsigned long long), s);
    if (n_rows > 0) {
        long long blocks = (n_rows + 255) / 256;
        if (blocks > 148 * 16) blocks = 148 * 16;
        row_length_histogram_kernel<<<(unsigned)blocks, 256, 0, s>>>(d_row_ptr, n_rows, d_bins, d_bins + 33);
        int rc = check_launch("row_length_histogram_kernel");
        if (rc) { cudaFree(d_bins); return rc; }
    }
    unsigned long long h[34];
    cudaError_t e = cudaMemcpyAsync(h, d_bins, sizeof h, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(d_bins);
    if (e != cudaSuccess) return fail(B200_ECUDA, "csr plan: %s", cudaGetErrorString(e));
    memcpy(plan->hist, h, 33 * sizeof(unsigned long long));
    plan->max_row_len = h[33];
    plan->mean_row_len = n_rows > 0 ? (double)nnz / (double)n_rows : 0.0;
    // Scheme pick from the histogram.  Every 32-row group decides at run time between lane-per-row
    // (bit-exact k order, x gathered one group ahead) and warp-per-row; the histogram picks the ring:
    //   * median row <= 8 entries: the small ring (variant 0, 4 x 128 entries, 32 warps/SM) -- a group
    //     of 32 such rows fits it, everything runs lane-per-row at full occupancy;
    //   * median row 9..32 entries: the large ring (variant 6, 8 x 128 entries, 16 warps/SM): groups of
    //     up to 768 entries (24 per row) still run lane-per-row instead of falling to warp-per-row,
    //     where 9..15-entry rows would leave most lanes idle;
    //   * longer rows: warp-per-row throughout, small ring.
    // Rows longer than vector_threshold (32 = the ring's wrap mirror) always go warp-per-row.
    unsigned long long acc = 0, half = (unsigned long long)(0.5 * (double)n_rows);
    int median_bin = 0;
    for (int b = 0; b < 33; b++) { acc += h[b]; if (acc >= half) { median_bin = b; break; } }
    plan->variant = (median_bin == 4 || median_bin == 5) ? 6 : 0;
    plan->rows_per_block = 32 * kCsrGroupsPerItem;
    plan->window = kCsrVariants[plan->variant].win;
    plan->vector_threshold = 32;
    return B200_OK;
}

extern "C" int b200_spmv_csr(const b200_csr_plan* plan, const int* d_row_ptr, const int* d_col_idx,
                             const double* d_values, const double* d_x, double* d_y, long long n_rows, double alpha,
                             double beta, b200_stream stream) {
    if (!plan || !d_row_ptr || !d_x || !d_y) return fail(B200_EINVAL, "csr: NULL argument");
    CsrArgs a;
    memset(&a, 0, sizeof a);
    a.row_ptr = d_row_ptr; a.col_idx = d_col_idx; a.values = d_values; a.x = d_x; a.y = d_y;
    a.n_rows = n_rows; a.ell_width = 0; a.vector_threshold = plan->vector_threshold > 0 ? plan->vector_threshold : 32;
    a.alpha = alpha; a.beta = beta;
    return launch_csr(a, plan->variant, (cudaStream_t)stream);
}

extern "C" long long b200_csr_dot_partials_capacity(long long n_rows) {
    // upper bound of the item count for any tuning variant: one item per 32-row group
    return n_rows > 0 ? (n_rows + 31) / 32 : 0;
}

extern "C" int b200_spmv_csr_dot(const b200_csr_plan* plan, const int* d_row_ptr, const int* d_col_idx,
                                 const double* d_values, const double* d_x, double* d_y, long long n_rows,
                                 double* d_partials, long long partials_capacity, int* n_partials_out,
                                 const void* d_scalars, b200_stream stream) {
    if (!plan || !d_row_ptr || !d_x || !d_y || !d_partials) return fail(B200_EINVAL, "csr_dot: NULL argument");
    CsrArgs a;
    memset(&a, 0, sizeof a);
    a.row_ptr = d_row_ptr; a.col_idx = d_col_idx; a.values = d_values; a.x = d_x; a.y = d_y;
    a.n_rows = n_rows; a.ell_width = 0; a.vector_threshold = plan->vector_threshold > 0 ? plan->vector_threshold : 32;
    a.alpha = 1.0; a.beta = 0.0;
    a.dot_partials = d_partials; a.dot_capacity = partials_capacity;
    if (d_scalars) a.converged = &static_cast<const CGScalars*>(d_scalars)->converged;
    int rc = launch_csr(a, plan->variant, (cudaStream_t)stream);
    if (rc == B200_OK && n_partials_out) *n_partials_out = (int)g_csr_last_items;
    return rc;
}

extern "C" int b200_spmv_ellpack_dot(const int* d_indices, const double* d_values, const double* d_x, double* d_y,
                                     long long n_rows, int width, double* d_partials, long long partials_capacity,
                                     int* n_partials_out, const void* d_scalars, b200_stream stream) {
    if (!d_indices || !d_values || !d_x || !d_y || !d_partials) return fail(B200_EINVAL, "ellpack_dot: NULL argument");
    if (width < 1 || width > 1000) return fail(B200_EINVAL, "ellpack: width outside [1, MAX_WIDTH]");
    CsrArgs a;
    memset(&a, 0, sizeof a);
    a.row_ptr = nullptr; a.col_idx = d_indices; a.values = d_values; a.x = d_x; a.y = d_y;
    a.n_rows = n_rows; a.ell_width = width; a.vector_threshold = 1 << 20;
    a.alpha = 1.0; a.beta = 0.0;
    a.dot_partials = d_partials; a.dot_capacity = partials_capacity;
    if (d_scalars) a.converged = &static_cast<const CGScalars*>(d_scalars)->converged;
    int rc = launch_csr(a, 0, (cudaStream_t)stream);
    if (rc == B200_OK && n_partials_out) *n_partials_out = (int)g_csr_last_items;
    return rc;
}

extern "C" int b200_spmv_ellpack(const int* d_indices, const double* d_values, const double* d_x, double* d_y,
                                 long long n_rows, int width, double alpha, double beta, b200_stream stream) {
    if (!d_indices || !d_values || !d_x || !d_y) return fail(B200_EINVAL, "ellpack: NULL argument");
    if (width < 1 || width > 1000) return fail(B200_EINVAL, "ellpack: width outside [1, MAX_WIDTH]");
    CsrArgs a;
    memset(&a, 0, sizeof a);
    a.row_ptr = nullptr; a.col_idx = d_indices; a.values = d_values; a.x = d_x; a.y = d_y;
    a.n_rows = n_rows; a.ell_width = width; a.vector_threshold = 1 << 20;  // ELLPACK rows are uniform: always stream
    a.alpha = alpha; a.beta = beta;
    // widths 9..24: the large ring keeps the rows lane-per-row (see b200_csr_plan_build)
    const int variant = (csr_default_variant() == 0 && width > 8 && width <= 24) ? 6 : 0;
    return launch_csr(a, variant, (cudaStream_t)stream);
}

extern "C" int b200_spmv_stencil5_ellpack(const double* d_values, const int* d_col_indices, const double* d_x,
                                          double* d_y, int num_rows, int width, double alpha, double beta,
                                          int grid_size, b200_stream stream) {
    // fast path: the closed-form interior kernel on the width-5 layout when it is a plain product
    if (width == 5 && alpha == 1.0 && beta == 0.0 && (long long)grid_size * grid_size == num_rows) {
        b200_band b;
        memset(&b, 0, sizeof b);
        b.d_col_idx = d_col_indices; b.d_values = d_values; b.values_len = 5LL * num_rows;
        b.row_offset = 0; b.n_local = num_rows; b.grid_size = grid_size; b.layout = 1;
        return b200_stencil5_spmv(&b, d_x, d_y, stream);
    }
    return b200_spmv_ellpack(d_col_indices, d_values, d_x, d_y, num_rows, width, alpha, beta, stream);
}

// ------------------------------------------------------------------------------------------------
// fused CG steps
// ------------------------------------------------------------------------------------------------
extern "C" size_t b200_cg_scalars_bytes(void) { return sizeof(CGScalars); }
extern "C" size_t b200_cg_status_bytes(void) { return sizeof(CGStatus); }
extern "C" size_t b200_xchg_bytes(void) { return sizeof(XchgArea); }
extern "C" size_t b200_xchg_flag_prev_offset(void) { return offsetof(XchgArea, halo_flag_prev); }
extern "C" size_t b200_xchg_flag_next_offset(void) { return offsetof(XchgArea, halo_flag_next); }
extern "C" size_t b200_xchg_halo_seq_offset(void) { return offsetof(XchgArea, halo_seq); }

namespace {
constexpr int kBlas1CtasPerSm = 8;  // resident 256-thread CTAs per SM, one wave
// K2 / K2r (two or four input streams, unrolled double2 loads): 2 CTAs per SM measured best on B200
// (1.503 vs 1.555 ms for 24 B/row at 20k x 20k).  Both kernels MUST use the same grid: their r.r
// partials are summed in the same order, which keeps the two CG schedules bit-identical.
constexpr int kRrCtasPerSm = 2;
int sm_count() {  // per device (a process may drive several GPUs)
    static std::atomic<int> cache[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    int v = cache[dev].load(std::memory_order_relaxed);
    if (v == 0) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v < 1) v = 148;
        cache[dev].store(v, std::memory_order_relaxed);
    }
    return v;
}
inline int blas1_grid(long long n, int vec, int ctas_per_sm = kBlas1CtasPerSm) {
    const long long tile = 256LL * vec * 4;
    long long need = (n + tile - 1) / tile;
    if (need < 1) need = 1;
    const long long cap = (long long)sm_count() * ctas_per_sm;
    return (int)(need < cap ? need : cap);
}
inline bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

// The persistent BLAS-1 kernels (K2 / K2r / K3) split their work statically over sm_count * 2 CTAs, which
// only balances if every SM hosts exactly two of them.  A normal launch spreads the CTAs breadth-first; a
// programmatic (dependent) launch places them as resources free up and stacks three or four on some SMs
// (measured at 20k x 20k: K2r 1.507 -> 1.583 ms; capping the residency with 100 KB of dummy shared
// memory per CTA shrinks L1 and costs more: 1.87 ms).  They are therefore launched the normal way: one
// launch gap per iteration stays, in front of K2 / K2r.
template <typename... KArgs, typename... Args>
cudaError_t launch_plain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl_mode() & 4) ? 1 : 0;  // bit 2 (off by default): A/B switch for the statement above
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

// device view of a reduction context.  grid = CTAs of the producing launch (one partial each).
int make_tail(const b200_reduce_ctx* c, int which, int two, long long grid, TailArgs* t, const char* who) {
    memset(t, 0, sizeof *t);
    t->which = -1;
    if (!c) return fail(B200_EINVAL, "%s: NULL reduction context", who);
    if (!c->d_partials || grid > c->capacity || (two && !c->d_partials_b))
        return fail(B200_EINVAL, "%s: partials buffer missing or too small", who);
    if (which != RED_SUM && !c->d_scalars) return fail(B200_EINVAL, "%s: NULL scalars", who);
    t->partials = c->d_partials; t->partials_b = c->d_partials_b; t->two = two;
    t->sc = static_cast<CGScalars*>(c->d_scalars);
    if (c->phases == 0) return B200_OK;  // partials only: tickets stay NULL
    if (!c->d_group_sums || !c->d_tickets) return fail(B200_EINVAL, "%s: NULL group sums / tickets", who);
    if (c->world < 1 || c->world > B200_MAX_RANKS || c->rank < 0 || c->rank >= c->world)
        return fail(B200_EINVAL, "%s: bad rank / world", who);
    if (c->world > 1 && !c->d_peer_xchg) return fail(B200_EINVAL, "%s: NULL peer table", who);
    if (which == RED_SUM && !c->d_out) return fail(B200_EINVAL, "%s: NULL out", who);
    if (c->phases != 3 && !c->d_stash) return fail(B200_EINVAL, "%s: split phases need a stash", who);
    t->which = which; t->phases = c->phases; t->tol = c->tol;
    t->status = static_cast<CGStatus*>(c->h_status_mapped);
    t->out = c->d_out; t->stash = c->d_stash;
    t->gsum = c->d_group_sums; t->tickets = c->d_tickets;
    t->cap_groups = (int)((c->capacity + B200_RED_GROUP - 1) / B200_RED_GROUP);
    t->rank = c->rank; t->world = c->world;
    if (c->world > 1) {
        for (int r = 0; r < c->world; r++) t->peer_xchg[r] = static_cast<XchgArea*>(c->d_peer_xchg[r]);
        t->my_xchg = t->peer_xchg[c->rank];
    }
    return B200_OK;
}

// the stand-alone reduction over n partials (one warp-sized CTA per group of 256)
int launch_reduce(const TailArgs& t, int n_partials, cudaStream_t s) {
    ReduceArgs a;
    a.t = t;
    a.n_partials = n_partials;
    int grid = (n_partials + B200_RED_GROUP - 1) / B200_RED_GROUP;
    if (grid < 1) grid = 1;
    launch_pdl(cg_reduce_kernel, grid, 32, 0, s, a);
    return check_launch("cg_reduce_kernel");
}

// STENCIL5 producer + its reduction: the kernel writes one partial per CTA, cg_reduce_kernel follows
// (programmatic dependent launch) unless the caller asked for the partials only
template <int MODE>
int launch_stencil_cg(const b200_band* band, Geometry& g, cudaStream_t s) {
    TailArgs t = g.tail;
    int rc = launch_stencil<MODE>(band, g, s);
    if (rc || t.tickets == nullptr) return rc;
    return launch_reduce(t, g.grid, s);
}

int make_push(const b200_halo_push_args* h, long long n, const double* v, const void* scalars, HaloPushArgs* a,
              const char* who) {
    memset(a, 0, sizeof *a);
    if (!h) return B200_OK;
    if (!h->d_my_xchg || h->halo < 1 || n < h->halo) return fail(B200_EINVAL, "%s: bad halo push arguments", who);
    if ((h->d_dst_prev && !h->d_flag_prev) || (h->d_dst_next && !h->d_flag_next)) return fail(B200_EINVAL, "%s: NULL flag", who);
    a->v_local = v; a->n_local = n; a->halo = h->halo; a->dst_prev = h->d_dst_prev; a->dst_next = h->d_dst_next;
    a->flag_prev = h->d_flag_prev; a->flag_next = h->d_flag_next;
    a->my_xchg = static_cast<XchgArea*>(h->d_my_xchg);
    a->sc = static_cast<const CGScalars*>(scalars);
    return B200_OK;
}
}  // namespace

extern "C" void b200_cg_set_pdl(int mode) { g_pdl.store(mode & 7, std::memory_order_relaxed); }

extern "C" int b200_cg_max_partials(const b200_band* band) {
    int n = b200_stencil5_num_partials(band);
    if (n < 0) return n;
    const int blas1 = sm_count() * kBlas1CtasPerSm;
    return n > blas1 ? n : blas1;
}

extern "C" int b200_cg_residual_init(const b200_band* band, const double* d_x, const double* d_b, double* d_r,
                                     double* d_p, const b200_reduce_ctx* ctx, b200_stream stream) {
    Geometry g;
    int rc = build_geometry(band, d_x, &g);
    if (rc) return rc;
    if (!d_b || !d_r || !d_p) return fail(B200_EINVAL, "cg_residual_init: NULL argument");
    choose_cg_kernel(g);
    if ((rc = make_tail(ctx, RED_RR0, 0, g.grid, &g.tail, "cg_residual_init"))) return rc;
    g.a.y = d_r; g.a.y2 = d_p; g.a.b = d_b; g.a.partials = ctx->d_partials;
    g.a.error_word = &static_cast<CGScalars*>(ctx->d_scalars)->error;
    return launch_stencil_cg<ST_RESID>(band, g, (cudaStream_t)stream);
}

extern "C" int b200_cg_spmv_dot(const b200_band* band, const double* d_p, double* d_Ap, const b200_reduce_ctx* ctx,
                                b200_stream stream) {
    Geometry g;
    int rc = build_geometry(band, d_p, &g);
    if (rc) return rc;
    if (!d_Ap) return fail(B200_EINVAL, "cg_spmv_dot: NULL argument");
    choose_cg_kernel(g);
    if ((rc = make_tail(ctx, RED_PAP, 0, g.grid, &g.tail, "cg_spmv_dot"))) return rc;
    g.a.y = d_Ap; g.a.partials = ctx->d_partials;
    CGScalars* sc = static_cast<CGScalars*>(ctx->d_scalars);
    g.a.converged = &sc->converged;
    g.a.error_word = &sc->error;
    return launch_stencil_cg<ST_DOT>(band, g, (cudaStream_t)stream);
}

namespace {
int update_xr(long long n, const double* d_p, const double* d_Ap, double* d_x, double* d_r, const b200_reduce_ctx* ctx,
              int which, b200_stream stream);
}
extern "C" int b200_cg_update_xr(long long n, const double* d_p, const double* d_Ap, double* d_x, double* d_r,
                                 const b200_reduce_ctx* ctx, b200_stream stream) {
    return update_xr(n, d_p, d_Ap, d_x, d_r, ctx, RED_RR, stream);
}
// same pass, but the r.r tail only tests convergence: beta comes from the r.z tail behind the preconditioner solve
extern "C" int b200_pcg_update_xr_stored_z(long long n, const double* d_p, const double* d_Ap, double* d_x, double* d_r,
                                           const b200_reduce_ctx* ctx, b200_stream stream) {
    return update_xr(n, d_p, d_Ap, d_x, d_r, ctx, RED_RRC, stream);
}
namespace {
int update_xr(long long n, const double* d_p, const double* d_Ap, double* d_x, double* d_r, const b200_reduce_ctx* ctx,
              int which, b200_stream stream) {
    if (!d_p || !d_Ap || !d_x || !d_r) return fail(B200_EINVAL, "cg_update_xr: NULL argument");
    const bool v2 = aligned16(d_p) && aligned16(d_Ap) && aligned16(d_x) && aligned16(d_r);
    const int grid = blas1_grid(n, v2 ? 2 : 1, kRrCtasPerSm);
    TailArgs t;
    int rc = make_tail(ctx, which, 0, grid, &t, "cg_update_xr");
    if (rc) return rc;
    const CGScalars* sc = static_cast<const CGScalars*>(ctx->d_scalars);
    if (v2) launch_plain(cg_update_xr_kernel<2>, grid, 256, 0, (cudaStream_t)stream, n, sc, d_p, d_Ap, d_x, d_r, t);
    else launch_plain(cg_update_xr_kernel<1>, grid, 256, 0, (cudaStream_t)stream, n, sc, d_p, d_Ap, d_x, d_r, t);
    return check_launch("cg_update_xr_kernel");
}
}  // namespace

extern "C" int b200_cg_update_p(long long n, const void* d_scalars, const double* d_r, double* d_p,
                                b200_stream stream) {
    if (!d_scalars || !d_r || !d_p) return fail(B200_EINVAL, "cg_update_p: NULL argument");
    const bool v2 = aligned16(d_r) && aligned16(d_p);
    const int grid = blas1_grid(n, v2 ? 2 : 1, kRrCtasPerSm);
    const CGScalars* sc = static_cast<const CGScalars*>(d_scalars);
    if (v2) launch_plain(cg_update_p_kernel<2>, grid, 256, 0, (cudaStream_t)stream, n, sc, d_r, d_p);
    else launch_plain(cg_update_p_kernel<1>, grid, 256, 0, (cudaStream_t)stream, n, sc, d_r, d_p);
    return check_launch("cg_update_p_kernel");
}

extern "C" int b200_cg_update_p_push(long long n, const void* d_scalars, const double* d_r, double* d_p,
                                     const b200_halo_push_args* h, b200_stream stream) {
    if (!d_scalars || !d_r || !d_p || !h) return fail(B200_EINVAL, "cg_update_p_push: bad argument");
    HaloPushArgs a;
    int rc = make_push(h, n, d_p, d_scalars, &a, "cg_update_p_push");
    if (rc) return rc;
    const int grid = blas1_grid(n, 1);
    launch_pdl(cg_update_p_push_kernel, grid, 256, 0, (cudaStream_t)stream, n, a.sc, d_r, d_p, a);
    return check_launch("cg_update_p_push_kernel");
}

extern "C" int b200_cg_spmv_fused_nx(const b200_band* band, const double* d_p_old, const double* const* d_p_older, int nx,
                                     const double* d_r, double* d_p_new, double* d_x, double* d_Ap,
                                     const b200_reduce_ctx* ctx, b200_stream stream) {
    Geometry g;
    int rc = build_geometry(band, d_p_old, &g);
    if (rc) return rc;
    if (!d_r || !d_p_new || !d_Ap || !ctx || !ctx->d_scalars) return fail(B200_EINVAL, "cg_spmv_fused: NULL argument");
    if (nx < 0 || nx > ST_MAX_NX || (nx > 0 && !d_x) || (nx > 1 && !d_p_older)) return fail(B200_EINVAL, "cg_spmv_fused: bad x retirement arguments");
    if (d_p_new == d_p_old) return fail(B200_EINVAL, "cg_spmv_fused: p_new must not alias p_old");
    for (int k = 0; k + 1 < nx; k++)
        if (!d_p_older[k] || d_p_older[k] == d_p_new) return fail(B200_EINVAL, "cg_spmv_fused: an older direction is NULL or aliases p_new");
    choose_cg_kernel(g, nx != 1);  // the ring kernel has no registers for an extra direction stream
    if ((rc = make_tail(ctx, RED_PAP, 0, g.grid, &g.tail, "cg_spmv_fused"))) return rc;
    const CGScalars* sc = static_cast<const CGScalars*>(ctx->d_scalars);
    g.a.y = d_Ap; g.a.y2 = d_p_new; g.a.r = d_r; g.a.xs = d_x; g.a.partials = ctx->d_partials;
    g.a.ab = &sc->alpha;
    static_assert(offsetof(CGScalars, beta) == offsetof(CGScalars, alpha) + sizeof(double), "alpha, beta adjacent");
    for (int k = 0; k + 1 < nx; k++) g.a.xp[k] = d_p_older[k];
    g.a.alpha_hist = sc->alpha_hist;
    g.a.iter_ptr = &sc->iterations;
    g.a.converged = &sc->converged;
    g.a.error_word = &const_cast<CGScalars*>(sc)->error;
    cudaStream_t s = (cudaStream_t)stream;
    switch (nx) {
        case 0: return launch_stencil_cg<ST_FUSED_X0>(band, g, s);
        case 1: return launch_stencil_cg<ST_FUSED>(band, g, s);
        case 2: return launch_stencil_cg<ST_FUSED_X2>(band, g, s);
        case 3: return launch_stencil_cg<ST_FUSED_X3>(band, g, s);
        default: return launch_stencil_cg<ST_FUSED_X4>(band, g, s);
    }
}
extern "C" int b200_cg_spmv_fused(const b200_band* band, const double* d_p_old, const double* d_r, double* d_p_new,
                                  double* d_x, double* d_Ap, const b200_reduce_ctx* ctx, b200_stream stream) {
    if (!d_x) return fail(B200_EINVAL, "cg_spmv_fused: NULL argument");
    return b200_cg_spmv_fused_nx(band, d_p_old, nullptr, 1, d_r, d_p_new, d_x, d_Ap, ctx, stream);
}

extern "C" int b200_cg_update_r(long long n, const double* d_Ap, double* d_r, const b200_halo_push_args* push,
                                const b200_reduce_ctx* ctx, b200_stream stream) {
    if (!d_Ap || !d_r) return fail(B200_EINVAL, "cg_update_r: NULL argument");
    const bool v2 = aligned16(d_Ap) && aligned16(d_r);
    const int grid = blas1_grid(n, v2 ? 2 : 1, kRrCtasPerSm);
    TailArgs t;
    int rc = make_tail(ctx, RED_RR, 0, grid, &t, "cg_update_r");
    if (rc) return rc;
    HaloPushArgs h;
    if ((rc = make_push(push, n, d_r, ctx->d_scalars, &h, "cg_update_r"))) return rc;
    if (push) {
        if (!t.tickets) return fail(B200_EINVAL, "cg_update_r: the halo push needs a fused reduction (phases != 0)");
        t.publish = 1; t.flag_prev = h.dst_prev ? h.flag_prev : nullptr; t.flag_next = h.dst_next ? h.flag_next : nullptr;
        t.my_xchg = h.my_xchg;
    }
    const CGScalars* sc = static_cast<const CGScalars*>(ctx->d_scalars);
    cudaStream_t s = (cudaStream_t)stream;
    if (push) {
        if (v2) launch_plain(cg_update_r_kernel<2, true>, grid, 256, 0, s, n, sc, d_Ap, d_r, h, t);
        else launch_plain(cg_update_r_kernel<1, true>, grid, 256, 0, s, n, sc, d_Ap, d_r, h, t);
    } else {
        if (v2) launch_plain(cg_update_r_kernel<2, false>, grid, 256, 0, s, n, sc, d_Ap, d_r, h, t);
        else launch_plain(cg_update_r_kernel<1, false>, grid, 256, 0, s, n, sc, d_Ap, d_r, h, t);
    }
    return check_launch("cg_update_r_kernel");
}

extern "C" int b200_cg_halo_dir(const double* d_r_prev, const double* d_r_next, const double* d_pold_prev,
                                const double* d_pold_next, double* d_pnew_prev, double* d_pnew_next, int halo,
                                const uint32_t* d_flag_prev, const uint32_t* d_flag_next, const void* d_my_xchg,
                                void* d_scalars, int beta_zero, b200_stream stream) {
    if (!d_scalars || halo < 1 || !d_my_xchg) return fail(B200_EINVAL, "cg_halo_dir: bad argument");
    if ((d_r_prev && (!d_pnew_prev || !d_flag_prev || (!beta_zero && !d_pold_prev))) ||
        (d_r_next && (!d_pnew_next || !d_flag_next || (!beta_zero && !d_pold_next))))
        return fail(B200_EINVAL, "cg_halo_dir: NULL buffer");
    if (!d_r_prev && !d_r_next) return B200_OK;
    HaloDirArgs a;
    memset(&a, 0, sizeof a);
    a.r_prev = d_r_prev; a.r_next = d_r_next; a.pold_prev = d_pold_prev; a.pold_next = d_pold_next;
    a.pnew_prev = d_pnew_prev; a.pnew_next = d_pnew_next; a.halo = halo;
    a.flag_prev = d_flag_prev; a.flag_next = d_flag_next;
    a.seq_ptr = &static_cast<const XchgArea*>(d_my_xchg)->halo_seq;
    a.sc = static_cast<CGScalars*>(d_scalars); a.beta_zero = beta_zero;
    int grid = (halo + 255) / 256;
    if (grid > 64) grid = 64;
    launch_pdl(cg_halo_dir_kernel, grid, 256, 0, (cudaStream_t)stream, a);
    return check_launch("cg_halo_dir_kernel");
}

extern "C" int b200_cg_finish_x_depth(long long n, const void* d_scalars, const double* const* d_pbuf, int nbuf, int depth,
                                      int only_if_converged, double* d_x, b200_stream stream) {
    if (!d_scalars || !d_pbuf || !d_x) return fail(B200_EINVAL, "cg_finish_x: NULL argument");
    if (depth < 1 || depth > ST_MAX_NX || nbuf < 1 || nbuf > 5 || (nbuf < depth + 1 && !(depth == 1 && nbuf == 1)))
        return fail(B200_EINVAL, "cg_finish_x: depth d needs d + 1 direction buffers");
    FinishXArgs a;
    memset(&a, 0, sizeof a);
    for (int k = 0; k < nbuf; k++) {
        if (!d_pbuf[k]) return fail(B200_EINVAL, "cg_finish_x: NULL direction buffer");
        a.pbuf[k] = d_pbuf[k];
    }
    a.nbuf = nbuf; a.depth = depth; a.only_if_converged = only_if_converged;
    const int grid = blas1_grid(n, 1);
    launch_pdl(cg_finish_x_kernel, grid, 256, 0, (cudaStream_t)stream, n, static_cast<const CGScalars*>(d_scalars), a, d_x);
    return check_launch("cg_finish_x_kernel");
}
extern "C" int b200_cg_finish_x(long long n, const void* d_scalars, const double* d_p0, const double* d_p1,
                                double* d_x, int only_if_converged, b200_stream stream) {
    if (!d_p0 || !d_p1) return fail(B200_EINVAL, "cg_finish_x: NULL argument");
    const double* bufs[2] = {d_p0, d_p1};
    return b200_cg_finish_x_depth(n, d_scalars, bufs, d_p0 == d_p1 ? 1 : 2, 1, only_if_converged, d_x, stream);
}

extern "C" int b200_cg_update_px_nx(long long n, const void* d_scalars, const double* d_r, const double* d_p_old,
                                    const double* const* d_p_older, int nx, double* d_p_new, double* d_x, b200_stream stream) {
    if (!d_scalars || !d_r || !d_p_old || !d_p_new || (nx > 0 && !d_x) || (nx > 1 && !d_p_older))
        return fail(B200_EINVAL, "cg_update_px_nx: NULL argument");
    if (nx < 0 || nx > ST_MAX_NX || d_p_new == d_p_old) return fail(B200_EINVAL, "cg_update_px_nx: bad x retirement arguments");
    UpdatePxArgs a;
    memset(&a, 0, sizeof a);
    for (int k = 0; k + 1 < nx; k++) {
        if (!d_p_older[k] || d_p_older[k] == d_p_new) return fail(B200_EINVAL, "cg_update_px_nx: an older direction is NULL or aliases p_new");
        a.older[k] = d_p_older[k];
    }
    const int grid = blas1_grid(n, 1, kRrCtasPerSm);
    const CGScalars* sc = static_cast<const CGScalars*>(d_scalars);
    cudaStream_t s = (cudaStream_t)stream;
    switch (nx) {
        case 0: launch_plain(cg_update_px_depth_kernel<0>, grid, 256, 0, s, n, sc, d_r, d_p_old, d_p_new, d_x, a); break;
        case 1: launch_plain(cg_update_px_depth_kernel<1>, grid, 256, 0, s, n, sc, d_r, d_p_old, d_p_new, d_x, a); break;
        case 2: launch_plain(cg_update_px_depth_kernel<2>, grid, 256, 0, s, n, sc, d_r, d_p_old, d_p_new, d_x, a); break;
        case 3: launch_plain(cg_update_px_depth_kernel<3>, grid, 256, 0, s, n, sc, d_r, d_p_old, d_p_new, d_x, a); break;
        default: launch_plain(cg_update_px_depth_kernel<4>, grid, 256, 0, s, n, sc, d_r, d_p_old, d_p_new, d_x, a); break;
    }
    return check_launch("cg_update_px_depth_kernel");
}

extern "C" int b200_cg_update_px(long long n, const void* d_scalars, const double* d_r, double* d_p, double* d_x,
                                 b200_stream stream) {
    if (!d_scalars || !d_r || !d_p || !d_x) return fail(B200_EINVAL, "cg_update_px: NULL argument");
    const bool v2 = aligned16(d_r) && aligned16(d_p) && aligned16(d_x);
    const int grid = blas1_grid(n, v2 ? 2 : 1, kRrCtasPerSm);
    const CGScalars* sc = static_cast<const CGScalars*>(d_scalars);
    if (v2) launch_plain(cg_update_px_kernel<2>, grid, 256, 0, (cudaStream_t)stream, n, sc, d_r, d_p, d_x);
    else launch_plain(cg_update_px_kernel<1>, grid, 256, 0, (cudaStream_t)stream, n, sc, d_r, d_p, d_x);
    return check_launch("cg_update_px_kernel");
}

extern "C" int b200_cg_reduce(const b200_reduce_ctx* ctx, int which, int n_partials, int two_sums, int phases,
                              b200_stream stream) {
    if (!ctx || n_partials < 0) return fail(B200_EINVAL, "cg_reduce: bad argument");
    if (phases < 1 || phases > 3) return fail(B200_EINVAL, "cg_reduce: phases must be 1, 2 or 3");
    b200_reduce_ctx c = *ctx;
    c.phases = phases;
    if (n_partials == 0 && !c.d_partials) c.d_partials = c.d_stash;  // exchange without data (barrier)
    TailArgs t;
    int rc = make_tail(&c, which, two_sums, n_partials, &t, "cg_reduce");
    if (rc) return rc;
    return launch_reduce(t, n_partials, (cudaStream_t)stream);
}

extern "C" int b200_cg_read_tail_times(const void* d_scalars, b200_cg_tail_times* h_out, b200_stream stream) {
    if (!d_scalars || !h_out) return fail(B200_EINVAL, "cg_read_tail_times: NULL argument");
    CGScalars h;
    cudaError_t e = cudaMemcpyAsync(&h, d_scalars, sizeof h, cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
    if (e != cudaSuccess) return fail(B200_ECUDA, "cg_read_tail_times: %s", cudaGetErrorString(e));
    for (int k = 0; k < 8; k++) { h_out->ns[k] = h.tail_ns[k]; h_out->count[k] = h.tail_cnt[k]; h_out->gap_ns[k] = h.gap_ns[k]; }
    h_out->error = h.error;
    return B200_OK;
}

// ---- Jacobi-preconditioned CG ----
extern "C" int b200_pcg_diag_inv(const int* d_row_ptr, const int* d_col_idx, const double* d_values, long long n_local,
                                 long long row_offset, int ell_width, double* d_dinv, int* d_err, b200_stream stream) {
    if (!d_col_idx || !d_values || !d_dinv || !d_err || n_local < 0) return fail(B200_EINVAL, "pcg_diag_inv: bad argument");
    if (!d_row_ptr && ell_width < 1) return fail(B200_EINVAL, "pcg_diag_inv: ELLPACK needs a width");
    if (n_local == 0) return B200_OK;
    pcg_diag_inv_kernel<<<blas1_grid(n_local, 1), 256, 0, (cudaStream_t)stream>>>(n_local, row_offset, d_row_ptr, ell_width,
                                                                                 d_col_idx, d_values, d_dinv, d_err);
    return check_launch("pcg_diag_inv_kernel");
}

extern "C" int b200_pcg_init(long long n, const double* d_r, const double* d_dinv, const double* d_z, double* d_p,
                             const b200_reduce_ctx* ctx, b200_stream stream) {
    if (!d_r || (!d_dinv && !d_z) || !d_p) return fail(B200_EINVAL, "pcg_init: NULL argument");
    const int grid = blas1_grid(n, 1);
    TailArgs t;
    int rc = make_tail(ctx, RED_RZ0, 0, grid, &t, "pcg_init");
    if (rc) return rc;
    pcg_init_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n, d_r, d_dinv, d_z, d_p, t);
    return check_launch("pcg_init_kernel");
}

extern "C" int b200_pcg_update_xr(long long n, const double* d_p, const double* d_Ap, const double* d_dinv, double* d_x,
                                  double* d_r, const b200_reduce_ctx* ctx, b200_stream stream) {
    if (!d_p || !d_Ap || !d_dinv || !d_x || !d_r) return fail(B200_EINVAL, "pcg_update_xr: NULL argument");
    const int grid = blas1_grid(n, 1);
    TailArgs t;
    int rc = make_tail(ctx, RED_PCG, 1, grid, &t, "pcg_update_xr");
    if (rc) return rc;
    pcg_update_xr_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n, static_cast<const CGScalars*>(ctx->d_scalars), d_p, d_Ap,
                                                                 d_dinv, d_x, d_r, t);
    return check_launch("pcg_update_xr_kernel");
}

extern "C" int b200_pcg_update_p(long long n, const void* d_scalars, const double* d_r, const double* d_dinv, double* d_p,
                                 const b200_halo_push_args* push, b200_stream stream) {
    if (!d_scalars || !d_r || !d_p) return fail(B200_EINVAL, "pcg_update_p: NULL argument");  // d_dinv NULL: d_r holds z
    HaloPushArgs h;
    int rc = make_push(push, n, d_p, d_scalars, &h, "pcg_update_p");
    if (rc) return rc;
    pcg_update_p_kernel<<<blas1_grid(n, 1), 256, 0, (cudaStream_t)stream>>>(n, static_cast<const CGScalars*>(d_scalars), d_r,
                                                                            d_dinv, d_p, h);
    return check_launch("pcg_update_p_kernel");
}

namespace {
int make_line_blocks(long long n_local, long long row_offset, int grid_size, LineBlocks* lb) {
    if (n_local < 1 || row_offset < 0 || grid_size < 1) return fail(B200_EINVAL, "block-Jacobi: bad band");
    lb->row_offset = row_offset; lb->n_local = n_local; lb->n = grid_size;
    lb->first_grid_row = row_offset / grid_size;
    const long long last = (row_offset + n_local - 1) / grid_size;
    lb->n_blocks = (int)(last - lb->first_grid_row + 1);
    return B200_OK;
}
}  // namespace

extern "C" int b200_bj_factor(const int* d_row_ptr, const int* d_col_idx, const double* d_values, long long n_local,
                              long long row_offset, int grid_size, int ell_width, double* d_m, double* d_invd, double* d_c,
                              int* d_err, b200_stream stream) {
    if (!d_col_idx || !d_values || !d_m || !d_invd || !d_c || !d_err) return fail(B200_EINVAL, "bj_factor: NULL argument");
    if (!d_row_ptr && ell_width < 1) return fail(B200_EINVAL, "bj_factor: ELLPACK needs a width");
    LineBlocks lb;
    int rc = make_line_blocks(n_local, row_offset, grid_size, &lb);
    if (rc) return rc;
    bj_factor_kernel<<<(lb.n_blocks + 127) / 128, 128, 0, (cudaStream_t)stream>>>(lb, d_row_ptr, ell_width, d_col_idx, d_values,
                                                                                 d_m, d_invd, d_c, d_err);
    return check_launch("bj_factor_kernel");
}

extern "C" int b200_bj_solve(long long n_local, long long row_offset, int grid_size, const void* d_scalars, const double* d_m,
                             const double* d_invd, const double* d_c, const double* d_r, double* d_z, b200_stream stream) {
    if (!d_m || !d_invd || !d_c || !d_r || !d_z) return fail(B200_EINVAL, "bj_solve: NULL argument");
    LineBlocks lb;
    int rc = make_line_blocks(n_local, row_offset, grid_size, &lb);
    if (rc) return rc;
    bj_solve_kernel<<<(lb.n_blocks + 127) / 128, 128, 0, (cudaStream_t)stream>>>(lb, static_cast<const CGScalars*>(d_scalars), d_m,
                                                                                d_invd, d_c, d_r, d_z);
    return check_launch("bj_solve_kernel");
}

extern "C" int b200_dot_partials(long long n, const double* d_x, const double* d_y, int which, const b200_reduce_ctx* ctx,
                                 b200_stream stream) {
    if (!d_x || !d_y) return fail(B200_EINVAL, "dot_partials: NULL argument");
    if (which != RED_PAP && which != RED_RZ && which != RED_RZ0 && which != RED_SUM) return fail(B200_EINVAL, "dot_partials: bad `which`");
    const int grid = blas1_grid(n, 1);
    TailArgs t;
    int rc = make_tail(ctx, which, 0, grid, &t, "dot_partials");
    if (rc) return rc;
    dot_partials_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n, static_cast<const CGScalars*>(ctx->d_scalars), d_x, d_y, t);
    return check_launch("dot_partials_kernel");
}

extern "C" int b200_residual_init_generic(long long n, const double* d_b, const double* d_Ap, double* d_r,
                                          double* d_p, const b200_reduce_ctx* ctx, b200_stream stream) {
    if (!d_b || !d_Ap || !d_r || !d_p) return fail(B200_EINVAL, "residual_init: NULL argument");
    const int grid = blas1_grid(n, 1);
    TailArgs t;
    int rc = make_tail(ctx, RED_RR0, 0, grid, &t, "residual_init");
    if (rc) return rc;
    residual_init_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n, d_b, d_Ap, d_r, d_p, t);
    return check_launch("residual_init_kernel");
}

extern "C" int b200_checksum(long long n, const double* d_x, const b200_reduce_ctx* ctx, b200_stream stream) {
    if (!d_x) return fail(B200_EINVAL, "checksum: NULL argument");
    const int grid = blas1_grid(n, 1);
    TailArgs t;
    int rc = make_tail(ctx, RED_SUM, 1, grid, &t, "checksum");
    if (rc) return rc;
    checksum_partials_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n, d_x, t);
    return check_launch("checksum_partials_kernel");
}

extern "C" int b200_halo_push(const double* d_v_local, long long n_local, const b200_halo_push_args* h,
                              const void* d_scalars, b200_stream stream) {
    if (!d_v_local || !h) return fail(B200_EINVAL, "halo_push: bad argument");
    HaloPushArgs a;
    int rc = make_push(h, n_local, d_v_local, d_scalars, &a, "halo_push");
    if (rc) return rc;
    const int per_dir = 8;
    halo_push_kernel<<<2 * per_dir, 256, 0, (cudaStream_t)stream>>>(a);
    return check_launch("halo_push_kernel");
}

// ------------------------------------------------------------------------------------------------
// device-side matrix construction
// ------------------------------------------------------------------------------------------------
extern "C" long long b200_stencil5_nnz_before(long long row, long long grid_size) {
    return stencil5_nnz_before(row, grid_size);
}

static unsigned gen_grid(long long n) {
    long long b = (n + 255) / 256;
    if (b < 1) b = 1;
    if (b > 148LL * 32) b = 148LL * 32;
    return (unsigned)b;
}

extern "C" int b200_gen_stencil5_csr(int grid_size, long long row_offset, long long n_local, double center,
                                     double neighbour, int* d_row_ptr, int* d_col_idx, double* d_values,
                                     b200_stream stream) {
    if (!d_row_ptr || !d_col_idx || !d_values || grid_size < 1) return fail(B200_EINVAL, "gen csr: bad argument");
    // CSR column ids are stored modulo 2^32 and read back as unsigned (csrc/stencil5.cuh): the grid may
    // exceed 2^31 rows as long as it stays below 2^32 and the band below 2^31 non-zeros
    if ((long long)grid_size * grid_size >= 4294967295LL) return fail(B200_EINVAL, "gen csr: column ids exceed 32 bits");
    if (n_local < 0 || 5 * n_local >= 2147483647LL) return fail(B200_EINVAL, "gen csr: band exceeds 2^31 non-zeros");
    gen_stencil5_csr_kernel<<<gen_grid(n_local + 1), 256, 0, (cudaStream_t)stream>>>(grid_size, row_offset, n_local, center,
                                                                                    neighbour, d_row_ptr, d_col_idx, d_values);
    return check_launch("gen_stencil5_csr_kernel");
}

extern "C" int b200_gen_stencil5_ellpack(int grid_size, long long row_offset, long long n_local, double center,
                                         double neighbour, int* d_indices, double* d_values, b200_stream stream) {
    if (!d_indices || !d_values || grid_size < 1) return fail(B200_EINVAL, "gen ell: bad argument");
    gen_stencil5_ell_kernel<<<gen_grid(n_local), 256, 0, (cudaStream_t)stream>>>(grid_size, row_offset, n_local, center,
                                                                                neighbour, d_indices, d_values);
    return check_launch("gen_stencil5_ell_kernel");
}

extern "C" int b200_gen_stencil5_entries(int grid_size, long long row_offset, long long n_local, double center,
                                         double neighbour, void* d_entries, b200_stream stream) {
    if (!d_entries || grid_size < 1) return fail(B200_EINVAL, "gen entries: bad argument");
    gen_stencil5_entries_kernel<<<gen_grid(n_local), 256, 0, (cudaStream_t)stream>>>(
        grid_size, row_offset, n_local, center, neighbour, static_cast<EntryPOD*>(d_entries));
    return check_launch("gen_stencil5_entries_kernel");
}

extern "C" int b200_fill(double* d_p, long long n, double value, b200_stream stream) {
    if (!d_p && n > 0) return fail(B200_EINVAL, "fill: NULL");
    if (n <= 0) return B200_OK;
    fill_kernel<<<gen_grid(n), 256, 0, (cudaStream_t)stream>>>(d_p, n, value);
    return check_launch("fill_kernel");
}

// ------------------------------------------------------------------------------------------------
// device-side ingest: Matrix Market entry text -> COO, COO -> CSR  
// ------------------------------------------------------------------------------------------------
namespace {
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 8); }
    template <class T> T* as() { return static_cast<T*>(p); }
};

// exclusive scan of n ints/long longs into d_out (long long); *d_total (device) receives the sum
template <typename T>
int exclusive_scan(const T* d_in, long long* d_out, long long n, long long* d_total, cudaStream_t s) {
    const long long tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    DevBuf sums;
    if (sums.alloc((size_t)(tiles > 0 ? tiles : 1) * sizeof(long long)) != cudaSuccess) return fail(B200_ENOMEM, "scan: cudaMalloc");
    if (tiles > 0) {
        scan_tiles_kernel<T><<<(unsigned)tiles, 256, 0, s>>>(d_in, d_out, n, sums.as<long long>());
        int rc = check_launch("scan_tiles_kernel");
        if (rc) return rc;
    }
    scan_tile_sums_kernel<<<1, 1024, 0, s>>>(sums.as<long long>(), tiles, d_total);
    int rc = check_launch("scan_tile_sums_kernel");
    if (rc) return rc;
    if (tiles > 0) {
        scan_add_offsets_kernel<<<(unsigned)tiles, 256, 0, s>>>(d_out, n, sums.as<long long>());
        rc = check_launch("scan_add_offsets_kernel");
        if (rc) return rc;
    }
    cudaError_t e = cudaStreamSynchronize(s);  // sums is freed on return
    if (e != cudaSuccess) return fail(B200_ECUDA, "scan: %s", cudaGetErrorString(e));
    return B200_OK;
}
}  // namespace

extern "C" int b200_patch_entry_values(void* d_entries, const long long* h_pairs, int n, b200_stream stream) {
    if (n <= 0) return B200_OK;
    if (!d_entries || !h_pairs) return fail(B200_EINVAL, "patch_entry_values: NULL argument");
    cudaStream_t s = (cudaStream_t)stream;
    DevBuf list;
    if (list.alloc((size_t)n * 16) != cudaSuccess) { cudaGetLastError(); return fail(B200_ENOMEM, "patch_entry_values: cudaMalloc"); }
    cudaError_t e = cudaMemcpyAsync(list.p, h_pairs, (size_t)n * 16, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return fail(B200_ECUDA, "patch_entry_values: %s", cudaGetErrorString(e));
    patch_entry_values_kernel<<<(n + 255) / 256, 256, 0, s>>>(static_cast<EntryPOD*>(d_entries), list.as<long long>(), n);
    int rc = check_launch("patch_entry_values_kernel");
    if (rc) return rc;
    e = cudaStreamSynchronize(s);  // list is freed on return
    return e == cudaSuccess ? B200_OK : fail(B200_ECUDA, "patch_entry_values: %s", cudaGetErrorString(e));
}

extern "C" int b200_coo_to_csr(const void* d_entries, long long nnz, int rows, int cols, int* d_row_ptr, int* d_col_idx,
                               double* d_values, b200_stream stream) {
    if ((!d_entries && nnz > 0) || !d_row_ptr || rows < 0 || nnz < 0 || nnz > 2147483647LL)
        return fail(B200_EINVAL, "coo_to_csr: bad argument");
    if (nnz > 0 && (!d_col_idx || !d_values)) return fail(B200_EINVAL, "coo_to_csr: NULL output");
    cudaStream_t s = (cudaStream_t)stream;
    const EntryPOD* e = static_cast<const EntryPOD*>(d_entries);
    DevBuf counts, cursor, scan, pos, ctmp, vtmp, misc;
    if (counts.alloc((size_t)rows * 4) != cudaSuccess || cursor.alloc((size_t)rows * 4) != cudaSuccess ||
        scan.alloc((size_t)rows * 8) != cudaSuccess || pos.alloc((size_t)nnz * 4) != cudaSuccess ||
        ctmp.alloc((size_t)nnz * 4) != cudaSuccess || vtmp.alloc((size_t)nnz * 8) != cudaSuccess ||
        misc.alloc(16) != cudaSuccess) {
        cudaGetLastError();
        return fail(B200_ENOMEM, "coo_to_csr: cudaMalloc");
    }
    cudaMemsetAsync(counts.p, 0, (size_t)rows * 4, s);
    cudaMemsetAsync(misc.p, 0, 16, s);
    int* bad = misc.as<int>() + 2;
    long long* total = misc.as<long long>();
    int rc;
    if (nnz > 0) {
        coo_count_rows_kernel<<<gen_grid(nnz), 256, 0, s>>>(e, nnz, rows, cols, counts.as<int>(), bad);
        if ((rc = check_launch("coo_count_rows_kernel"))) return rc;
    }
    if ((rc = exclusive_scan<int>(counts.as<int>(), scan.as<long long>(), rows, total, s))) return rc;
    int h_bad = 0;
    cudaMemcpyAsync(&h_bad, bad, 4, cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    if (h_bad == 1) return fail(B200_EINVAL, "coo_to_csr: entry with a row index outside [0, rows)");
    if (h_bad) return fail(B200_EINVAL, "coo_to_csr: entry with a column index outside [0, cols)");
    coo_finish_row_ptr_kernel<<<gen_grid(rows + 1), 256, 0, s>>>(scan.as<long long>(), nnz, rows, d_row_ptr, cursor.as<int>());
    if ((rc = check_launch("coo_finish_row_ptr_kernel"))) return rc;
    if (nnz > 0) {
        coo_scatter_kernel<<<gen_grid(nnz), 256, 0, s>>>(e, nnz, rows, cursor.as<int>(), d_col_idx, d_values, pos.as<int>());
        if ((rc = check_launch("coo_scatter_kernel"))) return rc;
        const long long warps = ((long long)rows + 31) / 32;
        const long long blocks = (warps * 32 + 255) / 256;
        csr_sort_rows_kernel<48><<<(unsigned)blocks, 256, 0, s>>>(d_row_ptr, rows, d_col_idx, d_values, pos.as<int>(),
                                                               ctmp.as<int>(), vtmp.as<double>());
        if ((rc = check_launch("csr_sort_rows_kernel"))) return rc;
    }
    cudaError_t err = cudaStreamSynchronize(s);  // temporaries are freed on return
    if (err != cudaSuccess) return fail(B200_ECUDA, "coo_to_csr: %s", cudaGetErrorString(err));
    return B200_OK;
}

extern "C" int b200_parse_mtx_entries(const void* d_text, long long n_bytes, long long max_entries, void* d_entries,
                                      long long* n_lines_out, int* n_inexact_out, long long* h_inexact_pairs,
                                      int inexact_cap, int* malformed_out, b200_stream stream) {
    if (!d_text || n_bytes < 0 || !d_entries || !n_lines_out || !n_inexact_out || !malformed_out || inexact_cap < 0)
        return fail(B200_EINVAL, "parse_mtx: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    const long long chunks = (n_bytes + PARSE_CHUNK - 1) / PARSE_CHUNK;
    DevBuf counts, first, misc, list;
    if (counts.alloc((size_t)chunks * 4) != cudaSuccess || first.alloc((size_t)chunks * 8) != cudaSuccess ||
        misc.alloc(32) != cudaSuccess || list.alloc((size_t)inexact_cap * 16) != cudaSuccess) {
        cudaGetLastError();
        return fail(B200_ENOMEM, "parse_mtx: cudaMalloc");
    }
    cudaMemsetAsync(misc.p, 0, 32, s);
    long long* total = misc.as<long long>();
    int* n_inexact = misc.as<int>() + 2;
    int* bad = misc.as<int>() + 3;
    const unsigned char* t = static_cast<const unsigned char*>(d_text);
    int rc;
    if (chunks > 0) {
        mtx_count_lines_kernel<<<(unsigned)((chunks + 127) / 128), 128, 0, s>>>(t, n_bytes, counts.as<int>());
        if ((rc = check_launch("mtx_count_lines_kernel"))) return rc;
    }
    if ((rc = exclusive_scan<int>(counts.as<int>(), first.as<long long>(), chunks, total, s))) return rc;
    if (chunks > 0) {
        mtx_parse_lines_kernel<<<(unsigned)((chunks + 127) / 128), 128, 0, s>>>(
            t, n_bytes, first.as<long long>(), max_entries, static_cast<EntryPOD*>(d_entries), n_inexact,
            list.as<long long>(), inexact_cap, bad);
        if ((rc = check_launch("mtx_parse_lines_kernel"))) return rc;
    }
    long long h_total = 0;
    int h_misc[2] = {0, 0};
    cudaMemcpyAsync(&h_total, total, 8, cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(h_misc, n_inexact, 8, cudaMemcpyDeviceToHost, s);
    cudaError_t err = cudaStreamSynchronize(s);
    if (err != cudaSuccess) return fail(B200_ECUDA, "parse_mtx: %s", cudaGetErrorString(err));
    *n_lines_out = h_total;
    *n_inexact_out = h_misc[0];
    *malformed_out = h_misc[1];
    const int ncopy = h_misc[0] < inexact_cap ? h_misc[0] : inexact_cap;
    if (ncopy > 0 && h_inexact_pairs)
        cudaMemcpy(h_inexact_pairs, list.p, (size_t)ncopy * 16, cudaMemcpyDeviceToHost);
    return B200_OK;
}
