// b200_abi.cu -- extern "C" launchers: the only translation unit that includes kernel headers.
// See include/b200_kernels.h for the contract of every entry point.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <atomic>

#include "../../include/b200_kernels.h"
#include "cg_kernels.cuh"
#include "csr_ell.cuh"
#include "generate.cuh"
#include "ingest.cuh"
#include "stencil5.cuh"
#include "stencil_layout.h"

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

// ------------------------------------------------------------------------------------------------
// STENCIL5
// ------------------------------------------------------------------------------------------------
namespace {

struct Variant {
    int cols, warps, stages;
    const char* info;
};
// keep in sync with the dispatch switch below
const Variant kVariants[] = {
    {4, 4, 2, "v0 (default): 128-col strips (4 cols/lane), 4 warps/CTA, 2-stage ring"},
    {1, 4, 4, "v1: 32-col strips (1 col/lane), 4 warps/CTA, 4-stage ring"},
    {2, 8, 4, "v2: 64-col strips, 8 warps/CTA, 4-stage ring"},
    {4, 4, 3, "v3: 128-col strips (4 cols/lane), 4 warps/CTA, 3-stage ring"},
    {2, 4, 3, "v4: 64-col strips, 4 warps/CTA, 3-stage ring"},
    {2, 2, 4, "v5: 64-col strips, 2 warps/CTA, 4-stage ring"},
    {4, 2, 3, "v6: 128-col strips, 2 warps/CTA, 3-stage ring"},
    {2, 4, 4, "v7: 64-col strips, 4 warps/CTA, 4-stage ring"},
    {2, 1, 4, "v8: 64-col strips, 1 warp/CTA, 4-stage ring"},
    {2, 2, 3, "v9: 64-col strips (2 cols/lane), 2 warps/CTA, 3-stage ring (round-1 default until the K1F sweep)"},
    {4, 1, 3, "v10: 128-col strips, 1 warp/CTA, 3-stage ring"},
    {2, 4, 2, "v11: 64-col strips, 4 warps/CTA, 2-stage ring"},
    {4, 8, 2, "v12: 128-col strips, 8 warps/CTA, 2-stage ring"},
    {4, 2, 2, "v13: 128-col strips, 2 warps/CTA, 2-stage ring"},
};
const int kNumVariants = (int)(sizeof(kVariants) / sizeof(kVariants[0]));
const int kDefaultRowsPerItem = 8;

struct Geometry {
    Stencil5Args a;
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
    const int v = (b->variant >= 0 && b->variant < kNumVariants) ? b->variant : 0;
    const Variant& V = kVariants[v];
    Stencil5Args& a = g->a;
    memset(&a, 0, sizeof a);
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
    g->threads = V.warps * 32;
    const int nb = (a.n_boundary_rows + g->threads - 1) / g->threads;
    g->grid = a.n_interior_ctas + nb;
    g->smem = (size_t)V.warps * V.stages * (5 * W + 2) * 8 + (size_t)V.warps * V.stages * 8;
    return B200_OK;
}

template <int MODE, int COLS, int WARPS, int STAGES, bool CG>
int launch_one(const Geometry& g, cudaStream_t s) {
    auto k = stencil5_kernel<MODE, COLS, WARPS, STAGES, CG>;
    static bool attr_set = false;  // per instantiation
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) {
            cudaGetLastError();
            snprintf(g_err, sizeof g_err, "stencil5: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return (e == cudaErrorNoKernelImageForDevice || e == cudaErrorInvalidDeviceFunction) ? B200_ENODEV
                                                                                                 : B200_ECUDA;
        }
        attr_set = true;
    }
    k<<<g.grid, g.threads, g.smem, s>>>(g.a);
    return check_launch("stencil5_kernel");
}

template <int MODE, bool CG>
int launch_variant(int v, const Geometry& g, cudaStream_t s) {
    switch (v) {
        case 1: return launch_one<MODE, 1, 4, 4, CG>(g, s);
        case 2: return launch_one<MODE, 2, 8, 4, CG>(g, s);
        case 3: return launch_one<MODE, 4, 4, 3, CG>(g, s);
        case 4: return launch_one<MODE, 2, 4, 3, CG>(g, s);
        case 5: return launch_one<MODE, 2, 2, 4, CG>(g, s);
        case 6: return launch_one<MODE, 4, 2, 3, CG>(g, s);
        case 7: return launch_one<MODE, 2, 4, 4, CG>(g, s);
        case 8: return launch_one<MODE, 2, 1, 4, CG>(g, s);
        case 9: return launch_one<MODE, 2, 2, 3, CG>(g, s);
        case 10: return launch_one<MODE, 4, 1, 3, CG>(g, s);
        case 11: return launch_one<MODE, 2, 4, 2, CG>(g, s);
        case 12: return launch_one<MODE, 4, 8, 2, CG>(g, s);
        case 13: return launch_one<MODE, 4, 2, 2, CG>(g, s);
        default: return launch_one<MODE, 4, 4, 2, CG>(g, s);
    }
}

template <int MODE>
int launch_stencil(const b200_band* b, Geometry& g, cudaStream_t s) {
    if (g.grid == 0) return B200_OK;
    const int v = (b->variant >= 0 && b->variant < kNumVariants) ? b->variant : 0;
    // peer-written halos must be read through L2 (ld.global.cg); single-GPU uses the read-only path
    const bool cg = (b->d_halo_prev != nullptr || b->d_halo_next != nullptr);
    return cg ? launch_variant<MODE, true>(v, g, s) : launch_variant<MODE, false>(v, g, s);
}

}  // namespace

extern "C" const char* b200_stencil5_variant_info(int v) {
    return (v >= 0 && v < kNumVariants) ? kVariants[v].info : nullptr;
}

extern "C" int b200_stencil5_num_partials(const b200_band* band) {
    Geometry g;
    static const double dummy = 0;
    if (build_geometry(band, &dummy, &g) != B200_OK) return -1;
    return g.grid;
}

extern "C" int b200_stencil5_spmv(const b200_band* band, const double* d_x, double* d_y, b200_stream stream) {
    Geometry g;
    int rc = build_geometry(band, d_x, &g);
    if (rc) return rc;
    if (!d_y) return fail(B200_EINVAL, "stencil5: NULL y");
    g.a.y = d_y;
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
// keep in sync with the dispatch switch in launch_csr
const CsrVariant kCsrVariants[] = {
    {8, 4, 128, "c0 (default): 8 warps/CTA, ring of 4 x 128-entry windows, 4 CTAs/SM"},
    {8, 4, 128, "c1: 8 warps/CTA, ring of 4 x 128-entry windows, 3 CTAs/SM"},
    {8, 4, 256, "c2: 8 warps/CTA, ring of 4 x 256-entry windows, 2 CTAs/SM"},
    {4, 4, 256, "c3: 4 warps/CTA, ring of 4 x 256-entry windows, 4 CTAs/SM"},
    {8, 8, 64, "c4: 8 warps/CTA, ring of 8 x 64-entry windows, 4 CTAs/SM"},
    {4, 4, 128, "c5: 4 warps/CTA, ring of 4 x 128-entry windows, 8 CTAs/SM"},
    {8, 8, 128, "c6: 8 warps/CTA, ring of 8 x 128-entry windows, 2 CTAs/SM"},
    {16, 4, 128, "c7: 16 warps/CTA, ring of 4 x 128-entry windows, 2 CTAs/SM"},
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
    static int resident_ctas = 0;  // per instantiation: CTAs that fit the whole GPU at once
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
        case 1: return launch_csr_ring<8, 4, 128, ELL, 3>(a, gpw, s);
        case 2: return launch_csr_ring<8, 4, 256, ELL, 2>(a, gpw, s);
        case 3: return launch_csr_ring<4, 4, 256, ELL, 4>(a, gpw, s);
        case 4: return launch_csr_ring<8, 8, 64, ELL, 4>(a, gpw, s);
        case 5: return launch_csr_ring<4, 4, 128, ELL, 8>(a, gpw, s);
        case 6: return launch_csr_ring<8, 8, 128, ELL, 2>(a, gpw, s);
        case 7: return launch_csr_ring<16, 4, 128, ELL, 2>(a, gpw, s);
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
    if (variant >= kNumCsrVariants) variant = 0;
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

namespace {
constexpr int kBlas1Ctas = 148 * 8;  // 8 resident 256-thread CTAs per SM, one wave
inline int blas1_grid(long long n, int vec, int cap = kBlas1Ctas) {
    const long long tile = 256LL * vec * 4;
    long long need = (n + tile - 1) / tile;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}
// K2 / K2r (two or four input streams, unrolled double2 loads): 2 CTAs per SM measured best on B200
// (1.503 vs 1.555 ms for 24 B/row at 20k x 20k).  Both kernels MUST use the same grid: their r.r
// partials are summed in the same order, which keeps the two CG schedules bit-identical.
constexpr int kRrCtas = 148 * 2;
inline bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }
}  // namespace

extern "C" int b200_cg_max_partials(const b200_band* band) {
    int n = b200_stencil5_num_partials(band);
    if (n < 0) return n;
    return n > kBlas1Ctas ? n : kBlas1Ctas;
}

extern "C" int b200_cg_residual_init(const b200_band* band, const double* d_x, const double* d_b, double* d_r,
                                     double* d_p, double* d_partials, void* d_scalars, b200_stream stream) {
    Geometry g;
    int rc = build_geometry(band, d_x, &g);
    if (rc) return rc;
    if (!d_b || !d_r || !d_p || !d_partials) return fail(B200_EINVAL, "cg_residual_init: NULL argument");
    g.a.y = d_r; g.a.y2 = d_p; g.a.b = d_b; g.a.partials = d_partials;
    if (d_scalars) g.a.error_word = &static_cast<CGScalars*>(d_scalars)->error;
    return launch_stencil<ST_RESID>(band, g, (cudaStream_t)stream);
}

extern "C" int b200_cg_spmv_dot(const b200_band* band, const double* d_p, double* d_Ap, double* d_partials,
                                const void* d_scalars, b200_stream stream) {
    Geometry g;
    int rc = build_geometry(band, d_p, &g);
    if (rc) return rc;
    if (!d_Ap || !d_partials) return fail(B200_EINVAL, "cg_spmv_dot: NULL argument");
    g.a.y = d_Ap; g.a.partials = d_partials;
    if (d_scalars) {
        g.a.converged = &static_cast<const CGScalars*>(d_scalars)->converged;
        g.a.error_word = &const_cast<CGScalars*>(static_cast<const CGScalars*>(d_scalars))->error;
    }
    return launch_stencil<ST_DOT>(band, g, (cudaStream_t)stream);
}

extern "C" int b200_cg_update_xr(long long n, const void* d_scalars, const double* d_p, const double* d_Ap,
                                 double* d_x, double* d_r, double* d_partials, int* n_partials_out,
                                 b200_stream stream) {
    if (!d_scalars || !d_p || !d_Ap || !d_x || !d_r || !d_partials) return fail(B200_EINVAL, "cg_update_xr: NULL argument");
    const bool v2 = aligned16(d_p) && aligned16(d_Ap) && aligned16(d_x) && aligned16(d_r);
    const int grid = blas1_grid(n, v2 ? 2 : 1, kRrCtas);
    if (n_partials_out) *n_partials_out = grid;
    const CGScalars* sc = static_cast<const CGScalars*>(d_scalars);
    if (v2) cg_update_xr_kernel<2><<<grid, 256, 0, (cudaStream_t)stream>>>(n, sc, d_p, d_Ap, d_x, d_r, d_partials);
    else cg_update_xr_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(n, sc, d_p, d_Ap, d_x, d_r, d_partials);
    return check_launch("cg_update_xr_kernel");
}

extern "C" int b200_cg_update_p(long long n, const void* d_scalars, const double* d_r, double* d_p,
                                b200_stream stream) {
    if (!d_scalars || !d_r || !d_p) return fail(B200_EINVAL, "cg_update_p: NULL argument");
    const bool v2 = aligned16(d_r) && aligned16(d_p);
    const int grid = blas1_grid(n, v2 ? 2 : 1, kRrCtas);
    const CGScalars* sc = static_cast<const CGScalars*>(d_scalars);
    if (v2) cg_update_p_kernel<2><<<grid, 256, 0, (cudaStream_t)stream>>>(n, sc, d_r, d_p);
    else cg_update_p_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(n, sc, d_r, d_p);
    return check_launch("cg_update_p_kernel");
}

extern "C" int b200_cg_update_p_push(long long n, const void* d_scalars, const double* d_r, double* d_p, int halo,
                                     double* d_dst_prev, double* d_dst_next, uint32_t* d_flag_prev,
                                     uint32_t* d_flag_next, uint32_t epoch, void* d_my_xchg, b200_stream stream) {
    if (!d_scalars || !d_r || !d_p || !d_my_xchg || halo < 1 || n < halo) return fail(B200_EINVAL, "cg_update_p_push: bad argument");
    if ((d_dst_prev && !d_flag_prev) || (d_dst_next && !d_flag_next)) return fail(B200_EINVAL, "cg_update_p_push: NULL flag");
    HaloPushArgs a;
    a.v_local = d_p; a.n_local = n; a.halo = halo; a.dst_prev = d_dst_prev; a.dst_next = d_dst_next;
    a.flag_prev = d_flag_prev; a.flag_next = d_flag_next; a.epoch = epoch;
    a.push_count = static_cast<XchgArea*>(d_my_xchg)->push_count;
    a.sc = static_cast<const CGScalars*>(d_scalars);
    const int grid = blas1_grid(n, 1);
    cg_update_p_push_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n, a.sc, d_r, d_p, a);
    return check_launch("cg_update_p_push_kernel");
}

extern "C" int b200_cg_spmv_fused(const b200_band* band, const double* d_p_old, const double* d_r, double* d_p_new,
                                  double* d_x, double* d_Ap, double* d_partials, const void* d_scalars,
                                  b200_stream stream) {
    Geometry g;
    int rc = build_geometry(band, d_p_old, &g);
    if (rc) return rc;
    if (!d_r || !d_p_new || !d_x || !d_Ap || !d_partials || !d_scalars) return fail(B200_EINVAL, "cg_spmv_fused: NULL argument");
    if (d_p_new == d_p_old) return fail(B200_EINVAL, "cg_spmv_fused: p_new must not alias p_old");
    const CGScalars* sc = static_cast<const CGScalars*>(d_scalars);
    g.a.y = d_Ap; g.a.y2 = d_p_new; g.a.r = d_r; g.a.xs = d_x; g.a.partials = d_partials;
    g.a.ab = &sc->alpha;
    static_assert(offsetof(CGScalars, beta) == offsetof(CGScalars, alpha) + sizeof(double), "alpha, beta adjacent");
    g.a.converged = &sc->converged;
    g.a.error_word = &const_cast<CGScalars*>(sc)->error;
    return launch_stencil<ST_FUSED>(band, g, (cudaStream_t)stream);
}

namespace {
int launch_update_r(long long n, const void* d_scalars, const double* d_Ap, double* d_r, double* d_partials,
                    int* n_partials_out, const HaloPushArgs* h, cudaStream_t s) {
    const bool v2 = aligned16(d_Ap) && aligned16(d_r);
    const int grid = blas1_grid(n, v2 ? 2 : 1, kRrCtas);
    if (n_partials_out) *n_partials_out = grid;
    const CGScalars* sc = static_cast<const CGScalars*>(d_scalars);
    HaloPushArgs none;
    memset(&none, 0, sizeof none);
    if (h) {
        if (v2) cg_update_r_kernel<2, true><<<grid, 256, 0, s>>>(n, sc, d_Ap, d_r, d_partials, *h);
        else cg_update_r_kernel<1, true><<<grid, 256, 0, s>>>(n, sc, d_Ap, d_r, d_partials, *h);
    } else {
        if (v2) cg_update_r_kernel<2, false><<<grid, 256, 0, s>>>(n, sc, d_Ap, d_r, d_partials, none);
        else cg_update_r_kernel<1, false><<<grid, 256, 0, s>>>(n, sc, d_Ap, d_r, d_partials, none);
    }
    return check_launch("cg_update_r_kernel");
}
}  // namespace

extern "C" int b200_cg_update_r(long long n, const void* d_scalars, const double* d_Ap, double* d_r,
                                double* d_partials, int* n_partials_out, b200_stream stream) {
    if (!d_scalars || !d_Ap || !d_r || !d_partials) return fail(B200_EINVAL, "cg_update_r: NULL argument");
    return launch_update_r(n, d_scalars, d_Ap, d_r, d_partials, n_partials_out, nullptr, (cudaStream_t)stream);
}

extern "C" int b200_cg_update_r_push(long long n, const void* d_scalars, const double* d_Ap, double* d_r,
                                     double* d_partials, int* n_partials_out, int halo, double* d_dst_prev,
                                     double* d_dst_next, uint32_t* d_flag_prev, uint32_t* d_flag_next,
                                     uint32_t epoch, void* d_my_xchg, b200_stream stream) {
    if (!d_scalars || !d_Ap || !d_r || !d_partials || !d_my_xchg || halo < 1 || n < halo)
        return fail(B200_EINVAL, "cg_update_r_push: bad argument");
    if ((d_dst_prev && !d_flag_prev) || (d_dst_next && !d_flag_next)) return fail(B200_EINVAL, "cg_update_r_push: NULL flag");
    HaloPushArgs a;
    a.v_local = d_r; a.n_local = n; a.halo = halo; a.dst_prev = d_dst_prev; a.dst_next = d_dst_next;
    a.flag_prev = d_flag_prev; a.flag_next = d_flag_next; a.epoch = epoch;
    a.push_count = static_cast<XchgArea*>(d_my_xchg)->push_count;
    a.sc = static_cast<const CGScalars*>(d_scalars);
    return launch_update_r(n, d_scalars, d_Ap, d_r, d_partials, n_partials_out, &a, (cudaStream_t)stream);
}

extern "C" int b200_cg_halo_dir(const double* d_r_prev, const double* d_r_next, const double* d_pold_prev,
                                const double* d_pold_next, double* d_pnew_prev, double* d_pnew_next, int halo,
                                const uint32_t* d_flag_prev, const uint32_t* d_flag_next, uint32_t epoch,
                                void* d_scalars, int beta_zero, b200_stream stream) {
    if (!d_scalars || halo < 1) return fail(B200_EINVAL, "cg_halo_dir: bad argument");
    if ((d_r_prev && (!d_pnew_prev || !d_flag_prev || (!beta_zero && !d_pold_prev))) ||
        (d_r_next && (!d_pnew_next || !d_flag_next || (!beta_zero && !d_pold_next))))
        return fail(B200_EINVAL, "cg_halo_dir: NULL buffer");
    if (!d_r_prev && !d_r_next) return B200_OK;
    HaloDirArgs a;
    a.r_prev = d_r_prev; a.r_next = d_r_next; a.pold_prev = d_pold_prev; a.pold_next = d_pold_next;
    a.pnew_prev = d_pnew_prev; a.pnew_next = d_pnew_next; a.halo = halo;
    a.flag_prev = d_flag_prev; a.flag_next = d_flag_next; a.epoch = epoch;
    a.sc = static_cast<CGScalars*>(d_scalars); a.beta_zero = beta_zero;
    int grid = (halo + 255) / 256;
    if (grid > 64) grid = 64;
    cg_halo_dir_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    return check_launch("cg_halo_dir_kernel");
}

extern "C" int b200_cg_finish_x(long long n, const void* d_scalars, const double* d_p0, const double* d_p1,
                                double* d_x, b200_stream stream) {
    if (!d_scalars || !d_p0 || !d_p1 || !d_x) return fail(B200_EINVAL, "cg_finish_x: NULL argument");
    const int grid = blas1_grid(n, 1);
    cg_finish_x_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n, static_cast<const CGScalars*>(d_scalars), d_p0, d_p1, d_x);
    return check_launch("cg_finish_x_kernel");
}

namespace {
int launch_reduce(const double* d_partials, int n_partials, int which, int phases, double tol, void* d_scalars,
                  void* h_status_mapped, double* d_out, int rank, int world, uint32_t epoch, void* const* d_peer_xchg,
                  double* d_stash, const HaloDirArgs* hd, cudaStream_t stream) {
    if (!d_partials || n_partials < 0) return fail(B200_EINVAL, "cg_reduce: bad partials");
    if (which != RED_SUM && !d_scalars) return fail(B200_EINVAL, "cg_reduce: NULL scalars");
    if (which == RED_SUM && !d_out) return fail(B200_EINVAL, "cg_reduce: NULL out");
    if (world < 1 || world > B200_MAX_RANKS || rank < 0 || rank >= world) return fail(B200_EINVAL, "cg_reduce: bad rank/world");
    if (world > 1 && !d_peer_xchg) return fail(B200_EINVAL, "cg_reduce: NULL peer table");
    if (phases != 3 && !d_stash) return fail(B200_EINVAL, "cg_reduce: split phases need a stash");
    ReduceArgs a;
    memset(&a, 0, sizeof a);
    a.partials = d_partials; a.n_partials = n_partials; a.which = which; a.phases = phases; a.tol = tol;
    a.sc = static_cast<CGScalars*>(d_scalars);
    a.status = static_cast<CGStatus*>(h_status_mapped);
    a.out = d_out; a.rank = rank; a.world = world; a.epoch = epoch; a.stash = d_stash;
    if (world > 1) {
        for (int r = 0; r < world; r++) a.peer_xchg[r] = static_cast<XchgArea*>(d_peer_xchg[r]);
        a.my_xchg = a.peer_xchg[rank];
    }
    if (hd) { a.with_halo_dir = 1; a.hd = *hd; }
    cg_reduce_kernel<<<1, 1024, 0, stream>>>(a);
    return check_launch("cg_reduce_kernel");
}
}  // namespace

extern "C" int b200_cg_reduce(const double* d_partials, int n_partials, int which, int phases, double tol,
                              void* d_scalars, void* h_status_mapped, double* d_out, int rank, int world,
                              uint32_t epoch, void* const* d_peer_xchg, double* d_stash, b200_stream stream) {
    return launch_reduce(d_partials, n_partials, which, phases, tol, d_scalars, h_status_mapped, d_out, rank, world, epoch,
                         d_peer_xchg, d_stash, nullptr, (cudaStream_t)stream);
}

extern "C" int b200_cg_reduce_rr_dir(const double* d_partials, int n_partials, int phases, double tol, void* d_scalars,
                                     void* h_status_mapped, int rank, int world, uint32_t epoch,
                                     void* const* d_peer_xchg, double* d_stash, const double* d_r_prev,
                                     const double* d_r_next, const double* d_pold_prev, const double* d_pold_next,
                                     double* d_pnew_prev, double* d_pnew_next, int halo, const uint32_t* d_flag_prev,
                                     const uint32_t* d_flag_next, uint32_t halo_epoch, b200_stream stream) {
    if (halo < 1) return fail(B200_EINVAL, "cg_reduce_rr_dir: bad halo");
    if ((d_r_prev && (!d_pnew_prev || !d_flag_prev || !d_pold_prev)) || (d_r_next && (!d_pnew_next || !d_flag_next || !d_pold_next)))
        return fail(B200_EINVAL, "cg_reduce_rr_dir: NULL buffer");
    HaloDirArgs h;
    h.r_prev = d_r_prev; h.r_next = d_r_next; h.pold_prev = d_pold_prev; h.pold_next = d_pold_next;
    h.pnew_prev = d_pnew_prev; h.pnew_next = d_pnew_next; h.halo = halo;
    h.flag_prev = d_flag_prev; h.flag_next = d_flag_next; h.epoch = halo_epoch;
    h.sc = static_cast<CGScalars*>(d_scalars); h.beta_zero = 0;
    return launch_reduce(d_partials, n_partials, RED_RR, phases, tol, d_scalars, h_status_mapped, nullptr, rank, world,
                         epoch, d_peer_xchg, d_stash, (d_r_prev || d_r_next) ? &h : nullptr, (cudaStream_t)stream);
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

extern "C" int b200_pcg_init(long long n, const double* d_r, const double* d_dinv, double* d_p, double* d_partials,
                             int* n_partials_out, b200_stream stream) {
    if (!d_r || !d_dinv || !d_p || !d_partials) return fail(B200_EINVAL, "pcg_init: NULL argument");
    const int grid = blas1_grid(n, 1);
    if (n_partials_out) *n_partials_out = grid;
    pcg_init_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n, d_r, d_dinv, d_p, d_partials);
    return check_launch("pcg_init_kernel");
}

extern "C" int b200_pcg_update_xr(long long n, const void* d_scalars, const double* d_p, const double* d_Ap,
                                  const double* d_dinv, double* d_x, double* d_r, double* d_partials_rr,
                                  double* d_partials_rz, int* n_partials_out, b200_stream stream) {
    if (!d_scalars || !d_p || !d_Ap || !d_dinv || !d_x || !d_r || !d_partials_rr || !d_partials_rz)
        return fail(B200_EINVAL, "pcg_update_xr: NULL argument");
    const int grid = blas1_grid(n, 1);
    if (n_partials_out) *n_partials_out = grid;
    pcg_update_xr_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n, static_cast<const CGScalars*>(d_scalars), d_p, d_Ap, d_dinv,
                                                                 d_x, d_r, d_partials_rr, d_partials_rz);
    return check_launch("pcg_update_xr_kernel");
}

extern "C" int b200_pcg_update_p(long long n, const void* d_scalars, const double* d_r, const double* d_dinv, double* d_p,
                                 b200_stream stream) {
    if (!d_scalars || !d_r || !d_dinv || !d_p) return fail(B200_EINVAL, "pcg_update_p: NULL argument");
    pcg_update_p_kernel<<<blas1_grid(n, 1), 256, 0, (cudaStream_t)stream>>>(n, static_cast<const CGScalars*>(d_scalars), d_r,
                                                                            d_dinv, d_p);
    return check_launch("pcg_update_p_kernel");
}

extern "C" int b200_pcg_reduce(const double* d_partials_rr, const double* d_partials_rz, int n_partials, double tol,
                               void* d_scalars, void* h_status_mapped, b200_stream stream) {
    if (!d_partials_rr || !d_partials_rz || n_partials < 0 || !d_scalars) return fail(B200_EINVAL, "pcg_reduce: bad argument");
    ReduceArgs a;
    memset(&a, 0, sizeof a);
    a.partials = d_partials_rr; a.partials_b = d_partials_rz; a.n_partials = n_partials; a.which = RED_PCG; a.phases = 3;
    a.tol = tol; a.sc = static_cast<CGScalars*>(d_scalars); a.status = static_cast<CGStatus*>(h_status_mapped);
    a.rank = 0; a.world = 1;
    cg_reduce_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(a);
    return check_launch("cg_reduce_kernel");
}

extern "C" int b200_dot_partials(long long n, const void* d_scalars, const double* d_x, const double* d_y,
                                 double* d_partials, int* n_partials_out, b200_stream stream) {
    if (!d_x || !d_y || !d_partials) return fail(B200_EINVAL, "dot_partials: NULL argument");
    const int grid = blas1_grid(n, 1);
    if (n_partials_out) *n_partials_out = grid;
    dot_partials_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n, static_cast<const CGScalars*>(d_scalars), d_x, d_y,
                                                                d_partials);
    return check_launch("dot_partials_kernel");
}

extern "C" int b200_residual_init_generic(long long n, const double* d_b, const double* d_Ap, double* d_r,
                                          double* d_p, double* d_partials, int* n_partials_out, b200_stream stream) {
    if (!d_b || !d_Ap || !d_r || !d_p || !d_partials) return fail(B200_EINVAL, "residual_init: NULL argument");
    const int grid = blas1_grid(n, 1);
    if (n_partials_out) *n_partials_out = grid;
    residual_init_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n, d_b, d_Ap, d_r, d_p, d_partials);
    return check_launch("residual_init_kernel");
}

extern "C" int b200_checksum_partials(long long n, const double* d_x, double* d_psum, double* d_psq,
                                      int* n_partials_out, b200_stream stream) {
    if (!d_x || !d_psum || !d_psq) return fail(B200_EINVAL, "checksum: NULL argument");
    const int grid = blas1_grid(n, 1);
    if (n_partials_out) *n_partials_out = grid;
    checksum_partials_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n, d_x, d_psum, d_psq);
    return check_launch("checksum_partials_kernel");
}

extern "C" int b200_halo_push(const double* d_v_local, long long n_local, int halo, double* d_dst_prev,
                              double* d_dst_next, uint32_t* d_flag_prev, uint32_t* d_flag_next, uint32_t epoch,
                              void* d_my_xchg, const void* d_scalars, b200_stream stream) {
    if (!d_v_local || !d_my_xchg || halo < 1 || n_local < halo) return fail(B200_EINVAL, "halo_push: bad argument");
    if ((d_dst_prev && !d_flag_prev) || (d_dst_next && !d_flag_next)) return fail(B200_EINVAL, "halo_push: NULL flag");
    HaloPushArgs a;
    a.v_local = d_v_local; a.n_local = n_local; a.halo = halo; a.dst_prev = d_dst_prev; a.dst_next = d_dst_next;
    a.flag_prev = d_flag_prev; a.flag_next = d_flag_next; a.epoch = epoch;
    a.push_count = static_cast<XchgArea*>(d_my_xchg)->push_count;
    a.sc = static_cast<const CGScalars*>(d_scalars);
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

extern "C" int b200_coo_to_csr(const void* d_entries, long long nnz, int rows, int* d_row_ptr, int* d_col_idx,
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
        coo_count_rows_kernel<<<gen_grid(nnz), 256, 0, s>>>(e, nnz, rows, counts.as<int>(), bad);
        if ((rc = check_launch("coo_count_rows_kernel"))) return rc;
    }
    if ((rc = exclusive_scan<int>(counts.as<int>(), scan.as<long long>(), rows, total, s))) return rc;
    int h_bad = 0;
    cudaMemcpyAsync(&h_bad, bad, 4, cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    if (h_bad) return fail(B200_EINVAL, "coo_to_csr: entry with a row index outside [0, rows)");
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
