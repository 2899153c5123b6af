// operators.cpp -- the SpmvOperator table (reference include/spmv.h:125-150, src/spmv/spmv.cu:11-23)
// backed by the sm_100a kernels.
//
//   name (aliases)                     object                 device format / kernel
//   "cusparse-csr" ("csr")             SPMV_CSR               CSR, adaptive stream/vector kernel (no cuSPARSE)
//   "stencil5-csr" ("stencil5")        SPMV_STENCIL5_CSR      CSR arrays, closed-form interior + CSR boundary
//   "ellpack"                          SPMV_ELLPACK           row-major ELLPACK, generic kernel
//   "stencil5-ellpack"                 SPMV_STENCIL5_ELLPACK  ELLPACK width 5, closed-form interior
//   "stencil5-halo-mgpu"               SPMV_STENCIL_HALO_MGPU row bands over all visible GPUs + halos
//
// Contract kept from the reference operators (src/spmv/spmv_stencil_csr_direct.cu:194-302,
// src/spmv/spmv_cusparse_csr.cu:182-316): init() builds the host structure once per process
// (global csr_mat / ellpack_matrix) and uploads it; run_timed() takes HOST vectors, copies them
// itself and reports the event-timed kernel only; run_device() takes DEVICE vectors and launches
// asynchronously on the default stream; free() releases the device side only.
#include <vector>

#include "host_common.h"

using namespace b200host;

namespace b200host {

void DeviceBand::release() {
    cudaFree(d_row_ptr); cudaFree(d_col_idx); cudaFree(d_values);
    d_row_ptr = nullptr; d_col_idx = nullptr; d_values = nullptr;
    values_len = n_local = nnz_local = 0;
}

static int g_variant = 0, g_rows_per_item = 0;

void DeviceBand::describe(b200_band* b) const {
    memset(b, 0, sizeof *b);
    b->d_row_ptr = d_row_ptr; b->d_col_idx = d_col_idx; b->d_values = d_values;
    b->values_len = values_len; b->row_offset = row_offset; b->n_local = n_local;
    b->grid_size = grid; b->layout = layout;
    b->variant = g_variant; b->rows_per_item = g_rows_per_item;
}

int upload_band_csr(const MatrixData* mat, long long off, long long nl, DeviceBand* out, cudaStream_t s) {
    out->row_offset = off; out->n_local = nl; out->grid = mat->grid_size; out->layout = 0;
    if (is_synthetic(mat)) {
        const long long n = mat->grid_size;
        const long long lnnz = b200_stencil5_nnz_before(off + nl, n) - b200_stencil5_nnz_before(off, n);
        out->nnz_local = lnnz;
        out->values_len = lnnz + 2;  // padding so the 16-byte bulk copies never need a manual tail
        B200_CUDA(cudaMalloc(&out->d_row_ptr, (size_t)(nl + 1) * sizeof(int)));
        B200_CUDA(cudaMalloc(&out->d_col_idx, (size_t)(lnnz + 2) * sizeof(int)));
        B200_CUDA(cudaMalloc(&out->d_values, (size_t)(lnnz + 2) * sizeof(double)));
        B200_CUDA(cudaMemsetAsync(out->d_values + lnnz, 0, 2 * sizeof(double), s));
        B200_K(b200_gen_stencil5_csr((int)n, off, nl, 5.0, -1.0, out->d_row_ptr, out->d_col_idx, out->d_values, s));
        return 0;
    }
    if (csr_mat.row_ptr == nullptr) return 1;
    const int base = csr_mat.row_ptr[off];
    const long long lnnz = (long long)csr_mat.row_ptr[off + nl] - base;
    out->nnz_local = lnnz;
    out->values_len = lnnz + 2;
    std::vector<int> rp((size_t)nl + 1);
    for (long long i = 0; i <= nl; i++) rp[(size_t)i] = csr_mat.row_ptr[off + i] - base;  // rebase, cols stay global
    B200_CUDA(cudaMalloc(&out->d_row_ptr, (size_t)(nl + 1) * sizeof(int)));
    B200_CUDA(cudaMalloc(&out->d_col_idx, (size_t)(lnnz + 2) * sizeof(int)));
    B200_CUDA(cudaMalloc(&out->d_values, (size_t)(lnnz + 2) * sizeof(double)));
    B200_CUDA(cudaMemsetAsync(out->d_values + lnnz, 0, 2 * sizeof(double), s));
    B200_CUDA(cudaMemcpyAsync(out->d_row_ptr, rp.data(), (size_t)(nl + 1) * sizeof(int), cudaMemcpyHostToDevice, s));
    B200_CUDA(cudaMemcpyAsync(out->d_col_idx, csr_mat.col_indices + base, (size_t)lnnz * sizeof(int),
                              cudaMemcpyHostToDevice, s));
    B200_CUDA(cudaMemcpyAsync(out->d_values, csr_mat.values + base, (size_t)lnnz * sizeof(double),
                              cudaMemcpyHostToDevice, s));
    B200_CUDA(cudaStreamSynchronize(s));  // rp is a local
    return 0;
}

int upload_band_ell(const MatrixData* mat, long long off, long long nl, DeviceBand* out, cudaStream_t s) {
    out->row_offset = off; out->n_local = nl; out->grid = mat->grid_size; out->layout = 1;
    if (is_synthetic(mat)) {
        out->nnz_local = 5 * nl; out->values_len = 5 * nl + 2;
        B200_CUDA(cudaMalloc(&out->d_col_idx, (size_t)(5 * nl + 2) * sizeof(int)));
        B200_CUDA(cudaMalloc(&out->d_values, (size_t)(5 * nl + 2) * sizeof(double)));
        B200_K(b200_gen_stencil5_ellpack(mat->grid_size, off, nl, 5.0, -1.0, out->d_col_idx, out->d_values, s));
        return 0;
    }
    if (ellpack_matrix.indices == nullptr) return 1;
    const long long w = ellpack_matrix.ell_width;
    out->nnz_local = w * nl; out->values_len = w * nl + 2;
    B200_CUDA(cudaMalloc(&out->d_col_idx, (size_t)(w * nl + 2) * sizeof(int)));
    B200_CUDA(cudaMalloc(&out->d_values, (size_t)(w * nl + 2) * sizeof(double)));
    B200_CUDA(cudaMemcpyAsync(out->d_col_idx, ellpack_matrix.indices + off * w, (size_t)(w * nl) * sizeof(int),
                              cudaMemcpyHostToDevice, s));
    B200_CUDA(cudaMemcpyAsync(out->d_values, ellpack_matrix.values + off * w, (size_t)(w * nl) * sizeof(double),
                              cudaMemcpyHostToDevice, s));
    B200_CUDA(cudaStreamSynchronize(s));
    return 0;
}

}  // namespace b200host

extern "C" int b200_set_tuning(int variant, int rows_per_item) {
    b200host::g_variant = variant;
    b200host::g_rows_per_item = rows_per_item;
    return 0;
}
extern "C" void b200_get_tuning(int* variant, int* rows_per_item) {
    if (variant) *variant = b200host::g_variant;
    if (rows_per_item) *rows_per_item = b200host::g_rows_per_item;
}

// ------------------------------------------------------------------------------------------------
// one operator instance = device matrix + scratch vectors for run_timed
// ------------------------------------------------------------------------------------------------
namespace {

enum Kind { K_CSR, K_STENCIL_CSR, K_ELL, K_STENCIL_ELL };

struct OpState {
    Kind kind;
    DeviceBand band;
    b200_csr_plan plan;
    double* dX = nullptr;
    double* dY = nullptr;
    int rows = 0, cols = 0, ell_width = 0;
    bool ready = false;

    void reset() {
        band.release();
        cudaFree(dX); cudaFree(dY);
        dX = dY = nullptr;
        ready = false;
    }

    int init(MatrixData* mat) {
        if (!mat || mat->rows <= 0) return EXIT_FAILURE;
        reset();
        rows = mat->rows; cols = mat->cols;
        const bool stencil = (kind == K_STENCIL_CSR || kind == K_STENCIL_ELL);
        if (stencil) {
            const long long n = mat->grid_size;
            if (n < 1 || n * n != (long long)mat->rows || mat->rows != mat->cols) {
                fprintf(stderr, "[ERROR] stencil operator needs an n x n grid matrix (grid_size=%d, rows=%d)\n",
                        mat->grid_size, mat->rows);
                return EXIT_FAILURE;
            }
        }
        const bool ell = (kind == K_ELL || kind == K_STENCIL_ELL);
        if (!is_synthetic(mat)) {
            if (ell ? ensure_ellpack_structure_built(mat) != EXIT_SUCCESS : build_csr_struct(mat) != EXIT_SUCCESS)
                return EXIT_FAILURE;
        } else if (!stencil && !(mat->grid_size > 0)) {
            return EXIT_FAILURE;
        }
        int rc = ell ? upload_band_ell(mat, 0, mat->rows, &band, 0) : upload_band_csr(mat, 0, mat->rows, &band, 0);
        if (rc) { reset(); return EXIT_FAILURE; }
        ell_width = ell ? (is_synthetic(mat) ? 5 : ellpack_matrix.ell_width) : 0;
        if (kind == K_STENCIL_ELL && ell_width != 5) {
            fprintf(stderr, "[ERROR] stencil5-ellpack needs ELLPACK width 5 (got %d)\n", ell_width);
            reset();
            return EXIT_FAILURE;
        }
        if (kind == K_CSR) {
            rc = b200_csr_plan_build(band.d_row_ptr, rows, band.nnz_local, &plan, 0);
            if (rc) { fprintf(stderr, "[b200] %s\n", b200_last_error()); reset(); return EXIT_FAILURE; }
        }
        if (cudaMalloc(&dX, (size_t)cols * sizeof(double)) != cudaSuccess ||
            cudaMalloc(&dY, (size_t)rows * sizeof(double)) != cudaSuccess) {
            fprintf(stderr, "[ERROR] cudaMalloc failed for operator vectors\n");
            reset();
            return EXIT_FAILURE;
        }
        if (cudaDeviceSynchronize() != cudaSuccess) { reset(); return EXIT_FAILURE; }
        ready = true;
        return EXIT_SUCCESS;
    }

    int launch(const double* d_x, double* d_y, cudaStream_t s) {
        if (!ready) { fprintf(stderr, "[ERROR] operator used before init()\n"); return EXIT_FAILURE; }
        int rc;
        switch (kind) {
            case K_CSR:
                rc = b200_spmv_csr(&plan, band.d_row_ptr, band.d_col_idx, band.d_values, d_x, d_y, rows, 1.0, 0.0, s);
                break;
            case K_ELL:
                rc = b200_spmv_ellpack(band.d_col_idx, band.d_values, d_x, d_y, rows, ell_width, 1.0, 0.0, s);
                break;
            default: {
                b200_band b;
                band.describe(&b);
                rc = b200_stencil5_spmv(&b, d_x, d_y, s);
            }
        }
        if (rc) fprintf(stderr, "[b200] SpMV launch failed (%d): %s\n", rc, b200_last_error());
        return rc;
    }

    // y = A x fused with the partials of x.y (generic CSR / ELLPACK only); -1 = not available here
    int launch_dot(const double* d_x, double* d_y, double* d_partials, long long cap, int* np, const void* scalars,
                   cudaStream_t s) {
        if (!ready || rows != cols) return -1;
        int rc;
        if (kind == K_CSR)
            rc = b200_spmv_csr_dot(&plan, band.d_row_ptr, band.d_col_idx, band.d_values, d_x, d_y, rows, d_partials, cap, np,
                                   scalars, s);
        else if (kind == K_ELL)
            rc = b200_spmv_ellpack_dot(band.d_col_idx, band.d_values, d_x, d_y, rows, ell_width, d_partials, cap, np, scalars, s);
        else
            return -1;
        return rc == B200_OK ? 0 : -1;  // e.g. unaligned arrays: the caller falls back to SpMV + dot kernel
    }

    int run_timed(const double* x, double* y, double* ms) {
        if (!ready) return EXIT_FAILURE;
        cudaEvent_t e0, e1;
        B200_CUDA(cudaMemcpy(dX, x, (size_t)cols * sizeof(double), cudaMemcpyHostToDevice));
        B200_CUDA(cudaEventCreate(&e0));
        B200_CUDA(cudaEventCreate(&e1));
        B200_CUDA(cudaEventRecord(e0, 0));
        int rc = launch(dX, dY, 0);
        B200_CUDA(cudaEventRecord(e1, 0));
        B200_CUDA(cudaEventSynchronize(e1));
        float t = 0.f;
        B200_CUDA(cudaEventElapsedTime(&t, e0, e1));
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        if (rc) return rc;
        if (ms) *ms = (double)t;
        B200_CUDA(cudaMemcpy(y, dY, (size_t)rows * sizeof(double), cudaMemcpyDeviceToHost));
        return EXIT_SUCCESS;
    }
};

OpState g_csr{K_CSR}, g_st_csr{K_STENCIL_CSR}, g_ell{K_ELL}, g_st_ell{K_STENCIL_ELL};

#define OP_THUNKS(tag, st)                                                                    \
    int tag##_init(MatrixData* m) { return st.init(m); }                                      \
    int tag##_run_timed(const double* x, double* y, double* ms) { return st.run_timed(x, y, ms); } \
    int tag##_run_device(const double* dx, double* dy) { return st.launch(dx, dy, 0); }       \
    void tag##_free() { st.reset(); }

OP_THUNKS(csr, g_csr)
OP_THUNKS(st_csr, g_st_csr)
OP_THUNKS(ell, g_ell)
OP_THUNKS(st_ell, g_st_ell)

}  // namespace

SpmvOperator SPMV_CSR = {"cusparse-csr", csr_init, csr_run_timed, csr_run_device, csr_free};
SpmvOperator SPMV_STENCIL5_CSR = {"stencil5-csr", st_csr_init, st_csr_run_timed, st_csr_run_device, st_csr_free};
SpmvOperator SPMV_ELLPACK = {"ellpack", ell_init, ell_run_timed, ell_run_device, ell_free};
SpmvOperator SPMV_STENCIL5_ELLPACK = {"stencil5-ellpack", st_ell_init, st_ell_run_timed, st_ell_run_device, st_ell_free};

namespace b200host {
int operator_spmv_dot(const SpmvOperator* op, const double* d_x, double* d_y, double* d_partials, long long cap, int* np,
                      const void* scalars) {
    if (op == &SPMV_CSR) return g_csr.launch_dot(d_x, d_y, d_partials, cap, np, scalars, 0);
    if (op == &SPMV_ELLPACK) return g_ell.launch_dot(d_x, d_y, d_partials, cap, np, scalars, 0);
    return -1;
}
// device matrix of any of this library's four operators (Jacobi PCG reads the diagonal from it)
const DeviceBand* operator_matrix(const SpmvOperator* op, int* ell_width) {
    const OpState* st = op == &SPMV_CSR ? &g_csr : op == &SPMV_ELLPACK ? &g_ell : op == &SPMV_STENCIL5_CSR ? &g_st_csr
                        : op == &SPMV_STENCIL5_ELLPACK ? &g_st_ell : nullptr;
    if (!st || !st->ready) return nullptr;
    if (ell_width) *ell_width = st->ell_width;
    return &st->band;
}
const DeviceBand* operator_band(const SpmvOperator* op) {
    if (op == &SPMV_STENCIL5_CSR && g_st_csr.ready) return &g_st_csr.band;
    if (op == &SPMV_STENCIL5_ELLPACK && g_st_ell.ready) return &g_st_ell.band;
    return nullptr;
}
}  // namespace b200host

extern "C" int b200_operator_band(const SpmvOperator* op, b200_band* out) {
    const DeviceBand* b = operator_band(op);
    if (!b || !out) return 1;
    b->describe(out);
    return 0;
}

extern "C" SpmvOperator* get_operator(const char* mode) {
    if (!mode) return nullptr;
    if (!strcmp(mode, "cusparse-csr") || !strcmp(mode, "csr")) return &SPMV_CSR;
    if (!strcmp(mode, "stencil5-csr") || !strcmp(mode, "stencil5")) return &SPMV_STENCIL5_CSR;
    if (!strcmp(mode, "ellpack")) return &SPMV_ELLPACK;
    if (!strcmp(mode, "stencil5-ellpack")) return &SPMV_STENCIL5_ELLPACK;
    if (!strcmp(mode, "stencil5-halo-mgpu")) return &SPMV_STENCIL_HALO_MGPU;
    return nullptr;
}

// ------------------------------------------------------------------------------------------------
// Device-resident ingest: initialise a CSR-based operator from COO entries that
// already live on the GPU (e.g. parsed there by b200_load_matrix_market_device).  No host Entry[]
// and no host CSR are built; csr_mat only carries the sizes.
// ------------------------------------------------------------------------------------------------
extern "C" int b200_operator_init_device_coo(SpmvOperator* op, const MatrixData* meta, const void* d_entries) {
    OpState* st = (op == &SPMV_CSR) ? &g_csr : (op == &SPMV_STENCIL5_CSR) ? &g_st_csr : nullptr;
    if (!st || !meta || (!d_entries && meta->nnz > 0) || meta->rows <= 0) {
        fprintf(stderr, "[ERROR] b200_operator_init_device_coo: needs the cusparse-csr or stencil5-csr operator\n");
        return EXIT_FAILURE;
    }
    if (st->kind == K_STENCIL_CSR) {
        const long long n = meta->grid_size;
        if (n < 1 || n * n != (long long)meta->rows) return EXIT_FAILURE;
    }
    st->reset();
    st->rows = meta->rows; st->cols = meta->cols; st->ell_width = 0;
    DeviceBand& b = st->band;
    b.row_offset = 0; b.n_local = meta->rows; b.nnz_local = meta->nnz; b.values_len = (long long)meta->nnz + 2;
    b.grid = meta->grid_size; b.layout = 0;
    B200_CUDA(cudaMalloc(&b.d_row_ptr, ((size_t)meta->rows + 1) * sizeof(int)));
    B200_CUDA(cudaMalloc(&b.d_col_idx, ((size_t)meta->nnz + 2) * sizeof(int)));
    B200_CUDA(cudaMalloc(&b.d_values, ((size_t)meta->nnz + 2) * sizeof(double)));
    B200_CUDA(cudaMemset(b.d_values + meta->nnz, 0, 2 * sizeof(double)));
    int rc = b200_coo_to_csr(d_entries, meta->nnz, meta->rows, meta->cols, b.d_row_ptr, b.d_col_idx, b.d_values, 0);
    if (rc) { fprintf(stderr, "[b200] %s\n", b200_last_error()); st->reset(); return EXIT_FAILURE; }
    if (st->kind == K_CSR) {
        rc = b200_csr_plan_build(b.d_row_ptr, meta->rows, meta->nnz, &st->plan, 0);
        if (rc) { fprintf(stderr, "[b200] %s\n", b200_last_error()); st->reset(); return EXIT_FAILURE; }
    }
    B200_CUDA(cudaMalloc(&st->dX, (size_t)meta->cols * sizeof(double)));
    B200_CUDA(cudaMalloc(&st->dY, (size_t)meta->rows * sizeof(double)));
    B200_CUDA(cudaDeviceSynchronize());
    csr_mat.nb_rows = meta->rows; csr_mat.nb_cols = meta->cols; csr_mat.nb_nonzeros = meta->nnz;
    st->ready = true;
    return EXIT_SUCCESS;
}

// copies of the operator's device CSR arrays (tests / debugging)
extern "C" int b200_operator_device_csr(SpmvOperator* op, const int** d_row_ptr, const int** d_col_idx,
                                        const double** d_values, long long* nnz) {
    const DeviceBand* b = (op == &SPMV_CSR && g_csr.ready) ? &g_csr.band : operator_band(op);
    if (!b) return 1;
    if (d_row_ptr) *d_row_ptr = b->d_row_ptr;
    if (d_col_idx) *d_col_idx = b->d_col_idx;
    if (d_values) *d_values = b->d_values;
    if (nnz) *nnz = b->nnz_local;
    return 0;
}
