/*
 * oracle.c -- CPU restatement of the reference's SpMV + CG hot path (plain C).
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Citations are file:line in the
 * upstream reference tree.  Compile with -ffp-contract=off: every fused
 * multiply-add the reference's nvcc build performs is written as an explicit
 * fma() below, in the operand order read from the reference's PTX
 * (nvcc -O2, default -fmad=true), so that results are bit-comparable.
 */
#include "oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------ */
/* structure                                                           */
/* ------------------------------------------------------------------ */

/* nnz count of the n x n 5-point stencil: src/io/io.cu:327-340 (loop form there; closed form
 * 5n^2 - 4n here, checked against the loop in tests). */
long long orc_stencil5_nnz(int n) {
    return 5LL * n * n - 4LL * n;
}

/* Generator emission order: src/io/io.cu:362-392 -- row-major over the grid, per point
 * Center, Left (col>0), Right (col<n-1), Top (row>0), Bottom (row<n-1); indices here are the
 * 0-based ones the reader produces (src/io/io.cu:164-165). */
void orc_stencil5_entries(int n, double center, double neighbour, orc_entry* out) {
    long long k = 0;
    for (int row = 0; row < n; row++) {
        for (int col = 0; col < n; col++) {
            int idx = row * n + col;
            out[k].row = idx; out[k].col = idx; out[k].value = center; k++;
            if (col > 0) { out[k].row = idx; out[k].col = idx - 1; out[k].value = neighbour; k++; }
            if (col < n - 1) { out[k].row = idx; out[k].col = idx + 1; out[k].value = neighbour; k++; }
            if (row > 0) { out[k].row = idx; out[k].col = idx - n; out[k].value = neighbour; k++; }
            if (row < n - 1) { out[k].row = idx; out[k].col = idx + n; out[k].value = neighbour; k++; }
        }
    }
}

/* File writer: src/io/io.cu:322-399.  Header lines :348-351, entry lines :375-391.  The value
 * tokens are literals in the reference ("5.0"/"-1.0"; the bundled matrix/example81x81.mtx was
 * written by an older revision with "-4.0"), hence passed as text. */
int orc_write_mtx_stencil5(int n, const char* filename, const char* center_txt, const char* nb_txt) {
    FILE* f = fopen(filename, "w");
    if (!f) return 1;
    int N = n * n;
    long long nnz = orc_stencil5_nnz(n);
    fprintf(f, "%%%%MatrixMarket matrix coordinate real general\n");
    fprintf(f, "%% STENCIL_GRID_SIZE %d\n", n);
    fprintf(f, "%d %d %d\n", N, N, (int)nnz);
    for (int row = 0; row < n; row++) {
        for (int col = 0; col < n; col++) {
            int idx = row * n + col + 1;
            fprintf(f, "%d %d %s\n", idx, idx, center_txt);
            if (col > 0) fprintf(f, "%d %d %s\n", idx, idx - 1, nb_txt);
            if (col < n - 1) fprintf(f, "%d %d %s\n", idx, idx + 1, nb_txt);
            if (row > 0) fprintf(f, "%d %d %s\n", idx, idx - n, nb_txt);
            if (row < n - 1) fprintf(f, "%d %d %s\n", idx, idx + n, nb_txt);
        }
    }
    fclose(f);
    return 0;
}

/* Reader: src/io/io.cu:109-171 (general matrices).  Skips '%' lines, remembers the
 * "% STENCIL_GRID_SIZE n" comment (:130-132, default -1), first non-comment line is
 * "rows cols nnz" (:126-128), then nnz x "%d %d %le" with 1->0-based shift (:153-166).
 * Symmetric files: the reference's path (io.cu:189-310) never fills `mat` (a bug noted in
 * SURVEY.md); the oracle therefore only defines the general path. */
int orc_load_mtx(const char* filename, orc_matrix* mat) {
    FILE* file = fopen(filename, "r");
    if (!file) return 1;
    char buffer[1024];
    int rows = 0, cols = 0, nnz = 0, grid = -1;
    while (fgets(buffer, sizeof buffer, file) != NULL) {
        if (buffer[0] != '%') {
            sscanf(buffer, "%d %d %d", &rows, &cols, &nnz);
            break;
        } else if (strstr(buffer, "STENCIL_GRID_SIZE") != NULL) {
            sscanf(buffer, "%% STENCIL_GRID_SIZE %d", &grid);
        }
    }
    orc_entry* e = (orc_entry*)malloc((size_t)(nnz > 0 ? nnz : 1) * sizeof(orc_entry));
    if (!e) { fclose(file); return 2; }
    mat->entries = e; mat->rows = rows; mat->cols = cols; mat->nnz = nnz; mat->grid_size = grid;
    for (int i = 0; i < nnz; i++) {
        if (fscanf(file, "%d %d %le", &e[i].row, &e[i].col, &e[i].value) != 3) {
            fclose(file);
            return 3;
        }
        e[i].row--; e[i].col--;
    }
    fclose(file);
    return 0;
}

/* COO -> CSR: src/spmv/spmv_cusparse_csr.cu:85-157 -- count per row, inclusive prefix sum,
 * scatter in file order, then a per-row insertion sort by column index (stable, strict '>'). */
int orc_build_csr(const orc_matrix* mat, orc_csr* out) {
    int rows = mat->rows, nnz = mat->nnz;
    int* row_ptr = (int*)calloc((size_t)rows + 1, sizeof(int));
    int* col = (int*)malloc((size_t)(nnz > 0 ? nnz : 1) * sizeof(int));
    double* val = (double*)malloc((size_t)(nnz > 0 ? nnz : 1) * sizeof(double));
    int* cnt = (int*)calloc((size_t)(rows > 0 ? rows : 1), sizeof(int));
    if (!row_ptr || !col || !val || !cnt) return 1;
    for (int i = 0; i < nnz; ++i) row_ptr[mat->entries[i].row + 1]++;
    for (int i = 1; i <= rows; ++i) row_ptr[i] += row_ptr[i - 1];
    for (int i = 0; i < nnz; ++i) {
        int r = mat->entries[i].row;
        int dst = row_ptr[r] + cnt[r]++;
        col[dst] = mat->entries[i].col;
        val[dst] = mat->entries[i].value;
    }
    free(cnt);
    for (int r = 0; r < rows; ++r) {
        int s = row_ptr[r], e = row_ptr[r + 1];
        for (int i = s + 1; i < e; ++i) {
            int kc = col[i]; double kv = val[i];
            int j = i - 1;
            while (j >= s && col[j] > kc) { col[j + 1] = col[j]; val[j + 1] = val[j]; j--; }
            col[j + 1] = kc; val[j + 1] = kv;
        }
    }
    out->nb_rows = rows; out->nb_cols = mat->cols; out->nb_nonzeros = nnz;
    out->row_ptr = row_ptr; out->col_indices = col; out->values = val;
    return 0;
}

void orc_free_csr(orc_csr* c) {
    free(c->row_ptr); free(c->col_indices); free(c->values);
    c->row_ptr = NULL; c->col_indices = NULL; c->values = NULL;
}

/* Closed form of generator -> orc_build_csr for the stencil (sorted row = N,W,C,E,S where the
 * neighbour exists): the layout the reference's kernel assumes in
 * src/spmv/spmv_stencil_csr_direct.cu:50-67,95-109.  64-bit row_ptr so that n > 20724 works. */
/* non-zeros in rows [0, r) of the n x n stencil: 5 r minus the neighbours that fall off the grid */
static int64_t stencil5_nnz_before_row(int64_t r, int64_t n) {
    int64_t i = r / n, j = r % n;
    int64_t no_north = r < n ? r : n;
    int64_t no_south = r > (n - 1) * n ? r - (n - 1) * n : 0;
    int64_t no_west = i + (j > 0 ? 1 : 0);
    int64_t no_east = i;
    return 5 * r - no_north - no_south - no_west - no_east;
}

void orc_stencil5_csr_direct(int n, double center, double neighbour, int64_t* row_ptr64,
                             int* col_idx, double* values) {
    /* grid rows are independent once the closed-form start of each is known: OpenMP over i (the
     * 20k x 20k matrix of bench.py's CPU arm is 2e9 entries) */
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        int64_t k = stencil5_nnz_before_row((int64_t)i * n, n);
        for (int j = 0; j < n; j++) {
            int64_t r = (int64_t)i * n + j;
            row_ptr64[r] = k;
            if (i > 0) { col_idx[k] = (int)(r - n); values[k] = neighbour; k++; }
            if (j > 0) { col_idx[k] = (int)(r - 1); values[k] = neighbour; k++; }
            col_idx[k] = (int)r; values[k] = center; k++;
            if (j < n - 1) { col_idx[k] = (int)(r + 1); values[k] = neighbour; k++; }
            if (i < n - 1) { col_idx[k] = (int)(r + n); values[k] = neighbour; k++; }
        }
    }
    row_ptr64[(int64_t)n * n] = stencil5_nnz_before_row((int64_t)n * n, n);
}

/* src/spmv/spmv_stencil_csr_direct.cu:50-67 (int there; int64 here, identical while < 2^31). */
int64_t orc_interior_csr_offset(int64_t row, int grid_size) {
    int64_t i = row / grid_size, j = row % grid_size;
    int64_t row0_nnz = 3 + (int64_t)(grid_size - 2) * 4 + 3;
    int64_t interior_row_nnz = 4 + (int64_t)(grid_size - 2) * 5 + 4;
    return row0_nnz + (i - 1) * interior_row_nnz + 4 + (j - 1) * 5;
}

/* ELLPACK from CSR.  The reference only declares this (include/spmv_ellpack.h:28-51: width =
 * max row nnz, "indices ... (row-major)", values "aligned ... per row"); no definition and no
 * test exists => PARITY UNPINNED.  Defined here: slot k of row r lives at r*width + k, entries
 * in CSR order, padding = index -1 / value 0.0. */
int orc_build_ellpack(const orc_csr* csr, orc_ell* out, int* max_width) {
    int w = 0;
    for (int r = 0; r < csr->nb_rows; r++) {
        int len = csr->row_ptr[r + 1] - csr->row_ptr[r];
        if (len > w) w = len;
    }
    size_t tot = (size_t)csr->nb_rows * (size_t)(w > 0 ? w : 1);
    out->indices = (int*)malloc(tot * sizeof(int));
    out->values = (double*)malloc(tot * sizeof(double));
    if (!out->indices || !out->values) return 1;
    for (int r = 0; r < csr->nb_rows; r++) {
        int s = csr->row_ptr[r], len = csr->row_ptr[r + 1] - s;
        for (int k = 0; k < w; k++) {
            size_t d = (size_t)r * w + k;
            if (k < len) { out->indices[d] = csr->col_indices[s + k]; out->values[d] = csr->values[s + k]; }
            else { out->indices[d] = -1; out->values[d] = 0.0; }
        }
    }
    out->nb_rows = csr->nb_rows; out->nb_cols = csr->nb_cols; out->ell_width = w;
    out->grid_size = -1; out->nb_nonzeros = csr->nb_nonzeros;
    if (max_width) *max_width = w;
    return 0;
}

void orc_free_ell(orc_ell* e) {
    free(e->indices); free(e->values); e->indices = NULL; e->values = NULL;
}

/* ------------------------------------------------------------------ */
/* SpMV                                                                */
/* ------------------------------------------------------------------ */

/* Generic CSR semantics: src/solvers/cg_solver_mgpu_partitioned.cu:40-56 (csr_spmv_kernel) and
 * the boundary branch of src/spmv/spmv_stencil_csr_direct.cu:113-119: sum = 0; for k in row:
 * sum += v[k]*x[col[k]]  -> one DFMA per k (PTX: fma.rn.f64 sum, v, x, sum). */
void orc_csr_spmv(const int* row_ptr, const int* col, const double* val, const double* x,
                  double* y, int n_rows) {
#pragma omp parallel for schedule(static)
    for (int r = 0; r < n_rows; r++) {
        double sum = 0.0;
        for (int k = row_ptr[r]; k < row_ptr[r + 1]; k++) sum = fma(val[k], x[col[k]], sum);
        y[r] = sum;
    }
}

/* STENCIL5 CSR-direct: src/spmv/spmv_stencil_csr_direct.cu:76-123.  Interior rows use the
 * closed-form offset (:95) and the expression (:105-109)
 *   vW*xW + vC*xC + vE*xE + vN*xN + vS*xS
 * which nvcc contracts to  t = vC*xC; t = fma(vW,xW,t); fma(vE,xE,t); fma(vN,xN,t);
 * fma(vS,xS,t)  (read from the reference's PTX; the first product folded is the LEFT operand
 * of the first add).  y = alpha*sum with alpha = 1.0 (:33) is exact.  Boundary rows: CSR walk. */
void orc_stencil5_spmv(const int* row_ptr, const int* col, const double* val, const double* x,
                       double* y, int n_rows, int grid) {
#pragma omp parallel for schedule(static)
    for (int row = 0; row < n_rows; row++) {
        int i = row / grid, j = row % grid;
        double sum;
        if (i > 0 && i < grid - 1 && j > 0 && j < grid - 1) {
            int64_t o = orc_interior_csr_offset(row, grid);
            sum = val[o + 2] * x[row];
            sum = fma(val[o + 1], x[row - 1], sum);
            sum = fma(val[o + 3], x[row + 1], sum);
            sum = fma(val[o + 0], x[row - grid], sum);
            sum = fma(val[o + 4], x[row + grid], sum);
        } else {
            sum = 0.0;
            for (int k = row_ptr[row]; k < row_ptr[row + 1]; k++) sum = fma(val[k], x[col[k]], sum);
        }
        y[row] = 1.0 * sum;
    }
}

/* ELLPACK SpMV: unpinned (no reference implementation); defined as the CSR sum over the
 * non-padding slots in slot order, so y_ELL is bit-identical to orc_csr_spmv. */
void orc_ell_spmv(const orc_ell* e, const double* x, double* y) {
    int w = e->ell_width;
#pragma omp parallel for schedule(static)
    for (int r = 0; r < e->nb_rows; r++) {
        double sum = 0.0;
        for (int k = 0; k < w; k++) {
            int c = e->indices[(size_t)r * w + k];
            if (c >= 0) sum = fma(e->values[(size_t)r * w + k], x[c], sum);
        }
        y[r] = sum;
    }
}

static double halo_fetch(int64_t g, const double* xl, const double* hp, const double* hn,
                         int64_t off, int64_t nl, int grid) {
    if (g >= off && g < off + nl) return xl[g - off];
    if (hp != NULL && g >= off - grid && g < off) return hp[g - (off - grid)];
    if (hn != NULL && g >= off + nl && g < off + nl + grid) return hn[g - (off + nl)];
    return 0.0;
}

/* Band SpMV with halos: src/spmv/spmv_stencil_partitioned_halo_kernel.cu:17-98.  Interior test
 * additionally requires row length 5 (:36); N/S come from local or halo by range (:43-68; NULL
 * halo is only checked in the boundary branch :83-88, the interior branch can never need a
 * missing halo); same contraction order as the single-GPU kernel (:72-74). */
void orc_halo_spmv(const int* rp, const int* colg, const double* val, const double* xl,
                   const double* hp, const double* hn, double* y, int n_local, int64_t off,
                   int64_t N, int grid) {
    (void)N;
#pragma omp parallel for schedule(static)
    for (int lr = 0; lr < n_local; lr++) {
        int64_t row = off + lr;
        int64_t i = row / grid, j = row % grid;
        int s = rp[lr], e = rp[lr + 1];
        double sum;
        if (i > 0 && i < grid - 1 && j > 0 && j < grid - 1 && (e - s) == 5) {
            double xn = halo_fetch(row - grid, xl, hp, hn, off, n_local, grid);
            double xs = halo_fetch(row + grid, xl, hp, hn, off, n_local, grid);
            sum = val[s + 2] * xl[lr];
            sum = fma(val[s + 1], xl[lr - 1], sum);
            sum = fma(val[s + 3], xl[lr + 1], sum);
            sum = fma(val[s + 0], xn, sum);
            sum = fma(val[s + 4], xs, sum);
        } else {
            sum = 0.0;
            for (int k = s; k < e; k++)
                sum = fma(val[k], halo_fetch(colg[k], xl, hp, hn, off, n_local, grid), sum);
        }
        y[lr] = sum;
    }
}

/* ------------------------------------------------------------------ */
/* reductions                                                          */
/* ------------------------------------------------------------------ */

/* dot_kernel + final_sum_kernel: src/solvers/cg_solver.cu:110-132 and :384-409.
 * Stage 1: blocks of 256, sdata[t] = x[i]*y[i] (0 past n), halving tree s = 128..1.
 * Stage 2: one block of 256, thread t sums block_results[t], [t+256], ... in order, same tree. */
static double tree256(double* s) {
    for (int h = 128; h > 0; h >>= 1)
        for (int t = 0; t < h; t++) s[t] += s[t + h];
    return s[0];
}

double orc_dot_blocktree(int n, const double* x, const double* y) {
    int blocks = (n + 255) / 256;
    double* br = (double*)malloc((size_t)(blocks > 0 ? blocks : 1) * sizeof(double));
#pragma omp parallel for schedule(static)
    for (int b = 0; b < blocks; b++) {
        double s[256];
        for (int t = 0; t < 256; t++) {
            long long i = (long long)b * 256 + t;
            s[t] = (i < n) ? x[i] * y[i] : 0.0;
        }
        br[b] = tree256(s);
    }
    double s[256];
    for (int t = 0; t < 256; t++) {
        double a = 0.0;
        for (int i = t; i < blocks; i += 256) a += br[i];
        s[t] = a;
    }
    double r = tree256(s);
    free(br);
    return r;
}

/* cublasDdot stand-in for the mgpu path (src/solvers/cg_solver_mgpu_partitioned.cu:145-154):
 * the library's summation order is not published => unpinned; sequential fma chain here. */
double orc_dot_sequential(int n, const double* x, const double* y) {
    double s = 0.0;
    for (int i = 0; i < n; i++) s = fma(x[i], y[i], s);
    return s;
}

/* ------------------------------------------------------------------ */
/* CG                                                                  */
/* ------------------------------------------------------------------ */

static void apply_op(const orc_csr* A, int grid, int op, const double* x, double* y) {
    if (op == 1)
        orc_stencil5_spmv(A->row_ptr, A->col_indices, A->values, x, y, A->nb_rows, grid);
    else
        orc_csr_spmv(A->row_ptr, A->col_indices, A->values, x, y, A->nb_rows);
}

/* cg_solve_device: src/solvers/cg_solver.cu:436-706.
 *   r = 1*b + (-1)*Ap           axpby_kernel :505      fma(1,b,(-1)*Ap)
 *   p = r                       copy_kernel :512
 *   rr_old = dot(r,r)           :516-517 ; b_norm = sqrt(rr_old) :527-528 (it is ||r0||)
 *   loop (:538): Ap = A p; pAp = dot(Ap,p) :550-551; alpha = rr_old/pAp :560;
 *     x += alpha p :564 fma(alpha,p,x);  r -= alpha Ap :574 fma(-alpha,Ap,r);
 *     rr_new = dot(r,r) :584-585; residual = sqrt(rr_new); converged = residual/b_norm < tol
 *     :424-431,594-595; if converged { iter++; break } :613-621;
 *     beta = rr_new/rr_old :624; p = r + beta p :628 fma(beta,p,r); rr_old = rr_new :637.
 *   stats: iterations = iter, residual_norm = last residual (initially b_norm :536),
 *   converged recomputed on the host :656, checksums in index order :658-665 (host code: the
 *   x86-64 host compiler has no FMA at its default -march, so mul then add). */
int orc_cg_device(const orc_csr* A, int grid, int op, const double* b, double* x, int max_iters,
                  double tol, orc_cg_result* res, double* rel_hist, int rel_hist_cap) {
    int n = A->nb_rows;
    double* r = (double*)malloc((size_t)n * sizeof(double));
    double* p = (double*)malloc((size_t)n * sizeof(double));
    double* Ap = (double*)malloc((size_t)n * sizeof(double));
    if (!r || !p || !Ap) return 1;
    apply_op(A, grid, op, x, Ap);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) { r[i] = fma(1.0, b[i], -1.0 * Ap[i]); p[i] = r[i]; }
    double rr_old = orc_dot_blocktree(n, r, r);
    double b_norm = sqrt(rr_old);
    double final_res = b_norm;
    int iter;
    for (iter = 0; iter < max_iters; iter++) {
        apply_op(A, grid, op, p, Ap);
        double pAp = orc_dot_blocktree(n, Ap, p);
        double alpha = rr_old / pAp;
#pragma omp parallel for schedule(static)
        for (int i = 0; i < n; i++) { x[i] = fma(alpha, p[i], x[i]); r[i] = fma(-alpha, Ap[i], r[i]); }
        double rr_new = orc_dot_blocktree(n, r, r);
        double resn = sqrt(rr_new);
        final_res = resn;
        if (rel_hist && iter < rel_hist_cap) rel_hist[iter] = resn / b_norm;
        if (resn / b_norm < tol) { iter++; break; }
        double beta = rr_new / rr_old;
#pragma omp parallel for schedule(static)
        for (int i = 0; i < n; i++) p[i] = fma(beta, p[i], r[i]);
        rr_old = rr_new;
    }
    res->iterations = iter;
    res->residual_norm = final_res;
    res->b_norm = b_norm;
    res->converged = (final_res / b_norm < tol) ? 1 : 0;
    double s = 0.0, s2 = 0.0;
    for (int i = 0; i < n; i++) { s += x[i]; s2 += x[i] * x[i]; }
    res->solution_sum = s;
    res->solution_norm = sqrt(s2);
    free(r); free(p); free(Ap);
    return 0;
}

/* ------------------------------------------------------------------ */
/* partition                                                           */
/* ------------------------------------------------------------------ */

/* Jacobi-preconditioned CG.  NOT in the reference (its README / cg_solver.h:6-7 list preconditioning
 * as the next step; parity unpinned): the textbook recurrence laid over the reference's CG conventions
 * (cg_solver.cu:498-638): r0 = b - A x0, z = D^-1 r, p0 = z0, rho = r.z; per iteration alpha = rho / p.Ap,
 * x += alpha p, r -= alpha Ap, stop when ||r|| / ||r0|| < tol (checked before the p update, iterations
 * counted like cg_solve_device), z = D^-1 r, beta = rho_new / rho, p = z + beta p.  D = diag(A). */
int orc_pcg_device(const orc_csr* A, int grid, int op, const double* b, double* x, int max_iters,
                   double tol, orc_cg_result* res) {
    int n = A->nb_rows;
    double* r = (double*)malloc((size_t)n * sizeof(double));
    double* z = (double*)malloc((size_t)n * sizeof(double));
    double* p = (double*)malloc((size_t)n * sizeof(double));
    double* Ap = (double*)malloc((size_t)n * sizeof(double));
    double* dinv = (double*)malloc((size_t)n * sizeof(double));
    if (!r || !z || !p || !Ap || !dinv) return 1;
    for (int i = 0; i < n; i++) {
        double d = 0.0;
        for (int k = A->row_ptr[i]; k < A->row_ptr[i + 1]; k++)
            if (A->col_indices[k] == i) d = A->values[k];
        if (d == 0.0) { free(r); free(z); free(p); free(Ap); free(dinv); return 2; }
        dinv[i] = 1.0 / d;
    }
    apply_op(A, grid, op, x, Ap);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) { r[i] = fma(1.0, b[i], -1.0 * Ap[i]); z[i] = dinv[i] * r[i]; p[i] = z[i]; }
    double rr0 = orc_dot_blocktree(n, r, r);
    double rho = orc_dot_blocktree(n, r, z);
    double b_norm = sqrt(rr0);
    double final_res = b_norm;
    int iter;
    for (iter = 0; iter < max_iters; iter++) {
        apply_op(A, grid, op, p, Ap);
        double pAp = orc_dot_blocktree(n, Ap, p);
        double alpha = rho / pAp;
#pragma omp parallel for schedule(static)
        for (int i = 0; i < n; i++) {
            x[i] = fma(alpha, p[i], x[i]);
            r[i] = fma(-alpha, Ap[i], r[i]);
            z[i] = dinv[i] * r[i];
        }
        double rr_new = orc_dot_blocktree(n, r, r);
        double rho_new = orc_dot_blocktree(n, r, z);
        double resn = sqrt(rr_new);
        final_res = resn;
        if (resn / b_norm < tol) { iter++; break; }
        double beta = rho_new / rho;
#pragma omp parallel for schedule(static)
        for (int i = 0; i < n; i++) p[i] = fma(beta, p[i], z[i]);
        rho = rho_new;
    }
    res->iterations = iter;
    res->residual_norm = final_res;
    res->b_norm = b_norm;
    res->converged = (final_res / b_norm < tol) ? 1 : 0;
    double s = 0.0, s2 = 0.0;
    for (int i = 0; i < n; i++) { s += x[i]; s2 += x[i] * x[i]; }
    res->solution_sum = s;
    res->solution_norm = sqrt(s2);
    free(r); free(z); free(p); free(Ap); free(dinv);
    return 0;
}


/* Block-Jacobi-preconditioned CG with LINE blocks.  NOT in the reference (parity unpinned, like
 * orc_pcg_device whose conventions it shares): M = the tridiagonal part (W, C, E) of A inside every grid
 * row, clipped to the row bands of a P-rank partition (orc_partition), so that the multi-GPU solver and
 * this restatement use the same blocks.  z = M^-1 r: Thomas algorithm per block, operations contracted
 * the way nvcc contracts the device code (m_j = a_j / d'_{j-1}; d'_j = fma(-m_j, c_{j-1}, d_j);
 * y_j = fma(-m_j, y_{j-1}, r_j); z_j = fma(-c_j, z_{j+1}, y_j) * (1 / d'_j)).
 * Loop: alpha = rho / p.Ap; x += alpha p; r -= alpha Ap; stop on ||r|| / ||r0|| < tol; z = M^-1 r;
 * beta = r.z / rho; p = z + beta p. */
static double csr_coeff(const orc_csr* A, int row, int col) {
    double v = 0.0;
    for (int k = A->row_ptr[row]; k < A->row_ptr[row + 1]; k++)
        if (A->col_indices[k] == col) v = A->values[k];
    return v;
}

int orc_pcg_block_device(const orc_csr* A, int grid, int op, int P, const double* b, double* x, int max_iters,
                         double tol, orc_cg_result* res) {
    int n = A->nb_rows;
    if (grid < 1 || (long long)grid * grid != n || P < 1) return 3;
    double* r = (double*)malloc((size_t)n * sizeof(double));
    double* z = (double*)malloc((size_t)n * sizeof(double));
    double* p = (double*)malloc((size_t)n * sizeof(double));
    double* Ap = (double*)malloc((size_t)n * sizeof(double));
    double* m = (double*)malloc((size_t)n * sizeof(double));
    double* invd = (double*)malloc((size_t)n * sizeof(double));
    double* c = (double*)malloc((size_t)n * sizeof(double));
    int* blk_lo = (int*)malloc(((size_t)grid + 2 * (size_t)P + 2) * sizeof(int));
    if (!r || !z || !p || !Ap || !m || !invd || !c || !blk_lo) return 1;
    /* blocks: (grid row) x (band) intersections, in row order */
    int nblk = 0;
    for (int g = 0; g < P; g++) {
        int64_t nl, off;
        orc_partition(n, P, g, &nl, &off);
        int64_t s = off;
        while (s < off + nl) {
            int64_t e = (s / grid + 1) * grid;
            if (e > off + nl) e = off + nl;
            blk_lo[nblk++] = (int)s;
            s = e;
        }
    }
    blk_lo[nblk] = n;
    int bad = 0;
#pragma omp parallel for schedule(static)
    for (int bi = 0; bi < nblk; bi++) {
        double dprev = 1.0, cprev = 0.0;
        for (int i = blk_lo[bi]; i < blk_lo[bi + 1]; i++) {
            double d = csr_coeff(A, i, i);
            double a = i > blk_lo[bi] ? csr_coeff(A, i, i - 1) : 0.0;
            double cc = i + 1 < blk_lo[bi + 1] ? csr_coeff(A, i, i + 1) : 0.0;
            double mj = i > blk_lo[bi] ? a / dprev : 0.0;
            double dj = fma(-mj, cprev, d);
            if (dj == 0.0 || d == 0.0) bad = 1;
            m[i] = mj; invd[i] = 1.0 / dj; c[i] = cc;
            dprev = dj; cprev = cc;
        }
    }
    if (bad) { free(r); free(z); free(p); free(Ap); free(m); free(invd); free(c); free(blk_lo); return 2; }
#define ORC_BJ_SOLVE()                                                           \
    _Pragma("omp parallel for schedule(static)")                                 \
    for (int bi = 0; bi < nblk; bi++) {                                          \
        double y = 0.0;                                                          \
        for (int i = blk_lo[bi]; i < blk_lo[bi + 1]; i++) { y = fma(-m[i], y, r[i]); z[i] = y; } \
        double zn = 0.0;                                                         \
        for (int i = blk_lo[bi + 1] - 1; i >= blk_lo[bi]; i--) { zn = fma(-c[i], zn, z[i]) * invd[i]; z[i] = zn; } \
    }
    apply_op(A, grid, op, x, Ap);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) r[i] = fma(1.0, b[i], -1.0 * Ap[i]);
    ORC_BJ_SOLVE();
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) p[i] = z[i];
    double rr0 = orc_dot_blocktree(n, r, r);
    double rho = orc_dot_blocktree(n, r, z);
    double b_norm = sqrt(rr0);
    double final_res = b_norm;
    int iter;
    for (iter = 0; iter < max_iters; iter++) {
        apply_op(A, grid, op, p, Ap);
        double pAp = orc_dot_blocktree(n, Ap, p);
        double alpha = rho / pAp;
#pragma omp parallel for schedule(static)
        for (int i = 0; i < n; i++) { x[i] = fma(alpha, p[i], x[i]); r[i] = fma(-alpha, Ap[i], r[i]); }
        double rr_new = orc_dot_blocktree(n, r, r);
        double resn = sqrt(rr_new);
        final_res = resn;
        if (resn / b_norm < tol) { iter++; break; }
        ORC_BJ_SOLVE();
        double rho_new = orc_dot_blocktree(n, r, z);
        double beta = rho_new / rho;
#pragma omp parallel for schedule(static)
        for (int i = 0; i < n; i++) p[i] = fma(beta, p[i], z[i]);
        rho = rho_new;
    }
#undef ORC_BJ_SOLVE
    res->iterations = iter;
    res->residual_norm = final_res;
    res->b_norm = b_norm;
    res->converged = (final_res / b_norm < tol) ? 1 : 0;
    double s = 0.0, s2 = 0.0;
    for (int i = 0; i < n; i++) { s += x[i]; s2 += x[i] * x[i]; }
    res->solution_sum = s;
    res->solution_norm = sqrt(s2);
    free(r); free(z); free(p); free(Ap); free(m); free(invd); free(c); free(blk_lo);
    return 0;
}

/* src/solvers/cg_solver_mgpu_partitioned.cu:262-268: n_local = N / P (matrix rows),
 * row_offset = g * n_local, the last rank takes N - row_offset. */
void orc_partition(int64_t N, int P, int g, int64_t* n_local, int64_t* row_offset) {
    int64_t nl = N / P;
    int64_t off = (int64_t)g * nl;
    if (g == P - 1) nl = N - off;
    *n_local = nl; *row_offset = off;
}

/* :306-329: local nnz = rp[off+n_local]-rp[off]; row_ptr rebased to 0; col stays GLOBAL. */
int64_t orc_local_csr_slice(const orc_csr* A, int64_t off, int64_t nl, int* rp_out, int* col_out,
                            double* val_out) {
    int base = A->row_ptr[off];
    int64_t lnnz = A->row_ptr[off + nl] - base;
    for (int64_t i = 0; i <= nl; i++) rp_out[i] = A->row_ptr[off + i] - base;
    for (int64_t k = 0; k < lnnz; k++) { col_out[k] = A->col_indices[base + k]; val_out[k] = A->values[base + k]; }
    return lnnz;
}

/* :697-703 (and :450-456 / :489-495 for the setup exchanges): rank g sends its first `grid`
 * local elements to g-1 and its last `grid` local elements to g+1. */
void orc_halo_ranges(int64_t nl, int grid, int g, int P, int64_t* plo, int64_t* phi, int64_t* nlo,
                     int64_t* nhi) {
    *plo = *phi = *nlo = *nhi = 0;
    if (g > 0) { *plo = 0; *phi = grid; }
    if (g < P - 1) { *nlo = nl - grid; *nhi = nl; }
}

/* cg_solve_mgpu_partitioned: src/solvers/cg_solver_mgpu_partitioned.cu:236-908, with the P MPI
 * ranks run as P virtual ranks in one loop nest.
 *   setup: halo(x) :450-456; Ap = haloSpMV(x) :467-469; b += (-1)*Ap (axpy, fma(-1,Ap,b)) :475;
 *   r = b :476; halo(r) :489-495; p = r, p_halo = r_halo :505-516;
 *   rs_old = sum_g dot(r,r) :522-531; b_norm = sqrt :533.
 *   loop :542: Ap = haloSpMV(p) :550-552; pAp = sum_g dot(p,Ap) :567,583; alpha = rs_old/pAp
 *   :592; x = alpha*p + x :599; r = (-alpha)*Ap + r :613; rs_new :629,645;
 *   if sqrt(rs_new)/b_norm < tol { iter++; break } :654-670; beta = rs_new/rs_old :673;
 *   p = 1*r + beta*p (axpby_kernel :135-140 => fma(1,r,beta*p), TWO roundings, unlike the
 *   single-GPU update_p_kernel) :680; halo(p) :697-703; rs_old = rs_new :713.
 * Dots: cublasDdot + MPI_Allreduce -- order unpinned; here sequential per rank, ranks summed in
 * rank order. */
int orc_cg_mgpu(const orc_csr* A, int grid, int P, const double* b, double* x, int max_iters,
                double tol, orc_cg_result* res) {
    int64_t N = A->nb_rows;
    int64_t* nl = (int64_t*)malloc(sizeof(int64_t) * P);
    int64_t* off = (int64_t*)malloc(sizeof(int64_t) * P);
    int** rp = (int**)malloc(sizeof(int*) * P);
    int** cl = (int**)malloc(sizeof(int*) * P);
    double** vl = (double**)malloc(sizeof(double*) * P);
    for (int g = 0; g < P; g++) {
        orc_partition(N, P, g, &nl[g], &off[g]);
        if (P > 1 && nl[g] < grid) return 2; /* halo rule needs n_local >= grid */
        int64_t lnnz = A->row_ptr[off[g] + nl[g]] - A->row_ptr[off[g]];
        rp[g] = (int*)malloc(sizeof(int) * (size_t)(nl[g] + 1));
        cl[g] = (int*)malloc(sizeof(int) * (size_t)(lnnz > 0 ? lnnz : 1));
        vl[g] = (double*)malloc(sizeof(double) * (size_t)(lnnz > 0 ? lnnz : 1));
        orc_local_csr_slice(A, off[g], nl[g], rp[g], cl[g], vl[g]);
    }
    double* r = (double*)malloc(sizeof(double) * (size_t)N);
    double* p = (double*)malloc(sizeof(double) * (size_t)N);
    double* Ap = (double*)malloc(sizeof(double) * (size_t)N);
    double* bb = (double*)malloc(sizeof(double) * (size_t)N);
    memcpy(bb, b, sizeof(double) * (size_t)N);
#define HALO_PREV(v, g) ((g) > 0 ? (v) + off[g] - grid : NULL)
#define HALO_NEXT(v, g) ((g) < P - 1 ? (v) + off[g] + nl[g] : NULL)
    /* with all bands in one address space the halo buffers are simply the neighbouring
     * `grid` elements of the global vector */
    for (int g = 0; g < P; g++)
        orc_halo_spmv(rp[g], cl[g], vl[g], x + off[g], HALO_PREV(x, g), HALO_NEXT(x, g),
                      Ap + off[g], (int)nl[g], off[g], N, grid);
    for (int64_t i = 0; i < N; i++) { bb[i] = fma(-1.0, Ap[i], bb[i]); r[i] = bb[i]; p[i] = r[i]; }
    double rs_old = 0.0;
    for (int g = 0; g < P; g++) rs_old += orc_dot_sequential((int)nl[g], r + off[g], r + off[g]);
    double b_norm = sqrt(rs_old), final_res = b_norm;
    int iter;
    for (iter = 0; iter < max_iters; iter++) {
        for (int g = 0; g < P; g++)
            orc_halo_spmv(rp[g], cl[g], vl[g], p + off[g], HALO_PREV(p, g), HALO_NEXT(p, g),
                          Ap + off[g], (int)nl[g], off[g], N, grid);
        double pAp = 0.0;
        for (int g = 0; g < P; g++) pAp += orc_dot_sequential((int)nl[g], p + off[g], Ap + off[g]);
        double alpha = rs_old / pAp;
        for (int64_t i = 0; i < N; i++) { x[i] = fma(alpha, p[i], x[i]); r[i] = fma(-alpha, Ap[i], r[i]); }
        double rs_new = 0.0;
        for (int g = 0; g < P; g++) rs_new += orc_dot_sequential((int)nl[g], r + off[g], r + off[g]);
        final_res = sqrt(rs_new);
        if (final_res / b_norm < tol) { iter++; break; }
        double beta = rs_new / rs_old;
        for (int64_t i = 0; i < N; i++) p[i] = fma(1.0, r[i], beta * p[i]);
        rs_old = rs_new;
    }
#undef HALO_PREV
#undef HALO_NEXT
    res->iterations = iter; res->residual_norm = final_res; res->b_norm = b_norm;
    res->converged = (final_res / b_norm < tol) ? 1 : 0;
    double s = 0.0, s2 = 0.0;
    for (int64_t i = 0; i < N; i++) { s += x[i]; s2 += x[i] * x[i]; }
    res->solution_sum = s; res->solution_norm = sqrt(s2);
    for (int g = 0; g < P; g++) { free(rp[g]); free(cl[g]); free(vl[g]); }
    free(rp); free(cl); free(vl); free(nl); free(off); free(r); free(p); free(Ap); free(bb);
    return 0;
}

/* ------------------------------------------------------------------ */
/* bench statistics: src/spmv/benchmark_stats.cu:8-89                  */
/* ------------------------------------------------------------------ */
static int cmp_d(const void* a, const void* b) {
    double x = *(const double*)a, y = *(const double*)b;
    return (x > y) - (x < y);
}

int orc_bench_stats_from_times(const double* times, int n, orc_bench_stats* st) {
    if (n < 3) return -1;
    double mean = 0.0;
    for (int i = 0; i < n; i++) mean += times[i];
    mean /= n;
    double ss = 0.0;
    for (int i = 0; i < n; i++) { double d = times[i] - mean; ss += d * d; }
    double sd = sqrt(ss / n);
    double* f = (double*)malloc(sizeof(double) * n);
    int fc = 0;
    for (int i = 0; i < n; i++) if (fabs(times[i] - mean) <= 2.0 * sd) f[fc++] = times[i];
    double m2 = 0.0;
    for (int i = 0; i < fc; i++) m2 += f[i];
    m2 /= fc;
    double ss2 = 0.0;
    for (int i = 0; i < fc; i++) { double d = f[i] - m2; ss2 += d * d; }
    st->mean_ms = m2; st->std_dev_ms = sqrt(ss2 / fc);
    qsort(f, fc, sizeof(double), cmp_d);
    st->median_ms = (fc % 2 == 0) ? (f[fc / 2 - 1] + f[fc / 2]) / 2.0 : f[fc / 2];
    st->min_ms = f[0]; st->max_ms = f[fc - 1];
    st->valid_runs = fc; st->outliers_removed = n - fc;
    free(f);
    return 0;
}
