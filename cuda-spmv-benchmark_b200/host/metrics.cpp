// metrics.cpp -- performance metrics, GPU/system probe and the JSON / CSV exporters.
//
// The OUTPUT SCHEMA is the contract (the reference's scripts grep these files): key names, nesting,
// key order and printf formats follow reference src/spmv/spmv_metrics.cu:190-324 (SpMV JSON/CSV)
// and src/solvers/cg_metrics.cu:20-185 (CG JSON/CSV).  The writer itself is a small table-driven
// emitter.  calculate_spmv_metrics keeps the reference's traffic model (spmv_metrics.cu:46-102) so
// "bandwidth_gb_s" stays comparable with published numbers; the roofline figure in bench.py uses
// the algorithmic bytes instead.
#include <time.h>
#include <unistd.h>

#include <string>

#include "host_common.h"

namespace {

// ---- tiny JSON object writer: 2-space indentation, caller-controlled formats
class Json {
  public:
    explicit Json(FILE* f) : f_(f) {}
    void open(const char* key = nullptr) {
        sep();
        pad();
        if (key) fprintf(f_, "\"%s\": {\n", key);
        else fprintf(f_, "{\n");
        depth_++;
        first_ = true;
    }
    void close() {
        fprintf(f_, "\n");
        depth_--;
        pad();
        fprintf(f_, "}");
        first_ = false;
        if (depth_ == 0) fprintf(f_, "\n");
    }
    void str(const char* k, const char* v) { key(k); fprintf(f_, "\"%s\"", v); }
    void raw(const char* k, const char* v) { key(k); fputs(v, f_); }
    void i(const char* k, long long v) { key(k); fprintf(f_, "%lld", v); }
    void f(const char* k, const char* fmt, double v) { key(k); fprintf(f_, fmt, v); }

  private:
    void sep() { if (!first_) fprintf(f_, ",\n"); first_ = false; }
    void pad() { for (int d = 0; d < depth_; d++) fputs("  ", f_); }
    void key(const char* k) { sep(); pad(); fprintf(f_, "\"%s\": ", k); }
    FILE* f_;
    int depth_ = 0;
    bool first_ = true;
};

struct Traffic {
    double data, indices, vectors;
};

// reference traffic model: CSR operators are charged values + col_idx + row_ptr from csr_mat,
// anything else nnz*8 + 2*nnz*4 (spmv_metrics.cu:72-95)
Traffic model_traffic(const char* op, int rows, int cols, int nnz) {
    Traffic t;
    const bool csr_like = !strcmp(op, "cusparse-csr") || !strcmp(op, "stencil5-csr");
    if (csr_like) {
        const double n = csr_mat.row_ptr ? csr_mat.nb_nonzeros : nnz, r = csr_mat.row_ptr ? csr_mat.nb_rows : rows;
        t.data = n * 8.0;
        t.indices = n * 4.0 + (r + 1) * 4.0;
    } else {
        t.data = nnz * 8.0;
        t.indices = nnz * 4.0 * 2;
    }
    t.vectors = ((double)rows + cols) * 8.0;
    return t;
}

const char* bound_of(double ai) { return ai < 0.5 ? "memory-bound" : (ai < 2.0 ? "balanced" : "compute-bound"); }

void cpu_model(char* out, size_t cap) {
    snprintf(out, cap, "Unknown");
    FILE* f = fopen("/proc/cpuinfo", "r");
    if (!f) return;
    char line[256];
    while (fgets(line, sizeof line, f)) {
        if (strncmp(line, "model name", 10) == 0) {
            const char* c = strchr(line, ':');
            if (c) {
                c += (c[1] == ' ') ? 2 : 1;
                snprintf(out, cap, "%s", c);
                char* nl = strchr(out, '\n');
                if (nl) *nl = 0;
            }
            break;
        }
    }
    fclose(f);
}

void smi_probe(BenchmarkMetrics* m) {
    FILE* p = popen("nvidia-smi --query-gpu=temperature.gpu,temperature.memory,power.draw,power.limit,"
                    "persistence_mode,pcie.link.gen.current,pcie.link.width.current "
                    "--format=csv,noheader,nounits 2>/dev/null", "r");
    if (!p) return;
    char line[256];
    if (fgets(line, sizeof line, p)) {
        double tg = 0, tm = 0, pd = 0, pl = 0;
        char pers[32] = "";
        int gen = 0, width = 0;
        // temperature.memory may read "N/A": parse field by field
        char* save = nullptr;
        int fld = 0;
        for (char* tok = strtok_r(line, ",", &save); tok; tok = strtok_r(nullptr, ",", &save), fld++) {
            while (*tok == ' ') tok++;
            switch (fld) {
                case 0: tg = atof(tok); break;
                case 1: tm = atof(tok); break;
                case 2: pd = atof(tok); break;
                case 3: pl = atof(tok); break;
                case 4: sscanf(tok, "%31s", pers); break;
                case 5: gen = atoi(tok); break;
                case 6: width = atoi(tok); break;
            }
        }
        m->gpu_info.current_temp_c = (int)tg;
        m->gpu_info.max_temp_c = (int)(tm > tg ? tm : tg);
        m->gpu_info.power_draw_w = (int)pd;
        m->gpu_info.power_limit_w = (int)pl;
        snprintf(m->gpu_info.persistence_mode, sizeof m->gpu_info.persistence_mode, "%s", pers);
        snprintf(m->gpu_info.pcie_generation, sizeof m->gpu_info.pcie_generation, "Gen%d", gen);
        m->gpu_info.pcie_link_width = width;
    }
    pclose(p);
}

std::string now_stamp() {
    char buf[64];
    time_t t = time(nullptr);
    strftime(buf, sizeof buf, "%Y-%m-%d %H:%M:%S", localtime(&t));
    return buf;
}

// reference cg_metrics.cu:68 divides by time_spmv_ms unconditionally ("inf" in the JSON when the timers
// are off); the engine always fills that field from the device clock, the guard keeps the JSON valid anyway
double spmv_gflops(const MatrixData* mat, int iterations, double spmv_ms) {
    return spmv_ms > 0.0 ? (2.0 * mat->nnz * iterations) / (spmv_ms * 1e6) : 0.0;
}

template <class Stats>
void cg_json_common(Json& j, const char* solver, const char* mode, int num_gpus, const MatrixData* mat,
                    const BenchmarkStats* bs, const Stats* cs, double allreduce_ms, double allgather_ms) {
    j.open();
    j.str("timestamp", now_stamp().c_str());
    j.str("solver", solver);
    j.str("mode", mode);
    if (num_gpus > 0) j.i("num_gpus", num_gpus);
    j.open("matrix");
    j.i("rows", mat->rows); j.i("cols", mat->cols); j.i("nnz", mat->nnz); j.i("grid_size", mat->grid_size);
    j.close();
    j.open("convergence");
    j.raw("converged", cs->converged ? "true" : "false");
    j.i("iterations", cs->iterations);
    j.f("residual_norm", "%.15e", cs->residual_norm);
    j.close();
    j.open("timing");
    j.f("median_ms", "%.3f", bs->median_ms); j.f("mean_ms", "%.3f", bs->mean_ms);
    j.f("min_ms", "%.3f", bs->min_ms); j.f("max_ms", "%.3f", bs->max_ms);
    j.f("std_dev_ms", "%.3f", bs->std_dev_ms);
    j.f("spmv_ms", "%.3f", cs->time_spmv_ms); j.f("blas1_ms", "%.3f", cs->time_blas1_ms);
    j.f("reductions_ms", "%.3f", cs->time_reductions_ms);
    if (num_gpus > 0) { j.f("allreduce_ms", "%.3f", allreduce_ms); j.f("allgather_ms", "%.3f", allgather_ms); }
    j.close();
    j.open("statistics");
    j.i("valid_runs", bs->valid_runs); j.i("outliers_removed", bs->outliers_removed);
    j.close();
    j.open("performance");
    j.f("gflops_spmv", "%.3f", spmv_gflops(mat, cs->iterations, cs->time_spmv_ms));
    j.close();
    j.open("validation");
    j.f("solution_sum", "%.16e", cs->solution_sum); j.f("solution_norm", "%.16e", cs->solution_norm);
    j.close();
    j.close();
}

}  // namespace

extern "C" void calculate_spmv_metrics(double execution_time_ms, const MatrixData* mat, const char* operator_name,
                                       BenchmarkMetrics* m) {
    m->matrix_rows = mat->rows; m->matrix_cols = mat->cols; m->matrix_nnz = mat->nnz;
    m->grid_size = mat->grid_size;
    m->execution_time_ms = execution_time_ms;
    m->operator_name = operator_name;
    m->sparsity_ratio = (double)mat->nnz / ((double)mat->rows * mat->cols);
    const double secs = execution_time_ms / 1000.0;
    m->gflops = (2.0 * mat->nnz / secs) / 1e9;
    const Traffic t = model_traffic(operator_name, mat->rows, mat->cols, mat->nnz);
    m->bandwidth_gb_s = ((t.data + t.indices + t.vectors) / secs) / 1e9;
}

extern "C" int get_gpu_properties(BenchmarkMetrics* m) {
    int dev = 0;
    cudaDeviceProp prop;
    B200_CUDA(cudaGetDevice(&dev));
    B200_CUDA(cudaGetDeviceProperties(&prop, dev));
    snprintf(m->gpu_info.name, sizeof m->gpu_info.name, "%.*s", (int)sizeof m->gpu_info.name - 1, prop.name);
    m->gpu_info.memory_mb = (int)(prop.totalGlobalMem / (1024 * 1024));
    snprintf(m->gpu_info.compute_capability, sizeof m->gpu_info.compute_capability, "%d.%d", prop.major, prop.minor);
    m->gpu_info.multiprocessor_count = prop.multiProcessorCount;
    m->gpu_info.max_threads_per_block = prop.maxThreadsPerBlock;
    int khz = 0;
    m->gpu_info.memory_clock_khz = cudaDeviceGetAttribute(&khz, cudaDevAttrMemoryClockRate, dev) == cudaSuccess ? khz : 0;
    m->gpu_info.graphics_clock_mhz = cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev) == cudaSuccess ? khz / 1000 : 0;
    B200_CUDA(cudaRuntimeGetVersion(&m->gpu_info.cuda_runtime_version));
    B200_CUDA(cudaDriverGetVersion(&m->gpu_info.cuda_driver_version));
    m->gpu_info.cusparse_version = 0;  // not linked
    cpu_model(m->gpu_info.cpu_model, sizeof m->gpu_info.cpu_model);
    m->gpu_info.system_ram_gb = (int)(((long long)sysconf(_SC_PHYS_PAGES) * sysconf(_SC_PAGE_SIZE)) >> 30);
    smi_probe(m);
    return 0;
}

extern "C" void print_benchmark_metrics(const BenchmarkMetrics* m, FILE* output_file) {
    FILE* fp = output_file ? output_file : stdout;
    const double flops = 2.0 * m->matrix_nnz;
    const double bytes = m->matrix_nnz * 8.0 + m->matrix_nnz * 4.0 * 2 + ((double)m->matrix_rows + m->matrix_cols) * 8.0;
    const double ai = flops / bytes;
    fprintf(fp, "\n=== SpMV Performance Metrics ===\n");
    fprintf(fp, "Operator: %s\n", m->operator_name);
    fprintf(fp, "\n--- Matrix Characteristics ---\n");
    if (m->grid_size > 0) {
        fprintf(fp, "Grid size: %d x %d (2D stencil)\n", m->grid_size, m->grid_size);
        fprintf(fp, "Matrix dimensions: %d x %d (grid²)\n", m->matrix_rows, m->matrix_cols);
    } else {
        fprintf(fp, "Matrix dimensions: %d x %d\n", m->matrix_rows, m->matrix_cols);
    }
    fprintf(fp, "Non-zeros: %d\n", m->matrix_nnz);
    fprintf(fp, "Sparsity ratio: %.6f (%.4f%% non-zero)\n", m->sparsity_ratio, m->sparsity_ratio * 100.0);
    fprintf(fp, "\n--- Performance Metrics ---\n");
    fprintf(fp, "Execution time: %.3f ms (%.1f μs)\n", m->execution_time_ms, m->execution_time_ms * 1000.0);
    fprintf(fp, "GFLOPS: %.3f\n", m->gflops);
    fprintf(fp, "Memory bandwidth: %.3f GB/s\n", m->bandwidth_gb_s);
    fprintf(fp, "\n--- Performance Analysis ---\n");
    fprintf(fp, "Arithmetic intensity: %.3f FLOP/byte\n", ai);
    if (ai < 0.5) {
        fprintf(fp, "Classification: Memory-bound (low arithmetic intensity)\n");
        fprintf(fp, "Optimization focus: Memory access patterns, data locality\n");
    } else if (ai < 2.0) {
        fprintf(fp, "Classification: Balanced compute/memory\n");
        fprintf(fp, "Optimization focus: Both compute and memory optimization\n");
    } else {
        fprintf(fp, "Classification: Compute-bound (high arithmetic intensity)\n");
        fprintf(fp, "Optimization focus: Compute throughput, parallelization\n");
    }
    fprintf(fp, "=============================\n\n");
}

extern "C" void print_metrics_json(const BenchmarkMetrics* m, FILE* output_file) {
    FILE* fp = output_file ? output_file : stdout;
    const Traffic t = model_traffic(m->operator_name, m->matrix_rows, m->matrix_cols, m->matrix_nnz);
    const double flops = 2.0 * m->matrix_nnz, bytes = t.data + t.indices + t.vectors, ai = flops / bytes;
    char dims[64];
    snprintf(dims, sizeof dims, "%dx%d", m->grid_size, m->grid_size);
    Json j(fp);
    j.open();
    j.open("gpu");
    j.str("name", m->gpu_info.name);
    j.i("memory_mb", m->gpu_info.memory_mb);
    j.str("compute_capability", m->gpu_info.compute_capability);
    j.i("multiprocessor_count", m->gpu_info.multiprocessor_count);
    j.i("memory_clock_khz", m->gpu_info.memory_clock_khz);
    j.i("graphics_clock_mhz", m->gpu_info.graphics_clock_mhz);
    j.i("cuda_runtime_version", m->gpu_info.cuda_runtime_version);
    j.i("cuda_driver_version", m->gpu_info.cuda_driver_version);
    j.i("cusparse_version", m->gpu_info.cusparse_version);
    j.i("current_temp_c", m->gpu_info.current_temp_c);
    j.i("power_draw_w", m->gpu_info.power_draw_w);
    j.i("power_limit_w", m->gpu_info.power_limit_w);
    j.str("persistence_mode", m->gpu_info.persistence_mode);
    j.str("pcie_generation", m->gpu_info.pcie_generation);
    j.i("pcie_link_width", m->gpu_info.pcie_link_width);
    j.close();
    j.open("system");
    j.str("cpu_model", m->gpu_info.cpu_model);
    j.i("system_ram_gb", m->gpu_info.system_ram_gb);
    j.close();
    j.open("benchmark");
    j.str("operator", m->operator_name);
    j.open("matrix");
    if (m->grid_size > 0) { j.i("grid_size", m->grid_size); j.str("grid_dimensions", dims); }
    j.i("rows", m->matrix_rows); j.i("cols", m->matrix_cols); j.i("nnz", m->matrix_nnz);
    j.f("sparsity_ratio", "%.6f", m->sparsity_ratio);
    j.f("sparsity_percent", "%.4f", m->sparsity_ratio * 100.0);
    j.close();
    j.open("performance");
    j.f("execution_time_ms", "%.6f", m->execution_time_ms);
    j.f("execution_time_us", "%.1f", m->execution_time_ms * 1000.0);
    j.f("gflops", "%.6f", m->gflops);
    j.f("bandwidth_gb_s", "%.6f", m->bandwidth_gb_s);
    j.close();
    j.open("analysis");
    j.f("arithmetic_intensity", "%.6f", ai);
    j.f("total_flops", "%.0f", flops); j.f("total_bytes", "%.0f", bytes);
    j.f("matrix_data_bytes", "%.0f", t.data); j.f("matrix_indices_bytes", "%.0f", t.indices);
    j.f("vector_bytes", "%.0f", t.vectors);
    j.str("performance_bound", bound_of(ai));
    j.close();
    j.open("validation");
    j.f("sum_y", "%.16e", m->sum_y); j.f("norm2_y", "%.16e", m->norm2_y);
    j.close();
    j.close();
    j.close();
}

extern "C" void print_metrics_csv(const BenchmarkMetrics* m, FILE* output_file) {
    FILE* fp = output_file ? output_file : stdout;
    const double flops = 2.0 * m->matrix_nnz;
    const double ai = flops / (m->matrix_nnz * 12.0 + ((double)m->matrix_rows + m->matrix_cols) * 8.0);
    static int header_done = 0;  // once per process, like the reference (spmv_metrics.cu:302-309)
    if (!header_done) {
        fputs("operator,grid_size,matrix_rows,matrix_cols,matrix_nnz,sparsity_ratio,sparsity_percent,"
              "execution_time_ms,execution_time_us,gflops,bandwidth_gb_s,"
              "arithmetic_intensity,total_flops,performance_bound,sum_y,norm2_y\n", fp);
        header_done = 1;
    }
    fprintf(fp, "%s,%d,%d,%d,%d,%.6f,%.4f,%.6f,%.1f,%.6f,%.6f,%.6f,%.0f,%s,%.16e,%.16e\n", m->operator_name,
            m->grid_size, m->matrix_rows, m->matrix_cols, m->matrix_nnz, m->sparsity_ratio, m->sparsity_ratio * 100.0,
            m->execution_time_ms, m->execution_time_ms * 1000.0, m->gflops, m->bandwidth_gb_s, ai, flops, bound_of(ai),
            m->sum_y, m->norm2_y);
}

extern "C" void export_cg_json(const char* filename, const char* mode, const MatrixData* mat, const BenchmarkStats* bs,
                               const CGStats* cs) {
    FILE* fp = fopen(filename, "w");
    if (!fp) { fprintf(stderr, "Error: Could not open %s for writing\n", filename); return; }
    Json j(fp);
    cg_json_common(j, "CG", mode, 0, mat, bs, cs, 0.0, 0.0);
    fclose(fp);
    printf("Results exported to: %s\n", filename);
}

extern "C" void export_cg_mgpu_json(const char* filename, const char* mode, const MatrixData* mat,
                                    const BenchmarkStats* bs, const CGStatsMultiGPU* cs, int num_gpus) {
    FILE* fp = fopen(filename, "w");
    if (!fp) { fprintf(stderr, "Error: Could not open %s for writing\n", filename); return; }
    Json j(fp);
    cg_json_common(j, "CG Multi-GPU", mode, num_gpus, mat, bs, cs, cs->time_allreduce_ms, cs->time_allgather_ms);
    fclose(fp);
    printf("Results exported to: %s\n", filename);
}

extern "C" void export_cg_csv(const char* filename, const char* mode, const MatrixData* mat, const BenchmarkStats* bs,
                              const CGStats* cs, bool write_header) {
    FILE* fp = fopen(filename, write_header ? "w" : "a");
    if (!fp) { fprintf(stderr, "Error: Could not open %s for writing\n", filename); return; }
    if (write_header)
        fputs("mode,rows,cols,nnz,grid_size,converged,iterations,residual_norm,"
              "median_ms,mean_ms,min_ms,max_ms,std_dev_ms,spmv_ms,blas1_ms,reductions_ms,"
              "valid_runs,outliers_removed,gflops_spmv,solution_sum,solution_norm\n", fp);
    fprintf(fp, "%s,%d,%d,%d,%d,%d,%d,%.15e,%.3f,%.3f,%.3f,%.3f,%.3f,%.3f,%.3f,%.3f,%d,%d,%.3f,%.16e,%.16e\n", mode,
            mat->rows, mat->cols, mat->nnz, mat->grid_size, cs->converged, cs->iterations, cs->residual_norm,
            bs->median_ms, bs->mean_ms, bs->min_ms, bs->max_ms, bs->std_dev_ms, cs->time_spmv_ms, cs->time_blas1_ms,
            cs->time_reductions_ms, bs->valid_runs, bs->outliers_removed,
            spmv_gflops(mat, cs->iterations, cs->time_spmv_ms), cs->solution_sum, cs->solution_norm);
    fclose(fp);
    if (write_header) printf("Results exported to: %s\n", filename);
}
