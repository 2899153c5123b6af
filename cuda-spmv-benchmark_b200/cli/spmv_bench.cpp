// spmv_bench <file.mtx | --grid=n> --mode=a[,b] [--json=F] [--csv=F] [--device-ingest]
// reference src/main/main.cu:44-268: validate modes, load, x = 1, per mode init -> 5 warm-ups ->
// benchmark_with_stats(10) -> checksums -> metrics -> JSON/CSV "<base>_<op><ext>" -> Sum/Norm2 lines.
#include "cli_common.h"

int main(int argc, char** argv) {
    CliArgs a = parse_cli(argc, argv);
    if ((a.matrix.empty() && a.grid <= 0) || a.modes.empty()) {
        fprintf(stderr, "Usage: %s <matrix_file.mtx | --grid=n> --mode=<mode1[,mode2,...]> [--json=<file>] [--csv=<file>]\n", argv[0]);
        fprintf(stderr, "Available modes: cusparse-csr (csr), stencil5-csr (stencil5), ellpack, stencil5-ellpack, stencil5-halo-mgpu\n");
        fprintf(stderr, "Example: %s matrix.mtx --mode=cusparse-csr,stencil5-csr --json=results.json\n", argv[0]);
        return EXIT_FAILURE;
    }
    for (auto& m : a.modes)
        if (!get_operator(m.c_str())) {
            fprintf(stderr, "Unknown mode '%s'\n", m.c_str());
            return EXIT_FAILURE;
        }
    MatrixData mat;
    void* d_entries = nullptr;
    if (load_or_generate(a, &mat, &d_entries)) return EXIT_FAILURE;
    printf("Matrix loaded: %d rows, %d cols, %d nonzeros\n", mat.rows, mat.cols, mat.nnz);
    printf("Testing %zu mode(s): ", a.modes.size());
    for (size_t i = 0; i < a.modes.size(); i++) printf("%s%s", a.modes[i].c_str(), i + 1 < a.modes.size() ? ", " : "\n");
    std::vector<double> x((size_t)mat.cols, 1.0), y((size_t)mat.rows, 0.0);
    for (auto& m : a.modes) {
        SpmvOperator* op = get_operator(m.c_str());
        printf("\n=== Testing mode: %s ===\n", m.c_str());
        if (init_operator(op, &mat, d_entries) != 0) {
            fprintf(stderr, "Failed to initialize operator '%s'\n", op->name);
            return EXIT_FAILURE;
        }
        printf("Warmup (5 runs)...\n");
        double ms = 0;
        for (int w = 0; w < 5; w++)
            if (op->run_timed(x.data(), y.data(), &ms) != 0) return EXIT_FAILURE;
        printf("Running statistical benchmark (%d iterations)...\n", a.runs);
        BenchmarkStats bs;
        if (benchmark_with_stats(op->run_timed, x.data(), y.data(), a.runs, &bs) != 0) {
            fprintf(stderr, "Statistical benchmark failed for mode '%s'\n", op->name);
            op->free();
            return EXIT_FAILURE;
        }
        printf("Completed: %d valid runs, %d outliers removed\n", bs.valid_runs, bs.outliers_removed);
        printf("Kernel time: median %.3f ms, mean %.3f ms, min %.3f, max %.3f, std %.3f\n", bs.median_ms, bs.mean_ms,
               bs.min_ms, bs.max_ms, bs.std_dev_ms);
        double sum = 0.0, sq = 0.0;  // index order, host (main.cu:177-183)
        for (int i = 0; i < mat.rows; i++) { sum += y[i]; sq += y[i] * y[i]; }
        BenchmarkMetrics mt;
        memset(&mt, 0, sizeof mt);
        calculate_spmv_metrics(bs.median_ms, &mat, op->name, &mt);
        mt.sum_y = sum;
        mt.norm2_y = sqrt(sq);
        if (get_gpu_properties(&mt) != 0) fprintf(stderr, "Warning: Could not retrieve GPU properties\n");
        print_benchmark_metrics(&mt, stdout);
        // algorithmic-byte roofline figure (stencil: values + x + y; generic CSR adds indices)
        const bool st = strstr(op->name, "stencil5") != nullptr;
        const double alg = st ? 8.0 * mat.nnz + 16.0 * mat.rows
                              : (strstr(op->name, "ellpack") ? 76.0 * mat.rows : 12.0 * mat.nnz + 4.0 * (mat.rows + 1) + 16.0 * mat.rows);
        printf("Algorithmic traffic: %.4f GB -> %.1f GB/s\n", alg / 1e9, alg / (bs.median_ms * 1e-3) / 1e9);
        if (!a.json.empty()) {
            std::string fn = per_mode_name(a.json, op->name, ".json");
            FILE* fp = fopen(fn.c_str(), "w");
            if (fp) { print_metrics_json(&mt, fp); fclose(fp); printf("JSON exported to: %s\n", fn.c_str()); }
            else fprintf(stderr, "Could not open %s\n", fn.c_str());
        }
        if (!a.csv.empty()) {
            std::string fn = per_mode_name(a.csv, op->name, ".csv");
            FILE* fp = fopen(fn.c_str(), "w");
            if (fp) { print_metrics_csv(&mt, fp); fclose(fp); printf("CSV exported to: %s\n", fn.c_str()); }
            else fprintf(stderr, "Could not open %s\n", fn.c_str());
        }
        printf("SpMV completed successfully using mode: %s\n", op->name);
        printf("\n=== Output Checksum ===\n");
        printf("Sum(y):    %.16e\n", sum);
        printf("Norm2(y):  %.16e\n", mt.norm2_y);
        printf("=======================\n\n");
        op->free();
    }
    if (a.modes.size() > 1) printf("\n=== Multi-mode benchmark completed ===\n");
    free(mat.entries);
    if (d_entries) b200_free_device(d_entries);
    return EXIT_SUCCESS;
}
