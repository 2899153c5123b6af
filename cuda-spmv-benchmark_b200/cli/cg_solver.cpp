// cg_solver <file.mtx | --grid=n> [--mode=a[,b]] [--host] [--tol=] [--maxiter=] [--timers] [--json=] [--csv=]
//           [--precond=jacobi]   (extension: Jacobi-preconditioned CG, pcg_solve_device)
// reference src/main/cg_solver.cu:23-243: defaults stencil5-csr / device path / tol 1e-6 / 1000 it,
// b = 1, x0 = 0, 3 warm-up solves, cg_benchmark_with_stats_device(10), "<json>_<mode>.json", CSV appended.
#include "cli_common.h"

int main(int argc, char** argv) {
    CliArgs a = parse_cli(argc, argv);
    if (a.matrix.empty() && a.grid <= 0) {
        fprintf(stderr, "Usage: %s <matrix.mtx | --grid=n> [--mode=<m1[,m2]>] [--host] [--tol=<t>] [--maxiter=<n>] "
                        "[--timers] [--json=<file>] [--csv=<file>] [--precond=jacobi]\n", argv[0]);
        return EXIT_FAILURE;
    }
    if (a.modes.empty()) a.modes.push_back("stencil5-csr");
    for (auto& m : a.modes)
        if (!get_operator(m.c_str())) { fprintf(stderr, "Unknown mode '%s'\n", m.c_str()); return EXIT_FAILURE; }
    MatrixData mat;
    void* d_entries = nullptr;
    if (load_or_generate(a, &mat, &d_entries)) return EXIT_FAILURE;
    printf("Matrix: %d x %d, %d nonzeros, grid %d\n", mat.rows, mat.cols, mat.nnz, mat.grid_size);
    std::vector<double> b((size_t)mat.rows, 1.0), x((size_t)mat.rows, 0.0);
    CGConfig cfg = {a.maxiter, a.tol, 1, a.timers ? 1 : 0};
    bool first = true;
    for (auto& m : a.modes) {
        SpmvOperator* op = get_operator(m.c_str());
        printf("\n=== CG with operator: %s (%s interface) ===\n", op->name, a.host ? "host" : "device");
        if (!a.host && !op->run_device) {
            fprintf(stderr, "[ERROR] Operator '%s' does not support device-native interface\n", op->name);
            return EXIT_FAILURE;
        }
        if (init_operator(op, &mat, d_entries) != 0) { fprintf(stderr, "Failed to initialize operator '%s'\n", op->name); return EXIT_FAILURE; }
        CGStats st;
        CGConfig quiet = cfg;
        quiet.verbose = 0;
        printf("Warmup (3 runs)...\n");
        for (int w = 0; w < 3; w++) {
            std::fill(x.begin(), x.end(), 0.0);
            int rc = a.jacobi ? pcg_solve_device(op, &mat, b.data(), x.data(), quiet, &st)
                     : a.host ? cg_solve(op, &mat, b.data(), x.data(), quiet, &st)
                              : cg_solve_device(op, &mat, b.data(), x.data(), quiet, &st);
            if (rc != 0) { fprintf(stderr, "CG solve failed\n"); return EXIT_FAILURE; }
        }
        std::fill(x.begin(), x.end(), 0.0);
        printf("Running benchmark (%d runs)...\n", a.runs);
        BenchmarkStats bs;
        if (a.jacobi) {
            // no reference wrapper exists for the preconditioned solver: same protocol (runs solves from
            // x0 = 0, median of the solver's own device time), statistics without the outlier filter
            std::vector<double> t;
            for (int r = 0; r < a.runs; r++) {
                std::fill(x.begin(), x.end(), 0.0);
                if (pcg_solve_device(op, &mat, b.data(), x.data(), quiet, &st) != 0) { fprintf(stderr, "PCG solve failed\n"); return EXIT_FAILURE; }
                t.push_back(st.time_total_ms);
            }
            std::sort(t.begin(), t.end());
            memset(&bs, 0, sizeof bs);
            double sum = 0, sq = 0;
            for (double v : t) sum += v;
            bs.valid_runs = (int)t.size(); bs.min_ms = t.front(); bs.max_ms = t.back(); bs.median_ms = t[t.size() / 2];
            bs.mean_ms = sum / t.size();
            for (double v : t) sq += (v - bs.mean_ms) * (v - bs.mean_ms);
            bs.std_dev_ms = t.size() > 1 ? sqrt(sq / (t.size() - 1)) : 0.0;
            printf("(Jacobi-preconditioned CG)\n");
        } else if (cg_benchmark_with_stats_device(op, &mat, b.data(), x.data(), cfg, a.runs, &bs, &st) != 0) {
            fprintf(stderr, "CG benchmark failed for mode '%s'\n", op->name);
            return EXIT_FAILURE;
        }
        printf("\n=== CG Results (%s) ===\n", op->name);
        printf("Converged: %s in %d iterations, residual %.6e\n", st.converged ? "YES" : "NO", st.iterations, st.residual_norm);
        printf("Time: median %.3f ms (mean %.3f, min %.3f, max %.3f, std %.3f; %d runs, %d outliers)\n", bs.median_ms,
               bs.mean_ms, bs.min_ms, bs.max_ms, bs.std_dev_ms, bs.valid_runs, bs.outliers_removed);
        printf("Sum(x):    %.16e\n", st.solution_sum);
        printf("Norm2(x):  %.16e\n", st.solution_norm);
        if (!a.json.empty()) {
            std::string fn = a.json + "_" + m + ".json";
            export_cg_json(fn.c_str(), m.c_str(), &mat, &bs, &st);
        }
        if (!a.csv.empty()) export_cg_csv(a.csv.c_str(), m.c_str(), &mat, &bs, &st, first);
        first = false;
        op->free();
    }
    free(mat.entries);
    if (d_entries) b200_free_device(d_entries);
    return EXIT_SUCCESS;
}
