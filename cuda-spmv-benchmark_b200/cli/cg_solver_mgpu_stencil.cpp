// cg_solver_mgpu_stencil <file.mtx | --grid=n> [--gpus=P] [--timers] [--json=F] [--csv=F] [--precond=jacobi]
// reference src/main/cg_solver_mgpu_stencil.cu:22-197 (there: mpirun -np P, one rank per GPU).
// Here one process drives P GPUs (default: all visible); max_iters 1000, tol 1e-6, 3 warm-ups,
// cg_benchmark_with_stats_mgpu_partitioned(10), Sum(x)/Norm2(x) lines, export_cg_mgpu_json.
#include <cuda_profiler_api.h>
#include <cuda_runtime_api.h>

#include "cli_common.h"

int main(int argc, char** argv) {
    CliArgs a = parse_cli(argc, argv);
    if (a.matrix.empty() && a.grid <= 0) {
        fprintf(stderr, "Usage: %s <matrix.mtx | --grid=n> [--gpus=P] [--timers] [--json=<file>] [--csv=<file>]\n", argv[0]);
        return EXIT_FAILURE;
    }
    MatrixData mat;
    if (load_or_generate(a, &mat)) return EXIT_FAILURE;
    if (mat.grid_size <= 0) {
        fprintf(stderr, "This solver needs a stencil matrix (STENCIL_GRID_SIZE comment in the .mtx)\n");
        return EXIT_FAILURE;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { fprintf(stderr, "No CUDA device\n"); return EXIT_FAILURE; }
    int P = a.gpus > 0 ? a.gpus : ndev;
    std::vector<int> devs(P);
    for (int r = 0; r < P; r++) devs[r] = r % ndev;  // more ranks than GPUs = virtual ranks (testing)
    if (b200_mgpu_init_single_process(P, devs.data(), mat.grid_size) != 0) { fprintf(stderr, "multi-GPU init failed\n"); return EXIT_FAILURE; }
    printf("Multi-GPU CG: %d rank(s) on %d GPU(s), matrix %d x %d, %d nonzeros, grid %d\n", P, ndev < P ? ndev : P,
           mat.rows, mat.cols, mat.nnz, mat.grid_size);
    std::vector<double> b((size_t)mat.rows, 1.0), x((size_t)mat.rows, 0.0);
    CGConfigMultiGPU cfg = {a.maxiter, a.tol, 1, a.timers ? 1 : 0};
    CGConfigMultiGPU quiet = cfg;
    quiet.verbose = 0;
    CGStatsMultiGPU st;
    auto solve = a.jacobi ? pcg_solve_mgpu_partitioned : cg_solve_mgpu_partitioned;
    for (int w = 0; w < 3; w++) {
        std::fill(x.begin(), x.end(), 0.0);
        if (solve(nullptr, &mat, b.data(), x.data(), quiet, &st) != 0) { fprintf(stderr, "CG solve failed\n"); return EXIT_FAILURE; }
    }
    // one solve inside the profiler window (reference src/main/cg_solver_mgpu_stencil.cu:115-117)
    std::fill(x.begin(), x.end(), 0.0);
    cudaProfilerStart();
    if (solve(nullptr, &mat, b.data(), x.data(), quiet, &st) != 0) { fprintf(stderr, "CG solve failed\n"); return EXIT_FAILURE; }
    cudaProfilerStop();
    std::fill(x.begin(), x.end(), 0.0);
    BenchmarkStats bs;
    if (a.jacobi) {  // extension: the reference's bench wrapper only knows plain CG
        std::vector<double> t;
        for (int r = 0; r < a.runs; r++) {
            std::fill(x.begin(), x.end(), 0.0);
            if (solve(nullptr, &mat, b.data(), x.data(), r == 0 ? cfg : quiet, &st) != 0) { fprintf(stderr, "PCG solve failed\n"); return EXIT_FAILURE; }
            t.push_back(st.time_total_ms);
        }
        std::sort(t.begin(), t.end());
        bs.median_ms = t[t.size() / 2]; bs.mean_ms = 0; for (double v : t) bs.mean_ms += v / t.size();
        bs.min_ms = t.front(); bs.max_ms = t.back(); bs.std_dev_ms = 0; bs.valid_runs = (int)t.size(); bs.outliers_removed = 0;
    } else if (cg_benchmark_with_stats_mgpu_partitioned(nullptr, &mat, b.data(), x.data(), cfg, a.runs, &bs, &st) != 0) {
        fprintf(stderr, "CG benchmark failed\n");
        return EXIT_FAILURE;
    }
    printf("\n=== Multi-GPU CG Results ===\n");
    printf("Converged: %s in %d iterations\n", st.converged ? "YES" : "NO", st.iterations);
    printf("Residual norm: %.6e\n", st.residual_norm);
    printf("Time: median %.3f ms (mean %.3f, min %.3f, max %.3f, std %.3f; %d runs, %d outliers)\n", bs.median_ms,
           bs.mean_ms, bs.min_ms, bs.max_ms, bs.std_dev_ms, bs.valid_runs, bs.outliers_removed);
    printf("\n=== Output Checksum ===\n");
    printf("Sum(x):    %.16e\n", st.solution_sum);
    printf("Norm2(x):  %.16e\n", st.solution_norm);
    printf("=======================\n");
    if (!a.json.empty()) export_cg_mgpu_json(a.json.c_str(), a.jacobi ? "partitioned-halo-jacobi" : "partitioned-halo", &mat, &bs, &st, P);
    if (!a.csv.empty()) printf("CSV export is not implemented for the multi-GPU solver (as in the reference)\n");
    free(mat.entries);
    return EXIT_SUCCESS;
}
