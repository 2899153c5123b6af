// bench_stats.cpp -- repeated-run statistics behind benchmark_with_stats,
// cg_benchmark_with_stats_device and cg_benchmark_with_stats_mgpu_partitioned
// (reference src/spmv/benchmark_stats.cu:39-177, src/spmv/benchmark_stats_mgpu_partitioned.cu:42-128).
//
// Rule kept from the reference: collect the successful runs (at least 3, else -1), drop every run
// further than two population standard deviations from the mean, and report mean / sigma /
// median / min / max of the survivors.  The "final" CG statistics are those of the run at the
// middle POSITION of the surviving list in launch order (the reference's filtered_indices[count/2]
// quirk) -- not of the run that has the median time.
#include <math.h>

#include <algorithm>
#include <vector>

#include "host_common.h"

namespace {

struct Summary {
    BenchmarkStats st;
    int middle_run;  // index into the original run list
};

bool summarise(const std::vector<double>& t, Summary* out) {
    const int n = (int)t.size();
    if (n < 3) return false;
    double mean = 0.0;
    for (double v : t) mean += v;
    mean /= n;
    double ss = 0.0;
    for (double v : t) ss += (v - mean) * (v - mean);
    const double sd = sqrt(ss / n);
    std::vector<double> keep;
    std::vector<int> keep_idx;
    for (int i = 0; i < n; i++)
        if (fabs(t[i] - mean) <= 2.0 * sd) { keep.push_back(t[i]); keep_idx.push_back(i); }
    const int m = (int)keep.size();
    double mean2 = 0.0;
    for (double v : keep) mean2 += v;
    mean2 /= m;
    double ss2 = 0.0;
    for (double v : keep) ss2 += (v - mean2) * (v - mean2);
    out->middle_run = keep_idx[m / 2];
    std::sort(keep.begin(), keep.end());
    out->st.mean_ms = mean2;
    out->st.std_dev_ms = sqrt(ss2 / m);
    out->st.median_ms = (m % 2 == 0) ? (keep[m / 2 - 1] + keep[m / 2]) / 2.0 : keep[m / 2];
    out->st.min_ms = keep.front();
    out->st.max_ms = keep.back();
    out->st.valid_runs = m;
    out->st.outliers_removed = n - m;
    return true;
}

template <class Stats, class Solve>
int bench_cg(MatrixData* mat, double* x, int num_runs, BenchmarkStats* bench, Stats* final_stats, Solve solve) {
    if (!mat || !x || !bench || !final_stats || num_runs < 1) return -1;
    std::vector<double> x0(x, x + mat->rows);  // every run starts from the caller's initial guess
    std::vector<double> times;
    std::vector<Stats> all;
    for (int i = 0; i < num_runs; i++) {
        std::copy(x0.begin(), x0.end(), x);
        Stats s;
        if (solve(&s) == 0) { times.push_back(s.time_total_ms); all.push_back(s); }
    }
    Summary sm;
    if (!summarise(times, &sm)) return -1;
    *bench = sm.st;
    *final_stats = all[sm.middle_run];
    return 0;
}

}  // namespace

extern "C" int benchmark_with_stats(int (*run_func)(const double*, double*, double*), const double* x, double* y,
                                    int num_runs, BenchmarkStats* stats) {
    if (!run_func || !stats || num_runs < 1) return -1;
    std::vector<double> times;
    for (int i = 0; i < num_runs; i++) {
        double ms = 0.0;
        if (run_func(x, y, &ms) == 0) times.push_back(ms);
    }
    Summary sm;
    if (!summarise(times, &sm)) return -1;
    *stats = sm.st;
    return 0;
}

extern "C" int cg_benchmark_with_stats_device(SpmvOperator* spmv_op, MatrixData* mat, double* b, double* x,
                                              CGConfig config, int num_runs, BenchmarkStats* bench_stats,
                                              CGStats* final_stats) {
    config.verbose = 0;  // silent runs, as in the reference
    return bench_cg<CGStats>(mat, x, num_runs, bench_stats, final_stats,
                             [&](CGStats* s) { return cg_solve_device(spmv_op, mat, b, x, config, s); });
}

extern "C" int cg_benchmark_with_stats_mgpu_partitioned(SpmvOperator* spmv_op, MatrixData* mat, double* b, double* x,
                                                        CGConfigMultiGPU config, int num_runs,
                                                        BenchmarkStats* bench_stats, CGStatsMultiGPU* final_stats) {
    config.verbose = 0;
    return bench_cg<CGStatsMultiGPU>(mat, x, num_runs, bench_stats, final_stats, [&](CGStatsMultiGPU* s) {
        return cg_solve_mgpu_partitioned(spmv_op, mat, b, x, config, s);
    });
}
