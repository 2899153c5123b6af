#!/bin/bash
# Runs the REFERENCE's own CLI binaries (oracle/_ref/{generate_matrix,spmv_bench,cg_solver},
# compiled from the reference sources by oracle/Makefile) on the GPU box and collects their
# printed checksums / iteration counts into gpurun_out/ref_gpu.json -- the golden vectors that pin
# the oracle's numeric functions (tests/test_oracle_golden.py).  Test infrastructure only.
#   usage (on the GPU box, repo root):  bash oracle/run_ref_gpu.sh
#   REF_GPU_SIZES="10000:5" REF_GPU_OUT=ref_gpu_10k bash oracle/run_ref_gpu.sh
#       -> the reference's own kernels on THIS B200 at a BASELINE size (a 12 GB .mtx written by the reference's
#          generator, read back by its fscanf loader): the "A100 kernels recompiled for sm_100" bar of SURVEY 2a.
#          Note the reference cg_solver CLI keeps x between its warm-ups and the measured solves
#          (src/main/cg_solver.cu:155-173): its iteration count / median belong to warm restarts, so the
#          per-iteration time (median_ms / iterations) is the comparable figure.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/ref_gpu
mkdir -p "$OUT"
python - <<'PY'
import json, os, re, subprocess, sys, time
sys.path.insert(0, "oracle")
import orc
out = "gpurun_out/ref_gpu"
name = os.environ.get("REF_GPU_OUT", "ref_gpu")
sizes = [(int(t.split(":")[0]), float(t.split(":")[1])) for t in
         os.environ.get("REF_GPU_SIZES", "81:-4,64:5,512:5,1000:5,2000:5").split(",")]
cases = []
for n, center in sizes:
    mtx = os.path.join(out, "s%d.mtx" % n)
    if center == -4.0:
        orc.write_mtx_stencil5(n, mtx, "-4.0", "-1.0")   # byte-identical to matrix/example81x81.mtx
    else:
        subprocess.run(["oracle/_ref/generate_matrix", str(n), mtx], check=True, stdout=subprocess.DEVNULL)
    case = {"n": n, "center": center, "spmv": {}, "cg": {}, "mtx_bytes": os.path.getsize(mtx)}
    t_case = time.time()
    r = subprocess.run(["oracle/_ref/spmv_bench", mtx, "--mode=stencil5-csr,cusparse-csr",
                        "--json=%s/spmv_%d.json" % (out, n)], capture_output=True, text=True)
    open(os.path.join(out, "spmv_%d.log" % n), "w").write(r.stdout + r.stderr)
    for op in ("stencil5-csr", "cusparse-csr"):
        p = "%s/spmv_%d_%s.json" % (out, n, op)
        if os.path.exists(p):
            j = json.load(open(p))
            v = j["benchmark"]["validation"]
            case["spmv"][op] = {"sum_y": v["sum_y"], "norm2_y": v["norm2_y"],
                                "execution_time_ms": j["benchmark"]["performance"]["execution_time_ms"]}
    for op in ("stencil5-csr", "cusparse-csr"):
        r = subprocess.run(["oracle/_ref/cg_solver", mtx, "--mode=%s" % op, "--json=%s/cg_%d" % (out, n)],
                           capture_output=True, text=True)
        open(os.path.join(out, "cg_%d_%s.log" % (n, op)), "w").write(r.stdout + r.stderr)
        p = "%s/cg_%d_%s.json" % (out, n, op)
        if os.path.exists(p):
            j = json.load(open(p))
            case["cg"][op] = {"iterations": j["convergence"]["iterations"], "converged": j["convergence"]["converged"],
                              "residual_norm": j["convergence"]["residual_norm"],
                              "solution_sum": j["validation"]["solution_sum"],
                              "solution_norm": j["validation"]["solution_norm"], "median_ms": j["timing"]["median_ms"]}
    for op, c in case["cg"].items():
        c["ms_per_iteration"] = c["median_ms"] / max(c["iterations"], 1)
    if os.environ.get("REF_GPU_SAME_BOX_OURS") == "1":
        # this repo's CLIs on the SAME box and the SAME .mtx file, same protocol (5 warm-ups + 10 runs, median):
        # the only apples-to-apples comparison with the reference's kernels (GPUs of the pool differ by a few %)
        ours = {"spmv": {}, "cg": {}}
        bindir = "cuda-spmv-benchmark_b200/bin"
        r = subprocess.run([bindir + "/spmv_bench", mtx, "--mode=stencil5-csr,cusparse-csr", "--device-ingest",
                            "--json=%s/ours_spmv_%d.json" % (out, n)], capture_output=True, text=True)
        open(os.path.join(out, "ours_spmv_%d.log" % n), "w").write(r.stdout + r.stderr)
        for op in ("stencil5-csr", "cusparse-csr"):
            p = "%s/ours_spmv_%d_%s.json" % (out, n, op)
            if os.path.exists(p):
                j = json.load(open(p))
                ours["spmv"][op] = {"sum_y": j["benchmark"]["validation"]["sum_y"],
                                    "execution_time_ms": j["benchmark"]["performance"]["execution_time_ms"]}
        for op in ("stencil5-csr", "cusparse-csr"):
            r = subprocess.run([bindir + "/cg_solver", mtx, "--mode=%s" % op, "--device-ingest", "--json=%s/ours_cg_%d" % (out, n)],
                               capture_output=True, text=True)
            open(os.path.join(out, "ours_cg_%d_%s.log" % (n, op)), "w").write(r.stdout + r.stderr)
            p = "%s/ours_cg_%d_%s.json" % (out, n, op)
            if os.path.exists(p):
                j = json.load(open(p))
                ours["cg"][op] = {"iterations": j["convergence"]["iterations"], "median_ms": j["timing"]["median_ms"],
                                  "ms_per_iteration": j["timing"]["median_ms"] / max(j["convergence"]["iterations"], 1),
                                  "solution_sum": j["validation"]["solution_sum"]}
        case["this_repo_same_box"] = ours
    case["wall_s"] = round(time.time() - t_case, 1)
    cases.append(case)
    os.remove(mtx)
gpu = subprocess.run(["nvidia-smi", "--query-gpu=name,driver_version", "--format=csv,noheader"],
                     capture_output=True, text=True).stdout.strip()
json.dump({"generator": "oracle/run_ref_gpu.sh: reference spmv_bench / cg_solver (oracle/_ref, -arch=sm_100) on " + gpu,
           "cases": cases}, open("gpurun_out/%s.json" % name, "w"), indent=1)
print(json.dumps(cases)[:2000])
PY
