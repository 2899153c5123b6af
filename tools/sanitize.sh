#!/usr/bin/env bash
# sanitize.sh -- compute-sanitizer over the small-grid GPU parity tests (memcheck +
# racecheck on the known-answer configurations; plus synccheck and initcheck).
#
#   bash tools/sanitize.sh [out-dir]          (on a GPU box: gpurun -- 'bash tools/sanitize.sh profiles/sanitizer_r02')
#
# The kernels under test rest on hand-argued mbarrier / async-proxy ordering (csrc/stencil5.cuh,
# csrc/csr_ell.cuh), ticketed last-CTA reductions and cross-rank flag protocols (csrc/cg_kernels.cuh); the
# selected tests drive every one of them on grids of at most 130 x 130 (bundled 81 x 81 matrix, 3 x 3,
# virtual ranks 2..8, both CG schedules, generic CSR / ELLPACK lane-per-row, sub-warp and warp-per-row rows,
# device ingest).  One summary file per tool: the sanitizer's own "ERROR SUMMARY" line plus pytest's verdict.
set -uo pipefail
cd "$(dirname "$0")/.."
OUT="${1:-gpurun_out/sanitizer}"
mkdir -p "$OUT"
SEL='(test_cg_solve_device_matches_oracle and (3- or 81-)) or test_cg_from_mtx_file_bundled or (virtual_ranks and (81-2 or 64-8)) or (bit_identical_to_classic and (81-1000 or 64-8 or 130-3)) or (test_pcg_jacobi_matches_oracle and 40) or (test_pcg_mgpu and 40-2) or test_halo_mgpu_operator or (test_coo_to_csr_bit_exact and (stencil or long_rows)) or (test_stencil5_csr_bit_exact and (0-81 or 0-130 or 3-81 or 9-81 or 20-81 or 21-130)) or test_stencil5_nonstandard_values_bundled or (test_stencil5_halo_bands_bit_exact and (81-2 or 64-8)) or (test_generic_csr_and_ellpack and (0] or 6]) and (unbalanced or long_rows or uniform5)) or test_generic_spmv_with_fused_dot'
FILES="tests/test_gpu_cg.py tests/test_gpu_spmv.py tests/test_gpu_ingest.py"
rc_all=0
for tool in ${SANITIZE_TOOLS:-memcheck racecheck synccheck}; do
    log="$OUT/${tool}.log"
    extra=""
    [ "$tool" = memcheck ] && extra="--leak-check no"
    [ "$tool" = initcheck ] && extra="--track-unused-memory no"
    timeout ${SANITIZE_TIMEOUT:-900} compute-sanitizer --tool "$tool" $extra --target-processes all --error-exitcode 99 \
        python -m pytest $FILES -q -m gpu -x -k "$SEL" -p no:cacheprovider > "$log" 2>&1
    rc=$?
    {
        echo "tool: $tool    exit code: $rc    ($(date -u +%FT%TZ), $(nvidia-smi --query-gpu=name,driver_version --format=csv,noheader | head -1))"
        echo "selection: -k \"$SEL\""
        grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" "$log" | tail -8
        echo "--- first reports (if any) ---"
        grep -E "^=========" "$log" | grep -vE "COMPUTE-SANITIZER|ERROR SUMMARY|RACECHECK SUMMARY" | head -30
    } > "$OUT/${tool}_summary.txt"
    cat "$OUT/${tool}_summary.txt"
    [ $rc -ne 0 ] && rc_all=1
done
exit $rc_all
