#!/usr/bin/env bash
# run_reference_scripts.sh -- drive the REFERENCE's own orchestration scripts against this build
# (SURVEY.md section 8f-4).
#
# The reference scripts (scripts/run_all.sh, scripts/benchmarking/benchmark_weak_scaling.sh, ...) expect
# to sit in the reference tree: they `git checkout`, `make <target>` and then call ./bin/<tool> and
# `mpirun -np P ./bin/cg_solver_mgpu_stencil`.  This script builds a scratch "compat tree" that looks like
# that to them but resolves to THIS repo's binaries:
#
#   <tree>/scripts      -> the reference scripts (staged by oracle/Makefile into oracle/_ref/scripts when
#                          the reference tree is mounted; they are NOT part of this repo's history)
#   <tree>/Makefile     -> targets spmv_bench / generate_matrix / cg_solver / cg_solver_mgpu_stencil /
#                          clean that install this repo's CLIs into <tree>/bin
#   <tree>/shims        -> mpirun (cuda-spmv-benchmark_b200/scripts/mpirun: -np P -> --gpus=P, one process
#                          drives the P GPUs), mpic++ (presence check only), git (no-op checkout)
#
# usage: tools/run_reference_scripts.sh <out-dir> [run_all [--quick|--size=N] | weak [g:n ...]]
#   run_all : scripts/run_all.sh unchanged.
#   weak    : scripts/benchmarking/benchmark_weak_scaling.sh; its CONFIGURATION block ("EDIT THIS") is the
#             only thing touched: the default grids (5000..14142, a 24 GB .mtx at 8 GPUs) can be replaced by
#             the g:n pairs given on the command line.
set -euo pipefail
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
PKG="$ROOT/cuda-spmv-benchmark_b200"
SCRIPTS="${B200_REFERENCE_SCRIPTS:-$ROOT/oracle/_ref/scripts}"
OUT="$(mkdir -p "${1:?usage: run_reference_scripts.sh <out-dir> [run_all ...|weak ...]}" && cd "$1" && pwd)"
shift
WHAT="${1:-run_all}"
[ $# -gt 0 ] && shift
[ -f "$SCRIPTS/run_all.sh" ] || { echo "reference scripts not staged at $SCRIPTS (make -C oracle with the reference mounted)" >&2; exit 3; }

TREE="$(mktemp -d /tmp/b200_compat_XXXXXX)"
mkdir -p "$TREE/bin" "$TREE/shims" "$TREE/matrix" "$OUT"
cp -r "$SCRIPTS" "$TREE/scripts"
cat > "$TREE/Makefile" <<EOF
# compat Makefile: the reference's target names, this repo's binaries
PKG := $PKG
TOOLS := spmv_bench generate_matrix cg_solver cg_solver_mgpu_stencil
all: \$(TOOLS)
\$(TOOLS):
	@mkdir -p bin
	@cp \$(PKG)/bin/\$@ bin/\$@
	@echo "installed bin/\$@ (libspmv_b200)"
clean:
	@rm -rf bin
.PHONY: all clean \$(TOOLS)
EOF
# the installed CLIs find the library through their \$ORIGIN/.. rpath
cp "$PKG/libspmv_b200.so" "$TREE/libspmv_b200.so"
cp "$PKG/scripts/mpirun" "$TREE/shims/mpirun"
printf '#!/bin/sh\necho "mpic++ shim: one process drives all GPUs, nothing to compile" >&2\nexit 0\n' > "$TREE/shims/mpic++"
printf '#!/bin/sh\n[ "$1" = checkout ] && { echo "Already on %s"; exit 0; }\nexit 0\n' "'main'" > "$TREE/shims/git"
chmod +x "$TREE"/shims/*
export PATH="$TREE/shims:$PATH"
cd "$TREE"

case "$WHAT" in
  run_all)
    bash scripts/run_all.sh "$@" 2>&1 | tee "$OUT/run_all.log"
    cp -r results "$OUT/run_all_results" 2>/dev/null || true
    ;;
  weak)
    S=scripts/benchmarking/benchmark_weak_scaling.sh
    if [ $# -gt 0 ]; then  # replace the entries of WEAK_SCALING_CONFIGS=( ... ) -- the script's own edit point
      cfg=""; for c in "$@"; do cfg="$cfg    \"$c\"\n"; done
      python3 - "$S" "$cfg" <<'PY'
import re, sys
p, cfg = sys.argv[1], sys.argv[2].replace("\\n", "\n")
s = open(p).read()
s = re.sub(r"WEAK_SCALING_CONFIGS=\(\n.*?\n\)", "WEAK_SCALING_CONFIGS=(\n" + cfg + ")", s, count=1, flags=re.S)
open(p, "w").write(s)
PY
    fi
    bash "$S" 2>&1 | tee "$OUT/weak_scaling.log"
    for d in results_weak_scaling_*; do [ -d "$d" ] && cp -r "$d" "$OUT/"; done
    ;;
  *) echo "unknown mode $WHAT" >&2; exit 2 ;;
esac
echo "compat tree: $TREE   outputs: $OUT"
