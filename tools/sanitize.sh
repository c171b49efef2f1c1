#!/bin/bash
# compute-sanitizer over the CI-sized shapes: memcheck and racecheck of the golden random-effect
# suite on all four kernel paths (planner / general / global-X / cluster: the autouse fixture of
# tests/test_re_gpu_parity.py), memcheck of the fixed-effect kernels + device L-BFGS and of the partitioner kernels.
# Logs -> gpurun_out/sanitizer_*.log (summaries are copied to profiles/).
set -u
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
CS=/usr/local/cuda/bin/compute-sanitizer
run() {  # name, tool, timeout, pytest args...
    local name=$1 tool=$2 lim=$3; shift 3
    echo "== $name ($tool)" | tee "$OUT/sanitizer_$name.log"
    timeout "$lim" $CS --tool "$tool" --error-exitcode 7 --print-limit 20 \
        python -m pytest "$@" -x -q -p no:cacheprovider >> "$OUT/sanitizer_$name.log" 2>&1
    echo "exit code $?" | tee -a "$OUT/sanitizer_$name.log"
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" "$OUT/sanitizer_$name.log" | tail -4
}
run memcheck_re memcheck 900 "tests/test_re_gpu_parity.py" -k "test_golden_fit_matches_reference and 0] or test_golden_variances or test_deferred_entities"
run memcheck_fe memcheck 600 tests/test_fe_gpu.py
run memcheck_partition memcheck 600 tests/test_partition_gpu.py
run racecheck_re racecheck 900 "tests/test_re_gpu_parity.py" -k "test_golden_fit_matches_reference and 0]"
run racecheck_fe racecheck 600 tests/test_fe_gpu.py -k "golden or reference_restatement or planned_path_matches"
