#!/bin/bash
# compute-sanitizer over the CI-sized shapes: memcheck and racecheck of the golden random-effect
# suite on all five kernel paths (planner / general / global-X / cluster / warp-per-entity: the autouse fixture of
# tests/test_re_gpu_parity.py), memcheck + racecheck of the fixed-effect tiled objective and the device L-BFGS, memcheck
# of the partitioner / AUC kernels.  Logs -> $OUT/sanitizer_*.log (summaries are copied to profiles/).
set -u
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
CS=/usr/local/cuda/bin/compute-sanitizer
run() {  # name, tool, timeout, pytest args...
    local name=$1 tool=$2 lim=$3; shift 3
    echo "== $name ($tool): pytest $*" > "$OUT/sanitizer_$name.log"
    timeout "$lim" $CS --tool "$tool" --error-exitcode 7 --print-limit 10 \
        python -m pytest "$@" -x -q -p no:cacheprovider >> "$OUT/sanitizer_$name.log" 2>&1
    echo "exit code $?" >> "$OUT/sanitizer_$name.log"
    echo "-- $name: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' "$OUT/sanitizer_$name.log" | tail -1) | $(grep -E ' passed| failed' "$OUT/sanitizer_$name.log" | tail -1) | $(tail -1 "$OUT/sanitizer_$name.log")"
}
run memcheck_re memcheck 900 tests/test_re_gpu_parity.py -k "(test_golden_fit_matches_reference and 0]) or test_golden_variances or test_deferred_entities or test_warp_per_entity"
run memcheck_fe memcheck 900 tests/test_fe_gpu.py -k "not large_zipf"
run memcheck_partition memcheck 600 tests/test_partition_gpu.py
run racecheck_re racecheck 1200 tests/test_re_gpu_parity.py -k "test_golden_fit_matches_reference and 0]"
run racecheck_fe racecheck 900 tests/test_fe_gpu.py -k "reference_restatement or tiled_objective_edge or rows_of_every_size or device_lbfgs_retraces"
