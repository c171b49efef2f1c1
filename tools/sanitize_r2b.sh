#!/bin/bash
# compute-sanitizer over the kernels changed late in round 2 (fast kernel's m x m step on raw shared addresses, narrow
# row lengths / label bits, chunked scoring): memcheck of the random-effect suites that reach them, racecheck of the
# golden fits through the fast kernel.  Logs -> gpurun_out/sanitizer2_*.log
set -u
OUT=gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
run() { local name=$1 tool=$2 lim=$3; shift 3
  echo "== $name ($tool): pytest $*" > "$OUT/sanitizer2_$name.log"
  timeout "$lim" $CS --tool "$tool" --error-exitcode 7 --print-limit 10 python -m pytest "$@" -x -q -p no:cacheprovider >> "$OUT/sanitizer2_$name.log" 2>&1
  echo "exit code $?" >> "$OUT/sanitizer2_$name.log"
  echo "-- $name: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' "$OUT/sanitizer2_$name.log" | tail -1) | $(grep -E ' passed| failed' "$OUT/sanitizer2_$name.log" | tail -1) | $(tail -1 "$OUT/sanitizer2_$name.log")"
}
run memcheck_re memcheck 900 tests/test_re_gpu_parity.py -k "(test_golden_fit_matches_reference and 0]) or test_golden_variances or test_deferred_entities or test_narrow_row or test_scoring or test_logistic or test_golden_loss_grad"
run racecheck_re_fast racecheck 900 tests/test_re_gpu_parity.py -k "test_golden_fit_matches_reference and 0] and auto"
run memcheck_fe memcheck 900 tests/test_fe_gpu.py -k "not large_zipf"
run racecheck_fe racecheck 900 tests/test_fe_gpu.py -k "reference_restatement or tiled_objective_edge or rows_of_every_size or device_lbfgs_retraces"
run memcheck_partition memcheck 600 tests/test_partition_gpu.py
