#!/bin/bash
# e2e (gdmix_re_fit_host) against the chunk size of its pipeline: bash tools/e2e_chunk_sweep.sh 4096 8192 32768
for c in "$@"; do
  python bench.py --steps 3 --warmup 3 --no-sub --no-cpu-baseline --no-traffic-probe --parity-entities 0 --e2e-chunk "$c" 2>/dev/null \
    | python -c "import sys, json; d = json.loads(sys.stdin.readlines()[-1]); print('chunk', $c, round(d['e2e']['value']), round(d['e2e']['frac_of_h2d_bound'], 4))"
done
