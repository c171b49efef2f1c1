# GPU parity tests + a 200k-entity bench of the planner's path (and the general kernel with "both")
timeout 600 python -m pytest tests/test_re_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
for p in ${1:-auto}; do GDMIX_RE_PATH=$p timeout 300 python bench.py --entities 200000 --steps 3 --warmup 3 --no-cpu-baseline --e2e-entities 16384 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$p', round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), d['solve'])
    else: print(l, end='')
"; done
