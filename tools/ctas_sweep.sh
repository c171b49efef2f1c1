for k in 1 2 3; do GDMIX_FAST_CTAS=$k timeout 300 python bench.py --entities 100000 --steps 2 --warmup 2 --no-cpu-baseline --e2e-entities 1024 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('ctas $k', round(d['value']), round(d['ms_per_step'],2))
"; done
