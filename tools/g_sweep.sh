for g in 128 256; do timeout 300 python bench.py --entities 100000 --steps 2 --warmup 2 --no-cpu-baseline --e2e-entities 1024 --threads-per-entity $g 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('G $g', round(d['value']), round(d['ms_per_step'],2))
    else: print(l[:300])
"; done
