# usage: bash tools/variant_bench.sh [entities] -- benches every gdmix_b200/lib/exp/*.so and the in-tree library on C1
E=${1:-400000}
for so in gdmix_b200/lib/exp/*.so; do
  GDMIX_B200_LIB=$PWD/$so timeout 300 python bench.py --entities $E --steps 3 --warmup 2 --no-cpu-baseline --e2e-entities 8192 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$so', round(d['value']), round(d['ms_per_step'],2), d['roofline']['kernel'], d['solve'])
    else: print(l, end='')
"; done
