# Round evidence on one B200: all GPU tests, the default bench line, the reference arm, and the ncu launch list.
TAG=${1:-r1_b}
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 600 gpurun_out/bench_$TAG.err
cut -c1-400 gpurun_out/bench_$TAG.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>&1; cut -c1-300 gpurun_out/bench_ref_$TAG.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
   python bench.py --steps 2 --warmup 1 --entities 200000 --no-cpu-baseline --e2e-entities 16384 > gpurun_out/launches_$TAG.log 2>&1
tail -2 gpurun_out/launches_$TAG.csv | cut -c1-300
