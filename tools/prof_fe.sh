ncu --set full --import-source on --clock-control none -k regex:fe_ -c 12 -o gpurun_out/prof_fe_$1 -f python tools/fe_bench.py 4000000 100000 32 1 > gpurun_out/prof_fe_$1.log 2>&1
tail -3 gpurun_out/prof_fe_$1.log | cut -c1-300
