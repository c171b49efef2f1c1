# usage: bash tools/prof_shape.sh <tag> E n d k  -- ncu --set full of the fast kernel on a synthetic shape
TAG=$1; shift
ncu --set full --import-source on --clock-control none -k regex:re_fast_kernel -c 1 -o gpurun_out/prof_$TAG -f \
  python tools/shape_bench.py "$@" > gpurun_out/prof_$TAG.log 2>&1
tail -2 gpurun_out/prof_$TAG.log | cut -c1-200
