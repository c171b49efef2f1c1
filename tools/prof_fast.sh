# usage: bash tools/prof_fast.sh <tag> [kernel-regex]
TAG=${1:-fast}
KRE=${2:-re_fast_kernel}
python - <<'PY'
import sys; sys.path.insert(0, '.')
import torch, ctypes as C
from gdmix_b200 import _capi as capi
from gdmix_b200.synthetic import make_batch
hb = make_batch(2000, 128, 256, 32, seed=1)
out = capi.re_fit_device(capi.DeviceBatch(hb), capi.make_opts())
torch.cuda.synchronize()
print("plan", capi.last_plan(), "deferred", int(out["workspace"][:12].view(torch.int32)[2].item()))
PY
ncu --set full --import-source on --clock-control none -k regex:$KRE -c 1 -o gpurun_out/prof_$TAG -f \
  python bench.py --entities 30000 --steps 1 --warmup 0 --no-cpu-baseline --e2e-entities 1024 > gpurun_out/prof_$TAG.log 2>&1
tail -3 gpurun_out/prof_$TAG.log
