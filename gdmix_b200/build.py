"""Builds gdmix_b200/lib/libgdmix_b200.so from csrc/ with nvcc for sm_100a (in-tree, no torch needed)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.environ.get("GDMIX_LIB_OUT") or os.path.join(HERE, "lib", "libgdmix_b200.so")  # GDMIX_LIB_OUT / GDMIX_NVCC_FLAGS: kernel experiments
SOURCES = ["api.cu"]
HEADERS = ["re_kernel.cuh", "re_fast.cuh", "re_small.cuh", "re_variance.cuh", "partition.cuh", "host_lbfgs.h", "fe_lbfgs.cuh", "fe_plan.cuh", "fe_tile.cuh", "seqex_parser.h", "seqex_writer.h", "avro_writer.h", "re_common.cuh", "re_passes.cuh", "re_lbfgs.cuh", "linesearch.cuh", "aux_kernels.cuh", os.path.join("..", "..", "include", "gdmix_b200.h")]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-Xcompiler", "-fPIC,-fvisibility=hidden,-fopenmp", "-shared", "-cudart", "static", "-lgomp",
           "-Xptxas", "-v" if verbose else "-warn-spills",
           "-o", LIB] + os.environ.get("GDMIX_NVCC_FLAGS", "").split() + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libgdmix_b200.so")
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
