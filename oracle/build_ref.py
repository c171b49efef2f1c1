#!/usr/bin/env python
"""Recipe for oracle/_ref: stages the reference's OWN modules for the CPU arm of bench.py -- TEST INFRASTRUCTURE.

The reference's per-entity solver class imports with nothing but numpy / scipy / sklearn:
  gdmix-trainer/src/gdmix/models/custom/binary_logistic_regression.py   BinaryLogisticRegressionTrainer
  gdmix-trainer/src/gdmix/util/model_utils.py                           threshold_coefficients
  gdmix-trainer/src/gdmix/util/constants.py
This script copies those three files, UNMODIFIED, from /root/reference into oracle/_ref/gdmix/... (plus empty
package markers).  oracle/_ref/ is git-ignored -- reference sources never enter the history -- but not
gpurun-ignored, so the staged tree travels to the GPU box, where /root/reference does not exist, and
`bench.py --impl reference` times the reference's own class there (cpu_baseline.kind "reference").  Where
oracle/_ref is absent the arm falls back to oracle/scipy_port.py (kind "port").

Run by __graft_entry__.build() when /root/reference is present; `python oracle/build_ref.py` by hand.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("GDMIX_REFERENCE_SRC", "/root/reference/gdmix-trainer/src")
OUT = os.path.join(HERE, "_ref")
FILES = ["gdmix/models/custom/binary_logistic_regression.py", "gdmix/util/model_utils.py", "gdmix/util/constants.py"]
PACKAGES = ["gdmix", "gdmix/models", "gdmix/models/custom", "gdmix/util"]


def build():
    """-> path of the staged tree, or None when the reference is not mounted here."""
    if not os.path.isdir(REF_SRC):
        return OUT if available() else None
    for p in PACKAGES:
        os.makedirs(os.path.join(OUT, p), exist_ok=True)
        init = os.path.join(OUT, p, "__init__.py")
        if not os.path.exists(init):
            open(init, "w").close()
    for f in FILES:
        shutil.copyfile(os.path.join(REF_SRC, f), os.path.join(OUT, f))
    return OUT


def available():
    return all(os.path.exists(os.path.join(OUT, f)) for f in FILES)


def load():
    """-> (BinaryLogisticRegressionTrainer, threshold_coefficients) of the staged reference, with the one-line shim
    for scipy >= 1.15 (fmin_l_bfgs_b lost its `disp` keyword; the reference passes disp=0) -- the same shim
    oracle/gen_golden.py uses."""
    if not available():
        raise ImportError("oracle/_ref is not staged (python oracle/build_ref.py where /root/reference is mounted)")
    if OUT not in sys.path:
        sys.path.insert(0, OUT)
    import scipy.optimize
    from gdmix.models.custom import binary_logistic_regression as blr
    from gdmix.util.model_utils import threshold_coefficients
    blr.fmin_l_bfgs_b = lambda *a, disp=None, **k: scipy.optimize.fmin_l_bfgs_b(*a, **k)
    return blr.BinaryLogisticRegressionTrainer, threshold_coefficients


if __name__ == "__main__":
    print(build())
