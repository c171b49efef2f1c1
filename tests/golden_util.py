"""Loading helpers for the committed golden vectors (tests/golden/, made by oracle/gen_golden.py)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# A case is "pinned" when the reference itself reproduces its coefficients under a 1e-17 nudge of the
# start point (see gen_golden.py: self_sensitivity).  The few unpinned ones have their optimum at
# infinity (single sample + unregularised intercept) and are compared on the objective instead.
PINNED_SENSITIVITY = 1e-10


def load_re():
    arr = np.load(os.path.join(GOLDEN, "re_golden.npz"))
    man = json.load(open(os.path.join(GOLDEN, "re_golden.json")))
    return arr, man["cases"]


def load_fe():
    arr = np.load(os.path.join(GOLDEN, "fe_golden.npz"))
    man = json.load(open(os.path.join(GOLDEN, "fe_golden.json")))
    return arr, man["cases"]


def load_partition():
    return json.load(open(os.path.join(GOLDEN, "partition_golden.json")))


def is_pinned(case):
    return case["self_sensitivity"] <= PINNED_SENSITIVITY and not case["self_nit_changed"]
