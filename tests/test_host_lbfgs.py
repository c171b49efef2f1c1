"""gdmix_lbfgs_* (the replicated host-side solver of the fixed-effect path, product code) driven by the
oracle's objective: must retrace the reference's scipy trajectories recorded in tests/golden/."""
import numpy as np
import pytest

from gdmix_b200 import _capi as capi
from oracle import oracle as O
from tests.golden_util import is_pinned, load_fe, load_re

FE_ARR, FE_CASES = load_fe()
RE_ARR, RE_CASES = load_re()


def _solve(n, opts, fun, x0):
    x = np.array(x0, dtype=np.float64, copy=True)
    s = capi.HostLbfgs(n, opts)
    f, g = fun(x)
    while s.iterate(x, f, g) == capi.HostLbfgs.NEED_FG:
        f, g = fun(x)
    info = s.info()
    s.close()
    return x, info


@pytest.mark.parametrize("c", FE_CASES, ids=[c["name"] for c in FE_CASES])
def test_fixed_effect_golden(c):
    k = c["key"]
    rows = O.FeBlock(c["n"], c["D"], FE_ARR[k + "_rowptr"], FE_ARR[k + "_col"], FE_ARR[k + "_val"], FE_ARR[k + "_y"],
                     FE_ARR[k + "_w"], FE_ARR[k + "_off"], linear_regression=c["linear_regression"])
    oo = O.make_opts(l2=c["l2"], regularize_bias=True, has_intercept=c["has_intercept"], m=c["m"],
                     max_iter=c["max_iter"], factr=c["factr"])
    po = capi.make_opts(l2=c["l2"], regularize_bias=True, has_intercept=c["has_intercept"], m=c["m"],
                        max_iter=c["max_iter"], factr=c["factr"])
    x, info = _solve(c["D"] + (1 if c["has_intercept"] else 0), po, lambda x: O.fe_loss_grad(rows, oo, x),
                     FE_ARR[k + "_x0"])
    assert (info["nit"], info["nfev"], info["status"]) == (c["nit"], c["nfev"], c["warnflag"])
    np.testing.assert_allclose(x, FE_ARR[k + "_theta"], rtol=1e-8, atol=1e-10)


@pytest.mark.parametrize("c", RE_CASES[::4], ids=[c["name"] for c in RE_CASES[::4]])
def test_random_effect_golden_through_host_solver(c):
    k = c["key"]
    blk = O.EntityBlock(c["n"], c["d"], RE_ARR[k + "_rowptr"], RE_ARR[k + "_col"], RE_ARR[k + "_val"],
                        RE_ARR[k + "_y"], RE_ARR[k + "_w"], RE_ARR[k + "_off"])
    kw = dict(l2=c["l2"], regularize_bias=c["regularize_bias"], has_intercept=c["has_intercept"], m=c["m"],
              max_iter=c["max_iter"], tol=c["tol"])
    oo, po = O.make_opts(**kw), capi.make_opts(**kw)
    p = c["d"] + (1 if c["has_intercept"] else 0)
    x0 = RE_ARR[k + "_theta0"] if c["warm"] else np.zeros(p)
    x, info = _solve(p, po, lambda x: O.re_loss_grad(blk, oo, x), x0)
    if is_pinned(c):
        assert (info["nit"], info["nfev"], info["status"]) == (c["nit"], c["nfev"], c["warnflag"])
        ref = RE_ARR[k + "_theta"]
        assert np.linalg.norm(x - ref) <= 1e-9 * max(np.linalg.norm(ref), 1e-300)


def test_long_vectors_give_the_same_bits_for_any_number_of_host_threads():
    """Above 32 768 coefficients the solver sweeps its vectors with several host threads; inner products are block
    sums added in block order, so the trajectory must not depend on the thread count (ranks with different core
    counts keep bit-identical state).  Each thread count runs in its own process (the count is read once)."""
    import os
    import subprocess
    import sys
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r)
from gdmix_b200 import _capi as capi
n = 70001
rng = np.random.default_rng(0)
A = np.abs(rng.standard_normal(n)) + 0.5
b = rng.standard_normal(n)
h = capi.HostLbfgs(n, capi.make_opts(l2=0.0, max_iter=25))
x = np.zeros(n)
fg = lambda x: (0.5 * float((A * x * x).sum() - 2 * (b * x).sum()), A * x - b)
f, g = fg(x)
while h.iterate(x, f, g) == capi.HostLbfgs.NEED_FG:
    f, g = fg(x)
i = h.info()
print(i["nit"], i["nfev"], x.tobytes().hex()[:64], repr(float(x.sum())))
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for th in ("1", "3", "8"):
        env = dict(os.environ, GDMIX_HOST_THREADS=th)
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stderr[-500:]
        outs.append(r.stdout.strip())
    assert outs[0] == outs[1] == outs[2], outs
    assert int(outs[0].split()[0]) > 5
