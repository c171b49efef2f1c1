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
