"""CPU oracle (oracle/lr_oracle.c) against the golden vectors produced by the reference itself
(BinaryLogisticRegressionTrainer + scipy.optimize.fmin_l_bfgs_b, see oracle/gen_golden.py).
This is the pin that lets the GPU parity tests trust the oracle."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests.golden_util import is_pinned, load_fe, load_partition, load_re

ARR, CASES = load_re()


def _block(c):
    k = c["key"]
    return O.EntityBlock(c["n"], c["d"], ARR[k + "_rowptr"], ARR[k + "_col"], ARR[k + "_val"], ARR[k + "_y"],
                         ARR[k + "_w"], ARR[k + "_off"])


def _opts(c):
    return O.make_opts(l2=c["l2"], regularize_bias=c["regularize_bias"], has_intercept=c["has_intercept"],
                       m=c["m"], max_iter=c["max_iter"], tol=c["tol"])


@pytest.mark.parametrize("c", CASES, ids=[c["name"] for c in CASES])
def test_loss_gradient_match_reference(c):
    """a1/a2: _loss and _gradient at an arbitrary theta (binary_logistic_regression.py:84-131)."""
    k = c["key"]
    f, g = O.re_loss_grad(_block(c), _opts(c), ARR[k + "_probe"])
    assert abs(f - c["probe_f"]) <= 1e-14 * max(1.0, abs(c["probe_f"]))
    np.testing.assert_allclose(g, ARR[k + "_probe_g"], rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("c", CASES, ids=[c["name"] for c in CASES])
def test_fit_matches_reference(c):
    """a3: trajectory-exact L-BFGS-B restatement: same nit / nfev / warnflag, theta to 1e-9."""
    k = c["key"]
    th0 = ARR[k + "_theta0"] if c["warm"] else None
    theta, f, nit, nfev, status, g = O.re_fit(_block(c), _opts(c), th0)
    ref = ARR[k + "_theta"]
    if is_pinned(c):
        assert (nit, nfev, status) == (c["nit"], c["nfev"], c["warnflag"])
        rel = np.linalg.norm(theta - ref) / max(np.linalg.norm(ref), 1e-300)
        assert rel <= 1e-9, rel
        assert abs(f - c["f"]) <= 1e-12 * max(1.0, abs(c["f"]))
        np.testing.assert_allclose(O.threshold(theta), ARR[k + "_theta_thresholded"], rtol=1e-8, atol=1e-12)
    else:
        # the reference does not reproduce itself here (optimum at infinity); both must still stop on
        # the same rule with a vanishing objective
        assert status == c["warnflag"] == 0
        assert np.abs(g).max() <= 1e-5 or f <= c["f"] + 1e-5
        assert abs(f - c["f"]) <= 1e-5


@pytest.mark.parametrize("c", [c for c in CASES if c["has_var_simple"]],
                         ids=[c["name"] for c in CASES if c["has_var_simple"]])
def test_variance_matches_reference(c):
    """a4: _compute_variance SIMPLE / FULL (binary_logistic_regression.py:144-189)."""
    k = c["key"]
    blk, o = _block(c), _opts(c)
    v = O.re_variance(blk, o, ARR[k + "_theta"], "simple")
    np.testing.assert_allclose(v, ARR[k + "_var_simple"], rtol=1e-10)
    if c["has_var_full"]:
        ref = ARR[k + "_var_full"]
        if np.all(np.isfinite(ref)) and np.abs(ref).max() < 1e8:
            v = O.re_variance(blk, o, ARR[k + "_theta"], "full")
            np.testing.assert_allclose(v, ref, rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("c", CASES[::5], ids=[c["name"] for c in CASES[::5]])
def test_scores_match_reference(c):
    """a5/a9: logits = X1.theta + offset, per-coordinate = logits - offset."""
    k = c["key"]
    logit, per = O.re_score(_block(c), _opts(c), ARR[k + "_theta"])
    np.testing.assert_allclose(logit, ARR[k + "_logits"], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(per, ARR[k + "_logits"] - ARR[k + "_off"].astype(np.float64), rtol=1e-12, atol=1e-13)
    logit0, per0 = O.re_score(_block(c), _opts(c), None)
    np.testing.assert_array_equal(logit0, ARR[k + "_off"].astype(np.float64))
    np.testing.assert_array_equal(per0, np.zeros(c["n"]))


def test_appendix_d_known_answers():
    """The two entities of the reference fixture grouped_per_member_train/data.tfrecord with the RE
    test's parameters (l2=0.1, regularize_bias=True): spot values recorded in SURVEY.md App. D."""
    by_name = {c["name"]: c for c in CASES}
    c = by_name["fixture:100034/l2=0.1/maxiter=100"]
    assert c["unique_global_indices"] == [0, 7, 34, 57, 60, 80, 95]
    theta, f, nit, nfev, status, _ = O.re_fit(_block(c), _opts(c))
    assert (nit, nfev) == (14, 15)
    np.testing.assert_allclose(theta[:3], [0.4270237951733761, -0.07106015236644797, -0.14212030473289594],
                               rtol=1e-9)
    assert abs(f - 0.07477851128219332) < 1e-14
    c = by_name["fixture:100/l2=0.1/maxiter=100"]
    theta, f, nit, nfev, status, _ = O.re_fit(_block(c), _opts(c))
    assert (nit, nfev) == (6, 7)
    np.testing.assert_allclose(theta, [0.19931540705268397, -0.697603924684394, 0.45842542671707404], rtol=1e-9)


def test_threshold_rule():
    """model_utils.py:4-12: abs(x) <= 1e-4 -> 0.0 (boundary inclusive)."""
    x = np.array([1e-4, -1e-4, 1.0000001e-4, -2e-4, 0.0, 5e-5])
    np.testing.assert_array_equal(O.threshold(x), [0.0, 0.0, 1.0000001e-4, -2e-4, 0.0, 0.0])


FE_ARR, FE_CASES = load_fe()


@pytest.mark.parametrize("c", FE_CASES, ids=[c["name"] for c in FE_CASES])
def test_fixed_effect_matches_reference_test_restatement(c):
    """a11: the reference's own FE oracle (test_fixed_effect_lr_lbfgs_model.py:480-528)."""
    k = c["key"]
    rows = O.FeBlock(c["n"], c["D"], FE_ARR[k + "_rowptr"], FE_ARR[k + "_col"], FE_ARR[k + "_val"],
                     FE_ARR[k + "_y"], FE_ARR[k + "_w"], FE_ARR[k + "_off"],
                     linear_regression=c["linear_regression"])
    # the reference test regularises every coefficient including the intercept
    o = O.make_opts(l2=c["l2"], regularize_bias=True, has_intercept=c["has_intercept"], m=c["m"],
                    max_iter=c["max_iter"], factr=c["factr"])
    x, f, nit, nfev, status = O.fe_fit(rows, o, FE_ARR[k + "_x0"])
    assert (nit, nfev, status) == (c["nit"], c["nfev"], c["warnflag"])
    np.testing.assert_allclose(x, FE_ARR[k + "_theta"], rtol=1e-8, atol=1e-10)
    assert abs(f - c["f"]) <= 1e-11 * max(1.0, abs(c["f"]))


def test_partition_map_bit_exact():
    """abs(String.hashCode) % n with JVM semantics (PartitionUtils.scala:31-37)."""
    g = load_partition()
    for s, h in g["known_answers"].items():
        assert O.java_string_hash(s) == h
    for s, h in g["hash"].items():
        assert O.java_string_hash(s) == h, s
    for n, table in g["partition"].items():
        for s, pid in table.items():
            assert O.partition_id(s, int(n)) == pid, (s, n)
    # the Int.MinValue quirk: abs() stays negative, % keeps the sign
    assert O.partition_id("polygenelubricants", 3) == -2
    assert O.partition_id("polygenelubricants", 10) == -8


@pytest.mark.parametrize("c", CASES[::3], ids=[c["name"] for c in CASES[::3]])
def test_scipy_port_is_the_reference(c):
    """oracle/scipy_port.py (the CPU-baseline arm of bench.py) repeats the reference's library calls in the
    reference's order, so on pinned cases it must reproduce the reference's coefficients to the last bits."""
    from oracle import scipy_port as SP
    k = c["key"]
    th0 = ARR[k + "_theta0"] if c["warm"] else None
    theta, f, nit, nfev, wf = SP.fit_entity(c["n"], c["d"], ARR[k + "_rowptr"], ARR[k + "_col"], ARR[k + "_val"],
                                            ARR[k + "_y"], ARR[k + "_w"], ARR[k + "_off"], l2=c["l2"],
                                            regularize_bias=c["regularize_bias"], has_intercept=c["has_intercept"],
                                            m=c["m"], max_iter=c["max_iter"], tol=c["tol"], theta0=th0)
    if is_pinned(c):
        assert (nit, nfev, wf) == (c["nit"], c["nfev"], c["warnflag"])
        np.testing.assert_allclose(theta, ARR[k + "_theta"], rtol=1e-12, atol=1e-15)


def test_reference_arm_runs_the_staged_reference_class_and_never_loads_the_product():
    """bench.py --impl reference: oracle/ref_arm.py drives the reference's own BinaryLogisticRegressionTrainer when
    oracle/_ref is staged (oracle/build_ref.py), reproduces the golden coefficients, and neither it nor the numpy
    workload generator it uses imports gdmix_b200 (a reference arm must not map the product's library)."""
    import subprocess
    import sys
    from oracle import build_ref, ref_arm
    cold = [c for c in CASES if not c["warm"] and is_pinned(c) and c["has_intercept"] and c["max_iter"] == 100
            and c["m"] == 10 and c["tol"] == 1e-12][:6]
    assert cold
    for c in cold:
        k = c["key"]
        p = c["d"] + 1
        batch = {"ent_rowptr": np.array([0, c["n"]], np.int64), "rowptr": ARR[k + "_rowptr"], "col": ARR[k + "_col"],
                 "val": ARR[k + "_val"], "y": ARR[k + "_y"], "w": ARR[k + "_w"], "off": ARR[k + "_off"],
                 "theta_ptr": np.array([0, p], np.int64)}
        th = ref_arm.fit_entities(batch, 0, 1, l2=c["l2"], regularize_bias=c["regularize_bias"], has_intercept=True)[0]
        want = ARR[k + "_theta"].copy()
        want[np.abs(want) <= 1e-4] = 0.0          # the arm applies threshold_coefficients like TrainingJobConsumer
        np.testing.assert_allclose(th, want, rtol=1e-12, atol=1e-15)
    if os.path.isdir(build_ref.REF_SRC):
        assert build_ref.build() and ref_arm.kind() == "reference"
    code = ("import sys, bench; g = bench._synth_arrays(); a, _ = g.make_arrays(4, 8, 16, 4); "
            "from oracle import ref_arm; ref_arm.fit_entities(bench._oracle_dict(a), 0, 2); "
            "bad = [m for m in sys.modules if m.startswith('gdmix_b200')]; "
            "maps = open('/proc/self/maps').read(); "
            "assert not bad and 'libgdmix_b200' not in maps, (bad, 'libgdmix_b200' in maps)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run([sys.executable, "-c", code], cwd=root, check=True)
