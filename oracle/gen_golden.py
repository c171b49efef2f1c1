#!/usr/bin/env python
"""Generate tests/golden/* by running the REFERENCE ITSELF in this container.

TEST INFRASTRUCTURE.  Runs only where /root/reference is mounted (never on the GPU
box).  It imports the reference's own ``BinaryLogisticRegressionTrainer``
(gdmix-trainer/src/gdmix/models/custom/binary_logistic_regression.py) and drives it
exactly like ``TrainingJobConsumer.__call__`` does (job_consumers.py:36-63): scipy COO
matrix in entity-local index space, fp32 weights/offsets, ``fit`` -> L-BFGS-B through
``scipy.optimize.fmin_l_bfgs_b``.  scipy here is 1.18.1 (the reference pins 1.5.4,
not installable offline); its removed ``disp`` kwarg is dropped by a one-line shim,
nothing else in the reference is touched.

Outputs (committed):
  tests/golden/re_golden.npz / re_golden.json   random-effect solver cases
  tests/golden/fe_golden.npz / fe_golden.json   fixed-effect problem of the reference's
                                                own test (test_fixed_effect_lr_lbfgs_model.py:379-528)
  tests/golden/partition_golden.json            Java String.hashCode partition map values
Usage:  python oracle/gen_golden.py
"""
import json
import zlib
import os
import sys

import numpy as np
import scipy.optimize
import scipy.sparse as sparse
from scipy.special import expit

REF_SRC = "/root/reference/gdmix-trainer/src"
sys.path.insert(0, REF_SRC)
from gdmix.models.custom import binary_logistic_regression as blr  # noqa: E402
from gdmix.util.model_utils import threshold_coefficients  # noqa: E402

# scipy >= 1.15 dropped fmin_l_bfgs_b(disp=...); the reference passes disp=0.
blr.fmin_l_bfgs_b = lambda *a, disp=None, **k: scipy.optimize.fmin_l_bfgs_b(*a, **k)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")
EPS = np.finfo(float).eps


def run_reference(case):
    """case: dict(n,d,rowptr,col,val,y,w,off, l2, regularize_bias, has_intercept, m, max_iter, tol, theta0)."""
    n, d = case["n"], case["d"]
    rowptr, col, val = case["rowptr"], case["col"], case["val"]
    rows = np.repeat(np.arange(n), np.diff(rowptr))
    # job_consumers.py:247  X = coo_matrix((values, (rows, locally_indexed_cols)))  (fp32 values)
    X = sparse.coo_matrix((val.astype(np.float32), (rows, col)), shape=(n, d))
    y = case["y"].astype(np.int64)
    w = case["w"].astype(np.float32)
    off = case["off"].astype(np.float32)
    tr = blr.BinaryLogisticRegressionTrainer(lambda_l2=case["l2"], precision=case["tol"] / EPS,
                                             num_lbfgs_corrections=case["m"], max_iter=case["max_iter"],
                                             regularize_bias=case["regularize_bias"],
                                             has_intercept=case["has_intercept"])
    theta0 = None if case.get("theta0") is None else np.array(case["theta0"], dtype=np.float64)
    (theta, f, info), _ = tr.fit(X, y, weights=w, offsets=off, theta_initial=theta0, variance_mode=None)
    out = {"theta": np.array(theta), "f": float(f), "nit": int(info["nit"]), "nfev": int(info["funcalls"]),
           "warnflag": int(info["warnflag"]), "g": np.array(info["grad"])}
    out["theta_thresholded"] = threshold_coefficients(theta, 1e-4)
    # probe point for _loss/_gradient parity
    rng = np.random.RandomState(case["seed"] + 7)
    p = d + (1 if case["has_intercept"] else 0)
    probe = rng.normal(0, 0.7, size=p)
    X1 = tr._add_column_of_ones(X) if case["has_intercept"] else X
    out["probe"] = probe
    out["probe_f"] = float(tr._loss(probe, X1, y, w, off))
    out["probe_g"] = np.array(tr._gradient(probe, X1, y, w, off))
    out["logits"] = np.array(tr.predict_proba(X, off, custom_theta=theta, return_logits=True))
    # How reproducible is the reference itself?  Re-run it from a start point moved by 1e-17 (far below
    # one ulp of any coefficient of interest).  Entities whose optimum is at infinity (e.g. one sample and
    # an unregularised intercept) amplify that to O(1): their coefficients are not pinned by the reference
    # to 1e-5, so the parity tests compare objective values there instead (see tests/test_oracle.py).
    pert0 = (np.zeros(p) if theta0 is None else theta0) + 1e-17
    (theta_p, _, info_p), _ = tr.fit(X, y, weights=w, offsets=off, theta_initial=pert0, variance_mode=None)
    out["self_sensitivity"] = float(np.linalg.norm(theta_p - theta) / max(np.linalg.norm(theta), 1e-300))
    out["self_nit_changed"] = bool(info_p["nit"] != info["nit"])
    if case.get("variance", False) and p <= 64:
        out["var_simple"] = np.array(tr._compute_variance(X1, y, theta, w, off, "simple"))
        try:
            out["var_full"] = np.array(tr._compute_variance(X1, y, theta, w, off, "full"))
        except np.linalg.LinAlgError:
            pass
    return out


def synth_entity(seed, n, d, k, dense=False, weights=False, all_same_label=None, empty_rows=False,
                 unsorted_cols=False):
    rng = np.random.RandomState(seed)
    rowptr = [0]
    cols, vals = [], []
    for i in range(n):
        ki = d if dense else min(d, k)
        if empty_rows and i % 3 == 1 and i != n - 1:
            ki = 0
        c = rng.choice(d, size=ki, replace=False) if ki else np.zeros(0, int)
        if not unsorted_cols:
            c = np.sort(c)
        cols.append(c)
        vals.append(rng.normal(0, 1, size=ki).astype(np.float32))
        rowptr.append(rowptr[-1] + ki)
    col = np.concatenate(cols).astype(np.int32) if cols else np.zeros(0, np.int32)
    val = np.concatenate(vals).astype(np.float32) if vals else np.zeros(0, np.float32)
    theta_star = rng.normal(0, 0.5, size=d + 1)
    off = rng.normal(0, 1, size=n).astype(np.float32)
    rowsi = np.repeat(np.arange(n), np.diff(rowptr))
    z = np.bincount(rowsi, weights=val.astype(np.float64) * theta_star[1 + col], minlength=n) + theta_star[0] + off
    y = (rng.uniform(size=n) < expit(z)).astype(np.float32)
    if all_same_label is not None:
        y[:] = all_same_label
    w = rng.uniform(0.5, 2.0, size=n).astype(np.float32) if weights else np.ones(n, np.float32)
    return dict(n=n, d=d, rowptr=np.array(rowptr, np.int64), col=col, val=val, y=y, w=w, off=off, seed=seed)


def fixture_cases():
    """Entities of the reference's own test fixtures (decoded by hand from
    test/resources/grouped_per_member_train/data.tfrecord -- tests/test_tfrecord_io.py re-decodes
    the same bytes from a committed copy of the *decoded values* -- and the literal datasets at
    test_random_effect_lr_lbfgs_model.py:169-194)."""
    raw = [
        ("fixture:100034", [[0, 7, 60, 80, 95], [34, 57]], [[1, 2, 3, 5, 6.6], [1, 2]], [0, 1], [1, 2], [0.5, 0.75]),
        ("fixture:100", [[10, 11]], [[-3.5, 2.3]], [1], [1], [0.2]),
        ("dataset1:xyz", [[0, 2], [0, 1], [1, 2], [0, 2], [1, 2], [0, 1]],
         [[0.55, -0.95], [0.22, -1.05], [0.90, 0.50], [1.99, 0.48], [0.37, -1.64], [0.33, 0.17]],
         [1, 0, 1, 0, 0, 1], [1.0, 0.8, 2.0, 3.0, 2.1, 1.7], [1.0, 2.0, 3.0, -1.0, 0.3, -0.7]),
        ("dataset2:abc102", [[1, 5, 10], [1, 50, 99]], [[0.3, -2.3, 0.9], [1.4, 99.8, -1.2]], [1, 0], [1.0, 0.8],
         [1.0, 2.0]),
        ("dataset2:zyz234", [[1, 3], [2, 20]], [[1.23, 4.5], [-1.0, 3.0]], [0, 0], [0.5, 0.74], [-1.0, -2.0]),
    ]
    cases = []
    for name, idx, vals, y, w, off in raw:
        flat = np.concatenate([np.asarray(c) for c in idx])
        uniq, local = np.unique(flat, return_inverse=True)  # job_consumers.py:243
        rowptr = np.concatenate([[0], np.cumsum([len(c) for c in idx])]).astype(np.int64)
        base = dict(n=len(idx), d=len(uniq), rowptr=rowptr, col=local.astype(np.int32),
                    val=np.concatenate([np.asarray(v, np.float32) for v in vals]),
                    y=np.asarray(y, np.float32), w=np.asarray(w, np.float32), off=np.asarray(off, np.float32),
                    unique_global_indices=uniq.tolist(), seed=zlib.crc32(name.encode()) % 1000)
        # the RE test parameters: l2=0.1, LRParams default regularize_bias=True, intercept, cold start
        for l2, maxit in ((0.1, 100), (0.0, 100), (0.1, 1)):
            c = dict(base, name=f"{name}/l2={l2}/maxiter={maxit}", l2=l2, regularize_bias=True, has_intercept=True,
                     m=10, max_iter=maxit, tol=1e-12, variance=True)
            cases.append(c)
        cases.append(dict(base, name=f"{name}/nointercept", l2=0.1, regularize_bias=False, has_intercept=False,
                          m=10, max_iter=100, tol=1e-12, variance=True))
    return cases


def synthetic_cases():
    cases = []
    sid = 1000
    shapes = [(6, 3, 2), (1, 4, 3), (2, 1, 1), (16, 24, 4), (33, 17, 5), (64, 40, 8), (128, 256, 32),
              (200, 70, 16), (40, 300, 12)]
    for (n, d, k) in shapes:
        for l2 in (0.1, 1.0, 10.0, 100.0):
            for variant in range(2):
                sid += 1
                e = synth_entity(sid, n, d, k, weights=(variant == 1), unsorted_cols=(variant == 1))
                cases.append(dict(e, name=f"synth/{n}x{d}k{k}/l2={l2}/v{variant}", l2=l2,
                                  regularize_bias=(variant == 1), has_intercept=True, m=10, max_iter=100, tol=1e-12,
                                  variance=(d <= 40)))
    # more C1-shaped entities (the headline shape) at the MovieLens/C1 settings
    for i in range(12):
        sid += 1
        e = synth_entity(sid, 128, 256, 32)
        cases.append(dict(e, name=f"synth/c1/{i}", l2=1.0, regularize_bias=False, has_intercept=True, m=10,
                          max_iter=100, tol=1e-12))
    # edge cases
    sid += 1
    cases.append(dict(synth_entity(sid, 24, 12, 4, all_same_label=1.0), name="edge/all_ones", l2=1.0,
                      regularize_bias=False, has_intercept=True, m=10, max_iter=100, tol=1e-12, variance=True))
    sid += 1
    cases.append(dict(synth_entity(sid, 24, 12, 4, all_same_label=0.0), name="edge/all_zeros_unreg", l2=0.0,
                      regularize_bias=False, has_intercept=True, m=10, max_iter=100, tol=1e-12))
    sid += 1
    cases.append(dict(synth_entity(sid, 30, 10, 3, empty_rows=True), name="edge/empty_rows", l2=0.5,
                      regularize_bias=False, has_intercept=True, m=10, max_iter=100, tol=1e-12, variance=True))
    sid += 1
    cases.append(dict(synth_entity(sid, 50, 20, 20, dense=True), name="edge/dense", l2=1.0, regularize_bias=True,
                      has_intercept=True, m=10, max_iter=100, tol=1e-12, variance=True))
    sid += 1
    cases.append(dict(synth_entity(sid, 50, 20, 5), name="edge/m3", l2=0.01, regularize_bias=False,
                      has_intercept=True, m=3, max_iter=100, tol=1e-12))
    sid += 1
    cases.append(dict(synth_entity(sid, 50, 20, 5), name="edge/maxiter5", l2=0.01, regularize_bias=False,
                      has_intercept=True, m=10, max_iter=5, tol=1e-12))
    sid += 1
    cases.append(dict(synth_entity(sid, 50, 20, 5), name="edge/loose_tol", l2=0.1, regularize_bias=False,
                      has_intercept=True, m=10, max_iter=100, tol=1e-3))
    sid += 1
    cases.append(dict(synth_entity(sid, 60, 30, 6), name="edge/no_intercept", l2=1.0, regularize_bias=False,
                      has_intercept=False, m=10, max_iter=100, tol=1e-12, variance=True))
    # intercept-only model: X is an n x 1 zero column (job_consumers.py:213-218)
    sid += 1
    e = synth_entity(sid, 20, 1, 1)
    e["val"][:] = 0.0
    cases.append(dict(e, name="edge/intercept_only", l2=0.1, regularize_bias=True, has_intercept=True, m=10,
                      max_iter=100, tol=1e-12))
    # large offsets / separable-ish data (exercises the stable CE branch and longer line searches)
    sid += 1
    e = synth_entity(sid, 40, 8, 3)
    e["off"] = (e["off"] * 30).astype(np.float32)
    cases.append(dict(e, name="edge/huge_offsets", l2=0.001, regularize_bias=False, has_intercept=True, m=10,
                      max_iter=100, tol=1e-12))
    sid += 1
    e = synth_entity(sid, 40, 8, 3)
    e["val"] = (e["val"] * 25).astype(np.float32)
    cases.append(dict(e, name="edge/large_values", l2=0.001, regularize_bias=False, has_intercept=True, m=10,
                      max_iter=100, tol=1e-12))
    return cases


def warm_start_cases(cold_cases):
    """test_binary_logistic_regression.py:180-205 / test_random_effect...:231-350: one iteration
    from the converged model stays put; one iteration from zeros does not."""
    out = []
    for c in cold_cases[:6]:
        ref = run_reference(c)
        out.append(dict(c, name=c["name"] + "/warm1", theta0=ref["theta"].tolist(), max_iter=1))
        out.append(dict(c, name=c["name"] + "/warm_thr", theta0=ref["theta_thresholded"].tolist(), max_iter=100))
    return out


def pack(cases, prefix):
    arrays, manifest = {}, []
    for i, c in enumerate(cases):
        ref = run_reference(c)
        key = f"c{i:03d}"
        for k in ("rowptr", "col", "val", "y", "w", "off"):
            arrays[f"{key}_{k}"] = c[k]
        if c.get("theta0") is not None:
            arrays[f"{key}_theta0"] = np.asarray(c["theta0"], np.float64)
        for k, v in ref.items():
            if isinstance(v, np.ndarray):
                arrays[f"{key}_{k}"] = v
        manifest.append(dict(key=key, name=c["name"], n=int(c["n"]), d=int(c["d"]), l2=c["l2"],
                             regularize_bias=bool(c["regularize_bias"]), has_intercept=bool(c["has_intercept"]),
                             m=c["m"], max_iter=c["max_iter"], tol=c["tol"], warm=c.get("theta0") is not None,
                             f=ref["f"], nit=ref["nit"], nfev=ref["nfev"], warnflag=ref["warnflag"],
                             probe_f=ref["probe_f"], self_sensitivity=ref["self_sensitivity"],
                             self_nit_changed=ref["self_nit_changed"], has_var_simple="var_simple" in ref,
                             has_var_full="var_full" in ref,
                             unique_global_indices=c.get("unique_global_indices")))
    np.savez_compressed(os.path.join(OUT, prefix + ".npz"), **arrays)
    with open(os.path.join(OUT, prefix + ".json"), "w") as fh:
        json.dump({"scipy": scipy.__version__, "numpy": np.__version__,
                   "generator": "oracle/gen_golden.py (reference BinaryLogisticRegressionTrainer + scipy fmin_l_bfgs_b)",
                   "cases": manifest}, fh, indent=1)
    return manifest


# ---- fixed effect: the reference test's own numpy/scipy restatement ------------------

def fe_cases():
    """Follows test_fixed_effect_lr_lbfgs_model.py:379-421 (data, seed 0) and :480-528 (solver):
    100 x 10, density 0.1, intercept column LAST, l2 = 1, m = 10, factr = 1e-12 (sic: the test passes
    _PRECISION=1e-12 as factr), objective NOT divided by n."""
    out = []
    for (has_offset, has_intercept, model_type, seed) in ((True, True, "logistic_regression", 0),
                                                          (False, True, "logistic_regression", 0),
                                                          (True, False, "logistic_regression", 0),
                                                          (True, True, "linear_regression", 0),
                                                          (True, True, "logistic_regression", 3)):
        np.random.seed(seed)
        F = sparse.random(100, 10, density=0.1).toarray()
        _ = sparse.random(100, 10, density=0.1).toarray()  # validation features (consumes RNG like the test)
        labels = np.random.randint(2, size=100)
        _ = np.random.randint(2, size=100)
        if model_type == "linear_regression":
            labels = labels.astype(np.float64)
        offs = np.random.rand(100) if has_offset else np.zeros(100)
        F32 = F.astype(np.float32)          # what reaches the TFRecord (float_list)
        offs32 = offs.astype(np.float32)
        F1 = np.hstack((F32.astype(np.float64), np.ones((100, 1)))) if has_intercept else F32.astype(np.float64)

        def _loss(theta):
            pred = F1.dot(theta) + offs32
            if model_type == "logistic_regression":
                loss = np.maximum(pred, 0) - pred * labels + np.log(1 + np.exp(-np.absolute(pred)))
            else:
                loss = np.square(labels.astype(np.float64) - pred)
            return loss.sum() + 1.0 / 2.0 * theta.dot(theta)

        def _grad(theta):
            logit = F1.dot(theta) + offs32
            if model_type == "logistic_regression":
                cg = F1.T.dot(expit(logit) - labels)
            else:
                cg = 2.0 * F1.T.dot(logit - labels)
            return cg + 1.0 * theta

        for maxit, x0 in ((100, None), (1, "prev")):
            if x0 == "prev":
                x0v = out[-1]["theta"]
            else:
                x0v = np.zeros(F1.shape[1])
            res = scipy.optimize.fmin_l_bfgs_b(func=_loss, x0=x0v, approx_grad=False, fprime=_grad, m=10,
                                               factr=1e-12, maxiter=maxit)
            csr = sparse.csr_matrix(F32)
            out.append(dict(name=f"fe/off={has_offset}/icpt={has_intercept}/{model_type}/seed{seed}/maxiter{maxit}",
                            n=100, D=10, rowptr=csr.indptr.astype(np.int64), col=csr.indices.astype(np.int32),
                            val=csr.data.astype(np.float32), y=labels.astype(np.float32),
                            w=np.ones(100, np.float32), off=offs32, has_intercept=has_intercept,
                            linear_regression=(model_type == "linear_regression"), l2=1.0, m=10, max_iter=maxit,
                            factr=1e-12, x0=x0v, theta=np.array(res[0]), f=float(res[1]), nit=int(res[2]["nit"]),
                            nfev=int(res[2]["funcalls"]), warnflag=int(res[2]["warnflag"])))
    return out


def pack_fe(cases):
    arrays, manifest = {}, []
    for i, c in enumerate(cases):
        key = f"f{i:03d}"
        for k in ("rowptr", "col", "val", "y", "w", "off", "x0", "theta"):
            arrays[f"{key}_{k}"] = c[k]
        manifest.append({k: c[k] for k in ("name", "n", "D", "has_intercept", "linear_regression", "l2", "m",
                                           "max_iter", "factr", "f", "nit", "nfev", "warnflag")} | {"key": key})
    np.savez_compressed(os.path.join(OUT, "fe_golden.npz"), **arrays)
    with open(os.path.join(OUT, "fe_golden.json"), "w") as fh:
        json.dump({"scipy": scipy.__version__, "cases": manifest}, fh, indent=1)


# ---- partition map: values follow from the Java language specification ---------------

def java_hash(s):
    h = 0
    for u in np.frombuffer(s.encode("utf-16-le"), dtype=np.uint16):
        h = (31 * h + int(u)) & 0xFFFFFFFF
    return h - (1 << 32) if h >= (1 << 31) else h


def jvm_partition(s, n):
    h = java_hash(s)
    a = h if h == -(1 << 31) else abs(h)
    return int(np.fmod(a, n))  # sign of the dividend, like the JVM's %


def partition_golden():
    ids = ["", "0", "1", "abc", "100", "100034", "xyz", "abc102", "zyz234", "polygenelubricants", "memberId",
           "9223372036854775807", "-17", "user_000123", "été", "\U0001F600x"]
    ids += [str(i) for i in range(1000, 1040)]
    return {"spec": "JLS String.hashCode: s[0]*31^(n-1)+...+s[n-1] over UTF-16 code units, int32 wrap; "
                    "partition = abs(hash) % n with Math.abs(Int.MinValue) == Int.MinValue "
                    "(PartitionUtils.scala:31-37)",
            "known_answers": {"": 0, "abc": 96354, "polygenelubricants": -2147483648},
            "hash": {s: java_hash(s) for s in ids},
            "partition": {str(n): {s: jvm_partition(s, n) for s in ids} for n in (1, 3, 10, 64, 512)}}


def main():
    os.makedirs(OUT, exist_ok=True)
    cold = fixture_cases() + synthetic_cases()
    cases = cold + warm_start_cases(cold)
    man = pack(cases, "re_golden")
    print(f"re_golden: {len(man)} cases")
    fe = fe_cases()
    pack_fe(fe)
    print(f"fe_golden: {len(fe)} cases")
    with open(os.path.join(OUT, "partition_golden.json"), "w") as fh:
        json.dump(partition_golden(), fh, indent=1, ensure_ascii=True)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
