"""The reference's per-entity training path restated with the same third-party calls -- TEST
INFRASTRUCTURE and the CPU-baseline arm of bench.py (``--impl reference`` / ``cpu_baseline``).

/root/reference does not exist on the GPU box and its sources may not be copied, so this module
restates, op for op, what the reference executes per entity on its CPU path:

  TrainingJobConsumer.__call__            gdmix/models/custom/scipy/job_consumers.py:36-63
  BinaryLogisticRegressionTrainer.fit     gdmix/models/custom/binary_logistic_regression.py:191-239
      scipy.sparse.hstack ones column first (:133-142), fmin_l_bfgs_b(func=_loss, fprime=_gradient)
      with _loss (:84-110) and _gradient (:121-131) as separate callbacks (X.theta is computed twice
      per evaluation, as in the reference), scipy COO matvecs, scipy.special.expit
  threshold_coefficients                  gdmix/util/model_utils.py:4-12

It is pinned by tests/test_oracle.py::test_scipy_port_is_the_reference against the golden vectors the
real reference produced (bit-identical coefficients expected: same library calls in the same order).
The TF reader, Manager queue and Avro writer of the reference are NOT included, so this over-states
the reference's throughput.
"""
import os
import time

import numpy as np
import scipy.sparse
from scipy.optimize import fmin_l_bfgs_b
from scipy.special import expit

EPS = float(np.finfo(float).eps)


def _reg_loss(theta, l2, has_intercept, regularize_bias):
    if has_intercept and not regularize_bias:
        return (l2 / 2.0) * theta[1:].dot(theta[1:])
    return (l2 / 2.0) * theta.dot(theta)


def _loss(theta, X, y, weights, offsets, l2, has_intercept, regularize_bias):
    n = X.shape[0]
    pred = X.dot(theta) + offsets
    ce = np.maximum(pred, 0) - pred * y + np.log(1 + np.exp(-np.absolute(pred)))
    cost = weights * ce
    return (1.0 / n) * (cost.sum() + _reg_loss(theta, l2, has_intercept, regularize_bias))


def _gradient(theta, X, y, weights, offsets, l2, has_intercept, regularize_bias):
    n = X.shape[0]
    predictions = expit(X.dot(theta) + offsets)
    cost_grad = X.T.dot(weights * (predictions - y))
    reg = l2 * theta
    if has_intercept and not regularize_bias:
        reg[0] = 0
    return (1.0 / n) * (cost_grad + reg)


def fit_entity(n, d, rowptr, col, val, y, w, off, l2=1.0, regularize_bias=False, has_intercept=True, m=10,
               max_iter=100, tol=1e-12, theta0=None, threshold=None):
    """One entity, exactly the reference's call sequence.  -> (theta, f, nit, nfev, warnflag)."""
    rows = np.repeat(np.arange(n), np.diff(rowptr))
    X = scipy.sparse.coo_matrix((val, (rows, col)), shape=(n, d))           # job_consumers.py:247
    X1 = scipy.sparse.hstack((np.ones((n, 1)), X)) if has_intercept else X   # :133-142, :219
    if theta0 is None:
        theta0 = np.zeros(X1.shape[1])
    theta, f, info = fmin_l_bfgs_b(func=_loss, x0=theta0, approx_grad=False, fprime=_gradient, m=m,
                                   factr=tol / EPS, maxiter=max_iter,
                                   args=(X1, y.astype(np.int64), w, off, l2, has_intercept, regularize_bias))
    if threshold is not None:
        theta = np.array([0.0 if abs(x) <= threshold else x for x in theta])  # model_utils.py:12
    return theta, f, info["nit"], info["funcalls"], info["warnflag"]


def _fit_range(args):
    batch, e0, e1, kw = args
    hi = 1 if kw.get("has_intercept", True) else 0
    out = []
    nit_sum = 0
    for e in range(e0, e1):
        r0, r1 = batch["ent_rowptr"][e], batch["ent_rowptr"][e + 1]
        q0, q1 = batch["rowptr"][r0], batch["rowptr"][r1]
        p = batch["theta_ptr"][e + 1] - batch["theta_ptr"][e]
        th, f, nit, nfev, wf = fit_entity(r1 - r0, p - hi, batch["rowptr"][r0:r1 + 1] - q0, batch["col"][q0:q1],
                                          batch["val"][q0:q1], batch["y"][r0:r1], batch["w"][r0:r1],
                                          batch["off"][r0:r1], threshold=1e-4, **kw)
        out.append(th)
        nit_sum += nit
    return e1 - e0, nit_sum, out


_POOL_BATCH = None


def _pool_init(batch):
    global _POOL_BATCH
    _POOL_BATCH = batch
    # one BLAS/OpenMP thread per worker process: the pool already uses every core, and letting each of
    # them spawn a full BLAS team oversubscribes the box ~cores-fold (measured: 9x slower)
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:  # pragma: no cover
        pass


def _pool_fit(args):
    e0, e1, kw = args
    n, nit, _ = _fit_range((_POOL_BATCH, e0, e1, kw))
    return n, nit


def timed_fit(batch, n_entities, cores=None, grain=16, **kw):
    """Solve entities [0, n_entities) of `batch` on `cores` processes (default: all).  -> dict with
    entities/s.  Workers inherit the batch by fork, so no per-job pickling is charged (the reference
    pickles every Job twice through a Manager queue)."""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    jobs = [(e, min(n_entities, e + grain), kw) for e in range(0, n_entities, grain)]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_pool_init, initargs=(batch,)) as pool:
        pool.map(_pool_fit, [(0, 1, kw)] * cores)  # spin the workers up outside the timed region
        t0 = time.perf_counter()
        res = pool.map(_pool_fit, jobs, chunksize=1)
        dt = time.perf_counter() - t0
    done = sum(r[0] for r in res)
    return {"entities": done, "seconds": dt, "entities_per_sec": done / dt, "cores": cores,
            "mean_nit": sum(r[1] for r in res) / max(done, 1)}
