"""ctypes front-end of the CPU oracle (oracle/lr_oracle.c) -- TEST INFRASTRUCTURE.

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs import this module.  Nothing under ``gdmix_b200/``
does: the product path has no CPU fallback.

Besides the C restatement this module holds the small numpy restatements of the
reference's host-side slicing rules that the parity tests need:

* ``build_local_block``      job_consumers.py:243-258 (np.unique -> local columns)
* ``warm_start_theta``       job_consumers.py:262-288
* ``sparsify_theta``         job_consumers.py:87-99
* ``java_string_hash`` / ``partition_id``   PartitionUtils.scala:31-37
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liblr_oracle.so")
_SRC = os.path.join(_HERE, "lr_oracle.c")


def build(force=False):
    """Compile lr_oracle.c with gcc (no FMA contraction, so that the arithmetic
    is the plain IEEE sequence the restatement spells out)."""
    if not force and os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(_SRC):
        return _SO
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    cmd = ["gcc", "-O2", "-std=c99", "-fPIC", "-shared", "-ffp-contract=off", "-fvisibility=hidden",
           "-o", _SO, _SRC, "-lm"]
    subprocess.check_call(cmd)
    return _SO


class Opts(C.Structure):
    _fields_ = [("l2", C.c_double), ("regularize_bias", C.c_int32), ("has_intercept", C.c_int32),
                ("m", C.c_int32), ("max_iter", C.c_int32), ("max_ls", C.c_int32), ("max_fun", C.c_int32),
                ("factr", C.c_double), ("pgtol", C.c_double)]


class Block(C.Structure):
    _fields_ = [("n", C.c_int64), ("d", C.c_int64), ("rowptr", C.c_void_p), ("col", C.c_void_p),
                ("val", C.c_void_p), ("y", C.c_void_p), ("w", C.c_void_p), ("off", C.c_void_p)]


class FeRows(C.Structure):
    _fields_ = [("n", C.c_int64), ("D", C.c_int64), ("rowptr", C.c_void_p), ("col", C.c_void_p),
                ("val", C.c_void_p), ("y", C.c_void_p), ("w", C.c_void_p), ("off", C.c_void_p),
                ("linear_regression", C.c_int32), ("num_workers", C.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.oracle_re_loss_grad.restype = C.c_double
        _lib.oracle_fe_loss_grad.restype = C.c_double
        _lib.oracle_java_string_hash.restype = C.c_int32
        _lib.oracle_partition_id.restype = C.c_int32
    return _lib


EPS = float(np.finfo(float).eps)


def make_opts(l2=1.0, regularize_bias=False, has_intercept=True, m=10, max_iter=100, tol=1e-12,
              factr=None, pgtol=1e-5, max_ls=20, max_fun=15000):
    """Same defaults the reference hands to scipy (random_effect_lr_lbfgs_model.py:142-146;
    pgtol/maxls/maxfun are scipy's defaults because the reference never sets them)."""
    if factr is None:
        factr = tol / EPS
    return Opts(l2, int(bool(regularize_bias)), int(bool(has_intercept)), m, max_iter, max_ls, max_fun,
                float(factr), float(pgtol))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class EntityBlock:
    """Keeps the numpy arrays alive next to the C struct."""

    def __init__(self, n, d, rowptr, col, val, y, w=None, off=None):
        self.n, self.d = int(n), int(d)
        self.rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
        self.col = np.ascontiguousarray(col, dtype=np.int32)
        self.val = np.ascontiguousarray(val, dtype=np.float32)
        self.y = np.ascontiguousarray(y, dtype=np.float32)
        self.w = np.ones(self.n, np.float32) if w is None else np.ascontiguousarray(w, dtype=np.float32)
        self.off = np.zeros(self.n, np.float32) if off is None else np.ascontiguousarray(off, dtype=np.float32)
        assert self.rowptr.shape == (self.n + 1,)
        self.c = Block(self.n, self.d, _p(self.rowptr), _p(self.col), _p(self.val), _p(self.y), _p(self.w),
                       _p(self.off))


def re_loss_grad(block, opts, theta):
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    g = np.zeros_like(theta)
    f = lib().oracle_re_loss_grad(C.byref(block.c), C.byref(opts), _p(theta), _p(g))
    return f, g


def re_fit(block, opts, theta0=None):
    """-> (theta, f, nit, nfev, status, g)   (theta before thresholding)."""
    p = block.d + (1 if opts.has_intercept else 0)
    theta = np.zeros(p) if theta0 is None else np.array(theta0, dtype=np.float64, copy=True)
    assert theta.shape == (p,)
    f = C.c_double()
    info = np.zeros(4, np.int32)
    g = np.zeros(p)
    lib().oracle_re_fit(C.byref(block.c), C.byref(opts), _p(theta), C.byref(f), _p(info), _p(g))
    return theta, f.value, int(info[0]), int(info[1]), int(info[2]), g


def re_fit_batch(batch, opts, theta0=None, e0=0, e1=None):
    """batch: dict with ent_rowptr,rowptr,col,val,y,w,off,theta_ptr (numpy, host)."""
    E = len(batch["ent_rowptr"]) - 1
    e1 = E if e1 is None else e1
    tp = batch["theta_ptr"]
    theta = np.zeros(int(tp[-1]))
    f = np.zeros(E)
    nit = np.zeros(E, np.int32)
    nfev = np.zeros(E, np.int32)
    status = np.zeros(E, np.int32)
    t0 = None if theta0 is None else np.ascontiguousarray(theta0, dtype=np.float64)
    lib().oracle_re_fit_batch(C.c_int64(e0), C.c_int64(e1), _p(batch["ent_rowptr"]), _p(batch["rowptr"]),
                              _p(batch["col"]), _p(batch["val"]), _p(batch["y"]), _p(batch["w"]),
                              _p(batch["off"]), _p(tp), C.byref(opts), None if t0 is None else _p(t0),
                              _p(theta), _p(f), _p(nit), _p(nfev), _p(status))
    return theta, f, nit, nfev, status


def re_variance(block, opts, theta, mode):
    mode_i = {"simple": 1, "full": 2}[mode.lower()]
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    var = np.zeros_like(theta)
    rc = lib().oracle_re_variance(C.byref(block.c), C.byref(opts), _p(theta), mode_i, _p(var))
    if rc != 0:
        raise np.linalg.LinAlgError("singular Hessian")
    return var


def re_score(block, opts, theta):
    logit = np.zeros(block.n)
    per = np.zeros(block.n)
    th = None if theta is None else np.ascontiguousarray(theta, dtype=np.float64)
    lib().oracle_re_score(C.byref(block.c), C.byref(opts), None if th is None else _p(th), _p(logit), _p(per))
    return logit, per


def threshold(coef, thr=1e-4):
    out = np.array(coef, dtype=np.float64, copy=True)
    lib().oracle_threshold(_p(out), C.c_int64(out.size), C.c_double(thr))
    return out


class FeBlock:
    def __init__(self, n, D, rowptr, col, val, y, w=None, off=None, linear_regression=False, num_workers=1):
        self.n, self.D = int(n), int(D)
        self.rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
        self.col = np.ascontiguousarray(col, dtype=np.int32)
        self.val = np.ascontiguousarray(val, dtype=np.float32)
        self.y = np.ascontiguousarray(y, dtype=np.float32)
        self.w = np.ones(self.n, np.float32) if w is None else np.ascontiguousarray(w, dtype=np.float32)
        self.off = np.zeros(self.n, np.float32) if off is None else np.ascontiguousarray(off, dtype=np.float32)
        self.c = FeRows(self.n, self.D, _p(self.rowptr), _p(self.col), _p(self.val), _p(self.y), _p(self.w),
                        _p(self.off), int(linear_regression), int(num_workers))


def fe_loss_grad(rows, opts, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    g = np.zeros_like(x)
    f = lib().oracle_fe_loss_grad(C.byref(rows.c), C.byref(opts), _p(x), _p(g))
    return f, g


def fe_fit(rows, opts, x0=None):
    p = rows.D + (1 if opts.has_intercept else 0)
    x = np.zeros(p) if x0 is None else np.array(x0, dtype=np.float64, copy=True)
    f = C.c_double()
    info = np.zeros(4, np.int32)
    lib().oracle_fe_fit(C.byref(rows.c), C.byref(opts), _p(x), C.byref(f), _p(info))
    return x, f.value, int(info[0]), int(info[1]), int(info[2])


# ---- host-side slicing rules of the reference, restated in numpy -------------

def build_local_block(sample_cols, sample_vals, y, w=None, off=None):
    """job_consumers.py:243-258: per-sample lists of global feature ids / values ->
    (EntityBlock in local index space, unique_global_indices)."""
    n = len(sample_cols)
    rowptr = np.zeros(n + 1, np.int64)
    for i, c in enumerate(sample_cols):
        rowptr[i + 1] = rowptr[i] + len(c)
    cols = np.concatenate([np.asarray(c, np.int64) for c in sample_cols]) if n else np.zeros(0, np.int64)
    vals = np.concatenate([np.asarray(v, np.float32) for v in sample_vals]) if n else np.zeros(0, np.float32)
    uniq, local = np.unique(cols, return_inverse=True)
    return EntityBlock(n, len(uniq), rowptr, local.astype(np.int32), vals, y, w, off), uniq


def warm_start_theta(prior_theta, prior_indices, unique_global_indices, has_intercept=True):
    """job_consumers.py:262-288 (local indexing): keep the prior's intercept and the
    coefficients of features that are present in the current data; others start at 0."""
    hi = 1 if has_intercept else 0
    theta0 = np.zeros(len(unique_global_indices) + hi)
    if has_intercept:
        theta0[0] = prior_theta[0]
    prior = {int(u): float(v) for u, v in zip(prior_indices, prior_theta[hi:])}
    for i, u in enumerate(unique_global_indices):
        if int(u) in prior:
            theta0[hi + i] = prior[int(u)]
    return theta0


def java_string_hash(s):
    u = np.frombuffer(s.encode("utf-16-le"), dtype=np.uint16)
    u = np.ascontiguousarray(u)
    return int(lib().oracle_java_string_hash(_p(u) if u.size else None, C.c_int64(u.size)))


def partition_id(s, num_partitions):
    u = np.ascontiguousarray(np.frombuffer(s.encode("utf-16-le"), dtype=np.uint16))
    return int(lib().oracle_partition_id(_p(u) if u.size else None, C.c_int64(u.size), C.c_int32(num_partitions)))
