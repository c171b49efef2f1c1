"""GPU parity of the fixed-effect path: gdmix_fe_loss_grad / gdmix_fe_score against the oracle and the
reference test's own restatement (test_fixed_effect_lr_lbfgs_model.py:480-528) recorded in tests/golden/."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
if not torch.cuda.is_available():  # pragma: no cover
    pytest.skip("needs a CUDA device", allow_module_level=True)

from gdmix_b200 import _capi as capi  # noqa: E402
from gdmix_b200.fe_solver import FixedEffectSolver  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tests.golden_util import load_fe  # noqa: E402

FE_ARR, FE_CASES = load_fe()


def _rows(c, num_workers=1):
    k = c["key"]
    return capi.DeviceFeRows(FE_ARR[k + "_rowptr"], FE_ARR[k + "_col"], FE_ARR[k + "_val"], FE_ARR[k + "_y"],
                             FE_ARR[k + "_w"], FE_ARR[k + "_off"], c["D"],
                             linear_regression=c["linear_regression"], num_workers=num_workers)


def _opts(c, mod):
    return mod.make_opts(l2=c["l2"], regularize_bias=True, has_intercept=c["has_intercept"], m=c["m"],
                         max_iter=c["max_iter"], factr=c["factr"])


@pytest.mark.parametrize("c", FE_CASES, ids=[c["name"] for c in FE_CASES])
def test_fe_fit_matches_reference_restatement(c):
    solver = FixedEffectSolver(_rows(c), _opts(c, capi))
    x, info = solver.fit(FE_ARR[c["key"] + "_x0"])
    assert (info["nit"], info["nfev"], info["status"]) == (c["nit"], c["nfev"], c["warnflag"])
    # the reference asserts in fp32 with assertAllClose's default 1e-6 (test_fixed_effect...:449-464)
    np.testing.assert_allclose(x, FE_ARR[c["key"] + "_theta"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(x, FE_ARR[c["key"] + "_theta"], rtol=1e-9, atol=1e-11)


def test_fe_loss_grad_large_random_matches_oracle():
    rng = np.random.default_rng(5)
    n, D, k = 20000, 3000, 12
    rowptr = np.arange(n + 1, dtype=np.int64) * k
    col = rng.integers(0, D, size=n * k).astype(np.int32)      # duplicates and hot columns included
    col[: n * k // 4] = rng.integers(0, 8, size=n * k // 4)     # contended scatter targets
    val = rng.standard_normal(n * k).astype(np.float32)
    y = (rng.random(n) < 0.4).astype(np.float32)
    w = rng.uniform(0.5, 2, n).astype(np.float32)
    off = rng.standard_normal(n).astype(np.float32)
    x = rng.standard_normal(D + 1) * 0.3
    for lin in (False, True):
        for nw in (1, 8):
            rows = capi.DeviceFeRows(rowptr, col, val, y, w, off, D, linear_regression=lin, num_workers=nw)
            po = capi.make_opts(l2=0.7, regularize_bias=False, has_intercept=True)
            fg = capi.fe_loss_grad_device(rows, po, torch.from_numpy(x).cuda())
            torch.cuda.synchronize()
            fg = fg.cpu().numpy()
            ob = O.FeBlock(n, D, rowptr, col, val, y, w, off, linear_regression=lin, num_workers=nw)
            f, g = O.fe_loss_grad(ob, O.make_opts(l2=0.7, regularize_bias=False, has_intercept=True), x)
            assert abs(fg[0] - f) <= 1e-11 * abs(f)
            np.testing.assert_allclose(fg[1:], g, rtol=1e-9, atol=1e-9)


def test_fe_score_matches_definition():
    c = FE_CASES[0]
    rows = _rows(c)
    x = FE_ARR[c["key"] + "_theta"]
    solver = FixedEffectSolver(rows, _opts(c, capi))
    logit, per = solver.score(x)
    k = c["key"]
    dense = np.zeros((c["n"], c["D"]))
    rp, col, val = FE_ARR[k + "_rowptr"], FE_ARR[k + "_col"], FE_ARR[k + "_val"]
    for i in range(c["n"]):
        dense[i, col[rp[i]:rp[i + 1]]] = val[rp[i]:rp[i + 1]]
    z = dense @ x[:-1] + x[-1]
    np.testing.assert_allclose(per, z.astype(np.float32), rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(logit, (z + FE_ARR[k + "_off"]).astype(np.float32), rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("pa", [dict(), dict(hz=40, hg=90, tile_rows=50)], ids=["default", "split"])
def test_tiled_path_rows_of_every_size(pa):
    """fe_z_kernel stages 32 rows per warp in shared memory (1024 hot / 256 cold entries at a time): blocks that fit,
    blocks that do not (walked from global memory), single rows longer than the stage, empty rows, a ragged last block
    -- against the oracle; the scoring walk (gdmix_fe_score) over the same rows against the definition."""
    rng = np.random.default_rng(17)
    D = 5000
    lens = np.concatenate([np.full(64, 32), rng.integers(0, 90, 200), [3000, 0, 1024, 1025, 1, 0, 2047],
                           rng.integers(20, 60, 37), [1500, 1500], np.zeros(33, np.int64), rng.integers(0, 5, 11)])
    n = len(lens)
    rowptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    nnz = int(rowptr[-1])
    col = rng.integers(0, D, nnz).astype(np.int32)
    val = rng.standard_normal(nnz).astype(np.float32)
    y = (rng.random(n) < 0.5).astype(np.float32)
    w = rng.uniform(0.5, 2.0, n).astype(np.float32)
    off = (0.1 * rng.standard_normal(n)).astype(np.float32)
    x = rng.standard_normal(D + 1) * 0.02
    rows = capi.DeviceFeRows(rowptr, col, val, y, w, off, D)
    opts = capi.make_opts(l2=0.3, regularize_bias=False, has_intercept=True)
    plan = capi.DeviceFeTilePlan(rows, **pa)
    xd = torch.from_numpy(x).cuda()
    b = capi.fe_loss_grad_device(rows, opts, xd, plan=plan).cpu().numpy()
    c = capi.fe_loss_grad_device(rows, opts, xd, plan=plan).cpu().numpy()
    np.testing.assert_array_equal(b, c)
    blk = O.FeBlock(n, D, rowptr, col, val, y, w, off)
    f_o, g_o = O.fe_loss_grad(blk, O.make_opts(l2=0.3, regularize_bias=False, has_intercept=True), x)
    np.testing.assert_allclose(b[0], f_o, rtol=1e-12)
    np.testing.assert_allclose(b[1:], g_o, rtol=1e-10, atol=1e-10)
    logit, per = capi.fe_score_device(rows, opts, xd)
    z = np.array([np.dot(val[rowptr[i]:rowptr[i + 1]].astype(np.float64), x[col[rowptr[i]:rowptr[i + 1]]]) for i in range(n)])
    np.testing.assert_allclose(per.cpu().numpy(), (z + x[-1]).astype(np.float32), rtol=2e-6, atol=1e-6)
    np.testing.assert_allclose(logit.cpu().numpy(), (z + x[-1] + off).astype(np.float32), rtol=2e-6, atol=1e-6)


def test_solver_frequency_ranking_is_invisible_to_the_caller():
    """With more features than the rows kernel keeps of x in shared memory, FixedEffectSolver renumbers this
    shard's features by falling frequency internally; value and gradient come back in the caller's order and
    equal the oracle's, and a short solve equals the un-ranked one."""
    rng = np.random.default_rng(23)
    n, D, k = 6000, 20000, 10
    rowptr = np.arange(n + 1, dtype=np.int64) * k
    col = (D - 1 - np.minimum((D ** rng.random(n * k) - 1).astype(np.int64), D - 1)).astype(np.int32)  # hot ids HIGH
    val = rng.standard_normal(n * k).astype(np.float32)
    y = (rng.random(n) < 0.4).astype(np.float32)
    rows = capi.DeviceFeRows(rowptr, col, val, y, None, None, D)
    opts = capi.make_opts(l2=2.0, regularize_bias=True, has_intercept=True, max_iter=8)
    solver = FixedEffectSolver(rows, opts)
    x = rng.standard_normal(D + 1) * 0.05
    f, g = solver.loss_grad(x)
    assert solver._perm is not None
    blk = O.FeBlock(n, D, rowptr, col, val, y, np.ones(n, np.float32), np.zeros(n, np.float32))
    f_o, g_o = O.fe_loss_grad(blk, O.make_opts(l2=2.0, regularize_bias=True, has_intercept=True), x)
    np.testing.assert_allclose(f, f_o, rtol=1e-12)
    np.testing.assert_allclose(g, g_o, rtol=1e-10, atol=1e-10)
    xa, ia = solver.fit()
    plain = FixedEffectSolver(rows, opts)
    plain.plan = capi.DeviceFeTilePlan(rows)  # skips _prepare: no ranking
    xb, ib = plain.fit()
    assert (ia["nit"], ia["nfev"]) == (ib["nit"], ib["nfev"])
    np.testing.assert_allclose(xa, xb, rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("c", FE_CASES, ids=[c["name"] for c in FE_CASES])
def test_device_lbfgs_retraces_the_host_state_machine(c):
    """gdmix_fe_lbfgs_* (solver state in HBM) against gdmix_lbfgs_* (the same state machine on the host) on every
    golden fixed-effect case: same iterations, evaluations, stop status; coefficients to rounding."""
    dev = FixedEffectSolver(_rows(c), _opts(c, capi), solver="device")
    host = FixedEffectSolver(_rows(c), _opts(c, capi), solver="host")
    xd, idv = dev.fit(FE_ARR[c["key"] + "_x0"])
    xh, ih = host.fit(FE_ARR[c["key"] + "_x0"])
    assert (idv["nit"], idv["nfev"], idv["status"]) == (ih["nit"], ih["nfev"], ih["status"])
    np.testing.assert_allclose(xd, xh, rtol=1e-10, atol=1e-12)
    assert abs(idv["f"] - ih["f"]) <= 1e-12 * abs(ih["f"])


def _zipf_shard(rng, n, D, k):
    rowptr = np.arange(n + 1, dtype=np.int64) * k
    col = np.minimum((D ** rng.random(n * k) - 1).astype(np.int32), D - 1)
    val = rng.standard_normal(n * k).astype(np.float32)
    xs = rng.standard_normal(D + 1) * 0.3
    z = (val.reshape(n, k) * xs[col.reshape(n, k)]).sum(1) + xs[-1]
    y = (rng.random(n) < 1 / (1 + np.exp(-z))).astype(np.float32)
    return rowptr, col, val, y


@pytest.mark.parametrize("m,max_iter", [(10, 25), (3, 12), (0, 5)])
def test_device_lbfgs_on_a_ranked_shard(m, max_iter):
    """More features than the rows kernel keeps in shared memory (the solver then works in falling-frequency feature
    order), several blocks of solver vectors, memory wrap-around (m = 3), plain gradient descent (m = 0): device and
    host solvers agree, bitwise reproducible run to run, coefficients come back in the caller's feature order."""
    rng = np.random.default_rng(17)
    n, D, k = 30000, capi.FE_HEAD + 3000, 16
    rowptr, col, val, y = _zipf_shard(rng, n, D, k)
    rows = capi.DeviceFeRows(rowptr, col, val, y, None, None, D)
    opts = capi.make_opts(l2=1.0, regularize_bias=True, has_intercept=True, m=m, max_iter=max_iter)
    xd, idv = FixedEffectSolver(rows, opts, solver="device").fit()
    xd2, idv2 = FixedEffectSolver(rows, opts, solver="device").fit()
    xh, ih = FixedEffectSolver(rows, opts, solver="host").fit()
    np.testing.assert_array_equal(xd, xd2)
    assert (idv["nit"], idv["nfev"], idv["status"]) == (ih["nit"], ih["nfev"], ih["status"])
    np.testing.assert_allclose(xd, xh, rtol=1e-8, atol=1e-11)
    # the objective at the returned point, recomputed by the oracle in the CALLER's feature order
    ob = O.FeBlock(n, D, rowptr, col, val, y, np.ones(n, np.float32), np.zeros(n, np.float32))
    f_o, _ = O.fe_loss_grad(ob, O.make_opts(l2=1.0, regularize_bias=True, has_intercept=True), xd)
    assert abs(f_o - idv["f"]) <= 1e-10 * abs(f_o)


def test_device_lbfgs_stops_at_once_on_a_stationary_start():
    c = FE_CASES[0]
    solver = FixedEffectSolver(_rows(c), _opts(c, capi))
    x, info = solver.fit(FE_ARR[c["key"] + "_theta"])
    x2, info2 = FixedEffectSolver(_rows(c), _opts(c, capi), solver="host").fit(FE_ARR[c["key"] + "_theta"])
    assert (info["nit"], info["nfev"], info["status"]) == (info2["nit"], info2["nfev"], info2["status"])
    np.testing.assert_allclose(x, x2, rtol=1e-10, atol=1e-12)


def _ragged_shard(rng, n, D, max_len, long_rows=()):
    lens = rng.integers(0, max_len + 1, n)
    for i, L in long_rows:
        lens[i] = L
    rowptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    nnz = int(rowptr[-1])
    col = np.minimum((D ** rng.random(nnz) - 1).astype(np.int32), D - 1)   # skewed, duplicates inside rows
    val = rng.standard_normal(nnz).astype(np.float32)
    y = (rng.random(n) < 0.4).astype(np.float32)
    w = rng.uniform(0.5, 2.0, n).astype(np.float32)
    off = rng.standard_normal(n).astype(np.float32)
    return rowptr, col, val, y, w, off


TILE_PARAMS = [dict(), dict(hz=16, hg=32, tile_rows=64, l2_tile_rows=7000), dict(hz=1, hg=1, tile_rows=1, l2_tile_rows=100),
               dict(hz=700, hg=5, tile_rows=8192), dict(hz=3, hg=699, tile_rows=300, l2_tile_rows=1 << 22)]


@pytest.mark.parametrize("pa", TILE_PARAMS, ids=[str(i) for i in range(len(TILE_PARAMS))])
def test_tiled_objective_matches_atomic_path_and_oracle(pa):
    """gdmix_fe_loss_grad_tiled (hot / cold split of x and of the gradient, tiles of rows, reduce-by-key) against the
    single-pass atomic kernel and the oracle: skewed columns with duplicates, empty rows, rows longer than a warp's
    stage, an empty column, weights / offsets, no intercept, linear regression; every split forced by small hz / hg /
    tile sizes.  Bitwise reproducible."""
    rng = np.random.default_rng(31)
    n, D = 30000, 700
    rowptr, col, val, y, w, off = _ragged_shard(rng, n, D, 11, long_rows=[(7, 3000), (20001, 1500)])
    col[col == 5] = 6                                   # column 5 stays empty
    for hi, rb, lin in ((True, True, False), (True, False, False), (False, False, False), (True, True, True)):
        rows = capi.DeviceFeRows(rowptr, col, val, y, w, off, D, linear_regression=lin, num_workers=2)
        opts = capi.make_opts(l2=0.7, regularize_bias=rb, has_intercept=hi)
        x = torch.from_numpy(rng.standard_normal(D + (1 if hi else 0)) * 0.1).cuda()
        plan = capi.DeviceFeTilePlan(rows, **pa)
        a = capi.fe_loss_grad_device(rows, opts, x).cpu().numpy()
        b = capi.fe_loss_grad_device(rows, opts, x, plan=plan).cpu().numpy()
        c = capi.fe_loss_grad_device(rows, opts, x, plan=plan).cpu().numpy()
        np.testing.assert_array_equal(b, c)
        np.testing.assert_allclose(b, a, rtol=1e-11, atol=1e-9)
        blk = O.FeBlock(n, D, rowptr, col, val, y, w, off, linear_regression=lin, num_workers=2)
        f_o, g_o = O.fe_loss_grad(blk, O.make_opts(l2=0.7, regularize_bias=rb, has_intercept=hi), x.cpu().numpy())
        np.testing.assert_allclose(b[0], f_o, rtol=1e-12)
        np.testing.assert_allclose(b[1:], g_o, rtol=1e-10, atol=1e-10)
        plan.close()


def test_tiled_objective_edge_shapes():
    """No rows; no non-zeros; one row; a single column; every row empty but one."""
    opts = capi.make_opts(l2=0.5, regularize_bias=True, has_intercept=True)
    oo = O.make_opts(l2=0.5, regularize_bias=True, has_intercept=True)
    cases = []
    cases.append((np.zeros(1, np.int64), np.zeros(0, np.int32), np.zeros(0, np.float32), np.zeros(0, np.float32), 4))
    cases.append((np.zeros(6, np.int64), np.zeros(0, np.int32), np.zeros(0, np.float32), np.ones(5, np.float32), 4))
    cases.append((np.array([0, 3], np.int64), np.array([2, 0, 2], np.int32), np.array([1.5, -2, 0.25], np.float32),
                  np.ones(1, np.float32), 3))
    cases.append((np.arange(0, 41, dtype=np.int64), np.zeros(40, np.int32), np.linspace(-1, 1, 40).astype(np.float32),
                  (np.arange(40) % 2).astype(np.float32), 1))
    rp = np.zeros(101, np.int64); rp[51:] = 7
    cases.append((rp, np.array([0, 1, 2, 3, 2, 1, 0], np.int32), np.arange(7, dtype=np.float32),
                  (np.arange(100) % 3 == 0).astype(np.float32), 4))
    for rowptr, col, val, y, D in cases:
        n = rowptr.shape[0] - 1
        rows = capi.DeviceFeRows(rowptr, col, val, y, None, None, D)
        x = torch.from_numpy(np.linspace(-0.3, 0.4, D + 1)).cuda()
        for pa in (dict(), dict(hz=1, hg=1, tile_rows=2, l2_tile_rows=3)):
            plan = capi.DeviceFeTilePlan(rows, **pa)
            b = capi.fe_loss_grad_device(rows, opts, x, plan=plan).cpu().numpy()
            blk = O.FeBlock(n, D, rowptr, col, val, y, np.ones(n, np.float32), np.zeros(n, np.float32))
            f_o, g_o = O.fe_loss_grad(blk, oo, x.cpu().numpy())
            np.testing.assert_allclose(b[0], f_o, rtol=1e-12, atol=1e-14)
            np.testing.assert_allclose(b[1:], g_o, rtol=1e-10, atol=1e-12)


def test_tiled_objective_large_zipf_default_split():
    """The bench's shard shape at test scale with the library's own split (D above hz and hg: both cold paths live),
    through the solver's frequency ranking."""
    rng = np.random.default_rng(41)
    n, D, k = 200_000, 60_000, 32
    rowptr, col, val, y = _zipf_shard(rng, n, D, k)
    rows = capi.DeviceFeRows(rowptr, col, val, y, None, None, D)
    opts = capi.make_opts(l2=1.0, regularize_bias=True, has_intercept=True)
    solver = FixedEffectSolver(rows, opts)
    x = rng.standard_normal(D + 1) * 0.05
    f, g = solver.loss_grad(x)
    assert solver.plan.n_cold_z > 0 and solver.plan.n_cold_g > 0 and solver.plan.n_tiles > 1
    blk = O.FeBlock(n, D, rowptr, col, val, y, np.ones(n, np.float32), np.zeros(n, np.float32))
    f_o, g_o = O.fe_loss_grad(blk, O.make_opts(l2=1.0, regularize_bias=True, has_intercept=True), x)
    np.testing.assert_allclose(f, f_o, rtol=1e-12)
    np.testing.assert_allclose(g, g_o, rtol=1e-10, atol=1e-10)
    f2, g2 = solver.loss_grad(x)
    assert f2 == f and np.array_equal(g, g2)


@pytest.mark.parametrize("hi,rb", [(True, True), (True, False), (False, False)])
def test_fe_hessian_simple_and_full_match_a_dense_restatement(hi, rb):
    """gdmix_fe_hessian against the reference's accumulator H = X1^T diag(w rho (1 - rho)) X1 with the intercept column
    LAST (fixed_effect_lr_lbfgs_model.py:271-296), densified in numpy: SIMPLE = its diagonal, FULL = the matrix; then
    the variances the model derives from it (:451-463: + l2 on the diagonal except an unregularised intercept, + 1e-12,
    inverse) against numpy's inverse.  Weights, offsets; feature ids are unique inside a row (tf.sparse.to_dense
    refuses repeated indices, so the reference has no semantics for them)."""
    rng = np.random.default_rng(9)
    n, D, k = 4000, 37, 6
    rowptr = np.arange(n + 1, dtype=np.int64) * k
    col = np.stack([rng.permutation(D)[:k] for _ in range(n)]).reshape(-1).astype(np.int32)
    val = rng.standard_normal(n * k).astype(np.float32)
    y = (rng.random(n) < 0.5).astype(np.float32)
    w = rng.uniform(0.5, 2.0, n).astype(np.float32)
    off = (0.3 * rng.standard_normal(n)).astype(np.float32)
    P = D + (1 if hi else 0)
    x = rng.standard_normal(P) * 0.2
    rows = capi.DeviceFeRows(rowptr, col, val, y, w, off, D)
    opts = capi.make_opts(l2=0.8, regularize_bias=rb, has_intercept=hi)
    xd = torch.from_numpy(x).cuda()
    X1 = np.zeros((n, P))
    np.add.at(X1, (np.repeat(np.arange(n), k), col), val.astype(np.float64))     # to_dense sums duplicates
    if hi:
        X1[:, D] = 1.0
    rho = 1.0 / (1.0 + np.exp(-(X1 @ x + off)))
    H = X1.T @ (X1 * (rho * (1 - rho) * w)[:, None])
    h_simple = capi.fe_hessian_device(rows, opts, xd, capi.VARIANCE_SIMPLE).cpu().numpy()
    h_full = capi.fe_hessian_device(rows, opts, xd, capi.VARIANCE_FULL).cpu().numpy()
    np.testing.assert_allclose(h_simple, np.diag(H), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(h_full, H, rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(h_full, h_full.T, rtol=0, atol=1e-12)
    # the variances of the model file
    l2, eps = 0.8, 1e-12
    Hs = h_simple + l2
    Hf = h_full + np.diag([l2 + eps] * P)
    if hi and not rb:
        Hs[-1] -= l2
        Hf[-1, -1] -= l2
    want_full = np.diagonal(np.linalg.inv(H + np.diag([l2 + eps] * P) - (np.diag([0.0] * (P - 1) + [l2]) if hi and not rb else 0)))
    np.testing.assert_allclose(np.diagonal(np.linalg.inv(Hf)), want_full, rtol=1e-9)
    np.testing.assert_allclose(1.0 / (Hs + eps), 1.0 / (np.diag(H) + l2 * (1 - np.eye(P)[-1] * (1 if hi and not rb else 0)) + eps), rtol=1e-10)
