"""BASELINE.json configs[3] at test scale: fixed effect -> per-user random effect -> per-item random effect, every
stage's scores feeding the next stage's offsets (the GDMix coordinate-descent pass) -- entirely on
the device: FE solve, fp32 scores, group-by-user (radix sort) + local indexing, batched RE solve, scores, regroup by
item, RE solve.  Checked against the same chain run through the CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
if not torch.cuda.is_available():  # pragma: no cover
    pytest.skip("needs a CUDA device", allow_module_level=True)

from gdmix_b200 import _capi as capi, partition as P  # noqa: E402
from gdmix_b200.fe_solver import FixedEffectSolver  # noqa: E402
from oracle import oracle as O  # noqa: E402


def _sparse_rows(rng, n, D, k):
    col = np.sort(np.stack([rng.choice(D, k, replace=False) for _ in range(n)]), axis=1).astype(np.int32)
    val = rng.standard_normal((n, k)).astype(np.float32)
    return np.arange(n + 1, dtype=np.int64) * k, col.reshape(-1), val.reshape(-1)


def _oracle_re_stage(keys, rowptr, gcol, val, y, off, l2):
    """Group on the host (stable), local-index per entity, oracle fit, fp32 scores back in row order."""
    order = np.argsort(keys, kind="stable")
    uniq, counts = np.unique(keys, return_counts=True)
    starts = np.concatenate([[0], np.cumsum(counts)])
    scores = np.zeros(len(keys), np.float32)
    thetas = {}
    oo = O.make_opts(l2=l2, regularize_bias=False)
    for e, key in enumerate(uniq):
        rows = order[starts[e]:starts[e + 1]]
        sc = [gcol[rowptr[r]:rowptr[r + 1]] for r in rows]
        sv = [val[rowptr[r]:rowptr[r + 1]] for r in rows]
        blk, ug = O.build_local_block(sc, sv, y[rows], None, off[rows])
        th = O.re_fit(blk, oo)[0]
        thetas[int(key)] = (np.asarray(th), np.asarray(ug))
        logit, _ = O.re_score(blk, oo, th)
        scores[rows] = logit.astype(np.float32)
    return thetas, scores


def _device_re_stage(keys, rowptr, gcol, val, y, off, l2, D):
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    gb = P.GroupedBatch(P.regroup_batch(t(keys), t(rowptr), t(gcol), t(val), t(y), t(off), None, num_features=D))
    opts = capi.make_opts(l2=l2, regularize_bias=False)
    fit = capi.re_fit_device(gb, opts)
    logit, _ = capi.re_score_device(gb, opts, fit["theta"])
    torch.cuda.synchronize()
    assert (fit["status"].cpu().numpy() == 0).all()
    d = gb.d
    th, tp = fit["theta"].cpu().numpy(), d["theta_ptr"].cpu().numpy()
    up, ug, ids = d["uniq_ptr"].cpu().numpy(), d["uniq_global"].cpu().numpy(), d["entity_ids"].cpu().numpy()
    thetas = {int(ids[e]): (th[tp[e]:tp[e + 1]], ug[up[e]:up[e + 1]]) for e in range(len(ids))}
    return thetas, gb.scatter_to_input_order(logit).cpu().numpy()


def test_fixed_effect_then_user_then_item_random_effects():
    rng = np.random.default_rng(11)
    n, U, I, D, Du, Di = 6000, 150, 40, 300, 60, 50
    user = rng.integers(0, U, n).astype(np.int64) * 7 + 3
    item = (rng.zipf(1.5, n) % I).astype(np.int64)
    g_rp, g_col, g_val = _sparse_rows(rng, n, D, 8)          # global bag
    u_rp, u_col, u_val = _sparse_rows(rng, n, Du, 4)         # per-user bag (item features)
    i_rp, i_col, i_val = _sparse_rows(rng, n, Di, 4)         # per-item bag (user features)
    y = (rng.random(n) < 0.45).astype(np.float32)

    # ---- stage 1: fixed effect
    opts = capi.make_opts(l2=1.0, regularize_bias=True)
    rows = capi.DeviceFeRows(g_rp, g_col, g_val, y, None, None, D)
    solver = FixedEffectSolver(rows, opts, D)
    x, info = solver.fit()
    s0, _ = solver.score(x)
    x_o, f_o, nit_o, nfev_o, st_o = O.fe_fit(O.FeBlock(n, D, g_rp, g_col, g_val, y), O.make_opts(l2=1.0, regularize_bias=True))
    assert info["nit"] == nit_o
    np.testing.assert_allclose(x, x_o, rtol=1e-7, atol=1e-9)
    s0_o = np.array([g_val[g_rp[r]:g_rp[r + 1]].astype(np.float64) @ x_o[g_col[g_rp[r]:g_rp[r + 1]]] for r in range(n)])
    s0_o = (s0_o + x_o[-1]).astype(np.float32)
    np.testing.assert_allclose(s0, s0_o, rtol=2e-6, atol=2e-6)

    # ---- stage 2: per-user random effect on offset = FE score (identical fp32 offsets into both chains)
    th_u, s1 = _device_re_stage(user, u_rp, u_col, u_val, y, s0_o, 1.0, Du)
    th_u_o, s1_o = _oracle_re_stage(user, u_rp, u_col, u_val, y, s0_o, 1.0)
    assert set(th_u) == set(th_u_o)
    for k in th_u_o:
        np.testing.assert_array_equal(th_u[k][1], th_u_o[k][1])
        assert np.linalg.norm(th_u[k][0] - th_u_o[k][0]) <= 1e-5 * max(np.linalg.norm(th_u_o[k][0]), 1e-12)
    np.testing.assert_allclose(s1, s1_o, rtol=2e-6, atol=2e-6)

    # ---- stage 3: per-item random effect on offset = FE + per-user score
    th_i, s2 = _device_re_stage(item, i_rp, i_col, i_val, y, s1_o, 1.0, Di)
    th_i_o, s2_o = _oracle_re_stage(item, i_rp, i_col, i_val, y, s1_o, 1.0)
    assert set(th_i) == set(th_i_o)
    for k in th_i_o:
        assert np.linalg.norm(th_i[k][0] - th_i_o[k][0]) <= 1e-5 * max(np.linalg.norm(th_i_o[k][0]), 1e-12)
    np.testing.assert_allclose(s2, s2_o, rtol=2e-6, atol=2e-6)
    # the chain improves the fit: AUC rises stage by stage (device evaluator)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    aucs = [P.auc(t(s), t(y)) for s in (s0, s1, s2)]
    assert aucs[0] < aucs[1] < aucs[2]
