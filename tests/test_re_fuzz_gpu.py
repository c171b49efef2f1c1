"""Randomised differential test of the random-effect solve: batches with ragged sample counts (power-law, up to
thousands), ragged row lengths (empty rows included), a different number of local features per entity, optional
weights / offsets / intercept / bias regularisation, m in {3, 10} -- solved through the planner's cascade (typical
tier, largest-shape tier, general, global-X, clusters) and through each forced path, all against the CPU oracle:
identical iteration counts and stop status, coefficients <= 1e-5 relative (BASELINE.json north_star)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
if not torch.cuda.is_available():  # pragma: no cover
    pytest.skip("needs a CUDA device", allow_module_level=True)

from gdmix_b200 import _capi as capi  # noqa: E402
from oracle import oracle as O  # noqa: E402


def _random_batch(seed):
    rng = np.random.default_rng(seed)
    E = int(rng.integers(40, 160))
    hi = bool(rng.integers(0, 2)) or True
    has_intercept = bool(rng.random() < 0.8)
    d_max = int(rng.choice([3, 17, 64, 200, 600]))
    ent, rowptr, cols, vals, ys, ws, offs, tptr = [0], [0], [], [], [], [], [], [0]
    for e in range(E):
        n = int(np.clip(rng.pareto(1.2) * 6 + 1, 1, 3000))
        d = int(rng.integers(1, d_max + 1))
        kmax = int(min(d, rng.choice([1, 4, 12, 40])))
        lens = rng.integers(0, kmax + 1, n)
        lens[-1] = max(lens[-1], 1)                       # prepare_jobs needs the last sample non-empty (:229-232)
        for ln in lens:
            c = np.sort(rng.choice(d, ln, replace=False)).astype(np.int32)
            cols.append(c); vals.append(rng.standard_normal(ln).astype(np.float32))
            rowptr.append(rowptr[-1] + ln)
        ent.append(ent[-1] + n)
        tptr.append(tptr[-1] + d + (1 if has_intercept else 0))
        ys.append((rng.random(n) < 0.5).astype(np.float32))
        ws.append(rng.uniform(0.5, 2.0, n).astype(np.float32))
        offs.append((0.3 * rng.standard_normal(n)).astype(np.float32))
    use_w, use_off = rng.random() < 0.5, rng.random() < 0.7
    hb = capi.HostBatch(np.array(ent, np.int64), np.array(rowptr, np.int64), np.concatenate(cols), np.concatenate(vals),
                        np.concatenate(ys), np.concatenate(ws) if use_w else None,
                        np.concatenate(offs) if use_off else None, np.array(tptr, np.int64), has_intercept)
    opts = capi.make_opts(l2=float(rng.choice([0.1, 1.0, 10.0])), regularize_bias=bool(rng.integers(0, 2)),
                          has_intercept=has_intercept, m=int(rng.choice([3, 10])))
    return hb, opts


def _oracle(hb, opts, theta0=None):
    ob = {"ent_rowptr": hb.ent_rowptr, "rowptr": hb.rowptr, "col": hb.col, "val": hb.val, "y": hb.label,
          "w": hb.weight if hb.weight is not None else np.ones(hb.n_rows, np.float32),
          "off": hb.offset if hb.offset is not None else np.zeros(hb.n_rows, np.float32), "theta_ptr": hb.theta_ptr}
    oo = O.Opts(opts.l2, opts.regularize_bias, opts.has_intercept, opts.m, opts.max_iter, opts.max_ls, opts.max_fun,
                opts.factr, opts.pgtol)
    return O.re_fit_batch(ob, oo, theta0=theta0)


@pytest.mark.parametrize("seed", range(32))
def test_random_ragged_batches_match_the_oracle_on_every_path(seed, monkeypatch):
    hb, opts = _random_batch(1000 + seed)
    th_o, f_o, nit_o, nfev_o, st_o = _oracle(hb, opts)
    norm = lambda a, b: np.array([np.linalg.norm(a[s:e] - b[s:e]) / max(np.linalg.norm(b[s:e]), 1e-300)
                                  for s, e in zip(hb.theta_ptr[:-1], hb.theta_ptr[1:])])
    # An optimum at infinity (one or two samples and an unregularised intercept) is not pinned by the reference
    # itself: moving the solver's OWN start point by 1e-13 moves its answer by tens of percent (the same finding as
    # `self_sensitivity` in tests/golden/re_golden.json).  Such entities are compared on the stop status only.
    th_p = _oracle(hb, opts, theta0=1e-13 * np.random.default_rng(seed).standard_normal(th_o.shape[0]))[0]
    pinned = norm(th_p, th_o) < 1e-7
    assert pinned.mean() > 0.5
    for path in ("auto", "generic", "big", "giant", "small"):
        monkeypatch.delenv("GDMIX_RE_PATH", raising=False)
        monkeypatch.delenv("GDMIX_GIANT_ROWS", raising=False)
        if path == "giant":
            monkeypatch.setenv("GDMIX_RE_PATH", "big"); monkeypatch.setenv("GDMIX_GIANT_ROWS", "1")
        elif path != "auto":
            monkeypatch.setenv("GDMIX_RE_PATH", path)
        out = capi.re_fit_host(hb, opts)
        assert (out["status"] == st_o).all(), (path, seed)
        rel = norm(out["theta"], th_o)
        assert rel[pinned].max(initial=0.0) <= 1e-5, (path, seed, float(rel[pinned].max()))
        same = (out["nit"] == nit_o) & (out["nfev"] == nfev_o)
        assert same[pinned].mean() >= 0.99, (path, seed, float(same[pinned].mean()))
        np.testing.assert_allclose(out["f"][pinned], f_o[pinned], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("seed", range(6))
def test_random_ragged_batches_warm_start_variance_and_sweep(seed):
    """The same kind of batch through the planner's cascade with the other outputs of the path: a warm start
    (prepare_jobs builds theta0 from the previous model, job_consumers.py:262-288) retraces the oracle's warm
    start; SIMPLE variances equal the oracle's; a regularisation sweep equals separate fits bit for bit."""
    hb, opts = _random_batch(2000 + seed)
    th_o, f_o, nit_o, nfev_o, st_o = _oracle(hb, opts)
    norm = lambda a, b: np.array([np.linalg.norm(a[s:e] - b[s:e]) / max(np.linalg.norm(b[s:e]), 1e-300)
                                  for s, e in zip(hb.theta_ptr[:-1], hb.theta_ptr[1:])])
    th_p = _oracle(hb, opts, theta0=1e-13 * np.random.default_rng(seed).standard_normal(th_o.shape[0]))[0]
    pinned = norm(th_p, th_o) < 1e-7
    # warm start half way between zero and the optimum
    theta0 = 0.5 * th_o
    w_o = _oracle(hb, opts, theta0=theta0)
    w_p = _oracle(hb, opts, theta0=theta0 * (1 + 1e-13))[0]
    wpinned = pinned & (norm(w_p, w_o[0]) < 1e-7)
    w_d = capi.re_fit_host(hb, opts, theta0=theta0)
    assert (w_d["status"] == w_o[4]).all()
    assert norm(w_d["theta"], w_o[0])[wpinned].max(initial=0.0) <= 1e-5
    assert ((w_d["nit"] == w_o[2]) & (w_d["nfev"] == w_o[3]))[wpinned].mean() >= 0.99
    # SIMPLE variance at the optimum
    vopts = capi.make_opts(l2=opts.l2, regularize_bias=bool(opts.regularize_bias), has_intercept=bool(opts.has_intercept),
                           m=opts.m, variance_mode=capi.VARIANCE_SIMPLE)
    v_d = capi.re_fit_host(hb, vopts, want_variance=True)
    oo = O.Opts(opts.l2, opts.regularize_bias, opts.has_intercept, opts.m, opts.max_iter, opts.max_ls, opts.max_fun,
                opts.factr, opts.pgtol)
    w_all = hb.weight if hb.weight is not None else np.ones(hb.n_rows, np.float32)
    off_all = hb.offset if hb.offset is not None else np.zeros(hb.n_rows, np.float32)
    hi = 1 if opts.has_intercept else 0
    for e in np.flatnonzero(pinned)[::7]:
        r0, r1 = hb.ent_rowptr[e], hb.ent_rowptr[e + 1]
        q0, q1 = hb.rowptr[r0], hb.rowptr[r1]
        t0, t1 = hb.theta_ptr[e], hb.theta_ptr[e + 1]
        blk = O.EntityBlock(r1 - r0, t1 - t0 - hi, hb.rowptr[r0:r1 + 1] - q0, hb.col[q0:q1], hb.val[q0:q1],
                            hb.label[r0:r1], w_all[r0:r1], off_all[r0:r1])
        # (1 / a nearly vanishing sum of rho (1 - rho) when the entity is almost separable: 1e-6, not 1e-9)
        np.testing.assert_allclose(v_d["variance"][t0:t1], O.re_variance(blk, oo, v_d["theta"][t0:t1], "simple"),
                                   rtol=1e-6)
    # sweep == separate fits, bit for bit
    db = capi.DeviceBatch(hb)
    l2s = [0.3, 3.0, 30.0]
    sw = capi.re_fit_sweep_device(db, opts, l2s)
    for j, l2 in enumerate(l2s):
        o1 = capi.make_opts(l2=l2, regularize_bias=bool(opts.regularize_bias), has_intercept=bool(opts.has_intercept), m=opts.m)
        one = capi.re_fit_device(db, o1)
        torch.cuda.synchronize()
        for key in ("theta", "f", "nit", "nfev", "status"):
            assert torch.equal(sw[key][j], one[key]), (key, l2)
