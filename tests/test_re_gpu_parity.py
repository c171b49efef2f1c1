"""GPU parity: the CUDA random-effect path (through the C ABI) against the CPU oracle and against the
golden vectors the reference itself produced.  Tolerance per BASELINE.json north_star: coefficients within
1e-5 relative of the reference solver (we assert much tighter where the reference pins its own answer)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():  # pragma: no cover
    pytest.skip("needs a CUDA device", allow_module_level=True)

from gdmix_b200 import _capi as capi  # noqa: E402
from gdmix_b200.synthetic import make_batch  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tests.golden_util import is_pinned, load_re  # noqa: E402

REL_TOL = 1e-5  # north_star tolerance


@pytest.fixture(autouse=True, params=["auto", "generic", "big", "giant", "small"])
def re_path(request, monkeypatch):
    """Every test runs five times: through the planner's choice (the warp-per-entity kernel for batches of small
    entities, the sliced-ELL fast kernel when the batch qualifies), through the general staged kernel alone, through
    the kernel that leaves X in global memory (the one entities too large for the chip take), through that kernel
    launched as thread-block clusters (eight CTAs share an entity's samples -- what entities with tens of thousands of
    samples take), and with the warp-per-entity kernel forced in front of the cascade (entities above its slices
    defer to the rest)."""
    monkeypatch.delenv("GDMIX_GIANT_ROWS", raising=False)
    if request.param == "giant":
        monkeypatch.setenv("GDMIX_RE_PATH", "big")
        monkeypatch.setenv("GDMIX_GIANT_ROWS", "1")
    elif request.param != "auto":
        monkeypatch.setenv("GDMIX_RE_PATH", request.param)
    else:
        monkeypatch.delenv("GDMIX_RE_PATH", raising=False)
    return request.param


def _oracle_opts(o):
    return O.Opts(o.l2, o.regularize_bias, o.has_intercept, o.m, o.max_iter, o.max_ls, o.max_fun, o.factr, o.pgtol)


def _oracle_batch(hb):
    return {"ent_rowptr": hb.ent_rowptr, "rowptr": hb.rowptr, "col": hb.col, "val": hb.val, "y": hb.label,
            "w": hb.weight if hb.weight is not None else np.ones(hb.n_rows, np.float32),
            "off": hb.offset if hb.offset is not None else np.zeros(hb.n_rows, np.float32),
            "theta_ptr": hb.theta_ptr}


def _rel_per_entity(a, b, ptr):
    out = np.zeros(len(ptr) - 1)
    for e in range(len(ptr) - 1):
        x, y = a[ptr[e]:ptr[e + 1]], b[ptr[e]:ptr[e + 1]]
        out[e] = np.linalg.norm(x - y) / max(np.linalg.norm(y), 1e-300)
    return out


def _golden_batch(cases, arr):
    """Packs golden cases that share solver options into one HostBatch."""
    ent, rowptr, cols, vals, ys, ws, offs, tptr = [0], [0], [], [], [], [], [], [0]
    for c in cases:
        k = c["key"]
        rp = arr[k + "_rowptr"]
        rowptr.extend((rp[1:] + rowptr[-1]).tolist())
        ent.append(ent[-1] + c["n"])
        cols.append(arr[k + "_col"]); vals.append(arr[k + "_val"]); ys.append(arr[k + "_y"])
        ws.append(arr[k + "_w"]); offs.append(arr[k + "_off"])
        tptr.append(tptr[-1] + c["d"] + (1 if c["has_intercept"] else 0))
    return capi.HostBatch(np.array(ent), np.array(rowptr), np.concatenate(cols), np.concatenate(vals),
                          np.concatenate(ys), np.concatenate(ws), np.concatenate(offs), np.array(tptr))


def _group_key(c):
    return (c["l2"], c["regularize_bias"], c["has_intercept"], c["m"], c["max_iter"], c["tol"])


ARR, CASES = load_re()
GROUPS = {}
for _c in CASES:
    GROUPS.setdefault(_group_key(_c), []).append(_c)


@pytest.mark.parametrize("threads", [0, 32, 64, 128, 256])
def test_golden_fit_matches_reference(threads):
    """Every golden case (reference fixtures + synthetic + edge cases + warm starts) through gdmix_re_fit_host."""
    worst = 0.0
    for key, cases in GROUPS.items():
        l2, rb, hi, m, maxit, tol = key
        hb = _golden_batch(cases, ARR)
        theta0 = np.concatenate([ARR[c["key"] + "_theta0"] if c["warm"] else
                                 np.zeros(c["d"] + (1 if hi else 0)) for c in cases])
        opts = capi.make_opts(l2=l2, regularize_bias=rb, has_intercept=hi, m=m, max_iter=maxit, tol=tol,
                              threads_per_entity=threads)
        out = capi.re_fit_host(hb, opts, theta0=theta0)
        for i, c in enumerate(cases):
            th = out["theta"][hb.theta_ptr[i]:hb.theta_ptr[i + 1]]
            ref = ARR[c["key"] + "_theta"]
            if is_pinned(c):
                rel = np.linalg.norm(th - ref) / max(np.linalg.norm(ref), 1e-300)
                worst = max(worst, rel)
                assert rel <= 1e-7, (c["name"], rel)
                assert (out["nit"][i], out["nfev"][i], out["status"][i]) == (c["nit"], c["nfev"], c["warnflag"]), c["name"]
                assert abs(out["f"][i] - c["f"]) <= 1e-11 * max(1.0, abs(c["f"])), c["name"]
            else:
                assert out["status"][i] == 0 and abs(out["f"][i] - c["f"]) <= 1e-5, c["name"]
    print("worst relative coefficient error vs reference:", worst)


def test_golden_loss_grad_matches_reference():
    """K1 seam: f and g at the reference's probe points (binary_logistic_regression.py:84-131)."""
    for key, cases in GROUPS.items():
        l2, rb, hi, m, maxit, tol = key
        hb = _golden_batch(cases, ARR)
        probe = np.concatenate([ARR[c["key"] + "_probe"] for c in cases])
        opts = capi.make_opts(l2=l2, regularize_bias=rb, has_intercept=hi, m=m, max_iter=maxit, tol=tol)
        db = capi.DeviceBatch(hb)
        f, g = capi.re_loss_grad_device(db, opts, torch.from_numpy(probe).cuda())
        torch.cuda.synchronize()
        f, g = f.cpu().numpy(), g.cpu().numpy()
        for i, c in enumerate(cases):
            assert abs(f[i] - c["probe_f"]) <= 1e-13 * max(1.0, abs(c["probe_f"])), c["name"]
            np.testing.assert_allclose(g[hb.theta_ptr[i]:hb.theta_ptr[i + 1]], ARR[c["key"] + "_probe_g"],
                                       rtol=1e-11, atol=1e-13, err_msg=c["name"])


@pytest.mark.parametrize("shape", [(128, 256, 32, False), (32, 64, 8, False), (64, 64, 16, True),
                                   (106, 20, 6, True), (16, 24, 4, False)])
def test_synthetic_batch_matches_oracle(shape):
    """Seeded synthetic entities (C1 / C3 / C4 / MovieLens-like shapes): CUDA vs oracle on the same bytes."""
    n, d, k, ragged = shape
    E = 400
    hb = make_batch(E, n, d, k, seed=7 + n, ragged=ragged, weights=ragged)
    opts = capi.make_opts(l2=1.0, regularize_bias=False, has_intercept=True)
    out = capi.re_fit_host(hb, opts, chunk_entities=150)  # exercises the chunked pipeline too
    th_o, f_o, nit_o, nfev_o, st_o = O.re_fit_batch(_oracle_batch(hb), _oracle_opts(opts))
    rel = _rel_per_entity(out["theta"], th_o, hb.theta_ptr)
    frac = float((rel <= REL_TOL).mean())
    print(f"shape {shape}: max rel {rel.max():.3e}, frac<=1e-5 {frac:.4f}, nit equal {(out['nit'] == nit_o).mean():.4f}")
    assert rel.max() <= REL_TOL
    assert (out["nit"] == nit_o).all() and (out["nfev"] == nfev_o).all() and (out["status"] == st_o).all()
    np.testing.assert_allclose(out["f"], f_o, rtol=1e-11)


def test_device_api_matches_host_api(re_path):
    hb = make_batch(300, 64, 64, 16, seed=3)
    opts = capi.make_opts(l2=0.5)
    host = capi.re_fit_host(hb, opts)
    assert capi.last_plan()["fast"] == (1 if re_path in ("auto", "small") else 0)
    assert ("small" in capi.last_plan()) == (re_path == "small")
    dev = capi.re_fit_device(capi.DeviceBatch(hb), opts)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(dev["theta"].cpu().numpy(), host["theta"])
    np.testing.assert_array_equal(dev["nit"].cpu().numpy(), host["nit"])


def test_run_to_run_bitwise_reproducible():
    hb = make_batch(500, 128, 256, 32, seed=11)
    opts = capi.make_opts()
    a = capi.re_fit_host(hb, opts)
    b = capi.re_fit_host(hb, opts, chunk_entities=97)
    np.testing.assert_array_equal(a["theta"], b["theta"])
    np.testing.assert_array_equal(a["f"], b["f"])


def test_threshold_and_variance_simple(re_path):
    """threshold_coefficients (model_utils.py:4-12) and SIMPLE variance (binary_logistic_regression.py:171-177)."""
    if re_path == "small":
        pytest.skip("compares bit for bit runs that the warp-per-entity tier only takes in part (no variance / sweep there)")
    hb = make_batch(64, 40, 24, 6, seed=5, weights=True)
    raw = capi.re_fit_host(hb, capi.make_opts(l2=1.0))
    thr = capi.re_fit_host(hb, capi.make_opts(l2=1.0, sparsity_threshold=1e-4, variance_mode=capi.VARIANCE_SIMPLE),
                           want_variance=True)
    expect = np.where(np.abs(raw["theta"]) <= 1e-4, 0.0, raw["theta"])
    np.testing.assert_array_equal(thr["theta"], expect)
    ob = _oracle_batch(hb)
    oo = _oracle_opts(capi.make_opts(l2=1.0))
    for e in range(0, 64, 7):
        r0, r1 = hb.ent_rowptr[e], hb.ent_rowptr[e + 1]
        q0, q1 = hb.rowptr[r0], hb.rowptr[r1]
        blk = O.EntityBlock(r1 - r0, 24, hb.rowptr[r0:r1 + 1] - q0, hb.col[q0:q1], hb.val[q0:q1], hb.label[r0:r1],
                            ob["w"][r0:r1], ob["off"][r0:r1])
        th = raw["theta"][hb.theta_ptr[e]:hb.theta_ptr[e + 1]]
        np.testing.assert_allclose(thr["variance"][hb.theta_ptr[e]:hb.theta_ptr[e + 1]],
                                   O.re_variance(blk, oo, th, "simple"), rtol=1e-10)


def test_scoring_matches_oracle():
    """InferenceJobConsumer (job_consumers.py:138-152): logits incl. offset, per-coordinate, entities w/o model."""
    hb = make_batch(200, 32, 64, 8, seed=9)
    opts = capi.make_opts()
    fit = capi.re_fit_host(hb, opts)
    has_model = (np.arange(200) % 5 != 0).astype(np.uint8)
    logit, per = capi.re_score_host(hb, opts, fit["theta"], has_model)
    ob, oo = _oracle_batch(hb), _oracle_opts(opts)
    for e in range(0, 200, 9):
        r0, r1 = hb.ent_rowptr[e], hb.ent_rowptr[e + 1]
        q0, q1 = hb.rowptr[r0], hb.rowptr[r1]
        blk = O.EntityBlock(r1 - r0, 64, hb.rowptr[r0:r1 + 1] - q0, hb.col[q0:q1], hb.val[q0:q1], hb.label[r0:r1],
                            ob["w"][r0:r1], ob["off"][r0:r1])
        th = fit["theta"][hb.theta_ptr[e]:hb.theta_ptr[e + 1]] if has_model[e] else None
        lo, po = O.re_score(blk, oo, th)
        np.testing.assert_allclose(logit[r0:r1], lo.astype(np.float32), rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(per[r0:r1], po.astype(np.float32), rtol=1e-6, atol=1e-7)


def test_scoring_in_chunks_is_the_same(monkeypatch):
    """gdmix_re_score_host cuts the partition into bounded chunks of entities: any chunking gives the same bits."""
    hb = make_batch(300, 24, 48, 6, seed=21)
    opts = capi.make_opts()
    fit = capi.re_fit_host(hb, opts)
    has_model = (np.arange(300) % 7 != 0).astype(np.uint8)
    whole = capi.re_score_host(hb, opts, fit["theta"], has_model)
    for budget in ("1", "20000", "150000"):     # one entity per chunk, a few entities, a few chunks
        monkeypatch.setenv("GDMIX_SCORE_CHUNK_BYTES", budget)
        part = capi.re_score_host(hb, opts, fit["theta"], has_model)
        np.testing.assert_array_equal(part[0], whole[0])
        np.testing.assert_array_equal(part[1], whole[1])


def test_logistic_terms_accuracy():
    """The fast kernel's own exp(-|z|), log(1 + t), 1 / (1 + t) (re_common.cuh: logistic_terms) against the CUDA
    math library: within 2 ulp for |z| <= 708 (and against numpy in higher precision terms: 4e-16 relative)."""
    import ctypes as C
    rng = np.random.default_rng(3)
    z = np.concatenate([rng.uniform(-40, 40, 200000), rng.uniform(-708, 708, 100000), rng.standard_normal(200000) * 1e-3,
                        rng.standard_normal(100000) * 1e-9, np.array([0.0, -0.0, 1e-300, 708.0, -708.0, 0.34657359, 0.8813736]),
                        np.linspace(-2, 2, 100001)])
    zd = torch.from_numpy(z).cuda()
    out = torch.empty(6 * z.size, dtype=torch.float64, device="cuda")
    capi.check(capi.lib.gdmix_selftest_logistic(C.c_void_p(zd.data_ptr()), C.c_int64(z.size), C.c_void_p(out.data_ptr()),
                                                C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    o = out.cpu().numpy().reshape(-1, 6)
    for k, name in ((0, "exp(-|z|)"), (2, "1/(1+t)")):
        mine, lib = o[:, k], o[:, 3 + k]
        ulp = np.abs(mine - lib) / np.spacing(np.abs(lib))
        assert ulp.max() <= 2.0, (name, ulp.max(), z[ulp.argmax()])
    # log(1 + t) goes through u = fl(1 + t) in the reference's formula too: one ulp of t can move u by one ulp of 1,
    # i.e. the value by 2.2e-16 whatever its size -- so: the log of MY u within 2 ulp, and the library's within that + eps
    u = 1.0 + o[:, 0]
    ref = np.log(u.astype(np.longdouble))
    err = np.abs(o[:, 1] - ref)
    assert np.all(err <= 2.0 * np.spacing(np.abs(o[:, 1])) + 1e-300), (err / np.spacing(np.abs(o[:, 1]))).max()
    assert np.max(np.abs(o[:, 1] - o[:, 4])) <= 4.5e-16
    # beyond the clamp: t stays a tiny positive number, the other two terms are exact
    zd2 = torch.tensor([750.0, -1e6, 1e300], dtype=torch.float64, device="cuda")
    out2 = torch.empty(18, dtype=torch.float64, device="cuda")
    capi.check(capi.lib.gdmix_selftest_logistic(C.c_void_p(zd2.data_ptr()), C.c_int64(3), C.c_void_p(out2.data_ptr()),
                                                C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    o2 = out2.cpu().numpy().reshape(-1, 6)
    assert (o2[:, 0] > 0).all() and (o2[:, 0] < 1e-307).all() and (o2[:, 2] == 1.0).all() and (o2[:, 1] < 1e-307).all()


def test_narrow_row_lengths_and_label_bits_are_bit_identical(monkeypatch):
    """gdmix_re_batch.row_len16 / label_bits (2 bytes and 1 bit per row across PCIe instead of 8 + 4 bytes): the row
    pointers rebuilt on the device by a scan and the labels expanded from bits give the fits the int64 / fp32 arrays
    give, bit for bit -- with ragged rows, empty rows and chunk boundaries that do not fall on a multiple of 8 rows."""
    import ctypes as C
    hb = make_batch(157, 19, 40, 5, seed=13, ragged=True)
    opts = capi.make_opts()
    E, T = hb.n_entities, hb.n_coef

    def run(cb, chunk):
        theta = np.zeros(T); nit = np.zeros(E, np.int32)
        capi.check(capi.lib.gdmix_re_fit_host(C.byref(cb), C.byref(opts), None, theta.ctypes.data_as(C.c_void_p), None,
                                              nit.ctypes.data_as(C.c_void_p), None, None, None, C.c_int64(chunk)))
        return theta, nit
    wide = hb.c_struct(narrow=False)
    assert wide.row_len16 is None and wide.label_bits is None
    assert hb.c_struct(narrow=True).row_len16 is None
    narrow = hb.c_struct(narrow=True, narrow_rows=True)
    assert narrow.row_len16 is not None and narrow.label_bits is not None
    only_bits = hb.c_struct(narrow=True, narrow_rows=True)      # label bits alone (label pointer NULL)
    only_bits.label = None
    for chunk in (0, 37, 1):                  # (the launch plan follows a chunk's shapes: compare like with like)
        ref = run(wide, chunk)
        for cb in (narrow, only_bits):
            got = run(cb, chunk)
            np.testing.assert_array_equal(got[0], ref[0])
            np.testing.assert_array_equal(got[1], ref[1])


def test_bad_column_index_is_rejected():
    hb = make_batch(8, 16, 24, 4, seed=1)
    hb.col[5] = 9999
    with pytest.raises(capi.GdmixError):
        capi.re_fit_host(hb, capi.make_opts())


@pytest.mark.parametrize("m", [1, 4, 15, 32])
def test_history_sizes_match_oracle(m):
    """Curvature-pair counts other than the default 10 (m > 10 runs the MT=32 kernel instantiation)."""
    hb = make_batch(120, 48, 40, 8, seed=21 + m, weights=True)
    opts = capi.make_opts(l2=0.3, m=m)
    out = capi.re_fit_host(hb, opts)
    th_o, f_o, nit_o, nfev_o, st_o = O.re_fit_batch(_oracle_batch(hb), _oracle_opts(opts))
    rel = _rel_per_entity(out["theta"], th_o, hb.theta_ptr)
    assert rel.max() <= REL_TOL, rel.max()
    assert (out["nit"] == nit_o).all() and (out["nfev"] == nfev_o).all()


def test_deferred_entities_take_the_general_kernel(monkeypatch, re_path):
    """Entities whose sliced form does not fit the fast kernel's shared memory are drained by the general
    kernel from the deferral list: same answers either way."""
    if re_path != "auto":
        pytest.skip("fast path only")
    hb = make_batch(300, 64, 64, 16, seed=31, ragged=True)
    opts = capi.make_opts(l2=1.0)
    ref = capi.re_fit_host(hb, opts)
    assert capi.last_plan()["fast"] == 1
    db = capi.DeviceBatch(hb)
    monkeypatch.setenv("GDMIX_FAST_CAP_STEPS", "20")   # mean entity needs ~24 steps: most are deferred
    dev = capi.re_fit_device(db, opts)
    torch.cuda.synchronize()
    deferred = int(dev["workspace"][:12].view(torch.int32)[2].item())
    print("deferred", deferred, "of", hb.n_entities)
    assert 0 < deferred < hb.n_entities
    mixed = capi.re_fit_host(hb, opts)
    monkeypatch.setenv("GDMIX_FAST_CAP_STEPS", "1")    # everything deferred
    alld = capi.re_fit_host(hb, opts)
    monkeypatch.setenv("GDMIX_RE_PATH", "generic")
    gen = capi.re_fit_host(hb, opts)
    np.testing.assert_array_equal(alld["theta"], gen["theta"])
    for k in ("nit", "nfev", "status"):
        np.testing.assert_array_equal(mixed[k], ref[k])
        np.testing.assert_array_equal(alld[k], ref[k])
    rel = _rel_per_entity(mixed["theta"], ref["theta"], hb.theta_ptr)
    assert rel.max() <= 1e-9, rel.max()


@pytest.mark.parametrize("shape", [(40, 24, 6), (128, 256, 32)])
def test_variance_full_matches_oracle(shape, re_path):
    """FULL variance = diag((X1^T D X1 + (l2 + 1e-12) I - l2 e0 e0^T)^-1) at the un-thresholded optimum
    (binary_logistic_regression.py:178-186); the matrix lives on chip for the small shape, in the workspace for C1."""
    if re_path == "small":
        pytest.skip("compares bit for bit runs that the warp-per-entity tier only takes in part (no variance / sweep there)")
    n, d, k = shape
    E = 48
    hb = make_batch(E, n, d, k, seed=15, weights=True)
    raw = capi.re_fit_host(hb, capi.make_opts(l2=1.0))
    out = capi.re_fit_host(hb, capi.make_opts(l2=1.0, sparsity_threshold=1e-4, variance_mode=capi.VARIANCE_FULL),
                           want_variance=True)
    np.testing.assert_array_equal(out["theta"], np.where(np.abs(raw["theta"]) <= 1e-4, 0.0, raw["theta"]))
    ob = _oracle_batch(hb)
    oo = _oracle_opts(capi.make_opts(l2=1.0))
    for e in range(0, E, 5):
        r0, r1 = hb.ent_rowptr[e], hb.ent_rowptr[e + 1]
        q0, q1 = hb.rowptr[r0], hb.rowptr[r1]
        blk = O.EntityBlock(r1 - r0, d, hb.rowptr[r0:r1 + 1] - q0, hb.col[q0:q1], hb.val[q0:q1], hb.label[r0:r1],
                            ob["w"][r0:r1], ob["off"][r0:r1])
        th = raw["theta"][hb.theta_ptr[e]:hb.theta_ptr[e + 1]]
        np.testing.assert_allclose(out["variance"][hb.theta_ptr[e]:hb.theta_ptr[e + 1]],
                                   O.re_variance(blk, oo, th, "full"), rtol=1e-9)


def test_golden_variances_match_reference():
    """SIMPLE and FULL variances the reference itself computed for the golden cases (re_golden: *_var_simple,
    *_var_full), cold-start cases with l2 > 0."""
    checked = 0
    for key, cases in GROUPS.items():
        l2, rb, hi, m, maxit, tol = key
        cases = [c for c in cases if not c["warm"] and is_pinned(c) and (c["key"] + "_var_full") in ARR]
        if l2 <= 0 or not cases:
            continue
        hb = _golden_batch(cases, ARR)
        for mode, suffix in ((capi.VARIANCE_SIMPLE, "_var_simple"), (capi.VARIANCE_FULL, "_var_full")):
            opts = capi.make_opts(l2=l2, regularize_bias=rb, has_intercept=hi, m=m, max_iter=maxit, tol=tol,
                                  variance_mode=mode)
            out = capi.re_fit_host(hb, opts, want_variance=True)
            for i, c in enumerate(cases):
                np.testing.assert_allclose(out["variance"][hb.theta_ptr[i]:hb.theta_ptr[i + 1]],
                                           ARR[c["key"] + suffix], rtol=2e-6, err_msg=c["name"] + suffix)
                checked += 1
    assert checked > 50


def test_oversized_entities_are_solved_not_rejected(re_path):
    """One entity with 200k non-zeros (more than the chip holds) and one with 70 000 rows (more than the 16-bit
    on-chip row index) in a batch of ordinary ones: nobody is rejected, the ordinary ones keep their kernel."""
    if re_path != "auto":
        pytest.skip("planner's choice only")
    small = make_batch(60, 48, 40, 8, seed=41, weights=True)
    rng = np.random.default_rng(9)

    def entity(n, d, k):
        col = np.sort(np.stack([rng.choice(d, k, replace=False) for _ in range(n)]), axis=1).astype(np.int32)
        val = rng.standard_normal((n, k)).astype(np.float32)
        th = rng.standard_normal(d + 1) * 0.3
        z = (val * th[1 + col]).sum(1) + th[0]
        y = (rng.random(n) < 1 / (1 + np.exp(-z))).astype(np.float32)
        return col.reshape(-1), val.reshape(-1), y, n, k

    parts = [entity(3200, 40, 40 - 1), entity(70000, 40, 3)]
    ent = small.ent_rowptr.tolist()
    rowptr = small.rowptr.tolist()
    cols, vals, ys = [small.col], [small.val], [small.label]
    ws, offs = [small.weight], [small.offset]
    tptr = small.theta_ptr.tolist()
    for col, val, y, n, k in parts:
        rowptr.extend((rowptr[-1] + k * np.arange(1, n + 1)).tolist())
        ent.append(ent[-1] + n)
        cols.append(col); vals.append(val); ys.append(y)
        ws.append(np.ones(n, np.float32)); offs.append(np.zeros(n, np.float32))
        tptr.append(tptr[-1] + 41)
    hb = capi.HostBatch(np.array(ent), np.array(rowptr), np.concatenate(cols), np.concatenate(vals),
                        np.concatenate(ys), np.concatenate(ws), np.concatenate(offs), np.array(tptr))
    opts = capi.make_opts(l2=1.0, variance_mode=capi.VARIANCE_SIMPLE)
    out = capi.re_fit_host(hb, opts, want_variance=True)
    assert capi.last_plan()["fast"] == 1
    assert (out["status"] >= 0).all()
    th_o, f_o, nit_o, nfev_o, st_o = O.re_fit_batch(_oracle_batch(hb), _oracle_opts(opts))
    rel = _rel_per_entity(out["theta"], th_o, hb.theta_ptr)
    assert rel.max() <= REL_TOL, rel
    np.testing.assert_array_equal(out["nit"][:60], nit_o[:60])
    assert abs(int(out["nit"][60]) - int(nit_o[60])) <= 1 and abs(int(out["nit"][61]) - int(nit_o[61])) <= 1
    ob, oo = _oracle_batch(hb), _oracle_opts(opts)
    for e in (60, 61):
        r0, r1 = hb.ent_rowptr[e], hb.ent_rowptr[e + 1]
        q0, q1 = hb.rowptr[r0], hb.rowptr[r1]
        blk = O.EntityBlock(r1 - r0, 40, hb.rowptr[r0:r1 + 1] - q0, hb.col[q0:q1], hb.val[q0:q1], hb.label[r0:r1],
                            ob["w"][r0:r1], ob["off"][r0:r1])
        np.testing.assert_allclose(out["variance"][hb.theta_ptr[e]:hb.theta_ptr[e + 1]],
                                   O.re_variance(blk, oo, th_o[hb.theta_ptr[e]:hb.theta_ptr[e + 1]], "simple"), rtol=1e-6)


@pytest.mark.parametrize("shape", [(300, 64, 64, 16), (120, 128, 256, 32)])
def test_l2_sweep_equals_separate_fits(shape, re_path):
    """BASELINE.json configs[4]: a sweep over l2_reg_weight solved from ONE staged copy of every entity block.
    Model j must be bit-for-bit what gdmix_re_fit returns with l2 = l2_values[j] (in the reference a sweep is
    n separate training jobs), and match the oracle run with that weight."""
    if re_path == "small":
        pytest.skip("compares bit for bit runs that the warp-per-entity tier only takes in part (no variance / sweep there)")
    E, n, d, k = shape
    hb = make_batch(E, n, d, k, seed=77, weights=True)
    db = capi.DeviceBatch(hb)
    l2s = [0.1, 1.0, 10.0, 100.0]
    sw = capi.re_fit_sweep_device(db, capi.make_opts(l2=123.0), l2s)
    torch.cuda.synchronize()
    for j, l2 in enumerate(l2s):
        opts = capi.make_opts(l2=l2)
        one = capi.re_fit_device(db, opts)
        torch.cuda.synchronize()
        for key in ("theta", "f", "nit", "nfev", "status"):
            assert torch.equal(sw[key][j], one[key]), (key, l2)
        th_o, f_o, nit_o, nfev_o, st_o = O.re_fit_batch(_oracle_batch(hb), _oracle_opts(opts), e0=0, e1=60)
        tp = hb.theta_ptr
        rel = _rel_per_entity(sw["theta"][j].cpu().numpy()[:tp[60]], th_o[:tp[60]], tp[:61])
        assert rel.max() <= REL_TOL, (l2, rel.max())
        assert (sw["nit"][j].cpu().numpy()[:60] == nit_o[:60]).all()


def test_l2_sweep_with_deferred_entities(monkeypatch, re_path):
    """Entities the fast kernel defers are swept by the general kernel, one launch per weight, same results."""
    if re_path == "small":
        pytest.skip("compares bit for bit runs that the warp-per-entity tier only takes in part (no variance / sweep there)")
    if re_path == "auto":
        monkeypatch.setenv("GDMIX_FAST_CAP_STEPS", "30")
    hb = make_batch(200, 64, 64, 16, seed=78, ragged=True)
    db = capi.DeviceBatch(hb)
    l2s = [0.5, 5.0, 50.0]
    sw = capi.re_fit_sweep_device(db, capi.make_opts(), l2s)
    for j, l2 in enumerate(l2s):
        one = capi.re_fit_device(db, capi.make_opts(l2=l2))
        torch.cuda.synchronize()
        for key in ("theta", "f", "nit", "nfev", "status"):
            assert torch.equal(sw[key][j], one[key]), (key, l2)


def test_l2_sweep_rejects_bad_arguments():
    hb = make_batch(8, 16, 24, 4, seed=1)
    db = capi.DeviceBatch(hb)
    with pytest.raises(capi.GdmixError):
        capi.re_fit_sweep_device(db, capi.make_opts(), [1.0] * 17)
    with pytest.raises(capi.GdmixError):
        capi.re_fit_sweep_device(db, capi.make_opts(), [1.0, -2.0])


def test_ragged_batch_takes_two_fast_tiers(monkeypatch, re_path):
    """A batch whose largest entity is far above the typical one is planned twice: a first fast launch for
    2.5 x the mean sample count at high residency, whose deferrals the launch planned for the largest entity
    drains.  Same answers as the single-tier plan and as the oracle."""
    if re_path != "auto":
        pytest.skip("fast path only")
    small, large = make_batch(380, 40, 64, 16, seed=91, ragged=True), make_batch(20, 600, 64, 16, seed=92)
    hb = capi.HostBatch(np.concatenate([small.ent_rowptr, large.ent_rowptr[1:] + small.ent_rowptr[-1]]),
                        np.concatenate([small.rowptr, large.rowptr[1:] + small.rowptr[-1]]),
                        np.concatenate([small.col, large.col]), np.concatenate([small.val, large.val]),
                        np.concatenate([small.label, large.label]), None,
                        np.concatenate([small.offset, large.offset]),
                        np.concatenate([small.theta_ptr, large.theta_ptr[1:] + small.theta_ptr[-1]]))
    n_e = np.diff(hb.ent_rowptr)
    assert n_e.max() > 3 * n_e.mean()
    opts = capi.make_opts(l2=1.0)
    db = capi.DeviceBatch(hb)
    two = capi.re_fit_device(db, opts)
    torch.cuda.synchronize()
    plan = capi.last_plan()
    assert plan["fast"] == 1 and "typical" in plan, plan
    counters = two["workspace"][:32].view(torch.int32).cpu().numpy()
    print("plan", plan, "tier-1 deferred", counters[7], "tier-2 deferred", counters[2])
    assert 0 < counters[7] < hb.n_entities and counters[2] == 0
    monkeypatch.setenv("GDMIX_FAST_TIERS", "1")
    one = capi.re_fit_device(db, opts)
    torch.cuda.synchronize()
    assert "typical" not in capi.last_plan()
    for k in ("nit", "nfev", "status"):
        assert torch.equal(two[k], one[k]), k
    rel = _rel_per_entity(two["theta"].cpu().numpy(), one["theta"].cpu().numpy(), hb.theta_ptr)
    assert rel.max() <= 1e-9, rel.max()
    th_o, f_o, nit_o, nfev_o, st_o = O.re_fit_batch(_oracle_batch(hb), _oracle_opts(opts))
    rel = _rel_per_entity(two["theta"].cpu().numpy(), th_o, hb.theta_ptr)
    assert rel.max() <= REL_TOL, rel.max()
    assert (two["nit"].cpu().numpy() == nit_o).all()


def test_narrow_column_transfers_are_equivalent():
    """gdmix_re_batch.col8 / col16 (1 or 2 bytes per local column index across PCIe) give bit-identical fits to
    the int32 columns; 1003 non-zeros per entity block exercises the 16-at-a-time widening kernel's tail."""
    import ctypes as C
    hb = make_batch(37, 17, 59, 3, seed=5)
    opts = capi.make_opts()
    outs = []
    for mode in ("i32", "u16", "u8"):
        cb = hb.c_struct(narrow=False)
        keep = None
        if mode == "u16":
            keep = hb.col.astype(np.uint16); cb.col16 = keep.ctypes.data; cb.col = None
        elif mode == "u8":
            keep = hb.col.astype(np.uint8); cb.col8 = keep.ctypes.data; cb.col = None
        theta = np.zeros(hb.n_coef); nit = np.zeros(hb.n_entities, np.int32)
        capi.check(capi.lib.gdmix_re_fit_host(C.byref(cb), C.byref(opts), None, theta.ctypes.data_as(C.c_void_p), None,
                                              nit.ctypes.data_as(C.c_void_p), None, None, None, C.c_int64(0)))
        outs.append((theta, nit))
    for t, n in outs[1:]:
        np.testing.assert_array_equal(t, outs[0][0])
        np.testing.assert_array_equal(n, outs[0][1])
    assert hb.c_struct().col8 is not None and hb.c_struct().col16 is None


def test_c1_shape_at_scale_properties():
    """BASELINE.json configs[1] shape at 100 000 entities generated on the device (the bench's generator): every
    entity converges, the objective the solver reports equals gdmix_re_loss_grad at the returned coefficients,
    the gradient there passes scipy's pgtol test or the step was stopped by factr, and a random sample of entities
    pulled back to the host matches the oracle (coefficients <= 1e-5 relative, identical iteration counts)."""
    import ctypes as C
    from gdmix_b200.synthetic import make_device_batch
    E, n, d, k = 100_000, 128, 256, 32
    data = make_device_batch(E, n, d, k, seed=424242)
    cb = capi.ReBatch(E, data["n_rows"], data["nnz"], data["ent_rowptr"].data_ptr(), data["rowptr"].data_ptr(),
                      data["col"].data_ptr(), data["val"].data_ptr(), data["label"].data_ptr(), None,
                      data["offset"].data_ptr(), data["theta_ptr"].data_ptr(), data["max_rows"], data["max_nnz"],
                      data["max_coef"], 0)
    opts = capi.make_opts(l2=1.0, regularize_bias=False)
    ws = torch.empty(max(capi.re_workspace_size(cb, opts), 256), dtype=torch.uint8, device="cuda")
    theta = torch.empty(data["n_coef"], dtype=torch.float64, device="cuda")
    f = torch.empty(E, dtype=torch.float64, device="cuda")
    nit = torch.empty(E, dtype=torch.int32, device="cuda")
    nfev = torch.empty_like(nit)
    status = torch.empty_like(nit)
    P = lambda t: C.c_void_p(t.data_ptr())
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    capi.check(capi.lib.gdmix_re_fit(C.byref(cb), C.byref(opts), None, P(theta), P(f), P(nit), P(nfev), P(status), None,
                                     P(ws), C.c_size_t(ws.numel()), st))
    f2 = torch.empty_like(f)
    g2 = torch.empty_like(theta)
    capi.check(capi.lib.gdmix_re_loss_grad(C.byref(cb), C.byref(opts), P(theta), P(f2), P(g2), P(ws),
                                           C.c_size_t(ws.numel()), st))
    torch.cuda.synchronize()
    assert bool((status == 0).all())
    assert torch.allclose(f, f2, rtol=1e-13, atol=0.0)
    gmax = g2.abs().view(E, d + 1).max(1).values
    small = float((gmax <= 1e-5).double().mean().item())
    print("max|g| <= pgtol for", small, "of the entities; mean nit", float(nit.float().mean()))
    assert small > 0.99 and float(gmax.max()) < 1e-3     # any rest stopped on factr, a hair above pgtol
    # a sample against the oracle
    rng = np.random.default_rng(0)
    pick = np.sort(rng.choice(E, 48, replace=False))
    rows = (pick[:, None] * n + np.arange(n)[None, :]).reshape(-1)
    nz = (rows[:, None] * k + np.arange(k)[None, :]).reshape(-1)
    tr = torch.from_numpy(rows).cuda()
    tz = torch.from_numpy(nz).cuda()
    hb = capi.HostBatch(np.arange(len(pick) + 1, dtype=np.int64) * n, np.arange(len(rows) + 1, dtype=np.int64) * k,
                        data["col"][tz].cpu().numpy(), data["val"][tz].cpu().numpy(), data["label"][tr].cpu().numpy(),
                        None, data["offset"][tr].cpu().numpy(), np.arange(len(pick) + 1, dtype=np.int64) * (d + 1))
    th_o, f_o, nit_o, nfev_o, st_o = O.re_fit_batch(_oracle_batch(hb), _oracle_opts(opts))
    th_d = theta.view(E, d + 1)[torch.from_numpy(pick).cuda()].cpu().numpy().reshape(-1)
    rel = _rel_per_entity(th_d, th_o, hb.theta_ptr)
    assert rel.max() <= REL_TOL, rel.max()
    assert (nit.cpu().numpy()[pick] == nit_o).all() and (nfev.cpu().numpy()[pick] == nfev_o).all()


def test_warp_per_entity_tier_solves_what_fits_and_defers_the_rest(monkeypatch, re_path):
    """GDMIX_RE_PATH=small: the warp-per-entity kernel (re_small.cuh) takes the entities that fit its slices -- sized
    for twice the mean entity -- and appends the others to the list the rest of the cascade drains; answers equal the
    oracle's either way (identical iteration counts), warm starts and weights included."""
    if re_path != "small":
        pytest.skip("one configuration is enough")
    hb = make_batch(600, 24, 40, 6, seed=77, ragged=True, weights=True)     # 8 .. 1024 samples: the long ones defer
    opts = capi.make_opts(l2=0.3, regularize_bias=True)
    rng = np.random.default_rng(5)
    theta0 = 0.1 * rng.standard_normal(hb.n_coef)
    db = capi.DeviceBatch(hb)
    out = capi.re_fit_device(db, opts, theta0=torch.from_numpy(theta0).cuda())
    torch.cuda.synchronize()
    plan = capi.last_plan()
    deferred = int(out["workspace"][:64].view(torch.int32)[11].item())
    assert "small" in plan and 0 < deferred < hb.n_entities
    th_o, f_o, nit_o, nfev_o, st_o = O.re_fit_batch(_oracle_batch(hb), _oracle_opts(opts), theta0=theta0)
    rel = _rel_per_entity(out["theta"].cpu().numpy(), th_o, hb.theta_ptr)
    assert rel.max() <= 1e-9, rel.max()
    assert (out["nit"].cpu().numpy() == nit_o).all() and (out["status"].cpu().numpy() == st_o).all()


def test_pinned_pool_reuses_blocks_and_narrow_columns_checks_its_range():
    """_capi.pinned_empty: page-locked blocks come back from the pool when the last view of them dies and are handed
    out again; gdmix_narrow_columns refuses a value that does not fit the width."""
    import gc
    a = capi.pinned_empty(1 << 20, np.float32)
    assert torch.from_numpy(a).is_pinned() or capi._pinned_pool.available   # page-locked (torch sees cudaHostAlloc memory)
    addr = a.ctypes.data
    a[:] = 1.0
    view = a[10:20]
    del a
    gc.collect()
    b = capi.pinned_empty(1 << 20, np.float32)
    bdata = b.ctypes.data
    assert bdata != addr                    # the first block is still alive behind `view`
    del view, b
    gc.collect()
    c = capi.pinned_empty(1 << 20, np.float32)
    assert c.ctypes.data in {addr, bdata}   # ... and now one of the two freed blocks is handed out again
    col = np.array([0, 5, 255, 256], np.int32)
    with pytest.raises(capi.GdmixError):
        capi.narrow_columns(col, 1)
    np.testing.assert_array_equal(capi.narrow_columns(col, 2), col.astype(np.uint16))
    np.testing.assert_array_equal(capi.narrow_columns(col[:3], 1), col[:3].astype(np.uint8))
