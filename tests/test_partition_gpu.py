"""Device partitioner / evaluator (csrc/partition.cuh) against numpy: bit-exact for the integer work (stable
sort, group-by, CSR regrouping, entity -> partition), 1e-12 for AUC; plus the known answers of the reference's
EvaluatorTest.scala:19-33 and the chained use: regroup on the device -> gdmix_re_fit -> same coefficients as
grouping on the host."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
if not torch.cuda.is_available():  # pragma: no cover
    pytest.skip("needs a CUDA device", allow_module_level=True)

from gdmix_b200 import _capi as capi, partition as P  # noqa: E402
from gdmix_b200.synthetic import make_batch  # noqa: E402


@pytest.mark.parametrize("n,bits", [(1, 7), (33, 3), (4096, 16), (4097, 1), (100_003, 40), (1_000_000, 23)])
def test_sort_is_stable_and_exact(n, bits):
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 1 << bits, n, dtype=np.int64)
    ks, perm = P.sort_pairs(torch.from_numpy(keys).cuda())
    order = np.argsort(keys, kind="stable")
    np.testing.assert_array_equal(perm.cpu().numpy(), order.astype(np.int32))
    np.testing.assert_array_equal(ks.cpu().numpy(), keys[order])


def test_group_by_entity_matches_numpy():
    rng = np.random.default_rng(1)
    n = 300_000
    ent = rng.zipf(1.3, n).astype(np.int64) % 50_000       # skewed entity sizes, many singletons, one huge
    perm, seg_ptr, seg_key = P.group_by_entity(torch.from_numpy(ent).cuda())
    uniq, counts = np.unique(ent, return_counts=True)
    np.testing.assert_array_equal(seg_key.cpu().numpy(), uniq)
    np.testing.assert_array_equal(np.diff(seg_ptr.cpu().numpy()), counts)
    np.testing.assert_array_equal(perm.cpu().numpy(), np.argsort(ent, kind="stable").astype(np.int32))
    # empty and single-entity inputs
    p0, s0, k0 = P.group_by_entity(torch.zeros(0, dtype=torch.int64, device="cuda"))
    assert p0.numel() == 0 and s0.tolist() == [0] and k0.numel() == 0
    p1, s1, k1 = P.group_by_entity(torch.full((5000,), 7, dtype=torch.int64, device="cuda"))
    assert s1.tolist() == [0, 5000] and k1.tolist() == [7] and p1.tolist() == list(range(5000))


def test_partition_ids_match_the_string_hash():
    ids = np.array([0, 7, 10, 123456789, 9007199254740993, -5, 2 ** 62, 99999, 100034], np.int64)
    got = P.partition_ids(torch.from_numpy(ids).cuda(), 17).cpu().numpy()
    _, exp = capi.partition_ids([str(int(v)) for v in ids], 17)   # host version on the decimal strings
    np.testing.assert_array_equal(got, exp)


@pytest.mark.parametrize("score,label,auc", [
    ([0.1, 0.4, 0.35, 0.8], [0, 0, 1.0, 1.0], 0.75),
    ([0.5, 0.7, 0.3, 0.4, 0.45, 0.8], [0, 0, 1.0, 1.0, 0, 1.0], 0.3333333),
    ([0.5, 0.75, 0.8, 0.2, 0.3, 0.4, 0.45, 0.5], [0, 0, 0, 0, 1.0, 1.0, 0, 1.0], 0.3)])
def test_auc_reference_known_answers(score, label, auc):
    """EvaluatorTest.scala:19-33, delta 1e-5."""
    got = P.auc(torch.tensor(score, dtype=torch.float32).cuda(), torch.tensor(label, dtype=torch.float32).cuda())
    assert abs(got - auc) <= 1e-5


def test_auc_with_ties_matches_rank_formula():
    rng = np.random.default_rng(2)
    n = 700_000
    s = np.round(rng.standard_normal(n), 2).astype(np.float32)   # heavy ties, negative scores, +-0
    y = (rng.random(n) < 1 / (1 + np.exp(-2 * s))).astype(np.float32)
    got = P.auc(torch.from_numpy(s).cuda(), torch.from_numpy(y).cuda())
    # Mann-Whitney U with midranks
    order = np.argsort(s, kind="stable")
    ss = s[order]
    ranks = np.empty(n)
    i = 0
    heads = np.flatnonzero(np.concatenate([[True], ss[1:] != ss[:-1]]))
    ends = np.concatenate([heads[1:], [n]])
    mid = (heads + 1 + ends) / 2.0
    ranks[order] = np.repeat(mid, ends - heads)
    npos = float(y.sum(dtype=np.float64))
    nneg = n - npos
    exp = (ranks[y > 0].sum() - npos * (npos + 1) / 2) / (npos * nneg)
    assert abs(got - exp) <= 1e-12


def test_regroup_on_device_then_fit_matches_host_grouping():
    """Rows shuffled as they would arrive from a fixed-effect scoring pass -> device group-by + local indexing ->
    gdmix_re_fit: the same coefficients as the host-grouped batch, entity by entity."""
    E, n, d, k, D = 400, 24, 16, 5, 3000
    hb = make_batch(E, n, d, k, seed=5, ragged=True, weights=True)
    rng = np.random.default_rng(3)
    gmap = np.stack([np.sort(rng.choice(D, d, replace=False)) for _ in range(E)])
    ent_of_row = np.repeat(np.arange(E), np.diff(hb.ent_rowptr))
    ent_ids = (rng.permutation(10 * E)[:E]).astype(np.int64)           # arbitrary, unordered entity ids
    ent_of_nnz = np.repeat(ent_of_row, np.diff(hb.rowptr))
    gcol = gmap[ent_of_nnz, hb.col].astype(np.int32)
    shuffle = rng.permutation(hb.n_rows)
    lens = np.diff(hb.rowptr)[shuffle]
    rowptr_s = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    idx = np.concatenate([np.arange(hb.rowptr[r], hb.rowptr[r + 1]) for r in shuffle])
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    out = P.regroup_batch(t(ent_ids[ent_of_row][shuffle]), t(rowptr_s), t(gcol[idx]), t(hb.val[idx]),
                          t(hb.label[shuffle]), t(hb.offset[shuffle]), t(hb.weight[shuffle]), num_features=D)
    assert out["n_entities"] == E
    cb = capi.ReBatch(E, out["n_rows"], out["nnz"], out["ent_rowptr"].data_ptr(), out["rowptr"].data_ptr(),
                      out["col"].data_ptr(), out["val"].data_ptr(), out["label"].data_ptr(), out["weight"].data_ptr(),
                      out["offset"].data_ptr(), out["theta_ptr"].data_ptr(), out["max_rows"], out["max_nnz"],
                      out["max_coef"], 0, None)
    opts = capi.make_opts(l2=1.0)
    import ctypes as C
    ws = torch.empty(max(capi.re_workspace_size(cb, opts), 256), dtype=torch.uint8, device="cuda")
    theta = torch.zeros(out["n_coef"], dtype=torch.float64, device="cuda")
    status = torch.zeros(E, dtype=torch.int32, device="cuda")
    capi.check(capi.lib.gdmix_re_fit(C.byref(cb), C.byref(opts), None, C.c_void_p(theta.data_ptr()), None, None, None,
                                     C.c_void_p(status.data_ptr()), None, C.c_void_p(ws.data_ptr()),
                                     C.c_size_t(ws.numel()), None))
    torch.cuda.synchronize()
    assert (status.cpu().numpy() == 0).all()
    ref = capi.re_fit_host(hb, opts)
    theta = theta.cpu().numpy()
    tp = out["theta_ptr"].cpu().numpy()
    up, ug = out["uniq_ptr"].cpu().numpy(), out["uniq_global"].cpu().numpy()
    eid = out["entity_ids"].cpu().numpy()
    inv = {int(v): e for e, v in enumerate(ent_ids)}
    for g in range(0, E, 3):
        e = inv[int(eid[g])]
        th_ref = ref["theta"][hb.theta_ptr[e]:hb.theta_ptr[e + 1]]
        # features of the host batch that actually occur in the entity's rows (the device never sees the others)
        used = np.unique(hb.col[hb.rowptr[hb.ent_rowptr[e]]:hb.rowptr[hb.ent_rowptr[e + 1]]])
        np.testing.assert_array_equal(ug[up[g]:up[g + 1]], gmap[e][used])
        got = theta[tp[g]:tp[g + 1]]
        np.testing.assert_allclose(got[0], th_ref[0], rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(got[1:], th_ref[1 + used], rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("D", [20, 64, 100, 2048])
def test_bitmap_local_indexing_equals_the_pair_sort(D, monkeypatch):
    """Feature bags of up to 2048 ids are indexed per entity by presence bitmaps (gdmix_local_index_*); the result
    is the one the (entity, feature) pair sort gives, bit for bit: np.unique(cols, return_inverse=True) per entity
    (job_consumers.py:243)."""
    rng = np.random.default_rng(D)
    n, E = 5000, 300
    ent = rng.integers(0, E, n).astype(np.int64) * 3 + 1
    lens = rng.integers(0, 9, n)
    rowptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    nnz = int(rowptr[-1])
    gcol = rng.integers(0, D, nnz).astype(np.int32)            # repeats inside a row included
    val = rng.standard_normal(nnz).astype(np.float32)
    y = (rng.random(n) < 0.5).astype(np.float32)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    a = P.regroup_batch(t(ent), t(rowptr), t(gcol), t(val), t(y), None, None, num_features=D)
    monkeypatch.setattr(P, "FORCE_PAIR_SORT", True)
    b = P.regroup_batch(t(ent), t(rowptr), t(gcol), t(val), t(y), None, None, num_features=D)
    for k in ("ent_rowptr", "rowptr", "col", "val", "label", "theta_ptr", "perm", "entity_ids", "uniq_ptr", "uniq_global"):
        assert torch.equal(a[k], b[k]), k
    for k in ("n_entities", "n_rows", "nnz", "max_rows", "max_nnz", "max_coef", "n_coef"):
        assert a[k] == b[k], k
    # and against numpy, entity by entity
    col, up, ug = a["col"].cpu().numpy(), a["uniq_ptr"].cpu().numpy(), a["uniq_global"].cpu().numpy()
    er, rp, perm = a["ent_rowptr"].cpu().numpy(), a["rowptr"].cpu().numpy(), a["perm"].cpu().numpy().astype(np.int64)
    for e in range(0, a["n_entities"], 17):
        rows = perm[er[e]:er[e + 1]]
        g = np.concatenate([gcol[rowptr[r]:rowptr[r + 1]] for r in rows]) if len(rows) else np.zeros(0, np.int32)
        u, inv = np.unique(g, return_inverse=True)
        np.testing.assert_array_equal(ug[up[e]:up[e + 1]], u)
        np.testing.assert_array_equal(col[rp[er[e]]:rp[er[e + 1]]], inv)


def test_bitmap_local_indexing_rejects_ids_outside_the_bag():
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    with pytest.raises(capi.GdmixError):
        P.regroup_batch(t(np.array([5, 5, 7], np.int64)), t(np.array([0, 1, 2, 3], np.int64)),
                        t(np.array([1, 64, 2], np.int32)), t(np.ones(3, np.float32)), t(np.ones(3, np.float32)), None, None,
                        num_features=64)
