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


# ---- DataPartitioner's bounds / OffsetUpdater's join / the partitioned files ------------------------------------------
# the fixture of gdmix-data/src/test/scala/com/linkedin/gdmix/data/DataPartitionerTest.scala:25-45
REF_UID = np.arange(10, dtype=np.int64)
REF_ENTITY = np.array([0, 0, 0, 1, 1, 1, 1, 1, 1, 2], np.int64)
REF_LABEL = np.array([0, 0, 1, 1, 1, 0, 0, 1, 1, 1], np.float32)
REF_INDICES = [[0, 1], [0, 1, 2], [3, 4], [5, 6], [7, 8], [9, 10], [3, 4], [5, 9], [0], [0, 2]]
REF_VALUES = [[0, 1], [0, 1.0, 2.2], [3, 4.1], [5.5, 6.6], [7.7, 8.8], [9.3, 10.12], [0.3, 0.8], [0.8, 1.8], [0.0],
              [1.0, -2.2]]


def _np_group_ids(entity, uid, lower, upper):
    """numpy restatement of DataPartitioner.getGroupId (DataPartitioner.scala:335-379)."""
    ids, inv, cnt = np.unique(entity, return_inverse=True, return_counts=True)
    count = cnt[inv]
    groups = count // upper + 1 if upper else np.ones_like(count)
    gid = np.mod(uid, groups)                       # pmod: numpy's mod already has the divisor's sign
    if lower:
        gid = np.where(count < lower, -1, gid)
    return gid.astype(np.int32)


def test_group_ids_reference_fixture_and_random():
    """DataPartitionerTest.testGetGroupId: lowerBound 2, upperBound 4 -- entity 0 (3 rows) all active, entity 1 (6 rows)
    one or two groups, entity 2 (1 row) passive with id -1; then random ids / uids (negative uids: pmod) against the
    restatement, every combination of bounds."""
    from gdmix_b200 import partition as P
    ent, uid = torch.from_numpy(REF_ENTITY).cuda(), torch.from_numpy(REF_UID).cuda()
    gid = P.group_ids(ent, uid, 2, 4).cpu().numpy()
    assert (gid[REF_ENTITY == 0] == 0).all() and (gid[REF_ENTITY == 2] == -1).all()
    assert 1 <= len(set(gid[REF_ENTITY == 1])) <= 2
    np.testing.assert_array_equal(gid, _np_group_ids(REF_ENTITY, REF_UID, 2, 4))
    rng = np.random.default_rng(3)
    entity = rng.integers(0, 5000, 200_000) ** 2 % 7919
    uids = rng.integers(-1 << 40, 1 << 40, 200_000)
    e, u = torch.from_numpy(entity).cuda(), torch.from_numpy(uids).cuda()
    for lower, upper in ((None, None), (5, None), (None, 7), (3, 20), (40, 10)):
        got = P.group_ids(e, u, lower, upper).cpu().numpy()
        np.testing.assert_array_equal(got, _np_group_ids(entity, uids, lower, upper))


def test_offset_join_reference_fixture_and_random():
    """OffsetUpdaterTest.testUpdateOffset (1.0 / 2.0, and 0.9 / 1.8 with the per-coordinate score subtracted in fp32),
    then a shuffled score file that misses some uids: the inner join's surviving rows and their offsets."""
    from gdmix_b200 import partition as P
    uid = torch.tensor([1, 2], dtype=torch.int64).cuda()
    s_uid = torch.tensor([2, 1], dtype=torch.int64).cuda()
    score = torch.tensor([2.0, 1.0], dtype=torch.float32).cuda()
    pc = torch.tensor([0.2, 0.1], dtype=torch.float32).cuda()
    off, m = P.join_offsets(uid, s_uid, score)
    assert m.all() and off.cpu().tolist() == [1.0, 2.0]
    off, m = P.join_offsets(uid, s_uid, score, pc)
    np.testing.assert_array_equal(off.cpu().numpy(), np.array([1.0, 2.0], np.float32) - np.array([0.1, 0.2], np.float32))
    rng = np.random.default_rng(8)
    n = 100_000
    uids = rng.permutation(1 << 20)[:n].astype(np.int64) - 1000
    keep = rng.random(n) < 0.9
    order = rng.permutation(int(keep.sum()))
    s_uid_h = uids[keep][order]
    s_h = rng.standard_normal(s_uid_h.shape[0]).astype(np.float32)
    pc_h = rng.standard_normal(s_uid_h.shape[0]).astype(np.float32)
    off, m = P.join_offsets(torch.from_numpy(uids).cuda(), torch.from_numpy(s_uid_h).cuda(), torch.from_numpy(s_h).cuda(),
                            torch.from_numpy(pc_h).cuda())
    np.testing.assert_array_equal(m.cpu().numpy(), keep)
    want = dict(zip(s_uid_h.tolist(), (s_h - pc_h).tolist()))
    got = off.cpu().numpy()
    np.testing.assert_array_equal(got[keep], np.array([want[u] for u in uids[keep].tolist()], np.float32))


def test_partition_and_write_layout_matches_the_spark_job(tmp_path):
    """groupPartitionAndSaveDataset on the reference's fixture (DataPartitionerTest.scala:25-45, 194-224) and on random
    data: active|passive/partitionId=k files of one SequenceExample per (entity, group), rows of a record in arrival
    order, partition = abs(hashCode(str(entity))) % n, partitionList.txt = sorted partitions that hold a record,
    offsets joined by uid with unmatched rows dropped -- read back with the package's reader and compared with a numpy
    restatement."""
    from gdmix_b200 import partition as P
    from gdmix_b200.io import tfrecord as T
    from oracle import oracle as O

    def run(entity, uid, lens, cols, vals, label, nparts, lower, upper, scores=None, sub="a"):
        rowptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        out = tmp_path / sub
        to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
        sc = None if scores is None else tuple(to(x) for x in scores)
        plist = out / "partitionList.txt"
        P.partition_and_write(str(out), to(entity), to(uid), to(rowptr), to(cols.astype(np.int32)), to(vals), to(label), nparts,
                              scores=sc, lower_bound=lower, upper_bound=upper, entity_name="entityId", bag="global",
                              label_name="label", partition_list_file=str(plist))
        # restatement
        live = np.ones(len(entity), bool)
        offs = np.zeros(len(entity), np.float32)
        if scores is not None:
            lut = {int(u): float(s) for u, s in zip(scores[0][::-1], scores[1][::-1])}
            live = np.array([int(u) in lut for u in uid])
            offs = np.array([lut.get(int(u), 0.0) for u in uid], np.float32)
        gid = np.full(len(entity), 99, np.int32)
        gid[live] = _np_group_ids(entity[live], uid[live], lower, upper)
        want = {}
        for i in np.flatnonzero(live):
            k = O.partition_id(str(int(entity[i])), nparts)
            cls = "active" if gid[i] == 0 else "passive"
            want.setdefault((cls, k), {}).setdefault((int(entity[i]), int(gid[i])), []).append(i)
        assert sorted(int(x) for x in plist.read_text().split(",")) == sorted({k for (_, k) in want})
        for (cls, k), recs in want.items():
            f = out / cls / f"partitionId={k}" / "part-00000.tfrecord"
            assert f.exists(), (cls, k)
            d = capi.parse_entity_grouped(f.read_bytes(), "entityId", "uid", "label", "offset", None, "global_indices",
                                          "global_values")
            assert sum(1 for _ in T.read_records(str(f), verify_crc=True)) == len(recs)
            got = {}
            r = q = 0
            for e, n in enumerate(d["ent_rows"]):
                rows = []
                for i in range(n):
                    kk = d["row_len"][r + i]
                    rows.append((int(d["uid"][r + i]), float(d["label"][r + i]), float(d["offset"][r + i]),
                                 d["gcol"][q:q + kk].tolist(), d["val"][q:q + kk].tolist()))
                    q += kk
                r += n
                got.setdefault(int(d["entity_ids"][e]), []).append(rows)
            exp = {}
            for (ent, g), idx in recs.items():
                exp.setdefault(ent, []).append([(int(uid[i]), float(label[i]), float(offs[i]),
                                                 cols[rowptr[i]:rowptr[i + 1]].tolist(), vals[rowptr[i]:rowptr[i + 1]].tolist())
                                                for i in idx])
            assert {e: sorted(v) for e, v in got.items()} == {e: sorted(v) for e, v in exp.items()}
        files = {(p.parent.parent.name, int(p.parent.name.split("=")[1])) for p in out.glob("*/partitionId=*/*.tfrecord")}
        assert files == set(want)

    lens = np.array([len(x) for x in REF_INDICES])
    cols = np.concatenate(REF_INDICES).astype(np.int64)
    vals = np.concatenate(REF_VALUES).astype(np.float32)
    run(REF_ENTITY, REF_UID, lens, cols, vals, REF_LABEL, 3, 2, 4, sub="ref")
    rng = np.random.default_rng(12)
    n = 5000
    entity = (rng.integers(0, 300, n) ** 2 % 401).astype(np.int64)
    uid = rng.permutation(n).astype(np.int64)
    lens = rng.integers(0, 6, n)
    cols = rng.integers(0, 1000, int(lens.sum())).astype(np.int64)
    vals = rng.standard_normal(int(lens.sum())).astype(np.float32)
    label = rng.integers(0, 2, n).astype(np.float32)
    run(entity, uid, lens, cols, vals, label, 7, 4, 9, sub="rnd")
    keep = rng.random(n) < 0.8
    s_uid = uid[keep][rng.permutation(int(keep.sum()))]
    run(entity, uid, lens, cols, vals, label, 5, None, 6, scores=(s_uid, rng.standard_normal(s_uid.shape[0]).astype(np.float32)),
        sub="joined")
