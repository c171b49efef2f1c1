"""The library's SequenceExample reader (gdmix_seqex_count / gdmix_seqex_fill, csrc/seqex_parser.h) against the
pure-Python protobuf walk of gdmix_b200/io/tfrecord.py -- which tests/test_io_formats.py pins to files TensorFlow
wrote -- on the reference's own fixture and on generated files: int64 / bytes entity ids, int64 / float labels,
missing optional columns, empty samples, negative and large int64 values, packed and unpacked encodings, gzip."""
import os
import struct

import numpy as np
import pytest

from gdmix_b200 import _capi as capi, ingest
from gdmix_b200.io import tfrecord as T

FIX = os.path.join(os.path.dirname(__file__), "golden", "ref_fixtures")


def _python_reader(files, **kw):
    """read_entity_grouped's Python walk, reached by hiding the native path."""
    saved = ingest._read_entity_grouped_native
    ingest._read_entity_grouped_native = lambda *a, **k: None
    try:
        return ingest.read_entity_grouped(files, **kw)
    finally:
        ingest._read_entity_grouped_native = saved


def _same(a, b):
    assert a.entity_ids == b.entity_ids
    for k in ("ent_rowptr", "rowptr", "gcol", "val", "uid", "offset", "weight"):
        np.testing.assert_array_equal(getattr(a, k), getattr(b, k), err_msg=k)
        assert getattr(a, k).dtype == getattr(b, k).dtype, k
    assert (a.label is None) == (b.label is None)
    if a.label is not None:
        np.testing.assert_array_equal(a.label, b.label)
    assert a.has_weight_column == b.has_weight_column and a.num_features == b.num_features


class _Meta:
    def __init__(self, names):
        self.names = names

    def get_feature_names(self):
        return self.names


def test_reference_fixture_matches_appendix_d():
    kw = dict(metadata=_Meta(["memberId"]), entity_name="memberId", feature_bag="per_member", label_column="response",
              offset_column="offset", weight_column="weight", uid_column="uid", num_features=100)
    path = os.path.join(FIX, "re_data.tfrecord")
    nat = ingest.read_entity_grouped(path, **kw)
    _same(nat, _python_reader(path, **kw))
    assert nat.entity_ids == ["100034", "100"]
    np.testing.assert_array_equal(nat.gcol, [0, 7, 60, 80, 95, 34, 57, 10, 11])
    np.testing.assert_array_equal(nat.uid, [10, 20, 23])
    np.testing.assert_allclose(nat.val, np.array([1, 2, 3, 5, 6.6, 1, 2, -3.5, 2.3], np.float32))


def _write(path, rng, n_ent, bytes_ids=False, float_label=False, with_label=True, with_weight=True, with_offset=True,
           unpacked=False):
    with T.TFRecordWriter(path) as w:
        for e in range(n_ent):
            n = int(rng.integers(1, 9))
            ctx = {"ent": [("user-%d-é" % e).encode("utf-8")] if bytes_ids else [int(rng.integers(-5, 10**12))],
                   "uid": [int(x) for x in rng.integers(-2**62, 2**62, n)]}
            if with_label:
                ctx["y"] = [float(x) for x in rng.integers(0, 2, n)] if float_label else [int(x) for x in rng.integers(0, 2, n)]
            if with_weight:
                ctx["w"] = [float(x) for x in rng.uniform(0.5, 2, n)]
            if with_offset:
                ctx["off"] = [float(x) for x in rng.standard_normal(n)]
            lens = rng.integers(0, 6, n)
            fl = {"bag_indices": [[int(x) for x in np.sort(rng.choice(5000, k, replace=False))] for k in lens],
                  "bag_values": [[float(x) for x in rng.standard_normal(k)] for k in lens],
                  "other_indices": [[1, 2]] * n}
            payload = T.encode_sequence_example(ctx, fl)
            w.write(_unpack(payload) if unpacked else payload)


def _unpack(payload):
    """Re-encodes every packed FloatList / Int64List of a SequenceExample element by element (both are legal)."""
    def walk(buf, depth):
        out = bytearray()
        for fno, wt, v in T._fields(buf, 0, len(buf)):
            if wt == 2:
                sub = bytes(buf[v[0]:v[1]])
                if depth == "feature" and fno in (2, 3):
                    inner = bytearray()
                    for f2, w2, v2 in T._fields(sub, 0, len(sub)):
                        if f2 == 1 and w2 == 2:
                            body = sub[v2[0]:v2[1]]
                            if fno == 2:
                                for i in range(0, len(body), 4):
                                    inner += T._enc_varint((1 << 3) | 5) + body[i:i + 4]
                            else:
                                for x in T._packed_varints(body, 0, len(body)):
                                    inner += T._enc_varint((1 << 3) | 0) + T._enc_varint(int(x))
                    sub = bytes(inner)
                elif depth in ("top", "features", "lists", "entry", "flist"):
                    nxt = {"top": "features" if fno == 1 else "lists", "features": "entry", "lists": "entry",
                           "entry": "feature" if fno == 2 else None, "flist": "feature"}[depth]
                    if depth == "entry" and fno == 2 and _is_feature_list(sub):
                        nxt = "flist"
                    if nxt:
                        sub = walk(sub, nxt)
                out += T._enc_varint((fno << 3) | 2) + T._enc_varint(len(sub)) + sub
            elif wt == 0:
                out += T._enc_varint((fno << 3) | 0) + T._enc_varint(v)
            elif wt == 5:
                out += T._enc_varint((fno << 3) | 5) + struct.pack("<I", v)
        return bytes(out)

    def _is_feature_list(sub):
        # a FeatureList holds repeated Feature messages in field 1; a Feature holds a list in fields 1..3: tell
        # them apart by looking one level down (a Feature's payload never starts with another field-1 message
        # whose own payload parses as a Feature -- good enough for the files this test writes)
        try:
            f = list(T._fields(sub, 0, len(sub)))
            return bool(f) and all(fno == 1 and wt == 2 for fno, wt, _ in f) and \
                all(k in ("bytes", "float", "int64", None) for k in (T._parse_feature(sub, a, b)[0] for _, _, (a, b) in f))
        except Exception:
            return False
    return walk(payload, "top")


@pytest.mark.parametrize("variant", [dict(), dict(bytes_ids=True), dict(float_label=True), dict(with_label=False),
                                     dict(with_weight=False, with_offset=False), dict(unpacked=True)])
@pytest.mark.parametrize("suffix", [".tfrecord", ".tfrecord.gz"])
def test_generated_files_match_the_python_reader(tmp_path, variant, suffix):
    rng = np.random.default_rng(len(str(variant)))
    for part in range(2):
        _write(str(tmp_path / f"part-{part:05d}{suffix}"), rng, 40, **variant)
    kw = dict(metadata=_Meta(["ent"]), entity_name="ent", feature_bag="bag", label_column="y", offset_column="off",
              weight_column="w", uid_column="uid", num_features=5000)
    nat = ingest.read_entity_grouped(str(tmp_path), **kw)
    assert nat.n_entities == 80
    _same(nat, _python_reader(str(tmp_path), **kw))


def test_malformed_files_raise(tmp_path):
    rng = np.random.default_rng(1)
    good = str(tmp_path / "a.tfrecord")
    _write(good, rng, 5)
    img = T._read_all(good)
    spec = ("ent", "uid", "y", "off", "w", "bag_indices", "bag_values")
    capi.parse_entity_grouped(img, *spec)
    with pytest.raises(capi.GdmixError):
        capi.parse_entity_grouped(img[:-7], *spec)                     # truncated last record
    with pytest.raises(capi.GdmixError):
        capi.parse_entity_grouped(img, "nope", *spec[1:])              # entity column missing
    with pytest.raises(capi.GdmixError):
        capi.parse_entity_grouped(img, "ent", "uid", "y", "off", "w", "bag_indices", "other_indices")   # not floats
    # a sample column shorter than uid
    bad = str(tmp_path / "b.tfrecord")
    with T.TFRecordWriter(bad) as w:
        w.write(T.encode_sequence_example({"ent": [1], "uid": [1, 2, 3], "off": [0.5]},
                                          {"bag_indices": [[1], [2], [3]], "bag_values": [[1.0], [2.0], [3.0]]}))
    with pytest.raises(capi.GdmixError):
        capi.parse_entity_grouped(T._read_all(bad), *spec)
    # fewer index lists than samples
    bad2 = str(tmp_path / "c.tfrecord")
    with T.TFRecordWriter(bad2) as w:
        w.write(T.encode_sequence_example({"ent": [1], "uid": [1, 2]}, {"bag_indices": [[1]], "bag_values": [[1.0]]}))
    with pytest.raises(capi.GdmixError):
        capi.parse_entity_grouped(T._read_all(bad2), *spec)


def _records_equal(a, b):
    for k in ("rowptr", "col", "val", "uid", "offset", "weight"):
        np.testing.assert_array_equal(getattr(a, k), getattr(b, k), err_msg=k)
        assert getattr(a, k).dtype == getattr(b, k).dtype, k
    np.testing.assert_array_equal(a.label, b.label)        # NaN where a row has no label, in both
    assert a.has_weight_column == b.has_weight_column


def _per_record(files, native, **kw):
    saved = ingest.USE_NATIVE_READER
    ingest.USE_NATIVE_READER = native
    try:
        return ingest.read_per_record(files, **kw)
    finally:
        ingest.USE_NATIVE_READER = saved


def test_fixed_effect_fixture_rows_match_the_python_reader():
    kw = dict(feature_bag="global", label_column="response", offset_column="offset", weight_column="weight",
              uid_column="uid")
    path = os.path.join(FIX, "fe_test.tfrecord")
    nat = _per_record(path, True, **kw)
    assert nat.n_rows > 0 and nat.col.size > 0
    _records_equal(nat, _per_record(path, False, **kw))


@pytest.mark.parametrize("variant", ["full", "no_bag", "missing_columns", "float_label"])
def test_generated_example_files_match_the_python_reader(tmp_path, variant):
    rng = np.random.default_rng(7)
    path = str(tmp_path / "rows.tfrecord.gz")
    with T.TFRecordWriter(path) as w:
        for i in range(300):
            k = int(rng.integers(0, 7))
            ex = {"uid": [int(rng.integers(-2**60, 2**60))]}
            if variant != "no_bag":
                ex["g_indices"] = [int(x) for x in np.sort(rng.choice(100000, k, replace=False))]
                ex["g_values"] = [float(x) for x in rng.standard_normal(k)]
            if not (variant == "missing_columns" and i % 3 == 0):
                ex["y"] = [float(rng.integers(0, 2))] if variant == "float_label" else [int(rng.integers(0, 2))]
                ex["w"] = [float(rng.uniform(0.5, 2))]
            if variant != "missing_columns":
                ex["off"] = [float(rng.standard_normal())]
            w.write(T.encode_example(ex))
    kw = dict(feature_bag=None if variant == "no_bag" else "g", label_column="y", offset_column="off",
              weight_column="w", uid_column="uid")
    nat = _per_record(path, True, **kw)
    assert nat.n_rows == 300
    _records_equal(nat, _per_record(path, False, **kw))


@pytest.mark.parametrize("with_label", [True, False])
@pytest.mark.parametrize("with_weight", [True, False])
@pytest.mark.parametrize("n", [0, 1, 1024, 2500])
def test_native_score_writer_is_byte_identical_to_the_python_writer(tmp_path, with_label, with_weight, n):
    """model_io.write_scores (gdmix_avro_score_blocks) against avro.write_records one dict at a time, same sync
    marker: the files must be equal byte for byte, and decode to the columns that went in."""
    from types import SimpleNamespace
    from gdmix_b200.io import avro, model_io
    sp = SimpleNamespace(uid_column_name="uid", prediction_score_column_name="predictionScore",
                         label_column_name="label", weight_column_name="weight",
                         prediction_score_per_coordinate_column_name="predictionScorePerCoordinate")
    schema = model_io.get_inference_output_avro_schema({}, True, sp, has_weight=with_weight)
    rng = np.random.default_rng(n + 7)
    uid = rng.integers(-2**62, 2**62, n).astype(np.int64)
    if n:
        uid[0] = 0
    score, pc = rng.standard_normal(n).astype(np.float32), rng.standard_normal(n).astype(np.float32)
    label = rng.integers(0, 2, n).astype(np.float32) if with_label else None
    weight = rng.uniform(0.5, 2, n).astype(np.float32)
    sync = bytes(range(16))
    a, b = str(tmp_path / "native.avro"), str(tmp_path / "python.avro")
    assert model_io.write_scores(a, schema, sp, uid, score, pc, label=label, weight=weight, sync=sync) == n

    def records():
        for i in range(n):
            r = {"uid": int(uid[i]), "predictionScore": float(score[i]), "predictionScorePerCoordinate": float(pc[i]),
                 "weight": float(weight[i])}
            if with_label:
                r["label"] = float(label[i])
            yield r
    with avro.Writer(b, schema, "null", sync=sync) as w:
        batch = []
        for r in records():
            batch.append(r)
            if len(batch) == 1024:
                w.write_block(batch); batch = []
        w.write_block(batch)
    assert open(a, "rb").read() == open(b, "rb").read()
    back = list(avro.read_records(a))
    assert len(back) == n
    if n:
        assert back[0]["uid"] == 0 and back[-1]["uid"] == int(uid[-1])
        assert (back[5 % n]["label"] is None) == (not with_label)
        assert ("weight" in back[0]) == with_weight


@pytest.mark.parametrize("has_intercept", [True, False])
@pytest.mark.parametrize("with_variance", [True, False])
def test_native_model_writer_is_byte_identical_to_the_python_writer(tmp_path, has_intercept, with_variance):
    """model_io.export_random_effect_models (gdmix_avro_model_blocks) against export_linear_model_to_avro
    (gen_one_avro_model per entity), same sync marker: equal files, incl. thresholded coefficients, an entity whose
    every coefficient is thresholded, unicode names / ids, and more models than one block holds."""
    from gdmix_b200.io import avro, model_io
    rng = np.random.default_rng(3)
    D = 50
    ff = tmp_path / "features.csv"
    ff.write_text("".join(f"featé{j},term{j % 3 if j % 4 else ''}\n" for j in range(D)), encoding="utf-8")
    M = 1500
    ids, idx, vals, vars_, biases = [], [], [], [], []
    for m in range(M):
        d = int(rng.integers(0, 9))
        gi = np.sort(rng.choice(D, d, replace=False)).astype(np.int64)
        w = rng.standard_normal(d) * rng.choice([1.0, 1e-5], d)        # some fall under the 1e-4 threshold
        if m == 7:
            w[:] = 1e-6
        ids.append(f"user中{m}" if m % 2 else str(m * 97))
        idx.append(gi); vals.append(w); vars_.append(rng.uniform(0.1, 2, d)); biases.append((float(rng.standard_normal()), float(rng.uniform(0.1, 2))))
    sync = bytes(range(16, 32))
    a, b = str(tmp_path / "native.avro"), str(tmp_path / "python.avro")
    hi = 1 if has_intercept else 0
    coef = np.concatenate([np.concatenate([[biases[m][0]] if hi else [], vals[m]]) for m in range(M)])
    var = np.concatenate([np.concatenate([[biases[m][1]] if hi else [], vars_[m]]) for m in range(M)]) if with_variance else None
    coef_ptr = np.concatenate([[0], np.cumsum([len(v) + hi for v in vals])]).astype(np.int64)
    model_io.export_random_effect_models(ids, coef, var, coef_ptr, np.concatenate(idx), has_intercept, str(ff), a, sync=sync)
    # the reference-shaped writer, with the same sync marker
    feature_list = model_io.read_feature_list(str(ff))
    recs = [model_io.gen_one_avro_model(ids[m], model_io.LOGISTIC_MODEL_CLASS, idx[m],
                                        (vals[m], vars_[m]) if with_variance else vals[m],
                                        (biases[m] if with_variance else biases[m][0]) if hi else None, feature_list, 1e-4)
            for m in range(M)]
    with avro.Writer(b, model_io.BAYESIAN_LINEAR_MODEL_SCHEMA, "null", sync=sync) as w:
        for i in range(0, M, 1024):
            w.write_block(recs[i:i + 1024])
    assert open(a, "rb").read() == open(b, "rb").read()
    back = list(avro.read_records(a))
    assert len(back) == M and back[1]["modelId"] == ids[1]
    assert (back[3]["variances"] is None) == (not with_variance)


def _record_to_sparse_coefficients(has_intercept, rec, feature2global_id):
    """CHECKER (test infrastructure): one decoded BayesianLinearModelAvro record -> (modelId, TrainingResult), the
    per-record conversion the reference's _load_weights does (random_effect_lr_lbfgs_model.py:277-309)."""
    from gdmix_b200.random_effect import TrainingResult
    hi = 1 if has_intercept else 0
    means = rec["means"]
    if hi:
        assert (means[0]["name"], means[0]["term"]) == ("(INTERCEPT)", "")
    theta = [np.float64(m["value"]) for m in means]
    idx = [feature2global_id[(m["name"], m["term"])] for m in means[hi:]]
    var = None
    if rec.get("variances"):
        var = np.array([np.float64(v["value"]) for v in rec["variances"]])
        assert [feature2global_id[(v["name"], v["term"])] for v in rec["variances"][hi:]] == idx
    if feature2global_id is None:   # intercept-only model: one dummy feature
        assert not idx
        theta.append(np.float64(0.0)); idx.append(0)
    return rec["modelId"], TrainingResult(np.array(theta), var, np.array(idx, dtype=np.int64))


@pytest.mark.parametrize("has_intercept", [True, False])
@pytest.mark.parametrize("with_variance", [True, False])
@pytest.mark.parametrize("codec", ["null", "deflate"])
def test_native_model_reader_matches_the_python_conversion(tmp_path, has_intercept, with_variance, codec):
    """RandomEffectLRLBFGSModel._load_weights (gdmix_avro_model_decode per container block) against
    a per-record Python conversion (the checker above) over avro.read_records, on files written by the Python writer
    (both codecs) -- same entity ids, coefficients, variances, feature indices, dtypes."""
    from types import SimpleNamespace
    from gdmix_b200.io import avro, model_io
    from gdmix_b200.random_effect import RandomEffectLRLBFGSModel as M
    rng = np.random.default_rng(11)
    D = 40
    ff = tmp_path / "features.csv"
    ff.write_text("".join(f"f{j},{'t' if j % 2 else ''}\n" for j in range(D)), encoding="utf-8")
    feature_list = model_io.read_feature_list(str(ff))
    recs = []
    for m in range(2300):
        d = int(rng.integers(0, 7))
        gi = np.sort(rng.choice(D, d, replace=False))
        w, v = rng.standard_normal(d), rng.uniform(0.1, 2, d)
        bias = (float(rng.standard_normal()), float(rng.uniform(0.1, 2)))
        recs.append(model_io.gen_one_avro_model(f"é{m}" if m % 3 else str(m), model_io.LOGISTIC_MODEL_CLASS, gi,
                                                (w, v) if with_variance else w,
                                                (bias if with_variance else bias[0]) if has_intercept else None,
                                                feature_list, 0.0))
    path = str(tmp_path / "models.avro")
    avro.write_records(path, model_io.BAYESIAN_LINEAR_MODEL_SCHEMA, recs, codec=codec)
    fake = SimpleNamespace(feature_file=str(ff), has_intercept=has_intercept)
    got = M._load_weights_native(fake, path)
    fmap = model_io.get_feature_map(str(ff))
    want = dict(_record_to_sparse_coefficients(has_intercept, r, fmap) for r in avro.read_records(path))
    assert list(got.keys()) == list(want.keys())
    for k in want:
        np.testing.assert_array_equal(got[k].theta, want[k].theta)
        np.testing.assert_array_equal(got[k].unique_global_indices, want[k].unique_global_indices)
        assert got[k].unique_global_indices.dtype == want[k].unique_global_indices.dtype
        assert (got[k].variance is None) == (want[k].variance is None)
        if want[k].variance is not None:
            np.testing.assert_array_equal(got[k].variance, want[k].variance)
    # a feature the feature file does not know is an error, as the dict lookup of the reference is
    bad = dict(recs[0]); bad["means"] = list(bad["means"]) + [{"name": "nope", "term": "", "value": 1.0}]
    if with_variance:
        bad["variances"] = list(bad["variances"]) + [{"name": "nope", "term": "", "value": 1.0}]
    avro.write_records(str(tmp_path / "bad.avro"), model_io.BAYESIAN_LINEAR_MODEL_SCHEMA, [bad])
    with pytest.raises(KeyError):
        M._load_weights_native(fake, str(tmp_path / "bad.avro"))


@pytest.mark.parametrize("has_intercept", [True, False])
def test_vectorised_warm_start_equals_the_per_entity_loop(has_intercept):
    """ingest.warm_start_theta (one sorted merge over (entity, feature) keys) against the per-entity search that
    mirrors job_consumers.py:262-288: prior features absent now are dropped, new features start at 0, entities
    without a prior keep zeros, unsorted and repeated prior indices included."""
    from collections import namedtuple
    from types import SimpleNamespace
    TR = namedtuple("TR", "theta variance unique_global_indices")
    rng = np.random.default_rng(5)
    E, D = 400, 60
    hi = 1 if has_intercept else 0
    d_e = rng.integers(0, 12, E)
    uniq_ptr = np.concatenate([[0], np.cumsum(d_e)]).astype(np.int64)
    uniq_global = np.concatenate([np.sort(rng.choice(D, d, replace=False)) for d in d_e]).astype(np.int64)
    theta_ptr = np.concatenate([[0], np.cumsum(d_e + hi)]).astype(np.int64)
    hb = SimpleNamespace(n_coef=int(theta_ptr[-1]), theta_ptr=theta_ptr)
    ids = [f"e{e}" for e in range(E)]
    prior = {}
    for e in range(0, E, 2):
        k = int(rng.integers(0, 15))
        pidx = rng.integers(0, D + 20, k)                      # unsorted, repeats, ids the current data lacks
        prior[ids[e]] = TR(rng.standard_normal(k + hi), None, pidx.astype(np.int64))
    prior["someone else"] = TR(rng.standard_normal(3 + hi), None, np.array([1, 2, 3]))
    a = ingest.warm_start_theta(hb, uniq_ptr, uniq_global, ids, prior, has_intercept)
    b = ingest._warm_start_theta_per_entity(hb, uniq_ptr, uniq_global, ids, prior, has_intercept)
    np.testing.assert_array_equal(a[0], b[0])
    np.testing.assert_array_equal(a[1], b[1])
    assert a[1].sum() == E // 2 and np.count_nonzero(a[0]) > 0
    z = ingest.warm_start_theta(hb, uniq_ptr, uniq_global, ids, {}, has_intercept)
    assert not z[0].any() and not z[1].any()


def test_model_file_with_another_schema_is_refused(tmp_path):
    """The block decoder is laid out for BayesianLinearModelAvro as Photon-ML / GDMix write it: a file whose writer
    schema orders its fields differently must be refused, not decoded as garbage (fastavro would follow the schema)."""
    import copy
    from types import SimpleNamespace
    from gdmix_b200.io import avro, model_io
    from gdmix_b200.random_effect import RandomEffectLRLBFGSModel as M
    schema = copy.deepcopy(model_io.BAYESIAN_LINEAR_MODEL_SCHEMA)
    schema["fields"][0], schema["fields"][1] = schema["fields"][1], schema["fields"][0]
    rec = {"modelId": "7", "modelClass": None, "means": [{"name": "(INTERCEPT)", "term": "", "value": 1.0}],
           "variances": None, "lossFunction": None}
    path = str(tmp_path / "m.avro")
    avro.write_records(path, schema, [rec])
    ff = tmp_path / "f.csv"
    ff.write_text("a,\n", encoding="utf-8")
    with pytest.raises(ValueError, match="writer schema"):
        M._load_weights_native(SimpleNamespace(feature_file=str(ff), has_intercept=True), path)
    model_io.check_model_schema(model_io.BAYESIAN_LINEAR_MODEL_SCHEMA)   # the package's own schema passes


def test_saving_a_prior_model_without_variances_under_a_variance_mode_raises(tmp_path):
    """A model loaded from a file written without variances cannot be re-saved with variances: the reference fails
    (TypeError) rather than writing misaligned arrays; so does _save_model."""
    from types import SimpleNamespace
    from gdmix_b200.random_effect import RandomEffectLRLBFGSModel as M, TrainingResult
    ff = tmp_path / "f.csv"
    ff.write_text("a,\nb,\n", encoding="utf-8")
    fake = SimpleNamespace(model_params=SimpleNamespace(random_effect_variance_mode="SIMPLE", sparsity_threshold=1e-4),
                           has_intercept=True)
    coeffs = {"1": TrainingResult(np.array([0.5, 1.0]), np.array([0.1, 0.2]), np.array([0])),
              "2": TrainingResult(np.array([0.5, 1.0]), None, np.array([1]))}
    with pytest.raises(TypeError, match="no variances"):
        M._save_model(fake, str(tmp_path / "out.avro"), coeffs, 2, str(ff))
    bad = {"1": TrainingResult(np.array([0.5, 1.0]), np.array([0.1]), np.array([0]))}
    with pytest.raises(ValueError, match="variances for"):
        M._save_model(fake, str(tmp_path / "out.avro"), bad, 2, str(ff))


@pytest.mark.parametrize("string_ids", [False, True])
@pytest.mark.parametrize("label_as_int", [True, False])
def test_native_partition_writer_matches_the_python_encoder(tmp_path, string_ids, label_as_int):
    """gdmix_seqex_encode (the writer of DataPartitioner's per-entity TFRecord files) against the record-at-a-time
    Python encoder: identical bytes -- framing, masked crc32c, protobuf -- for int64 / string entity ids, int64 / float
    labels, empty samples, an entity without samples' worth of features, optional columns absent; and the native reader
    gives the arrays back."""
    from gdmix_b200.io import tfrecord as T
    rng = np.random.default_rng(5)
    E = 300
    ent_rows = rng.integers(1, 9, E)
    N = int(ent_rows.sum())
    row_len = rng.integers(0, 7, N)
    nnz = int(row_len.sum())
    gcol = rng.integers(0, 1 << 40, nnz)
    val = rng.standard_normal(nnz).astype(np.float32)
    uid = rng.integers(0, 1 << 50, N)
    label = rng.integers(0, 2, N).astype(np.float32)
    off = rng.standard_normal(N).astype(np.float32)
    ids = [f"m\u00e9{e}" for e in range(E)] if string_ids else [int(x) for x in rng.integers(0, 1 << 45, E)]
    for with_w in (True, False):
        w = rng.uniform(0.5, 2, N).astype(np.float32) if with_w else None
        img = capi.encode_entity_grouped(ent_rows, row_len, gcol, val, uid, entity_int=None if string_ids else ids,
                                         entity_str=ids if string_ids else None, label=label, label_as_int=label_as_int,
                                         offset=off, weight=w, entity="memberId", bag="per_member").tobytes()
        wr = T.TFRecordWriter(str(tmp_path / "x.tfrecord"))
        r = q = 0
        for e, n in enumerate(ent_rows):
            ctx = {"memberId": T.encode_feature([ids[e]], "bytes" if string_ids else "int64"),
                   "uid": T.encode_feature([int(x) for x in uid[r:r + n]], "int64"),
                   "response": T.encode_feature([int(x) for x in label[r:r + n]], "int64") if label_as_int
                   else T.encode_feature([float(x) for x in label[r:r + n]], "float"),
                   "offset": T.encode_feature([float(x) for x in off[r:r + n]], "float")}
            if with_w:
                ctx["weight"] = T.encode_feature([float(x) for x in w[r:r + n]], "float")
            fi, fv = [], []
            for i in range(n):
                k = row_len[r + i]
                fi.append(T.encode_feature([int(x) for x in gcol[q:q + k]], "int64"))
                fv.append(T.encode_feature([float(x) for x in val[q:q + k]], "float"))
                q += k
            wr.write(T.encode_sequence_example(ctx, {"per_member_indices": fi, "per_member_values": fv}))
            r += n
        assert b"".join(wr.chunks) == img
        assert len(list(T.read_records_from_bytes(img, verify_crc=True))) == E if hasattr(T, "read_records_from_bytes") else True
        d = capi.parse_entity_grouped(img, "memberId", "uid", "response", "offset", "weight" if with_w else None,
                                      "per_member_indices", "per_member_values")
        assert d["entity_ids"] == [str(x) for x in ids]
        np.testing.assert_array_equal(d["ent_rows"], ent_rows)
        np.testing.assert_array_equal(d["row_len"], row_len)
        np.testing.assert_array_equal(d["gcol"], gcol)
        np.testing.assert_array_equal(d["val"], val)
        np.testing.assert_array_equal(d["uid"], uid)
        np.testing.assert_array_equal(d["label"], label)
        np.testing.assert_array_equal(d["offset"], off)
    # no feature bag at all (intercept-only data)
    img = capi.encode_entity_grouped(ent_rows, None, None, None, uid, entity_int=None if string_ids else ids,
                                     entity_str=ids if string_ids else None, label=label, bag=None).tobytes()
    path = tmp_path / "nobag.tfrecord"
    path.write_bytes(img)
    assert sum(1 for _ in T.read_records(str(path), verify_crc=True)) == E


def _partition_file(tmp_path, rng, E, max_rows, max_len, id_range, name="part-00000.tfrecord", wide_entity=None):
    ent_rows = rng.integers(1, max_rows + 1, E)
    N = int(ent_rows.sum())
    row_len = rng.integers(0, max_len + 1, N)
    if wide_entity is not None:            # (entity, distinct features): that entity gets one long row
        e, width = wide_entity
        row_len[int(ent_rows[:e].sum())] = width
    nnz = int(row_len.sum())
    gcol = rng.integers(0, id_range, nnz)
    if wide_entity is not None:
        r = int(ent_rows[:wide_entity[0]].sum())
        q = int(row_len[:r].sum())
        gcol[q:q + wide_entity[1]] = rng.permutation(id_range)[:wide_entity[1]]
    val = rng.standard_normal(nnz).astype(np.float32)
    img = capi.encode_entity_grouped(ent_rows, row_len, gcol, val, np.arange(N, dtype=np.int64),
                                     entity_int=np.arange(E, dtype=np.int64), label=rng.integers(0, 2, N).astype(np.float32),
                                     offset=rng.standard_normal(N).astype(np.float32), entity="ent", bag="bag",
                                     label_name="y", offset_name="off")
    with open(str(tmp_path / name), "wb") as f:
        f.write(img.tobytes())
    return ent_rows, row_len, gcol, val


@pytest.mark.parametrize("shape", [(200, 6, 12, 50), (60, 40, 30, 5000), (30, 3, 0, 10)])
def test_fused_reader_equals_reader_plus_local_index(tmp_path, shape):
    """gdmix_seqex_fill_local (entity-local ranks while parsing) against gdmix_seqex_fill + gdmix_local_index_host and
    against np.unique per entity (job_consumers.py:243): same local indices, same distinct features, same batch."""
    E, max_rows, max_len, id_range = shape
    rng = np.random.default_rng(E)
    for k in range(3):
        _partition_file(tmp_path, rng, E, max_rows, max_len, id_range, name=f"part-{k:05d}.tfrecord")
    files = ingest.list_tfrecord_files(str(tmp_path), 1, 0)
    args = (files, "ent", "bag", "y", "off", "w", "uid", id_range, str(tmp_path))
    fused = ingest._read_entity_grouped_native(*args, fused=True)
    plain = ingest._read_entity_grouped_native(*args, fused=False)
    assert fused.local16 is not None and plain.local16 is None
    hb_f, up_f, ug_f = ingest.to_local_batch(fused)
    hb_p, up_p, ug_p = ingest.to_local_batch(plain)
    np.testing.assert_array_equal(up_f, up_p)
    np.testing.assert_array_equal(ug_f, ug_p)
    np.testing.assert_array_equal(hb_f.theta_ptr, hb_p.theta_ptr)
    np.testing.assert_array_equal(hb_f._col_narrow.astype(np.int32), hb_p.col[:hb_p.nnz])
    np.testing.assert_array_equal(fused.gcol, plain.gcol)          # rebuilt from ranks + distinct features
    for name in ("ent_rowptr", "rowptr", "val", "label", "offset", "weight", "uid"):
        np.testing.assert_array_equal(getattr(fused, name), getattr(plain, name))
    assert fused.entity_ids == plain.entity_ids
    for e in range(0, 3 * E, 7):
        q0, q1 = plain.rowptr[plain.ent_rowptr[e]], plain.rowptr[plain.ent_rowptr[e + 1]]
        u, inv = np.unique(plain.gcol[q0:q1], return_inverse=True)
        np.testing.assert_array_equal(u, ug_f[up_f[e]:up_f[e + 1]])
        np.testing.assert_array_equal(inv, hb_f._col_narrow[q0:q1])
    cb = hb_f.c_struct()
    assert cb.col16 is not None and cb.col is None and cb.col8 is None


def test_fused_reader_falls_back_beyond_65535_distinct_features(tmp_path):
    """An entity with more distinct features than 16-bit local indices can name: the reader goes back to int64
    columns + gdmix_local_index_host, and an out-of-range index is still refused."""
    rng = np.random.default_rng(2)
    _partition_file(tmp_path, rng, 5, 2, 4, 200000, wide_entity=(3, 70000))
    kw = dict(metadata=_Meta(["ent"]), entity_name="ent", feature_bag="bag", label_column="y", offset_column="off",
              weight_column="w", uid_column="uid", num_features=200000)
    d = ingest.read_entity_grouped(str(tmp_path), **kw)
    assert d.local16 is None and d.n_entities == 5
    hb, up, ug = ingest.to_local_batch(d)
    assert int(np.diff(up).max()) >= 70000
    with pytest.raises(ValueError):
        ingest.read_entity_grouped(str(tmp_path), **dict(kw, num_features=1000))


def test_model_writer_takes_the_readers_id_table(tmp_path):
    """avro_model_blocks with the ids as a (characters, offsets) table -- what the reader hands on -- writes the bytes
    it writes from the list of strings; the reader's table matches its own list of ids (ASCII and UTF-8 ids)."""
    rng = np.random.default_rng(4)
    for ids in ([str(x) for x in rng.integers(0, 1 << 40, 50)], [f"mémbre-{k}" for k in range(50)]):
        coef_ptr = np.arange(51, dtype=np.int64) * 4
        coef = rng.standard_normal(200)
        feat_idx = rng.integers(0, 30, 150).astype(np.int64)
        names, terms = [f"f{j}" for j in range(30)], [""] * 30
        sync = bytes(range(16))
        a = capi.avro_model_blocks(ids, coef, None, coef_ptr, feat_idx, True, 1e-4, names, terms, "cls", "(INTERCEPT)", sync)
        table = capi._string_table(ids)
        b = capi.avro_model_blocks(ids, coef, None, coef_ptr, feat_idx, True, 1e-4, names, terms, "cls", "(INTERCEPT)", sync,
                                   id_table=table)
        assert bytes(a) == bytes(b)
    # the reader's table
    _partition_file(tmp_path, rng, 40, 3, 5, 100)
    files = ingest.list_tfrecord_files(str(tmp_path), 1, 0)
    d = ingest._read_entity_grouped_native(files, "ent", "bag", "y", "off", "w", "uid", 100, str(tmp_path))
    chars, ptr = d.entity_id_table
    raw = chars.tobytes()
    assert [raw[ptr[e]:ptr[e + 1]].decode("utf-8") for e in range(d.n_entities)] == d.entity_ids
