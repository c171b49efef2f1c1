"""TF-free TFRecord / Example parser, Avro container codec, model-file helpers and CLI parameter groups.

Pins:
  * tests/golden/ref_fixtures/{re_data,fe_test}.tfrecord were written by TensorFlow for the reference's own tests;
    expected_tfrecord.json is their decode by google.protobuf (oracle/gen_io_golden.py), independent of our parser.
  * the record literals of test_io_utils.py:86-188 (gen_one_avro_model with threshold / variance).
  * the parameter behaviour the reference tests rely on (test_params.py, test_random_effect_lr_lbfgs_model.py:54-56).
"""
import json
import os

import numpy as np
import pytest

from gdmix_b200 import params as P
from gdmix_b200.io import avro, model_io, tfrecord
from gdmix_b200.io.dataset_metadata import DatasetMetadata

FIX = os.path.join(os.path.dirname(__file__), "golden", "ref_fixtures")
EXPECTED = json.load(open(os.path.join(FIX, "expected_tfrecord.json")))


def _same_feature(got, exp):
    kind, values = got
    if exp["kind"] is None:
        assert values is None or len(values) == 0
        return
    assert kind == exp["kind"]
    if kind == "bytes":
        assert [bytes(v).decode("latin-1") for v in values] == exp["values"]
    elif kind == "float":
        np.testing.assert_array_equal(np.asarray(values, np.float32), np.asarray(exp["values"], np.float32))
    else:
        np.testing.assert_array_equal(np.asarray(values, np.int64), np.asarray(exp["values"], np.int64))


def test_sequence_example_file_written_by_tensorflow():
    recs = list(tfrecord.read_records(os.path.join(FIX, "re_data.tfrecord"), verify_crc=True))
    assert len(recs) == len(EXPECTED["re_data.tfrecord"]) == 2
    for payload, exp in zip(recs, EXPECTED["re_data.tfrecord"]):
        ctx, lists = tfrecord.parse_sequence_example(payload)
        assert sorted(ctx) == sorted(exp["context"])
        for k, v in exp["context"].items():
            _same_feature(ctx[k], v)
        assert sorted(lists) == sorted(exp["feature_lists"])
        for k, feats in exp["feature_lists"].items():
            assert len(lists[k]) == len(feats)
            for g, x in zip(lists[k], feats):
                _same_feature(g, x)


def test_example_file_written_by_tensorflow():
    recs = list(tfrecord.read_records(os.path.join(FIX, "fe_test.tfrecord"), verify_crc=True))
    assert len(recs) == len(EXPECTED["fe_test.tfrecord"]) == 32
    for payload, exp in zip(recs, EXPECTED["fe_test.tfrecord"]):
        ex = tfrecord.parse_example(payload)
        assert sorted(ex) == sorted(exp)
        for k, v in exp.items():
            _same_feature(ex[k], v)


def test_crc32c_known_answers():
    # RFC 3720 B.4 test vectors
    assert tfrecord.crc32c(b"\x00" * 32) == 0x8A9136AA
    assert tfrecord.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert tfrecord.crc32c(bytes(range(32))) == 0x46DD794E
    assert tfrecord.crc32c(b"123456789") == 0xE3069283


@pytest.mark.parametrize("suffix", ["", ".gz", ".deflate"])
def test_tfrecord_round_trip_all_compressions(tmp_path, suffix):
    fn = str(tmp_path / f"part-0.tfrecord{suffix}")
    payloads = [tfrecord.encode_sequence_example(
        {"memberId": [1000 + i], "uid": [7 * i, 7 * i + 1], "offset": [0.5, -1.25], "response": [0, 1],
         "name": [b"abc"]},
        {"bag_indices": [[0, 3, 9], []], "bag_values": [[1.0, -2.0, 3.5], []]}) for i in range(5)]
    with tfrecord.TFRecordWriter(fn) as w:
        for p in payloads:
            w.write(p)
    back = list(tfrecord.read_records(fn, verify_crc=True))
    assert back == payloads
    ctx, lists = tfrecord.parse_sequence_example(back[3])
    assert ctx["memberId"] == ("int64", [1003]) or list(ctx["memberId"][1]) == [1003]
    np.testing.assert_array_equal(np.asarray(ctx["offset"][1], np.float32), np.float32([0.5, -1.25]))
    assert [len(f[1]) if f[1] is not None else 0 for f in lists["bag_indices"]] == [3, 0]


def test_tfrecord_detects_corruption(tmp_path):
    fn = str(tmp_path / "x.tfrecord")
    with tfrecord.TFRecordWriter(fn) as w:
        w.write(tfrecord.encode_example({"a": [1, 2, 3]}))
    raw = bytearray(open(fn, "rb").read())
    raw[3] ^= 0x40
    open(fn, "wb").write(bytes(raw))
    with pytest.raises(ValueError):
        list(tfrecord.read_records(fn))


def test_negative_int64_and_large_values_round_trip():
    ex = tfrecord.parse_example(tfrecord.encode_example({"v": [-1, -2 ** 63, 2 ** 63 - 1, 0]}))
    np.testing.assert_array_equal(np.asarray(ex["v"][1], np.int64), np.array([-1, -2 ** 63, 2 ** 63 - 1, 0]))


# ---- Avro ---------------------------------------------------------------------------------------------------

def test_reference_avro_fixture_decodes():
    schema, records = avro.read_container(os.path.join(FIX, "validate_data.avro"))
    records = list(records)
    assert schema["type"] == "record" and len(records) > 0
    names = [f["name"] for f in schema["fields"]]
    assert all(set(r) == set(names) for r in records)


@pytest.mark.parametrize("codec", ["null", "deflate"])
def test_model_file_round_trip(tmp_path, codec):
    recs = [model_io.gen_one_avro_model(str(i), model_io.LOGISTIC_MODEL_CLASS, np.arange(3),
                                        (np.array([1.5, -2.5, 1e-9 * i]), np.array([0.1, 0.2, 0.3])), (0.25 * i, 2.0),
                                        [("f1", "t1"), ("f2", ""), ("f3", "t,3")], 1e-4) for i in range(2500)]
    fn = str(tmp_path / "m.avro")
    assert avro.write_records(fn, model_io.BAYESIAN_LINEAR_MODEL_SCHEMA, recs, codec=codec) == 2500
    back = avro.read_records(fn)
    assert len(back) == 2500
    assert back[7]["modelId"] == "7" and back[7]["means"][0] == {"name": "(INTERCEPT)", "term": "", "value": 1.75}
    assert [m["name"] for m in back[7]["means"]] == ["(INTERCEPT)", "f1", "f2"]  # 7e-9 <= 1e-4 dropped
    assert [v["value"] for v in back[7]["variances"]] == [2.0, 0.1, 0.2]


def test_gen_one_avro_model_reference_literals():
    """test_io_utils.py:86-188."""
    cls = model_io.LOGISTIC_MODEL_CLASS
    fl = [("f1,2", "t1"), ("f2", ""), ("f3", "t3,3")]
    rec = model_io.gen_one_avro_model("1234", cls, np.arange(3), np.array([[1.2, 3.4, 5.6]]), 7.8, fl, 0.0)
    exp = {"modelId": "1234", "modelClass": cls, "means": [
        {"name": "(INTERCEPT)", "term": "", "value": 7.8}, {"name": "f1,2", "term": "t1", "value": 1.2},
        {"name": "f2", "term": "", "value": 3.4}, {"name": "f3", "term": "t3,3", "value": 5.6}], "lossFunction": ""}
    assert rec == exp
    rec = model_io.gen_one_avro_model("1234", cls, np.arange(3), np.array([[1.2, 3.4, 5.6]]), None, fl, 0.0)
    assert rec["means"] == exp["means"][1:]
    rec = model_io.gen_one_avro_model("1234", cls, np.arange(3), np.array([[1.2, 3.4, -5.6]]), 0.8, fl, 3.4)
    assert rec["means"] == [{"name": "(INTERCEPT)", "term": "", "value": 0.8},
                            {"name": "f3", "term": "t3,3", "value": -5.6}]
    rec = model_io.gen_one_avro_model("1234", cls, np.arange(3),
                                      (np.array([[1.2, 3.4, 5.6]]), np.array([[7.8, 9.0, 10.1]])), (-7.8, 1.2), fl, 0.0)
    assert [v["value"] for v in rec["variances"]] == [1.2, 7.8, 9.0, 10.1]
    assert [v["value"] for v in rec["means"]] == [-7.8, 1.2, 3.4, 5.6]


def test_export_and_load_fixed_effect_model(tmp_path):
    """export_linear_model_to_avro + load_linear_models_from_avro: intercept first in the file, LAST when loaded
    (io_utils.py:45-83)."""
    ff = str(tmp_path / "features.csv")
    open(ff, "w").write("f1,t1\nf2,\nf3,t3\n")
    out = str(tmp_path / "model" / "part-00000.avro")
    w = np.array([0.5, 0.0, -2.0])
    model_io.export_linear_model_to_avro(["global model"], [np.arange(3)], [w], [1.25], ff, out)
    (m,) = model_io.load_linear_models_from_avro(out, ff)
    np.testing.assert_array_equal(m, [0.5, 0.0, -2.0, 1.25])
    assert model_io.get_feature_map(ff) == {("f1", "t1"): 0, ("f2", ""): 1, ("f3", "t3"): 2}
    short = str(tmp_path / "short.csv")
    open(short, "w").write("f3,t3\n")
    (m,) = model_io.load_linear_models_from_avro(out, short)
    np.testing.assert_array_equal(m, [-2.0, 1.25])


def test_score_schema_and_records(tmp_path):
    sp = P.SchemaParams(uid_column_name="uid", weight_column_name="weight", label_column_name="response",
                        prediction_score_column_name="predictionScore")
    schema = model_io.get_inference_output_avro_schema({}, True, sp, has_weight=True)
    assert [f["name"] for f in schema["fields"]] == ["uid", "predictionScore", "response", "weight",
                                                     "predictionScorePerCoordinate"]
    recs = [{"uid": 10 ** 12 + i, "predictionScore": np.float32(0.1 * i), "response": 1.0, "weight": 2.0,
             "predictionScorePerCoordinate": -0.5, "extra": 1} for i in range(3)]
    fn = str(tmp_path / "s.avro")
    model_io.batched_write_avro(recs, fn, schema)
    back = avro.read_records(fn)
    assert [r["uid"] for r in back] == [10 ** 12, 10 ** 12 + 1, 10 ** 12 + 2]
    assert back[2]["predictionScore"] == float(np.float32(0.2)) and "extra" not in back[0]


# ---- metadata, parameters -------------------------------------------------------------------------------------

def test_dataset_metadata_reference_fixture():
    md = DatasetMetadata(os.path.join(FIX, "re_data.json"))
    assert md.get_feature_names() == ["per_member", "weight", "offset", "uid", "memberId"]
    assert md.get_label_names() == ["response"]
    assert md.get_feature_shape("per_member") == [100]
    with pytest.raises(ValueError):
        DatasetMetadata({"features": [{"name": "a", "dtype": "float", "shape": []},
                                      {"name": "a", "dtype": "float", "shape": []}]})
    with pytest.raises(ValueError):
        DatasetMetadata({"features": [{"name": "a", "dtype": "complex", "shape": []}]})


def test_params_parse_like_smart_arg():
    argv = ["--uid_column_name", "uid", "--weight_column_name=weight", "--label_column_name", "response",
            "--action", "train", "--stage", "random_effect", "--metadata_file", "m.json", "--output_model_dir", "out",
            "--feature_bag", "per_member", "--regularize_bias", "False", "--l2_reg_weight", "0.1",
            "--enable_local_indexing", "True", "--partition_entity", "memberId", "--__frozen__", "True",
            "--unknown_flag", "7"]
    p = P.Params.from_argv(argv)
    assert (p.stage, p.action, p.model_type) == ("random_effect", "train", "logistic_regression")
    re = P.REParams.__from_argv__(argv, error_on_unknown=False)
    assert re.regularize_bias is False and re.l2_reg_weight == 0.1 and re.enable_local_indexing is True
    assert re.num_of_lbfgs_curvature_pairs == 10 and re.lbfgs_tolerance == 1e-12 and re.sparsity_threshold == 1e-4
    assert re.random_effect_variance_mode is None and re.partition_entity == "memberId"
    with pytest.raises(ValueError):
        P.REParams.from_argv(argv, error_on_unknown=True)
    with pytest.raises(AssertionError):
        P.REParams.from_argv(argv + ["--num_of_consumers", "10"])       # queue size must exceed consumers
    with pytest.raises(AssertionError):
        P.REParams.from_argv(argv + ["--random_effect_variance_mode", "bogus"])
    with pytest.raises(AssertionError):
        P.Params.from_argv(["--uid_column_name", "uid", "--action", "train"])  # train needs a label column
    fe = P.FixedLRParams.from_argv(argv + ["--fixed_effect_variance_mode", "simple"])
    assert fe.fixed_effect_variance_mode == "simple" and fe.copy_to_local is True
    rt = P.REParams.from_argv(re.to_argv())
    assert rt == re
