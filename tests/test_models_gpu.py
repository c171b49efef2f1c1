"""The plugin classes end to end on the GPU: TFRecords in, Avro model + score files out.

Modelled on the reference's own tests (gdmix-trainer/test/models/custom/test_random_effect_lr_lbfgs_model.py and
test_fixed_effect_lr_lbfgs_model.py): same fixture (tests/golden/ref_fixtures/re_data.tfrecord is the reference's
grouped_per_member_train/data.tfrecord), same parameters, same assertions -- plus coefficient parity with what
the reference's solver produced for those entities (tests/golden/re_golden.*, cases "fixture:*")."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
if not torch.cuda.is_available():  # pragma: no cover
    pytest.skip("needs a CUDA device", allow_module_level=True)

from gdmix_b200 import RandomEffectLRLBFGSModel, FixedEffectLRModelLBFGS, FixedEffectLRLBFGSModel  # noqa: E402
from gdmix_b200 import constants, gdmix as cli  # noqa: E402
from gdmix_b200.io import avro, tfrecord  # noqa: E402
from gdmix_b200.params import Params, SchemaParams  # noqa: E402
from gdmix_b200.synthetic import make_batch  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tests.golden_util import load_re  # noqa: E402

FIX = os.path.join(os.path.dirname(__file__), "golden", "ref_fixtures")
ARR, CASES = load_re()
BY_NAME = {c["name"]: c for c in CASES}


def schema_params():
    return SchemaParams(uid_column_name="uid", weight_column_name="weight", label_column_name="response",
                        prediction_score_column_name="predictionScore")


def re_dataset(tmp_path):
    """The reference fixture laid out as a directory of TFRecords."""
    d = tmp_path / "data"
    d.mkdir(exist_ok=True)
    (d / "data.tfrecord").write_bytes(open(os.path.join(FIX, "re_data.tfrecord"), "rb").read())
    return str(d)


def re_params(tmp_path, out_dir, extra=(), feature_bag=True, has_intercept=True, l2="0.1"):
    p = ["--uid_column_name", "uid", "--weight_column_name", "weight", "--label_column_name", "response",
         "--metadata_file", os.path.join(FIX, "re_data.json"), "--output_model_dir", out_dir,
         "--offset_column_name", "offset", "--partition_entity", "memberId", "--l2_reg_weight", l2,
         "--feature_file", os.path.join(FIX, "re_feature_file.csv"), "--batch_size", "2",
         "--has_intercept", str(has_intercept)]
    if feature_bag:
        p += ["--feature_bag", "per_member"]
    if not has_intercept:
        p += ["--regularize_bias", "False"]
    return p + list(extra)


def _model_by_id(path):
    return {r["modelId"]: r for r in avro.read_records(path)}


def test_random_effect_train_and_predict(tmp_path):
    """test_train_and_predict (:98-167) + parity with the reference solver's coefficients."""
    data = re_dataset(tmp_path)
    out = str(tmp_path / "model")
    model = RandomEffectLRLBFGSModel(raw_model_params=re_params(tmp_path, out, ["--enable_local_indexing", "True"]))
    active, passive, valid = (str(tmp_path / n) for n in ("active.avro", "passive.avro", "valid.avro"))
    ctx = {constants.ACTIVE_TRAINING_OUTPUT_FILE: active, constants.PASSIVE_TRAINING_OUTPUT_FILE: passive,
           constants.VALIDATION_OUTPUT_FILE: valid, constants.PARTITION_INDEX: 0,
           constants.PASSIVE_TRAINING_DATA_DIR: data}
    model.train(training_data_dir=data, validation_data_dir=data, metadata_file=os.path.join(FIX, "re_data.json"),
                checkpoint_path=str(tmp_path / "ckpt"), execution_context=ctx, schema_params=schema_params())
    models = _model_by_id(os.path.join(out, "part-00000.avro"))
    assert set(models) == {"100034", "100"}
    # coefficients == what the reference's own trainer produced (scipy 1.18.1), thresholded at 1e-4
    feats = [tuple(l.strip().split(",")) for l in open(os.path.join(FIX, "re_feature_file.csv"))]
    for ent in ("100034", "100"):
        c = BY_NAME[f"fixture:{ent}/l2=0.1/maxiter=100"]
        ref = ARR[c["key"] + "_theta"]
        rec = next(r for r in json.load(open(os.path.join(FIX, "expected_tfrecord.json")))["re_data.tfrecord"]
                   if str(r["context"]["memberId"]["values"][0]) == ent)
        gidx = sorted({j for f in rec["feature_lists"]["per_member_indices"] for j in f["values"]})
        assert len(gidx) == len(ref) - 1
        means = models[ent]["means"]
        assert means[0]["name"] == "(INTERCEPT)" and abs(means[0]["value"] - ref[0]) <= 1e-7 * abs(ref[0])
        kept = [(j, v) for j, v in enumerate(ref[1:]) if abs(v) > 1e-4]
        assert len(means) - 1 == len(kept)
        for m, (j, v) in zip(means[1:], kept):
            assert abs(m["value"] - v) <= 1e-7 * abs(v)
            if gidx is not None:
                assert (m["name"], m["term"]) == feats[int(gidx[j])]
    for f in (active, passive, valid):
        recs = avro.read_records(f)
        assert len(recs) == 3 and all(isinstance(r, dict) for r in recs)
        assert set(recs[0]) == {"uid", "predictionScore", "response", "weight", "predictionScorePerCoordinate"}
    # cold prediction reproduces scoring-while-training (TEST 3)
    pred_dir = str(tmp_path / "pred")
    model.predict(output_dir=pred_dir, input_data_path=data, metadata_file=os.path.join(FIX, "re_data.json"),
                  checkpoint_path=out, execution_context=ctx, schema_params=schema_params())
    assert avro.read_records(os.path.join(pred_dir, "part-00000.avro")) == avro.read_records(active)
    # scores: logit = x.theta + b + offset, per-coordinate = logit - offset
    recs = {r["uid"]: r for r in avro.read_records(active)}
    assert abs((recs[10]["predictionScore"] - recs[10]["predictionScorePerCoordinate"]) - 0.5) < 1e-6
    # TEST 4 / 5: no validation set, no scoring
    model.train(training_data_dir=data, validation_data_dir=None, metadata_file=os.path.join(FIX, "re_data.json"),
                checkpoint_path=str(tmp_path / "ckpt"), execution_context=ctx, schema_params=schema_params())
    quiet = RandomEffectLRLBFGSModel(raw_model_params=re_params(
        tmp_path, str(tmp_path / "m2"), ["--disable_random_effect_scoring_after_training", "True"]))
    ctx2 = dict(ctx, **{constants.ACTIVE_TRAINING_OUTPUT_FILE: str(tmp_path / "never.avro")})
    quiet.train(training_data_dir=data, validation_data_dir=None, metadata_file=os.path.join(FIX, "re_data.json"),
                checkpoint_path=str(tmp_path / "ckpt"), execution_context=ctx2, schema_params=schema_params())
    assert not os.path.exists(str(tmp_path / "never.avro"))


def test_random_effect_fails_on_unknown_partition_entity(tmp_path):
    """test_train_should_fail_if_producer_or_consumer_fails (:74-96)."""
    data = re_dataset(tmp_path)
    p = re_params(tmp_path, str(tmp_path / "m"))
    p[p.index("--partition_entity") + 1] = "fake_partition_entity"
    model = RandomEffectLRLBFGSModel(raw_model_params=p)
    with pytest.raises(Exception):
        model.train(training_data_dir=data, validation_data_dir=data, metadata_file=os.path.join(FIX, "re_data.json"),
                    checkpoint_path=str(tmp_path / "c"), execution_context={constants.PARTITION_INDEX: 0},
                    schema_params=schema_params())


def test_random_effect_warm_start(tmp_path):
    """:231-328 -- one iteration from the previous optimum stays there; one iteration from zero does not."""
    data = re_dataset(tmp_path)
    out = str(tmp_path / "model")
    md = os.path.join(FIX, "re_data.json")
    ctx = {constants.PARTITION_INDEX: 0}
    RandomEffectLRLBFGSModel(raw_model_params=re_params(tmp_path, out)).train(
        data, None, md, str(tmp_path / "c"), ctx, schema_params())
    full = _model_by_id(os.path.join(out, "part-00000.avro"))
    RandomEffectLRLBFGSModel(raw_model_params=re_params(tmp_path, out, ["--num_of_lbfgs_iterations", "1"])).train(
        data, None, md, str(tmp_path / "c"), ctx, schema_params())
    warm = _model_by_id(os.path.join(out, "part-00000.avro"))
    cold_dir = str(tmp_path / "cold")
    RandomEffectLRLBFGSModel(raw_model_params=re_params(tmp_path, cold_dir, ["--num_of_lbfgs_iterations", "1"])).train(
        data, None, md, str(tmp_path / "c"), ctx, schema_params())
    cold = _model_by_id(os.path.join(cold_dir, "part-00000.avro"))
    for ent in full:
        a = np.array([m["value"] for m in full[ent]["means"]])
        b = np.array([m["value"] for m in warm[ent]["means"]])
        c = np.array([m["value"] for m in cold[ent]["means"]])
        np.testing.assert_allclose(b, a, rtol=1e-4, atol=1e-4)
        assert len(c) != len(a) or not np.allclose(c, a, rtol=1e-4, atol=1e-4)


def test_random_effect_intercept_only_and_variance(tmp_path):
    """Intercept-only model (:154-167: theta has length 2 with a zero dummy weight) and SIMPLE variance."""
    data = re_dataset(tmp_path)
    out = str(tmp_path / "model")
    model = RandomEffectLRLBFGSModel(raw_model_params=re_params(
        tmp_path, out, ["--random_effect_variance_mode", "simple"], feature_bag=False))
    model.train(data, None, os.path.join(FIX, "re_data.json"), str(tmp_path / "c"), {constants.PARTITION_INDEX: 0},
                schema_params())
    models = _model_by_id(os.path.join(out, "part-00000.avro"))
    for ent, rec in models.items():
        assert [m["name"] for m in rec["means"]] == ["(INTERCEPT)"]
        assert len(rec["variances"]) == 1 and rec["variances"][0]["value"] > 0
    loaded = model._load_weights(os.path.join(out, "part-00000.avro"))
    assert all(len(v.theta) == 2 and v.theta[1] == 0.0 for v in loaded.values())


def _write_partition(path, hb, gcols_of, ids):
    with tfrecord.TFRecordWriter(path) as w:
        for e in range(hb.n_entities):
            r0, r1 = int(hb.ent_rowptr[e]), int(hb.ent_rowptr[e + 1])
            idx, val = [], []
            for i in range(r0, r1):
                q0, q1 = int(hb.rowptr[i]), int(hb.rowptr[i + 1])
                idx.append(tfrecord.encode_feature(gcols_of(e, hb.col[q0:q1]).tolist(), "int64"))
                val.append(tfrecord.encode_feature(hb.val[q0:q1].tolist(), "float"))
            ctx = {"memberId": tfrecord.encode_feature([ids[e]]),
                   "uid": tfrecord.encode_feature(list(range(r0, r1)), "int64"),
                   "response": tfrecord.encode_feature(hb.label[r0:r1].astype(int).tolist(), "int64"),
                   "offset": tfrecord.encode_feature(hb.offset[r0:r1].tolist(), "float"),
                   "weight": tfrecord.encode_feature(hb.weight[r0:r1].tolist(), "float")}
            w.write(tfrecord.encode_sequence_example(ctx, {"per_member_indices": idx, "per_member_values": val}))


def test_random_effect_partition_matches_oracle(tmp_path):
    """A 300-entity partition with global feature ids and string entity ids through the whole class: every
    entity's coefficients vs the CPU oracle on the same bytes."""
    E, n, d, k, D = 300, 24, 16, 5, 500
    hb = make_batch(E, n, d, k, seed=77, ragged=True, weights=True)
    rng = np.random.default_rng(5)
    gmap = np.stack([np.sort(rng.choice(D, d, replace=False)) for _ in range(E)])
    ids = [f"member-{7 * e}" for e in range(E)]
    data = tmp_path / "part"
    data.mkdir()
    _write_partition(str(data / "part-0.tfrecord.gz"), hb, lambda e, c: gmap[e][c], ids)
    md = {"features": [{"name": "per_member", "dtype": "float", "shape": [D], "isSparse": True},
                       {"name": "weight", "dtype": "float", "shape": [], "isSparse": False},
                       {"name": "offset", "dtype": "float", "shape": [], "isSparse": False},
                       {"name": "uid", "dtype": "long", "shape": [], "isSparse": False},
                       {"name": "memberId", "dtype": "string", "shape": [], "isSparse": False}],
          "labels": [{"name": "response", "dtype": "int", "shape": [], "isSparse": False}]}
    mdf = str(tmp_path / "md.json")
    json.dump(md, open(mdf, "w"))
    ff = str(tmp_path / "features.csv")
    open(ff, "w").write("".join(f"f{j},t\n" for j in range(D)))
    out = str(tmp_path / "model")
    p = ["--uid_column_name", "uid", "--weight_column_name", "weight", "--label_column_name", "response",
         "--metadata_file", mdf, "--output_model_dir", out, "--partition_entity", "memberId", "--feature_bag",
         "per_member", "--feature_file", ff, "--l2_reg_weight", "1.0", "--regularize_bias", "False"]
    model = RandomEffectLRLBFGSModel(raw_model_params=p)
    model.train(str(data), None, mdf, str(tmp_path / "c"), {constants.PARTITION_INDEX: 3}, schema_params())
    models = _model_by_id(os.path.join(out, "part-00003.avro"))
    assert len(models) == E
    oo = O.make_opts(l2=1.0, regularize_bias=False)
    th, f, nit, nfev, st = O.re_fit_batch({"ent_rowptr": hb.ent_rowptr, "rowptr": hb.rowptr, "col": hb.col,
                                           "val": hb.val, "y": hb.label, "w": hb.weight, "off": hb.offset,
                                           "theta_ptr": hb.theta_ptr}, oo)
    np.testing.assert_array_equal(model.last_fit_info["nit"], nit)
    for e in range(0, E, 7):
        ref = th[hb.theta_ptr[e]:hb.theta_ptr[e + 1]]
        # local feature c of the synthetic batch is global id gmap[e][c]; unused local columns never reach the file
        rec = {(m["name"], m["term"]): m["value"] for m in models[ids[e]]["means"]}
        assert abs(rec[("(INTERCEPT)", "")] - ref[0]) <= 1e-5 * abs(ref[0])
        for c in range(d):
            v = ref[1 + c]
            got = rec.get((f"f{gmap[e][c]}", "t"), 0.0)
            if abs(v) > 1e-4:
                assert abs(got - v) <= 1e-5 * abs(v)
            else:
                assert got == 0.0


def fe_params(tmp_path, out, extra=()):
    return ["--uid_column_name", "uid", "--weight_column_name", "weight", "--label_column_name", "response",
            "--metadata_file", os.path.join(FIX, "fe_tensor_metadata.json"), "--output_model_dir", out,
            "--feature_bag", "global", "--feature_file", os.path.join(FIX, "fe_feature_list_global"),
            "--l2_reg_weight", "1.0", "--regularize_bias", "True"] + list(extra)


def fe_base(tmp_path, model_type="logistic_regression"):
    return Params(uid_column_name="uid", weight_column_name="weight", label_column_name="response",
                  prediction_score_column_name="predictionScore", action="train", stage="fixed_effect",
                  model_type=model_type, training_score_dir=str(tmp_path / "train_scores"),
                  validation_score_dir=str(tmp_path / "valid_scores"))


def _fe_fixture_rows():
    exp = json.load(open(os.path.join(FIX, "expected_tfrecord.json")))["fe_test.tfrecord"]
    cols = [np.array(r["global_indices"]["values"], np.int32) for r in exp]
    vals = [np.array(r["global_values"]["values"], np.float32) for r in exp]
    y = np.array([r["response"]["values"][0] for r in exp], np.float32)
    w = np.array([r["weight"]["values"][0] if "weight" in r else 1.0 for r in exp], np.float32)
    off = np.array([r["offset"]["values"][0] if "offset" in r and r["offset"]["values"] else 0.0 for r in exp],
                   np.float32)
    rowptr = np.concatenate([[0], np.cumsum([len(c) for c in cols])]).astype(np.int64)
    return rowptr, np.concatenate(cols), np.concatenate(vals), y, w, off


def test_fixed_effect_train_predict_on_reference_fixture(tmp_path):
    """FixedEffectLRModelLBFGS on the reference's fe_lbfgs fixture: coefficients vs the oracle's replay of the same
    rows, score files, model reload (test_fixed_effect_lr_lbfgs_model.py:379-470 in spirit)."""
    assert FixedEffectLRLBFGSModel is FixedEffectLRModelLBFGS
    data = tmp_path / "fe"
    data.mkdir()
    (data / "test.tfrecord").write_bytes(open(os.path.join(FIX, "fe_test.tfrecord"), "rb").read())
    out = str(tmp_path / "fe_model")
    model = FixedEffectLRModelLBFGS(raw_model_params=fe_params(tmp_path, out, ["--fixed_effect_variance_mode", "simple"]),
                                    base_training_params=fe_base(tmp_path))
    ctx = {constants.TASK_INDEX: 0, constants.NUM_WORKERS: 1, constants.IS_CHIEF: True}
    model.train(str(data), str(data), os.path.join(FIX, "fe_tensor_metadata.json"), out, ctx, schema_params())
    rowptr, col, val, y, w, off = _fe_fixture_rows()
    D = model.num_features
    oo = O.make_opts(l2=1.0, regularize_bias=True)
    x_ref, f_ref, nit_ref, nfev_ref, st_ref = O.fe_fit(O.FeBlock(len(y), D, rowptr, col, val, y, w, off), oo)
    assert model.fit_info["nit"] == nit_ref and model.fit_info["nfev"] == nfev_ref
    np.testing.assert_allclose(model.model_coefficients, np.where(np.abs(x_ref) <= 1e-4, 0.0, x_ref),
                               rtol=1e-7, atol=1e-9)
    # SIMPLE variance: 1 / (sum_i x_ij^2 rho_i (1 - rho_i) w_i + l2 + eps), intercept last
    xm = model.model_coefficients
    z = np.array([val[rowptr[i]:rowptr[i + 1]].astype(np.float64) @ xm[col[rowptr[i]:rowptr[i + 1]]]
                  for i in range(len(y))]) + xm[-1] + off
    rho = 1.0 / (1.0 + np.exp(-z))
    dd = rho * (1 - rho) * w
    H = np.zeros(D + 1)
    for i in range(len(y)):
        sl = slice(rowptr[i], rowptr[i + 1])
        np.add.at(H, col[sl], val[sl].astype(np.float64) ** 2 * dd[i])
    H[-1] = dd.sum()
    np.testing.assert_allclose(model.variances, 1.0 / (H + 1.0 + 1e-12), rtol=1e-9)
    for d_ in ("train_scores", "valid_scores"):
        recs = avro.read_records(str(tmp_path / d_ / "part-00000.avro"))
        assert len(recs) == len(y)
        np.testing.assert_allclose([r["predictionScorePerCoordinate"] for r in recs], (z - off).astype(np.float32),
                                   rtol=2e-6, atol=1e-6)
    # reload + predict
    loaded = model._load_model()
    np.testing.assert_allclose(loaded, model.model_coefficients, rtol=0, atol=0)
    pred = str(tmp_path / "pred")
    model.predict(pred, str(data), os.path.join(FIX, "fe_tensor_metadata.json"), out, ctx, schema_params())
    assert avro.read_records(os.path.join(pred, "part-00000.avro")) == \
        avro.read_records(str(tmp_path / "train_scores" / "part-00000.avro"))


def test_cli_random_effect_stage(tmp_path, monkeypatch):
    """`python -m gdmix.gdmix` flags of a random-effect train job (gdmix-workflow/test/test_workflow_generator.py:
    186-224): partition list, partitionId= directories, per-worker score file names."""
    root = tmp_path / "job"
    for sub in ("train/active/partitionId=0", "train/passive/partitionId=0", "valid/partitionId=0"):
        (root / sub).mkdir(parents=True)
        (root / sub / "data.tfrecord").write_bytes(open(os.path.join(FIX, "re_data.tfrecord"), "rb").read())
    (root / "partition_list.txt").write_text("0,1")
    monkeypatch.delenv("TF_CONFIG", raising=False)
    monkeypatch.delenv("RANK", raising=False)
    args = ["gdmix", "--action=train", "--stage=random_effect", "--model_type=logistic_regression",
            "--uid_column_name=uid", "--weight_column_name=weight", "--label_column_name=response",
            "--prediction_score_column_name=predictionScore",
            "--prediction_score_per_coordinate_column_name=predictionScorePerCoordinate",
            f"--training_score_dir={root}/scores/train", f"--validation_score_dir={root}/scores/valid",
            f"--partition_list_file={root}/partition_list.txt", f"--metadata_file={FIX}/re_data.json",
            f"--output_model_dir={root}/models", f"--training_data_dir={root}/train",
            f"--validation_data_dir={root}/valid", "--feature_bag=per_member",
            f"--feature_file={FIX}/re_feature_file.csv", "--regularize_bias=False", "--l2_reg_weight=0.1",
            "--lbfgs_tolerance=1e-12", "--num_of_lbfgs_curvature_pairs=10", "--num_of_lbfgs_iterations=100",
            "--has_intercept=True", "--offset_column_name=offset", "--batch_size=16", "--data_format=tfrecord",
            "--partition_entity=memberId", "--enable_local_indexing=True", "--max_training_queue_size=10",
            "--training_queue_timeout_in_seconds=300", "--num_of_consumers=2", "--__frozen__=True"]
    cli.run(args)
    # the model file is keyed by partition index inside output_model_dir (random_effect_lr_lbfgs_model.py:113)
    assert os.path.exists(f"{root}/models/part-00000.avro")
    assert not os.path.exists(f"{root}/models/part-00001.avro")   # partition 1 has no data: skipped
    for f in ("scores/train/partitionId=0/part-00000-active.avro", "scores/train/partitionId=0/part-00000-passive.avro",
              "scores/valid/partitionId=0/part-00000.avro"):
        assert len(avro.read_records(f"{root}/{f}")) == 3
