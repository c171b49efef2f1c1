"""Driver / factory logic that sits above the hot path (no GPU): worker identity from TF_CONFIG or RANK,
partition sharding ``partition_list[task::workers]`` (random_effect_driver.py:60-68), ``partitionId=`` anchoring,
score-file names (driver.py:191-216).  Mirrors gdmix-trainer/test/drivers/test_random_effect_driver.py and
test_fixed_effect_driver.py with a recording stand-in for the model."""
import json
import os

import pytest

from gdmix_b200 import constants
from gdmix_b200.drivers import DriverFactory, FixedEffectDriver, RandomEffectDriver
from gdmix_b200.params import Params, SchemaParams


class RecordingModel:
    def __init__(self, root):
        self.checkpoint_path = os.path.join(root, "models")
        self.training_data_dir = os.path.join(root, "train", "active")
        self.passive_training_data_dir = os.path.join(root, "train", "passive")
        self.validation_data_dir = os.path.join(root, "valid")
        self.metadata_file = os.path.join(root, "md.json")
        self.calls = []

    def train(self, **kw):
        self.calls.append(("train", kw))

    def predict(self, **kw):
        self.calls.append(("predict", kw))

    def export(self, output_model_dir):
        self.calls.append(("export", output_model_dir))


def _tf_config(index, n=5, task_type="worker"):
    return json.dumps({"task": {"type": task_type, "index": index},
                       "cluster": {"worker": [f"node{i}:1" for i in range(n)], "evaluator": ["node9:1"]}})


def _params(root, stage, action="train"):
    return Params(uid_column_name="uid", weight_column_name="weight", label_column_name="response",
                  prediction_score_column_name="predictionScore", action=action, stage=stage,
                  training_score_dir=os.path.join(root, "s_train"), validation_score_dir=os.path.join(root, "s_valid"),
                  partition_list_file=os.path.join(root, "partitions.txt"))


@pytest.fixture
def job(tmp_path):
    root = str(tmp_path)
    open(os.path.join(root, "partitions.txt"), "w").write("0,1,2,3,4,5,6,7,8,9,10,11\n")
    for p in (1, 6, 11):
        for sub in ("train/active", "train/passive", "valid"):
            d = os.path.join(root, sub, f"partitionId={p}")
            os.makedirs(d)
            open(os.path.join(d, "part-0.tfrecord"), "w").write("x")
    return root


def test_random_effect_driver_shards_partitions_by_worker(job, monkeypatch):
    monkeypatch.setenv(constants.TF_CONFIG, _tf_config(1))
    model = RecordingModel(job)
    drv = RandomEffectDriver(_params(job, constants.RANDOM_EFFECT), model)
    assert constants.TF_CONFIG not in os.environ          # random effect runs in local mode afterwards
    ctx = drv.execution_context
    assert (ctx[constants.TASK_INDEX], ctx[constants.NUM_WORKERS], ctx[constants.IS_CHIEF]) == (1, 5, False)
    assert drv._get_partition_list() == [1, 6, 11]
    drv.run_training(SchemaParams(uid_column_name="uid"), export_model=True)
    trained = [c for c in model.calls if c[0] == "train"]
    assert [c[1]["execution_context"][constants.PARTITION_INDEX] for c in trained] == [1, 6, 11]
    kw = trained[1][1]
    assert kw["training_data_dir"].endswith("train/active/partitionId=6")
    assert kw["validation_data_dir"].endswith("valid/partitionId=6")
    assert kw["checkpoint_path"].endswith("models/partitionId=6")
    ec = kw["execution_context"]
    assert ec[constants.ACTIVE_TRAINING_OUTPUT_FILE].endswith("s_train/partitionId=6/part-00001-active.avro")
    assert ec[constants.PASSIVE_TRAINING_OUTPUT_FILE].endswith("s_train/partitionId=6/part-00001-passive.avro")
    assert ec[constants.VALIDATION_OUTPUT_FILE].endswith("s_valid/partitionId=6/part-00001.avro")
    assert ec[constants.PASSIVE_TRAINING_DATA_DIR].endswith("train/passive/partitionId=6")
    assert not any(c[0] == "export" for c in model.calls)  # only the chief exports


def test_random_effect_driver_skips_empty_partitions_and_uses_rank_env(job, monkeypatch):
    monkeypatch.delenv(constants.TF_CONFIG, raising=False)
    monkeypatch.setenv("RANK", "0")
    monkeypatch.setenv("WORLD_SIZE", "5")
    model = RecordingModel(job)
    drv = RandomEffectDriver(_params(job, constants.RANDOM_EFFECT), model)
    assert drv._get_partition_list() == [0, 5, 10]
    drv.run_training(SchemaParams(uid_column_name="uid"))
    assert model.calls == []                               # partitions 0, 5, 10 hold no data


def test_random_effect_driver_requires_partition_list_and_job_name(job, monkeypatch):
    p = _params(job, constants.RANDOM_EFFECT)
    p.partition_list_file = None
    with pytest.raises(AssertionError):
        RandomEffectDriver(p, RecordingModel(job))
    monkeypatch.setenv(constants.TF_CONFIG, json.dumps({"cluster": {"worker": ["a:1"]}}))
    with pytest.raises(Exception):
        RandomEffectDriver(_params(job, constants.RANDOM_EFFECT), RecordingModel(job))


def test_fixed_effect_driver_context(job, monkeypatch):
    for k in ("RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):   # the driver derives these from TF_CONFIG:
        monkeypatch.setenv(k, "x"); monkeypatch.delenv(k)            # have monkeypatch restore them afterwards
    monkeypatch.setenv(constants.TF_CONFIG, _tf_config(3))
    drv = FixedEffectDriver(_params(job, constants.FIXED_EFFECT), RecordingModel(job))
    ctx = drv.execution_context
    assert (ctx[constants.TASK_INDEX], ctx[constants.NUM_WORKERS], ctx[constants.NUM_SHARDS],
            ctx[constants.SHARD_INDEX], ctx[constants.IS_CHIEF]) == (3, 5, 5, 3, False)
    assert drv._get_partition_list() == [3]
    assert drv._anchor_directory("/x/y", 3) == "/x/y"
    monkeypatch.delenv(constants.TF_CONFIG)
    assert os.environ["RANK"] == "3" and os.environ["WORLD_SIZE"] == "5" and os.environ["MASTER_ADDR"] == "node0"
    monkeypatch.delenv("RANK", raising=False)
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    local = FixedEffectDriver(_params(job, constants.FIXED_EFFECT), RecordingModel(job)).execution_context
    assert (local[constants.TASK_INDEX], local[constants.NUM_WORKERS], local[constants.IS_CHIEF]) == (0, 1, True)


def test_inference_runs_only_on_workers(job, monkeypatch):
    monkeypatch.setenv(constants.TF_CONFIG, _tf_config(0, task_type="evaluator"))
    model = RecordingModel(job)
    RandomEffectDriver(_params(job, constants.RANDOM_EFFECT, "inference"), model).run_inference(
        SchemaParams(uid_column_name="uid"))
    assert model.calls == []
    monkeypatch.setenv(constants.TF_CONFIG, _tf_config(1))
    RandomEffectDriver(_params(job, constants.RANDOM_EFFECT, "inference"), model).run_inference(
        SchemaParams(uid_column_name="uid"))
    preds = [c[1] for c in model.calls if c[0] == "predict"]
    assert len(preds) == 6 and preds[0]["output_dir"].endswith("s_train/partitionId=1")


def test_factory_rejects_what_is_out_of_scope(job):
    with pytest.raises(Exception):
        DriverFactory.get_driver(Params(uid_column_name="uid", label_column_name="y", stage=constants.RANDOM_EFFECT,
                                        model_type=constants.LINEAR_REGRESSION, partition_list_file="p"), [])
    with pytest.raises(Exception):
        DriverFactory.get_driver(Params(uid_column_name="uid", label_column_name="y", model_type=constants.DETEXT), [])


def test_torch_env_is_derived_from_tf_config():
    """A fixed-effect job launched the reference's way (TF_CONFIG only) gets the rendezvous variables
    torch.distributed needs; variables torch.distributed.run already set are left alone."""
    from gdmix_b200.drivers import torch_env_from_tf_config
    env = {constants.TF_CONFIG: _tf_config(3)}
    got = torch_env_from_tf_config(env)
    assert got["RANK"] == "3" and got["WORLD_SIZE"] == "5" and env["MASTER_ADDR"] and int(env["MASTER_PORT"]) > 0
    env2 = {constants.TF_CONFIG: _tf_config(2), "RANK": "7", "WORLD_SIZE": "9", "MASTER_ADDR": "127.0.0.1",
            "MASTER_PORT": "1234"}
    assert torch_env_from_tf_config(env2) == {} and env2["RANK"] == "7"
    assert torch_env_from_tf_config({}) == {}


def test_select_device_without_a_gpu_is_a_no_op(monkeypatch):
    import torch
    from gdmix_b200.drivers import select_device
    if not torch.cuda.is_available():
        assert select_device(3) is None
    else:
        monkeypatch.setenv("LOCAL_RANK", "0")
        assert select_device(3) == 0
