"""Drivers and factories: what ``python -m gdmix.gdmix`` instantiates in the reference, re-hosted without
TensorFlow (gdmix-trainer/src/gdmix/drivers/{driver,random_effect_driver,fixed_effect_driver}.py,
factory/{driver_factory,model_factory}.py).

Worker identity comes from ``TF_CONFIG`` exactly as in the reference (``task.index`` / number of ``worker``
entries, random_effect_driver.py:28-58); without it, from ``RANK`` / ``WORLD_SIZE`` (one process per GPU under
torch.distributed.run), else a single worker.  Random effect: worker r trains partitions
``partition_list[r::num_workers]`` (random_effect_driver.py:60-68) with no communication.  Fixed effect: rows
are sharded by file ``files[r::num_workers]`` and one all-reduce per objective evaluation joins the workers.
"""
import abc
import glob
import json
import logging
import os

from . import constants
from .ingest import is_empty_directory

logger = logging.getLogger(__name__)

TASK_TYPE_WORKER = constants.WORKER


def _cluster_from_env():
    """-> (task_type, task_index, num_workers) from TF_CONFIG, or from RANK/WORLD_SIZE, or (worker, 0, 1)."""
    tf_config = os.environ.get(constants.TF_CONFIG)
    if tf_config:
        cfg = json.loads(tf_config)
        task = cfg.get("task", {})
        workers = cfg.get("cluster", {}).get(constants.WORKER, [])
        return task.get("type"), task.get("index"), len(workers), True
    if "RANK" in os.environ and "WORLD_SIZE" in os.environ:
        return constants.WORKER, int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), False
    return constants.WORKER, 0, 1, False


def torch_env_from_tf_config(env=None):
    """TF_CONFIG -> the rendezvous variables torch.distributed reads, for jobs launched the reference's way (one
    process per worker with TF_CONFIG only, fixed_effect_driver.py:24-58): worker 0's "host:port" is the
    rendezvous address, task.index the rank, the number of workers the world size.  Variables that are already set
    (torch.distributed.run) win.  -> dict of the values derived (also written into `env`)."""
    env = os.environ if env is None else env
    tf_config = env.get(constants.TF_CONFIG)
    if not tf_config:
        return {}
    cfg = json.loads(tf_config)
    workers = cfg.get("cluster", {}).get(constants.WORKER, [])
    task = cfg.get("task", {})
    if not workers or task.get("index") is None:
        return {}
    host, _, port = str(workers[0]).rpartition(":")
    derived = {"MASTER_ADDR": host or str(workers[0]), "MASTER_PORT": port or "29500",
               "RANK": str(int(task["index"])), "WORLD_SIZE": str(len(workers))}
    out = {}
    for k, v in derived.items():
        if k not in env:
            env[k] = v
            out[k] = v
    return out


def select_device(task_index):
    """One process per GPU: LOCAL_RANK when torch.distributed.run set it, else task index modulo the visible GPUs.
    Without this every worker of a multi-worker job lands on cuda:0 (NCCL then refuses the duplicate GPU for the fixed
    effect and the random-effect workers serialise on one device)."""
    import torch
    if not torch.cuda.is_available():
        return None
    n = torch.cuda.device_count()
    local = os.environ.get("LOCAL_RANK")
    idx = int(local) if local is not None else int(task_index or 0) % max(n, 1)
    torch.cuda.set_device(idx)
    return idx


class Driver(abc.ABC):
    """driver.py:13-216."""

    def __init__(self, base_training_params, model, effect_name):
        self.base_training_params = base_training_params
        self.model = model
        self._validate_params()
        self.execution_context = self._setup_cluster()
        self.effect_name = effect_name

    @abc.abstractmethod
    def _validate_params(self):
        raise NotImplementedError

    @abc.abstractmethod
    def _setup_cluster(self):
        raise NotImplementedError

    @abc.abstractmethod
    def _get_partition_list(self):
        raise NotImplementedError

    @abc.abstractmethod
    def _anchor_directory(self, directory_path, partition_index):
        raise NotImplementedError

    def run_training(self, schema_params, export_model=False, output_model_dir=None):
        logger.info(f"Commencing {self.effect_name} training")
        logger.info(f"Execution context : {self.execution_context}")
        partition_index_list = self._get_partition_list()
        logger.info(f"This worker on work on the following list of partitions : {partition_index_list}")
        for partition_index in partition_index_list:
            checkpoint_path = self._anchor_directory(self.model.checkpoint_path, partition_index)
            training_data_dir = self._anchor_directory(self.model.training_data_dir, partition_index)
            validation_data_dir = self._anchor_directory(self.model.validation_data_dir, partition_index) \
                if self.model.validation_data_dir else None
            if is_empty_directory(training_data_dir):
                logger.info(f"{training_data_dir} is empty, no dataset to train on.")
                continue
            self.execution_context[constants.PARTITION_INDEX] = partition_index
            self.model.train(training_data_dir=training_data_dir, validation_data_dir=validation_data_dir,
                             metadata_file=self.model.metadata_file, checkpoint_path=checkpoint_path,
                             execution_context=self._prepare_training_context(partition_index),
                             schema_params=schema_params)
            if export_model and self.execution_context[constants.IS_CHIEF]:
                self.model.export(output_model_dir=output_model_dir)

    def run_inference(self, schema_params):
        logger.info(f"Commencing {self.effect_name} inference")
        if self.execution_context[constants.TASK_TYPE] != TASK_TYPE_WORKER:
            logger.info("Only workers should run inference. Exiting")
            return
        for partition_index in self._get_partition_list():
            self.execution_context[constants.PARTITION_INDEX] = partition_index
            for input_path, output_path in (
                    (self.model.training_data_dir, self.base_training_params.training_score_dir),
                    (self.model.validation_data_dir, self.base_training_params.validation_score_dir)):
                if input_path and output_path:
                    data_path = self._anchor_directory(input_path, partition_index)
                    output_dir = os.path.join(self._anchor_directory(output_path, partition_index))
                    if is_empty_directory(input_path):
                        logger.info(f"{input_path} is empty, no dataset to inference on.")
                        continue
                    self.model.predict(output_dir=output_dir, input_data_path=data_path,
                                       metadata_file=self.model.metadata_file,
                                       checkpoint_path=self.model.checkpoint_path,
                                       execution_context=self.execution_context, schema_params=schema_params)
        logger.info("Inference complete")

    def export_model(self, output_model_dir):
        self.model.export(output_model_dir=output_model_dir)

    def _prepare_training_context(self, partition_index):
        """Score-file names of the random-effect stage (driver.py:191-216)."""
        if self.base_training_params.stage != constants.RANDOM_EFFECT:
            return self.execution_context
        bp, task = self.base_training_params, self.execution_context[constants.TASK_INDEX]
        ctx = dict(self.execution_context)
        tdir = self._anchor_directory(bp.training_score_dir, partition_index)
        ctx[constants.ACTIVE_TRAINING_OUTPUT_FILE] = os.path.join(tdir, f"part-{task:05d}-active.avro")
        ctx[constants.PASSIVE_TRAINING_OUTPUT_FILE] = os.path.join(tdir, f"part-{task:05d}-passive.avro")
        ctx[constants.VALIDATION_OUTPUT_FILE] = os.path.join(
            self._anchor_directory(bp.validation_score_dir, partition_index),
            f"part-{task:05d}.avro") if bp.validation_score_dir else None
        passive = self._anchor_directory(self.model.passive_training_data_dir, partition_index)
        if os.path.exists(passive) and len(glob.glob(os.path.join(passive, "[!.]*"))) != 0:
            ctx[constants.PASSIVE_TRAINING_DATA_DIR] = passive
        return ctx


class RandomEffectDriver(Driver):
    """random_effect_driver.py:13-73."""
    _RANDOM_EFFECT_PARTITION_DIR_PREFIX = "partitionId="

    def __init__(self, base_training_params, model):
        super().__init__(base_training_params, model, constants.RANDOM_EFFECT)

    def _validate_params(self):
        assert self.base_training_params.model_type == constants.LOGISTIC_REGRESSION, \
            "Random effect supports logistic_regression only"
        assert self.base_training_params.partition_list_file is not None, "Random effect requires partition list file"

    def _setup_cluster(self):
        task_type, task_index, num_workers, from_tf_config = _cluster_from_env()
        if from_tf_config:
            if task_type is None or task_index is None:
                raise Exception("No job name found")
            if num_workers < 1:
                raise Exception("No worker found")
            os.environ.pop(constants.TF_CONFIG, None)  # random effect runs in local mode
        select_device(task_index)
        return {constants.TASK_TYPE: task_type, constants.TASK_INDEX: task_index, constants.CLUSTER_SPEC: None,
                constants.NUM_WORKERS: num_workers, constants.NUM_SHARDS: 1, constants.SHARD_INDEX: 0,
                constants.IS_CHIEF: task_index == 0}

    def _get_partition_list(self):
        with open(self.base_training_params.partition_list_file) as f:
            line = f.readline()
        all_partitions = [int(x) for x in line.split(",")]
        return all_partitions[self.execution_context[constants.TASK_INDEX]::
                              self.execution_context[constants.NUM_WORKERS]]

    def _anchor_directory(self, directory_path, partition_index):
        return os.path.join(directory_path, self._RANDOM_EFFECT_PARTITION_DIR_PREFIX + str(partition_index))


class FixedEffectDriver(Driver):
    """fixed_effect_driver.py:14-67: one "partition" (the whole dataset), every worker takes part."""

    def __init__(self, base_training_params, model):
        super().__init__(base_training_params, model, constants.FIXED_EFFECT)

    def _validate_params(self):
        pass

    def _setup_cluster(self):
        task_type, task_index, num_workers, from_tf_config = _cluster_from_env()
        if from_tf_config and (task_type is None or task_index is None):
            raise Exception("No job name found")
        if from_tf_config and num_workers > 1:
            torch_env_from_tf_config()     # RANK / WORLD_SIZE / MASTER_* for the NCCL rendezvous of the all-reduce
        select_device(task_index)
        return {constants.TASK_TYPE: task_type, constants.TASK_INDEX: task_index, constants.CLUSTER_SPEC: None,
                constants.NUM_WORKERS: num_workers, constants.NUM_SHARDS: num_workers,
                constants.SHARD_INDEX: task_index, constants.IS_CHIEF: task_index == 0}

    def _get_partition_list(self):
        return [self.execution_context[constants.TASK_INDEX]]  # partition index == task index (:60-62)

    def _anchor_directory(self, directory_path, partition_index):
        return directory_path


class ModelFactory:
    """model_factory.py:25-54.  DeText models are outside this package's scope."""

    @staticmethod
    def get_model(base_training_params, raw_model_params):
        from .fixed_effect import FixedEffectLRModelLBFGS
        from .random_effect import RandomEffectLRLBFGSModel
        model_type, stage = base_training_params.model_type, base_training_params.stage
        if model_type in (constants.LOGISTIC_REGRESSION, constants.LINEAR_REGRESSION):
            if stage == constants.FIXED_EFFECT:
                return FixedEffectLRModelLBFGS(raw_model_params=raw_model_params,
                                               base_training_params=base_training_params)
            if model_type == constants.LINEAR_REGRESSION:
                raise Exception("Does not support random effect model for plain linear regression")
            return RandomEffectLRLBFGSModel(raw_model_params=raw_model_params)
        if model_type == constants.DETEXT:
            raise Exception("DeText models are not part of gdmix_b200 (the LR hot path only)")
        raise Exception(f"Unknown training models {model_type}")


class DriverFactory:
    """driver_factory.py:14-37."""
    drivers = {constants.FIXED_EFFECT: FixedEffectDriver, constants.RANDOM_EFFECT: RandomEffectDriver}

    @staticmethod
    def get_driver(base_training_params, raw_model_params):
        driver = DriverFactory.drivers[base_training_params.stage]
        model = ModelFactory.get_model(base_training_params, raw_model_params)
        return driver(base_training_params=base_training_params, model=model)
