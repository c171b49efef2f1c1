"""FixedEffectLRModelLBFGS -- the reference's fixed-effect LR / linear-regression plugin on the GPU.

Same constructor ``(raw_model_params, base_training_params)``, attributes, ``train`` / ``predict`` / ``export``
and files as gdmix-trainer/src/gdmix/models/custom/fixed_effect_lr_lbfgs_model.py:74-812; underneath:

  reference                                                     here
  ------------------------------------------------------------  ----------------------------------------------
  per_record_input_fn over this worker's files (TF)             ingest.read_per_record over ``files[rank::world]``
  _train_model_fn: TF while_loop, sum of loss / gradient        gdmix_fe_loss_grad (one CUDA pass over the shard)
  two collective_ops.all_reduce per evaluation (:382-390)       one all_reduce of [value | gradient] (NCCL)
  scipy fmin_l_bfgs_b replicated on every worker (:635-643)     gdmix_lbfgs_* replicated on every rank
  _scoring_fn + Hessian accumulator H (:214-307)                gdmix_fe_score, gdmix_fe_hessian (+ all_reduce)
  chief writes the Photon-ML Avro model (:690-728)              same record, io.model_io

``FixedEffectLRLBFGSModel`` (the north-star's spelling) is an alias of ``FixedEffectLRModelLBFGS`` (the
reference's).  ``copy_to_local``, ``num_server_creation_retries``, ``retry_interval`` and
``delayed_exit_in_seconds`` are accepted for CLI compatibility and unused (no TF server to create or drain).
"""
import glob
import logging
import os

import numpy as np

from . import _capi as capi
from . import constants, ingest
from .api import Model
from .fe_solver import FixedEffectSolver
from .io import model_io
from .io.dataset_metadata import DatasetMetadata, read_json_file
from .params import FixedLRParams

logger = logging.getLogger(__name__)

LINEAR_MODEL_CLASS = "com.linkedin.photon.ml.supervised.regression.LinearRegressionModel"


class FixedEffectLRModelLBFGS(Model):
    """Global linear model: L-BFGS on the all-reduced objective of every worker's shard."""

    def __init__(self, raw_model_params, base_training_params):
        super().__init__(raw_model_params)
        self.model_params: FixedLRParams = self._parse_parameters(raw_model_params)
        mp = self.model_params
        self.training_output_dir = base_training_params.training_score_dir
        self.validation_output_dir = base_training_params.validation_score_dir
        self.model_type = base_training_params.model_type
        self.training_data_dir = mp.training_data_dir
        self.validation_data_dir = mp.validation_data_dir
        self.metadata_file = mp.metadata_file
        self.checkpoint_path = mp.output_model_dir
        self.data_format = mp.data_format
        self.offset_column_name = mp.offset_column_name
        self.feature_bag_name = mp.feature_bag
        self.feature_file = mp.feature_file if self.feature_bag_name else None
        self.num_correction_pairs = mp.num_of_lbfgs_curvature_pairs
        self.has_intercept = mp.has_intercept
        self.is_regularize_bias = mp.regularize_bias
        self.max_iteration = mp.num_of_lbfgs_iterations
        self.l2_reg_weight = mp.l2_reg_weight
        self.sparsity_threshold = mp.sparsity_threshold
        if self.model_type == constants.LOGISTIC_REGRESSION:
            self.disable_fixed_effect_scoring_after_training = mp.disable_fixed_effect_scoring_after_training
        else:
            self.disable_fixed_effect_scoring_after_training = True  # no scoring after plain linear regression
        assert os.path.exists(self.metadata_file), "metadata file %s does not exist" % self.metadata_file
        self.metadata = read_json_file(self.metadata_file)
        self.tensor_metadata = DatasetMetadata(self.metadata_file)
        self.num_features = self._get_num_features()
        self.model_coefficients = None
        self.variances = None
        self.fixed_effect_variance_mode = mp.fixed_effect_variance_mode
        self.epsilon = 1.0e-12
        self.fit_info = None
        assert self.feature_file is None or os.path.exists(self.feature_file), \
            f"feature file {self.feature_file} doesn't exist."
        if self.fixed_effect_variance_mode is not None:
            assert self.model_type == constants.LOGISTIC_REGRESSION, \
                f"doesn't support variance computation for model type {self.model_type}."

    # ---- helpers ------------------------------------------------------------------------------------------
    def _parse_parameters(self, raw_model_parameters):
        params = FixedLRParams.__from_argv__(raw_model_parameters, error_on_unknown=False)
        logger.info(params)
        return params

    def _get_num_features(self):
        if self.feature_bag_name is None:
            return 1  # intercept-only model: one dummy all-zero feature
        n = self.tensor_metadata.get_feature_shape(self.feature_bag_name)[0]
        assert n > 0, "number of features must > 0"
        return n

    def _has_feature(self, name):
        return name in self.tensor_metadata.get_feature_names()

    def _has_label(self, name):
        return name in self.tensor_metadata.get_label_names()

    @staticmethod
    def _get_assigned_files(input_data_path, num_shards, shard_index):
        """util/distribution_utils.py:11-47: file-level sharding only."""
        all_files = ingest.list_tfrecord_files(input_data_path)
        assert len(all_files) >= num_shards, \
            "Doesn't support sample level sharding,number of files must >= number of workers"
        return all_files[shard_index::num_shards]

    def _opts(self):
        mp = self.model_params
        return capi.make_opts(l2=mp.l2_reg_weight, regularize_bias=mp.regularize_bias,
                              has_intercept=self.has_intercept, m=mp.num_of_lbfgs_curvature_pairs,
                              max_iter=mp.num_of_lbfgs_iterations, tol=mp.lbfgs_tolerance)

    def _device_rows(self, files, schema_params, num_workers):
        import torch
        data = ingest.read_per_record(files, self.feature_bag_name, schema_params.label_column_name,
                                      self.offset_column_name, schema_params.weight_column_name,
                                      schema_params.uid_column_name)
        if data.col.size and (data.col.min() < 0 or data.col.max() >= self.num_features):
            raise ValueError(f"feature index outside [0, {self.num_features})")
        label = np.where(np.isnan(data.label), 0.0, data.label).astype(np.float32)
        rows = capi.DeviceFeRows(data.rowptr, data.col, data.val, label, data.weight, data.offset,
                                 self.num_features, linear_regression=self.model_type == constants.LINEAR_REGRESSION,
                                 num_workers=num_workers, device=torch.device("cuda", torch.cuda.current_device()))
        return data, rows

    @staticmethod
    def _group(num_workers):
        """The process group of the job when it runs on more than one worker (torch.distributed over NCCL)."""
        if num_workers <= 1:
            return None
        import torch.distributed as dist
        if not dist.is_initialized():
            import torch
            dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
        assert dist.get_world_size() == num_workers, "TF_CONFIG / WORLD_SIZE disagree about the number of workers"
        return dist.group.WORLD

    # ---- training (:509-688) ---------------------------------------------------------------------------------
    def train(self, training_data_dir, validation_data_dir, metadata_file, checkpoint_path, execution_context,
              schema_params):
        logger.info("Kicking off fixed effect LR LBFGS training")
        import torch
        task_index = execution_context[constants.TASK_INDEX]
        num_workers = execution_context[constants.NUM_WORKERS]
        is_chief = execution_context[constants.IS_CHIEF]
        group = self._group(num_workers)
        files = self._get_assigned_files(training_data_dir, num_workers, task_index)
        data, rows = self._device_rows(files, schema_params, num_workers)
        opts = self._opts()
        solver = FixedEffectSolver(rows, opts, self.num_features, group=group)

        prev_model = self._load_model(catch_exception=True)
        expected = self.num_features + 1 if self.has_intercept else self.num_features
        if prev_model is None or len(prev_model) != expected:
            logger.info("No usable initial model found, use all zeros instead.")
            x0 = np.zeros(expected)
        else:
            logger.info("Found a previous model,  loaded as the initial point for training")
            x0 = np.asarray(prev_model, dtype=np.float64)
        x, info = solver.fit(x0)
        self.fit_info = info
        logger.info(f"f_min: {info['f']}  num of funcalls: {info['nfev']}  iterations: {info['nit']}")
        # threshold_coefficients (util/model_utils.py:4-12)
        self.model_coefficients = np.where(np.abs(x) <= self.sparsity_threshold, 0.0, x)

        if self.fixed_effect_variance_mode is not None:
            mode = capi.VARIANCE_SIMPLE if self.fixed_effect_variance_mode == constants.SIMPLE else capi.VARIANCE_FULL
            xd = torch.from_numpy(self.model_coefficients).to(rows.val.device)
            H = capi.fe_hessian_device(rows, opts, xd, mode)
            if group is not None:
                torch.distributed.all_reduce(H, group=group)
            H = H.cpu().numpy()
            if mode == capi.VARIANCE_SIMPLE:
                H = H + self.l2_reg_weight
                if self.has_intercept and not self.is_regularize_bias:
                    H[-1] -= self.l2_reg_weight
                self.variances = 1.0 / (H + self.epsilon)
            else:
                H = H + np.diag([self.l2_reg_weight + self.epsilon] * H.shape[0])
                if self.has_intercept and not self.is_regularize_bias:
                    H[-1][-1] -= self.l2_reg_weight
                self.variances = np.diagonal(np.linalg.inv(H)).copy()
        if not self.disable_fixed_effect_scoring_after_training:
            self._score_and_write(solver, data, self.model_coefficients, task_index, schema_params,
                                  self.training_output_dir)
        if validation_data_dir:
            vfiles = self._get_assigned_files(validation_data_dir, num_workers, task_index)
            vdata, vrows = self._device_rows(vfiles, schema_params, num_workers)
            self._score_and_write(FixedEffectSolver(vrows, opts, self.num_features), vdata, self.model_coefficients,
                                  task_index, schema_params, self.validation_output_dir)
        if group is not None:
            torch.distributed.barrier(group=group)
        if is_chief:
            self._save_model()

    # ---- scoring (:406-473) ------------------------------------------------------------------------------------
    def _score_and_write(self, solver, data, x, task_index, schema_params, output_dir):
        logit, per_coordinate = solver.score(x)
        schema = model_io.get_inference_output_avro_schema(
            self.metadata, True, schema_params, has_weight=self._has_feature(schema_params.weight_column_name))
        sp = schema_params
        has_label = self._has_label(sp.label_column_name)
        has_weight = self._has_feature(sp.weight_column_name)

        os.makedirs(output_dir, exist_ok=True)
        # the reference writes int(weight) here (:426): the truncation is kept
        model_io.write_scores(os.path.join(output_dir, f"part-{task_index:05d}.avro"), schema, sp, data.uid, logit,
                              per_coordinate, label=data.label if has_label else None,
                              weight=np.trunc(data.weight))   # used only if the schema has the field

    # ---- model files (:690-750) -----------------------------------------------------------------------------------
    def _save_model(self):
        with_var = self.fixed_effect_variance_mode is not None
        if self.has_intercept:
            bias = (self.model_coefficients[-1], self.variances[-1]) if with_var else self.model_coefficients[-1]
        else:
            bias = None
        if self.feature_bag_name is None:
            indices = values = None
        else:
            weights = self.model_coefficients[:-1] if self.has_intercept else self.model_coefficients
            var = (self.variances[:-1] if self.has_intercept else self.variances) if with_var else None
            values = [weights] if var is None else [(weights, var)]
            indices = [np.arange(weights.shape[0])]
        model_class = model_io.LOGISTIC_MODEL_CLASS if self.model_type == constants.LOGISTIC_REGRESSION \
            else LINEAR_MODEL_CLASS
        model_io.export_linear_model_to_avro(model_ids=["global model"], list_of_weight_indices=indices,
                                             list_of_weight_values=values,
                                             biases=None if bias is None else [bias], feature_file=self.feature_file,
                                             output_file=os.path.join(self.checkpoint_path, "part-00000.avro"),
                                             model_class=model_class, sparsity_threshold=self.sparsity_threshold)

    def _load_model(self, catch_exception=False):
        model = None
        if self.checkpoint_path and os.path.exists(self.checkpoint_path):
            model_file = glob.glob(f"{self.checkpoint_path}/*.avro")
            if len(model_file) == 1:
                model = model_io.load_linear_models_from_avro(model_file[0], self.feature_file)[0]
            elif not catch_exception:
                raise ValueError("Load model failed, no model file or multiple model files found in the model "
                                 f"diretory {self.checkpoint_path}")
        elif not catch_exception:
            raise FileNotFoundError(f"checkpoint path {self.checkpoint_path} doesn't exist")
        if self.feature_bag_name is None and model is not None:
            model = model_io.add_dummy_weight([model])[0]  # intercept-only: a dummy zero weight in front
        return model

    def export(self, output_model_dir):
        logger.info("No need model export for LR model. ")

    # ---- inference (:752-807) ----------------------------------------------------------------------------------------
    def predict(self, output_dir, input_data_path, metadata_file, checkpoint_path, execution_context, schema_params):
        logger.info("Kicking off fixed effect LR predict")
        task_index = execution_context[constants.TASK_INDEX]
        num_workers = execution_context[constants.NUM_WORKERS]
        files = self._get_assigned_files(input_data_path, num_workers, task_index)
        data, rows = self._device_rows(files, schema_params, num_workers)
        x = np.asarray(self._load_model(), dtype=np.float64)
        self._score_and_write(FixedEffectSolver(rows, self._opts(), self.num_features), data, x, task_index,
                              schema_params, output_dir)


# the spelling BASELINE.json's north_star uses
FixedEffectLRLBFGSModel = FixedEffectLRModelLBFGS
