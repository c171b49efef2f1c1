"""RandomEffectLRLBFGSModel -- the reference's random-effect plugin with its per-entity scipy loop replaced by
one batched CUDA solve per partition.

Same constructor, attributes, ``train`` / ``predict`` / ``export`` signatures, files and error behaviour as
gdmix-trainer/src/gdmix/models/custom/random_effect_lr_lbfgs_model.py:56-317; what changed underneath:

  reference                                                      here
  -------------------------------------------------------------  ---------------------------------------------
  per_entity_grouped_input_fn (TF) + prepare_jobs: one Job per   ingest.read_entity_grouped + to_local_batch: the
  entity through a Manager().Queue (job_consumers.py:161-296)    whole partition as one flat entity-local CSR batch
  TrainingJobConsumer: BinaryLogisticRegressionTrainer.fit per   ONE gdmix_re_fit_host call (C ABI): every entity
  entity in a process pool (job_consumers.py:36-63)              solved on the GPU, thresholded, SIMPLE variance
  InferenceJobConsumer (job_consumers.py:138-152)                ONE gdmix_re_score_host call
  fastavro model / score writers                                 io.model_io / io.avro (same schemas, same records)

The reference always fits in the index space ``enable_local_indexing`` selects; both give the same
coefficients (features an entity never sees keep a zero coefficient: SURVEY.md section 0), so the device always
works entity-locally and ``enable_local_indexing`` is accepted and ignored.  ``num_of_consumers``,
``max_training_queue_size`` and ``training_queue_timeout_in_seconds`` are accepted for CLI compatibility; there
is no queue.  There is no CPU fallback: without the CUDA library the package does not import.
"""
import logging
import os
import time
from collections import namedtuple

import numpy as np

from . import _capi as capi
from . import constants, ingest
from .api import Model
from .io import model_io
from .io.dataset_metadata import DatasetMetadata, read_json_file
from .params import REParams

logger = logging.getLogger(__name__)

# job_consumers.py:19
TrainingResult = namedtuple("TrainingResult", ("theta", "variance", "unique_global_indices"))


class FlatModels:
    """The models of one partition as the solve returned them -- flat arrays -- behind the mapping interface the
    reference passes around (entity id -> TrainingResult, job_consumers.py:19).  A partition of a million entities
    then never becomes a million Python tuples: the model writer and the scoring pass take the arrays as they are, and
    a TrainingResult is only built for an entity somebody asks for."""

    def __init__(self, entity_ids, theta, variance, theta_ptr, uniq_ptr, uniq_global, id_table=None):
        self.entity_ids = list(entity_ids)
        self.source_ids = entity_ids      # the parsed partition's own list: identity tells "same partition" in O(1)
        self.id_table = id_table          # (utf-8 characters, offsets) of the ids as the reader left them, or None
        self.theta, self.variance = theta, variance
        self.theta_ptr, self.uniq_ptr, self.uniq_global = theta_ptr, uniq_ptr, uniq_global
        self._index = None

    @property
    def index(self):
        if self._index is None:
            self._index = {eid: e for e, eid in enumerate(self.entity_ids)}
        return self._index

    def _result(self, e):
        a, b = int(self.theta_ptr[e]), int(self.theta_ptr[e + 1])
        return TrainingResult(theta=self.theta[a:b].copy(), variance=None if self.variance is None else self.variance[a:b].copy(),
                              unique_global_indices=self.uniq_global[int(self.uniq_ptr[e]):int(self.uniq_ptr[e + 1])].copy())

    def __len__(self):
        return len(self.entity_ids)

    def __bool__(self):
        return bool(self.entity_ids)

    def __contains__(self, eid):
        return eid in self.index

    def __getitem__(self, eid):
        return self._result(self.index[eid])

    def get(self, eid, default=None):
        e = self.index.get(eid)
        return default if e is None else self._result(e)

    def keys(self):
        return list(self.entity_ids)

    def __iter__(self):
        return iter(self.entity_ids)

    def items(self):
        return ((eid, self._result(e)) for e, eid in enumerate(self.entity_ids))


class RandomEffectLRLBFGSModel(Model):
    """Entity-based logistic regression: all entities of a partition are trained in one GPU batch."""

    def __init__(self, raw_model_params):
        super().__init__(raw_model_params)
        self.model_params: REParams = self._parse_parameters(raw_model_params)
        self.checkpoint_path = os.path.join(self.model_params.output_model_dir)
        self.metadata_file = self.model_params.metadata_file
        self.feature_bag_name = self.model_params.feature_bag
        self.has_intercept = self.model_params.has_intercept
        # an intercept-only model has no feature file
        self.feature_file = None if self.feature_bag_name is None else self.model_params.feature_file
        if self.model_params.training_data_dir is not None:
            self.training_data_dir = os.path.join(self.model_params.training_data_dir, constants.ACTIVE)
            self.passive_training_data_dir = os.path.join(self.model_params.training_data_dir, constants.PASSIVE)
        else:
            self.training_data_dir = None
            self.passive_training_data_dir = None
        self.validation_data_dir = self.model_params.validation_data_dir
        self.disable_random_effect_scoring_after_training = \
            self.model_params.disable_random_effect_scoring_after_training
        self.last_fit_info = None  # nit / nfev / status arrays of the latest _train (diagnostics, tests)
        self.last_timing = {}      # seconds per phase of the latest train() (diagnostics, bench)
        self._parsed = None        # input path -> parsed partition, alive during one _action call

    # ---- Model API ------------------------------------------------------------------------------------
    def train(self, training_data_dir, validation_data_dir, metadata_file, checkpoint_path, execution_context,
              schema_params):
        logger.info("Kicking off random effect custom LR training")
        self._action(constants.ACTION_TRAIN, (training_data_dir, validation_data_dir), metadata_file,
                     checkpoint_path, execution_context, schema_params)

    def predict(self, output_dir, input_data_path, metadata_file, checkpoint_path, execution_context, schema_params):
        logger.info(f"Running inference on dataset : {input_data_path}, results to be written to path : {output_dir}")
        self._action(constants.ACTION_INFERENCE, (output_dir, input_data_path), metadata_file, checkpoint_path,
                     execution_context, schema_params)

    def export(self, output_model_dir):
        logger.info("Model export is done as part of the training() API for random effect LR LBFGS training. "
                    "Skipping.")

    def _parse_parameters(self, raw_model_parameters) -> REParams:
        params = REParams.__from_argv__(raw_model_parameters, error_on_unknown=False)
        logger.info(params)
        return params

    # ---- orchestration (random_effect_lr_lbfgs_model.py:92-138) ------------------------------------------
    def _action(self, action, action_context, metadata_file, checkpoint_path, execution_context, schema_params):
        partition_index = execution_context[constants.PARTITION_INDEX]
        self._parsed = {}
        try:
            return self._action_impl(action, action_context, metadata_file, checkpoint_path, execution_context,
                                     schema_params, partition_index)
        finally:
            self._parsed = None
            self._join_writer()

    def _action_impl(self, action, action_context, metadata_file, checkpoint_path, execution_context, schema_params,
                     partition_index):
        metadata = read_json_file(metadata_file)
        tensor_metadata = DatasetMetadata(metadata)
        # an intercept-only model is padded with one dummy (all-zero) feature
        num_features = 1 if self.feature_bag_name is None \
            else tensor_metadata.get_feature_shape(self.feature_bag_name)[0]
        logger.info(f"Found {num_features} features in feature bag {self.feature_bag_name}")
        assert num_features > 0, "number of features must > 0"
        avro_filename = f"part-{partition_index:05d}.avro"
        common = dict(metadata=metadata, tensor_metadata=tensor_metadata, schema_params=schema_params,
                      num_features=num_features)
        if action == constants.ACTION_INFERENCE:
            output_dir, input_data_path = action_context
            model_weights = self._load_weights(os.path.join(checkpoint_path, avro_filename))
            self._predict(input_path=input_data_path, output_file=os.path.join(output_dir, avro_filename),
                          model_weights=model_weights, **common)
        elif action == constants.ACTION_TRAIN:
            training_data_dir, validation_data_dir = action_context
            model_file = os.path.join(self.model_params.output_model_dir, avro_filename)
            model_weights = self._load_weights(model_file, True)  # warm start when a model is already there
            model_weights = self._train(training_data_dir, tensor_metadata, model_weights, num_features,
                                        schema_params, model_file)
            if validation_data_dir:
                o = execution_context.get(constants.VALIDATION_OUTPUT_FILE, None)
                o and self._predict(input_path=validation_data_dir, output_file=o, model_weights=model_weights,
                                    **common)
            if not self.disable_random_effect_scoring_after_training:
                o = execution_context.get(constants.ACTIVE_TRAINING_OUTPUT_FILE, None)
                o and self._predict(input_path=training_data_dir, output_file=o, model_weights=model_weights,
                                    **common)
                i = execution_context.get(constants.PASSIVE_TRAINING_DATA_DIR, None)
                o = execution_context.get(constants.PASSIVE_TRAINING_OUTPUT_FILE, None)
                i and o and self._predict(input_path=i, output_file=o, model_weights=model_weights, **common)
        else:
            raise ValueError(f"Invalid action {action!r}.")

    def _read(self, input_path, tensor_metadata, schema_params, num_features, need_label):
        assert self.model_params.data_format == constants.TFRECORD
        # scoring after training reads the partition the training just parsed: the parsed arrays (and their entity-local
        # form) are kept for the duration of one train() / predict() call and reused when the same path comes again
        cached = self._parsed.get(input_path) if self._parsed is not None else None
        if cached is not None:
            data = cached["data"]
            if need_label and data.label is None:
                raise ValueError(f"label column {schema_params.label_column_name!r} not found in {input_path}")
            return data
        data = self._read_uncached(input_path, tensor_metadata, schema_params, num_features, need_label)
        if self._parsed is not None:
            self._parsed[input_path] = {"data": data}
        return data

    def _local(self, input_path, data):
        """ingest.to_local_batch, once per parsed partition."""
        slot = self._parsed.get(input_path) if self._parsed is not None else None
        if slot is not None and "local" in slot and slot["data"] is data:
            return slot["local"]
        loc = ingest.to_local_batch(data, self.has_intercept)
        if slot is not None and slot["data"] is data:
            slot["local"] = loc
        return loc

    def _read_uncached(self, input_path, tensor_metadata, schema_params, num_features, need_label):
        data = ingest.read_entity_grouped(
            input_path, tensor_metadata, entity_name=self.model_params.partition_entity,
            feature_bag=self.feature_bag_name, label_column=schema_params.label_column_name,
            offset_column=self.model_params.offset_column_name, weight_column=schema_params.weight_column_name,
            uid_column=schema_params.uid_column_name, num_features=num_features)
        if need_label and data.label is None:
            raise ValueError(f"label column {schema_params.label_column_name!r} not found in {input_path}")
        return data

    def _opts(self, **kw):
        mp = self.model_params
        return capi.make_opts(l2=mp.l2_reg_weight, regularize_bias=mp.regularize_bias,
                              has_intercept=self.has_intercept, m=mp.num_of_lbfgs_curvature_pairs,
                              max_iter=mp.num_of_lbfgs_iterations, tol=mp.lbfgs_tolerance, **kw)

    # ---- training (random_effect_lr_lbfgs_model.py:140-167 + job_consumers.py:36-99) -----------------------
    def _train(self, input_path, tensor_metadata, model_weights: dict, num_features, schema_params,
               output_model_file):
        logger.info(f"Start training with "
                    f"{f'loaded {len(model_weights)} previous models' if model_weights else 'zeros'} "
                    f"as the model initial value.")
        mp = self.model_params
        tm = self.last_timing = {}
        t0 = time.perf_counter()
        data = self._read(input_path, tensor_metadata, schema_params, num_features, need_label=True)
        tm["read_s"] = time.perf_counter() - t0
        if data.n_entities:
            labels = data.label
            if not np.all((labels == 0) | (labels == 1)):
                raise ValueError("labels must be 0/1 for logistic regression")  # binary_logistic_regression.py:208
            t0 = time.perf_counter()
            hb, uniq_ptr, uniq_global = self._local(input_path, data)
            theta0, has_model = ingest.warm_start_theta(hb, uniq_ptr, uniq_global, data.entity_ids, model_weights,
                                                        self.has_intercept)
            tm["local_index_s"] = time.perf_counter() - t0
            mode = mp.random_effect_variance_mode
            if mode == constants.FULL:
                vmode = capi.VARIANCE_FULL
            elif mode == constants.SIMPLE:
                vmode = capi.VARIANCE_SIMPLE
            else:
                vmode = capi.VARIANCE_NONE
            opts = self._opts(sparsity_threshold=mp.sparsity_threshold, variance_mode=vmode)
            t0 = time.perf_counter()
            out = capi.re_fit_host(hb, opts, theta0=theta0 if has_model.any() else None,
                                   want_variance=vmode != capi.VARIANCE_NONE)
            tm["fit_s"] = time.perf_counter() - t0
            t0 = time.perf_counter()
            self.last_fit_info = {k: out[k] for k in ("nit", "nfev", "status", "f")}
            results = FlatModels(data.entity_ids, out["theta"], out["variance"] if vmode != capi.VARIANCE_NONE else None,
                                 hb.theta_ptr, uniq_ptr, uniq_global, id_table=getattr(data, "entity_id_table", None))
            if model_weights:
                # prior-only entities survive; prior-only features of a retrained entity do not (:161)
                model_weights = dict(model_weights.items())
                model_weights.update(dict(results.items()))
            else:
                model_weights = results
            tm["results_s"] = time.perf_counter() - t0
        logger.info(f"{len(model_weights)} models in total after training/refreshing.")
        # the model file is encoded and written (host threads of the library) while the scoring passes that follow use
        # the GPU; _action joins the writer before it returns, and its exception, if any, is raised there
        import threading

        def save():
            t0 = time.perf_counter()
            try:
                self._save_model(output_model_file, model_coefficients=model_weights, num_features=num_features,
                                 feature_file=self.feature_file)
            except BaseException as ex:      # noqa: BLE001  (re-raised by _join_writer)
                self._writer_error = ex
            tm["save_model_s"] = time.perf_counter() - t0

        self._writer_error = None
        self._writer = threading.Thread(target=save, name="gdmix-model-writer")
        self._writer.start()
        return model_weights

    def _join_writer(self):
        w, self._writer = getattr(self, "_writer", None), None
        if w is not None:
            w.join()
            err, self._writer_error = getattr(self, "_writer_error", None), None
            if err is not None:
                raise err

    # ---- inference (random_effect_lr_lbfgs_model.py:169-190 + job_consumers.py:102-152) --------------------
    def _predict(self, input_path, metadata, tensor_metadata, output_file, schema_params, num_features,
                 model_weights):
        logger.info(f"Start inference for {input_path}.")
        tm = self.last_timing
        t0 = time.perf_counter()
        data = self._read(input_path, tensor_metadata, schema_params, num_features, need_label=False)
        has_weight = any(schema_params.weight_column_name == f.name for f in tensor_metadata.get_features())
        schema = model_io.get_inference_output_avro_schema(metadata, True, schema_params, has_weight=has_weight)
        if data.n_entities:
            hb, uniq_ptr, uniq_global = self._local(input_path, data)
            theta, has_model = ingest.warm_start_theta(hb, uniq_ptr, uniq_global, data.entity_ids, model_weights,
                                                       self.has_intercept)
            tm["predict_prepare_s"] = tm.get("predict_prepare_s", 0.0) + time.perf_counter() - t0
            t0 = time.perf_counter()
            logit, per_coordinate = capi.re_score_host(hb, self._opts(), theta, has_model)
            tm["predict_score_s"] = tm.get("predict_score_s", 0.0) + time.perf_counter() - t0
        else:
            logit = per_coordinate = np.zeros(0, np.float32)
        sp = schema_params
        t0 = time.perf_counter()
        model_io.write_scores(output_file, schema, sp, data.uid, logit, per_coordinate, label=data.label,
                              weight=data.weight)
        tm["predict_write_s"] = tm.get("predict_write_s", 0.0) + time.perf_counter() - t0
        logger.info(f"Inference complete: {input_path}.")

    # ---- model files (random_effect_lr_lbfgs_model.py:219-309) ---------------------------------------------
    def _save_model(self, output_file, model_coefficients, num_features, feature_file):
        """One Photon-ML BayesianLinearModelAvro record per entity (random_effect_lr_lbfgs_model.py:219-260 ->
        export_linear_model_to_avro): intercept first, features above the sparsity threshold by name / term,
        variances aligned when a variance mode is set.  The records are encoded by the library from flat arrays."""
        with_variance = self.model_params.random_effect_variance_mode is not None
        hi = 1 if self.has_intercept else 0
        if feature_file is None:
            assert num_features == 1          # intercept-only model: nothing but the bias is written
        if isinstance(model_coefficients, FlatModels) and feature_file is not None and \
                (not with_variance or model_coefficients.variance is not None):
            fm = model_coefficients
            os.makedirs(os.path.dirname(output_file) or ".", exist_ok=True)
            model_io.export_random_effect_models(fm.entity_ids, fm.theta, fm.variance if with_variance else None,
                                                 fm.theta_ptr, fm.uniq_global, self.has_intercept, feature_file,
                                                 output_file, sparsity_threshold=self.model_params.sparsity_threshold,
                                                 id_table=getattr(fm, "id_table", None))
            return
        model_ids = list(model_coefficients.keys())
        means, variances, indices = [], [], []
        for entity_id, (mean, variance, unique_global_indices) in model_coefficients.items():
            mean = np.asarray(mean, dtype=np.float64).ravel()
            keep = hi if feature_file is None else mean.shape[0]
            means.append(mean[:keep])
            if with_variance:
                if variance is None:
                    # a prior-only entity loaded from a model file without variances: the reference fails here too
                    # (zip(means, None) in gen_one_avro_model, io_utils.py:102-160) -- never write misaligned arrays
                    raise TypeError(f"model {entity_id!r} has no variances but random_effect_variance_mode="
                                    f"{self.model_params.random_effect_variance_mode} asks for them")
                variance = np.asarray(variance, dtype=np.float64).ravel()
                if variance.shape[0] != mean.shape[0]:
                    raise ValueError(f"model {entity_id!r}: {variance.shape[0]} variances for {mean.shape[0]} means")
                variances.append(variance[:keep])
            if feature_file is not None:
                indices.append(np.asarray(unique_global_indices, dtype=np.int64).ravel())
        cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt)
        coef_ptr = np.concatenate([[0], np.cumsum([m.shape[0] for m in means])]).astype(np.int64)
        os.makedirs(os.path.dirname(output_file) or ".", exist_ok=True)
        model_io.export_random_effect_models(model_ids, cat(means, np.float64),
                                             cat(variances, np.float64) if with_variance else None, coef_ptr,
                                             cat(indices, np.int64), self.has_intercept, feature_file, output_file,
                                             sparsity_threshold=self.model_params.sparsity_threshold)

    def _load_weights(self, model_file, catch_exception=False):
        logger.info(f"Loading model from {model_file}")
        if not os.path.exists(model_file):
            if catch_exception:
                logger.info(f"No model found at {model_file}.")
                return {}
            raise FileNotFoundError(f"Model file {model_file} does not exist")
        return self._load_weights_native(model_file)

    def _load_weights_native(self, model_file):
        """entity id -> TrainingResult, decoded block by block by the library (gdmix_avro_model_decode); the same
        dict the reference builds one fastavro record at a time (random_effect_lr_lbfgs_model.py:262-309)."""
        from .io import avro
        feature_list = model_io.read_feature_list(self.feature_file) if self.feature_file else []
        fmap = capi.FeatureMap([f[0] for f in feature_list], [f[1] for f in feature_list], constants.INTERCEPT)
        hi = 1 if self.has_intercept else 0
        out = {}
        try:
            schema, blocks = avro.read_blocks(model_file)
            model_io.check_model_schema(schema, model_file)   # the block decoder is laid out for this schema only
            for n, data in blocks:
                try:
                    d = fmap.decode_models(data, n)
                except capi.GdmixError as ex:
                    raise KeyError(f"{model_file}: {ex}") from None
                ids = d["id_chars"].tobytes()
                ip, mp = d["id_ptr"], d["mean_ptr"]
                for m in range(n):
                    a, b = int(mp[m]), int(mp[m + 1])
                    feat = d["mean_feat"][a:b]
                    if hi and not (b > a and feat[0] == -1):
                        raise ValueError(f"{model_file}: the first mean of a model with intercept must be the intercept")
                    if (feat[hi:] < 0).any():
                        raise ValueError(f"{model_file}: an intercept among the feature coefficients")
                    theta, idx = d["mean_val"][a:b].copy(), feat[hi:].copy()
                    if self.feature_file is None:
                        # intercept-only model: one dummy feature
                        if idx.size:
                            raise ValueError(f"{model_file}: feature coefficients in an intercept-only model")
                        theta, idx = np.append(theta, 0.0), np.zeros(1, np.int64)
                    var = d["var_val"][a:b].copy() if d["has_var"][m] else None
                    out[ids[ip[m]:ip[m + 1]].decode("utf-8")] = TrainingResult(theta=theta, variance=var,
                                                                               unique_global_indices=idx)
        finally:
            fmap.close()
        return out
