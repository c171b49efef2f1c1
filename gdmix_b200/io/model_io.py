"""Photon-ML Avro model files and score files -- the outputs of the path.

Mirrors gdmix-trainer/src/gdmix/util/io_utils.py of the reference:
  gen_one_avro_model              :102-160   (intercept always first, coefficients with abs(w) <= 1e-4 dropped)
  export_linear_model_to_avro     :163-212
  load_linear_models_from_avro    :45-83     (fixed effect: intercept moved to the END)
  read_feature_list / get_feature_map :215-239
  get_inference_output_avro_schema    :367-375
  batched_write_avro              :299-334
and the schema at models/schemas.py:3-51.
"""
import csv
import json
import os

import numpy as np

from ..constants import INTERCEPT
from . import avro

BAYESIAN_LINEAR_MODEL_SCHEMA = {
    "type": "record",
    "name": "BayesianLinearModelAvro",
    "namespace": "com.linkedin.photon.avro.generated",
    "doc": "a generic schema to describe a Bayesian linear model with means and variances",
    "fields": [
        {"name": "modelId", "type": "string"},
        {"name": "modelClass", "type": ["null", "string"], "default": None,
         "doc": "The fully-qualified class name of enclosing GLM model class. E.g.: "
                "com.linkedin.photon.ml.supervised.classification.LogisticRegressionModel"},
        {"name": "means", "type": {"type": "array", "items": {
            "type": "record", "name": "NameTermValueAvro",
            "doc": "A tuple of name, term and value. Used as feature or model coefficient",
            "fields": [{"name": "name", "type": "string"}, {"name": "term", "type": "string"},
                       {"name": "value", "type": "double"}]}}},
        {"name": "variances", "type": ["null", {"type": "array", "items": "NameTermValueAvro"}], "default": None},
        {"name": "lossFunction", "type": ["null", "string"], "default": None,
         "doc": "The loss function used for training as the class name. E.g.: "
                "com.linkedin.photon.ml.function.LogisticLossFunction"},
    ],
}
LOGISTIC_MODEL_CLASS = "com.linkedin.photon.ml.supervised.classification.LogisticRegressionModel"


def _field_shapes(schema):
    """(name, type) of every top-level field with named types reduced to their names and docs / defaults dropped."""
    def shape(t):
        if isinstance(t, list):
            return [shape(x) for x in t]
        if isinstance(t, dict):
            if t.get("type") == "array":
                return {"array": shape(t["items"])}
            if t.get("type") == "record":
                return {"record": t.get("name"), "fields": [(f["name"], shape(f["type"])) for f in t["fields"]]}
            return shape(t.get("type"))
        return str(t).rsplit(".", 1)[-1]
    return [(f["name"], shape(f["type"])) for f in schema.get("fields", [])]


def check_model_schema(writer_schema, where=""):
    """The library's block decoder (gdmix_avro_model_decode) is laid out for BAYESIAN_LINEAR_MODEL_SCHEMA: a file
    whose writer schema orders or types its fields differently would decode as garbage, so it is refused.
    (fastavro in the reference follows the writer schema; files the reference or this package wrote match.)"""
    if isinstance(writer_schema, (bytes, str)):
        writer_schema = json.loads(writer_schema)
    if _field_shapes(writer_schema) != _field_shapes(BAYESIAN_LINEAR_MODEL_SCHEMA):
        raise ValueError(f"{where}: writer schema is not BayesianLinearModelAvro as Photon-ML / GDMix write it "
                         f"(fields {[f.get('name') for f in writer_schema.get('fields', [])]})")


_FEATURE_LIST_CACHE = {}


def read_feature_list(feature_file):
    """CSV rows ``name,term``; the row number is the global feature index (intercept not included).  The parsed list is
    kept per (path, mtime, size): a job reads the same million-row file for every partition's model and score pass."""
    st = os.stat(feature_file)
    key = (os.path.abspath(feature_file), st.st_mtime_ns, st.st_size)
    hit = _FEATURE_LIST_CACHE.get(key)
    if hit is not None:
        return hit
    result = _read_feature_list(feature_file)
    _FEATURE_LIST_CACHE.clear()
    _FEATURE_LIST_CACHE[key] = result
    return result


def _read_feature_list(feature_file):
    result = []
    with open(feature_file, newline="") as f:
        for row in csv.reader(f):
            assert len(row) == 2, f"Each feature name should have exactly name and term only, but I got {row}."
            result.append(tuple(row))
    return result


def get_feature_map(feature_file):
    return {feature: index for index, feature in enumerate(read_feature_list(feature_file))}


def gen_one_avro_model(model_id, model_class, weight_indices, weight_values, bias, feature_list, sparsity_threshold):
    """One model record.  `weight_values` / `bias` are plain values, or (mean, variance) tuples."""
    has_bias = bias is not None
    if isinstance(bias, tuple) and len(bias) == 2 and bias[1] is not None:
        has_variance = True
    elif weight_values is not None and isinstance(weight_values, tuple) and len(weight_values) == 2 \
            and weight_values[1] is not None:
        has_variance = True
    else:
        has_variance = False
    record = {"modelId": model_id, "modelClass": model_class, "means": [], "lossFunction": ""}
    if has_bias:
        record["means"].append({"name": INTERCEPT, "term": "", "value": float(bias[0] if has_variance else bias)})
    if has_variance:
        record["variances"] = []
        if has_bias:
            record["variances"].append({"name": INTERCEPT, "term": "", "value": float(bias[1])})
    if weight_indices is not None and weight_values is not None:
        if has_variance:
            mean, variance = weight_values
            variance = np.asarray(variance).flatten()
        else:
            mean, variance = weight_values, None
        for i, (w_i, w_v) in enumerate(zip(np.asarray(weight_indices).flatten(), np.asarray(mean).flatten())):
            if abs(w_v) > sparsity_threshold:
                name, term = feature_list[int(w_i)][0], feature_list[int(w_i)][1]
                record["means"].append({"name": name, "term": term, "value": float(w_v)})
                if has_variance:
                    record["variances"].append({"name": name, "term": term, "value": float(variance[i])})
    return record


def export_linear_model_to_avro(model_ids, list_of_weight_indices, list_of_weight_values, biases, feature_file,
                                output_file, model_log_interval=1000, model_class=LOGISTIC_MODEL_CLASS,
                                sparsity_threshold=1.0e-4):
    feature_list = read_feature_list(feature_file) if feature_file else None
    num_models = len(list_of_weight_indices) if biases is None else len(biases)

    def gen_records():
        no_weights = list_of_weight_indices is None or list_of_weight_values is None or feature_list is None
        for i in range(num_models):
            current_bias = None if biases is None else biases[i]
            if no_weights:
                yield gen_one_avro_model(str(model_ids[i]), model_class, None, None, current_bias, feature_list,
                                         sparsity_threshold)
            else:
                yield gen_one_avro_model(str(model_ids[i]), model_class, list_of_weight_indices[i],
                                         list_of_weight_values[i], current_bias, feature_list, sparsity_threshold)

    avro.write_records(output_file, BAYESIAN_LINEAR_MODEL_SCHEMA, gen_records())


_FEATURE_COLUMNS = {}


def _feature_columns(feature_file):
    """(names, terms) of the feature file as two lists, the same list objects for the same parsed file (so that the
    binding's string tables are built once per job, not once per partition)."""
    if not feature_file:
        return [], []
    fl = read_feature_list(feature_file)
    hit = _FEATURE_COLUMNS.get(id(fl))
    if hit is not None and hit[0] is fl:
        return hit[1], hit[2]
    names, terms = [f[0] for f in fl], [f[1] for f in fl]
    _FEATURE_COLUMNS.clear()
    _FEATURE_COLUMNS[id(fl)] = (fl, names, terms)
    return names, terms


def export_random_effect_models(model_ids, coef, var, coef_ptr, feat_idx, has_intercept, feature_file, output_file,
                                model_class=LOGISTIC_MODEL_CLASS, sparsity_threshold=1.0e-4, sync=None, id_table=None):
    """The same file export_linear_model_to_avro writes for per-entity models, from flat arrays: model m owns
    coef[coef_ptr[m]:coef_ptr[m+1]] (intercept first when has_intercept; var aligned or None) and feat_idx lists the
    global feature ids of its other coefficients.  Records are encoded by the library (gdmix_avro_model_blocks)."""
    from .. import _capi as capi
    names, terms = _feature_columns(feature_file)
    with avro.Writer(output_file, BAYESIAN_LINEAR_MODEL_SCHEMA, "null", sync=sync) as w:
        body = capi.avro_model_blocks(model_ids, coef, var, coef_ptr, feat_idx, has_intercept, sparsity_threshold,
                                      names, terms, model_class, INTERCEPT, w.sync, id_table=id_table)
        w.f.write(body)
        w.count += len(model_ids)
        return w.count


def load_linear_models_from_avro(model_file, feature_file):
    """Fixed-effect loader: dense coefficient arrays with the intercept moved to the end."""
    feature_map = None if feature_file is None else get_feature_map(feature_file)

    def one(model_record):
        num_features = 0 if feature_map is None else len(feature_map)
        coef = np.zeros(num_features + 1, dtype=np.float64)
        has_bias = 0
        for ntv in model_record["means"]:
            name, term, value = ntv["name"], ntv["term"], np.float64(ntv["value"])
            if name == INTERCEPT and term == "":
                coef[num_features] = value
                has_bias = 1
            elif feature_map is not None:
                idx = feature_map.get((name, term))
                if idx is not None:
                    coef[idx] = value
        return coef[:num_features + has_bias]

    return tuple(one(r) for r in avro.read_records(model_file))


def add_dummy_weight(models):
    """Intercept-only models get a zero first weight (io_utils.py:86-99)."""
    out = []
    for m in models:
        c = np.zeros(2, dtype=np.float64)
        c[1] = m[0]
        out.append(c)
    return tuple(out)


def get_inference_output_avro_schema(metadata, has_logits_per_coordinate, schema_params, has_weight=False):
    fields = [{"name": schema_params.uid_column_name, "type": "long"},
              {"name": schema_params.prediction_score_column_name, "type": "float"},
              {"name": schema_params.label_column_name, "type": ["null", "float"], "default": None}]
    if has_weight or metadata.get(schema_params.weight_column_name) is not None:
        fields.append({"name": schema_params.weight_column_name, "type": "float"})
    if has_logits_per_coordinate:
        fields.append({"name": schema_params.prediction_score_per_coordinate_column_name, "type": "float"})
    return {"name": "validation_result", "type": "record", "fields": fields}


def batched_write_avro(records, output_file, schema, write_frequency=1000, batch_size=1024):
    return avro.write_records(output_file, schema, records, batch_size=batch_size)


def write_scores(output_file, schema, schema_params, uid, score, per_coordinate, label=None, weight=None, sync=None):
    """The score file of a train / predict pass (records of get_inference_output_avro_schema in blocks of 1024, as
    batched_write_avro writes them) from whole columns: the records are encoded by the library
    (gdmix_avro_score_blocks) instead of one Python dict at a time -- the same bytes, ~100x faster.
    label None -> the union's null branch; weight is written only when the schema has the field."""
    from .. import _capi as capi
    names = [f["name"] for f in schema["fields"]]
    sp = schema_params
    expect = [sp.uid_column_name, sp.prediction_score_column_name, sp.label_column_name]
    has_w = sp.weight_column_name in names
    if has_w:
        expect.append(sp.weight_column_name)
    has_pc = sp.prediction_score_per_coordinate_column_name in names
    if has_pc:
        expect.append(sp.prediction_score_per_coordinate_column_name)
    if names != expect:
        raise ValueError(f"unexpected score schema field order {names}")
    if has_w and weight is None:
        raise ValueError("the score schema has a weight field but no weights were given")
    with avro.Writer(output_file, schema, "null", sync=sync) as w:
        body = capi.avro_score_blocks(uid, score, label, weight if has_w else None, per_coordinate if has_pc else None,
                                      w.sync)
        w.f.write(body)
        w.count += int(np.asarray(uid).shape[0])
        return w.count


def dumps_schema(schema):
    return json.dumps(schema)
