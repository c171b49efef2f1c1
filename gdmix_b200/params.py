"""CLI parameter groups, same field names / defaults / checks as the reference's smart_arg dataclasses:

  GDMixParams / SchemaParams / Params   gdmix-trainer/src/gdmix/params.py:12-54
  LRParams                              models/custom/base_lr_params.py:5-42
  REParams                              models/custom/random_effect_lr_lbfgs_model.py:34-53
  FixedLRParams                         models/custom/fixed_effect_lr_lbfgs_model.py:55-71

smart_arg is not available here; ``from_argv`` implements the part of its behaviour the DAG relies on:
``--name=value`` or ``--name value`` pairs, every group picks its own fields out of one shared argv and ignores the
rest (``error_on_unknown=False``), booleans arrive as the strings ``True`` / ``False``
(test_random_effect_lr_lbfgs_model.py:54-56), ``Optional`` fields accept ``None``.
"""
import dataclasses
import typing
from dataclasses import dataclass
from typing import Optional

from . import constants

_ACTIONS = (constants.ACTION_INFERENCE, constants.ACTION_TRAIN)
_STAGES = (constants.FIXED_EFFECT, constants.RANDOM_EFFECT)
_MODEL_TYPES = (constants.LOGISTIC_REGRESSION, constants.LINEAR_REGRESSION, constants.DETEXT)
_VARIANCE_MODE = (constants.FULL, constants.SIMPLE)
_MISSING = object()


def _convert(tp, raw):
    origin = typing.get_origin(tp)
    if origin is typing.Union:  # Optional[X]
        inner = [a for a in typing.get_args(tp) if a is not type(None)]
        if raw is None or raw == "None":
            return None
        return _convert(inner[0], raw)
    if tp is bool:
        if isinstance(raw, bool):
            return raw
        if raw in ("True", "true", "1"):
            return True
        if raw in ("False", "false", "0"):
            return False
        raise ValueError(f"cannot parse {raw!r} as bool")
    if tp is int:
        return int(raw)
    if tp is float:
        return float(raw)
    return raw


def _argv_to_dict(argv):
    out = {}
    argv = list(argv)
    i = 0
    while i < len(argv):
        a = argv[i]
        if isinstance(a, str) and a.startswith("--"):
            if "=" in a:
                k, v = a[2:].split("=", 1)
                out[k] = v
            elif i + 1 < len(argv) and not str(argv[i + 1]).startswith("--"):
                out[a[2:]] = argv[i + 1]
                i += 1
            else:
                out[a[2:]] = "True"
        i += 1
    return out


class _ArgSuite:
    @classmethod
    def from_argv(cls, argv, error_on_unknown=False):
        given = _argv_to_dict(argv)
        hints = typing.get_type_hints(cls)
        kwargs = {}
        names = set()
        for f in dataclasses.fields(cls):
            names.add(f.name)
            if f.name in given:
                kwargs[f.name] = _convert(hints[f.name], given[f.name])
            elif f.default is dataclasses.MISSING and f.default_factory is dataclasses.MISSING:
                raise ValueError(f"{cls.__name__}: required argument --{f.name} is missing")
        if error_on_unknown:
            unknown = set(given) - names
            if unknown:
                raise ValueError(f"{cls.__name__}: unknown arguments {sorted(unknown)}")
        return cls(**kwargs)

    # the reference's spelling (smart_arg)
    __from_argv__ = from_argv

    def to_argv(self):
        out = []
        for f in dataclasses.fields(self):
            v = getattr(self, f.name)
            if v is not None:
                out.extend([f"--{f.name}", str(v)])
        return out

    __to_argv__ = to_argv


@dataclass
class SchemaParams(_ArgSuite):
    uid_column_name: str = dataclasses.field(default=_MISSING)
    weight_column_name: Optional[str] = None
    label_column_name: Optional[str] = None
    prediction_score_column_name: Optional[str] = None
    prediction_score_per_coordinate_column_name: str = "predictionScorePerCoordinate"

    def __post_init__(self):
        if self.uid_column_name is _MISSING:
            raise ValueError("SchemaParams: required argument --uid_column_name is missing")


@dataclass
class Params(SchemaParams):
    """GDMix driver parameters (GDMixParams + SchemaParams)."""
    action: str = constants.ACTION_TRAIN
    stage: str = constants.FIXED_EFFECT
    model_type: str = constants.LOGISTIC_REGRESSION
    training_score_dir: Optional[str] = None
    validation_score_dir: Optional[str] = None
    partition_list_file: Optional[str] = None

    def __post_init__(self):
        super().__post_init__()
        assert self.action in _ACTIONS, f"Action: {self.action} must be in {_ACTIONS}"
        assert self.stage in _STAGES, f"Stage: {self.stage} must be in {_STAGES}"
        assert self.model_type in _MODEL_TYPES, f"Model type: {self.model_type} must be in {_MODEL_TYPES}"
        assert (self.action == constants.ACTION_TRAIN and self.label_column_name) or \
               (self.action == constants.ACTION_INFERENCE and self.prediction_score_column_name)


@dataclass
class LRParams(_ArgSuite):
    """Base linear model parameters."""
    metadata_file: str = dataclasses.field(default=_MISSING)
    output_model_dir: str = dataclasses.field(default=_MISSING)
    training_data_dir: Optional[str] = None
    validation_data_dir: Optional[str] = None
    feature_bag: Optional[str] = None
    feature_file: Optional[str] = None
    regularize_bias: bool = True
    l2_reg_weight: float = 1.0
    lbfgs_tolerance: float = 1e-12
    num_of_lbfgs_curvature_pairs: int = 10
    num_of_lbfgs_iterations: int = 100
    has_intercept: bool = True
    offset_column_name: str = "offset"
    batch_size: int = 16
    data_format: str = "tfrecord"
    # not a CLI field in the reference either (no annotation there): always 1e-4
    sparsity_threshold = 1.0e-4

    def __post_init__(self):
        for name in ("metadata_file", "output_model_dir"):
            if getattr(self, name) is _MISSING:
                raise ValueError(f"{type(self).__name__}: required argument --{name} is missing")
        assert self.batch_size > 0, "Batch size must be positive number"
        if self.regularize_bias:
            assert self.has_intercept, "Intercept must be used when it is regularized"
        assert self.feature_bag or self.has_intercept, "Either intercept or feature bag much be used"


@dataclass
class REParams(LRParams):
    """Random-effect LR parameters."""
    partition_entity: Optional[str] = None
    enable_local_indexing: bool = False
    max_training_queue_size: int = 10
    training_queue_timeout_in_seconds: int = 300
    num_of_consumers: int = 2
    random_effect_variance_mode: Optional[str] = None
    disable_random_effect_scoring_after_training: bool = False

    def __post_init__(self):
        super().__post_init__()
        assert self.max_training_queue_size > self.num_of_consumers, \
            "queue size limit must be larger than the number of consumers"
        assert self.random_effect_variance_mode is None or self.random_effect_variance_mode in _VARIANCE_MODE, \
            f"Action: {self.random_effect_variance_mode} must be in {_VARIANCE_MODE}"


@dataclass
class FixedLRParams(LRParams):
    """Fixed-effect LR / linear-regression parameters."""
    copy_to_local: bool = True
    num_server_creation_retries: int = 50
    retry_interval: int = 2
    delayed_exit_in_seconds: int = 60
    disable_fixed_effect_scoring_after_training: bool = False
    fixed_effect_variance_mode: Optional[str] = None

    def __post_init__(self):
        super().__post_init__()
        assert self.fixed_effect_variance_mode is None or self.fixed_effect_variance_mode in _VARIANCE_MODE, \
            f"Action: {self.fixed_effect_variance_mode} must be in {_VARIANCE_MODE}"
