"""``tensor_metadata.json`` reader: same accessors and validation as the reference's ``DatasetMetadata``
(gdmix-trainer/src/gdmix/io/dataset_metadata.py:5-130), with numpy dtypes in place of TF dtypes."""
import json
import os
from collections import namedtuple

import numpy as np

MetadataInfo = namedtuple("MetadataInfo", ["name", "dtype", "shape", "isSparse"])


def read_json_file(file_path):
    if not os.path.exists(file_path):
        raise IOError(f"Path {file_path!r} does not exist.")
    try:
        with open(file_path) as f:
            return json.load(f)
    except Exception as e:
        raise ValueError(f"Failed loading file {file_path!r}.") from e


class DatasetMetadata:
    TO_NP_DTYPE = {"int": np.int32, "long": np.int64, "float": np.float32, "double": np.float64,
                   "bytes": bytes, "string": bytes}
    FEATURES, LABELS, INDICES, VALUES = "features", "labels", "indices", "values"
    SUPPORTED_TYPES = frozenset(["int", "long", "float", "double", "bytes", "string"])
    METADATA_FIELDS = frozenset(["name", "dtype", "shape", "isSparse"])

    def __init__(self, path_or_metadata):
        if isinstance(path_or_metadata, str):
            path_or_metadata = read_json_file(path_or_metadata)
        md = path_or_metadata
        if not isinstance(md.get(self.FEATURES, []), list):
            raise TypeError(f"Features must be a list. Type {type(md[self.FEATURES])} detected.")
        if not isinstance(md.get(self.LABELS, []), list):
            raise TypeError(f"Labels must be a list. Type {type(md[self.LABELS])} detected.")

        def parse(key):
            tensors = {}
            for entity in md.get(key, []):
                name = entity.get("name")
                if name in tensors:
                    raise ValueError(f"Invalid field: Tensor name in your metadata appears more than once:{name}")
                tensors[name] = self._build(dict(entity))
            return tensors

        feats, labels = parse(self.FEATURES), parse(self.LABELS)
        self._tensors = {**feats, **labels}
        self._features, self._labels = list(feats.values()), list(labels.values())
        self._feature_names, self._label_names = list(feats.keys()), list(labels.keys())

    @classmethod
    def _build(cls, d):
        d.setdefault("isSparse", False)
        if not cls.METADATA_FIELDS.issubset(d.keys()):
            raise ValueError(f"Invalid field: Required metadata fields are {','.join(sorted(cls.METADATA_FIELDS))}. "
                             f"Provided fields are {','.join(d.keys())}")
        if d["name"] is None or not isinstance(d["name"], str):
            raise ValueError("Invalid field: Feature name can not be None and must be str")
        if d["dtype"] not in cls.SUPPORTED_TYPES:
            raise ValueError(f"Invalid field: User provided dtype '{d['dtype']}' is not supported. "
                             f"Supported types are '{sorted(cls.SUPPORTED_TYPES)}'.")
        if d["shape"] is None or not isinstance(d["shape"], list):
            raise ValueError("Invalid field: Feature shape can not be None and must be a list")
        return MetadataInfo(d["name"], cls.TO_NP_DTYPE[d["dtype"]], d["shape"], bool(d["isSparse"]))

    def get_features(self):
        return list(self._features)

    def get_labels(self):
        return list(self._labels)

    def get_label_names(self):
        return list(self._label_names)

    def get_feature_names(self):
        return list(self._feature_names)

    def get_feature_shape(self, feature_name):
        return next(x for x in self._features if x.name == feature_name).shape

    def get_tensors(self):
        return dict(self._tensors)
