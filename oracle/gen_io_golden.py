#!/usr/bin/env python
"""Pins the TF-free TFRecord / Example parser against files TensorFlow itself wrote.

Test infrastructure (like everything under oracle/): run in the build container, where /root/reference exists.

  * copies the reference's small DATA fixtures (no source code) into tests/golden/ref_fixtures/:
      gdmix-trainer/test/resources/grouped_per_member_train/{data.tfrecord, data.json, fake_feature_file.csv}
      gdmix-trainer/test/resources/fe_lbfgs/{training_data/test.tfrecord, metadata/tensor_metadata.json,
                                             featureList/global}
      gdmix-trainer/test/resources/validate/data.avro
  * decodes the two TFRecord files with google.protobuf (message classes built at run time from the published
    tf.train.Example / SequenceExample schema, i.e. independently of gdmix_b200.io.tfrecord) and writes what it
    finds to tests/golden/ref_fixtures/expected_tfrecord.json.  tests/test_io_formats.py compares
    gdmix_b200.io.tfrecord against that file.
"""
import json
import os
import shutil
import struct
import sys

from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

REF = "/root/reference/gdmix-trainer/test/resources"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "ref_fixtures")
COPIES = {
    "grouped_per_member_train/data.tfrecord": "re_data.tfrecord",
    "grouped_per_member_train/data.json": "re_data.json",
    "grouped_per_member_train/fake_feature_file.csv": "re_feature_file.csv",
    "fe_lbfgs/training_data/test.tfrecord": "fe_test.tfrecord",
    "fe_lbfgs/metadata/tensor_metadata.json": "fe_tensor_metadata.json",
    "fe_lbfgs/featureList/global": "fe_feature_list_global",
    "validate/data.avro": "validate_data.avro",
}


def _messages():
    """tf.train.{Example, SequenceExample} from tensorflow/core/example/{feature,example}.proto (public schema)."""
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name = "tf_example_restated.proto"
    fd.package = "tfx"
    fd.syntax = "proto3"
    T = descriptor_pb2.FieldDescriptorProto

    def msg(name):
        m = fd.message_type.add(); m.name = name; return m

    def field(m, name, number, ftype, label=T.LABEL_OPTIONAL, type_name=None, packed=None, oneof=None):
        f = m.field.add(); f.name = name; f.number = number; f.type = ftype; f.label = label
        if type_name: f.type_name = ".tfx." + type_name
        if packed is not None: f.options.packed = packed
        if oneof is not None: f.oneof_index = oneof
        return f

    m = msg("BytesList"); field(m, "value", 1, T.TYPE_BYTES, T.LABEL_REPEATED)
    m = msg("FloatList"); field(m, "value", 1, T.TYPE_FLOAT, T.LABEL_REPEATED, packed=True)
    m = msg("Int64List"); field(m, "value", 1, T.TYPE_INT64, T.LABEL_REPEATED, packed=True)
    m = msg("Feature"); m.oneof_decl.add().name = "kind"
    field(m, "bytes_list", 1, T.TYPE_MESSAGE, type_name="BytesList", oneof=0)
    field(m, "float_list", 2, T.TYPE_MESSAGE, type_name="FloatList", oneof=0)
    field(m, "int64_list", 3, T.TYPE_MESSAGE, type_name="Int64List", oneof=0)
    m = msg("Features")
    e = m.nested_type.add(); e.name = "FeatureEntry"; e.options.map_entry = True
    field(e, "key", 1, T.TYPE_STRING); field(e, "value", 2, T.TYPE_MESSAGE, type_name="Feature")
    field(m, "feature", 1, T.TYPE_MESSAGE, T.LABEL_REPEATED, type_name="Features.FeatureEntry")
    m = msg("FeatureList"); field(m, "feature", 1, T.TYPE_MESSAGE, T.LABEL_REPEATED, type_name="Feature")
    m = msg("FeatureLists")
    e = m.nested_type.add(); e.name = "FeatureListEntry"; e.options.map_entry = True
    field(e, "key", 1, T.TYPE_STRING); field(e, "value", 2, T.TYPE_MESSAGE, type_name="FeatureList")
    field(m, "feature_list", 1, T.TYPE_MESSAGE, T.LABEL_REPEATED, type_name="FeatureLists.FeatureListEntry")
    m = msg("Example"); field(m, "features", 1, T.TYPE_MESSAGE, type_name="Features")
    m = msg("SequenceExample")
    field(m, "context", 1, T.TYPE_MESSAGE, type_name="Features")
    field(m, "feature_lists", 2, T.TYPE_MESSAGE, type_name="FeatureLists")
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = lambda n: message_factory.GetMessageClass(pool.FindMessageTypeByName("tfx." + n))
    return get("Example"), get("SequenceExample")


def _records(path):
    buf = open(path, "rb").read()
    pos = 0
    while pos < len(buf):
        (n,) = struct.unpack_from("<Q", buf, pos)
        yield buf[pos + 12:pos + 12 + n]
        pos += 12 + n + 4


def _feature(f):
    kind = f.WhichOneof("kind")
    if kind == "bytes_list":
        return {"kind": "bytes", "values": [v.decode("latin-1") for v in f.bytes_list.value]}
    if kind == "float_list":
        return {"kind": "float", "values": [float(v) for v in f.float_list.value]}
    if kind == "int64_list":
        return {"kind": "int64", "values": [int(v) for v in f.int64_list.value]}
    return {"kind": None, "values": []}


def main():
    if not os.path.isdir(REF):
        sys.exit("reference tree not present: run this in the build container")
    os.makedirs(OUT, exist_ok=True)
    for src, dst in COPIES.items():
        shutil.copyfile(os.path.join(REF, src), os.path.join(OUT, dst))
        os.chmod(os.path.join(OUT, dst), 0o644)
    Example, SequenceExample = _messages()
    expected = {"re_data.tfrecord": [], "fe_test.tfrecord": []}
    for payload in _records(os.path.join(OUT, "re_data.tfrecord")):
        m = SequenceExample(); m.ParseFromString(payload)
        expected["re_data.tfrecord"].append({
            "context": {k: _feature(v) for k, v in sorted(m.context.feature.items())},
            "feature_lists": {k: [_feature(f) for f in v.feature] for k, v in sorted(m.feature_lists.feature_list.items())}})
    for payload in _records(os.path.join(OUT, "fe_test.tfrecord")):
        m = Example(); m.ParseFromString(payload)
        expected["fe_test.tfrecord"].append({k: _feature(v) for k, v in sorted(m.features.feature.items())})
    with open(os.path.join(OUT, "expected_tfrecord.json"), "w") as f:
        json.dump(expected, f)
    print({k: len(v) for k, v in expected.items()})


if __name__ == "__main__":
    main()
