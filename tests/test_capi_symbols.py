"""The C-ABI library loads without a GPU and exports every symbol include/gdmix_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "gdmix_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"GDMIX_API\s+[^;(]*?\b(gdmix_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    import __graft_entry__ as entry
    entry.build()
    from gdmix_b200 import _capi as capi
    declared = _declared_symbols()
    assert len(declared) >= 15, declared
    lib = ctypes.CDLL(capi.LIB_PATH)
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    # and the Python binding knows about each of them
    assert sorted(capi.SYMBOLS) == declared


def test_no_torch_types_in_the_abi():
    text = open(os.path.join(ROOT, "include", "gdmix_b200.h")).read()
    assert "torch" not in text.lower() and "at::" not in text and "#include <cuda" not in text


def test_version_and_error_string_without_gpu():
    from gdmix_b200 import _capi as capi
    assert b"sm_100a" in capi.lib.gdmix_version()
    # argument validation happens before any CUDA call
    rc = capi.lib.gdmix_partition_ids(None, None, ctypes.c_int64(1), ctypes.c_int32(0), None, None)
    assert rc == capi.GDMIX_ERR_INVALID and b"gdmix_partition_ids" in capi.lib.gdmix_last_error()


def test_partition_map_bit_exact_through_the_abi():
    """abs(String.hashCode) % n, JVM semantics (PartitionUtils.scala:31-37) -- product code vs golden values."""
    from gdmix_b200 import _capi as capi
    from tests.golden_util import load_partition
    g = load_partition()
    ids = list(g["hash"].keys())
    for n, table in g["partition"].items():
        h, p = capi.partition_ids(ids, int(n))
        assert [int(v) for v in h] == [g["hash"][s] for s in ids]
        assert [int(v) for v in p] == [table[s] for s in ids]
    h, p = capi.partition_ids(["polygenelubricants"], 10)
    assert int(h[0]) == -2147483648 and int(p[0]) == -8


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under gdmix_b200/ may import or load it."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "gdmix_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"(from|import)\s+oracle\b|oracle/|lr_oracle|scipy_port", src):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad
