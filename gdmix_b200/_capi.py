"""ctypes binding of include/gdmix_b200.h.

The library is the product: if ``lib/libgdmix_b200.so`` is missing or does not load, importing
this module raises -- there is no CPU fallback anywhere in the package.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GDMIX_B200_LIB") or os.path.join(_HERE, "lib", "libgdmix_b200.so")  # override: kernel experiments

GDMIX_OK = 0
GDMIX_ERR_INVALID = -1
GDMIX_ERR_CUDA = -2
GDMIX_ERR_WORKSPACE = -3
GDMIX_ERR_TOO_LARGE = -4
GDMIX_ERR_NO_DEVICE = -5
GDMIX_MAX_M = 32
GDMIX_MAX_SWEEP = 16

VARIANCE_NONE, VARIANCE_SIMPLE, VARIANCE_FULL = 0, 1, 2
EPS = float(np.finfo(np.float64).eps)

# every symbol include/gdmix_b200.h declares (checked by tests/test_capi_symbols.py)
SYMBOLS = ["gdmix_last_error", "gdmix_version", "gdmix_device_info", "gdmix_re_workspace_size",
           "gdmix_re_loss_grad", "gdmix_re_fit", "gdmix_re_score", "gdmix_fe_loss_grad", "gdmix_fe_score",
           "gdmix_fe_hessian",
           "gdmix_re_fit_host", "gdmix_re_score_host", "gdmix_host_register", "gdmix_host_unregister",
           "gdmix_host_release", "gdmix_partition_ids", "gdmix_lbfgs_create", "gdmix_lbfgs_iterate",
           "gdmix_lbfgs_info", "gdmix_lbfgs_destroy", "gdmix_launch_count", "gdmix_re_last_plan",
           "gdmix_partition_workspace_size", "gdmix_sort_pairs_u64", "gdmix_group_by_key", "gdmix_csr_gather_rows",
           "gdmix_gather_f32", "gdmix_partition_ids_i64", "gdmix_auc", "gdmix_re_fit_sweep",
           "gdmix_re_last_plan_typical", "gdmix_re_last_plan_small", "gdmix_local_index_mark", "gdmix_local_index_apply",
           "gdmix_seqex_count", "gdmix_seqex_fill", "gdmix_example_count", "gdmix_example_fill", "gdmix_avro_score_blocks", "gdmix_avro_model_blocks",
           "gdmix_feature_map_create", "gdmix_feature_map_destroy", "gdmix_avro_model_decode",
           "gdmix_fe_lbfgs_create", "gdmix_fe_lbfgs_reset", "gdmix_fe_lbfgs_step", "gdmix_fe_lbfgs_poll",
           "gdmix_fe_lbfgs_destroy", "gdmix_fe_column_counts", "gdmix_remap_i32", "gdmix_group_ids", "gdmix_offset_join", "gdmix_seqex_encode", "gdmix_local_index_host", "gdmix_fe_tile_plan_create",
           "gdmix_fe_tile_plan_destroy", "gdmix_fe_tile_plan_info", "gdmix_fe_loss_grad_tiled",
           "gdmix_pinned_alloc", "gdmix_pinned_free", "gdmix_narrow_columns", "gdmix_selftest_logistic",
           "gdmix_seqex_fill_local", "gdmix_avro_model_blocks_alloc", "gdmix_buffer_free"]


class SeqexSpec(C.Structure):
    _fields_ = [(n, C.c_char_p) for n in ("entity", "uid", "label", "offset", "weight", "bag_indices", "bag_values")]


class SeqexSizes(C.Structure):
    _fields_ = [("n_entities", C.c_int64), ("n_rows", C.c_int64), ("nnz", C.c_int64), ("id_bytes", C.c_int64),
                ("all_labelled", C.c_int32), ("saw_weight", C.c_int32), ("min_index", C.c_int64), ("max_index", C.c_int64)]


class ModelTable(C.Structure):
    _fields_ = [("n_models", C.c_int64), ("id_chars", C.c_void_p), ("id_ptr", C.c_void_p), ("model_class", C.c_char_p),
                ("coef", C.c_void_p), ("var", C.c_void_p), ("coef_ptr", C.c_void_p), ("feat_idx", C.c_void_p),
                ("has_intercept", C.c_int32), ("reserved", C.c_int32), ("threshold", C.c_double),
                ("intercept_name", C.c_char_p), ("name_chars", C.c_void_p), ("name_ptr", C.c_void_p),
                ("term_chars", C.c_void_p), ("term_ptr", C.c_void_p), ("n_features", C.c_int64)]


class GdmixError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"gdmix_b200 error {code}: {msg}")
        self.code = code


class ReBatch(C.Structure):
    _fields_ = [("n_entities", C.c_int64), ("n_rows", C.c_int64), ("nnz", C.c_int64),
                ("ent_rowptr", C.c_void_p), ("rowptr", C.c_void_p), ("col", C.c_void_p), ("val", C.c_void_p),
                ("label", C.c_void_p), ("weight", C.c_void_p), ("offset", C.c_void_p), ("theta_ptr", C.c_void_p),
                ("max_rows", C.c_int32), ("max_nnz", C.c_int32), ("max_coef", C.c_int32), ("reserved", C.c_int32),
                ("col16", C.c_void_p), ("col8", C.c_void_p), ("row_len16", C.c_void_p), ("label_bits", C.c_void_p)]


class LrOpts(C.Structure):
    _fields_ = [("l2", C.c_double), ("factr", C.c_double), ("pgtol", C.c_double),
                ("sparsity_threshold", C.c_double), ("regularize_bias", C.c_int32), ("has_intercept", C.c_int32),
                ("m", C.c_int32), ("max_iter", C.c_int32), ("max_ls", C.c_int32), ("max_fun", C.c_int32),
                ("variance_mode", C.c_int32), ("threads_per_entity", C.c_int32)]


class FeRows(C.Structure):
    _fields_ = [("n_rows", C.c_int64), ("nnz", C.c_int64), ("n_features", C.c_int64), ("rowptr", C.c_void_p),
                ("col", C.c_void_p), ("val", C.c_void_p), ("label", C.c_void_p), ("weight", C.c_void_p),
                ("offset", C.c_void_p), ("linear_regression", C.c_int32), ("num_workers", C.c_int32)]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a).  gdmix_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.gdmix_last_error.restype = C.c_char_p
    lib.gdmix_version.restype = C.c_char_p
    lib.gdmix_launch_count.restype = C.c_int64
    lib.gdmix_host_release.restype = None
    lib.gdmix_re_last_plan.restype = None
    lib.gdmix_re_last_plan_typical.restype = None
    lib.gdmix_re_last_plan_small.restype = None
    lib.gdmix_lbfgs_create.restype = C.c_void_p
    lib.gdmix_feature_map_create.restype = C.c_void_p
    lib.gdmix_feature_map_destroy.restype = None
    lib.gdmix_feature_map_destroy.argtypes = [C.c_void_p]
    lib.gdmix_lbfgs_create.argtypes = [C.c_int64, C.c_void_p]
    lib.gdmix_lbfgs_iterate.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
    lib.gdmix_lbfgs_info.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.gdmix_lbfgs_destroy.argtypes = [C.c_void_p]
    lib.gdmix_lbfgs_destroy.restype = None
    lib.gdmix_fe_tile_plan_create.restype = C.c_void_p
    lib.gdmix_fe_tile_plan_create.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_void_p]
    lib.gdmix_fe_tile_plan_destroy.argtypes = [C.c_void_p]
    lib.gdmix_fe_tile_plan_destroy.restype = None
    lib.gdmix_fe_tile_plan_info.argtypes = [C.c_void_p, C.c_void_p]
    lib.gdmix_fe_loss_grad_tiled.argtypes = [C.c_void_p] * 6
    lib.gdmix_fe_lbfgs_create.restype = C.c_void_p
    lib.gdmix_fe_lbfgs_create.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.gdmix_fe_lbfgs_reset.argtypes = [C.c_void_p, C.c_void_p]
    lib.gdmix_fe_lbfgs_step.argtypes = [C.c_void_p, C.c_void_p]
    lib.gdmix_fe_lbfgs_poll.argtypes = [C.c_void_p] * 7
    lib.gdmix_fe_lbfgs_destroy.argtypes = [C.c_void_p]
    lib.gdmix_fe_lbfgs_destroy.restype = None
    for name in SYMBOLS:
        getattr(lib, name)  # AttributeError if the build is stale
    return lib


lib = _load()


def check(rc):
    if rc != GDMIX_OK:
        raise GdmixError(rc, lib.gdmix_last_error().decode())


def make_opts(l2=1.0, regularize_bias=False, has_intercept=True, m=10, max_iter=100, tol=1e-12, factr=None,
              pgtol=1e-5, max_ls=20, max_fun=15000, sparsity_threshold=0.0, variance_mode=VARIANCE_NONE,
              threads_per_entity=0):
    """Defaults are what the reference hands to scipy (random_effect_lr_lbfgs_model.py:142-146); pgtol,
    maxls and maxfun are scipy's own defaults because the reference never sets them."""
    if factr is None:
        factr = tol / EPS
    return LrOpts(float(l2), float(factr), float(pgtol), float(sparsity_threshold), int(bool(regularize_bias)),
                  int(bool(has_intercept)), int(m), int(max_iter), int(max_ls), int(max_fun), int(variance_mode),
                  int(threads_per_entity))


def device_info():
    sm, smem, cc = C.c_int32(), C.c_int32(), C.c_int32()
    check(lib.gdmix_device_info(C.byref(sm), C.byref(smem), C.byref(cc)))
    return {"sm_count": sm.value, "smem_per_block_optin": smem.value, "cc": cc.value}


def last_plan():
    """Plan of this thread's latest random-effect launch (gdmix_re_last_plan)."""
    a = (C.c_int32 * 8)()
    lib.gdmix_re_last_plan(a)
    keys = ["fast", "threads", "ept", "ctas_per_sm", "cap_steps", "smem", "grid", "hist_global"]
    d = dict(zip(keys, list(a)))
    lib.gdmix_re_last_plan_typical(a)
    if a[0]:
        d["typical"] = dict(zip(["threads", "ept", "ctas_per_sm", "cap_steps", "smem", "max_rows"], list(a)[1:7]))
    lib.gdmix_re_last_plan_small(a)
    if a[0]:
        d["small"] = dict(zip(["warps_per_cta", "slots", "ctas_per_sm", "cap_rows", "cap_nnz", "smem", "grid"], list(a)[1:8]))
    return d


def launch_count():
    return int(lib.gdmix_launch_count())


def _np_ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class HostBatch:
    """A batch of entities in host memory (numpy), entity-local CSR -- gdmix_re_batch with host pointers."""

    def __init__(self, ent_rowptr, rowptr, col, val, label, weight=None, offset=None, theta_ptr=None,
                 has_intercept=True, col_narrow=None):
        """col: int32 entity-local column indices, or None when col_narrow (uint8 / uint16, what crosses PCIe) is
        given ready-made (the fused reader produces it directly)."""
        self.ent_rowptr = np.ascontiguousarray(ent_rowptr, dtype=np.int64)
        self.rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
        if col is None and col_narrow is None:
            raise ValueError("HostBatch needs col or col_narrow")
        self.col = None if col is None else np.ascontiguousarray(col, dtype=np.int32)
        self._col_narrow = None
        if col_narrow is not None:
            if col_narrow.dtype not in (np.uint8, np.uint16):
                raise ValueError("col_narrow must be uint8 or uint16")
            self._col_narrow = np.ascontiguousarray(col_narrow)
        self.val = np.ascontiguousarray(val, dtype=np.float32)
        self.label = np.ascontiguousarray(label, dtype=np.float32)
        self.weight = None if weight is None else np.ascontiguousarray(weight, dtype=np.float32)
        self.offset = None if offset is None else np.ascontiguousarray(offset, dtype=np.float32)
        self.has_intercept = bool(has_intercept)
        self.n_entities = len(self.ent_rowptr) - 1
        self.n_rows = int(self.ent_rowptr[-1]) if self.n_entities >= 0 else 0
        self.nnz = int(self.rowptr[self.n_rows])
        assert self.rowptr.shape[0] == self.n_rows + 1
        assert (self.col if self.col is not None else self._col_narrow).shape[0] >= self.nnz
        self.theta_ptr = np.ascontiguousarray(theta_ptr, dtype=np.int64)
        assert self.theta_ptr.shape[0] == self.n_entities + 1
        rows = np.diff(self.ent_rowptr)
        nnz_e = self.rowptr[self.ent_rowptr[1:]] - self.rowptr[self.ent_rowptr[:-1]]
        coef = np.diff(self.theta_ptr)
        self.max_rows = int(rows.max()) if self.n_entities else 0
        self.max_nnz = int(nnz_e.max()) if self.n_entities else 0
        self.max_coef = int(coef.max()) if self.n_entities else 0
        self.n_coef = int(self.theta_ptr[-1])

    def c_struct(self, narrow=None, narrow_rows=False):
        """narrow=True sends the local column indices as uint16 (half the PCIe bytes for them); default: whenever
        every entity has fewer than 65536 local features.  narrow_rows=True also builds the 16-bit row lengths and
        the label bits gdmix_re_fit_host can take instead of the int64 row pointers / fp32 labels (-5 % of the bytes:
        worth it for a caller that keeps the batch, not for one numpy pass per partition)."""
        if narrow is None:
            narrow = (self.max_coef < 65536 and self.nnz > 0) or self.col is None
        if not narrow and self.col is None:
            raise ValueError("this HostBatch only has narrow column indices")
        c16 = c8 = None
        if narrow:
            if getattr(self, "_col_narrow", None) is None:
                # one byte per index when every entity has at most 256 local features, else two
                # (an entity's local indices are < its coefficient count minus the intercept)
                width = 1 if self.max_coef - (1 if self.has_intercept else 0) <= 256 else 2
                try:
                    self._col_narrow = narrow_columns(self.col[:self.nnz], width)
                except GdmixError:
                    if width == 2:
                        raise
                    self._col_narrow = narrow_columns(self.col[:self.nnz], 2)
            c8 = self._col_narrow if self._col_narrow.dtype == np.uint8 else None
            c16 = self._col_narrow if c8 is None else None
        rl16 = bits = None
        if narrow and narrow_rows and self.n_rows > 0:
            # what else gdmix_re_fit_host can take narrow: 16-bit row lengths instead of the int64 row pointers
            # (-6 bytes per row on PCIe) and the 0/1 labels as bits (-3.9 bytes per row)
            if getattr(self, "_rows_narrow", None) is None:
                lens = np.diff(self.rowptr)
                r16 = None
                if lens.size and int(lens.max()) <= 65535:
                    r16 = pinned_empty(lens.shape[0], np.uint16)
                    r16[:] = lens
                lab = self.label
                b = None
                if lab is not None and bool(np.all((lab == 0) | (lab == 1))):
                    packed = np.packbits(lab != 0, bitorder="little")
                    b = pinned_empty(packed.shape[0] + 1, np.uint8)
                    b[:packed.shape[0]] = packed
                    b[packed.shape[0]:] = 0
                self._rows_narrow = (r16, b)
            rl16, bits = self._rows_narrow
        return ReBatch(self.n_entities, self.n_rows, self.nnz, _np_ptr(self.ent_rowptr), _np_ptr(self.rowptr),
                       _np_ptr(self.col), _np_ptr(self.val), _np_ptr(self.label), _np_ptr(self.weight),
                       _np_ptr(self.offset), _np_ptr(self.theta_ptr), self.max_rows, self.max_nnz, self.max_coef, 0,
                       _np_ptr(c16), _np_ptr(c8), _np_ptr(rl16), _np_ptr(bits))

    def algorithmic_bytes(self, warm=False):
        """8 B/nnz + 16 B/sample + 8 B/coef out (+8 in when warm) + 4 B/feature index map."""
        hi_total = self.n_coef
        return 8 * self.nnz + 16 * self.n_rows + 8 * hi_total * (2 if warm else 1) + 4 * hi_total


def re_fit_host(batch, opts, theta0=None, want_variance=False, chunk_entities=0):
    """gdmix_re_fit_host: numpy in, numpy out (theta, f, nit, nfev, status[, variance])."""
    E, T = batch.n_entities, batch.n_coef
    theta = pinned_empty(T, np.float64)      # D2H targets: page-locked, from the library's pool; every coefficient of
    #                                           every entity is written by the call (or it fails), so no zero fill
    f = np.zeros(E, np.float64)
    nit = np.zeros(E, np.int32)
    nfev = np.zeros(E, np.int32)
    status = np.zeros(E, np.int32)
    var = None
    if want_variance:
        var = pinned_empty(T, np.float64)
    t0 = None if theta0 is None else np.ascontiguousarray(theta0, dtype=np.float64)
    cb = batch.c_struct()
    check(lib.gdmix_re_fit_host(C.byref(cb), C.byref(opts), _np_ptr(t0), _np_ptr(theta), _np_ptr(f), _np_ptr(nit),
                                _np_ptr(nfev), _np_ptr(status), _np_ptr(var), C.c_int64(chunk_entities)))
    out = {"theta": theta, "f": f, "nit": nit, "nfev": nfev, "status": status}
    if want_variance:
        out["variance"] = var
    return out


def re_score_host(batch, opts, theta, has_model=None):
    logit = pinned_empty(batch.n_rows, np.float32)       # every row is written by the call
    per = pinned_empty(batch.n_rows, np.float32)
    th = None if theta is None else np.ascontiguousarray(theta, dtype=np.float64)
    hm = None if has_model is None else np.ascontiguousarray(has_model, dtype=np.uint8)
    cb = batch.c_struct()
    check(lib.gdmix_re_score_host(C.byref(cb), C.byref(opts), _np_ptr(th), _np_ptr(hm), _np_ptr(logit),
                                  _np_ptr(per)))
    return logit, per


def partition_ids(ids, num_partitions):
    """abs(String.hashCode(id)) % num_partitions for a list of str ids -> (hash int32[], partition int32[])."""
    enc = [np.frombuffer(s.encode("utf-16-le"), dtype=np.uint16) for s in ids]
    ptr = np.zeros(len(ids) + 1, np.int64)
    if ids:
        ptr[1:] = np.cumsum([len(u) for u in enc])
    units = np.ascontiguousarray(np.concatenate(enc)) if ids and ptr[-1] > 0 else np.zeros(1, np.uint16)
    h = np.zeros(len(ids), np.int32)
    p = np.zeros(len(ids), np.int32)
    check(lib.gdmix_partition_ids(_np_ptr(units), _np_ptr(ptr), C.c_int64(len(ids)), C.c_int32(num_partitions),
                                  _np_ptr(h), _np_ptr(p)))
    return h, p


# ---- device-pointer entry points (torch tensors supply memory and the stream; plumbing only) ------

def _tptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class DeviceBatch:
    """gdmix_re_batch over torch CUDA tensors."""

    def __init__(self, host: HostBatch, device="cuda"):
        import torch
        self.host = host
        to = lambda a, pin=False: None if a is None else torch.from_numpy(a).to(device, non_blocking=False)
        self.ent_rowptr, self.rowptr, self.theta_ptr = to(host.ent_rowptr), to(host.rowptr), to(host.theta_ptr)
        self.col, self.val, self.label = to(host.col), to(host.val), to(host.label)
        self.weight, self.offset = to(host.weight), to(host.offset)

    def c_struct(self):
        h = self.host
        return ReBatch(h.n_entities, h.n_rows, h.nnz, _tptr(self.ent_rowptr), _tptr(self.rowptr), _tptr(self.col),
                       _tptr(self.val), _tptr(self.label), _tptr(self.weight), _tptr(self.offset),
                       _tptr(self.theta_ptr), h.max_rows, h.max_nnz, h.max_coef, 0, None)


def re_workspace_size(cb, opts):
    n = C.c_size_t()
    check(lib.gdmix_re_workspace_size(C.byref(cb), C.byref(opts), C.byref(n)))
    return n.value


def _stream_ptr(stream):
    import torch
    s = torch.cuda.current_stream() if stream is None else stream
    return C.c_void_p(s.cuda_stream)


def re_fit_device(dbatch, opts, theta0=None, want_variance=False, workspace=None, out=None, stream=None):
    """gdmix_re_fit on device tensors; asynchronous on the (current) torch stream."""
    import torch
    h = dbatch.host
    dev = dbatch.val.device
    cb = dbatch.c_struct()
    need = re_workspace_size(cb, opts)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(max(need, 256), dtype=torch.uint8, device=dev)
    if out is None:
        out = {"theta": torch.empty(h.n_coef, dtype=torch.float64, device=dev),
               "f": torch.empty(h.n_entities, dtype=torch.float64, device=dev),
               "nit": torch.empty(h.n_entities, dtype=torch.int32, device=dev),
               "nfev": torch.empty(h.n_entities, dtype=torch.int32, device=dev),
               "status": torch.empty(h.n_entities, dtype=torch.int32, device=dev)}
        if want_variance:
            out["variance"] = torch.empty(h.n_coef, dtype=torch.float64, device=dev)
    check(lib.gdmix_re_fit(C.byref(cb), C.byref(opts), _tptr(theta0), _tptr(out["theta"]), _tptr(out["f"]),
                           _tptr(out["nit"]), _tptr(out["nfev"]), _tptr(out["status"]),
                           _tptr(out.get("variance")), _tptr(workspace), C.c_size_t(workspace.numel()),
                           _stream_ptr(stream)))
    out["workspace"] = workspace
    return out


def re_fit_sweep_device(dbatch, opts, l2_values, theta0=None, workspace=None, stream=None):
    """gdmix_re_fit_sweep: len(l2_values) models per entity from one staged copy of each entity's block.
    Returns tensors with a leading axis over l2_values: theta [n_l2, n_coef], f / nit / nfev / status [n_l2, E]."""
    import torch
    h = dbatch.host
    dev = dbatch.val.device
    cb = dbatch.c_struct()
    l2 = np.ascontiguousarray(l2_values, dtype=np.float64)
    n_l2 = int(l2.shape[0])
    need = re_workspace_size(cb, opts)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(max(need, 256), dtype=torch.uint8, device=dev)
    E = h.n_entities
    out = {"theta": torch.empty((n_l2, h.n_coef), dtype=torch.float64, device=dev),
           "f": torch.empty((n_l2, E), dtype=torch.float64, device=dev),
           "nit": torch.empty((n_l2, E), dtype=torch.int32, device=dev),
           "nfev": torch.empty((n_l2, E), dtype=torch.int32, device=dev),
           "status": torch.empty((n_l2, E), dtype=torch.int32, device=dev)}
    check(lib.gdmix_re_fit_sweep(C.byref(cb), C.byref(opts), _np_ptr(l2), C.c_int32(n_l2), _tptr(theta0),
                                 _tptr(out["theta"]), C.c_int64(h.n_coef), _tptr(out["f"]), _tptr(out["nit"]),
                                 _tptr(out["nfev"]), _tptr(out["status"]), _tptr(workspace),
                                 C.c_size_t(workspace.numel()), _stream_ptr(stream)))
    out["workspace"] = workspace
    return out


def re_loss_grad_device(dbatch, opts, theta, stream=None):
    import torch
    h = dbatch.host
    dev = dbatch.val.device
    cb = dbatch.c_struct()
    need = re_workspace_size(cb, opts)
    ws = torch.empty(max(need, 256), dtype=torch.uint8, device=dev)
    f = torch.empty(h.n_entities, dtype=torch.float64, device=dev)
    g = torch.empty(h.n_coef, dtype=torch.float64, device=dev)
    check(lib.gdmix_re_loss_grad(C.byref(cb), C.byref(opts), _tptr(theta), _tptr(f), _tptr(g), _tptr(ws),
                                 C.c_size_t(ws.numel()), _stream_ptr(stream)))
    return f, g


def re_score_device(dbatch, opts, theta, has_model=None, stream=None):
    import torch
    h = dbatch.host
    dev = dbatch.val.device
    logit = torch.empty(h.n_rows, dtype=torch.float32, device=dev)
    per = torch.empty(h.n_rows, dtype=torch.float32, device=dev)
    cb = dbatch.c_struct()
    check(lib.gdmix_re_score(C.byref(cb), C.byref(opts), _tptr(theta), _tptr(has_model), _tptr(logit), _tptr(per),
                             _stream_ptr(stream)))
    return logit, per


class DeviceFeRows:
    def __init__(self, rowptr, col, val, label, weight, offset, n_features, linear_regression=False, num_workers=1,
                 device="cuda"):
        import torch
        to = lambda a, dt: None if a is None else torch.as_tensor(np.ascontiguousarray(a, dtype=dt)).to(device)
        self.rowptr, self.col, self.val = to(rowptr, np.int64), to(col, np.int32), to(val, np.float32)
        self.label, self.weight, self.offset = to(label, np.float32), to(weight, np.float32), to(offset, np.float32)
        self.n_rows = int(self.rowptr.numel() - 1)
        self.nnz = int(self.col.numel())
        self.n_features = int(n_features)
        self.linear_regression, self.num_workers = bool(linear_regression), int(num_workers)

    def c_struct(self):
        return FeRows(self.n_rows, self.nnz, self.n_features, _tptr(self.rowptr), _tptr(self.col), _tptr(self.val),
                      _tptr(self.label), _tptr(self.weight), _tptr(self.offset), int(self.linear_regression),
                      self.num_workers)


FE_HEAD = 8192   # above this many features the solver renumbers them by falling frequency (hot ranks first)


class DeviceFeTilePlan:
    """gdmix_fe_tile_plan_*: the shard laid out for the tiled objective (row-major hot / cold copy for z = X x,
    column-major tiled copy for g = X^T dz), built on the device by the library once per training run.
    hz / hg / tile_rows / l2_tile_rows: 0 = the library's choice (as much of x / of the gradient in shared memory as
    the kernels hold); the tests pass small values to exercise the cold paths."""

    def __init__(self, rows, hz=0, hg=0, tile_rows=0, l2_tile_rows=0, stream=None):
        cs = rows.c_struct()
        self._rows = rows
        self._h = lib.gdmix_fe_tile_plan_create(C.cast(C.pointer(cs), C.c_void_p), C.c_int32(hz), C.c_int32(hg),
                                                C.c_int32(tile_rows), C.c_int64(l2_tile_rows), _stream_ptr(stream))
        if not self._h:
            raise GdmixError(GDMIX_ERR_INVALID, lib.gdmix_last_error().decode())
        a = (C.c_int64 * 8)()
        check(lib.gdmix_fe_tile_plan_info(self._h, C.cast(a, C.c_void_p)))
        (self.hz, self.hg, self.tile_rows, self.n_tiles, self.n_cold_z, self.n_cold_g, self.bytes,
         self.n_l2_tiles) = [int(v) for v in a]

    def close(self):
        if self._h:
            lib.gdmix_fe_tile_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def fe_column_counts(rows, stream=None):
    """Non-zeros per feature of this rank's shard -> int64 CUDA tensor [D] (gdmix_fe_column_counts)."""
    import torch
    counts = torch.empty(rows.n_features, dtype=torch.int64, device=rows.val.device)
    check(lib.gdmix_fe_column_counts(_tptr(rows.col), C.c_int64(rows.nnz), C.c_int64(rows.n_features), _tptr(counts),
                                     _stream_ptr(stream)))
    return counts


def rank_by_count(counts, stream=None):
    """Features by falling count, ties by feature id -> int32 CUDA tensor (the library's stable radix sort)."""
    import torch
    n = counts.numel()
    mx = int(counts.max().item()) if n else 0
    keys = (mx - counts).contiguous()
    b = C.c_size_t()
    check(lib.gdmix_partition_workspace_size(C.c_int64(max(n, 1)), C.byref(b)))
    ws = torch.empty(b.value, dtype=torch.uint8, device=counts.device)
    out = torch.empty_like(keys)
    perm = torch.empty(n, dtype=torch.int32, device=counts.device)
    check(lib.gdmix_sort_pairs_u64(_tptr(keys), C.c_int64(n), C.c_int32(max(1, mx.bit_length())), _tptr(out), _tptr(perm),
                                   _tptr(ws), C.c_size_t(ws.numel()), _stream_ptr(stream)))
    return perm


def remap_columns(col, rank_of, stream=None):
    """rank_of[col] as a new int32 tensor (gdmix_remap_i32)."""
    import torch
    out = torch.empty_like(col)
    check(lib.gdmix_remap_i32(_tptr(col), _tptr(rank_of), C.c_int64(col.numel()), _tptr(out), _stream_ptr(stream)))
    return out


def fe_loss_grad_device(rows, opts, x, fg=None, stream=None, plan=None):
    """-> fg tensor [1 + D + has_intercept]: value then gradient (this rank's partial).  With a DeviceFeTilePlan the
    atomics-free tiled path runs; without, the single-pass kernel with fp64 atomics (tests, one-off evaluations)."""
    import torch
    n = 1 + rows.n_features + (1 if opts.has_intercept else 0)
    if fg is None:
        fg = torch.empty(n, dtype=torch.float64, device=x.device)
    cs = rows.c_struct()
    if plan is not None:
        check(lib.gdmix_fe_loss_grad_tiled(C.cast(C.pointer(cs), C.c_void_p), plan._h,
                                           C.cast(C.pointer(opts), C.c_void_p), _tptr(x), _tptr(fg),
                                           _stream_ptr(stream)))
    else:
        check(lib.gdmix_fe_loss_grad(C.byref(cs), C.byref(opts), _tptr(x), _tptr(fg), _stream_ptr(stream)))
    return fg


def fe_hessian_device(rows, opts, x, mode, stream=None):
    """-> h tensor: [D+hi] (SIMPLE) or [(D+hi), (D+hi)] (FULL); this rank's partial sum."""
    import torch
    P = rows.n_features + (1 if opts.has_intercept else 0)
    h = torch.empty(P * P if mode == VARIANCE_FULL else P, dtype=torch.float64, device=x.device)
    cs = rows.c_struct()
    check(lib.gdmix_fe_hessian(C.byref(cs), C.byref(opts), _tptr(x), C.c_int32(mode), _tptr(h), _stream_ptr(stream)))
    return h.view(P, P) if mode == VARIANCE_FULL else h


def fe_score_device(rows, opts, x, stream=None):
    import torch
    logit = torch.empty(rows.n_rows, dtype=torch.float32, device=x.device)
    per = torch.empty(rows.n_rows, dtype=torch.float32, device=x.device)
    cs = rows.c_struct()
    check(lib.gdmix_fe_score(C.byref(cs), C.byref(opts), _tptr(x), _tptr(logit), _tptr(per), _stream_ptr(stream)))
    return logit, per


class DeviceLbfgs:
    """gdmix_fe_lbfgs_*: the replicated L-BFGS-B state of the fixed-effect solve, resident on the device.  `x` and `fg`
    are the caller's CUDA tensors (float64, [n] and [1 + n]); step() only enqueues kernels on the stream."""

    DONE, NEED_FG, AGAIN = 0, 1, 2

    def __init__(self, x, fg, opts):
        n = x.numel()
        assert fg.numel() == n + 1 and x.is_cuda and fg.is_cuda
        self._keep = (x, fg)
        self._h = lib.gdmix_fe_lbfgs_create(C.c_int64(n), C.cast(C.pointer(opts), C.c_void_p), _tptr(x), _tptr(fg))
        if not self._h:
            raise GdmixError(GDMIX_ERR_INVALID, lib.gdmix_last_error().decode())

    def reset(self, stream=None):
        check(lib.gdmix_fe_lbfgs_reset(self._h, _stream_ptr(stream)))

    def step(self, stream=None):
        check(lib.gdmix_fe_lbfgs_step(self._h, _stream_ptr(stream)))

    def poll(self, stream=None):
        """Synchronises the stream.  -> dict(task, nit, nfev, status, f)."""
        task, nit, nfev, status, f = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32(), C.c_double()
        check(lib.gdmix_fe_lbfgs_poll(self._h, _stream_ptr(stream), C.cast(C.byref(task), C.c_void_p),
                                      C.cast(C.byref(nit), C.c_void_p), C.cast(C.byref(nfev), C.c_void_p),
                                      C.cast(C.byref(status), C.c_void_p), C.cast(C.byref(f), C.c_void_p)))
        return {"task": task.value, "nit": nit.value, "nfev": nfev.value, "status": status.value, "f": f.value}

    def close(self):
        if self._h:
            lib.gdmix_fe_lbfgs_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class HostLbfgs:
    """gdmix_lbfgs_*: the replicated L-BFGS-B state of the fixed-effect solve (host side, reverse communication)."""

    NEED_FG, DONE = 1, 0

    def __init__(self, n, opts):
        self._h = lib.gdmix_lbfgs_create(C.c_int64(n), C.cast(C.pointer(opts), C.c_void_p))
        if not self._h:
            raise GdmixError(GDMIX_ERR_INVALID, lib.gdmix_last_error().decode())

    def iterate(self, x, f, g):
        """x: float64 numpy array, updated in place; g: float64 numpy array.  -> NEED_FG or DONE."""
        assert x.dtype == np.float64 and g.dtype == np.float64 and x.flags.c_contiguous and g.flags.c_contiguous
        rc = lib.gdmix_lbfgs_iterate(self._h, _np_ptr(x), C.c_double(f), _np_ptr(g))
        if rc < 0:
            check(rc)
        return rc

    def info(self):
        nit, nfev, st, f = C.c_int32(), C.c_int32(), C.c_int32(), C.c_double()
        check(lib.gdmix_lbfgs_info(self._h, C.byref(nit), C.byref(nfev), C.byref(st), C.byref(f)))
        return {"nit": nit.value, "nfev": nfev.value, "status": st.value, "f": f.value}

    def close(self):
        if self._h:
            lib.gdmix_lbfgs_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def parse_entity_grouped(file_image, entity, uid, label, offset, weight, bag_indices, bag_values):
    """One uncompressed TFRecord file image of entity-grouped SequenceExamples -> dict of flat numpy arrays
    (gdmix_seqex_count + gdmix_seqex_fill; host code of the library, no device involved)."""
    enc = lambda x: None if x is None else x.encode("utf-8")
    spec = SeqexSpec(enc(entity), enc(uid), enc(label), enc(offset), enc(weight), enc(bag_indices), enc(bag_values))
    buf = np.frombuffer(file_image, dtype=np.uint8)
    sz = SeqexSizes()
    check(lib.gdmix_seqex_count(_np_ptr(buf), C.c_int64(buf.size), C.byref(spec), C.byref(sz)))
    E, N, Z = sz.n_entities, sz.n_rows, sz.nnz
    out = {"ent_rows": np.zeros(E, np.int64), "row_len": np.zeros(N, np.int64), "gcol": np.zeros(Z, np.int64),
           "val": np.zeros(Z, np.float32), "uid": np.zeros(N, np.int64), "label": np.zeros(N, np.float32),
           "offset": np.zeros(N, np.float32), "weight": np.zeros(N, np.float32),
           "id_chars": np.zeros(max(sz.id_bytes, 1), np.uint8), "id_ptr": np.zeros(E + 1, np.int64)}
    check(lib.gdmix_seqex_fill(_np_ptr(buf), C.c_int64(buf.size), C.byref(spec), _np_ptr(out["ent_rows"]),
                               _np_ptr(out["row_len"]), _np_ptr(out["gcol"]), _np_ptr(out["val"]), _np_ptr(out["uid"]),
                               _np_ptr(out["label"]), _np_ptr(out["offset"]), _np_ptr(out["weight"]),
                               _np_ptr(out["id_chars"]), _np_ptr(out["id_ptr"]), None))
    raw = out["id_chars"].tobytes()
    ip = out["id_ptr"]
    out["entity_ids"] = [raw[ip[e]:ip[e + 1]].decode("utf-8") for e in range(E)]
    out["all_labelled"], out["saw_weight"] = bool(sz.all_labelled), bool(sz.saw_weight)
    return out


class _PinnedPool:
    """Page-locked host blocks (gdmix_pinned_alloc), kept for reuse: a worker trains partition after partition and
    pinning a few GB costs about as long as copying them.  A block returns here when the last numpy view of it dies."""

    def __init__(self):
        import threading
        self.lock = threading.Lock()
        self.free = []          # (capacity, pointer)
        self.available = None   # None: not tried yet; False: no CUDA device (host-only use of the readers)
        self.limit = int(os.environ.get("GDMIX_PINNED_POOL_MB", "24576")) << 20

    def take(self, nbytes):
        """-> (pointer, capacity) or None when page-locked memory cannot be had."""
        if self.available is False:
            return None
        with self.lock:
            best = None
            for k, (cap, _) in enumerate(self.free):
                if nbytes <= cap <= 2 * nbytes + (1 << 20) and (best is None or cap < self.free[best][0]):
                    best = k
            if best is not None:
                cap, ptr = self.free.pop(best)
                return ptr, cap
        cap = (max(int(nbytes), 1) + (1 << 21) - 1) & ~((1 << 21) - 1)
        out = C.c_void_p()
        rc = lib.gdmix_pinned_alloc(C.c_size_t(cap), C.byref(out))
        if rc != 0 or not out.value:
            if self.available is None:
                self.available = False
            return None
        self.available = True
        return out.value, cap

    def give(self, ptr, cap):
        with self.lock:
            self.free.append((cap, ptr))
            total = sum(c for c, _ in self.free)
            while total > self.limit and self.free:
                c, p = self.free.pop(0)
                lib.gdmix_pinned_free(C.c_void_p(p))
                total -= c

    def release(self):
        with self.lock:
            for _, p in self.free:
                lib.gdmix_pinned_free(C.c_void_p(p))
            self.free = []


_pinned_pool = _PinnedPool()


class _PinnedBlock:
    """Owner of one pool block behind a numpy array (array.base): gives the block back when collected."""

    def __init__(self, ptr, cap, n, dtype):
        self.ptr, self.cap = ptr, cap
        self.__array_interface__ = {"data": (ptr, False), "shape": (int(n),), "typestr": np.dtype(dtype).str, "version": 3}

    def __del__(self):
        try:
            _pinned_pool.give(self.ptr, self.cap)
        except Exception:      # interpreter shutdown
            pass


class _LibBuffer:
    """Owner of a malloc'ed buffer the library returned (gdmix_buffer_free when collected), as a uint8 array."""

    def __init__(self, ptr, n):
        self.ptr = ptr
        self.__array_interface__ = {"data": (ptr, False), "shape": (int(n),), "typestr": "|u1", "version": 3}

    def __del__(self):
        try:
            lib.gdmix_buffer_free(C.c_void_p(self.ptr))
        except Exception:
            pass


def pinned_empty(n, dtype):
    """np.empty(n, dtype) in page-locked memory of the library's pool (plain np.empty when there is no CUDA device:
    the readers are host code and are used without one)."""
    dtype = np.dtype(dtype)
    got = _pinned_pool.take(int(n) * dtype.itemsize)
    if got is None:
        return np.empty(int(n), dtype)
    return np.asarray(_PinnedBlock(got[0], got[1], n, dtype))


def narrow_columns(col, width):
    """int32 local column indices -> uint8 (width 1) / uint16 (width 2) in pinned memory, all host threads."""
    col = np.ascontiguousarray(col, dtype=np.int32)
    out = pinned_empty(col.shape[0], np.uint8 if width == 1 else np.uint16)
    check(lib.gdmix_narrow_columns(_np_ptr(col), C.c_int64(col.shape[0]), C.c_int32(width), _np_ptr(out)))
    return out


def local_index_host(ent_rowptr, rowptr, gcol):
    """np.unique per entity over a parsed partition (gdmix_local_index_host, all host threads).
    -> (local int32[nnz], d_e int64[E], uniq_ptr int64[E+1], uniq_global int64[sum d_e])"""
    ent_rowptr = np.ascontiguousarray(ent_rowptr, dtype=np.int64)
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
    gcol = np.ascontiguousarray(gcol, dtype=np.int64)
    E = ent_rowptr.shape[0] - 1
    nnz = int(rowptr[ent_rowptr[-1]]) if E >= 0 else 0
    local = np.empty(max(nnz, 1), np.int32)
    d_e = np.zeros(max(E, 1), np.int64)
    scratch = np.empty(max(nnz, 1), np.int64)
    check(lib.gdmix_local_index_host(_np_ptr(ent_rowptr), _np_ptr(rowptr), _np_ptr(gcol), C.c_int64(E), _np_ptr(local),
                                     _np_ptr(d_e), _np_ptr(scratch), None, None))
    d_e = d_e[:E]
    uniq_ptr = np.zeros(E + 1, np.int64)
    np.cumsum(d_e, out=uniq_ptr[1:])
    uniq_global = np.empty(max(int(uniq_ptr[-1]), 1), np.int64)
    check(lib.gdmix_local_index_host(_np_ptr(ent_rowptr), _np_ptr(rowptr), _np_ptr(gcol), C.c_int64(E), None,
                                     _np_ptr(np.ascontiguousarray(d_e)), _np_ptr(scratch), _np_ptr(uniq_ptr), _np_ptr(uniq_global)))
    return local[:nnz], d_e, uniq_ptr, uniq_global[:int(uniq_ptr[-1])]


def encode_entity_grouped(ent_rows, row_len, gcol, val, uid, entity_int=None, entity_str=None, label=None,
                          label_as_int=True, offset=None, weight=None, entity="entity", uid_name="uid",
                          label_name="response", offset_name="offset", weight_name="weight", bag=None):
    """Flat arrays -> the bytes of one uncompressed TFRecord file of entity-grouped SequenceExamples
    (gdmix_seqex_encode; host code of the library).  Record e = ent_rows[e] consecutive samples; the entity id is
    entity_int[e] (int64) or entity_str[e].  bag: name of the feature bag (columns <bag>_indices / <bag>_values), None:
    no features are written."""
    enc = lambda x: None if x is None else x.encode("utf-8")
    E = int(len(ent_rows))
    ent_rows = np.ascontiguousarray(ent_rows, dtype=np.int64)
    N = int(ent_rows.sum())
    row_len = None if bag is None else np.ascontiguousarray(row_len, dtype=np.int64)
    gcol = None if bag is None else np.ascontiguousarray(gcol, dtype=np.int64)
    val = None if bag is None else np.ascontiguousarray(val, dtype=np.float32)
    uid = np.ascontiguousarray(uid, dtype=np.int64)
    assert uid.shape[0] == N and (bag is None or row_len.shape[0] == N)
    f32 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float32)
    label, offset, weight = f32(label), f32(offset), f32(weight)
    ei = idc = idp = None
    if entity_int is not None:
        ei = np.ascontiguousarray(entity_int, dtype=np.int64)
    else:
        idc, idp = _string_table([str(x) for x in entity_str])
    spec = SeqexSpec(enc(entity), enc(uid_name), enc(label_name) if label is not None else None,
                     enc(offset_name) if offset is not None else None, enc(weight_name) if weight is not None else None,
                     enc(bag + "_indices") if bag else None, enc(bag + "_values") if bag else None)
    need = C.c_int64()
    args = [C.byref(spec), C.c_int64(E), _np_ptr(ent_rows), _np_ptr(ei), _np_ptr(idc), _np_ptr(idp), _np_ptr(row_len),
            _np_ptr(gcol), _np_ptr(val), _np_ptr(uid), _np_ptr(label), C.c_int32(1 if label_as_int else 0), _np_ptr(offset),
            _np_ptr(weight)]
    check(lib.gdmix_seqex_encode(*args, None, C.c_int64(0), C.byref(need)))
    out = np.empty(max(need.value, 1), np.uint8)
    written = C.c_int64()
    check(lib.gdmix_seqex_encode(*args, _np_ptr(out), C.c_int64(out.size), C.byref(written)))
    return out[:written.value]


def seqex_spec(entity, uid, label, offset, weight, bag_indices, bag_values):
    enc = lambda x: None if x is None else x.encode("utf-8")
    return SeqexSpec(enc(entity), enc(uid), enc(label), enc(offset), enc(weight), enc(bag_indices), enc(bag_values))


def seqex_count(buf, spec):
    """gdmix_seqex_count over a uint8 numpy view of one uncompressed file -> SeqexSizes."""
    sz = SeqexSizes()
    check(lib.gdmix_seqex_count(_np_ptr(buf), C.c_int64(buf.size), C.byref(spec), C.byref(sz)))
    return sz


def seqex_fill_into(buf, spec, out, e0, r0, q0, id_chars, id_ptr):
    """gdmix_seqex_fill of one file straight into the partition-wide arrays `out` at entity e0 / row r0 / non-zero q0;
    the file's entity-id strings go to its own id_chars / id_ptr (small)."""
    at = lambda a, i: None if a is None else C.c_void_p(a.ctypes.data + i * a.itemsize)
    rng = (C.c_int64 * 2)()
    check(lib.gdmix_seqex_fill(_np_ptr(buf), C.c_int64(buf.size), C.byref(spec), at(out["ent_rows"], e0),
                               at(out["row_len"], r0), at(out["gcol"], q0), at(out["val"], q0), at(out["uid"], r0),
                               at(out["label"], r0), at(out["offset"], r0), at(out["weight"], r0),
                               _np_ptr(id_chars), _np_ptr(id_ptr), rng))
    return int(rng[0]), int(rng[1])      # smallest / largest feature index written


def seqex_fill_local_into(buf, spec, out, e0, r0, q0, id_chars, id_ptr):
    """gdmix_seqex_fill_local of one file into the partition-wide arrays: as seqex_fill_into, with out["local16"],
    out["d_e"], out["uniq_scratch"] instead of out["gcol"]."""
    at = lambda a, i: None if a is None else C.c_void_p(a.ctypes.data + i * a.itemsize)
    rng = (C.c_int64 * 2)()
    check(lib.gdmix_seqex_fill_local(_np_ptr(buf), C.c_int64(buf.size), C.byref(spec), at(out["ent_rows"], e0),
                                     at(out["row_len"], r0), at(out["local16"], q0), at(out["d_e"], e0),
                                     at(out["uniq_scratch"], q0), at(out["val"], q0), at(out["uid"], r0),
                                     at(out["label"], r0), at(out["offset"], r0), at(out["weight"], r0),
                                     _np_ptr(id_chars), _np_ptr(id_ptr), rng))
    return int(rng[0]), int(rng[1])


def local_index_gather(ent_rowptr, rowptr, d_e, scratch):
    """Second call of gdmix_local_index_host: the distinct indices parked in scratch -> (uniq_ptr, uniq_global)."""
    E = ent_rowptr.shape[0] - 1
    d_e = np.ascontiguousarray(d_e, dtype=np.int64)
    uniq_ptr = np.zeros(E + 1, np.int64)
    np.cumsum(d_e[:E], out=uniq_ptr[1:])
    uniq_global = np.empty(max(int(uniq_ptr[-1]), 1), np.int64)
    check(lib.gdmix_local_index_host(_np_ptr(ent_rowptr), _np_ptr(rowptr), None, C.c_int64(E), None, _np_ptr(d_e),
                                     _np_ptr(scratch), _np_ptr(uniq_ptr), _np_ptr(uniq_global)))
    return uniq_ptr, uniq_global[:int(uniq_ptr[-1])]


def parse_per_record(file_image, uid, label, offset, weight, bag_indices, bag_values):
    """One uncompressed TFRecord file image of tf.train.Example rows -> dict of flat numpy arrays
    (gdmix_example_count + gdmix_example_fill; host code of the library)."""
    enc = lambda x: None if x is None else x.encode("utf-8")
    spec = SeqexSpec(None, enc(uid), enc(label), enc(offset), enc(weight), enc(bag_indices), enc(bag_values))
    buf = np.frombuffer(file_image, dtype=np.uint8)
    sz = SeqexSizes()
    check(lib.gdmix_example_count(_np_ptr(buf), C.c_int64(buf.size), C.byref(spec), C.byref(sz)))
    N, Z = sz.n_rows, sz.nnz
    out = {"row_len": np.zeros(N, np.int64), "col": np.zeros(max(Z, 1), np.int32)[:Z], "val": np.zeros(max(Z, 1), np.float32)[:Z],
           "uid": np.zeros(N, np.int64), "label": np.zeros(N, np.float32), "offset": np.zeros(N, np.float32),
           "weight": np.zeros(N, np.float32)}
    colbuf, valbuf = np.zeros(max(Z, 1), np.int32), np.zeros(max(Z, 1), np.float32)
    rl = np.zeros(max(N, 1), np.int64); u = np.zeros(max(N, 1), np.int64)
    lab, off, w = (np.zeros(max(N, 1), np.float32) for _ in range(3))
    check(lib.gdmix_example_fill(_np_ptr(buf), C.c_int64(buf.size), C.byref(spec), _np_ptr(rl), _np_ptr(colbuf),
                                 _np_ptr(valbuf), _np_ptr(u), _np_ptr(lab), _np_ptr(off), _np_ptr(w)))
    out = {"row_len": rl[:N], "col": colbuf[:Z], "val": valbuf[:Z], "uid": u[:N], "label": lab[:N], "offset": off[:N],
           "weight": w[:N], "saw_weight": bool(sz.saw_weight)}
    return out


def avro_score_blocks(uid, score, label, weight, per_coordinate, sync, records_per_block=1024):
    """-> the blocks of an Avro container holding these score records (gdmix_avro_score_blocks), as a bytes-like
    memoryview."""
    f32 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float32)
    uid = np.ascontiguousarray(uid, dtype=np.int64)
    score, label, weight, per_coordinate = f32(score), f32(label), f32(weight), f32(per_coordinate)
    n = int(uid.shape[0])
    blocks = (n + records_per_block - 1) // records_per_block
    out = np.empty(max(n * 31 + blocks * 36, 1), np.uint8)
    written = C.c_int64()
    sync_arr = np.frombuffer(bytes(sync), dtype=np.uint8)
    assert sync_arr.size == 16
    check(lib.gdmix_avro_score_blocks(_np_ptr(uid), _np_ptr(score), _np_ptr(label), _np_ptr(weight),
                                      _np_ptr(per_coordinate), C.c_int64(n), C.c_int32(records_per_block),
                                      _np_ptr(sync_arr), _np_ptr(out), C.c_int64(out.size), C.byref(written)))
    return memoryview(out)[:written.value]      # no copy: the caller writes it to the file


def _string_table(strings):
    """list of str -> (uint8 array of the concatenated utf-8 bytes, int64 offsets [n + 1])"""
    enc = [x.encode("utf-8") for x in strings]
    ptr = np.zeros(len(enc) + 1, np.int64)
    if enc:
        np.cumsum([len(b) for b in enc], out=ptr[1:])
    chars = np.frombuffer(b"".join(enc), dtype=np.uint8) if enc and ptr[-1] else np.zeros(1, np.uint8)
    return np.ascontiguousarray(chars), ptr


_FEATURE_TABLES = {}


def _feature_tables(feature_names, feature_terms):
    """String tables of the feature file's names / terms, kept for the last pair of lists seen (same objects)."""
    key = (id(feature_names), id(feature_terms), len(feature_names))
    hit = _FEATURE_TABLES.get(key)
    if hit is not None and hit[0] is feature_names and hit[1] is feature_terms:
        return hit[2]
    nc, npt = _string_table(list(feature_names))
    tc, tpt = _string_table(list(feature_terms))
    _FEATURE_TABLES.clear()
    _FEATURE_TABLES[key] = (feature_names, feature_terms, (nc, npt, tc, tpt))
    return nc, npt, tc, tpt


def avro_model_blocks(model_ids, coef, var, coef_ptr, feat_idx, has_intercept, threshold, feature_names, feature_terms,
                      model_class, intercept_name, sync, records_per_block=1024, id_table=None):
    """-> bytes: the blocks of an Avro container holding these BayesianLinearModelAvro records
    (gdmix_avro_model_blocks; the layout of its arguments is documented in include/gdmix_b200.h).
    id_table = (uint8 utf-8 characters, int64 offsets [n + 1]) of the model ids when the caller has them as arrays
    already (the reader's own table): no per-id Python work."""
    if id_table is not None and id_table[1].shape[0] == len(model_ids) + 1:
        idc, idp = np.ascontiguousarray(id_table[0], dtype=np.uint8), np.ascontiguousarray(id_table[1], dtype=np.int64)
        if idc.size == 0:
            idc = np.zeros(1, np.uint8)
    else:
        idc, idp = _string_table([str(m) for m in model_ids])
    nc, npt, tc, tpt = _feature_tables(feature_names, feature_terms)
    coef = np.ascontiguousarray(coef, dtype=np.float64)
    var = None if var is None else np.ascontiguousarray(var, dtype=np.float64)
    coef_ptr = np.ascontiguousarray(coef_ptr, dtype=np.int64)
    feat_idx = np.ascontiguousarray(feat_idx, dtype=np.int64)
    hi = 1 if has_intercept else 0
    n_coef = int(coef_ptr[-1]) if coef_ptr.size else 0
    # the encoder indexes var and feat_idx with coefficient offsets and checks no lengths itself
    if coef_ptr.shape[0] != len(model_ids) + 1 or coef.shape[0] < n_coef:
        raise ValueError("coef / coef_ptr do not describe len(model_ids) models")
    if var is not None and var.shape[0] != coef.shape[0]:
        raise ValueError(f"{var.shape[0]} variances for {coef.shape[0]} coefficients")
    if len(feature_names) and feat_idx.shape[0] != n_coef - hi * len(model_ids):
        raise ValueError(f"{feat_idx.shape[0]} feature indices for {n_coef - hi * len(model_ids)} feature coefficients")
    if feat_idx.size == 0:
        feat_idx = np.zeros(1, np.int64)
    t = ModelTable(len(model_ids), _np_ptr(idc), _np_ptr(idp), model_class.encode("utf-8"), _np_ptr(coef), _np_ptr(var),
                   _np_ptr(coef_ptr), _np_ptr(feat_idx), 1 if has_intercept else 0, 0, float(threshold),
                   intercept_name.encode("utf-8"), _np_ptr(nc), _np_ptr(npt), _np_ptr(tc), _np_ptr(tpt), len(feature_names))
    sync_arr = np.frombuffer(bytes(sync), dtype=np.uint8)
    assert sync_arr.size == 16
    ptr, written = C.c_void_p(), C.c_int64()
    check(lib.gdmix_avro_model_blocks_alloc(C.byref(t), C.c_int32(records_per_block), _np_ptr(sync_arr), C.byref(ptr),
                                            C.byref(written)))
    # no copy: a view of the library's buffer (freed when the view's owner is collected); the caller writes it out
    return memoryview(np.asarray(_LibBuffer(ptr.value, written.value)))


class FeatureMap:
    """(name, term) -> feature-file row, held by the library for gdmix_avro_model_decode."""

    def __init__(self, feature_names, feature_terms, intercept_name):
        self._keep = (_string_table(list(feature_names)), _string_table(list(feature_terms)))
        (nc, npt), (tc, tpt) = self._keep
        self.h = lib.gdmix_feature_map_create(_np_ptr(nc), _np_ptr(npt), _np_ptr(tc), _np_ptr(tpt),
                                              C.c_int64(len(feature_names)), intercept_name.encode("utf-8"))
        if not self.h:
            raise GdmixError(GDMIX_ERR_INVALID, lib.gdmix_last_error().decode())

    def decode_models(self, block, n_records):
        """One uncompressed container block of BayesianLinearModelAvro records -> flat arrays (dict)."""
        buf = np.frombuffer(block, dtype=np.uint8)
        nm, ib = C.c_int64(), C.c_int64()
        h = C.c_void_p(self.h)
        check(lib.gdmix_avro_model_decode(h, _np_ptr(buf), C.c_int64(buf.size), C.c_int64(n_records), C.byref(nm),
                                          C.byref(ib), None, None, None, None, None, None, None))
        n = int(n_records)
        out = {"id_chars": np.zeros(max(ib.value, 1), np.uint8), "id_ptr": np.zeros(n + 1, np.int64),
               "mean_ptr": np.zeros(n + 1, np.int64), "mean_feat": np.zeros(max(nm.value, 1), np.int64),
               "mean_val": np.zeros(max(nm.value, 1), np.float64), "var_val": np.zeros(max(nm.value, 1), np.float64),
               "has_var": np.zeros(max(n, 1), np.uint8)}
        check(lib.gdmix_avro_model_decode(h, _np_ptr(buf), C.c_int64(buf.size), C.c_int64(n_records), C.byref(nm),
                                          C.byref(ib), _np_ptr(out["id_chars"]), _np_ptr(out["id_ptr"]),
                                          _np_ptr(out["mean_ptr"]), _np_ptr(out["mean_feat"]), _np_ptr(out["mean_val"]),
                                          _np_ptr(out["var_val"]), _np_ptr(out["has_var"])))
        return out

    def close(self):
        if getattr(self, "h", None):
            lib.gdmix_feature_map_destroy(C.c_void_p(self.h))
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
