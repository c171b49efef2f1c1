"""TFRecords -> the flat CSR batches the C ABI takes (the reference's producer side).

Random effect: one ``SequenceExample`` per entity (context = entity id + per-sample dense columns as
variable-length lists; feature_lists = ``<bag>_indices`` / ``<bag>_values`` with one Feature per sample) as read by
``per_entity_grouped_input_fn`` (gdmix-trainer/src/gdmix/io/input_data_pipeline.py:223-332) and sliced per entity
by ``prepare_jobs`` (models/custom/scipy/job_consumers.py:161-296):

  * entity id -> ``str`` (bytes decoded as utf-8, integers through ``str()``)            :235-239
  * global feature ids -> entity-local columns = rank among the entity's sorted unique ids  :243
  * intercept-only model (no feature bag): X is an n x 1 zero column                        :213-218
  * warm start: prior intercept + prior coefficients of the features present now            :262-288

Fixed effect: one ``Example`` per row (``per_record_input_fn``, :129-220).

Files: a directory is globbed for ``*.tfrecord``, then ``*.tfrecord.deflate``, then ``*.tfrecord.gz``
(:88-126); the sorted list is sharded ``files[shard_index::num_shards]`` (util/distribution_utils.py:36-47).
"""
import glob
import os

import numpy as np

from . import constants
from ._capi import HostBatch
from .io import tfrecord
from .io.dataset_metadata import DatasetMetadata

INDICES_SUFFIX = "_indices"
VALUES_SUFFIX = "_values"


def list_tfrecord_files(input_path, num_shards=1, shard_index=0):
    if isinstance(input_path, (list, tuple)):
        files = sorted(input_path)
    elif os.path.isdir(input_path):
        files = []
        for suffix in ("", ".deflate", ".gz"):
            files = sorted(glob.glob(os.path.join(input_path, constants.TFRECORD_GLOB_PATTERN + suffix)))
            if files:
                break
        if not files:
            files = sorted(f for f in glob.glob(os.path.join(input_path, "*"))
                           if os.path.isfile(f) and not os.path.basename(f).startswith((".", "_")))
    else:
        files = sorted(glob.glob(input_path)) or ([input_path] if os.path.exists(input_path) else [])
    return files[shard_index::num_shards]


def is_empty_directory(path):
    return not os.path.isdir(path) or not any(not n.startswith((".", "_")) for n in os.listdir(path))


class EntityGroupedData:
    """All entities of one partition, flattened.  Columns are GLOBAL feature ids here."""

    def __init__(self):
        self.entity_ids = []
        self.ent_rowptr = np.zeros(1, np.int64)
        self.rowptr = np.zeros(1, np.int64)
        self._gcol = np.zeros(0, np.int64)
        # the fused reader's output instead of gcol: entity-local indices (uint16), distinct features per entity
        self.local16 = self.uniq_ptr = self.uniq_global = None
        self.val = np.zeros(0, np.float32)
        self.label = None
        self.weight = None
        self.offset = None
        self.uid = None
        self.has_weight_column = False
        self.num_features = 1

    @property
    def gcol(self):
        """GLOBAL feature id of every non-zero; rebuilt from the local indices when the reader never wrote it."""
        if self._gcol is None:
            nnz_e = self.rowptr[self.ent_rowptr[1:]] - self.rowptr[self.ent_rowptr[:-1]]
            base = np.repeat(self.uniq_ptr[:-1], nnz_e)
            self._gcol = self.uniq_global[base + self.local16.astype(np.int64)]
        return self._gcol

    @gcol.setter
    def gcol(self, value):
        self._gcol = value

    @property
    def n_entities(self):
        return len(self.entity_ids)

    @property
    def n_rows(self):
        return int(self.ent_rowptr[-1])


def _entity_id_to_str(kind, values):
    v = values[0]
    if kind == "bytes":
        return v.decode("utf-8")
    return str(int(v)) if kind == "int64" else str(v)


def _read_entity_grouped_native(files, entity_name, feature_bag, label_column, offset_column, weight_column,
                                uid_column, num_features, input_path, fused=None):
    """The library's own SequenceExample reader (csrc/seqex_parser.h through gdmix_seqex_count / _fill): ~50x the
    pure-Python protobuf walk below, same arrays.  Returns None for the one case it leaves to that walk (float
    entity ids); malformed files raise ValueError like the Python reader does."""
    from . import _capi as capi
    if fused is None:
        fused = os.environ.get("GDMIX_INGEST_FUSED", "1") != "0"
    # Two passes over the partition's files, each file on its own thread (the parser is native code: the calls release
    # the GIL): count, then fill straight into the partition-wide arrays -- no per-file arrays, no concatenation.
    # Uncompressed files are memory-mapped; compressed ones are inflated first.
    import mmap
    from concurrent.futures import ThreadPoolExecutor
    spec = capi.seqex_spec(entity_name, uid_column, label_column, offset_column, weight_column,
                           feature_bag + INDICES_SUFFIX, feature_bag + VALUES_SUFFIX)
    keep = []

    def image(fn):
        if tfrecord.compression_of(fn) or os.path.getsize(fn) == 0:
            return np.frombuffer(tfrecord._read_all(fn), dtype=np.uint8)
        with open(fn, "rb") as f:
            mm = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
        keep.append(mm)
        return np.frombuffer(mm, dtype=np.uint8)

    def count(fn):
        buf = image(fn)
        try:
            return buf, capi.seqex_count(buf, spec)
        except capi.GdmixError as ex:
            if "Python reader" in str(ex):
                return buf, None
            raise ValueError(f"{fn}: {ex}") from None

    workers = max(1, min(len(files), os.cpu_count() or 1))
    with ThreadPoolExecutor(max_workers=workers) as pool:
        counted = list(pool.map(count, files))
        if any(sz is None for _, sz in counted):
            return None
        E = sum(sz.n_entities for _, sz in counted)
        N = sum(sz.n_rows for _, sz in counted)
        Z = sum(sz.nnz for _, sz in counted)
        # what the device calls upload afterwards (values, labels, offsets, weights) is parsed straight into
        # page-locked memory of the library's pool; the global column ids and uids stay on the host
        # Fused (default): an entity's indices are ranked among its distinct ones while the reader still has them in
        # cache (np.unique per entity, job_consumers.py:243) -- the int64 global columns are never written or re-read.
        pin = capi.pinned_empty
        out = {"ent_rows": np.empty(E, np.int64), "row_len": np.empty(N, np.int64),
               "val": pin(Z, np.float32), "uid": np.empty(N, np.int64), "label": pin(N, np.float32),
               "offset": pin(N, np.float32), "weight": pin(N, np.float32)}
        if fused:
            out.update({"local16": pin(Z, np.uint16), "d_e": np.zeros(max(E, 1), np.int64),
                        "uniq_scratch": np.empty(max(Z, 1), np.int64)})
        else:
            out["gcol"] = np.empty(Z, np.int64)
        starts, e0, r0, q0 = [], 0, 0, 0
        for _, sz in counted:
            starts.append((e0, r0, q0))
            e0 += sz.n_entities; r0 += sz.n_rows; q0 += sz.nnz

        def fill(k):
            buf, sz = counted[k]
            idc, idp = np.zeros(max(sz.id_bytes, 1), np.uint8), np.zeros(sz.n_entities + 1, np.int64)
            try:
                rng = (capi.seqex_fill_local_into if fused else capi.seqex_fill_into)(buf, spec, out, *starts[k], idc, idp)
            except capi.GdmixError as ex:
                if fused and ex.code == capi.GDMIX_ERR_TOO_LARGE:
                    return None, None, None
                raise ValueError(f"{files[k]}: {ex}") from None
            raw = idc.tobytes()
            if raw.isascii():      # byte offsets are character offsets: one decode, then slices
                txt, pl = raw.decode("ascii"), idp.tolist()
                ids_k = [txt[pl[e]:pl[e + 1]] for e in range(sz.n_entities)]
            else:
                ids_k = [raw[idp[e]:idp[e + 1]].decode("utf-8") for e in range(sz.n_entities)]
            return ids_k, rng, (idc[:sz.id_bytes], idp)

        filled = list(pool.map(fill, range(len(files))))
        if any(f[0] is None for f in filled):
            # an entity with more than 65535 distinct features: the unfused reader + gdmix_local_index_host
            return _read_entity_grouped_native(files, entity_name, feature_bag, label_column, offset_column,
                                               weight_column, uid_column, num_features, input_path, fused=False)
        ids = [f[0] for f in filled]
        ranges = [f[1] for f, (_, sz) in zip(filled, counted) if sz.nnz]
        # the ids once more as one (characters, offsets) table: what the model writer takes without touching a str
        id_chars = np.concatenate([f[2][0] for f in filled]) if filled else np.zeros(0, np.uint8)
        id_ptr = np.zeros(E + 1, np.int64)
        at, base = 0, 0
        for f, (_, sz) in zip(filled, counted):
            id_ptr[at:at + sz.n_entities + 1] = f[2][1][:sz.n_entities + 1] + base
            at += sz.n_entities
            base += sz.id_bytes
    all_labelled = all(sz.all_labelled for _, sz in counted)
    index_lo = min((r[0] for r in ranges), default=0)      # tracked by the parser's filling pass
    index_hi = max((r[1] for r in ranges), default=0)
    saw_weight = any(sz.saw_weight for _, sz in counted)
    del counted
    d = EntityGroupedData()
    d.num_features = int(num_features)
    d.entity_ids = [i for part in ids for i in part]
    d.entity_id_table = (id_chars, id_ptr)
    d.has_weight_column = saw_weight
    d.ent_rowptr = np.zeros(E + 1, np.int64)
    np.cumsum(out["ent_rows"], out=d.ent_rowptr[1:])
    d.rowptr = capi.pinned_empty(N + 1, np.int64)
    d.rowptr[0] = 0
    np.cumsum(out["row_len"], out=d.rowptr[1:])
    d.val = out["val"]
    if fused:
        d._gcol = None
        d.local16 = out["local16"]
        d.uniq_ptr, d.uniq_global = capi.local_index_gather(d.ent_rowptr, d.rowptr, out["d_e"][:E], out["uniq_scratch"])
    else:
        d.gcol = out["gcol"]
    d.uid, d.offset, d.weight = out["uid"], out["offset"], out["weight"]
    d.label = out["label"] if (files and all_labelled and label_column is not None) else None
    for mm in keep:
        try:
            mm.close()
        except BufferError:      # a numpy view is still alive somewhere: the mapping goes with it
            pass
    if Z and (index_lo < 0 or index_hi >= d.num_features):
        raise ValueError(f"feature index outside [0, {d.num_features}) in {input_path}")
    return d


def read_entity_grouped(input_path, metadata, entity_name, feature_bag, label_column, offset_column, weight_column,
                        uid_column, num_features, num_shards=1, shard_index=0):
    """-> EntityGroupedData.  `metadata` is a DatasetMetadata (used to check that the entity column exists, as the
    reference does) ."""
    if isinstance(metadata, str):
        metadata = DatasetMetadata(metadata)
    if entity_name not in metadata.get_feature_names():
        raise ValueError(f"entity name {entity_name} is not found among the features")
    files = list_tfrecord_files(input_path, num_shards, shard_index)
    if feature_bag is not None:
        d = _read_entity_grouped_native(files, entity_name, feature_bag, label_column, offset_column, weight_column,
                                        uid_column, num_features, input_path)
        if d is not None:
            return d
    ids, n_per_entity = [], []
    row_len, gcols, vals, labels, weights, offsets, uids = [], [], [], [], [], [], []
    saw_weight = False
    for fn in files:
        for payload in tfrecord.read_records(fn):
            ctx, lists = tfrecord.parse_sequence_example(payload)
            if entity_name not in ctx or ctx[entity_name][0] is None:
                raise ValueError(f"record without entity column {entity_name!r} in {fn}")
            ids.append(_entity_id_to_str(*ctx[entity_name]))
            uid = np.asarray(ctx[uid_column][1], dtype=np.int64)
            n = uid.shape[0]
            n_per_entity.append(n)
            uids.append(uid)
            if feature_bag is None:
                # intercept-only: one explicit zero in column 0 per sample
                row_len.append(np.ones(n, np.int64))
                gcols.append(np.zeros(n, np.int64))
                vals.append(np.zeros(n, np.float32))
            else:
                fi = lists.get(feature_bag + INDICES_SUFFIX, [])
                fv = lists.get(feature_bag + VALUES_SUFFIX, [])
                if len(fi) != n or len(fv) != n:
                    raise ValueError(f"entity {ids[-1]}: {len(fi)} index lists / {len(fv)} value lists for {n} samples")
                for (_, ci), (_, vi) in zip(fi, fv):
                    ci = np.zeros(0, np.int64) if ci is None else np.asarray(ci, np.int64)
                    vi = np.zeros(0, np.float32) if vi is None else np.asarray(vi, np.float32)
                    if ci.shape[0] != vi.shape[0]:
                        raise ValueError(f"entity {ids[-1]}: indices / values length mismatch")
                    gcols.append(ci)
                    vals.append(vi)
                row_len.append(np.array([len(c[1]) if c[1] is not None else 0 for c in fi], np.int64))
            if label_column is not None and label_column in ctx:
                labels.append(np.asarray(ctx[label_column][1], dtype=np.float32))
            offsets.append(np.asarray(ctx[offset_column][1], dtype=np.float32) if offset_column in ctx
                           else np.zeros(n, np.float32))
            if weight_column is not None and weight_column in ctx:
                saw_weight = True
                weights.append(np.asarray(ctx[weight_column][1], dtype=np.float32))
            else:
                weights.append(np.ones(n, np.float32))
    d = EntityGroupedData()
    d.entity_ids = ids
    d.num_features = int(num_features)
    d.has_weight_column = saw_weight
    d.ent_rowptr = np.concatenate([[0], np.cumsum(n_per_entity)]).astype(np.int64)
    rl = np.concatenate(row_len) if row_len else np.zeros(0, np.int64)
    d.rowptr = np.concatenate([[0], np.cumsum(rl)]).astype(np.int64)
    d.gcol = np.concatenate(gcols).astype(np.int64) if gcols else np.zeros(0, np.int64)
    d.val = np.concatenate(vals).astype(np.float32) if vals else np.zeros(0, np.float32)
    cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt)
    d.uid = cat(uids, np.int64)
    d.offset = cat(offsets, np.float32)
    d.weight = cat(weights, np.float32)
    d.label = cat(labels, np.float32) if labels and len(labels) == len(ids) else None
    if d.gcol.size and (d.gcol.min() < 0 or d.gcol.max() >= d.num_features):
        raise ValueError(f"feature index outside [0, {d.num_features}) in {input_path}")
    return d


def to_local_batch(data, has_intercept=True):
    """np.unique per entity (job_consumers.py:243) for the whole partition, by the library (all host threads).
    -> (HostBatch with entity-local columns, uniq_ptr int64[E+1], uniq_global int64[sum d_e])"""
    from . import _capi as capi
    hi = 1 if has_intercept else 0
    label = data.label if data.label is not None else np.zeros(data.n_rows, np.float32)
    if data.local16 is not None:      # the fused reader ranked the indices already
        uniq_ptr, uniq_global = data.uniq_ptr, data.uniq_global
        theta_ptr = uniq_ptr + hi * np.arange(uniq_ptr.shape[0], dtype=np.int64)
        hb = HostBatch(data.ent_rowptr, data.rowptr, None, data.val, label, data.weight, data.offset, theta_ptr,
                       has_intercept, col_narrow=data.local16)
        return hb, uniq_ptr, uniq_global
    local, d_e, uniq_ptr, uniq_global = capi.local_index_host(data.ent_rowptr, data.rowptr, data.gcol)
    theta_ptr = np.concatenate([[0], np.cumsum(d_e + hi)]).astype(np.int64)
    hb = HostBatch(data.ent_rowptr, data.rowptr, local, data.val, label, data.weight, data.offset, theta_ptr,
                   has_intercept)
    return hb, uniq_ptr, uniq_global


def warm_start_theta(hb, uniq_ptr, uniq_global, entity_ids, model_weights, has_intercept=True):
    """theta0 for every entity of the batch plus a has_model flag (job_consumers.py:262-288: the prior's
    intercept and the prior coefficients of features that occur in the current data; everything else 0).
    One sorted merge over (entity, feature) keys for the whole batch instead of a search per entity."""
    hi = 1 if has_intercept else 0
    E = len(entity_ids)
    flat = bool(model_weights) and hasattr(model_weights, "theta_ptr") and hasattr(model_weights, "index")
    if flat and getattr(model_weights, "source_ids", None) is entity_ids and model_weights.uniq_global is uniq_global and \
            model_weights.theta.shape[0] == hb.n_coef:
        # the very arrays the models were trained from (scoring the partition just trained, parsed once): the
        # coefficients line up one to one and are handed on as they are (read-only downstream)
        return model_weights.theta, np.ones(E, np.uint8)
    theta0 = np.zeros(hb.n_coef, np.float64)
    has_model = np.zeros(E, np.uint8)
    if not model_weights:
        return theta0, has_model
    if flat:
        return _warm_start_from_flat(hb, uniq_ptr, uniq_global, entity_ids, model_weights, hi, theta0, has_model)
    ents, idxs, coefs = [], [], []
    for e, eid in enumerate(entity_ids):
        prior = model_weights.get(eid)
        if prior is None:
            continue
        has_model[e] = 1
        ptheta = np.asarray(prior.theta, np.float64)
        if has_intercept:
            theta0[hb.theta_ptr[e]] = ptheta[0]
        pidx = np.asarray(prior.unique_global_indices, np.int64)
        if pidx.size:
            ents.append(np.full(pidx.size, e, np.int64)); idxs.append(pidx); coefs.append(ptheta[hi:hi + pidx.size])
    if not idxs or uniq_global.size == 0:
        return theta0, has_model
    p_ent, p_idx, p_coef = np.concatenate(ents), np.concatenate(idxs), np.concatenate(coefs)
    width = np.int64(max(int(p_idx.max()), int(uniq_global.max())) + 1)
    p_key = p_ent * width + p_idx
    order = np.argsort(p_key, kind="stable")            # a repeated prior feature keeps its first occurrence
    p_key, p_coef = p_key[order], p_coef[order]
    d_e = np.diff(uniq_ptr)
    c_ent = np.repeat(np.arange(E, dtype=np.int64), d_e)
    c_key = c_ent * width + uniq_global
    pos = np.minimum(np.searchsorted(p_key, c_key), p_key.size - 1)
    hit = p_key[pos] == c_key
    k = np.flatnonzero(hit)
    theta0[hb.theta_ptr[c_ent[k]] + hi + (k - uniq_ptr[c_ent[k]])] = p_coef[pos[k]]
    return theta0, has_model


def _warm_start_from_flat(hb, uniq_ptr, uniq_global, entity_ids, fm, hi, theta0, has_model):
    """warm_start_theta when the prior models are a random_effect.FlatModels (the models a train() just produced):
    the same merge over (entity, feature) keys, with the prior side taken from the flat arrays instead of one
    TrainingResult per entity."""
    E = len(entity_ids)
    index = fm.index
    src = np.fromiter((index.get(eid, -1) for eid in entity_ids), dtype=np.int64, count=E)   # prior model of entity e
    hit_e = np.flatnonzero(src >= 0)
    has_model[hit_e] = 1
    if hit_e.size == 0:
        return theta0, has_model
    ptp, pup = np.asarray(fm.theta_ptr, np.int64), np.asarray(fm.uniq_ptr, np.int64)
    if hi:
        theta0[hb.theta_ptr[hit_e]] = fm.theta[ptp[src[hit_e]]]
    if E == len(fm.entity_ids) and np.array_equal(src, np.arange(E)) and np.array_equal(pup, uniq_ptr) and \
            np.array_equal(fm.uniq_global, uniq_global):
        # the very partition the models were trained on (scoring after training): coefficients line up one to one
        theta0[:] = fm.theta
        return theta0, has_model
    d_prior = np.diff(pup)[src[hit_e]]
    p_ent = np.repeat(hit_e, d_prior)
    start = np.repeat(pup[src[hit_e]], d_prior)
    within = np.arange(p_ent.size, dtype=np.int64) - np.repeat(np.cumsum(d_prior) - d_prior, d_prior)
    p_idx = np.asarray(fm.uniq_global, np.int64)[start + within]
    p_coef = fm.theta[np.repeat(ptp[src[hit_e]] + hi, d_prior) + within]
    if p_idx.size == 0 or uniq_global.size == 0:
        return theta0, has_model
    width = np.int64(max(int(p_idx.max()), int(uniq_global.max())) + 1)
    p_key = p_ent * width + p_idx                        # ascending already: entities ascending, features ascending inside
    c_ent = np.repeat(np.arange(E, dtype=np.int64), np.diff(uniq_ptr))
    c_key = c_ent * width + uniq_global
    pos = np.minimum(np.searchsorted(p_key, c_key), p_key.size - 1)
    k = np.flatnonzero(p_key[pos] == c_key)
    theta0[hb.theta_ptr[c_ent[k]] + hi + (k - uniq_ptr[c_ent[k]])] = p_coef[pos[k]]
    return theta0, has_model


def _warm_start_theta_per_entity(hb, uniq_ptr, uniq_global, entity_ids, model_weights, has_intercept=True):
    """The same, one entity at a time (the shape of the reference's loop); kept as the test's cross-check."""
    hi = 1 if has_intercept else 0
    theta0 = np.zeros(hb.n_coef, np.float64)
    has_model = np.zeros(len(entity_ids), np.uint8)
    if not model_weights:
        return theta0, has_model
    for e, eid in enumerate(entity_ids):
        prior = model_weights.get(eid)
        if prior is None:
            continue
        has_model[e] = 1
        t0 = hb.theta_ptr[e]
        ptheta = np.asarray(prior.theta, np.float64)
        if has_intercept:
            theta0[t0] = ptheta[0]
        pidx = np.asarray(prior.unique_global_indices, np.int64)
        pcoef = ptheta[hi:]
        cur = uniq_global[uniq_ptr[e]:uniq_ptr[e + 1]]
        if pidx.size and cur.size:
            order = np.argsort(pidx, kind="stable")
            ps, pc = pidx[order], pcoef[order]
            pos = np.searchsorted(ps, cur)
            pos_c = np.minimum(pos, ps.size - 1)
            hit = ps[pos_c] == cur
            theta0[t0 + hi + np.flatnonzero(hit)] = pc[pos_c[hit]]
    return theta0, has_model


class RecordData:
    """Rows of a per-record (fixed-effect) dataset as CSR with GLOBAL feature ids."""

    def __init__(self):
        self.rowptr = np.zeros(1, np.int64)
        self.col = np.zeros(0, np.int32)
        self.val = np.zeros(0, np.float32)
        self.label = None
        self.weight = None
        self.offset = None
        self.uid = None
        self.has_weight_column = False

    @property
    def n_rows(self):
        return len(self.rowptr) - 1


USE_NATIVE_READER = True   # tests flip it to compare the library's reader with the Python walk below


def _read_per_record_native(files, feature_bag, label_column, offset_column, weight_column, uid_column):
    """The library's tf.train.Example reader (gdmix_example_count / _fill), same arrays as the Python walk."""
    from . import _capi as capi
    parts = []
    for fn in files:
        try:
            parts.append(capi.parse_per_record(tfrecord._read_all(fn), uid_column, label_column, offset_column,
                                               weight_column,
                                               None if feature_bag is None else feature_bag + INDICES_SUFFIX,
                                               None if feature_bag is None else feature_bag + VALUES_SUFFIX))
        except capi.GdmixError as ex:
            raise ValueError(f"{fn}: {ex}") from None
    cat = lambda key, dt: (np.concatenate([p[key] for p in parts]).astype(dt) if parts else np.zeros(0, dt))
    d = RecordData()
    d.rowptr = np.concatenate([[0], np.cumsum(cat("row_len", np.int64))]).astype(np.int64)
    d.col, d.val = cat("col", np.int32), cat("val", np.float32)
    d.uid, d.label = cat("uid", np.int64), cat("label", np.float32)
    d.offset, d.weight = cat("offset", np.float32), cat("weight", np.float32)
    d.has_weight_column = any(p["saw_weight"] for p in parts)
    return d


def read_per_record(input_path_or_files, feature_bag, label_column, offset_column, weight_column, uid_column,
                    num_shards=1, shard_index=0):
    files = list_tfrecord_files(input_path_or_files, num_shards, shard_index)
    if USE_NATIVE_READER:
        return _read_per_record_native(files, feature_bag, label_column, offset_column, weight_column, uid_column)
    row_len, cols, vals, labels, weights, offsets, uids = [], [], [], [], [], [], []
    saw_weight = False
    for fn in files:
        for payload in tfrecord.read_records(fn):
            ex = tfrecord.parse_example(payload)
            if feature_bag is not None:
                ci = ex.get(feature_bag + INDICES_SUFFIX, (None, None))[1]
                vi = ex.get(feature_bag + VALUES_SUFFIX, (None, None))[1]
                ci = np.zeros(0, np.int64) if ci is None else np.asarray(ci, np.int64)
                vi = np.zeros(0, np.float32) if vi is None else np.asarray(vi, np.float32)
                cols.append(ci); vals.append(vi); row_len.append(ci.shape[0])
            else:
                row_len.append(0)
            one = lambda name, default: ex[name][1][0] if name is not None and name in ex and len(ex[name][1]) \
                else default
            uids.append(one(uid_column, 0))
            labels.append(one(label_column, np.nan))
            offsets.append(one(offset_column, 0.0))
            if weight_column is not None and weight_column in ex:
                saw_weight = True
            weights.append(one(weight_column, 1.0))
    d = RecordData()
    d.rowptr = np.concatenate([[0], np.cumsum(row_len)]).astype(np.int64)
    d.col = np.concatenate(cols).astype(np.int32) if cols else np.zeros(0, np.int32)
    d.val = np.concatenate(vals).astype(np.float32) if vals else np.zeros(0, np.float32)
    d.uid = np.asarray(uids, np.int64)
    d.label = np.asarray(labels, np.float32)
    d.offset = np.asarray(offsets, np.float32)
    d.weight = np.asarray(weights, np.float32)
    d.has_weight_column = saw_weight
    return d
