"""Device-side partitioner / evaluator: what gdmix-data's Spark jobs do before and after the hot path, without
leaving HBM (kernels in csrc/partition.cuh, C ABI in include/gdmix_b200.h).

  group_by_entity   DataPartitioner.boundAndGroupData's groupBy(entity) (DataPartitioner.scala:296-379)
  group_ids         DataPartitioner.getGroupId: active / passive bounds per entity (DataPartitioner.scala:335-379)
  join_offsets      OffsetUpdater.updateOffset: the previous coordinate's scores joined by uid (OffsetUpdater.scala:105-129)
  partition_and_write   groupPartitionAndSaveDataset: the active|passive/partitionId=k TFRecord layout and
                    partitionList.txt that RandomEffectDriver consumes (DataPartitioner.scala:95-120, 203-280)
  regroup_batch     rows in arrival order (+ global feature ids) -> the entity-local CSR batch gdmix_re_fit takes
                    (prepare_jobs' np.unique per entity, job_consumers.py:243, for the whole dataset at once)
  partition_ids     PartitionUtils.getPartitionIdUDF for integer entity ids (PartitionUtils.scala:31-37)
  auc               Evaluator.calculateMetric(..., "auc") (Evaluator.scala:29-45)

torch supplies device memory only; O(nnz) work is done by the library's kernels, O(#entities) glue (a cumsum over
segment lengths) by torch.
"""
import ctypes as C

import numpy as np

from . import _capi as capi
from ._capi import _stream_ptr, _tptr, check, lib


def _ws(n, device):
    import torch
    b = C.c_size_t()
    check(lib.gdmix_partition_workspace_size(C.c_int64(max(int(n), 1)), C.byref(b)))
    return torch.empty(b.value, dtype=torch.uint8, device=device)


def _key_bits(keys):
    mx = int(keys.max().item()) if keys.numel() else 0
    return max(1, mx.bit_length())


def group_by_entity(keys, key_bits=None, stream=None):
    """keys: int64 CUDA tensor of non-negative entity keys, one per row.
    -> (perm int32[n]: rows in grouped order, original order kept inside an entity; seg_ptr int64[G+1];
        seg_key int64[G]: the entities in ascending key order)"""
    import torch
    n = keys.numel()
    assert keys.dtype == torch.int64 and keys.is_cuda
    if key_bits is None:
        key_bits = _key_bits(keys)
    ws = _ws(n, keys.device)
    perm = torch.empty(n, dtype=torch.int32, device=keys.device)
    keys_sorted = torch.empty(n, dtype=torch.int64, device=keys.device)
    seg_ptr = torch.empty(n + 1, dtype=torch.int64, device=keys.device)
    seg_key = torch.empty(max(n, 1), dtype=torch.int64, device=keys.device)
    ng = torch.zeros(1, dtype=torch.int64, device=keys.device)
    check(lib.gdmix_group_by_key(_tptr(keys), C.c_int64(n), C.c_int32(key_bits), _tptr(keys_sorted), _tptr(perm),
                                 _tptr(seg_ptr), _tptr(seg_key), _tptr(ng), _tptr(ws), C.c_size_t(ws.numel()),
                                 _stream_ptr(stream)))
    g = int(ng.item())
    return perm, seg_ptr[:g + 1], seg_key[:g]


def sort_pairs(keys, key_bits=None, stream=None):
    """Stable ascending sort of non-negative int64 keys -> (keys_sorted, perm int32)."""
    import torch
    n = keys.numel()
    if key_bits is None:
        key_bits = _key_bits(keys)
    ws = _ws(n, keys.device)
    out = torch.empty_like(keys)
    perm = torch.empty(n, dtype=torch.int32, device=keys.device)
    check(lib.gdmix_sort_pairs_u64(_tptr(keys), C.c_int64(n), C.c_int32(key_bits), _tptr(out), _tptr(perm), _tptr(ws),
                                   C.c_size_t(ws.numel()), _stream_ptr(stream)))
    return out, perm


def gather_rows(rowptr, col, val, perm, stream=None):
    """CSR rows in the order perm gives -> (rowptr', col', val')."""
    import torch
    n = perm.numel()
    ws = _ws(n, perm.device)
    rp = torch.empty(n + 1, dtype=torch.int64, device=perm.device)
    co = torch.empty_like(col)
    va = torch.empty_like(val)
    check(lib.gdmix_csr_gather_rows(_tptr(rowptr), _tptr(col), _tptr(val), _tptr(perm), C.c_int64(n), _tptr(rp),
                                    _tptr(co), _tptr(va), _tptr(ws), C.c_size_t(ws.numel()), _stream_ptr(stream)))
    return rp, co, va


def gather_f32(x, perm, stream=None):
    import torch
    out = torch.empty(perm.numel(), dtype=torch.float32, device=perm.device)
    check(lib.gdmix_gather_f32(_tptr(x), _tptr(perm), C.c_int64(perm.numel()), _tptr(out), _stream_ptr(stream)))
    return out


def partition_ids(ids, num_partitions, stream=None):
    """abs(str(id).hashCode()) % num_partitions for an int64 CUDA tensor of entity ids -> int32 tensor."""
    import torch
    out = torch.empty(ids.numel(), dtype=torch.int32, device=ids.device)
    check(lib.gdmix_partition_ids_i64(_tptr(ids), C.c_int64(ids.numel()), C.c_int32(num_partitions), _tptr(out),
                                      _stream_ptr(stream)))
    return out


def auc(score, label, stream=None):
    """Area under the ROC curve of fp32 CUDA tensors (label > 0 positive), ties counted half."""
    import torch
    n = score.numel()
    ws = _ws(n, score.device)
    out = torch.zeros(3, dtype=torch.float64, device=score.device)
    check(lib.gdmix_auc(_tptr(score.contiguous()), _tptr(label.contiguous()), C.c_int64(n), _tptr(out), _tptr(ws),
                        C.c_size_t(ws.numel()), _stream_ptr(stream)))
    a, p, q = out.cpu().tolist()
    return a


def regroup_batch(entity, rowptr, gcol, val, label, offset=None, weight=None, has_intercept=True, num_features=None):
    """Rows in arrival order -> the entity-local CSR batch of a random-effect stage, all on the device.

    entity int64[n] (non-negative ids), rowptr int64[n+1], gcol int32[nnz] GLOBAL feature ids, val fp32[nnz],
    label / offset / weight fp32[n].  Returns a dict with the gdmix_re_batch arrays (ent_rowptr, rowptr, col =
    entity-LOCAL feature index, val, label, offset, weight, theta_ptr), the bounds (max_rows, max_nnz, max_coef),
    `perm` (row i of the batch is input row perm[i]), `entity_ids` and `uniq_ptr` / `uniq_global` (entity e's
    local feature j is global feature uniq_global[uniq_ptr[e] + j]) -- what prepare_jobs derives per entity
    (job_consumers.py:243-258), for the whole dataset in a handful of launches."""
    import torch
    dev = entity.device
    n = entity.numel()
    perm, ent_rowptr, entity_ids = group_by_entity(entity)
    E = entity_ids.numel()
    rp, gc, va = gather_rows(rowptr, gcol, val, perm)
    lab = gather_f32(label, perm)
    off = gather_f32(offset, perm) if offset is not None else None
    wt = gather_f32(weight, perm) if weight is not None else None
    nnz = gc.numel()
    hi = 1 if has_intercept else 0
    w32 = (int(num_features) + 31) // 32 if num_features else 0
    if num_features and num_features <= BITMAP_MAX_FEATURES and E * w32 <= (1 << 28) and not FORCE_PAIR_SORT:
        # small feature bags: per-entity presence bitmaps instead of sorting (entity, feature) pairs
        ent_rowptr = ent_rowptr.contiguous()
        bitmap = torch.empty(E * w32 + 1, dtype=torch.int32, device=dev)   # + the out-of-range flag
        wprefix = torch.empty(E * w32, dtype=torch.int32, device=dev)
        d_e = torch.empty(E, dtype=torch.int64, device=dev)
        check(lib.gdmix_local_index_mark(_tptr(ent_rowptr), C.c_int64(E), _tptr(rp), _tptr(gc), C.c_int64(n),
                                         C.c_int32(int(num_features)), _tptr(bitmap), _tptr(wprefix), _tptr(d_e),
                                         _stream_ptr(None)))
        uniq_ptr = torch.zeros(E + 1, dtype=torch.int64, device=dev)
        uniq_ptr[1:] = torch.cumsum(d_e, 0)
        uniq_global = torch.empty(int(uniq_ptr[-1].item()), dtype=torch.int64, device=dev)
        local = torch.empty(nnz, dtype=torch.int32, device=dev)
        check(lib.gdmix_local_index_apply(_tptr(ent_rowptr), C.c_int64(E), _tptr(rp), _tptr(gc), C.c_int64(n),
                                          C.c_int32(int(num_features)), _tptr(bitmap), _tptr(wprefix), _tptr(uniq_ptr),
                                          _tptr(local), _tptr(uniq_global), _stream_ptr(None)))
        return _finish_regroup(ent_rowptr, rp, local, va, lab, off, wt, d_e, hi, E, n, nnz, perm, entity_ids,
                               uniq_ptr, uniq_global)
    # entity index of every row / non-zero of the grouped batch (O(n) expansion of O(E) segments)
    ent_of_row = torch.repeat_interleave(torch.arange(E, device=dev), ent_rowptr[1:] - ent_rowptr[:-1])
    ent_of_nnz = torch.repeat_interleave(ent_of_row, rp[1:] - rp[:-1])
    fbits = max(1, int(num_features - 1).bit_length()) if num_features else _key_bits(gc.to(torch.int64))
    pair = (ent_of_nnz << fbits) | gc.to(torch.int64)
    pperm, pseg, pkey = group_by_entity(pair, key_bits=fbits + max(1, int(E - 1).bit_length()))
    # group g = one distinct (entity, feature); groups are ordered by entity, then by feature id
    uniq_global = (pkey & ((1 << fbits) - 1)).to(torch.int64)
    ent_of_group = pkey >> fbits
    d_e = torch.bincount(ent_of_group, minlength=E)
    uniq_ptr = torch.zeros(E + 1, dtype=torch.int64, device=dev)
    uniq_ptr[1:] = torch.cumsum(d_e, 0)
    group_of_sorted = torch.repeat_interleave(torch.arange(pkey.numel(), device=dev), pseg[1:] - pseg[:-1])
    local = torch.empty(nnz, dtype=torch.int32, device=dev)
    local[pperm.long()] = (group_of_sorted - uniq_ptr[ent_of_group[group_of_sorted]]).to(torch.int32)
    return _finish_regroup(ent_rowptr, rp, local, va, lab, off, wt, d_e, hi, E, n, nnz, perm, entity_ids, uniq_ptr,
                           uniq_global)


BITMAP_MAX_FEATURES = 2048   # feature bags up to this size are indexed by presence bitmaps (gdmix_local_index_*)
FORCE_PAIR_SORT = False      # tests: take the (entity, feature) pair sort regardless


def _finish_regroup(ent_rowptr, rp, local, va, lab, off, wt, d_e, hi, E, n, nnz, perm, entity_ids, uniq_ptr,
                    uniq_global):
    import torch
    dev = rp.device
    theta_ptr = torch.zeros(E + 1, dtype=torch.int64, device=dev)
    theta_ptr[1:] = torch.cumsum(d_e + hi, 0)
    rows_e = ent_rowptr[1:] - ent_rowptr[:-1]
    nnz_e = rp[ent_rowptr[1:]] - rp[ent_rowptr[:-1]]
    return {"ent_rowptr": ent_rowptr.contiguous(), "rowptr": rp, "col": local, "val": va, "label": lab, "offset": off,
            "weight": wt, "theta_ptr": theta_ptr, "n_entities": E, "n_rows": n, "nnz": nnz,
            "max_rows": int(rows_e.max().item()) if E else 0, "max_nnz": int(nnz_e.max().item()) if E else 0,
            "max_coef": int(d_e.max().item()) + hi if E else hi, "n_coef": int(theta_ptr[-1].item()),
            "perm": perm, "entity_ids": entity_ids, "uniq_ptr": uniq_ptr, "uniq_global": uniq_global}


class GroupedBatch:
    """A regroup_batch() result with the interface capi.re_fit_device / re_score_device expect of a DeviceBatch."""

    def __init__(self, d):
        from types import SimpleNamespace
        self.d = d
        self.ent_rowptr, self.rowptr, self.theta_ptr = d["ent_rowptr"], d["rowptr"], d["theta_ptr"]
        self.col, self.val, self.label = d["col"], d["val"], d["label"]
        self.weight, self.offset = d["weight"], d["offset"]
        self.host = SimpleNamespace(n_entities=d["n_entities"], n_rows=d["n_rows"], nnz=d["nnz"], n_coef=d["n_coef"],
                                    max_rows=d["max_rows"], max_nnz=d["max_nnz"], max_coef=d["max_coef"])

    def c_struct(self):
        h = self.host
        return capi.ReBatch(h.n_entities, h.n_rows, h.nnz, _tptr(self.ent_rowptr), _tptr(self.rowptr), _tptr(self.col),
                            _tptr(self.val), _tptr(self.label), _tptr(self.weight), _tptr(self.offset),
                            _tptr(self.theta_ptr), h.max_rows, h.max_nnz, h.max_coef, 0, None)

    def scatter_to_input_order(self, per_row):
        """per_row[i] belongs to grouped row i = input row perm[i] -> tensor in input row order."""
        import torch
        out = torch.empty_like(per_row)
        out[self.d["perm"].long()] = per_row
        return out


# ---- DataPartitioner's bounds, OffsetUpdater's join, and the partitioned files ------------------------------------------

def group_ids(entity, uid, lower_bound=None, upper_bound=None, stream=None):
    """Group id of every row (DataPartitioner.getGroupId): 0 = active; -1 = the entity has fewer than lower_bound rows;
    otherwise pmod(uid, rows // upper_bound + 1).  No bounds: all zeros.  int64 CUDA tensors in, int32 tensor out."""
    import torch
    n = entity.numel()
    out = torch.zeros(n, dtype=torch.int32, device=entity.device)
    if (lower_bound is None and upper_bound is None) or n == 0:
        return out
    perm, seg_ptr, _ = group_by_entity(entity, stream=stream)
    check(lib.gdmix_group_ids(_tptr(seg_ptr.contiguous()), C.c_int64(seg_ptr.numel() - 1), _tptr(perm), _tptr(uid.contiguous()),
                              C.c_int64(n), C.c_int32(int(lower_bound or 0)), C.c_int32(int(upper_bound or 0)), _tptr(out),
                              _stream_ptr(stream)))
    return out


def join_offsets(uid, score_uid, score, per_coordinate=None, stream=None):
    """OffsetUpdater.updateOffset on the device: -> (offset fp32[n], matched bool[n]); a data row without a score row of
    the same uid is one the reference's inner join drops."""
    import torch
    n, m = uid.numel(), score_uid.numel()
    skey, sperm = sort_pairs(score_uid.contiguous(), key_bits=64, stream=stream)
    off = torch.empty(n, dtype=torch.float32, device=uid.device)
    matched = torch.empty(n, dtype=torch.uint8, device=uid.device)
    check(lib.gdmix_offset_join(_tptr(uid.contiguous()), C.c_int64(n), _tptr(skey), _tptr(sperm), C.c_int64(m),
                                _tptr(score.contiguous()), _tptr(None if per_coordinate is None else per_coordinate.contiguous()),
                                _tptr(off), _tptr(matched), _stream_ptr(stream)))
    return off, matched.bool()


def partition_and_write(out_dir, entity, uid, rowptr, gcol, val, label, num_partitions, offset=None, weight=None,
                        scores=None, lower_bound=None, upper_bound=None, split=True, save_passive=True,
                        entity_name="entity", uid_name="uid", label_name="response", offset_name="offset",
                        weight_name="weight", bag="features", label_as_int=True, partition_list_file=None):
    """DataPartitioner.groupPartitionAndSaveDataset on the device + the files on disk.

    Rows in arrival order (int64 entity ids and uids, CSR feature bag with GLOBAL ids, fp32 label / offset / weight;
    CUDA tensors) -> `<out_dir>/active/partitionId=k/part-00000.tfrecord` (+ `passive/...` when bounds are given and
    save_passive; split=False -- validation data -- writes `<out_dir>/partitionId=k/...` without the split), one
    SequenceExample per (entity, group), and `partition_list_file` with the sorted partition ids that hold a record.
    scores = (uid, predictionScore[, predictionScorePerCoordinate]) CUDA tensors of the previous coordinate: rows are
    inner-joined on uid and the joined score becomes the offset column.  The whole regrouping -- join, bounds,
    (class, partition, entity, group) sort, CSR gather -- runs on the device; the host only encodes and writes the files.
    Returns a dict with the record counts per file."""
    import os
    import torch
    dev = entity.device
    n = entity.numel()
    keep = None
    if scores is not None:
        s_uid, s_score = scores[0], scores[1]
        s_pc = scores[2] if len(scores) > 2 else None
        offset, matched = join_offsets(uid, s_uid, s_score, s_pc)
        if not bool(matched.all().item()):
            # inner join: unmatched rows leave; a stable one-bit sort keeps the others in arrival order
            _, order = sort_pairs((~matched).to(torch.int64), key_bits=1)
            keep = order[:int(matched.sum().item())]
    if keep is not None:
        rowptr, gcol, val = gather_rows(rowptr, gcol, val, keep)
        entity, uid = entity[keep.long()], uid[keep.long()]
        label = gather_f32(label, keep)
        offset = gather_f32(offset, keep)
        weight = gather_f32(weight, keep) if weight is not None else None
        n = entity.numel()
    if n == 0:
        raise ValueError("no rows to partition")
    gid = group_ids(entity, uid, lower_bound, upper_bound)
    part = partition_ids(entity, num_partitions)
    if int(part.min().item()) < 0:
        raise ValueError("negative partition id (an entity whose String.hashCode is Int.MinValue): Spark would write partitionId=-k")
    _, seg_ptr, ent_ids = group_by_entity(entity)
    # entity rank of every row: position of its id among the sorted distinct ids
    rank = torch.searchsorted(ent_ids, entity)
    gmax = int(gid.max().item()) + 2
    gb, eb, pb = max(1, (gmax - 1).bit_length()), max(1, int(ent_ids.numel() - 1).bit_length()), max(1, int(num_partitions - 1).bit_length())
    if gb + eb + pb + 1 > 63:
        raise ValueError("too many entities x groups x partitions for one 64-bit sort key")
    passive = (gid != 0).to(torch.int64) if split else torch.zeros(n, dtype=torch.int64, device=dev)
    key = (passive << (pb + eb + gb)) | (part.to(torch.int64) << (eb + gb)) | (rank << gb) | (gid.to(torch.int64) + 1)
    perm, rec_ptr, rec_key = group_by_entity(key, key_bits=gb + eb + pb + 1)
    rp, gc, va = gather_rows(rowptr, gcol, val, perm)
    cols = {"uid": uid[perm.long()].cpu().numpy(), "label": gather_f32(label, perm).cpu().numpy(),
            "offset": None if offset is None else gather_f32(offset, perm).cpu().numpy(),
            "weight": None if weight is None else gather_f32(weight, perm).cpu().numpy()}
    rp_h, gc_h, va_h = rp.cpu().numpy(), gc.cpu().numpy().astype(np.int64), va.cpu().numpy()
    rec_ptr_h, rec_key_h = rec_ptr.cpu().numpy(), rec_key.cpu().numpy()
    ent_h = ent_ids.cpu().numpy()
    rec_passive = rec_key_h >> (pb + eb + gb)
    rec_part = (rec_key_h >> (eb + gb)) & ((1 << pb) - 1)
    rec_ent = ent_h[(rec_key_h >> gb) & ((1 << eb) - 1)]
    row_len = np.diff(rp_h)
    written = {}
    file_key = rec_passive * (1 << pb) + rec_part
    starts = np.flatnonzero(np.diff(file_key, prepend=-1))
    ends = np.append(starts[1:], file_key.shape[0])
    for a, b in zip(starts, ends):
        pas, k = int(rec_passive[a]), int(rec_part[a])
        if pas and not (save_passive and (lower_bound is not None or upper_bound is not None)):
            continue
        r0, r1 = int(rec_ptr_h[a]), int(rec_ptr_h[b])
        q0, q1 = int(rp_h[r0]), int(rp_h[r1])
        sub = os.path.join(out_dir, ("passive" if pas else "active") if split else "", f"partitionId={k}")
        os.makedirs(sub, exist_ok=True)
        image = capi.encode_entity_grouped(
            np.diff(rec_ptr_h[a:b + 1]), row_len[r0:r1], gc_h[q0:q1], va_h[q0:q1], cols["uid"][r0:r1],
            entity_int=rec_ent[a:b], label=cols["label"][r0:r1], label_as_int=label_as_int,
            offset=None if cols["offset"] is None else cols["offset"][r0:r1],
            weight=None if cols["weight"] is None else cols["weight"][r0:r1], entity=entity_name, uid_name=uid_name,
            label_name=label_name, offset_name=offset_name, weight_name=weight_name, bag=bag)
        with open(os.path.join(sub, "part-00000.tfrecord"), "wb") as f:
            f.write(image.tobytes())
        written[(("passive" if pas else "active") if split else "all", k)] = int(b - a)
    if partition_list_file:
        ids = sorted({int(k) for k in np.unique(rec_part)})
        os.makedirs(os.path.dirname(partition_list_file) or ".", exist_ok=True)
        with open(partition_list_file, "w") as f:
            f.write(",".join(str(k) for k in ids))
    return written
