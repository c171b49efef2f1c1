#!/usr/bin/env python
"""BASELINE.json configs[0]: the MovieLens-100k workflow of the reference -- global fixed effect -> per-user random effect
-> per-movie random effect, each coordinate trained on the previous one's scores as offsets -- on a seeded LOOK-ALIKE of
ml-100k (the dataset needs a download, the box has no network): 943 users, 1 682 movies, 100 000 ratings, 80 / 20 split,
the feature bags of scripts/download_process_movieLens_data.py:16-25,104-112 (global = 24 user + 20 movie features,
per_user = the 20 movie features, per_movie = the 24 user features; age / 100, release year / 2000, one-hot gender,
occupation and genres), label = rating > 3, weight 1, and the hyper-parameters of
gdmix-workflow/examples/movielens-100k/lr-movieLens.yaml (l2 1.0, bias unregularised, tolerance 1e-12, m = 10, 100
iterations, one partition).

Everything goes through the plugin classes and the files the DAG exchanges: per-record TFRecords ->
FixedEffectLRModelLBFGS.train -> score Avro -> partition.partition_and_write (DataPartitioner: offset join by uid,
group by entity, partitionId=0 layout) -> RandomEffectLRLBFGSModel.train -> score Avro -> ... ; validation AUC after
every coordinate by gdmix_auc (the Spark Evaluator's role).  `oracle=True` replays the same chain with the CPU oracle
(tests only).  The reference publishes 0.6237 / 0.7058 / 0.7599 on the real data (README.md:295-299); a look-alike cannot
reproduce those digits -- what must hold is that the AUC rises coordinate by coordinate and that the GPU chain and the
CPU chain agree.
Usage: python tools/movielens_lookalike.py [out_dir]"""
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

N_USERS, N_MOVIES, N_RATINGS = 943, 1682, 100_000
N_OCC, N_GENRE = 21, 19
D_USER, D_MOVIE = 3 + N_OCC, N_GENRE + 1          # 24, 20
README_AUC = {"global": 0.6237, "per-user": 0.7058, "per-movie": 0.7599}


def make_data(seed=7):
    """-> dict of per-rating arrays: user, movie, label, train mask, and CSR feature bags (global, per_user, per_movie)."""
    rng = np.random.default_rng(seed)
    age = rng.normal(34, 12, N_USERS).clip(7, 73) / 100.0
    male = rng.random(N_USERS) < 0.71
    occ = rng.integers(0, N_OCC, N_USERS)
    genres = rng.random((N_MOVIES, N_GENRE)) < (1.72 / N_GENRE)
    genres[np.arange(N_MOVIES), rng.integers(1, N_GENRE, N_MOVIES)] = True     # at least one genre
    year = rng.integers(1930, 1999, N_MOVIES) / 2000.0
    # activity: 20 .. 737 ratings per user (mean 106), popularity skew over movies
    act = rng.lognormal(0.0, 1.0, N_USERS)
    act = 20 + (act / act.sum()) * (N_RATINGS - 20 * N_USERS)
    user = np.repeat(np.arange(N_USERS), np.floor(act).astype(int))
    user = np.concatenate([user, rng.integers(0, N_USERS, N_RATINGS - user.size)])
    pop = rng.zipf(1.6, N_MOVIES).astype(np.float64)
    movie = rng.choice(N_MOVIES, N_RATINGS, p=pop / pop.sum())
    order = rng.permutation(N_RATINGS)
    user, movie = user[order], movie[order]
    # planted taste: a weak global part, a user bias + user x genre affinity, a movie quality + movie x occupation part
    w_genre = rng.normal(0, 0.25, N_GENRE)
    b_user = rng.normal(0, 0.9, N_USERS)
    a_user = rng.normal(0, 0.7, (N_USERS, N_GENRE))
    b_movie = rng.normal(0, 1.0, N_MOVIES)
    a_movie = rng.normal(0, 0.5, (N_MOVIES, N_OCC))
    z = 0.25 + genres[movie] @ w_genre + 0.8 * (age[user] - 0.34) + b_user[user] + \
        (a_user[user] * genres[movie]).sum(1) / np.sqrt(np.maximum(genres[movie].sum(1), 1)) + b_movie[movie] + \
        a_movie[movie, occ[user]]
    label = (rng.random(N_RATINGS) < 1.0 / (1.0 + np.exp(-z))).astype(np.float32)
    train = rng.random(N_RATINGS) < 0.8

    def user_bag(u):      # indices in USER_FEATURE_VALUES order: age, M, F, occupations
        return [0, 1 if male[u] else 2, 3 + int(occ[u])], [float(np.float32(age[u])), 1.0, 1.0]

    def movie_bag(m):     # genres then release_date
        g = np.flatnonzero(genres[m]).tolist()
        return g + [N_GENRE], [1.0] * len(g) + [float(np.float32(year[m]))]

    ub = [user_bag(u) for u in range(N_USERS)]
    mb = [movie_bag(m) for m in range(N_MOVIES)]
    bags = {"per_movie": ([ub[u][0] for u in user], [ub[u][1] for u in user], D_USER),
            "per_user": ([mb[m][0] for m in movie], [mb[m][1] for m in movie], D_MOVIE),
            "global": ([ub[u][0] + [D_USER + j for j in mb[m][0]] for u, m in zip(user, movie)],
                       [ub[u][1] + mb[m][1] for u, m in zip(user, movie)], D_USER + D_MOVIE)}
    out = {"user": user.astype(np.int64), "movie": movie.astype(np.int64), "label": label, "train": train,
           "uid": np.arange(N_RATINGS, dtype=np.int64)}
    for name, (idx, val, D) in bags.items():
        lens = np.array([len(r) for r in idx])
        out[name] = {"rowptr": np.concatenate([[0], np.cumsum(lens)]).astype(np.int64),
                     "col": np.concatenate(idx).astype(np.int32), "val": np.concatenate(val).astype(np.float32), "D": D}
    return out


def _csr_rows(bag, sel):
    rp, col, val = bag["rowptr"], bag["col"], bag["val"]
    lens = np.diff(rp)[sel]
    idx = np.concatenate([np.arange(rp[i], rp[i + 1]) for i in sel]) if len(sel) else np.zeros(0, np.int64)
    return np.concatenate([[0], np.cumsum(lens)]).astype(np.int64), col[idx], val[idx]


def _auc_np(score, label):
    order = np.argsort(score, kind="mergesort")
    s, y = score[order], label[order]
    ranks = np.empty(len(s))
    i = 0
    while i < len(s):
        j = i
        while j + 1 < len(s) and s[j + 1] == s[i]:
            j += 1
        ranks[i:j + 1] = 0.5 * (i + j) + 1
        i = j + 1
    pos = y > 0
    n1, n0 = pos.sum(), (~pos).sum()
    return float((ranks[pos].sum() - n1 * (n1 + 1) / 2) / (n1 * n0))


def run_oracle(data):
    """The chain on the CPU oracle: same objective, same solver options, fp32 scores between coordinates."""
    from oracle import oracle as O
    tr, va = np.flatnonzero(data["train"]), np.flatnonzero(~data["train"])
    y = data["label"]
    out = {}
    # global
    g = data["global"]
    rp, col, val = _csr_rows(g, tr)
    oo = O.make_opts(l2=1.0, regularize_bias=False, has_intercept=True)
    x, *_ = O.fe_fit(O.FeBlock(len(tr), g["D"], rp, col, val, y[tr]), oo)
    x = np.where(np.abs(x) <= 1e-4, 0.0, x)

    def fe_score(sel):
        rp, col, val = _csr_rows(g, sel)
        return np.array([val[rp[i]:rp[i + 1]].astype(np.float64) @ x[col[rp[i]:rp[i + 1]]] for i in range(len(sel))]) + x[-1]
    score = np.zeros(len(y), np.float32)
    score[tr], score[va] = fe_score(tr).astype(np.float32), fe_score(va).astype(np.float32)
    out["global"] = _auc_np(score[va], y[va])
    for name, key, bag in (("per-user", "user", "per_user"), ("per-movie", "movie", "per_movie")):
        b = data[bag]
        new = score.copy()
        ent = data[key]
        for e in np.unique(ent[tr]):
            rows = tr[ent[tr] == e]
            rp, col, val = _csr_rows(b, rows)
            uniq, local = np.unique(col, return_inverse=True)
            blk = O.EntityBlock(len(rows), len(uniq), rp, local.astype(np.int32), val, y[rows], None, score[rows])
            th, *_ = O.re_fit(blk, oo)
            th = O.threshold(th, 1e-4)
            lut = {int(u): j for j, u in enumerate(uniq)}
            for sel in (rows, va[ent[va] == e]):
                for i in sel:
                    sl = slice(b["rowptr"][i], b["rowptr"][i + 1])
                    zz = th[0] + sum(float(v) * th[1 + lut[int(c)]] for c, v in zip(b["col"][sl], b["val"][sl]) if int(c) in lut)
                    new[i] = np.float32(zz + float(score[i]))
        score = new
        out[name] = _auc_np(score[va], y[va])
    return out, score


def run_plugin(data, root):
    """The chain through the plugin classes and the DAG's files."""
    import torch
    from gdmix_b200 import FixedEffectLRModelLBFGS, RandomEffectLRLBFGSModel, constants
    from gdmix_b200 import partition as P
    from gdmix_b200.io import avro, tfrecord as T
    from gdmix_b200.params import Params, SchemaParams
    dev = torch.device("cuda", torch.cuda.current_device())
    sp = SchemaParams(uid_column_name="uid", weight_column_name="weight", label_column_name="response",
                      prediction_score_column_name="predictionScore")
    tr, va = np.flatnonzero(data["train"]), np.flatnonzero(~data["train"])
    y, uid = data["label"], data["uid"]
    aucs = {}

    def meta(path, bag, D, entity=None):
        feats = [{"name": bag, "dtype": "float", "shape": [D], "isSparse": True},
                 {"name": "weight", "dtype": "float", "shape": [], "isSparse": False},
                 {"name": "offset", "dtype": "float", "shape": [], "isSparse": False},
                 {"name": "uid", "dtype": "long", "shape": [], "isSparse": False}]
        if entity:
            feats.append({"name": entity, "dtype": "long", "shape": [], "isSparse": False})
        os.makedirs(os.path.dirname(path), exist_ok=True)
        json.dump({"features": feats, "labels": [{"name": "response", "dtype": "int", "shape": [], "isSparse": False}]},
                  open(path, "w"))

    def feature_file(path, D):
        os.makedirs(os.path.dirname(path), exist_ok=True)
        open(path, "w").write("".join(f"f{j},\n" for j in range(D)))

    def read_scores(path):
        recs = avro.read_records(path)
        return (np.array([r["uid"] for r in recs], np.int64), np.array([r["predictionScore"] for r in recs], np.float32))

    # ---- global fixed effect: per-record Examples ----
    g = data["global"]
    for sub, sel in (("trainingData", tr), ("validationData", va)):
        d = os.path.join(root, "global", sub)
        os.makedirs(d, exist_ok=True)
        with T.TFRecordWriter(os.path.join(d, "part-00000.tfrecord")) as w:
            for i in sel:
                sl = slice(g["rowptr"][i], g["rowptr"][i + 1])
                w.write(T.encode_example({"uid": T.encode_feature([int(uid[i])], "int64"),
                                          "response": T.encode_feature([int(y[i])], "int64"),
                                          "weight": T.encode_feature([1.0], "float"),
                                          "global_indices": T.encode_feature([int(c) for c in g["col"][sl]], "int64"),
                                          "global_values": T.encode_feature([float(v) for v in g["val"][sl]], "float")}))
    meta(os.path.join(root, "global", "metadata", "tensor_metadata.json"), "global", g["D"])
    feature_file(os.path.join(root, "global", "featureList", "global"), g["D"])
    base = Params(uid_column_name="uid", weight_column_name="weight", label_column_name="response",
                  prediction_score_column_name="predictionScore", action="train", stage="fixed_effect",
                  model_type="logistic_regression", training_score_dir=os.path.join(root, "global", "train_scores"),
                  validation_score_dir=os.path.join(root, "global", "validation_scores"))
    fe = FixedEffectLRModelLBFGS(raw_model_params=[
        "--uid_column_name", "uid", "--weight_column_name", "weight", "--label_column_name", "response",
        "--metadata_file", os.path.join(root, "global", "metadata", "tensor_metadata.json"),
        "--output_model_dir", os.path.join(root, "global", "models"), "--feature_bag", "global",
        "--feature_file", os.path.join(root, "global", "featureList", "global"), "--l2_reg_weight", "1.0",
        "--regularize_bias", "False"], base_training_params=base)
    fe.train(os.path.join(root, "global", "trainingData"), os.path.join(root, "global", "validationData"),
             os.path.join(root, "global", "metadata", "tensor_metadata.json"), os.path.join(root, "global", "models"),
             {constants.TASK_INDEX: 0, constants.NUM_WORKERS: 1, constants.IS_CHIEF: True}, sp)
    prev_train = read_scores(os.path.join(root, "global", "train_scores", "part-00000.avro"))
    prev_valid = read_scores(os.path.join(root, "global", "validation_scores", "part-00000.avro"))
    lab_of = lambda uids: torch.from_numpy(y[uids]).to(dev)
    aucs["global"] = P.auc(torch.from_numpy(prev_valid[1]).to(dev), lab_of(prev_valid[0]))

    # ---- random effects: DataPartitioner on the device, then the trainer ----
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    for name, key, bag in (("per-user", "user", "per_user"), ("per-movie", "movie", "per_movie")):
        b = data[bag]
        cdir = os.path.join(root, bag)
        for sub, sel, scores, split in (("trainingData", tr, prev_train, True), ("validationData", va, prev_valid, False)):
            rp, col, val = _csr_rows(b, sel)
            P.partition_and_write(os.path.join(cdir, sub), to(data[key][sel]), to(uid[sel]), to(rp), to(col), to(val),
                                  to(y[sel]), 1, weight=to(np.ones(len(sel), np.float32)),
                                  scores=(to(scores[0]), to(scores[1])), split=split, entity_name=key + "_id", bag=bag,
                                  partition_list_file=os.path.join(cdir, "partitionList.txt") if split else None)
        meta(os.path.join(cdir, "metadata", "tensor_metadata.json"), bag, b["D"], entity=key + "_id")
        feature_file(os.path.join(cdir, "featureList", bag), b["D"])
        re = RandomEffectLRLBFGSModel(raw_model_params=[
            "--uid_column_name", "uid", "--weight_column_name", "weight", "--label_column_name", "response",
            "--metadata_file", os.path.join(cdir, "metadata", "tensor_metadata.json"),
            "--output_model_dir", os.path.join(cdir, "models"), "--offset_column_name", "offset",
            "--partition_entity", key + "_id", "--l2_reg_weight", "1.0", "--regularize_bias", "False",
            "--feature_file", os.path.join(cdir, "featureList", bag), "--feature_bag", bag,
            "--enable_local_indexing", "False", "--training_data_dir", os.path.join(cdir, "trainingData")])
        ctx = {constants.PARTITION_INDEX: 0,
               constants.ACTIVE_TRAINING_OUTPUT_FILE: os.path.join(cdir, "train_scores", "partitionId=0", "part-00000-active.avro"),
               constants.VALIDATION_OUTPUT_FILE: os.path.join(cdir, "validation_scores", "partitionId=0", "part-00000.avro")}
        for f in (ctx[constants.ACTIVE_TRAINING_OUTPUT_FILE], ctx[constants.VALIDATION_OUTPUT_FILE]):
            os.makedirs(os.path.dirname(f), exist_ok=True)
        re.train(os.path.join(cdir, "trainingData", "active", "partitionId=0"),
                 os.path.join(cdir, "validationData", "partitionId=0"),
                 os.path.join(cdir, "metadata", "tensor_metadata.json"), os.path.join(cdir, "models"), ctx, sp)
        prev_train = read_scores(ctx[constants.ACTIVE_TRAINING_OUTPUT_FILE])
        prev_valid = read_scores(ctx[constants.VALIDATION_OUTPUT_FILE])
        aucs[name] = P.auc(torch.from_numpy(prev_valid[1]).to(dev), lab_of(prev_valid[0]))
    final = np.zeros(len(y), np.float32)
    final[prev_train[0]] = prev_train[1]
    final[prev_valid[0]] = prev_valid[1]
    return aucs, final


if __name__ == "__main__":
    root = sys.argv[1] if len(sys.argv) > 1 else tempfile.mkdtemp(prefix="gdmix_ml100k_")
    data = make_data()
    aucs, _ = run_plugin(data, root)
    print(json.dumps({"workload": "MovieLens-100k look-alike (943 users, 1682 movies, 100000 ratings, 80/20 split), global -> "
                                  "per-user -> per-movie through the plugin classes", "validation_auc": aucs,
                      "reference_readme_auc_on_real_data": README_AUC, "rows_train": int(data["train"].sum()),
                      "rows_validation": int((~data["train"]).sum())}))
