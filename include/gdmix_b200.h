/*
 * gdmix_b200.h -- C ABI of the B200-native random-effect / fixed-effect LR trainer.
 *
 * This is the drop-in boundary for ONE hot path of linkedin/gdmix: the per-entity
 * L2-regularised logistic-regression solves of the random-effect trainer, the
 * fixed-effect objective/gradient that feeds the global L-BFGS step, and the
 * scoring pass that follows both.  The reference has no FFI on this path -- its
 * numerical seam is Python calling scipy once per entity -- so every entry point
 * below names the reference call it replaces (paths relative to
 * gdmix-trainer/src/gdmix/ in the reference tree):
 *
 *   gdmix_re_fit         TrainingJobConsumer.__call__            models/custom/scipy/job_consumers.py:36-63
 *                        -> BinaryLogisticRegressionTrainer.fit  models/custom/binary_logistic_regression.py:191-239
 *                        -> scipy.optimize.fmin_l_bfgs_b         (L-BFGS-B 3.0, un-vendored dependency)
 *                        + threshold_coefficients                util/model_utils.py:4-12
 *                        + _compute_variance (SIMPLE)            binary_logistic_regression.py:144-189
 *   gdmix_re_fit_sweep   the same for several l2_reg_weight values at once (one training job per value in the
 *                        reference: base_lr_params.py:19 via the workflow's hyper-parameter loop)
 *   gdmix_re_loss_grad   _loss / _gradient                       binary_logistic_regression.py:84-131
 *   gdmix_re_score       InferenceJobConsumer.__call__           job_consumers.py:138-152
 *                        -> predict_proba(return_logits=True)    binary_logistic_regression.py:241-262
 *   gdmix_fe_loss_grad   _train_model_fn (per-worker partial)    models/custom/fixed_effect_lr_lbfgs_model.py:309-381
 *                        (the all-reduce at :382-390 stays with the caller: NCCL on the same stream)
 *   gdmix_fe_tile_plan_*, gdmix_fe_loss_grad_tiled   the same partial over a shard laid out once per training run
 *   gdmix_fe_lbfgs_*     fmin_l_bfgs_b replicated on every worker (:635-643), state resident on the device
 *   gdmix_fe_score       _scoring_fn                             fixed_effect_lr_lbfgs_model.py:214-307
 *   gdmix_partition_ids  getPartitionIdUDF                       gdmix-data/.../utils/PartitionUtils.scala:31-37
 *   gdmix_group_by_key, gdmix_csr_gather_rows, gdmix_local_index_*   groupBy(entity) + per-entity np.unique
 *                        (DataPartitioner.scala:296-379, job_consumers.py:243) for chained coordinates on the device
 *   gdmix_auc            Evaluator.calculateMetric("auc")         gdmix-data/.../evaluation/Evaluator.scala:29-45
 *   gdmix_avro_score_blocks, gdmix_avro_model_blocks   batched_write_avro / export_linear_model_to_avro (fastavro)
 *                                                                util/io_utils.py:102-212, 299-375
 *   gdmix_seqex_*, gdmix_example_*   per_entity_grouped_input_fn / per_record_input_fn (tf.data readers)
 *                                                                io/input_data_pipeline.py:223-273
 *
 * Conventions
 *   - All array pointers in gdmix_re_batch / gdmix_fe_rows are DEVICE pointers owned by
 *     the caller, except in the *_host entry points which take host pointers and do
 *     their own pinned staging, H2D/D2H copies and chunking.
 *   - Functions enqueue work on `stream` (a cudaStream_t passed as void*) and return
 *     without synchronising unless stated.  They are re-entrant per stream as long as
 *     each stream uses its own workspace.
 *   - Return value: 0 (GDMIX_OK) or a negative gdmix_status.  gdmix_last_error() gives a
 *     thread-local message.  No exceptions cross this boundary; nothing is allocated
 *     behind the caller's back by the device-pointer entry points.
 *   - Coefficient vectors: random effect = intercept FIRST (if has_intercept), then the
 *     entity's local features in ascending local index; fixed effect = intercept LAST
 *     (the reference's conventions, binary_logistic_regression.py:133-142 vs
 *     fixed_effect_lr_lbfgs_model.py:345-351).
 *   - Arithmetic: objective, gradient and all solver state are fp64 (trajectory parity
 *     with scipy requires it); feature values / labels / weights / offsets are fp32 in
 *     HBM, exactly what the TFRecords hold.
 */
#ifndef GDMIX_B200_H
#define GDMIX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define GDMIX_API
#else
#define GDMIX_API __attribute__((visibility("default")))
#endif

typedef enum gdmix_status {
    GDMIX_OK = 0,
    GDMIX_ERR_INVALID = -1,     /* bad argument (null pointer, negative size, m > GDMIX_MAX_M ...) */
    GDMIX_ERR_CUDA = -2,        /* a CUDA runtime call failed; message has cudaGetErrorString */
    GDMIX_ERR_WORKSPACE = -3,   /* workspace too small */
    GDMIX_ERR_TOO_LARGE = -4,   /* an entity exceeds what the kernels index (rows/features >= 65535) */
    GDMIX_ERR_NO_DEVICE = -5    /* no sm_100 device */
} gdmix_status;

#define GDMIX_MAX_SWEEP 16      /* most regularisation weights one gdmix_re_fit_sweep call takes */
#define GDMIX_MAX_M 32          /* largest supported number of L-BFGS curvature pairs */

/* Per-entity solver status, mirrors scipy's warnflag (fmin_l_bfgs_b). */
#define GDMIX_SOLVE_CONVERGED 0 /* pgtol or factr test passed */
#define GDMIX_SOLVE_MAXITER 1   /* maxiter / maxfun reached */
#define GDMIX_SOLVE_ABNORMAL 2  /* line search failed with empty memory */

#define GDMIX_VARIANCE_NONE 0
#define GDMIX_VARIANCE_SIMPLE 1
#define GDMIX_VARIANCE_FULL 2   /* diag(H^-1): a dense p x p inversion per entity after the solve */

/* A batch of entities in entity-local CSR form: what prepare_jobs
 * (job_consumers.py:161-296) hands to the consumers, flattened.  Entity e owns samples
 * [ent_rowptr[e], ent_rowptr[e+1]); sample i owns non-zeros [rowptr[i], rowptr[i+1]);
 * col holds the entity-LOCAL feature index (rank of the global id among the entity's
 * sorted unique ids, job_consumers.py:243); coefficients of entity e live at
 * theta[theta_ptr[e] .. theta_ptr[e+1]) with length d_e + has_intercept. */
typedef struct gdmix_re_batch {
    int64_t n_entities;
    int64_t n_rows;
    int64_t nnz;
    const int64_t *ent_rowptr; /* [n_entities + 1] */
    const int64_t *rowptr;     /* [n_rows + 1] */
    const int32_t *col;        /* [nnz] */
    const float *val;          /* [nnz] */
    const float *label;        /* [n_rows] 0/1 */
    const float *weight;       /* [n_rows] or NULL (= 1) */
    const float *offset;       /* [n_rows] or NULL (= 0) */
    const int64_t *theta_ptr;  /* [n_entities + 1] */
    /* Upper bounds over the batch, known to the host from ingest.  They size shared
     * memory and the launch geometry; the kernels re-check each entity against them. */
    int32_t max_rows;          /* max samples of one entity */
    int32_t max_nnz;           /* max non-zeros of one entity */
    int32_t max_coef;          /* max coefficients (d_e + has_intercept) of one entity */
    int32_t reserved;
    /* Optional, *_host entry points only: the same local column indices as 16-bit values (valid when every
     * entity has fewer than 65536 local features).  When set, `col` may be NULL: 2 instead of 4 bytes per
     * non-zero cross PCIe and are widened on the device. */
    const uint16_t *col16;
    /* Likewise as 8-bit values (valid when every entity has at most 256 local features -- the C1 shape): one
     * byte per non-zero crosses PCIe.  Takes precedence over col16. */
    const uint8_t *col8;
    /* Optional, gdmix_re_fit_host only: the non-zeros of every row as 16-bit values (valid when no row has more than
     * 65535).  When set, 2 instead of 8 bytes per row cross PCIe and the row pointers are rebuilt on the device by a
     * scan; `rowptr` stays required on the host side (the chunks are cut with it). */
    const uint16_t *row_len16;
    /* Likewise the 0/1 labels as bits, row i at bit (i & 7) of byte (i >> 3): 1/8 instead of 4 bytes per row.  When
     * set, `label` may be NULL. */
    const uint8_t *label_bits;
} gdmix_re_batch;

/* LRParams / scipy knobs (base_lr_params.py:5-42; scipy defaults for the rest). */
typedef struct gdmix_lr_opts {
    double l2;                  /* l2_reg_weight */
    double factr;               /* lbfgs_tolerance / eps */
    double pgtol;               /* 1e-5: scipy default, the reference never overrides it */
    double sparsity_threshold;  /* |theta| <= thr -> 0 on output; 0 disables */
    int32_t regularize_bias;
    int32_t has_intercept;
    int32_t m;                  /* num_of_lbfgs_curvature_pairs */
    int32_t max_iter;           /* num_of_lbfgs_iterations */
    int32_t max_ls;             /* 20 */
    int32_t max_fun;            /* 15000 */
    int32_t variance_mode;      /* GDMIX_VARIANCE_* */
    int32_t threads_per_entity; /* 0 = choose; else 32/64/128/256 */
} gdmix_lr_opts;

/* Rows of the fixed-effect shard this rank owns (per_record_input_fn output, flattened). */
typedef struct gdmix_fe_rows {
    int64_t n_rows;
    int64_t nnz;
    int64_t n_features;       /* D: x has D + has_intercept entries, intercept LAST */
    const int64_t *rowptr;    /* [n_rows + 1] */
    const int32_t *col;       /* [nnz] global feature index < D */
    const float *val;         /* [nnz] */
    const float *label;       /* [n_rows] */
    const float *weight;      /* [n_rows] or NULL */
    const float *offset;      /* [n_rows] or NULL */
    int32_t linear_regression; /* 0: logistic loss, 1: squared error (fixed_effect_lr_lbfgs_model.py:356-358) */
    int32_t num_workers;       /* the reference adds l2/num_workers per worker before the all-reduce (:375-381) */
} gdmix_fe_rows;

GDMIX_API const char *gdmix_last_error(void);
GDMIX_API const char *gdmix_version(void);

/* Device facts the host side needs for its own planning: SM count, max opt-in shared
 * memory per block, compute capability major*10+minor.  Any pointer may be NULL. */
GDMIX_API int gdmix_device_info(int32_t *sm_count, int32_t *smem_per_block_optin, int32_t *cc);

/* Bytes of device scratch gdmix_re_fit / gdmix_re_loss_grad need for this batch shape
 * (only n_entities, max_rows, max_nnz, max_coef and opts are read). */
GDMIX_API int gdmix_re_workspace_size(const gdmix_re_batch *batch, const gdmix_lr_opts *opts, size_t *bytes);

/* K1 test seam: f[e], g[theta_ptr[e]..] at a caller-supplied theta for every entity. */
GDMIX_API int gdmix_re_loss_grad(const gdmix_re_batch *batch, const gdmix_lr_opts *opts, const double *theta,
                                 double *f, double *g, void *workspace, size_t workspace_bytes, void *stream);

/* The hot path: stage each entity's block once into shared memory and run L-BFGS-B
 * (as scipy.fmin_l_bfgs_b runs it without bounds) to completion on chip.
 *   theta0      NULL = cold start (zeros), else warm-start coefficients (same layout as theta_out)
 *   theta_out   [theta_ptr[E]] coefficients (thresholded if opts->sparsity_threshold > 0)
 *   f_out,nit,nfev,status   [E], any may be NULL
 *   var_out     [theta_ptr[E]] or NULL; needs opts->variance_mode == GDMIX_VARIANCE_SIMPLE or _FULL
 *               (binary_logistic_regression.py:144-189; taken at the un-thresholded optimum) */
GDMIX_API int gdmix_re_fit(const gdmix_re_batch *batch, const gdmix_lr_opts *opts, const double *theta0,
                           double *theta_out, double *f_out, int32_t *nit, int32_t *nfev, int32_t *status,
                           double *var_out, void *workspace, size_t workspace_bytes, void *stream);

/* Regularisation sweep (BASELINE.json configs[4]): n_l2 models per entity, model j being exactly what
 * gdmix_re_fit returns with opts->l2 = l2_values[j] (same theta0 for every j), but each entity's block is read
 * from HBM and staged on chip ONCE for all of them.  In the reference a sweep over l2_reg_weight is n_l2 separate
 * training jobs over the same partitions (gdmix-workflow hyper-parameter loop; params l2_reg_weight at
 * base_lr_params.py:19).  l2_values is a HOST array, 1 <= n_l2 <= GDMIX_MAX_SWEEP.
 *   theta_out   [n_l2][coef_stride], coef_stride >= theta_ptr[E]; model j of entity e at
 *               theta_out + j * coef_stride + theta_ptr[e]
 *   f_out,nit,nfev,status   [n_l2][E], any may be NULL
 * Workspace: gdmix_re_workspace_size of the same batch / opts. */
GDMIX_API int gdmix_re_fit_sweep(const gdmix_re_batch *batch, const gdmix_lr_opts *opts, const double *l2_values,
                                 int32_t n_l2, const double *theta0, double *theta_out, int64_t coef_stride,
                                 double *f_out, int32_t *nit, int32_t *nfev, int32_t *status, void *workspace,
                                 size_t workspace_bytes, void *stream);

/* Scoring: logit = x.theta (+ intercept) + offset, per_coordinate = logit - offset, both
 * rounded to fp32 as the reference's Avro `float` fields are.  has_model[e] == 0 (or
 * theta == NULL) means "no model for this entity": logit = offset (job_consumers.py:145-146). */
GDMIX_API int gdmix_re_score(const gdmix_re_batch *batch, const gdmix_lr_opts *opts, const double *theta,
                             const uint8_t *has_model, float *logit, float *logit_per_coordinate, void *stream);

/* Fixed effect: fg[0] = sum_rows w*loss + l2/2*|x_reg|^2/num_workers, fg[1..D+has_intercept] =
 * gradient (same partial the reference all-reduces).  fg is zeroed by the call.  The caller
 * all-reduces fg (NCCL, same stream) and feeds it to the replicated L-BFGS step. */
GDMIX_API int gdmix_fe_loss_grad(const gdmix_fe_rows *rows, const gdmix_lr_opts *opts, const double *x, double *fg,
                                 void *stream);
/* The objective / gradient over a shard laid out once per training run (the shard does not change between the ~100
 * evaluations of an L-BFGS run).  The columns of `rows` are expected in falling-frequency order (rank 0 = most frequent
 * feature; gdmix_fe_column_counts + gdmix_remap_i32 produce that order) -- any order is correct, this one is fast.
 *   _create   builds, on the device: the row-major copy split into a HOT part (rank < hz, the coefficients the z pass
 *             keeps in shared memory; fp32 value + 16-bit rank) and a COLD part; the column-major copy of the ranks < hg
 *             tiled by `tile_rows` rows (fp32 value + 16-bit row + 16-bit rank; the g pass keeps a tile's dz and hg fp64
 *             accumulators in shared memory) and of the ranks >= hg tiled by `l2_tile_rows` rows (their dz gathers stay in
 *             L2).  0 for any of hz / hg / tile_rows / l2_tile_rows = choose.  Synchronises the stream; NULL on error.
 *   gdmix_fe_loss_grad_tiled   fg as gdmix_fe_loss_grad, without atomics: every sum has a fixed order (bitwise
 *             reproducible); enqueue-only, capturable in a CUDA graph.  `rows` supplies label / weight / offset.
 *   _info     out8 = { hz, hg, tile_rows, tiles, cold non-zeros of the z side, of the g side, plan bytes, L2 tiles } */
typedef struct gdmix_fe_tile_plan gdmix_fe_tile_plan;
GDMIX_API gdmix_fe_tile_plan *gdmix_fe_tile_plan_create(const gdmix_fe_rows *rows, int32_t hz, int32_t hg, int32_t tile_rows,
                                                        int64_t l2_tile_rows, void *stream);
GDMIX_API void gdmix_fe_tile_plan_destroy(gdmix_fe_tile_plan *plan);
GDMIX_API int gdmix_fe_tile_plan_info(const gdmix_fe_tile_plan *plan, int64_t *out8);
GDMIX_API int gdmix_fe_loss_grad_tiled(const gdmix_fe_rows *rows, const gdmix_fe_tile_plan *plan, const gdmix_lr_opts *opts,
                                       const double *x, double *fg, void *stream);

/* Set-up helpers of the planned objective (device pointers; once per training run, the shard does not change between
 * the evaluations of an L-BFGS run): non-zeros per feature (synchronises the stream; an index outside [0, n_features)
 * is an error), and out[i] = map[in[i]] (renumbering the shard's columns by falling frequency). */
GDMIX_API int gdmix_fe_column_counts(const int32_t *col, int64_t nnz, int64_t n_features, int64_t *counts, void *stream);
GDMIX_API int gdmix_remap_i32(const int32_t *in, const int32_t *map, int64_t n, int32_t *out, void *stream);

/* Hessian of the logistic loss over this rank's rows at x, X1^T diag(w rho (1-rho)) X1 with the intercept column
 * LAST -- the accumulator H of _scoring_fn (fixed_effect_lr_lbfgs_model.py:271-296; the reference keeps it in
 * fp32, here fp64).  mode GDMIX_VARIANCE_SIMPLE: h[D+hi] = diagonal; GDMIX_VARIANCE_FULL: h[(D+hi)^2] row-major.
 * h is zeroed by the call.  The caller all-reduces h, adds the L2 term and inverts (:451-463). */
GDMIX_API int gdmix_fe_hessian(const gdmix_fe_rows *rows, const gdmix_lr_opts *opts, const double *x, int32_t mode,
                               double *h, void *stream);
GDMIX_API int gdmix_fe_score(const gdmix_fe_rows *rows, const gdmix_lr_opts *opts, const double *x, float *logit,
                             float *logit_per_coordinate, void *stream);

/* Host-buffer form of gdmix_re_fit: every pointer (batch arrays, theta0, outputs) is HOST
 * memory.  The library owns the device buffers, splits the batch into chunks of about
 * `chunk_entities` entities (0 = choose), copies each chunk straight out of / into the caller's
 * arrays and overlaps H2D copy, solve and D2H copy of alternating chunks on two streams.  Synchronous: returns when the outputs are in host memory.
 * This is the call the Python plugin classes make. */
GDMIX_API int gdmix_re_fit_host(const gdmix_re_batch *host_batch, const gdmix_lr_opts *opts, const double *theta0,
                                double *theta_out, double *f_out, int32_t *nit, int32_t *nfev, int32_t *status,
                                double *var_out, int64_t chunk_entities);
GDMIX_API int gdmix_re_score_host(const gdmix_re_batch *host_batch, const gdmix_lr_opts *opts, const double *theta,
                                  const uint8_t *has_model, float *logit, float *logit_per_coordinate);
/* Page-locks / unlocks a caller buffer (cudaHostRegister) so that the *_host entry points can copy from
 * and to it asynchronously.  Optional: pageable buffers work, pinned ones overlap copy and compute. */
GDMIX_API int gdmix_host_register(void *ptr, size_t bytes);
GDMIX_API int gdmix_host_unregister(void *ptr);
/* Self-test of the solver's own exp(-|z|) / log(1 + t) / 1 / (1 + t) (csrc/re_common.cuh: logistic_terms) against
 * the CUDA math library on device arrays: out[6 i ..] = the three terms by logistic_terms, then by the library. */
GDMIX_API int gdmix_selftest_logistic(const double *z, int64_t n, double *out, void *stream);
/* Page-locked host memory of the library's own (cudaHostAlloc / cudaFreeHost): what the plugin's readers parse a
 * partition INTO, so that the *_host calls that follow copy at the link's rate instead of through the driver's
 * pageable staging (the reference has no counterpart: its arrays are numpy's). */
GDMIX_API int gdmix_pinned_alloc(size_t bytes, void **out);
GDMIX_API int gdmix_pinned_free(void *ptr);
/* int32 entity-local column indices -> 1 or 2 bytes each (width = 1 | 2), all host threads; the narrow form is what
 * gdmix_re_batch.col8 / col16 carry across PCIe. */
GDMIX_API int gdmix_narrow_columns(const int32_t *col, int64_t n, int32_t width, void *out);
/* Releases the cached device buffers and streams of the *_host entry points. */
GDMIX_API void gdmix_host_release(void);

/* Entity -> partition map, bit-exact with the JVM: abs(String.hashCode(id)) % num_partitions
 * (Math.abs(Int.MinValue) stays negative and % keeps the dividend's sign).  ids are UTF-16
 * code units, entity e owning units[id_ptr[e] .. id_ptr[e+1]).  Host function. */
GDMIX_API int gdmix_partition_ids(const uint16_t *units, const int64_t *id_ptr, int64_t n_ids,
                                  int32_t num_partitions, int32_t *hash_out, int32_t *partition_out);

/* ---- either side of the path, on the device (the Spark jobs DataPartitioner / Evaluator of gdmix-data) --------
 * All pointers are device pointers; `ws` is caller-owned scratch of gdmix_partition_workspace_size(n) bytes.
 *
 * gdmix_sort_pairs_u64    stable LSD radix sort of n keys (the low key_bits bits are significant); perm_out[i] is
 *                         the original index of the i-th smallest key.
 * gdmix_group_by_key      groupBy(entity): sort + segments.  perm brings rows of one entity together (original
 *                         order kept inside an entity), seg_ptr[g] .. seg_ptr[g+1] are entity g's rows in the
 *                         permuted order, seg_key[g] its key, *n_groups_dev the number of entities
 *                         (DataPartitioner.scala:296-379 without the bounds).
 * gdmix_csr_gather_rows   the sample block in the new row order (rowptr_out[n_rows+1], col_out / val_out[nnz]);
 *                         gdmix_gather_f32 does the same for labels / offsets / weights.
 * gdmix_partition_ids_i64 abs(String.valueOf(id).hashCode) % num_partitions for integer entity ids
 *                         (PartitionUtils.scala:31-37), bit-exact with gdmix_partition_ids on the decimal string.
 * gdmix_auc               area under the ROC curve with ties counted half (Evaluator.scala:29-45 ->
 *                         BinaryClassificationMetrics.areaUnderROC); label > 0 is positive.
 *                         out3 = { auc, positives, negatives }.  Synchronises the stream once.
 * gdmix_group_ids         DataPartitioner.getGroupId (DataPartitioner.scala:335-379): rows grouped by entity (perm /
 *                         seg_ptr of gdmix_group_by_key) -> group id per INPUT row: 0 active; -1 the entity has fewer than
 *                         lower_bound rows; otherwise pmod(uid, rows / upper_bound + 1) (bounds <= 0: absent).
 * gdmix_offset_join       OffsetUpdater.updateOffset (OffsetUpdater.scala:105-129): offset = float(predictionScore)
 *                         [- predictionScorePerCoordinate] of the score row with the same uid; matched[i] = 0: no such
 *                         row (the inner join drops the data row).  score_uid_sorted / score_perm: the scores' uids
 *                         (as u64) through gdmix_sort_pairs_u64. */
GDMIX_API int gdmix_group_ids(const int64_t *seg_ptr, int64_t n_groups, const uint32_t *perm, const int64_t *uid, int64_t n,
                              int32_t lower_bound, int32_t upper_bound, int32_t *group_id, void *stream);
GDMIX_API int gdmix_offset_join(const int64_t *uid, int64_t n, const uint64_t *score_uid_sorted, const uint32_t *score_perm,
                                int64_t m, const float *score, const float *per_coordinate, float *offset_out,
                                uint8_t *matched, void *stream);
GDMIX_API int gdmix_partition_workspace_size(int64_t n, size_t *bytes);
GDMIX_API int gdmix_sort_pairs_u64(const uint64_t *keys_in, int64_t n, int32_t key_bits, uint64_t *keys_out,
                                   uint32_t *perm_out, void *ws, size_t ws_bytes, void *stream);
GDMIX_API int gdmix_group_by_key(const uint64_t *keys, int64_t n, int32_t key_bits, uint64_t *keys_sorted, uint32_t *perm,
                                 int64_t *seg_ptr, uint64_t *seg_key, int64_t *n_groups_dev, void *ws, size_t ws_bytes,
                                 void *stream);
GDMIX_API int gdmix_csr_gather_rows(const int64_t *rowptr_in, const int32_t *col_in, const float *val_in,
                                    const uint32_t *perm, int64_t n_rows, int64_t *rowptr_out, int32_t *col_out,
                                    float *val_out, void *ws, size_t ws_bytes, void *stream);
GDMIX_API int gdmix_gather_f32(const float *in, const uint32_t *perm, int64_t n, float *out, void *stream);
GDMIX_API int gdmix_partition_ids_i64(const int64_t *ids, int64_t n, int32_t num_partitions, int32_t *partition_out,
                                      void *stream);
GDMIX_API int gdmix_auc(const float *score, const float *label, int64_t n, double *out3, void *ws, size_t ws_bytes,
                        void *stream);
/* Entity-local feature indexing (np.unique(cols, return_inverse=True) per entity, job_consumers.py:243) for feature
 * bags of at most a few thousand ids, by per-entity presence bitmaps instead of an (entity, feature) pair sort.
 * The batch is already grouped: entity e owns samples [ent_rowptr[e], ent_rowptr[e+1]).
 *   mark:   bitmap[E * W32 + 1] (W32 = ceil(num_features / 32); the last word is the call's out-of-range flag),
 *           word_prefix[E * W32], d_e[E] = distinct features per entity;
 *           synchronises the stream once (ids outside [0, num_features) are an error)
 *   apply:  with uniq_ptr[E+1] = exclusive scan of d_e (the caller's), local_col[nnz] = rank of each non-zero's feature
 *           among its entity's distinct ids (may alias gcol), uniq_global[uniq_ptr[E]] = those ids, ascending */
GDMIX_API int gdmix_local_index_mark(const int64_t *ent_rowptr, int64_t n_entities, const int64_t *rowptr,
                                     const int32_t *gcol, int64_t n_rows, int32_t num_features, uint32_t *bitmap,
                                     uint32_t *word_prefix, int64_t *d_e, void *stream);
GDMIX_API int gdmix_local_index_apply(const int64_t *ent_rowptr, int64_t n_entities, const int64_t *rowptr,
                                      const int32_t *gcol, int64_t n_rows, int32_t num_features, const uint32_t *bitmap,
                                      const uint32_t *word_prefix, const int64_t *uniq_ptr, int32_t *local_col,
                                      int64_t *uniq_global, void *stream);

/* Host-side reader of entity-grouped TFRecord files (no TensorFlow): one uncompressed file image -> the flat arrays
 * of the random-effect ingest, i.e. what per_entity_grouped_input_fn + prepare_jobs build per entity in the reference
 * (input_data_pipeline.py:244-273, job_consumers.py:161-258).  One record = one tf.train.SequenceExample = one entity:
 * context holds the entity id (int64 or bytes, one value) and one list per sample column; feature_lists hold
 * <bag>_indices (int64 list per sample) and <bag>_values (float list per sample).  Names are NUL-terminated strings;
 * label / offset / weight may be NULL (column not wanted).  Two calls over the same buffer:
 *   gdmix_seqex_count  -> sizes (entities, samples, non-zeros, bytes of all entity-id strings, flags)
 *   gdmix_seqex_fill   -> ent_rows[E] samples per entity, row_len[N], gcol[nnz] GLOBAL feature ids, val[nnz],
 *                         uid[N], label[N] (meaningful when sizes.all_labelled), offset[N] (0 when the column is
 *                         absent), weight[N] (1 when absent), id_chars[id_bytes] + id_ptr[E+1] (entity ids as
 *                         decimal / utf-8 strings, the modelId the trainer writes)
 * Malformed input is an error (gdmix_last_error), never a silent skip.  Pure host code: no device is touched. */
typedef struct gdmix_seqex_spec {
    const char *entity, *uid, *label, *offset, *weight, *bag_indices, *bag_values;
} gdmix_seqex_spec;
typedef struct gdmix_seqex_sizes {
    int64_t n_entities, n_rows, nnz, id_bytes;
    int32_t all_labelled, saw_weight;
    int64_t min_index, max_index;   /* smallest / largest feature index DECODED (INT64_MAX / INT64_MIN when none was:
                                     * the counting pass skips over packed index lists; gdmix_seqex_fill reports the
                                     * range of what it wrote through index_range[2], which may be NULL) */
} gdmix_seqex_sizes;
GDMIX_API int gdmix_seqex_count(const uint8_t *file_image, int64_t len, const gdmix_seqex_spec *spec,
                                gdmix_seqex_sizes *sizes);
GDMIX_API int gdmix_seqex_fill(const uint8_t *file_image, int64_t len, const gdmix_seqex_spec *spec, int64_t *ent_rows,
                               int64_t *row_len, int64_t *gcol, float *val, int64_t *uid, float *label, float *offset,
                               float *weight, char *id_chars, int64_t *id_ptr, int64_t *index_range);
/* gdmix_seqex_fill with the entity-local indexing of gdmix_local_index_host fused into it (an entity's indices are
 * ranked while they are still in cache, and the int64 global columns are never written): local16[nnz] = rank of every
 * index among its entity's distinct indices, d_e[E] their number, uniq_scratch[nnz] = the sorted distinct indices of
 * every entity parked at its first non-zero (gather them with gdmix_local_index_host's second call).
 * GDMIX_ERR_TOO_LARGE when an entity has more than 65535 distinct features: use the two separate calls. */
GDMIX_API int gdmix_seqex_fill_local(const uint8_t *file_image, int64_t len, const gdmix_seqex_spec *spec,
                                     int64_t *ent_rows, int64_t *row_len, uint16_t *local16, int64_t *d_e,
                                     int64_t *uniq_scratch, float *val, int64_t *uid, float *label, float *offset,
                                     float *weight, char *id_chars, int64_t *id_ptr, int64_t *index_range);

/* Entity-local feature indexing of a parsed partition on the host (np.unique(cols, return_inverse=True) per entity,
 * job_consumers.py:243), all host threads.  Two calls: uniq_global == NULL -> local_col[nnz], d_e[E] and the distinct
 * ids parked in scratch (int64[nnz]); then, with uniq_ptr[E+1] = exclusive scan of d_e, uniq_global[uniq_ptr[E]]. */
GDMIX_API int gdmix_local_index_host(const int64_t *ent_rowptr, const int64_t *rowptr, const int64_t *gcol,
                                     int64_t n_entities, int32_t *local_col, int64_t *d_e, int64_t *scratch,
                                     const int64_t *uniq_ptr, int64_t *uniq_global);

/* The writer of the same files (what DataPartitioner's Spark job saves, DataPartitioner.scala:203-280 ->
 * IoUtils.saveDataFrame with recordType SequenceExample): n_entities records, record e = ent_rows[e] consecutive samples;
 * the entity id is entity_int[e] (int64 list) or the utf-8 string id_chars[id_ptr[e] .. id_ptr[e+1]) (bytes list); spec
 * names the columns to write (NULL: not written); label_as_int writes the label as an int64 list.  out == NULL:
 * *written = bytes needed.  TFRecord framing with masked crc32c; all host threads.  Pure host code. */
GDMIX_API int gdmix_seqex_encode(const gdmix_seqex_spec *spec, int64_t n_entities, const int64_t *ent_rows,
                                 const int64_t *entity_int, const char *id_chars, const int64_t *id_ptr,
                                 const int64_t *row_len, const int64_t *gcol, const float *val, const int64_t *uid,
                                 const float *label, int32_t label_as_int, const float *offset, const float *weight,
                                 uint8_t *out, int64_t capacity, int64_t *written);

/* The same for the fixed effect's input, one tf.train.Example per row (per_record_input_fn,
 * input_data_pipeline.py:223-243): spec->bag_indices / bag_values name two lists of the features map (NULL:
 * intercept-only rows), uid / label / offset / weight scalar columns (first element; absent: 0 / NaN / 0 / 1);
 * spec->entity is ignored.  sizes.n_rows rows, sizes.nnz non-zeros. */
GDMIX_API int gdmix_example_count(const uint8_t *file_image, int64_t len, const gdmix_seqex_spec *spec,
                                  gdmix_seqex_sizes *sizes);
GDMIX_API int gdmix_example_fill(const uint8_t *file_image, int64_t len, const gdmix_seqex_spec *spec, int64_t *row_len,
                                 int32_t *col, float *val, int64_t *uid, float *label, float *offset, float *weight);

/* Avro encoding of the score records both trainers write (`validation_result`, util/io_utils.py:367-375, appended in
 * blocks of 1024 by batched_write_avro :299-334): the BODY of an object-container file -- per block <count> <size>
 * <records> <sync> -- for n records with fields uid (long), predictionScore (float), label ([null, float]; label ==
 * NULL writes the null branch), weight (float; NULL: the schema has no such field), predictionScorePerCoordinate
 * (float; NULL likewise).  The caller writes the container header (magic, schema, codec "null", sync) in front.
 * capacity >= n * 31 + blocks * 36.  Host code. */
GDMIX_API int gdmix_avro_score_blocks(const int64_t *uid, const float *score, const float *label, const float *weight,
                                      const float *per_coordinate, int64_t n, int32_t records_per_block,
                                      const uint8_t *sync16, uint8_t *out, int64_t capacity, int64_t *written);

/* Avro encoding of model files: Photon-ML BayesianLinearModelAvro records as gen_one_avro_model writes them
 * (util/io_utils.py:102-160, models/schemas.py:3-51): modelId, modelClass, means = the intercept first (always), then
 * every feature with |coefficient| > threshold as (name, term, value); variances aligned with means, or null when
 * var == NULL; lossFunction "".  Model m owns coef[coef_ptr[m] .. coef_ptr[m+1]) (intercept first when has_intercept)
 * and its non-intercept coefficients' GLOBAL feature ids sit in feat_idx in the same order (intercepts not listed);
 * names / terms come from the feature file as concatenated utf-8 strings with offset tables.  Blocks of
 * records_per_block records as in gdmix_avro_score_blocks.  out == NULL: *written = bytes needed. */
typedef struct gdmix_model_table {
    int64_t n_models;
    const char *id_chars; const int64_t *id_ptr;
    const char *model_class;
    const double *coef, *var; const int64_t *coef_ptr;
    const int64_t *feat_idx;
    int32_t has_intercept; int32_t reserved;
    double threshold;
    const char *intercept_name;
    const char *name_chars; const int64_t *name_ptr; const char *term_chars; const int64_t *term_ptr;
    int64_t n_features;
} gdmix_model_table;
GDMIX_API int gdmix_avro_model_blocks(const gdmix_model_table *table, int32_t records_per_block, const uint8_t *sync16,
                                      uint8_t *out, int64_t capacity, int64_t *written);
/* The same records into a buffer the library allocates to the exact size (one sizing pass and one writing pass
 * instead of a sizing call plus a writing call that sizes again): *out is released with gdmix_buffer_free. */
GDMIX_API int gdmix_avro_model_blocks_alloc(const gdmix_model_table *table, int32_t records_per_block,
                                            const uint8_t *sync16, uint8_t **out, int64_t *written);
GDMIX_API void gdmix_buffer_free(void *ptr);

/* Reading model files back (warm starts, the predict action; io_utils.py:163-212 / _load_weights,
 * random_effect_lr_lbfgs_model.py:262-309): the records of ONE container block (uncompressed) -> flat arrays.
 * A feature map handle holds the feature file's (name, term) -> row index table.  Call once with id_chars == NULL
 * for the sizes (*n_means, *id_bytes), then with the arrays: id_chars / id_ptr[n+1] (modelId strings), mean_ptr[n+1],
 * mean_feat (feature row, -1 = the intercept) / mean_val per entry, var_val aligned (0 where has_var[m] == 0).
 * A (name, term) missing from the feature file, or variances not aligned with means, is an error. */
typedef struct gdmix_feature_map gdmix_feature_map;
GDMIX_API gdmix_feature_map *gdmix_feature_map_create(const char *name_chars, const int64_t *name_ptr,
                                                      const char *term_chars, const int64_t *term_ptr,
                                                      int64_t n_features, const char *intercept_name);
GDMIX_API void gdmix_feature_map_destroy(gdmix_feature_map *map);
GDMIX_API int gdmix_avro_model_decode(const gdmix_feature_map *map, const uint8_t *block, int64_t len, int64_t n_records,
                                      int64_t *n_means, int64_t *id_bytes, char *id_chars, int64_t *id_ptr,
                                      int64_t *mean_ptr, int64_t *mean_feat, double *mean_val, double *var_val,
                                      uint8_t *has_var);

/* Replicated host-side solver state of the fixed-effect solve: L-BFGS-B without bounds, reverse
 * communication, the role scipy.optimize.fmin_l_bfgs_b plays at fixed_effect_lr_lbfgs_model.py:635-643.
 *   h = gdmix_lbfgs_create(n, opts)            (uses opts->m, max_iter, max_ls, max_fun, factr, pgtol)
 *   task = gdmix_lbfgs_iterate(h, x, f, g)     f, g = all-reduced objective / gradient at x (host memory)
 *       1: evaluate at the x just written and call again;  0: finished;  < 0: gdmix_status
 *   gdmix_lbfgs_info(h, &nit, &nfev, &status, &f)   scipy's nit / funcalls / warnflag / final f */
typedef struct gdmix_lbfgs gdmix_lbfgs;
GDMIX_API gdmix_lbfgs *gdmix_lbfgs_create(int64_t n, const gdmix_lr_opts *opts);
GDMIX_API int gdmix_lbfgs_iterate(gdmix_lbfgs *h, double *x, double f, const double *g);
GDMIX_API int gdmix_lbfgs_info(const gdmix_lbfgs *h, int32_t *nit, int32_t *nfev, int32_t *status, double *f);
GDMIX_API void gdmix_lbfgs_destroy(gdmix_lbfgs *h);

/* The same replicated solver with its state RESIDENT ON THE DEVICE (x, gradient, direction, previous iterate, the m
 * curvature pairs): what fixed_effect_lr_lbfgs_model.py:635-643 keeps in every worker's scipy instance.  An
 * evaluation then never leaves the GPU: gdmix_fe_loss_grad_planned -> all-reduce of fg on the same stream ->
 * gdmix_fe_lbfgs_step -> next evaluation; only a 24-byte status record crosses to the host (gdmix_fe_lbfgs_poll).
 *   x_dev[n], fg_dev[1 + n]   caller-owned device arrays: the iterate (start point in, solution out) and the
 *                             all-reduced [value | gradient] at x_dev
 *   _reset   (re)starts a solve from the x_dev contents;  enqueue-only
 *   _step    consumes fg_dev at x_dev and writes the next trial point into x_dev (or finishes); enqueue-only, a fixed
 *            sequence of 9 + 2m launches whatever the branch taken, so the caller may capture it in a CUDA graph
 *   _poll    synchronises the stream; *task = 1: evaluate at x_dev and step again, 0: finished, 2: step again without
 *            evaluating (restart from steepest descent after a failed line search); nit / nfev / status / f as
 *            gdmix_lbfgs_info
 * Every inner product has a fixed summation order, so ranks fed the same reduced fg keep bit-identical state. */
typedef struct gdmix_fe_lbfgs gdmix_fe_lbfgs;
GDMIX_API gdmix_fe_lbfgs *gdmix_fe_lbfgs_create(int64_t n, const gdmix_lr_opts *opts, double *x_dev, double *fg_dev);
GDMIX_API int gdmix_fe_lbfgs_reset(gdmix_fe_lbfgs *h, void *stream);
GDMIX_API int gdmix_fe_lbfgs_step(gdmix_fe_lbfgs *h, void *stream);
GDMIX_API int gdmix_fe_lbfgs_poll(gdmix_fe_lbfgs *h, void *stream, int32_t *task, int32_t *nit, int32_t *nfev,
                                  int32_t *status, double *f);
GDMIX_API void gdmix_fe_lbfgs_destroy(gdmix_fe_lbfgs *h);

/* Launch plan of this thread's most recent gdmix_re_fit / gdmix_re_loss_grad (diagnostics, tests, bench):
 * out8 = { fast kernel used, threads per entity, features per thread (fast) , CTAs per SM,
 *          sliced-ELL capacity in steps (fast), dynamic shared memory per CTA, grid, history in global arena }.
 * After a fast-path call the number of entities it deferred to the general kernel is the third int32 of the
 * workspace. */
GDMIX_API void gdmix_re_last_plan(int32_t *out8);
/* Ragged batches (largest entity well above 2.5 x the mean sample count) get a first fast-kernel launch planned
 * for the typical entity, ahead of the one gdmix_re_last_plan describes (which then only drains what the first
 * deferred): out8 = { planned (0/1), threads per entity, features per thread, CTAs per SM, sliced-ELL capacity,
 * dynamic shared memory per CTA, rows per entity it is planned for, 0 }.  Entities it deferred: eighth int32 of
 * the workspace. */
GDMIX_API void gdmix_re_last_plan_typical(int32_t *out8);
/* Batches of small entities (typically a few hundred non-zeros, at most 96 coefficients; fits of one model without
 * variance output) get a launch of the warp-per-entity kernel ahead of everything else; the launches above then only
 * drain what it deferred: out8 = { used (0/1), warps per CTA, coefficient slots per lane, CTAs per SM, rows a slice
 * holds, non-zeros a slice holds, dynamic shared memory per CTA, grid }.  Entities it deferred: twelfth int32 of the
 * workspace. */
GDMIX_API void gdmix_re_last_plan_small(int32_t *out8);

/* Number of kernel launches issued by this library since load (for bench accounting). */
GDMIX_API int64_t gdmix_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* GDMIX_B200_H */
