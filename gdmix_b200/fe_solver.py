"""Fixed-effect LR solve: CUDA objective/gradient + all-reduce + replicated L-BFGS-B.

Mirrors the numerical core of the reference's ``FixedEffectLRModelLBFGS.train``
(gdmix-trainer/src/gdmix/models/custom/fixed_effect_lr_lbfgs_model.py):

  reference                                              here
  -----------------------------------------------------  ------------------------------------------------
  _train_model_fn: tf.while_loop over the worker's       gdmix_fe_loss_grad: one streaming CUDA pass over
  batches, sum of loss and gradient (:309-381)           this rank's rows -> fg = [value | gradient]
  collective_ops.all_reduce(value), (gradients)          ONE torch.distributed.all_reduce(fg) (NCCL over
  (:382-390, two collectives)                            NVLink; gloo in the CPU tests)
  fmin_l_bfgs_b replicated on every worker (:635-643)    gdmix_lbfgs_* replicated on every rank (same
                                                         algorithm as the per-entity device solver)
  threshold_coefficients (:648-649)                      abs(x) <= 1e-4 -> 0
  _scoring_fn (:214-270)                                 gdmix_fe_score

Coefficient layout is the reference's: features first, intercept LAST.  The per-rank L2 term is divided by
the number of workers before the reduction exactly as the reference does (:375-381).
"""
import numpy as np

from . import _capi as capi


def shard_rows(n_items, rank, world):
    """The reference shards by file, ``files[rank::world]`` (util/distribution_utils.py:46-47); the same
    strided rule applied to any list of units (files, row blocks)."""
    return list(range(n_items))[rank::world]


class FixedEffectSolver:
    """One rank's view of the fixed-effect problem.

    rows        capi.DeviceFeRows holding this rank's shard (device memory), or None when `local_eval` is given
    opts        capi.LrOpts (has_intercept, regularize_bias, l2, m, max_iter, factr, ...)
    n_features  D (x has D + has_intercept entries)
    group       torch.distributed process group, or None for single-process
    local_eval  optional callable x(np.float64[D+hi]) -> np.float64[1+D+hi]; the tests inject a CPU evaluator here
                to exercise the multi-rank plumbing without a GPU.  The product path never sets it.
    """

    def __init__(self, rows, opts, n_features=None, group=None, local_eval=None, device=None):
        import torch
        self.torch = torch
        self.rows = rows
        self.opts = opts
        self.group = group
        self.local_eval = local_eval
        self.n_features = int(n_features if n_features is not None else rows.n_features)
        self.hi = 1 if opts.has_intercept else 0
        self.n_coef = self.n_features + self.hi
        self.dist = torch.distributed if (torch.distributed.is_available() and
                                          torch.distributed.is_initialized()) else None
        self.world = self.dist.get_world_size(group) if self.dist else 1
        self.nfev = 0
        if local_eval is None:
            if rows is None:
                raise ValueError("FixedEffectSolver needs device rows (there is no CPU path)")
            self.device = rows.val.device
            self._x_dev = torch.empty(self.n_coef, dtype=torch.float64, device=self.device)
            self._fg_dev = torch.empty(1 + self.n_coef, dtype=torch.float64, device=self.device)
            self._fg_host = torch.empty(1 + self.n_coef, dtype=torch.float64).pin_memory()
            self.plan = None  # column-major copy + work items, built on the first evaluation
            self._ranked = None  # (rows with columns renumbered by falling frequency, permutation) -- see _prepare
        else:
            self.device = torch.device("cpu") if device is None else device

    # ---- objective ------------------------------------------------------------------------------------
    def loss_grad(self, x):
        """All-reduced (f, g) at x -- the reference's _compute_loss_and_gradients (:394-404)."""
        torch = self.torch
        self.nfev += 1
        if self.local_eval is not None:
            fg = torch.from_numpy(np.ascontiguousarray(self.local_eval(x), dtype=np.float64)).to(self.device)
            if self.dist and self.world > 1:
                self.dist.all_reduce(fg, group=self.group)
            fg = fg.cpu().numpy()
            return float(fg[0]), fg[1:].copy()
        self._x_dev.copy_(torch.from_numpy(x), non_blocking=False)
        if self.plan is None:
            self._prepare()
        if self._ranked is None:
            capi.fe_loss_grad_device(self.rows, self.opts, self._x_dev, fg=self._fg_dev, plan=self.plan)
        else:
            # evaluate in the shard's own frequency order, hand the result back in the caller's feature order
            # (two D-sized gathers; the all-reduce below needs one common order across ranks)
            rows_r, perm_x, perm_fg = self._ranked
            torch.index_select(self._x_dev, 0, perm_x, out=self._x_rank)
            capi.fe_loss_grad_device(rows_r, self.opts, self._x_rank, fg=self._fg_rank, plan=self.plan)
            self._fg_dev.index_copy_(0, perm_fg, self._fg_rank)
        if self.dist and self.world > 1:
            self.dist.all_reduce(self._fg_dev, group=self.group)  # value and gradient in one collective
        self._fg_host.copy_(self._fg_dev, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        fg = self._fg_host.numpy()
        return float(fg[0]), fg[1:].copy()

    def _prepare(self):
        """Once per training run: the column-major copy of the shard (capi.DeviceFePlan) and -- when x is larger
        than the part of it the rows kernel keeps in shared memory (capi.FE_HEAD, 8192 coefficients) -- this
        shard's features renumbered by falling frequency, so that the coefficients in shared memory are the ones
        most non-zeros multiply whatever order the feature file happens to list them in.  The permutation never
        leaves this object: x comes in and fg goes out in the caller's feature order."""
        torch = self.torch
        rows, D = self.rows, self.n_features
        if D > capi.FE_HEAD and rows.nnz > 0:
            counts = torch.bincount(rows.col.to(torch.int64), minlength=D)
            by_freq = torch.sort(counts, descending=True, stable=True).indices      # feature at rank r
            rank_of = torch.empty(D, dtype=torch.int32, device=self.device)
            rank_of[by_freq] = torch.arange(D, dtype=torch.int32, device=self.device)
            ranked = capi.DeviceFeRows.__new__(capi.DeviceFeRows)
            ranked.__dict__.update(rows.__dict__)
            ranked.col = rank_of[rows.col.to(torch.int64)].contiguous()
            tail = torch.arange(D, self.n_coef, dtype=torch.int64, device=self.device)  # the intercept stays last
            perm_x = torch.cat([by_freq, tail])
            perm_fg = torch.cat([torch.zeros(1, dtype=torch.int64, device=self.device), 1 + perm_x])
            self._ranked = (ranked, perm_x, perm_fg)
            self._x_rank = torch.empty_like(self._x_dev)
            self._fg_rank = torch.empty_like(self._fg_dev)
            self.plan = capi.DeviceFePlan(ranked)
        else:
            self.plan = capi.DeviceFePlan(rows)

    # ---- solve ----------------------------------------------------------------------------------------
    def fit(self, x0=None, threshold=None):
        """-> (x, info).  x0: previous model (same length) or None for zeros (:606-623)."""
        x = np.zeros(self.n_coef) if x0 is None else np.array(x0, dtype=np.float64, copy=True)
        assert x.shape == (self.n_coef,)
        solver = capi.HostLbfgs(self.n_coef, self.opts)
        f, g = self.loss_grad(x)
        while solver.iterate(x, f, g) == capi.HostLbfgs.NEED_FG:
            f, g = self.loss_grad(x)
        info = solver.info()
        solver.close()
        if threshold is not None:
            x = np.where(np.abs(x) <= threshold, 0.0, x)
        return x, info

    def score(self, x):
        """-> (logits incl. offset, per-coordinate logits) as fp32 numpy arrays for this rank's rows."""
        torch = self.torch
        xd = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).to(self.device)
        logit, per = capi.fe_score_device(self.rows, self.opts, xd)
        torch.cuda.synchronize()
        return logit.cpu().numpy(), per.cpu().numpy()
