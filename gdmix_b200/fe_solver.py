"""Fixed-effect LR solve: CUDA objective/gradient + all-reduce + replicated L-BFGS-B, all resident on the device.

Mirrors the numerical core of the reference's ``FixedEffectLRModelLBFGS.train``
(gdmix-trainer/src/gdmix/models/custom/fixed_effect_lr_lbfgs_model.py):

  reference                                              here
  -----------------------------------------------------  ------------------------------------------------
  _train_model_fn: tf.while_loop over the worker's       gdmix_fe_loss_grad_planned: CUDA passes over this
  batches, sum of loss and gradient (:309-381)           rank's rows -> fg = [value | gradient]
  collective_ops.all_reduce(value), (gradients)          ONE torch.distributed.all_reduce(fg) on the same
  (:382-390, two collectives)                            stream (NCCL over NVLink; gloo in the CPU tests)
  fmin_l_bfgs_b replicated on every worker (:635-643)    gdmix_fe_lbfgs_* replicated on every rank, state in
                                                         HBM: per evaluation only a 24-byte status record
                                                         crosses to the host (solver="host": gdmix_lbfgs_*,
                                                         the same state machine on the host)
  threshold_coefficients (:648-649)                      abs(x) <= 1e-4 -> 0
  _scoring_fn (:214-270)                                 gdmix_fe_score

Coefficient layout is the reference's: features first, intercept LAST.  The per-rank L2 term is divided by
the number of workers before the reduction exactly as the reference does (:375-381).

Feature order on the device: when x is larger than the part of it the rows kernel keeps in shared memory
(capi.FE_HEAD coefficients) the features are renumbered by falling GLOBAL frequency (non-zero counts all-reduced
over the ranks, so every rank uses the same order and the all-reduce of fg needs no per-evaluation permutation).
The permutation never leaves this object: x0 comes in and x goes out in the caller's feature order.
"""
import numpy as np

from . import _capi as capi


def shard_rows(n_items, rank, world):
    """The reference shards by file, ``files[rank::world]`` (util/distribution_utils.py:46-47); the same
    strided rule applied to any list of units (files, row blocks)."""
    return list(range(n_items))[rank::world]


class FixedEffectSolver:
    """One rank's view of the fixed-effect problem.

    rows        capi.DeviceFeRows holding this rank's shard (device memory)
    opts        capi.LrOpts (has_intercept, regularize_bias, l2, m, max_iter, factr, ...)
    n_features  D (x has D + has_intercept entries)
    group       torch.distributed process group, or None for single-process
    solver      "device" (default): L-BFGS state in HBM (gdmix_fe_lbfgs_*); "host": gdmix_lbfgs_* fed through a
                pinned copy of fg per evaluation (the round-1 path, kept as a cross-check of the device solver)
    profile     record CUDA-event times of the three phases of every evaluation (kernels / all-reduce / step)
    """

    def __init__(self, rows, opts, n_features=None, group=None, solver="device", profile=False, plan_args=None):
        import torch
        self.torch = torch
        self.rows = rows
        self.opts = opts
        self.group = group
        self.solver = solver
        self.profile = profile
        self.plan_args = dict(plan_args or {})   # hz / hg / tile_rows / l2_tile_rows of capi.DeviceFeTilePlan (tests)
        self.n_features = int(n_features if n_features is not None else rows.n_features)
        self.hi = 1 if opts.has_intercept else 0
        self.n_coef = self.n_features + self.hi
        self.dist = torch.distributed if (torch.distributed.is_available() and
                                          torch.distributed.is_initialized()) else None
        self.world = self.dist.get_world_size(group) if self.dist else 1
        self.nfev = 0
        self.phase_ms = []          # profile: [(kernels, allreduce, step)] per evaluation
        self._setup()

    # ---- device state --------------------------------------------------------------------------------
    def _setup(self):
        torch = self.torch
        if self.rows is None:
            raise ValueError("FixedEffectSolver needs device rows (there is no CPU path)")
        self.device = self.rows.val.device
        self._x_dev = torch.zeros(self.n_coef, dtype=torch.float64, device=self.device)     # device feature order
        self._fg_dev = torch.zeros(1 + self.n_coef, dtype=torch.float64, device=self.device)
        self._fg_host = torch.empty(1 + self.n_coef, dtype=torch.float64).pin_memory()
        self.plan = None      # column-major copy + work items, built on the first evaluation
        self._perm = None     # device order -> caller's order (None: identity)
        self._rows_eval = self.rows

    def _prepare(self):
        """Once per training run: the feature order of the device (see the module docstring) and the column-major
        layout of the shard for the tiled objective (capi.DeviceFeTilePlan)."""
        torch = self.torch
        rows, D = self.rows, self.n_features
        if D > capi.FE_HEAD:
            counts = capi.fe_column_counts(rows)                         # int64[D], this rank's non-zeros per feature
            if self.dist and self.world > 1:
                self.dist.all_reduce(counts, group=self.group)
            by_freq = capi.rank_by_count(counts)                         # feature at rank r (stable: ties by feature id)
            rank_of = torch.empty(D, dtype=torch.int32, device=self.device)
            rank_of[by_freq] = torch.arange(D, dtype=torch.int32, device=self.device)
            ranked = capi.DeviceFeRows.__new__(capi.DeviceFeRows)
            ranked.__dict__.update(rows.__dict__)
            ranked.col = capi.remap_columns(rows.col, rank_of)
            tail = torch.arange(D, self.n_coef, dtype=torch.int64, device=self.device)  # the intercept stays last
            self._perm = torch.cat([by_freq.to(torch.int64), tail])
            self._rows_eval = ranked
        self.plan = capi.DeviceFeTilePlan(self._rows_eval, **self.plan_args)

    def _to_device_order(self, x_np):
        x = self.torch.from_numpy(np.ascontiguousarray(x_np, dtype=np.float64)).to(self.device)
        return x if self._perm is None else x[self._perm]

    def _to_caller_order(self, v):
        if self._perm is None:
            return v.clone()
        out = self.torch.empty_like(v)
        out[self._perm] = v
        return out

    def _partial(self):
        """This rank's [value | gradient] at self._x_dev -> self._fg_dev (enqueue only)."""
        capi.fe_loss_grad_device(self._rows_eval, self.opts, self._x_dev, fg=self._fg_dev, plan=self.plan)

    def _evaluate(self):
        """All-reduced fg at self._x_dev, left in self._fg_dev (enqueue only)."""
        torch = self.torch
        self.nfev += 1
        if self.plan is None:
            self._prepare()
        if self.profile:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
        self._partial()
        if self.profile:
            ev[1].record()
        if self.dist and self.world > 1:
            self.dist.all_reduce(self._fg_dev, group=self.group)  # value and gradient in one collective
        if self.profile:
            ev[2].record()
            return ev
        return None

    # ---- objective ------------------------------------------------------------------------------------
    def loss_grad(self, x):
        """All-reduced (f, g) at x (numpy, caller's feature order) -- the reference's
        _compute_loss_and_gradients (:394-404)."""
        torch = self.torch
        if self.plan is None:
            self._prepare()
        self._x_dev.copy_(self._to_device_order(x))
        self._evaluate()
        fg = self._fg_dev
        g = self._to_caller_order(fg[1:])
        torch.cuda.current_stream().synchronize()
        return float(fg[0].item()), g.cpu().numpy()

    # ---- solve ----------------------------------------------------------------------------------------
    def fit(self, x0=None, threshold=None):
        """-> (x, info).  x0: previous model (same length) or None for zeros (:606-623)."""
        x = np.zeros(self.n_coef) if x0 is None else np.array(x0, dtype=np.float64, copy=True)
        assert x.shape == (self.n_coef,)
        if self.solver == "host":
            x, info = self._fit_host(x)
        else:
            x, info = self._fit_device(x)
        if threshold is not None:
            x = np.where(np.abs(x) <= threshold, 0.0, x)
        return x, info

    def _fit_device(self, x):
        torch = self.torch
        if self.plan is None:
            self._prepare()
        self._x_dev.copy_(self._to_device_order(x))
        lb = capi.DeviceLbfgs(self._x_dev, self._fg_dev, self.opts)
        try:
            lb.reset()
            task = capi.DeviceLbfgs.NEED_FG
            while task != capi.DeviceLbfgs.DONE:
                ev = self._evaluate() if task == capi.DeviceLbfgs.NEED_FG else None
                lb.step()
                if ev is not None:
                    end = torch.cuda.Event(enable_timing=True)
                    end.record()
                info = lb.poll()
                task = info["task"]
                if ev is not None:
                    self.phase_ms.append((ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]),
                                          ev[2].elapsed_time(end)))
        finally:
            lb.close()
        out = self._to_caller_order(self._x_dev).cpu().numpy()
        return out, {k: info[k] for k in ("nit", "nfev", "status", "f")}

    def _fit_host(self, x):
        """The host state machine (gdmix_lbfgs_*): fg crosses PCIe every evaluation."""
        solver = capi.HostLbfgs(self.n_coef, self.opts)
        f, g = self.loss_grad(x)
        while solver.iterate(x, f, g) == capi.HostLbfgs.NEED_FG:
            f, g = self.loss_grad(x)
        info = solver.info()
        solver.close()
        return x, info

    def score(self, x):
        """-> (logits incl. offset, per-coordinate logits) as fp32 numpy arrays for this rank's rows."""
        torch = self.torch
        xd = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).to(self.device)
        logit, per = capi.fe_score_device(self.rows, self.opts, xd)
        torch.cuda.synchronize()
        return logit.cpu().numpy(), per.cpu().numpy()
