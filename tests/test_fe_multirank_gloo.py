"""World-size-2 gloo test of the fixed-effect multi-rank plumbing on CPU: rows sharded [rank::world], one
all-reduce of [value | gradient] per evaluation, solver state replicated and identical on every rank
(fixed_effect_lr_lbfgs_model.py:382-390, :635-643 of the reference).  The per-rank evaluator injected here is the
CPU oracle (tests may use it); on the GPU box the same class runs the CUDA kernel (tests/test_fe_gpu.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gdmix_b200 import _capi as capi
    from gdmix_b200.fe_solver import FixedEffectSolver, shard_rows
    from oracle import oracle as O
    from tests.golden_util import load_fe
    arr, cases = load_fe()
    c = cases[0]
    k = c["key"]
    rowptr, col, val = arr[k + "_rowptr"], arr[k + "_col"], arr[k + "_val"]
    mine = shard_rows(c["n"], rank, world)
    # this rank's rows as its own CSR
    lens = np.diff(rowptr)[mine]
    rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    idx = np.concatenate([np.arange(rowptr[i], rowptr[i + 1]) for i in mine]).astype(np.int64)
    rows = O.FeBlock(len(mine), c["D"], rp, col[idx], val[idx], arr[k + "_y"][mine], arr[k + "_w"][mine],
                     arr[k + "_off"][mine], linear_regression=c["linear_regression"], num_workers=world)
    kw = dict(l2=c["l2"], regularize_bias=True, has_intercept=c["has_intercept"], m=c["m"], max_iter=c["max_iter"],
              factr=c["factr"])
    oo, po = O.make_opts(**kw), capi.make_opts(**kw)

    import torch

    class CpuEvalSolver(FixedEffectSolver):
        """Test seam: this rank's partial [value | gradient] comes from the CPU oracle instead of the CUDA passes, so
        that the multi-rank plumbing of the product solver (one all-reduce of fg per evaluation, replicated solver
        state) runs on gloo without a GPU.  Lives here, not in the package: the product has no CPU path."""

        def _setup(self):
            pass

        def loss_grad(self, x):
            self.nfev += 1
            f, g = O.fe_loss_grad(rows, oo, x)
            fg = torch.from_numpy(np.concatenate([[f], g]))
            if self.dist and self.world > 1:
                self.dist.all_reduce(fg, group=self.group)
            fg = fg.numpy()
            return float(fg[0]), fg[1:].copy()

    solver = CpuEvalSolver(None, po, n_features=c["D"], solver="host")
    x, info = solver.fit(arr[k + "_x0"])
    np.save(os.path.join(out_dir, f"x{rank}.npy"), x)
    np.save(os.path.join(out_dir, f"info{rank}.npy"), np.array([info["nit"], info["nfev"], info["status"]]))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_ranks_match_single_rank_and_golden(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    from tests.golden_util import load_fe
    arr, cases = load_fe()
    c = cases[0]
    x0, x1 = np.load(tmp_path / "x0.npy"), np.load(tmp_path / "x1.npy")
    i0, i1 = np.load(tmp_path / "info0.npy"), np.load(tmp_path / "info1.npy")
    np.testing.assert_array_equal(x0, x1)            # replicated state stays bit-identical
    np.testing.assert_array_equal(i0, i1)
    assert tuple(i0) == (c["nit"], c["nfev"], c["warnflag"])
    np.testing.assert_allclose(x0, arr[c["key"] + "_theta"], rtol=1e-8, atol=1e-10)


def test_shard_rule_is_the_references():
    """files[rank::world] (util/distribution_utils.py:46-47; test_distribution_utils.py:32-53)."""
    from gdmix_b200.fe_solver import shard_rows
    assert shard_rows(10, 0, 3) == [0, 3, 6, 9]
    assert shard_rows(10, 2, 3) == [2, 5, 8]
    assert sorted(sum((shard_rows(7, r, 4) for r in range(4)), [])) == list(range(7))
