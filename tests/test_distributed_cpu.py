"""World-size-2 checks on CPU (gloo) of the data-parallel contract: contiguous particle
shards, one all-reduce(sum) of [kernel-level gradient | ELBO terms], identical result to
the unsharded computation.  The arithmetic of each shard is done by the float64
restatement (oracle/kernel_math.py); the partition function is the product's."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import kernel_math as KM
from tests.helpers import make_problem


def _partition(N, world, rank):
    """The product's own partition function (cbf_ssm_b200.model.cbfssm.CBFSSM._shard) called on a stand-in
    object: the model class cannot be constructed without a GPU, the method itself is plain arithmetic."""
    from types import SimpleNamespace
    from cbf_ssm_b200.model.cbfssm import CBFSSM
    stub = SimpleNamespace(dims=SimpleNamespace(samples=1), world=world, rank=rank)
    return CBFSSM._shard(stub, N)


def _flat(kl):
    parts = []
    for tag in ("f", "b"):
        for nm in ("P", "alpha", "S", "Z", "ell"):
            parts.append(np.ravel(kl[tag][nm]))
        parts.append(np.asarray([kl[tag]["sig2"]]))
    parts += [kl["var_x"], kl["var_y"]]
    return np.concatenate(parts)


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg, params, u, y, eb, zb, ef = make_problem(4, 2, 2, 6, 5, 3, 9, 2, 1.0, (10.0, 0.5), seed=4, strong=True)
    pn = {k: v.numpy() for k, v in params.items()}
    B, S, T = 3, 5, 9
    N = B * S
    n0, nl = _partition(N, world, rank)
    # a shard is a set of whole particles; kernel_math is vectorised over particles, so feed it the
    # shard as B'=1 "sequence" per particle with S'=1 (u, y repeated per particle)
    un, yn = np.repeat(u, S, axis=0)[n0:n0 + nl], np.repeat(y, S, axis=0)[n0:n0 + nl]
    import copy
    c1 = copy.copy(cfg)
    c1.samples = 1
    c1.loss_factors = tuple(v / S for v in cfg.loss_factors)       # weights are l/S of the full problem
    ebs = eb.reshape(2, T, N)[:, :, n0:n0 + nl].reshape(2, T, nl, 1)
    zbs = zb.reshape(2, T, N)[:, :, n0:n0 + nl].reshape(2, T, nl, 1)
    efs = ef.reshape(T - 1, N)[:, n0:n0 + nl].reshape(T - 1, nl, 1)
    out, _ = KM.elbo_value_and_grad(c1, pn, un, yn, ebs, zbs, efs, True)
    vec = np.concatenate((_flat(out["kernel_level"]), [out["loglik"], out["kl_x"], out["entropy"]]))
    t = torch.tensor(vec)
    dist.all_reduce(t)                                             # the one collective of a step
    # every rank must shard the same minibatch: the product's load_ds broadcasts rank 0's shuffled order
    from types import SimpleNamespace
    from cbf_ssm_b200.model.base_model import BaseModel
    stub = SimpleNamespace(_group=dist.group.WORLD, world=world, engine=SimpleNamespace(device="cpu"))
    mine = np.random.RandomState(100 + rank).permutation(37)        # ranks disagree before the broadcast
    agreed = BaseModel._same_order_on_every_rank(stub, mine)
    if rank == 0:
        np.save(os.path.join(out_dir, "reduced.npy"), t.numpy())
    np.save(os.path.join(out_dir, f"order{rank}.npy"), agreed)
    dist.destroy_process_group()


def test_two_rank_shards_all_reduce_to_the_unsharded_result(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    red = np.load(tmp_path / "reduced.npy")
    cfg, params, u, y, eb, zb, ef = make_problem(4, 2, 2, 6, 5, 3, 9, 2, 1.0, (10.0, 0.5), seed=4, strong=True)
    out, _ = KM.elbo_value_and_grad(cfg, {k: v.numpy() for k, v in params.items()}, u, y, eb, zb, ef, True)
    ref = np.concatenate((_flat(out["kernel_level"]), [out["loglik"], out["kl_x"], out["entropy"]]))
    assert np.allclose(red, ref, rtol=1e-10, atol=1e-12)
    o0, o1 = np.load(tmp_path / "order0.npy"), np.load(tmp_path / "order1.npy")
    assert np.array_equal(o0, o1) and np.array_equal(o0, np.random.RandomState(100).permutation(37))


@pytest.mark.parametrize("N,world", [(1600, 8), (100, 8), (7, 4), (3, 8)])
def test_partition_covers_every_particle_once(N, world):
    seen = []
    for r in range(world):
        n0, nl = _partition(N, world, r)
        assert nl >= 0
        seen += list(range(n0, n0 + nl))
    assert seen == list(range(N))
