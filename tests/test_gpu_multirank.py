"""Two ranks on two GPUs (NCCL) through the product's own ``CBFSSM(config, group=...)`` path: the particles of one
minibatch are split contiguously over the ranks (a sequence's particles end up on both), gradients and ELBO terms
are all-reduced once per step, prediction moments once per fetch -- and every rank ends with what one GPU
computes alone.  Skipped on a single-GPU box (the driver's 8-GPU tier and ``gpurun --gpus 2`` run it)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = {"register_m20": 20, "tensor_m100": 100}


def _config(M):
    class DS:
        dim_u, dim_y = 2, 2
    return {"ds": DS, "batch_size": 3, "shuffle": 1, "dim_x": 4, "ind_pnt_num": M, "samples": 10, "learning_rate": 0.01,
            "loss_factors": np.asarray([10.0, 0.3]), "k_factor": 1.0, "recog_len": 4, "zeta_pos": 2.0, "zeta_mean": 0.3,
            "zeta_var": 0.05, "var_x": np.full(4, 0.01), "var_y": np.full(4, 1.0), "gp_var": 0.5, "gp_len": 1.0}


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from cbf_ssm_b200.model import CBFSSM
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    g = np.random.default_rng(5)
    B, T, S = 3, 12, 10
    u, y = g.standard_normal((B, T, 2)), g.standard_normal((B, T, 2))
    draws = [(g.standard_normal((2, T, B, S)), g.standard_normal((2, T, B, S)), g.standard_normal((T - 1, B, S)))
             for _ in range(2)]
    for name, M in CASES.items():
        models = [("sharded", CBFSSM(_config(M), device=dev, group=dist.group.WORLD, seed=3))]
        if rank == 0:
            models.append(("single", CBFSSM(_config(M), device=dev, group=None, seed=3)))
        for tag, m in models:
            m.engine.flags = 12
            m.inject_draws(*draws[0])
            loss = m.evaluate_batch(u, y, ["train", "loss"], True)[1]
            grad = m.engine.grad.cpu().numpy().copy()
            theta = m.engine.theta.cpu().numpy().copy()
            m.inject_draws(*draws[1])
            pm, pv, im, iv, mse = m.evaluate_batch(u, y, ["pred_mean", "pred_var", "internal_mean", "internal_var", "mse"],
                                                   False)
            np.savez(os.path.join(out_dir, f"{name}_{tag}_{rank}.npz"), loss=loss, grad=grad, theta=theta, pm=pm, pv=pv,
                     im=im, iv=iv, mse=mse)
        n0, nl = models[0][1]._shard(B)
        assert (n0, nl) == (rank * 15, 15)          # sequence 1 (particles 10..19) is split between the ranks
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_step_and_prediction_equal_the_single_gpu_result(tmp_path):
    import torch.multiprocessing as mp
    port = 29700 + (os.getpid() % 1500)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for name in CASES:
        one = np.load(tmp_path / f"{name}_single_0.npz")
        for rank in (0, 1):
            sh = np.load(tmp_path / f"{name}_sharded_{rank}.npz")
            assert abs(float(sh["loss"]) - float(one["loss"])) <= 2e-5 * abs(float(one["loss"])), (name, rank)
            assert np.max(np.abs(sh["grad"] - one["grad"])) <= 2e-5 * np.max(np.abs(one["grad"])), (name, rank)
            assert np.max(np.abs(sh["theta"] - one["theta"])) <= 1e-5, (name, rank)       # same Adam update everywhere
            for k in ("pm", "pv", "im", "iv"):
                assert np.max(np.abs(sh[k] - one[k])) <= 2e-5 * np.max(np.abs(one[k])), (name, rank, k)
            assert float(sh["mse"]) == pytest.approx(float(one["mse"]), rel=1e-4)
        a, b = np.load(tmp_path / f"{name}_sharded_0.npz"), np.load(tmp_path / f"{name}_sharded_1.npz")
        assert np.array_equal(a["grad"], b["grad"]) and np.array_equal(a["theta"], b["theta"])   # ranks stay in lock-step
