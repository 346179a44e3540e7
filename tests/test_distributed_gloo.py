"""World-size-2 `gloo` test of the N > 1 host logic (CPU): shard arithmetic, the packed (P*D + P) all-reduce, and the
finalisation with the GLOBAL pre-factor.  The per-rank evaluator is the oracle here (no GPU); on the GPU the same
host code wraps pacoh_meta_mll_fwd_bwd (engine.meta_log_prob_and_score)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pacoh_oracle as orc


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from meta_learning_pacoh_b200 import engine as eng
    torch.set_num_threads(1)
    train = orc.sinusoid_tasks(9, 6, seed=4)
    stats = orc.normalization_stats(train)
    tasks = [orc.prepare_task(x, y, stats, torch.float64) for x, y in train]
    lay = orc.Layout(1, mean_layers=(8,), kernel_layers=(8,))
    mu, sigma = orc.hyper_prior_params(lay, 0.5, 3.0, torch.float64)
    g = torch.Generator().manual_seed(11)
    theta = mu + sigma * torch.randn(3, lay.D, generator=g, dtype=torch.float64)
    idx = np.random.RandomState(12).choice(9, size=11)            # same stream on every rank, odd length, repeats
    lo, hi = eng.shard_bounds(len(idx), rank, world)
    th = theta.clone().requires_grad_(True)
    mll = torch.stack([orc.task_mll(th, lay, *tasks[i]) for i in idx[lo:hi]], -1).sum(-1)     # (P,)
    (dth,) = torch.autograd.grad(mll.sum(), th)
    packed = torch.cat([dth.reshape(-1), mll.detach()])           # dtheta || mll_sum: ONE buffer, ONE all-reduce
    dist.all_reduce(packed)
    P, D = theta.shape
    pre = eng.pre_factor([6] * len(idx))                          # global batch
    prior_grad = -(theta - mu) / sigma ** 2
    score = 0.01 * prior_grad + pre * packed[:P * D].view(P, D)
    logp = 0.01 * orc.hyper_prior_log_prob(theta, mu, sigma) + pre * packed[P * D:]
    torch.save({"score": score, "logp": logp, "bounds": (lo, hi)}, os.path.join(out_dir, "rank%d.pt" % rank))
    dist.destroy_process_group()


def test_task_sharded_allreduce_equals_full_batch(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    outs = [torch.load(os.path.join(tmp_path, "rank%d.pt" % r)) for r in range(world)]
    assert outs[0]["bounds"] == (0, 5) and outs[1]["bounds"] == (5, 11)
    assert torch.equal(outs[0]["score"], outs[1]["score"]) and torch.equal(outs[0]["logp"], outs[1]["logp"])
    train = orc.sinusoid_tasks(9, 6, seed=4)
    stats = orc.normalization_stats(train)
    tasks = [orc.prepare_task(x, y, stats, torch.float64) for x, y in train]
    lay = orc.Layout(1, mean_layers=(8,), kernel_layers=(8,))
    mu, sigma = orc.hyper_prior_params(lay, 0.5, 3.0, torch.float64)
    g = torch.Generator().manual_seed(11)
    theta = mu + sigma * torch.randn(3, lay.D, generator=g, dtype=torch.float64)
    idx = np.random.RandomState(12).choice(9, size=11)
    logp, score, _ = orc.meta_log_prob_and_grad(theta, lay, [tasks[i] for i in idx], 0.01, mu, sigma)
    assert torch.allclose(outs[0]["logp"], logp, rtol=1e-12, atol=1e-12)
    assert torch.allclose(outs[0]["score"], score, rtol=1e-10, atol=1e-12)


def test_shard_bounds_cover_batch_exactly():
    from meta_learning_pacoh_b200 import engine as eng
    for T in (1, 7, 20, 4096):
        for world in (1, 2, 3, 4, 8):
            cuts = [eng.shard_bounds(T, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == T
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in cuts) - min(h - l for l, h in cuts) <= 1
