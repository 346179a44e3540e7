"""Validation of the peer-memory all-reduce fused into the finalize kernel (engine.PeerAllReduce) against the NCCL path.
Run under torchrun on >= 2 GPUs of one node:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/manual/peer_allreduce_check.py"""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import pacoh_oracle as orc
from meta_learning_pacoh_b200 import engine as eng

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
P, T, n = 16, 256, 50
rs = np.random.RandomState(3)
x = rs.uniform(-5, 5, size=(T, n, 1)).astype(np.float32)
y = (np.sin(x[..., 0]) + 0.1 * rs.normal(size=(T, n))).astype(np.float32)
lay, arch = orc.Layout(1), eng.GPArch(1)
mu, sigma = orc.hyper_prior_params(lay, 0.5, 3.0)
mu, sigma = mu.to(dev), sigma.to(dev)
e = eng.MetaMLLEngine(arch, x, y, dev)
peer = eng.PeerAllReduce(dist.group.WORLD, P, lay.D, dev)
worst = 0.0
for it in range(6):
    theta = (mu.cpu() + sigma.cpu() * torch.randn(P, lay.D, generator=torch.Generator().manual_seed(30 + it))).to(dev)
    idx = np.random.RandomState(31 + it).choice(T, size=T).astype(np.int32)
    lo, hi = eng.shard_bounds(T, rank, world)
    shard = torch.from_numpy(idx[lo:hi]).to(dev)
    pre = eng.pre_factor([n] * T)
    la, ga, _ = eng.meta_log_prob_and_score(theta, e, shard, mu, sigma, 0.01, pre, group=dist.group.WORLD)
    lb, gb, _ = eng.meta_log_prob_and_score(theta, e, shard, mu, sigma, 0.01, pre, group=dist.group.WORLD, peer=peer)
    err = max(((la - lb).abs().max() / la.abs().max()).item(), ((ga - gb).abs().max() / ga.abs().max()).item())
    # every rank must hold bitwise the same result (fixed rank-order sum)
    chk = gb.double().sum().reshape(1).clone()
    gathered = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(gathered, chk)
    same = all(torch.equal(g, gathered[0]) for g in gathered)
    worst = max(worst, err)
    if rank == 0:
        print("iter %d: peer vs nccl rel diff %.2e, identical across ranks: %s" % (it, err, same))
    assert same
# latency
torch.cuda.synchronize(); dist.barrier()
for name, pr in (("nccl", None), ("peer", peer)):
    t0 = time.perf_counter()
    for _ in range(50):
        eng.meta_log_prob_and_score(theta, e, shard, mu, sigma, 0.01, pre, group=dist.group.WORLD, peer=pr)
    torch.cuda.synchronize(); dist.barrier()
    if rank == 0:
        print("%s path: %.1f us per call (MLL kernels included)" % (name, (time.perf_counter() - t0) / 50 * 1e6))
if rank == 0:
    print("OK" if worst <= 1e-6 else "MISMATCH", worst)
dist.destroy_process_group()
