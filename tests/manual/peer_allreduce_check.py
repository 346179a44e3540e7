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
# latency of the cross-rank sum + finalize alone, at the config #4 buffer size (64 x 2342 + 64 floats), CUDA events
P4 = 64
peer4 = eng.PeerAllReduce(dist.group.WORLD, P4, lay.D, dev)
theta4 = (mu.cpu() + sigma.cpu() * torch.randn(P4, lay.D, generator=torch.Generator().manual_seed(1))).to(dev)
packed4 = torch.randn(P4 * lay.D + P4, device=dev)
def run_nccl():
    buf = packed4.clone()
    dist.all_reduce(buf)
    return eng.logprob_finalize(theta4, mu, sigma, 0.01, 0.5, buf)
def run_peer():
    peer4.out_buffer().copy_(packed4)
    return peer4.finalize(theta4, mu, sigma, 0.01, 0.5)
for name, fn in (("nccl all_reduce + pacoh_logprob_finalize", run_nccl), ("pacoh_peer_allreduce_finalize", run_peer)):
    for _ in range(10):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 200 * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("%-44s %.1f us per step (incl. a %d-float device copy), world %d" % (name, t.item(), packed4.numel(), world))
if rank == 0:
    print("OK" if worst <= 1e-6 else "MISMATCH", worst)
dist.destroy_process_group()
