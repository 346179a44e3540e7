"""Config #4 (64 x 4096 x 50) gradient of the summed MLL against the fp64 oracle for a few particles (the CPU oracle
loops over tasks: ~1 min).  Run on a GPU box: python tests/manual/fullsize_fp64_check.py [n_particles_checked]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import pacoh_oracle as orc
from meta_learning_pacoh_b200 import engine as eng
P, T, n = 64, 4096, 50
K = int(sys.argv[1]) if len(sys.argv) > 1 else 2
rs = np.random.RandomState(3)
x = rs.uniform(-5, 5, size=(T, n, 1)).astype(np.float32)
y = (np.sin(x[..., 0]) + 0.1 * rs.normal(size=(T, n))).astype(np.float32)
lay, arch = orc.Layout(1), eng.GPArch(1)
mu, sigma = orc.hyper_prior_params(lay, 0.5, 3.0)
theta = (mu + sigma * torch.randn(P, lay.D, generator=torch.Generator().manual_seed(30)))
e = eng.MetaMLLEngine(arch, x, y, "cuda:0")
idx = np.random.RandomState(31).choice(T, size=T).astype(np.int32)
_, packed, info = e.mll_fwd_bwd(theta.cuda(), torch.from_numpy(idx).cuda())
g = packed[:P * lay.D].view(P, lay.D).cpu().double()
msum = packed[P * lay.D:].cpu().double()
t0 = time.time()
th64 = theta[:K].double().requires_grad_(True)
tot = torch.zeros(K, dtype=torch.float64)
xs, ys = torch.from_numpy(x).double(), torch.from_numpy(y).double()
for t in idx:
    tot = tot + orc.task_mll(th64, lay, xs[t], ys[t])
tot.sum().backward()
g64 = th64.grad
print("oracle %.0f s; sum mll rel err %.2e; grad err / max|g| per particle:" % (time.time() - t0, ((msum[:K] - tot.detach()).abs() / tot.detach().abs()).max().item()),
      [(float((g[k] - g64[k]).abs().max() / g64[k].abs().max())) for k in range(K)])
