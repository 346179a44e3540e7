"""Full-size (config #4) self-consistency of the gradient: sum over 4 task shards vs the whole batch, relative to max |g|.
Run on a GPU box: python tests/manual/linearity_check.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import pacoh_oracle as orc
from meta_learning_pacoh_b200 import engine as eng
P, T, n = 64, 4096, 50
rs = np.random.RandomState(3)
x = rs.uniform(-5, 5, size=(T, n, 1)).astype(np.float32)
y = (np.sin(x[..., 0]) + 0.1 * rs.normal(size=(T, n))).astype(np.float32)
lay, arch = orc.Layout(1), eng.GPArch(1)
mu, sigma = orc.hyper_prior_params(lay, 0.5, 3.0)
theta = (mu + sigma * torch.randn(P, lay.D, generator=torch.Generator().manual_seed(30))).cuda()
e = eng.MetaMLLEngine(arch, x, y, "cuda:0")
idx = torch.from_numpy(np.random.RandomState(31).choice(T, size=T).astype(np.int32)).cuda()
_, full, _ = e.mll_fwd_bwd(theta, idx)
acc = torch.zeros_like(full, dtype=torch.float64)
for s in range(4):
    _, pk, _ = e.mll_fwd_bwd(theta, idx[s * 1024:(s + 1) * 1024].contiguous())
    acc += pk.double()
print("waves", os.environ.get("PACOH_BWD_WAVES"), "linearity err / max|g| = %.3e" % ((acc - full.double()).abs().max().item() / full.abs().max().item()))
