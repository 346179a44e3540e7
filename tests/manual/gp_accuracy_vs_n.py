"""Worst per-group gradient error of the GP path against the fp64 oracle as a function of n (and of the fp32 oracle, i.e.
the reference's own precision, for comparison).  Run on a GPU box: python tests/manual/gp_accuracy_vs_n.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import pacoh_oracle as orc
from meta_learning_pacoh_b200 import engine as eng
DEV = "cuda:0"
lay, arch = orc.Layout(1), eng.GPArch(1)
for n in (32, 48, 64, 65, 80, 97, 112, 128):
    worst, worst32 = 0.0, 0.0
    for seed in range(3):
        rs = np.random.RandomState(100 * n + seed)
        x = rs.uniform(-2, 2, size=(5, n, 1)).astype(np.float32)
        y = (np.sin(2 * x[..., 0]) + 0.1 * rs.normal(size=(5, n))).astype(np.float32)
        mu, sigma = orc.hyper_prior_params(lay, 0.5, 3.0)
        theta = (mu + sigma * torch.randn(3, lay.D, generator=torch.Generator().manual_seed(seed + n))).numpy()
        idx = [4, 0, 0, 3, 1, 4, 2]
        e = eng.MetaMLLEngine(arch, x, y, DEV)
        th = torch.from_numpy(theta).to(DEV)
        _, packed, info = e.mll_fwd_bwd(th, torch.tensor(idx, dtype=torch.int32, device=DEV))
        _, dth = eng.logprob_finalize(th, mu.to(DEV), sigma.to(DEV), 0.01, eng.pre_factor([n] * len(idx)), packed)
        out = {}
        for dt in (torch.float64, torch.float32):
            tasks = [(torch.from_numpy(x[i]).to(dt), torch.from_numpy(y[i]).to(dt)) for i in range(5)]
            m_, s_ = orc.hyper_prior_params(lay, 0.5, 3.0, dt)
            _, g, _ = orc.meta_log_prob_and_grad(torch.from_numpy(theta).to(dt), lay, [tasks[i] for i in idx], 0.01, m_, s_)
            out[dt] = g.double().numpy()
        g64, g32, got = out[torch.float64], out[torch.float32], dth.cpu().double().numpy()
        scale = np.abs(g64).max()
        for name, (a, b) in arch.entries().items():
            den = max(np.abs(g64[:, a:b]).max(), 1e-2 * scale)
            worst = max(worst, np.abs(got[:, a:b] - g64[:, a:b]).max() / den)
            worst32 = max(worst32, np.abs(g32[:, a:b] - g64[:, a:b]).max() / den)
    print("n = %3d: kernels vs fp64 %.2e   fp32 oracle vs fp64 %.2e   (info %d)" % (n, worst, worst32, int(info.abs().max())))
