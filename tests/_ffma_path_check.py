"""Helper run in a subprocess by test_gpu_parity.py with the kernel-selection environment variables set (they are read
once per process): `ffma` -> PACOH_MLP_FWD=ffma PACOH_MLP_BWD=ffma, the CUDA-core MLP kernels of mlp.cu; `gpwarp` ->
PACOH_GP=warp, the register (warp-per-matrix) GP kernel for 32 < n <= 64.  Checks them against the fp64 oracle on a few
shapes.  Exit code 0 = parity holds."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pacoh_oracle as orc  # noqa: E402
from meta_learning_pacoh_b200 import engine as eng  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "ffma"
if mode == "ffma":
    assert os.environ.get("PACOH_MLP_FWD") == "ffma" and os.environ.get("PACOH_MLP_BWD") == "ffma"
else:
    assert os.environ.get("PACOH_GP") == "warp"
dev = "cuda:0"
worst = 0.0
for kw, n in ((dict(input_dim=1), 50), (dict(input_dim=2, mean_layers=(16,), kernel_layers=(8, 24, 16)), 13),
              (dict(input_dim=1, mean_layers=(32,) * 4, kernel_layers=(32,) * 4), 20),
              (dict(input_dim=3, mean_layers=(32, 32), kernel_layers=(32, 32), feature_dim=4), 41)):
    rs = np.random.RandomState(n)
    x = rs.uniform(-2, 2, size=(6, n, kw["input_dim"])).astype(np.float32)
    y = (np.sin(2 * x[..., 0]) + 0.1 * rs.normal(size=(6, n))).astype(np.float32)
    lay, arch = orc.Layout(**kw), eng.GPArch(**kw)
    mu, sigma = orc.hyper_prior_params(lay, 0.5, 3.0)
    theta = (mu + sigma * torch.randn(4, lay.D, generator=torch.Generator().manual_seed(n))).numpy()
    idx = [5, 1, 1, 0, 2, 3, 4]
    tasks = [(torch.from_numpy(x[i]).double(), torch.from_numpy(y[i]).double()) for i in range(6)]
    mu64, s64 = orc.hyper_prior_params(lay, 0.5, 3.0, torch.float64)
    logp64, g64, _ = orc.meta_log_prob_and_grad(torch.from_numpy(theta).double(), lay, [tasks[i] for i in idx], 0.01, mu64, s64)
    e = eng.MetaMLLEngine(arch, x, y, dev)
    th = torch.from_numpy(theta).to(dev)
    _, packed, info = e.mll_fwd_bwd(th, torch.tensor(idx, dtype=torch.int32, device=dev))
    logp, dth = eng.logprob_finalize(th, mu.to(dev), sigma.to(dev), 0.01, eng.pre_factor([n] * len(idx)), packed)
    err_l = (logp.cpu().double() - logp64).abs().max().item() / logp64.abs().max().item()
    err_g = (dth.cpu().double() - g64).abs().max().item() / g64.abs().max().item()
    worst = max(worst, err_l, err_g)
    print(mode, "path", kw, "logp rel %.2e grad rel %.2e" % (err_l, err_g))
sys.exit(0 if worst <= 1e-4 else 1)
