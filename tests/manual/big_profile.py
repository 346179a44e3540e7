"""One batched MLL forward+backward at BASELINE config #5 shapes (PACOH-MAP, P = 1), for ncu / timing runs.
    python tests/manual/big_profile.py n T [reps]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from meta_learning_pacoh_b200 import engine as eng  # noqa: E402

n, T = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = torch.device("cuda:0")
arch = eng.GPArch(1, outputscale=True, noise_floor=1e-3)
rs = np.random.RandomState(0)
x = rs.uniform(-2, 2, size=(T, n, 1)).astype(np.float32)
y = (np.sin(2 * x[..., 0]) + 0.1 * rs.normal(size=(T, n))).astype(np.float32)
g = torch.Generator().manual_seed(30)
theta = ((torch.rand(1, arch.D, generator=g) * 2 - 1) * 0.17)
for name, (a, b) in arch.entries().items():
    if name.endswith("_raw"):
        theta[:, a:b] = 0.0
e = eng.MetaMLLEngine(arch, x, y, dev)
th = theta.to(dev)
idx = torch.arange(T, dtype=torch.int32, device=dev)
for _ in range(reps):
    mll, packed, info = e.mll_fwd_bwd(th, idx)
torch.cuda.synchronize()
print("n=%d T=%d mll mean %.5f info max %d" % (n, T, float(mll.mean()), int(info.max())))
