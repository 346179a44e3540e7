"""Generates tests/golden/svgd_imq.npz: the SVGD direction with the reference's IMQSteinKernel (meta_learn/svgd.py:63-99)
on the particles / scores of the committed svgd_cfg2.npz and svgd_n20.npz fixtures (not duplicated here).

    python tests/golden/make_golden_imq.py

Uses the LIVE reference module through oracle/ref_shim.py; phi is assembled exactly as SVGD.phi does it
(svgd.py:18-21: K_XX = K(X, X.detach()); grad_K = -autograd.grad(K_XX.sum(), X); phi = (K_XX.detach() @ score + grad_K) / P),
so the gradient flows through the per-dimension median bandwidth like in the reference."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_shim  # noqa: E402

torch.set_num_threads(1)
ref = ref_shim.load_reference()


def phi_reference(particles, score, bandwidth):
    X = torch.from_numpy(particles).clone().requires_grad_(True)
    kernel = ref.svgd.IMQSteinKernel(bandwidth=bandwidth)
    K = kernel(X, X.detach())
    grad_K = -torch.autograd.grad(K.sum(), X)[0]
    phi = (K.detach().matmul(torch.from_numpy(score)) + grad_K) / X.size(0)
    return phi.numpy(), K.detach().numpy()


out = {}
for tag, name in (("cfg2", "svgd_cfg2.npz"), ("n20", "svgd_n20.npz")):
    fx = np.load(os.path.join(HERE, name))
    out["phi_median_" + tag], out["K_median_" + tag] = phi_reference(fx["particles"], fx["score"], None)
    out["phi_bw_" + tag], out["K_bw_" + tag] = phi_reference(fx["particles"], fx["score"], 0.7)
out["bandwidth"] = np.float32(0.7)
np.savez_compressed(os.path.join(HERE, "svgd_imq.npz"), **out)
print({k: v.shape for k, v in out.items()})
