"""Initial values drawn from torch's global CPU generator in exactly the order the reference consumes it, so that a
given ``random_seed`` produces the same particles / posterior initialisation as the reference (host logic, no kernels).

RNG consumption of the reference before anything is sampled ("RNG parity"):
  RandomGPMeta.__init__ -> VectorizedGP.__init__ (random_gp.py:22-52) builds the mean net, then the kernel net; every
  LinearVectorized.__init__ (models.py:279-293) draws  normal(in*out)  then  uniform_(in*out)  for the weight and
  uniform_(out) for the bias.  Those tensors are discarded (the particles are re-sampled from the hyper-prior), but
  they advance the generator.
"""
import math

import torch


def _consume_linear_vectorized(in_dim, out_dim):
    w = torch.normal(0, 1, size=(in_dim * out_dim,))      # models.py:283
    w.uniform_(-1.0, 1.0)                                  # models.py:288 -> :389-394
    torch.zeros(out_dim).uniform_(-1.0, 1.0)               # models.py:289-293


def consume_vectorized_gp_init(arch):
    """Advance the global generator as VectorizedGP.__init__ does for this architecture."""
    def net(widths, out_dim):
        prev = arch.input_dim
        for w in widths:
            _consume_linear_vectorized(prev, w)
            prev = w
        _consume_linear_vectorized(prev, out_dim)

    if arch.mean_kind == "NN":
        net(arch.mean_layers, 1)
    if arch.covar_kind == "NN":
        net(arch.kernel_layers, arch.feature_dim)


def sample_params_from_prior(arch, num, weight_prior_std, bias_prior_std):
    """CatDist.sample((num,)) (models.py:149-184 over random_gp.py:125-151): one torch.normal call per parameter
    group, concatenated -- same draws as the reference for the same generator state."""
    mu, sigma = arch.hyper_prior(weight_prior_std, bias_prior_std)
    parts = []
    for _, (a, b) in arch.entries().items():
        parts.append(torch.normal(mu[a:b].expand(num, b - a), sigma[a:b].expand(num, b - a)))
    return torch.cat(parts, dim=-1)


def init_diag_posterior(D, init_std=0.1):
    """RandomGPPosterior.__init__ (random_gp.py:244-248): loc ~ N(0, init_std), scale (log-std) ~ N(log 0.1, init_std)."""
    loc = torch.normal(0.0, init_std, size=(D,))
    scale = torch.normal(math.log(0.1), init_std, size=(D,))
    return loc, scale


def init_full_posterior(D, init_std=0.1):
    """cov_type='full' (random_gp.py:244-251): loc as above, tril_cov = diag(U[0.05, 0.1])."""
    loc = torch.normal(0.0, init_std, size=(D,))
    tril = torch.diag(torch.ones(D).uniform_(0.05, 0.1))
    return loc, tril
