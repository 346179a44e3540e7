"""CPU oracle for the PACOH meta-training hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (``meta_learning_pacoh_b200``)
imports this module; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may.

It is a plain-torch restatement (dtype generic: run it in float64 for ground truth, in
float32 to mimic the reference) of the reference's arithmetic for SURVEY.md section 8(a):

  a1  per-particle MLP                      meta_learn/models.py:295-317, 343-349
  a3  flat parameter layout                 meta_learn/models.py:266-277, 319-323, 351-383
                                            meta_learn/random_gp.py:33-51, 97-111
  a4  softplus transforms                   meta_learn/random_gp.py:69-73
  a5  SE Gram                               meta_learn/models.py:428-446
  a6  noise on the diagonal                 meta_learn/models.py:448-487
  a8  marginal log-likelihood / n           meta_learn/random_gp.py:83-85  (gpytorch ExactMarginalLogLikelihood)
  a9  factorised Gaussian hyper-prior       meta_learn/random_gp.py:118-157, models.py:159-181
  a10 meta log-density                      meta_learn/random_gp.py:206-222
  a11 SVGD phi, RBF kernel, median          meta_learn/svgd.py:12-59, 103-107
  a13 diagonal Gaussian VI posterior        meta_learn/random_gp.py:224-263, GPR_meta_vi.py:216-224
  a14 PACOH-MAP loop                        meta_learn/GPR_meta_mll.py:104-119, 207-264
  a15 eval-mode GP posterior                meta_learn/GPR_meta_svgd.py:203-212 (gpytorch ExactGP.eval)
  a16 eval metrics                          meta_learn/abstract.py:134-181, 260-272

The arithmetic that the reference delegates to gpytorch (unpinned, not installed here;
circumstantially 1.0.x) is restated from its published semantics: dense Cholesky always,
``mll = MVN(m, K + s2 I).log_prob(y) / n``.

PINNING.  MAP path: pinned numerically against the reference's own logged run in
demo.ipynb (cells 6 and 8) -- see tests/test_oracle_pinning.py and
tests/golden/demo_trajectory.json.  SVGD phi / hyper-prior / vectorised MLP / VI posterior:
pinned against the *live* reference modules (svgd.py, models.py, random_gp.py imported from
/root/reference through oracle/ref_shim.py) via the fixtures produced by
tests/golden/make_golden.py.  The reference itself has no test and no logged output for
``VectorizedGP.forward`` on the SVGD/VI path ("parity unpinned" by the reference); its MLL is
the same restated function that the demo.ipynb anchor pins (s=1, no noise floor, P-batched).
"""
import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

LOG_2PI = math.log(2.0 * math.pi)


# --------------------------------------------------------------------------------------
# a3: parameter layout
# --------------------------------------------------------------------------------------
class Layout:
    """Flat layout of one particle's parameter vector.

    Order (random_gp.py:33-51): mean module, covariance module, ``noise_raw``; inside an MLP
    (models.py:351-383) layers ``fc_1..fc_L, out``; inside a layer bias precedes weight
    (models.py:319-323); weight is the row-major ``(out, in)`` matrix (models.py:307).
    ``outputscale_raw`` (MAP only, GPR_meta_mll.py:218) is appended last when requested.
    """

    def __init__(self, input_dim, mean_kind="NN", covar_kind="NN", mean_layers=(32, 32),
                 kernel_layers=(32, 32), feature_dim=2, outputscale=False, noise_floor=0.0):
        assert mean_kind in ("NN", "constant", "zero") and covar_kind in ("NN", "SE")
        self.input_dim, self.mean_kind, self.covar_kind = int(input_dim), mean_kind, covar_kind
        self.mean_layers, self.kernel_layers = tuple(mean_layers), tuple(kernel_layers)
        self.feature_dim = int(feature_dim) if covar_kind == "NN" else int(input_dim)
        self.outputscale, self.noise_floor = bool(outputscale), float(noise_floor)
        self.entries = OrderedDict()
        off = 0

        def add(name, size):
            nonlocal off
            self.entries[name] = (off, off + size)
            off += size

        def add_mlp(prefix, sizes_hidden, out_dim):
            prev = self.input_dim
            for i, h in enumerate(sizes_hidden):
                add("%s.fc_%d.bias" % (prefix, i + 1), h)
                add("%s.fc_%d.weight" % (prefix, i + 1), h * prev)
                prev = h
            add("%s.out.bias" % prefix, out_dim)
            add("%s.out.weight" % prefix, out_dim * prev)

        if mean_kind == "NN":
            add_mlp("mean_nn", self.mean_layers, 1)
        elif mean_kind == "constant":
            add("constant_mean", 1)
        if covar_kind == "NN":
            add_mlp("kernel_nn", self.kernel_layers, self.feature_dim)
        add("lengthscale_raw", self.feature_dim)
        add("noise_raw", 1)
        if self.outputscale:
            add("outputscale_raw", 1)
        self.D = off

    def get(self, theta, name):
        a, b = self.entries[name]
        return theta[..., a:b]


def mlp_forward(theta, layout, prefix, x):
    """a1. theta (P, D); x (P, n, d) or (n, d) -> (P, n, out).  h <- tanh(h W^T + b)."""
    P = theta.shape[0]
    if x.ndim == 2:
        x = x.unsqueeze(0).expand(P, -1, -1)
    sizes = layout.mean_layers if prefix == "mean_nn" else layout.kernel_layers
    h, prev = x, layout.input_dim
    for i, width in enumerate(sizes):
        W = layout.get(theta, "%s.fc_%d.weight" % (prefix, i + 1)).reshape(P, width, prev)
        b = layout.get(theta, "%s.fc_%d.bias" % (prefix, i + 1))
        h = torch.tanh(torch.bmm(h, W.transpose(1, 2)) + b[:, None, :])
        prev = width
    b = layout.get(theta, "%s.out.bias" % prefix)
    W = layout.get(theta, "%s.out.weight" % prefix).reshape(P, b.shape[-1], prev)
    return torch.bmm(h, W.transpose(1, 2)) + b[:, None, :]


# --------------------------------------------------------------------------------------
# a5, a6, a8: Gram, noise, marginal log-likelihood
# --------------------------------------------------------------------------------------
def se_gram(z1, lengthscale, z2=None, outputscale=None):
    """a5. K_ab = s * exp(-1/2 sum_f ((z1_af - z2_bf) / l_f)^2).  z (P, n, F); lengthscale (P, F)."""
    ls = lengthscale.reshape(lengthscale.shape[0], 1, -1)
    u1 = z1 / ls
    u2 = u1 if z2 is None else z2 / ls
    d2 = ((u1[:, :, None, :] - u2[:, None, :, :]) ** 2).sum(-1)
    K = torch.exp(-0.5 * d2)
    if outputscale is not None:
        K = outputscale.reshape(-1, 1, 1) * K
    return K


def mvn_mll(mean, K, noise, y):
    """a8. [log N(y | mean, K + noise I)] / n with a dense Cholesky.  mean (P, n); K (P, n, n); noise (P,)."""
    P, n = mean.shape
    Kt = K + noise.reshape(P, 1, 1) * torch.eye(n, dtype=K.dtype)
    L = torch.linalg.cholesky(Kt)
    r = (y - mean).unsqueeze(-1)
    alpha = torch.cholesky_solve(r, L)
    quad = (r * alpha).sum((-2, -1))
    logdet = 2.0 * torch.log(torch.diagonal(L, dim1=-2, dim2=-1)).sum(-1)
    return (-0.5 * quad - 0.5 * logdet - 0.5 * n * LOG_2PI) / n


def gp_components(theta, layout, x):
    """mean (P, n), features (P, n, F), lengthscale (P, F), noise (P,), outputscale (P,) or None."""
    P = theta.shape[0]
    xb = x.unsqueeze(0).expand(P, -1, -1) if x.ndim == 2 else x
    if layout.mean_kind == "NN":
        mean = mlp_forward(theta, layout, "mean_nn", xb).squeeze(-1)
    elif layout.mean_kind == "constant":
        mean = layout.get(theta, "constant_mean").expand(P, xb.shape[1])
    else:
        mean = torch.zeros(P, xb.shape[1], dtype=theta.dtype)
    z = mlp_forward(theta, layout, "kernel_nn", xb) if layout.covar_kind == "NN" else xb
    ls = F.softplus(layout.get(theta, "lengthscale_raw"))
    noise = F.softplus(layout.get(theta, "noise_raw")).reshape(P) + layout.noise_floor
    osc = F.softplus(layout.get(theta, "outputscale_raw")).reshape(P) if layout.outputscale else None
    return mean, z, ls, noise, osc


def task_mll(theta, layout, x, y):
    """a4+a7+a8: one task, all particles.  x (n, d); y (n,) -> (P,)."""
    mean, z, ls, noise, osc = gp_components(theta, layout, x)
    return mvn_mll(mean, se_gram(z, ls, outputscale=osc), noise, y)


# --------------------------------------------------------------------------------------
# a9, a10: hyper-prior and meta log-density
# --------------------------------------------------------------------------------------
def hyper_prior_params(layout, weight_prior_std=0.5, bias_prior_std=3.0, dtype=torch.float32):
    """a9. Per-coordinate (mu, sigma) of the factorised Gaussian hyper-prior (random_gp.py:125-151)."""
    mu, sigma = torch.zeros(layout.D, dtype=dtype), torch.ones(layout.D, dtype=dtype)
    for name, (a, b) in layout.entries.items():
        if name == "noise_raw":
            mu[a:b] = -1.0
        elif "mean_nn" in name or "kernel_nn" in name:
            sigma[a:b] = weight_prior_std if name.endswith("weight") else bias_prior_std
    return mu, sigma


def hyper_prior_log_prob(theta, mu, sigma):
    return (-0.5 * ((theta - mu) / sigma) ** 2 - torch.log(sigma) - 0.5 * LOG_2PI).sum(-1)


def pre_factor(task_sizes):
    """random_gp.py:209-212: harmonic-mean n over (harmonic-mean n + number of tasks in the batch)."""
    sizes = np.asarray(task_sizes, dtype=np.float64)
    hm = 1.0 / np.mean(1.0 / sizes)
    return hm / (hm + len(sizes))


def meta_log_prob(theta, layout, tasks, prior_factor, mu, sigma, return_mll=False):
    """a10. tasks: list of (x (n_t, d), y (n_t,)); duplicates count twice.  Reference loop structure."""
    pre = pre_factor([x.shape[0] for x, _ in tasks])
    mlls = torch.stack([task_mll(theta, layout, x, y) for x, y in tasks], dim=-1)  # (P, T)
    logp = prior_factor * hyper_prior_log_prob(theta, mu, sigma) + pre * mlls.sum(-1)
    return (logp, mlls) if return_mll else logp


def meta_log_prob_and_grad(theta, layout, tasks, prior_factor, mu, sigma):
    """score function used by SVGD.phi (svgd.py:13-16): d sum_p logp_p / d theta."""
    th = theta.detach().clone().requires_grad_(True)
    logp, mlls = meta_log_prob(th, layout, tasks, prior_factor, mu, sigma, return_mll=True)
    (g,) = torch.autograd.grad(logp.sum(), th)
    return logp.detach(), g, mlls.detach()


# --------------------------------------------------------------------------------------
# a11: SVGD
# --------------------------------------------------------------------------------------
def norm_sq(X, Y):
    """svgd.py:103-107."""
    XX, XY, YY = X.matmul(X.t()), X.matmul(Y.t()), Y.matmul(Y.t())
    return -2 * XY + XX.diag().unsqueeze(1) + YY.diag().unsqueeze(0)


def rbf_gamma(d2, bandwidth=None):
    """svgd.py:44-57: median heuristic over ALL P*P entries (numpy median), gamma = 1/(1e-8 + 2 h)."""
    if bandwidth is None:
        a = d2.detach().cpu().numpy()
        h = np.median(a) / (2 * np.log(a.shape[0] + 1))
        bw = np.sqrt(h).item()
    else:
        bw = bandwidth
    return 1.0 / (1e-8 + 2 * bw ** 2)


def svgd_phi(theta, score, bandwidth=None):
    """a11 closed form: phi = (K s + 2 gamma (rowsum(K) * theta - K theta)) / P  (== svgd.py:18-21)."""
    d2 = norm_sq(theta, theta)
    gamma = rbf_gamma(d2, bandwidth)
    K = torch.exp(-gamma * d2)
    phi = (K.matmul(score) + 2.0 * gamma * (K.sum(1, keepdim=True) * theta - K.matmul(theta))) / theta.shape[0]
    return phi, gamma


def svgd_phi_autograd(theta, score, bandwidth=None):
    """svgd.py:18-21 verbatim structure (autograd through the kernel), used to check the closed form."""
    X = theta.detach().clone().requires_grad_(True)
    d2 = norm_sq(X, X.detach())
    gamma = rbf_gamma(d2, bandwidth)
    K = torch.exp(-gamma * d2)
    grad_K = -torch.autograd.grad(K.sum(), X)[0]
    return (K.detach().matmul(score) + grad_K) / X.shape[0], gamma


def imq_kernel_matrix(X, Y, alpha=0.5, beta=-0.5, bandwidth=None):
    """IMQSteinKernel.forward (svgd.py:63-99): K_ij = (alpha + sum_d (X_jd - Y_id)^2 / h_d)^beta with, for
    ``bandwidth=None``, h_d = median over the strict upper triangle {i < j} of (X_jd - Y_id)^2 (torch.median: the LOWER
    median) divided by log(P + 1).  Differentiable in X, including through the median."""
    ns = (X.unsqueeze(0) - Y.unsqueeze(1)) ** 2                       # [i, j, d]
    if bandwidth is None:
        P = ns.shape[0]
        idx = torch.arange(P)
        h = ns[idx > idx.unsqueeze(-1), ...].median(dim=0)[0] / math.log(P + 1)
    else:
        h = bandwidth
    return torch.exp(beta * torch.log(alpha + torch.sum(ns / h, dim=-1)))


def svgd_phi_imq_autograd(theta, score, bandwidth=None, alpha=0.5, beta=-0.5):
    """svgd.py:18-21 verbatim structure with the IMQ kernel (autograd through the kernel AND its median bandwidth)."""
    X = theta.detach().clone().requires_grad_(True)
    K = imq_kernel_matrix(X, X.detach(), alpha, beta, bandwidth)
    grad_K = -torch.autograd.grad(K.sum(), X)[0]
    return (K.detach().matmul(score) + grad_K) / X.shape[0], K.detach()


def svgd_phi_imq(theta, score, bandwidth=None, alpha=0.5, beta=-0.5):
    """a11, IMQ closed form (what the CUDA kernels compute).  With ns_ijd = (x_jd - x_id)^2, base_ij = alpha + sum_d ns_ijd / h_d,
    K = base^beta, A = beta base^(beta - 1):
        d sum(K) / d x_md = (2 / h_d) sum_i A_im (x_md - x_id)                                   [direct]
                          - [m == j*_d] (sum_ij A_ij ns_ijd / h_d^2) 2 (x_j*d - x_i*d) / log(P + 1)   [through the median]
    where (i*_d < j*_d) is the pair whose squared distance is the (lower) median of dimension d; only the FIRST kernel
    argument is differentiated (svgd.py:18: K(X, X.detach())), hence only x_j*.  phi = (K s - d sum(K)/dx) / P."""
    X = theta.detach()
    P, D = X.shape
    ns = (X.unsqueeze(0) - X.unsqueeze(1)) ** 2                       # [i, j, d]
    if bandwidth is None:
        iu, ju = torch.triu_indices(P, P, offset=1)
        vals = ns[iu, ju, :]                                          # (P(P-1)/2, D), row order = the reference's mask order
        med, arg = vals.median(dim=0)
        h = med / math.log(P + 1)
        i_star, j_star = iu[arg], ju[arg]
    else:
        h = torch.full((D,), float(bandwidth), dtype=X.dtype)
    base = alpha + (ns / h).sum(-1)
    K = base ** beta
    A = beta * base ** (beta - 1.0)
    # direct term: m is the SECOND index of ns (the differentiated argument)
    grad = (2.0 / h) * (A.sum(0).unsqueeze(1) * X - A.t().matmul(X))
    if bandwidth is None:
        c = torch.einsum("ij,ijd->d", A, ns) / h ** 2                 # -d sum(K) / d h_d
        dd = torch.arange(D)
        grad[j_star, dd] -= c * 2.0 * (X[j_star, dd] - X[i_star, dd]) / math.log(P + 1)
    return (K.matmul(score) - grad) / P, K


# --------------------------------------------------------------------------------------
# a13: VI
# --------------------------------------------------------------------------------------
def vi_neg_elbo_and_grad(loc, scale, eps, layout, tasks, prior_factor, mu, sigma):
    """GPR_meta_vi.py:216-224 with a diagonal posterior N(loc, exp(scale)^2) and caller-supplied eps (S, D).

    Returns loss, dloss/dloc, dloss/dscale, theta (S, D).
    """
    loc = loc.detach().clone().requires_grad_(True)
    scale = scale.detach().clone().requires_grad_(True)
    theta = loc + torch.exp(scale) * eps
    logq = (-0.5 * eps ** 2 - scale - 0.5 * LOG_2PI).sum(-1)
    elbo = meta_log_prob(theta, layout, tasks, prior_factor, mu, sigma) - prior_factor * logq
    loss = -elbo.mean()
    gl, gs = torch.autograd.grad(loss, (loc, scale))
    return loss.detach(), gl, gs, theta.detach()


# --------------------------------------------------------------------------------------
# a15, a16: posterior and metrics
# --------------------------------------------------------------------------------------
def gp_posterior(theta, layout, xc, yc, xs):
    """a15. Predictive mean (P, n*) and covariance (P, n*, n*) INCLUDING observation noise, normalised space."""
    mean_c, z_c, ls, noise, osc = gp_components(theta, layout, xc)
    mean_s, z_s, _, _, _ = gp_components(theta, layout, xs)
    P, nc = mean_c.shape
    Kcc = se_gram(z_c, ls, outputscale=osc) + noise.reshape(P, 1, 1) * torch.eye(nc, dtype=theta.dtype)
    Kcs = se_gram(z_c, ls, z_s, outputscale=osc)
    Kss = se_gram(z_s, ls, outputscale=osc)
    L = torch.linalg.cholesky(Kcc)
    alpha = torch.cholesky_solve((yc - mean_c).unsqueeze(-1), L)
    mu = mean_s + (Kcs.transpose(1, 2) @ alpha).squeeze(-1)
    V = torch.linalg.solve_triangular(L, Kcs, upper=False)
    cov = Kss - V.transpose(1, 2) @ V + noise.reshape(P, 1, 1) * torch.eye(mean_s.shape[1], dtype=theta.dtype)
    return mu, cov


def calib_error(cdf_vals):
    """abstract.py:260-272."""
    conf = torch.linspace(0.05, 0.95, 20, dtype=cdf_vals.dtype)
    emp = (cdf_vals[:, None] <= conf).sum(0).to(cdf_vals.dtype) / cdf_vals.shape[0]
    return torch.sqrt(torch.mean((emp - conf) ** 2))


def eval_metrics(mu, cov, y_test, y_mean, y_std):
    """abstract.py:134-163 for the predictive mixture over P members (P may be 1).

    mu (P, n*), cov (P, n*, n*) in normalised space; y_test (n*,) in original units.
    Returns (avg joint log-likelihood / n*, rmse, calibration error).
    """
    P, ns = mu.shape
    yn = (y_test - y_mean) / y_std
    mvn = torch.distributions.MultivariateNormal(mu, covariance_matrix=cov)
    lp = mvn.log_prob(yn) - ns * math.log(y_std)                 # affine Jacobian, models.py:15-31
    ll = (torch.logsumexp(lp, 0) - math.log(P)) / ns             # models.py:121-126
    m_orig = mu * y_std + y_mean
    s_orig = torch.sqrt(torch.diagonal(cov, dim1=-2, dim2=-1)) * y_std
    rmse = torch.sqrt(torch.mean((m_orig.mean(0) - y_test) ** 2))
    cdf = torch.distributions.Normal(m_orig, s_orig).cdf(y_test).mean(0)
    return ll.item(), rmse.item(), calib_error(cdf).item()


def mixture_mean_std(mu, cov, y_mean, y_std):
    """models.py:90-115 after the affine un-normalisation (models.py:33-43)."""
    m = mu * y_std + y_mean
    var = torch.diagonal(cov, dim1=-2, dim2=-1) * y_std ** 2
    return m.mean(0), torch.sqrt(((m - m.mean(0)) ** 2).mean(0) + var.mean(0))


# --------------------------------------------------------------------------------------
# data preparation (a17) and the reference's loop structures
# --------------------------------------------------------------------------------------
def normalization_stats(meta_train_data):
    """abstract.py:212-222."""
    X = np.concatenate([np.asarray(x).reshape(len(x), -1) for x, _ in meta_train_data], 0)
    Y = np.concatenate([np.asarray(y).reshape(len(y), -1) for _, y in meta_train_data], 0)
    return X.mean(0), X.std(0) + 1e-8, Y.mean(0), Y.std(0) + 1e-8


def prepare_task(x, y, stats, dtype=torch.float32):
    """abstract.py:224-258: normalise in float64 numpy, then cast."""
    xm, xs, ym, ys = stats
    x = np.asarray(x).reshape(len(x), -1)
    xn = torch.from_numpy((x - xm[None]) / xs[None]).to(dtype)
    if y is None:
        return xn
    y = np.asarray(y).reshape(len(y), -1)
    return xn, torch.from_numpy(((y - ym[None]) / ys[None]).flatten()).to(dtype)


class SVGDOracle:
    """GPR_meta_svgd.py:82-121 + svgd.py:25-28 with the reference's loop structure (Adam on the (P, D) tensor)."""

    def __init__(self, tasks, layout, particles, prior_factor=0.01, weight_prior_std=0.5, bias_prior_std=3.0,
                 lr=1e-3, bandwidth=None, task_batch_size=-1, seed=None):
        self.tasks, self.layout, self.prior_factor, self.bandwidth = tasks, layout, prior_factor, bandwidth
        self.mu, self.sigma = hyper_prior_params(layout, weight_prior_std, bias_prior_std, particles.dtype)
        self.particles = particles.clone()
        self.optimizer = torch.optim.Adam([self.particles], lr=lr)
        self.task_batch_size = len(tasks) if task_batch_size < 1 else min(task_batch_size, len(tasks))
        self.rds = np.random.RandomState(seed + 1) if seed is not None else np.random

    def step(self, idx=None):
        if idx is None:
            idx = self.rds.choice(len(self.tasks), size=self.task_batch_size)
        batch = [self.tasks[i] for i in idx]
        _, score, _ = meta_log_prob_and_grad(self.particles, self.layout, batch, self.prior_factor, self.mu, self.sigma)
        phi, _ = svgd_phi(self.particles.detach(), score, self.bandwidth)
        self.optimizer.zero_grad()
        self.particles.grad = -phi
        self.optimizer.step()
        return idx


class MAPOracle:
    """PACOH-MAP restatement (a14) used for the demo.ipynb anchor.

    Module construction order = RNG order (GPR_meta_mll.py:207-251): kernel net first, then mean net,
    both ``torch.nn.Linear`` with default init; raw hypers start at 0;
    noise = 1e-3 + softplus(raw_noise) (GPR_meta_mll.py:54-55); AdamW(lr, weight_decay) over every
    group (GPR_meta_mll.py:253-257).
    """

    def __init__(self, meta_train_data, weight_decay=0.0, lr=1e-3, feature_dim=2, mean_layers=(32, 32),
                 kernel_layers=(32, 32), task_batch_size=5, seed=None, lr_decay=1.0):
        import torch.nn as nn
        if seed is not None:
            torch.manual_seed(seed)
            self.rds = np.random.RandomState(seed + 1)
        else:
            self.rds = np.random
        self.stats = normalization_stats(meta_train_data)
        self.y_mean, self.y_std = float(self.stats[2][0]), float(self.stats[3][0])
        d = np.asarray(meta_train_data[0][0]).reshape(len(meta_train_data[0][0]), -1).shape[1]

        def mlp(sizes, out):
            layers, prev = [], d
            for h in sizes:
                layers += [nn.Linear(prev, h), nn.Tanh()]
                prev = h
            return nn.Sequential(*layers, nn.Linear(prev, out))

        self.kernel_nn = mlp(kernel_layers, feature_dim)
        self.mean_nn = mlp(mean_layers, 1)
        self.raw_lengthscale = nn.Parameter(torch.zeros(1, feature_dim))
        self.raw_outputscale = nn.Parameter(torch.zeros(()))
        self.raw_noise = nn.Parameter(torch.zeros(1))
        groups = [{"params": self.kernel_nn.parameters(), "lr": lr, "weight_decay": weight_decay},
                  {"params": self.mean_nn.parameters(), "lr": lr, "weight_decay": weight_decay},
                  {"params": [self.raw_lengthscale, self.raw_outputscale], "lr": lr},
                  {"params": [self.raw_noise], "lr": lr}]
        self.optimizer = torch.optim.AdamW(groups, lr=lr, weight_decay=weight_decay)
        self.scheduler = torch.optim.lr_scheduler.StepLR(self.optimizer, 1000, gamma=lr_decay) if lr_decay < 1.0 else None
        self.tasks = [prepare_task(x, y, self.stats) for x, y in meta_train_data]
        self.task_batch_size = task_batch_size

    def _hypers(self):
        return (F.softplus(self.raw_lengthscale).reshape(1, -1), F.softplus(self.raw_outputscale).reshape(1),
                1e-3 + F.softplus(self.raw_noise).reshape(1))

    def mll(self, x, y):
        ls, osc, noise = self._hypers()
        mean = self.mean_nn(x).squeeze(-1).unsqueeze(0)
        z = self.kernel_nn(x).unsqueeze(0)
        return mvn_mll(mean, se_gram(z, ls, outputscale=osc), noise, y)[0]

    def step(self):
        idx = self.rds.choice(len(self.tasks), size=self.task_batch_size)
        self.optimizer.zero_grad()
        loss = 0.0
        for i in idx:
            loss = loss - self.mll(*self.tasks[i])
        loss.backward()
        self.optimizer.step()
        if self.scheduler is not None:
            self.scheduler.step()
        return float(loss.item()), idx

    @torch.no_grad()
    def posterior(self, xc, yc, xs):
        ls, osc, noise = self._hypers()
        n = xc.shape[0]
        zc, zs = self.kernel_nn(xc).unsqueeze(0), self.kernel_nn(xs).unsqueeze(0)
        Kcc = se_gram(zc, ls, outputscale=osc) + noise * torch.eye(n)
        Kcs, Kss = se_gram(zc, ls, zs, outputscale=osc), se_gram(zs, ls, outputscale=osc)
        L = torch.linalg.cholesky(Kcc)
        alpha = torch.cholesky_solve((yc - self.mean_nn(xc).squeeze(-1)).reshape(1, n, 1), L)
        mu = self.mean_nn(xs).squeeze(-1).unsqueeze(0) + (Kcs.transpose(1, 2) @ alpha).squeeze(-1)
        V = torch.linalg.solve_triangular(L, Kcs, upper=False)
        cov = Kss - V.transpose(1, 2) @ V + noise * torch.eye(xs.shape[0])
        return mu, cov

    def eval_datasets(self, test_tuples):
        out = []
        for xc, yc, xs, ys in test_tuples:
            xcn, ycn = prepare_task(xc, yc, self.stats)
            xsn = prepare_task(xs, None, self.stats)
            mu, cov = self.posterior(xcn, ycn, xsn)
            yt = torch.from_numpy(np.asarray(ys).flatten()).float()
            out.append(eval_metrics(mu, cov, yt, self.y_mean, self.y_std))
        return tuple(float(np.mean(c)) for c in zip(*out))

    def flat_parameters(self, layout):
        """Pack the nn.Linear parameters into the engine's flat layout (a3 order, outputscale appended)."""
        theta = torch.zeros(layout.D)
        for prefix, net in (("mean_nn", self.mean_nn), ("kernel_nn", self.kernel_nn)):
            lin = [m for m in net if isinstance(m, torch.nn.Linear)]
            names = ["fc_%d" % (i + 1) for i in range(len(lin) - 1)] + ["out"]
            for nm, l in zip(names, lin):
                a, b = layout.entries["%s.%s.bias" % (prefix, nm)]
                theta[a:b] = l.bias.detach()
                a, b = layout.entries["%s.%s.weight" % (prefix, nm)]
                theta[a:b] = l.weight.detach().reshape(-1)
        a, b = layout.entries["lengthscale_raw"]; theta[a:b] = self.raw_lengthscale.detach().reshape(-1)
        a, b = layout.entries["noise_raw"]; theta[a:b] = self.raw_noise.detach()
        a, b = layout.entries["outputscale_raw"]; theta[a:b] = self.raw_outputscale.detach().reshape(1)
        return theta.unsqueeze(0)


def sinusoid_tasks(n_tasks, n_samples, seed=26, n_test=0):
    """Re-statement of experiments/data_sim.py:203-248 (SinusoidDataset defaults) so that bench.py / smoke()
    can build the BASELINE inputs on a box without /root/reference.  Checked index-for-index against the
    reference generator in tests/test_oracle_pinning.py.
    Returns train tuples [(x (n,1), y (n,1))] (and test 4-tuples when n_test > 0, drawn AFTER the train set
    from the same generator, as demo.py:14-18 does)."""
    rs = np.random.RandomState(seed)

    def sample_fn():
        amp = rs.uniform(0.7, 1.3)
        x_shift = rs.normal(loc=0.0, scale=0.1)
        y_shift = rs.normal(loc=5.0, scale=0.1)
        slope = rs.normal(loc=0.5, scale=0.2)
        period = rs.uniform(1.5, 1.5)
        return lambda x: slope * x + amp * np.sin(period * (x - x_shift)) + y_shift

    train = []
    for _ in range(n_tasks):
        f = sample_fn()
        X = rs.uniform(-5, 5, size=(n_samples, 1))
        Y = f(X) + 0.1 * rs.normal(size=f(X).shape)
        train.append((X, Y))
    if n_test <= 0:
        return train
    test = []
    for _ in range(n_tasks):
        f = sample_fn()
        X = rs.uniform(-5, 5, size=(n_samples + n_test, 1))
        Y = f(X) + 0.1 * rs.normal(size=f(X).shape)
        test.append((X[:n_samples], Y[:n_samples], X[n_samples:], Y[n_samples:]))
    return train, test
