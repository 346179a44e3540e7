"""GPU tests (``-m gpu``) of the predictive path: pacoh_gp_posterior / pacoh_pred_metrics (csrc/gp_post.cu) against the
oracle's eval-mode posterior (oracle.gp_posterior = gpytorch ExactGP.eval + likelihood restated, pinned by the
demo.ipynb validation lines) and its metrics (abstract.py:134-181).  Tolerance: 1e-4 relative on mean / std / LL / RMSE."""
import numpy as np
import pytest
import torch

from oracle import pacoh_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
RTOL = 1e-4


@pytest.fixture(scope="module")
def eng():
    from meta_learning_pacoh_b200 import engine
    assert torch.cuda.is_available()
    return engine


def _prior_particles(lay, P, seed):
    mu, sigma = orc.hyper_prior_params(lay, 0.5, 3.0)
    g = torch.Generator().manual_seed(seed)
    return mu + sigma * torch.randn(P, lay.D, generator=g)


def _tasks(Tt, nc_hi, ns_hi, d, seed, ragged=True):
    rs = np.random.RandomState(seed)
    out = []
    for t in range(Tt):
        nc = rs.randint(max(1, nc_hi // 3), nc_hi + 1) if ragged and t else nc_hi
        ns = rs.randint(max(1, ns_hi // 2), ns_hi + 1) if ragged and t else ns_hi
        xc, xs = rs.uniform(-2, 2, size=(nc, d)), rs.uniform(-2, 2, size=(ns, d))
        f = lambda x: np.sin(2 * x[:, 0]) + 0.3 * x.sum(-1)      # noqa: E731
        out.append((xc.astype(np.float32), (f(xc) + 0.1 * rs.normal(size=nc)).astype(np.float32),
                    xs.astype(np.float32), (f(xs) + 0.1 * rs.normal(size=ns)).astype(np.float32)))
    return out


@pytest.mark.parametrize("kw,P,Tt,nc,ns", [
    (dict(input_dim=1), 10, 6, 5, 50),                                                         # the demo / config #2 evaluation shape
    (dict(input_dim=1), 3, 4, 40, 20),                                                         # joint sets of 60 points: tensor-memory MLL kernel
    (dict(input_dim=1), 2, 3, 50, 500),                                                        # joint sets > 64 points: blocked Cholesky values-only
    (dict(input_dim=3, mean_layers=(32, 32), kernel_layers=(32, 32), feature_dim=4), 4, 5, 128, 130),   # full context tile, F = 4
    (dict(input_dim=2, mean_kind="constant", covar_kind="SE"), 3, 4, 17, 33),
    (dict(input_dim=1, outputscale=True, noise_floor=1e-3), 1, 7, 5, 50),                      # PACOH-MAP variant
    (dict(input_dim=2, mean_kind="zero", covar_kind="NN"), 2, 2, 1, 1),                        # one context point, one test point
])
def test_posterior_mean_var_cov_and_joint_ll_match_oracle(eng, kw, P, Tt, nc, ns):
    lay, arch = orc.Layout(**kw), eng.GPArch(**kw)
    theta = _prior_particles(lay, P, 40 + nc)
    tasks = _tasks(Tt, nc, ns, kw["input_dim"], seed=nc * 7 + ns)
    post = eng.gp_posterior_batch(arch, theta.to(DEV), [(t[0], t[1]) for t in tasks], [t[2] for t in tasks],
                                  targets=[t[3] for t in tasks], want_cov=True)
    assert int(post.info.min()) == 0 and int(post.info.max()) == 0
    for t, (xc, yc, xs, ys) in enumerate(tasks):
        mu64, cov64 = orc.gp_posterior(theta.double(), lay, torch.from_numpy(xc).double(), torch.from_numpy(yc).double(),
                                       torch.from_numpy(xs).double())
        n = len(xs)
        mu, var, cov = post.mu[:, t, :n].cpu().double(), post.var[:, t, :n].cpu().double(), post.cov[:, t, :n, :n].cpu().double()
        scale = max(mu64.abs().max().item(), 1.0)
        assert (mu - mu64).abs().max().item() <= RTOL * scale
        sd64 = torch.diagonal(cov64, dim1=-2, dim2=-1).sqrt()
        assert ((var.sqrt() - sd64).abs() / sd64).max().item() <= RTOL
        assert (cov - cov64).abs().max().item() <= RTOL * cov64.abs().max().item()
        lp64 = torch.distributions.MultivariateNormal(mu64, covariance_matrix=cov64).log_prob(torch.from_numpy(ys).double())
        assert (post.joint_ll[:, t].cpu().double() - lp64).abs().max().item() <= RTOL * max(lp64.abs().max().item(), float(n))
    # metrics kernel vs the oracle's (abstract.py:157-161)
    y_mean, y_std = 0.7, 1.9
    out = eng.pred_metrics(post, y_std).cpu().numpy()
    for t, (xc, yc, xs, ys) in enumerate(tasks):
        mu64, cov64 = orc.gp_posterior(theta.double(), lay, torch.from_numpy(xc).double(), torch.from_numpy(yc).double(),
                                       torch.from_numpy(xs).double())
        ll, rmse, cal = orc.eval_metrics(mu64, cov64, torch.from_numpy(ys).double() * y_std + y_mean, y_mean, y_std)
        assert abs(out[t, 0] - ll) <= RTOL * max(1.0, abs(ll)) and abs(out[t, 1] - rmse) <= RTOL * max(1.0, rmse)
        assert abs(out[t, 2] - cal) <= 1.0 / len(xs) + 1e-6        # empirical frequencies move in steps of 1 / n*


def test_predict_density_objects_behave_like_the_references(eng):
    from meta_learning_pacoh_b200.meta_learn import GPRegressionMetaLearned, GPRegressionMetaLearnedSVGD
    train, test = orc.sinusoid_tasks(20, 5, seed=26, n_test=50)
    xc, yc, xs, ys = test[0]
    lay = orc.Layout(1)
    for cls, kwargs, P in ((GPRegressionMetaLearnedSVGD, dict(num_particles=6), 6), (GPRegressionMetaLearned, dict(), 1)):
        m = cls(train, random_seed=30, **kwargs)
        d = m.predict(xc, yc, xs, return_density=True)
        mean, std = m.predict(xc, yc, xs)
        assert np.allclose(d.mean.numpy(), mean) and np.allclose(d.stddev.numpy(), std) and mean.shape == (50,)
        lp = d.log_prob(torch.from_numpy(ys.flatten()).float())
        assert lp.shape == () and np.isfinite(float(lp))
        ll, rmse, cal = m.eval(xc, yc, xs, ys)
        assert abs(float(lp) / 50 - ll) <= 1e-5 * max(1.0, abs(ll))                 # abstract.py:157: avg_log_likelihood = log_prob / n*
        assert abs(rmse - float(np.sqrt(np.mean((mean - ys.flatten()) ** 2)))) <= 1e-5
        base = d.dists.base_dist if P > 1 else d.base_dist
        cov = base.covariance_matrix
        assert cov.shape[-2:] == (50, 50) and torch.allclose(torch.diagonal(cov, dim1=-2, dim2=-1), base.variance, rtol=1e-5, atol=1e-7)
        ucb, lcb = m.confidence_intervals(xc, yc, xs.flatten(), confidence=0.9)
        assert (ucb.numpy() > mean).all() and (lcb.numpy() < mean).all()


def test_eval_datasets_is_one_batched_call_and_equals_per_task_eval(eng):
    from meta_learning_pacoh_b200.meta_learn import GPRegressionMetaLearnedSVGD
    train, test = orc.sinusoid_tasks(12, 8, seed=3, n_test=40)
    m = GPRegressionMetaLearnedSVGD(train, num_particles=5, random_seed=30)
    per_task = np.asarray([m.eval(*t) for t in test])
    ll, rmse, cal = m.eval_datasets(test)
    assert np.allclose([ll, rmse, cal], per_task.mean(0), rtol=1e-6, atol=1e-7)


def test_posterior_rejects_what_it_cannot_do(eng):
    from meta_learning_pacoh_b200._lib import PacohError
    lay, arch = orc.Layout(1), eng.GPArch(1)
    theta = _prior_particles(lay, 2, 1).to(DEV)
    big = _tasks(1, 129, 4, 1, seed=1, ragged=False)
    with pytest.raises(PacohError):
        eng.gp_posterior_batch(arch, theta, [(big[0][0], big[0][1])], [big[0][2]])
