"""Pins oracle/pacoh_oracle.py to the reference: demo.ipynb logged trajectory (MAP) and fixtures produced by the
live reference modules (tests/golden/make_golden.py).  CPU only."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import pacoh_oracle as orc

torch.set_num_threads(1)


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def _tasks(fx):
    return [(torch.from_numpy(fx["x"][i]), torch.from_numpy(fx["y"][i])) for i in range(fx["x"].shape[0])]


ARCH = {
    "svgd_cfg2.npz": dict(input_dim=1),
    "svgd_n20.npz": dict(input_dim=1),
    "svgd_arch.npz": dict(input_dim=2, mean_layers=(16,), kernel_layers=(8, 24, 16)),
    "const_se.npz": dict(input_dim=2, mean_kind="constant", covar_kind="SE"),
}


def test_layout_matches_reference(golden_dir):
    ref = json.load(open(os.path.join(golden_dir, "layout.json")))
    for key, kw in (("default_d1", ARCH["svgd_cfg2.npz"]), ("arch_d2", ARCH["svgd_arch.npz"]),
                    ("const_se_d2", ARCH["const_se.npz"])):
        lay = orc.Layout(**kw)
        assert list(lay.entries.keys()) == list(ref[key].keys())
        assert [b - a for a, b in lay.entries.values()] == [v[0] for v in ref[key].values()]
    lay = orc.Layout(1)
    assert lay.D == 2342 and lay.entries["kernel_nn.fc_2.weight"] == (1249, 2273)
    assert lay.entries["lengthscale_raw"] == (2339, 2341) and lay.entries["noise_raw"] == (2341, 2342)


@pytest.mark.parametrize("name", list(ARCH))
def test_meta_logprob_score_phi_match_reference(golden_dir, name):
    fx = _load(golden_dir, name)
    lay = orc.Layout(**ARCH[name])
    tasks = _tasks(fx)
    theta = torch.from_numpy(fx["particles"])
    mu, sigma = orc.hyper_prior_params(lay, 0.5, 3.0)
    np.testing.assert_allclose(orc.hyper_prior_log_prob(theta, mu, sigma).numpy(), fx["prior_logp"], rtol=2e-6)
    batch = [tasks[i] for i in fx["idx"]]
    logp, score, _ = orc.meta_log_prob_and_grad(theta, lay, batch, 0.01, mu, sigma)
    np.testing.assert_allclose(logp.numpy(), fx["logp"], rtol=2e-5, atol=1e-5)
    scale = np.abs(fx["score"]).max()
    np.testing.assert_allclose(score.numpy(), fx["score"], rtol=1e-4, atol=2e-5 * scale)
    bw = 0.7 if name == "svgd_arch.npz" else None
    phi, gamma = orc.svgd_phi(theta, torch.from_numpy(fx["score"]), bw)
    assert abs(gamma - float(fx["gamma"])) <= 1e-6 * abs(gamma)
    np.testing.assert_allclose(phi.numpy(), fx["phi"], rtol=1e-4, atol=1e-5 * np.abs(fx["phi"]).max())
    phi2, _ = orc.svgd_phi_autograd(theta, torch.from_numpy(fx["score"]), bw)
    np.testing.assert_allclose(phi2.numpy(), fx["phi"], rtol=1e-4, atol=1e-5 * np.abs(fx["phi"]).max())


def test_three_svgd_steps_match_reference(golden_dir):
    fx = _load(golden_dir, "svgd_cfg2.npz")
    lay = orc.Layout(1)
    s = orc.SVGDOracle(_tasks(fx), lay, torch.from_numpy(fx["particles"]), seed=30)
    for k in range(3):
        idx = s.step()
        assert np.array_equal(idx, fx["steps_idx"][k])
    np.testing.assert_allclose(s.particles.numpy(), fx["particles_after3"], rtol=0, atol=2e-5)


def test_vi_matches_reference(golden_dir):
    fx = _load(golden_dir, "vi_cfg3.npz")
    lay = orc.Layout(1)
    mu, sigma = orc.hyper_prior_params(lay, 0.5, 3.0)
    loss, gl, gs, theta = orc.vi_neg_elbo_and_grad(torch.from_numpy(fx["loc"]), torch.from_numpy(fx["scale"]),
                                                   torch.from_numpy(fx["eps"]), lay, _tasks(fx), 0.01, mu, sigma)
    np.testing.assert_allclose(theta.numpy(), fx["theta"], atol=1e-6)
    assert abs(loss.item() - float(fx["loss"])) <= 2e-5 * abs(float(fx["loss"]))
    np.testing.assert_allclose(gl.numpy(), fx["dloc"], rtol=1e-4, atol=2e-5 * np.abs(fx["dloc"]).max())
    np.testing.assert_allclose(gs.numpy(), fx["dscale"], rtol=1e-4, atol=2e-5 * np.abs(fx["dscale"]).max())


def test_sinusoid_generator_matches_reference(golden_dir):
    fx = _load(golden_dir, "sinusoid.npz")
    train, test = orc.sinusoid_tasks(20, 5, seed=26, n_test=50)
    assert np.array_equal(np.stack([x for x, _ in train]), fx["train_x"])
    assert np.array_equal(np.stack([y for _, y in train]), fx["train_y"])
    assert np.array_equal(np.stack([t[2] for t in test]), fx["test_xs"])
    assert np.array_equal(np.stack([t[3] for t in test]), fx["test_ys"])


def test_analytic_mll_gradient_fp64():
    """dL/dK = 1/2 (alpha alpha^T - K^-1), dL/dm = alpha (the formulas the CUDA backward uses)."""
    torch.manual_seed(0)
    n = 12
    A = torch.randn(n, n, dtype=torch.float64)
    K = (A @ A.T / n + 0.3 * torch.eye(n, dtype=torch.float64)).requires_grad_(True)
    m = torch.randn(n, dtype=torch.float64, requires_grad=True)
    y = torch.randn(n, dtype=torch.float64)
    L = torch.distributions.MultivariateNormal(m, covariance_matrix=K).log_prob(y)
    gK, gm = torch.autograd.grad(L, (K, m))
    Ki = torch.linalg.inv(K.detach())
    alpha = Ki @ (y - m.detach())
    np.testing.assert_allclose(0.5 * (gK + gK.T).numpy(), (0.5 * (torch.outer(alpha, alpha) - Ki)).numpy(), atol=1e-12)
    np.testing.assert_allclose(gm.numpy(), alpha.numpy(), atol=1e-12)


def test_demo_trajectory_map_anchor(golden_dir):
    """demo.ipynb cell 6: iteration-1 line exactly, iteration-1000 line to the logged precision."""
    traj = json.load(open(os.path.join(golden_dir, "demo_trajectory.json")))
    train, test = orc.sinusoid_tasks(20, 5, seed=26, n_test=50)
    m = orc.MAPOracle(train, weight_decay=0.2, seed=30)
    loss, idx = m.step()
    assert list(idx) == traj["first_task_indices"]
    first = traj["trajectory"][0]
    assert "%.6f" % loss == "%.6f" % first["loss"]
    ll, rmse, cal = m.eval_datasets(test)
    assert "%.3f" % ll == "%.3f" % first["valid_ll"] and "%.3f" % rmse == "%.3f" % first["valid_rmse"]
    assert "%.3f" % cal == "%.3f" % first["calib_err"]
    cum = 0.0
    for _ in range(2, 1001):
        l, _ = m.step()
        cum += l
    second = traj["trajectory"][1]
    assert abs(cum / 1000 - second["loss"]) < 5e-5
    ll, rmse, cal = m.eval_datasets(test)
    assert abs(ll - second["valid_ll"]) < 2e-3 and abs(rmse - second["valid_rmse"]) < 2e-3
    assert abs(cal - second["calib_err"]) < 2e-3


@pytest.mark.parametrize("tag,name", [("cfg2", "svgd_cfg2.npz"), ("n20", "svgd_n20.npz")])
@pytest.mark.parametrize("key,bw", [("median", None), ("bw", 0.7)])
def test_imq_stein_kernel_direction_matches_reference(golden_dir, tag, name, key, bw):
    """IMQSteinKernel (svgd.py:63-99) through SVGD.phi's tail (svgd.py:18-21), golden from the live reference
    (tests/golden/make_golden_imq.py): the autograd restatement is the reference's op sequence, the closed form (with the
    explicit gradient through the per-dimension median) is what the CUDA kernels implement."""
    fx = np.load(os.path.join(golden_dir, name))
    g = np.load(os.path.join(golden_dir, "svgd_imq.npz"))
    theta, score = torch.from_numpy(fx["particles"]), torch.from_numpy(fx["score"])
    ref_phi, ref_K = g["phi_%s_%s" % (key, tag)], g["K_%s_%s" % (key, tag)]
    phi_a, K_a = orc.svgd_phi_imq_autograd(theta, score, bw)
    np.testing.assert_allclose(phi_a.numpy(), ref_phi, rtol=1e-5, atol=1e-6 * np.abs(ref_phi).max())
    np.testing.assert_allclose(K_a.numpy(), ref_K, rtol=1e-6)
    phi_c, K_c = orc.svgd_phi_imq(theta, score, bw)
    np.testing.assert_allclose(phi_c.numpy(), ref_phi, rtol=1e-4, atol=1e-5 * np.abs(ref_phi).max())
    # closed form == autograd to rounding in fp64 (the derivative through torch.median included)
    a64, _ = orc.svgd_phi_imq_autograd(theta.double(), score.double(), bw)
    c64, _ = orc.svgd_phi_imq(theta.double(), score.double(), bw)
    assert (a64 - c64).abs().max().item() <= 1e-12 * a64.abs().max().item()

