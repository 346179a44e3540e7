"""Generates the committed golden fixtures in tests/golden/ (run once, in the build container).

    python tests/golden/make_golden.py

Everything here is produced by the LIVE, unmodified reference modules imported from /root/reference
through oracle/ref_shim.py (meta_learn/svgd.py, meta_learn/models.py, meta_learn/random_gp.py,
experiments/data_sim.py); only ``VectorizedGP.forward`` (the gpytorch call site, random_gp.py:54-89)
is replaced by the dense restatement -- see the shim's docstring.  The reference cannot travel to the
GPU box, so the vectors are committed as small .npz files.

Fixtures
  svgd_cfg2.npz   BASELINE config #2: Sinusoid 20 tasks x 5 samples, P=10 particles sampled by the reference's
                  own ``sample_params_from_prior`` at torch seed 30, task batch from RandomState(31).choice;
                  outputs of ``RandomGPMeta.log_prob`` + autograd score and ``SVGD.phi`` (fp32, reference code).
  svgd_n20.npz    same pipeline at P=8, 16 tasks x 20 samples, hidden (32,32).
  svgd_arch.npz   a non-default architecture: d=2 inputs, mean (16,), kernel (8, 24, 16), mean 'NN'/covar 'NN'.
  const_se.npz    mean_module='constant', covar_module='SE' (no MLP at all), d=2.
  vi_cfg3.npz     RandomGPPosterior(diag) rsample / log_prob / neg-ELBO gradients for 32 tasks x 20 samples, S=8.
  layout.json     parameter names/offsets of the reference's RandomGPMeta for the architectures above.
  sinusoid.npz    first tasks of SinusoidDataset(RandomState(26)) (pins oracle.sinusoid_tasks).
  demo_trajectory.json  the logged lines of demo.ipynb cells 6 and 8 (copied numbers, the MAP anchor).
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_shim  # noqa: E402

torch.set_num_threads(1)
ref = ref_shim.load_reference()


def _prep(tasks):
    """abstract.py:212-258 via numpy (the reference base class needs absl logging; arithmetic is three lines)."""
    X = np.concatenate([x for x, _ in tasks], 0)
    Y = np.concatenate([y for _, y in tasks], 0)
    xm, xs, ym, ys = X.mean(0), X.std(0) + 1e-8, Y.mean(0), Y.std(0) + 1e-8
    out = []
    for x, y in tasks:
        out.append((torch.from_numpy((x - xm[None]) / xs[None]).float(),
                    torch.from_numpy(((y - ym[None]) / ys[None]).flatten()).float()))
    return out


def svgd_fixture(name, tasks, P, seed, bandwidth=None, **gp_kwargs):
    torch.manual_seed(seed)
    d = tasks[0][0].shape[1]
    rgp = ref.random_gp.RandomGPMeta(size_in=d, prior_factor=0.01, weight_prior_std=0.5, bias_prior_std=3.0, **gp_kwargs)
    particles = rgp.sample_params_from_prior(shape=(P,))
    idx = np.random.RandomState(seed + 1).choice(len(tasks), size=len(tasks))
    tiled = []
    for i in idx:  # GPR_meta_svgd.py:190-197
        x, y = tasks[i]
        tiled.append((x.view((1,) + x.shape).repeat(P, 1, 1), y.view((1,) + y.shape).repeat(P, 1)))
    X = particles.detach().requires_grad_(True)
    logp = rgp.log_prob(X, tiled)
    score = torch.autograd.grad(logp.sum(), X)[0]
    prior_lp = rgp._log_prob_prior(particles)
    kernel = ref.svgd.RBF_Kernel(bandwidth=bandwidth)
    phi = ref.svgd.SVGD(rgp, kernel, optimizer=None).phi(particles, tiled)
    d2 = ref.svgd.norm_sq(particles, particles)
    bw = kernel._bandwidth(d2)
    # three reference SVGD steps with Adam (GPR_meta_svgd.py:100-104, svgd.py:25-28)
    part_run = particles.clone()
    opt = torch.optim.Adam([part_run], lr=1e-3)
    svgd = ref.svgd.SVGD(rgp, ref.svgd.RBF_Kernel(bandwidth=bandwidth), optimizer=opt)
    rs = np.random.RandomState(seed + 1)
    steps_idx = []
    for _ in range(3):
        bi = rs.choice(len(tasks), size=len(tasks))
        steps_idx.append(bi)
        tl = [(tasks[i][0].view((1,) + tasks[i][0].shape).repeat(P, 1, 1),
               tasks[i][1].view((1,) + tasks[i][1].shape).repeat(P, 1)) for i in bi]
        svgd.step(part_run, tl)
    np.savez_compressed(
        os.path.join(HERE, name),
        x=np.stack([x.numpy() for x, _ in tasks]), y=np.stack([y.numpy() for _, y in tasks]),
        idx=idx.astype(np.int32), particles=particles.numpy(), logp=logp.detach().numpy(),
        prior_logp=prior_lp.numpy(), score=score.numpy(), phi=phi.numpy(),
        gamma=np.float64(1.0 / (1e-8 + 2 * bw ** 2)),
        steps_idx=np.stack(steps_idx).astype(np.int32), particles_after3=part_run.detach().numpy())
    return rgp


def main():
    layouts = {}
    ds = ref.data_sim.SinusoidDataset(random_state=np.random.RandomState(26))
    raw20x5 = ds.generate_meta_train_data(n_tasks=20, n_samples=5)
    raw_test = ds.generate_meta_test_data(n_tasks=20, n_samples_context=5, n_samples_test=50)
    np.savez_compressed(os.path.join(HERE, "sinusoid.npz"),
                        train_x=np.stack([x for x, _ in raw20x5]), train_y=np.stack([y for _, y in raw20x5]),
                        test_xc=np.stack([t[0] for t in raw_test]), test_yc=np.stack([t[1] for t in raw_test]),
                        test_xs=np.stack([t[2] for t in raw_test]), test_ys=np.stack([t[3] for t in raw_test]))

    kw = dict(covar_module_str='NN', mean_module_str='NN', mean_nn_layers=(32, 32), kernel_nn_layers=(32, 32))
    rgp = svgd_fixture("svgd_cfg2.npz", _prep(raw20x5), P=10, seed=30, **kw)
    layouts["default_d1"] = {k: list(v) for k, v in rgp.parameter_shapes().items()}

    ds = ref.data_sim.SinusoidDataset(random_state=np.random.RandomState(27))
    svgd_fixture("svgd_n20.npz", _prep(ds.generate_meta_train_data(n_tasks=16, n_samples=20)), P=8, seed=31,
                 bandwidth=None, **kw)

    rs = np.random.RandomState(5)
    tasks2d = []
    for _ in range(6):
        x = rs.uniform(-2, 2, size=(12, 2))
        y = np.sin(x[:, :1] * 2) + 0.3 * x[:, 1:] ** 2 + 0.05 * rs.normal(size=(12, 1))
        tasks2d.append((x, y))
    rgp = svgd_fixture("svgd_arch.npz", _prep(tasks2d), P=5, seed=32, bandwidth=0.7, covar_module_str='NN',
                       mean_module_str='NN', mean_nn_layers=(16,), kernel_nn_layers=(8, 24, 16))
    layouts["arch_d2"] = {k: list(v) for k, v in rgp.parameter_shapes().items()}
    rgp = svgd_fixture("const_se.npz", _prep(tasks2d), P=7, seed=33, covar_module_str='SE', mean_module_str='constant')
    layouts["const_se_d2"] = {k: list(v) for k, v in rgp.parameter_shapes().items()}
    with open(os.path.join(HERE, "layout.json"), "w") as f:
        json.dump(layouts, f, indent=1)

    # ---- VI (GPR_meta_vi.py:216-224, random_gp.py:224-263) ----
    torch.manual_seed(34)
    ds = ref.data_sim.SinusoidDataset(random_state=np.random.RandomState(28))
    tasks = _prep(ds.generate_meta_train_data(n_tasks=32, n_samples=20))
    rgp = ref.random_gp.RandomGPMeta(size_in=1, prior_factor=0.01, weight_prior_std=0.5, bias_prior_std=3.0, **kw)
    post = ref.random_gp.RandomGPPosterior(rgp.parameter_shapes(), cov_type='diag')
    S = 8
    gen_state = torch.get_rng_state()
    theta = post.rsample(sample_shape=(S,))
    # recover eps exactly as the reparameterisation drew it (same generator state -> same normal draws)
    torch.set_rng_state(gen_state)
    eps = torch.distributions.Normal(torch.zeros_like(post.loc), torch.ones_like(post.loc)).sample((S,))
    assert torch.allclose(theta, post.loc + post.scale.exp() * eps, atol=1e-6)
    tiled = [(x.view((1,) + x.shape).repeat(S, 1, 1), y.view((1,) + y.shape).repeat(S, 1)) for x, y in tasks]
    elbo = rgp.log_prob(theta, tiled) - 0.01 * post.log_prob(theta)
    loss = -torch.mean(elbo)
    gl, gs = torch.autograd.grad(loss, (post.loc, post.scale))
    np.savez_compressed(os.path.join(HERE, "vi_cfg3.npz"),
                        x=np.stack([x.numpy() for x, _ in tasks]), y=np.stack([y.numpy() for _, y in tasks]),
                        loc=post.loc.detach().numpy(), scale=post.scale.detach().numpy(), eps=eps.numpy(),
                        theta=theta.detach().numpy(), logq=post.log_prob(theta).detach().numpy(),
                        loss=loss.detach().numpy(), dloc=gl.numpy(), dscale=gs.numpy())

    # ---- demo.ipynb logged trajectory (numbers copied from the notebook's stored outputs) ----
    nb = json.load(open(os.path.join(ref_shim.REFERENCE_ROOT, "demo.ipynb")))
    lines = []
    for cell in (nb["cells"][6], nb["cells"][8]):
        for o in cell.get("outputs", []):
            lines += "".join(o.get("text", "")).strip().splitlines()
    traj, final = [], {}
    for ln in lines:
        if "Iter" in ln:
            p = ln.split("Iter ")[1]
            it = int(p.split("/")[0])
            g = lambda key: float(p.split(key)[1].split()[0])  # noqa: E731
            traj.append({"iter": it, "loss": g("Loss: "), "valid_ll": g("Valid-LL: "), "valid_rmse": g("Valid-RMSE: "),
                         "calib_err": g("Calib-Err ")})
        elif ":" in ln:
            k, v = ln.split(":")
            final[k.strip()] = float(v)
    with open(os.path.join(HERE, "demo_trajectory.json"), "w") as f:
        json.dump({"source": "demo.ipynb cells 6 and 8 (data RandomState(26), model seed 30, weight_decay=0.2)",
                   "first_task_indices": [18, 16, 2, 6, 10], "trajectory": traj, "final": final}, f, indent=1)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
