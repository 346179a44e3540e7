"""GPU tests of the reference-facing API (``-m gpu``): the three meta-learners, driven exactly like the reference's
callers (demo.py:25-32, tests/test_GPR.py::TestGPR_mll_meta), checked against golden vectors from the live reference
and against the CPU oracle's loop-structured restatements."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import pacoh_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ml():
    from meta_learning_pacoh_b200 import meta_learn
    assert torch.cuda.is_available()
    return meta_learn


def _denorm_tasks(fx, n):
    """golden fixtures store normalised tasks; with normalize_data=False the learner sees them unchanged."""
    return [(fx["x"][i].astype(np.float64), fx["y"][i].reshape(-1, 1).astype(np.float64)) for i in range(n)]


# ---------------------------------------------------------------------------------------------- PACOH-SVGD
def test_svgd_three_steps_match_reference_run(ml, golden_dir):
    """BASELINE config #2 (10 particles, 20 tasks x 5): same seed => same initial particles, same batch index stream,
    and after three Adam-SVGD steps the same particles as the reference code (fixture particles_after3)."""
    fx = np.load(os.path.join(golden_dir, "svgd_cfg2.npz"))
    tasks = _denorm_tasks(fx, 20)
    m = ml.GPRegressionMetaLearnedSVGD(tasks, num_particles=10, normalize_data=False, random_seed=30)
    assert np.array_equal(m.particles.cpu().numpy(), fx["particles"])
    rs = np.random.RandomState(31)
    for k in range(3):
        idx = m._sample_task_indices()
        assert np.array_equal(idx, fx["steps_idx"][k]) and np.array_equal(idx, rs.choice(20, size=20))
        m.svgd_step(idx)
    got = m.particles.cpu().numpy()
    assert np.abs(got - fx["particles_after3"]).max() <= 5e-5
    assert np.abs(got - fx["particles"]).max() > 1e-3          # it did move


def test_svgd_meta_fit_predict_eval_like_reference_callers(ml):
    train, test = orc.sinusoid_tasks(20, 5, seed=26, n_test=50)
    m = ml.GPRegressionMetaLearnedSVGD(train, num_iter_fit=40, num_particles=10, random_seed=30)
    ll0, rmse0, cal0 = m.eval_datasets(test)
    m.meta_fit(valid_tuples=test[:3], verbose=False, log_period=20)
    assert m.fitted
    xc, yc, xs, ys = test[0]
    mean, std = m.predict(xc, yc, xs)
    assert mean.shape == (50,) and std.shape == (50,) and np.isfinite(mean).all() and (std > 0).all()
    ucb, lcb = m.confidence_intervals(xc, yc, xs.flatten(), confidence=0.9)
    assert (ucb.numpy() > lcb.numpy()).all()
    # predictive path vs the oracle's posterior on the same particles
    lay = orc.Layout(1)
    stats = (m.x_mean, m.x_std, m.y_mean, m.y_std)
    xcn, ycn = orc.prepare_task(xc, yc, stats, torch.float64)
    xsn = orc.prepare_task(xs, None, stats, torch.float64)
    mu64, cov64 = orc.gp_posterior(m.particles.cpu().double(), lay, xcn, ycn, xsn)
    ref_mean, ref_std = orc.mixture_mean_std(mu64, cov64, float(m.y_mean[0]), float(m.y_std[0]))
    assert np.abs(mean - ref_mean.numpy()).max() <= 1e-4 * np.abs(ref_mean.numpy()).max()
    assert np.abs(std - ref_std.numpy()).max() <= 1e-4 * np.abs(ref_std.numpy()).max()
    ll, rmse, cal = m.eval(xc, yc, xs, ys)
    ll64, rmse64, cal64 = orc.eval_metrics(mu64, cov64, torch.from_numpy(ys.flatten()), float(m.y_mean[0]), float(m.y_std[0]))
    assert abs(ll - ll64) <= 1e-4 * max(1.0, abs(ll64)) and abs(rmse - rmse64) <= 1e-4 * rmse64 and abs(cal - cal64) <= 0.021


def test_svgd_two_stage_direction_equals_single_call():
    """pacoh_svgd_kernel_matrix on a side stream + pacoh_svgd_phi_apply == pacoh_svgd_phi (the step overlaps the
    particle-only half with the MLL kernels); a prepare() for other particles must be ignored, not reused."""
    from meta_learning_pacoh_b200 import engine as eng
    g = torch.Generator().manual_seed(5)
    theta = torch.randn(16, 300, generator=g).cuda()
    score = torch.randn(16, 300, generator=g).cuda()
    ref = eng.SVGDDirection(16, 300, "cuda:0")(theta, score)
    sv = eng.SVGDDirection(16, 300, "cuda:0")
    sv.prepare(theta)
    assert torch.equal(sv(theta, score), ref)
    sv.prepare(theta * 2.0)                       # stale: different tensor
    assert torch.equal(sv(theta, score), ref)
    assert torch.equal(sv(theta, score), ref)     # and without any prepare


def test_svgd_imq_kernel_learner_steps(ml):
    """kernel='IMQ' (GPR_meta_svgd.py:176-177): the learner trains, and its update equals the oracle's closed-form IMQ
    direction applied to the engine's own score (the score path is checked elsewhere)."""
    from meta_learning_pacoh_b200 import engine as eng
    train, _ = orc.sinusoid_tasks(8, 5, seed=4, n_test=10)
    m = ml.GPRegressionMetaLearnedSVGD(train, num_particles=6, kernel='IMQ', random_seed=3, num_iter_fit=4, optimizer='SGD', lr=1e-2)
    before = m.particles.detach().clone()
    idx = m._sample_task_indices()
    pre = eng.pre_factor([m.engine.n] * len(idx))
    _, score, _ = eng.meta_log_prob_and_score(before, m.engine, torch.from_numpy(np.asarray(idx, dtype=np.int32)).cuda(),
                                              m._prior_mu, m._prior_sigma, m.prior_factor, pre)
    m.svgd_step(idx)
    phi, _ = orc.svgd_phi_imq(before.cpu().double(), score.cpu().double(), None)
    expect = before.cpu().double() + 1e-2 * phi                      # SGD: particles += lr * phi  (grad = -phi)
    got = m.particles.detach().cpu().double()
    assert (got - expect).abs().max().item() <= 1e-4 * (1e-2 * phi.abs().max().item()) + 1e-7
    m.meta_fit(verbose=False)
    assert torch.isfinite(m.particles).all()


def test_learners_accept_ragged_task_sets(ml):
    """meta_train_data with different numbers of points per task (what the reference's per-task loop takes, e.g. the
    real-world datasets of its experiments): all three learners train and predict; the SVGD logp of the first step equals
    the oracle's on the same normalised tasks (harmonic-mean pre-factor included)."""
    rs = np.random.RandomState(2)
    tasks = []
    for t in range(10):
        n_t = int(rs.randint(4, 41))
        x = rs.uniform(-5, 5, size=(n_t, 1))
        tasks.append((x, np.sin(x) * rs.uniform(0.7, 1.3) + 0.1 * rs.normal(size=(n_t, 1))))
    m = ml.GPRegressionMetaLearnedSVGD(tasks, num_particles=4, random_seed=5, num_iter_fit=3)
    assert m._ragged and m.engine.task_n is not None
    idx = m._sample_task_indices()
    theta0 = m.particles.detach().clone()
    logp = m.svgd_step(idx)
    norm_tasks = [(td["train_x"].cpu().double(), td["train_y"].cpu().double()) for td in m.task_dicts]
    lay = orc.Layout(1)
    mu64, s64 = orc.hyper_prior_params(lay, m.weight_prior_std, m.bias_prior_std, torch.float64)
    logp64 = orc.meta_log_prob(theta0.cpu().double(), lay, [norm_tasks[i] for i in idx], m.prior_factor, mu64, s64)
    assert (logp.cpu().double() - logp64).abs().max().item() <= 1e-4 * logp64.abs().max().item()
    m.meta_fit(verbose=False)
    x_c, y_c = tasks[0]
    mean, std = m.predict(x_c, y_c, np.linspace(-5, 5, 30))
    assert np.isfinite(mean).all() and np.isfinite(std).all()
    for cls, kw in ((ml.GPRegressionMetaLearnedVI, dict(svi_batch_size=4)), (ml.GPRegressionMetaLearned, dict(task_batch_size=4))):
        m2 = cls(tasks, random_seed=6, num_iter_fit=3, **kw)
        m2.meta_fit(verbose=False)
        mean, std = m2.predict(x_c, y_c, np.linspace(-5, 5, 30))
        assert np.isfinite(mean).all() and np.isfinite(std).all()


def test_svgd_seed_determinism(ml):
    """tests/test_GPR.py:173-187 style: two runs with the same seed are bit-identical."""
    train, test = orc.sinusoid_tasks(12, 8, seed=3, n_test=20)
    outs = []
    for _ in range(2):
        m = ml.GPRegressionMetaLearnedSVGD(train, num_iter_fit=15, num_particles=6, random_seed=22, task_batch_size=4)
        m.meta_fit(verbose=False)
        outs.append((m.particles.cpu().numpy().copy(), m.predict(*test[0][:3])))
    assert np.array_equal(outs[0][0], outs[1][0])
    assert np.array_equal(outs[0][1][0], outs[1][1][0]) and np.array_equal(outs[0][1][1], outs[1][1][1])


# ---------------------------------------------------------------------------------------------- PACOH-VI
def test_vi_neg_elbo_and_gradients_match_reference(ml, golden_dir):
    fx = np.load(os.path.join(golden_dir, "vi_cfg3.npz"))
    tasks = _denorm_tasks(fx, 32)
    m = ml.GPRegressionMetaLearnedVI(tasks, svi_batch_size=8, normalize_data=False, random_seed=34)
    assert np.array_equal(m.posterior.loc.detach().cpu().numpy(), fx["loc"])
    assert np.array_equal(m.posterior.scale.detach().cpu().numpy(), fx["scale"])
    loss = m.get_neg_elbo(np.arange(32))                       # eps from the global generator == reference's rsample
    assert abs(loss.item() - float(fx["loss"])) <= 1e-4 * abs(float(fx["loss"]))
    gl, gs = m.posterior.loc.grad.cpu().numpy(), m.posterior.scale.grad.cpu().numpy()
    assert np.abs(gl - fx["dloc"]).max() <= 1e-4 * np.abs(fx["dloc"]).max()
    assert np.abs(gs - fx["dscale"]).max() <= 1e-4 * np.abs(fx["dscale"]).max()


@pytest.mark.parametrize("cov_type", ["diag", "full"])
def test_vi_meta_fit_and_predict(ml, cov_type):
    train, test = orc.sinusoid_tasks(16, 10, seed=5, n_test=30)
    kw = dict(mean_nn_layers=(8,), kernel_nn_layers=(8,)) if cov_type == "full" else {}
    m = ml.GPRegressionMetaLearnedVI(train, num_iter_fit=30, svi_batch_size=4, cov_type=cov_type, random_seed=7, lr=1e-2, **kw)
    l0 = m.get_neg_elbo(np.arange(16)).item()
    last = m.meta_fit(valid_tuples=test[:2], verbose=False, log_period=15)
    assert np.isfinite(last) and np.isfinite(l0)
    for mode in ("Bayes", "MAP"):
        mean, std = m.predict(*test[0][:3], n_posterior_samples=12, mode=mode)
        assert mean.shape == (30,) and np.isfinite(mean).all() and (std > 0).all()
    ll, rmse, cal = m.eval_datasets(test[:3], n_posterior_samples=12)
    assert np.isfinite([ll, rmse, cal]).all()


def test_vi_full_covariance_gradient_matches_autograd_oracle(ml):
    """cov_type='full' goes through the MetaLogProb autograd.Function: check d loss / d loc against the oracle."""
    train = orc.sinusoid_tasks(6, 7, seed=9)
    m = ml.GPRegressionMetaLearnedVI(train, svi_batch_size=3, cov_type="full", random_seed=3, mean_nn_layers=(8,), kernel_nn_layers=(8,))
    D = m.arch.D
    eps = torch.randn(3, D, generator=torch.Generator().manual_seed(1))
    loss = m.get_neg_elbo(np.arange(6), eps=eps)
    lay = orc.Layout(1, mean_layers=(8,), kernel_layers=(8,))
    tasks = [(td["train_x"].cpu().double(), td["train_y"].cpu().double()) for td in m.task_dicts]
    mu, sigma = orc.hyper_prior_params(lay, 0.5, 3.0, torch.float64)
    loc = m.posterior.loc.detach().cpu().double().requires_grad_(True)
    tril = m.posterior.tril_cov.detach().cpu().double().requires_grad_(True)
    q = torch.distributions.MultivariateNormal(loc, scale_tril=torch.tril(tril))
    theta = loc + eps.double() @ torch.tril(tril).T
    ref = -(orc.meta_log_prob(theta, lay, tasks, 0.01, mu, sigma) - 0.01 * q.log_prob(theta)).mean()
    gl, gt = torch.autograd.grad(ref, (loc, tril))
    assert abs(loss.item() - ref.item()) <= 1e-4 * abs(ref.item())
    assert (m.posterior.loc.grad.cpu().double() - gl).abs().max() <= 1e-4 * gl.abs().max()
    assert (m.posterior.tril_cov.grad.cpu().double() - gt).abs().max() <= 2e-4 * gt.abs().max()


# ---------------------------------------------------------------------------------------------- PACOH-MAP
def test_map_reproduces_demo_trajectory_start(ml, golden_dir):
    """BASELINE config #1 / demo.py: first logged line of demo.ipynb exactly, then 60 iterations against the oracle loop."""
    traj = json.load(open(os.path.join(golden_dir, "demo_trajectory.json")))
    train, test = orc.sinusoid_tasks(20, 5, seed=26, n_test=50)
    m = ml.GPRegressionMetaLearned(train, weight_decay=0.2, num_iter_fit=12000, random_seed=30)
    ref = orc.MAPOracle(train, weight_decay=0.2, seed=30)
    ll, rmse, cal = m.eval_datasets(test)               # untrained model
    first = traj["trajectory"][0]
    losses, ref_losses = [], []
    for it in range(60):
        m.optimizer.zero_grad()
        idx = m.rds_numpy.choice(len(m.task_dicts), size=m.task_batch_size)
        loss = m._loss_and_grad(idx)
        m.optimizer.step()
        losses.append(loss.item())
        rl, ridx = ref.step()
        assert np.array_equal(idx, ridx)
        ref_losses.append(rl)
    assert "%.5f" % losses[0] == "%.5f" % first["loss"]
    assert np.abs(np.array(losses) - np.array(ref_losses)).max() <= 2e-4
    ll, rmse, cal = m.eval_datasets(test)
    rll, rrmse, rcal = ref.eval_datasets(test)
    assert abs(ll - rll) <= 2e-4 and abs(rmse - rrmse) <= 2e-4 and abs(cal - rcal) <= 0.011


def test_map_first_log_line_and_state_dict_roundtrip(ml, golden_dir):
    traj = json.load(open(os.path.join(golden_dir, "demo_trajectory.json")))["trajectory"][0]
    train, test = orc.sinusoid_tasks(20, 5, seed=26, n_test=50)
    m = ml.GPRegressionMetaLearned(train, weight_decay=0.2, num_iter_fit=1, random_seed=30)
    loss = m.meta_fit(valid_tuples=None, verbose=False)
    assert "%.5f" % loss == "%.5f" % traj["loss"]
    ll, rmse, cal = m.eval_datasets(test)                # after one step == the notebook's first validation line
    assert "%.3f" % ll == "%.3f" % traj["valid_ll"] and "%.3f" % rmse == "%.3f" % traj["valid_rmse"]
    assert "%.3f" % cal == "%.3f" % traj["calib_err"]
    # serialisation + continued-training determinism (tests/test_GPR.py:189-222)
    m.meta_fit(verbose=False, n_iter=5)
    sd = m.state_dict()
    assert "likelihood.noise_covar.raw_noise" in sd["model"] and "learned_kernel.fc_1.weight" in sd["model"]
    m2 = ml.GPRegressionMetaLearned(train, weight_decay=0.2, num_iter_fit=1, random_seed=25)
    m2.load_state_dict(sd)
    m.rds_numpy, m2.rds_numpy = np.random.RandomState(5), np.random.RandomState(5)
    m.meta_fit(verbose=False, n_iter=5)
    m2.meta_fit(verbose=False, n_iter=5)
    p1, p2 = m.predict(*test[0][:3]), m2.predict(*test[0][:3])
    assert np.array_equal(p1[0], p2[0]) and np.array_equal(p1[1], p2[1])


@pytest.mark.parametrize("kw", [dict(mean_module="constant", covar_module="SE"), dict(mean_module="zero", covar_module="NN", feature_dim=3),
                                dict(mean_nn_layers=(128,) * 4, kernel_nn_layers=(128,) * 4)])
def test_map_module_variants_train(ml, kw):
    train, test = orc.sinusoid_tasks(10, 6, seed=2, n_test=15)
    mode = "learn_kernel" if kw.get("mean_module") == "zero" else "both"
    m = ml.GPRegressionMetaLearned(train, learning_mode=mode, num_iter_fit=20, random_seed=4, **kw)
    l0 = m._loss_and_grad(np.arange(10)).item()
    m.meta_fit(verbose=False)
    l1 = m._loss_and_grad(np.arange(10)).item()
    assert np.isfinite([l0, l1]).all() and l1 < l0
    assert np.isfinite(m.eval_datasets(test)).all()


# ---------------------------------------------------------------------------------------------- CUDA-graph replayed steps
@pytest.mark.parametrize("kind", ["svgd", "vi", "map"])
def test_graph_replayed_steps_equal_eager_steps(ml, monkeypatch, kind):
    """run_steps replays GRAPH_STEPS steps per CUDA graph (device-side step state, pre-uploaded index / normal-draw
    streams: SURVEY 8(f).2); the result must be BITWISE what the eager step sequence gives, lr decay included."""
    train = orc.sinusoid_tasks(20, 5, seed=26)

    def make():
        if kind == "svgd":
            return ml.GPRegressionMetaLearnedSVGD(train, num_particles=10, lr=2e-3, lr_decay=0.5, random_seed=30)
        if kind == "map":
            return ml.GPRegressionMetaLearned(train, lr_params=2e-3, weight_decay=0.1, lr_decay=0.5, learning_mode="learn_kernel",
                                              mean_module="constant", random_seed=30)
        return ml.GPRegressionMetaLearnedVI(train, svi_batch_size=8, lr=2e-3, lr_decay=0.5, random_seed=30)

    def params(m):
        if kind == "map":
            return m._flat.clone()
        return m.particles if kind == "svgd" else torch.cat([m.posterior.loc.detach(), m.posterior.scale.detach()])

    monkeypatch.setenv("PACOH_GRAPH", "0")
    a = make()
    a.run_steps(37)
    assert a._graph is None
    monkeypatch.setenv("PACOH_GRAPH", "1")
    b = make()
    b.run_steps(1)
    b.run_steps(36)                      # 1 eager, (1 more eager to reach an even count for svgd) then graphs + eager remainder
    assert b._graph is not None
    assert torch.equal(params(a), params(b))
    assert a._state.steps == b._state.steps == 37
    if kind == "map":
        b._sync_step_tensors()
        assert float((b._flat - a._flat).abs().max()) == 0.0 and float(b.constant_mean.detach()) == 0.0     # frozen by the mask
    st = b.optimizer.state[next(iter(b.optimizer.state))]
    assert float(st["step"]) == 37
    torch.manual_seed(7)                                       # VI draws its normals from the (shared) global CPU generator
    a.meta_fit(verbose=False, log_period=10, n_iter=12)        # meta_fit goes through the same path and stays in sync
    torch.manual_seed(7)
    b.meta_fit(verbose=False, log_period=10, n_iter=12)
    assert torch.equal(params(a), params(b))


# ---------------------------------------------------------------------------------------------- single-task sibling
def test_single_task_learner_like_reference_tests(ml):
    """GPRegressionLearned (meta_learn/GPR_mll.py:11-216): the behavioural checks of the reference's tests/test_GPR.py::TestGPR_mll
    -- seed determinism (:35-48), learning improves the fit (:76-100) -- plus the first loss against the oracle's MLL."""
    rs = np.random.RandomState(22)
    x = rs.uniform(-3, 3, size=(30, 1)); t = np.sin(2 * x) + 0.1 * rs.normal(size=(30, 1)) + 0.5 * x
    xv = rs.uniform(-3, 3, size=(40, 1)); tv = np.sin(2 * xv) + 0.1 * rs.normal(size=(40, 1)) + 0.5 * xv
    a = ml.GPRegressionLearned(x, t, num_iter_fit=60, random_seed=22)
    b = ml.GPRegressionLearned(x, t, num_iter_fit=60, random_seed=22)
    ll0, rmse0, _ = a.eval(xv, tv)
    la = a.fit(valid_x=xv, valid_t=tv, verbose=False, log_period=30)
    lb = b.fit(valid_x=xv, valid_t=tv, verbose=False, log_period=30)
    assert la == lb and np.array_equal(a.predict(xv)[0], b.predict(xv)[0])            # seed determinism
    ll1, rmse1, _ = a.eval(xv, tv)
    assert ll1 > ll0 and a.fitted
    mean, std = a.predict(xv.flatten())
    assert mean.shape == (40,) and (std > 0).all()
    ucb, lcb = a.confidence_intervals(xv.flatten())
    assert (ucb.numpy() > lcb.numpy()).all()
    # first-iteration loss = - mll of the (normalised) training set under the initial parameters, noise floor 1e-4
    c = ml.GPRegressionLearned(x, t, random_seed=5)
    lay = orc.Layout(1, outputscale=True, noise_floor=1e-4)
    theta = c._pack().cpu().double()
    xn, tn = c._normalize_data(x, t)
    want = -orc.task_mll(theta, lay, torch.from_numpy(xn), torch.from_numpy(tn.flatten()))
    got = c._loss_and_grad(np.zeros(1, dtype=np.int32))
    assert abs(float(got) - float(want)) <= 1e-5 * abs(float(want))


def test_map_fused_adamw_equals_torch_adamw(ml, monkeypatch):
    """The fused device-state AdamW (pacoh_adamw_step_dev) against torch.optim.AdamW on the same gradients, with weight decay
    and a StepLR schedule, stepped through the learner's torch path (zero_grad / _loss_and_grad / optimizer.step).  Three steps:
    Adam amplifies last-bit differences of tiny gradients step by step, so longer runs only agree statistically."""
    monkeypatch.setenv("PACOH_GRAPH", "0")
    train = orc.sinusoid_tasks(20, 5, seed=26)
    a = ml.GPRegressionMetaLearned(train, weight_decay=0.2, lr_decay=0.9, random_seed=30)
    b = ml.GPRegressionMetaLearned(train, weight_decay=0.2, lr_decay=0.9, random_seed=30)
    for _ in range(3):
        idx = a.rds_numpy.choice(20, size=5)
        assert np.array_equal(idx, b.rds_numpy.choice(20, size=5))
        a.map_step(idx)                                   # fused path
        b.optimizer.zero_grad(); b._loss_and_grad(idx); b.optimizer.step(); b.lr_scheduler.step()     # torch path
    assert a._state.steps == b._state.steps == 3
    # the kernel net's output bias has a numerically ZERO gradient (shift invariance): Adam normalises that rounding noise to
    # +-lr steps, so two correct implementations random-walk it differently (the reference's own path does too): leave it out
    lo, hi = a.arch.entries()["kernel_nn.out.bias"]
    keep = torch.ones(a.arch.D, dtype=torch.bool, device=a._flat.device)
    keep[lo:hi] = False
    # measured 1.5e-8 after two steps (3e-3 moved); the bound is far below what a wrong bias correction (~1e-3) or a missing
    # decoupled decay (3 steps x lr x wd x |p| ~ 6e-5) would leave, with room for Adam's amplification of last-bit gradient noise
    assert float((a._flat - b._flat)[0, keep].abs().max()) <= 1e-5
    assert float((a._mflat - b._mflat)[keep].abs().max()) <= 1e-6           # measured 6e-8
    assert float((a._vflat - b._vflat)[keep].abs().max()) <= 1e-6           # measured 2.5e-8
    moved = float((a._flat - ml.GPRegressionMetaLearned(train, weight_decay=0.2, random_seed=30)._flat)[0, keep].abs().max())
    assert moved > 1e-3                                                     # the optimizer really stepped
