"""GPU parity tests proper (``-m gpu``): the CUDA path, called through the C ABI, against
  (a) the committed golden vectors produced by the live reference code (tests/golden/make_golden.py), and
  (b) the CPU oracle (fp64) on the same seeded inputs,
plus size-independent properties at BASELINE's full size (config #4: 64 particles x 4096 tasks x 50 points).

Tolerance (north_star): 1e-4 relative on MLL, gradients, in fp32.  Gradients are compared per parameter group in
the max-norm of the group (an entry whose true value is ~0, e.g. the kernel net's output bias, which the SE kernel
is invariant to, has no meaningful element-wise relative error).
"""
import os

import numpy as np
import pytest
import torch

from oracle import pacoh_oracle as orc

pytestmark = pytest.mark.gpu

RTOL = 1e-4


@pytest.fixture(scope="module")
def eng():
    from meta_learning_pacoh_b200 import engine
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return engine


DEV = "cuda:0"

ARCH = {
    "svgd_cfg2.npz": (dict(input_dim=1), dict(input_dim=1)),
    "svgd_n20.npz": (dict(input_dim=1), dict(input_dim=1)),
    "svgd_arch.npz": (dict(input_dim=2, mean_layers=(16,), kernel_layers=(8, 24, 16)),) * 2,
    "const_se.npz": (dict(input_dim=2, mean_kind="constant", covar_kind="SE"),) * 2,
}


def relmax(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def run_engine(eng, arch, x, y, theta, idx, prior_factor=0.01, wstd=0.5, bstd=3.0):
    e = eng.MetaMLLEngine(arch, x, y, DEV)
    th = torch.as_tensor(theta, dtype=torch.float32).contiguous().to(DEV)
    tidx = torch.as_tensor(np.asarray(idx), dtype=torch.int32).to(DEV)
    mll, packed, info = e.mll_fwd_bwd(th, tidx)
    mu, sigma = arch.hyper_prior(wstd, bstd)
    pre = eng.pre_factor([x.shape[1]] * len(idx))
    logp, dth = eng.logprob_finalize(th, mu.to(DEV), sigma.to(DEV), prior_factor, pre, packed)
    return mll.cpu().numpy(), logp.cpu().numpy(), dth.cpu().numpy(), info.cpu().numpy()


def oracle64(lay, x, y, theta, idx, prior_factor=0.01, wstd=0.5, bstd=3.0):
    tasks = [(torch.from_numpy(np.asarray(x[i])).double(), torch.from_numpy(np.asarray(y[i])).double()) for i in range(len(x))]
    mu, sigma = orc.hyper_prior_params(lay, wstd, bstd, torch.float64)
    logp, g, mll = orc.meta_log_prob_and_grad(torch.as_tensor(theta).double(), lay, [tasks[i] for i in idx], prior_factor, mu, sigma)
    return mll.numpy(), logp.numpy(), g.numpy()


# n > 64 runs the blocked Cholesky path (csrc/gp_big.cu) and holds the same 1e-4 bar as the small-matrix kernels (round 1
# tested 64 < n <= 128 at 5e-4 on an in-place Gauss-Jordan sweep; that kernel is now only reachable with PACOH_GP=tc128).
RTOL_N128 = RTOL
def assert_groups(arch, got, want, tol=RTOL, ref32=None):
    """Per parameter group, error in the group's max-norm, relative to max(|group|, 1e-3 * overall gradient scale).

    A group whose true gradient is numerically zero has no meaningful relative error of its own: the kernel net's output
    bias (the stationary SE kernel is invariant to a feature shift, so its likelihood gradient is an exact cancellation of
    O(scale) terms and only the tiny prior term survives).  What fp32 leaves there is summation residue -- the reference's
    own fp32 torch path included: on test_every_matrix_size_up_to_128[65] the fp32 oracle is off by 3.2e-6 (2.3e-4 of the
    floor) where this engine is off by 1.9e-6.  For such degenerate groups (below 1% of the scale) the bound is therefore
    the larger of the 1e-4 floor bound and 1.5 x the fp32 oracle's own error, when the caller supplies it (``ref32``)."""
    scale = np.abs(want).max()
    for name, (a, b) in arch.entries().items():
        ref = want[:, a:b]
        err = np.abs(got[:, a:b] - ref).max()
        gmax = np.abs(ref).max()
        bound = tol * max(gmax, 1e-3 * scale)
        if ref32 is not None and gmax < 1e-2 * scale:
            bound = max(bound, 1.5 * np.abs(ref32[:, a:b] - ref).max())
        assert err <= bound, (name, err, gmax, scale)


def oracle32(lay, x, y, theta, idx, prior_factor=0.01, wstd=0.5, bstd=3.0):
    """The same oracle in fp32: the precision the reference's own torch path runs in."""
    tasks = [(torch.from_numpy(np.asarray(x[i])).float(), torch.from_numpy(np.asarray(y[i])).float()) for i in range(len(x))]
    mu, sigma = orc.hyper_prior_params(lay, wstd, bstd, torch.float32)
    _, g, _ = orc.meta_log_prob_and_grad(torch.as_tensor(theta).float(), lay, [tasks[i] for i in idx], prior_factor, mu, sigma)
    return g.double().numpy()


# ------------------------------------------------------------------------------------------ golden vectors
@pytest.mark.parametrize("name", list(ARCH))
def test_logp_and_score_match_reference_golden(eng, golden_dir, name):
    fx = np.load(os.path.join(golden_dir, name))
    arch = eng.GPArch(**ARCH[name][0])
    mll, logp, score, info = run_engine(eng, arch, fx["x"], fx["y"], fx["particles"], fx["idx"])
    assert (info == 0).all()
    assert relmax(logp, fx["logp"]) <= RTOL
    assert_groups(arch, score, fx["score"])
    # and against the fp64 oracle, which is tighter than the reference's own fp32 run
    lay = orc.Layout(**ARCH[name][1])
    mll64, logp64, g64 = oracle64(lay, fx["x"], fx["y"], fx["particles"], fx["idx"])
    assert relmax(mll, mll64) <= RTOL and relmax(logp, logp64) <= RTOL
    assert_groups(arch, score, g64)


@pytest.mark.parametrize("name,bw", [("svgd_cfg2.npz", None), ("svgd_n20.npz", None), ("svgd_arch.npz", 0.7), ("const_se.npz", None)])
def test_svgd_phi_matches_reference_golden(eng, golden_dir, name, bw):
    fx = np.load(os.path.join(golden_dir, name))
    P, D = fx["particles"].shape
    sv = eng.SVGDDirection(P, D, DEV, bandwidth=bw)
    phi = sv(torch.from_numpy(fx["particles"]).to(DEV), torch.from_numpy(fx["score"]).to(DEV))
    assert abs(float(sv.gamma.item()) - float(fx["gamma"])) <= 1e-5 * float(fx["gamma"])
    assert relmax(phi.cpu().numpy(), fx["phi"]) <= RTOL


@pytest.mark.parametrize("tag,name", [("cfg2", "svgd_cfg2.npz"), ("n20", "svgd_n20.npz")])
@pytest.mark.parametrize("key,bw", [("median", None), ("bw", 0.7)])
def test_svgd_phi_imq_matches_reference_golden(eng, golden_dir, tag, name, key, bw):
    """kernel='IMQ' (IMQSteinKernel, svgd.py:63-99): per-dimension median bandwidth, gradient through the median."""
    fx = np.load(os.path.join(golden_dir, name))
    g = np.load(os.path.join(golden_dir, "svgd_imq.npz"))
    P, D = fx["particles"].shape
    sv = eng.SVGDDirection(P, D, DEV, bandwidth=bw, kernel="IMQ")
    theta, score = torch.from_numpy(fx["particles"]).to(DEV), torch.from_numpy(fx["score"]).to(DEV)
    phi = sv(theta, score)
    assert relmax(phi.cpu().numpy(), g["phi_%s_%s" % (key, tag)]) <= RTOL
    sv.prepare(theta)                                   # two-stage form gives the same
    assert torch.equal(sv(theta, score), phi)


def test_svgd_phi_imq_full_size_against_oracle(eng):
    """64 particles x 2342 parameters (config #4's particle matrix): 2016 pairs per dimension through the on-device sort."""
    g = torch.Generator().manual_seed(11)
    theta = torch.randn(64, 2342, generator=g) * 0.5
    score = torch.randn(64, 2342, generator=g)
    phi = eng.SVGDDirection(64, 2342, DEV, kernel="IMQ")(theta.to(DEV), score.to(DEV))
    ref, _ = orc.svgd_phi_imq(theta.double(), score.double(), None)
    assert relmax(phi.cpu().numpy(), ref.numpy()) <= RTOL


def test_vi_sample_and_gradient_match_reference_golden(eng, golden_dir):
    fx = np.load(os.path.join(golden_dir, "vi_cfg3.npz"))
    arch = eng.GPArch(1)
    loc, scale, eps = (torch.from_numpy(fx[k]).to(DEV) for k in ("loc", "scale", "eps"))
    theta, logq = eng.vi_sample(loc, scale, eps)
    assert np.abs(theta.cpu().numpy() - fx["theta"]).max() <= 1e-6
    assert relmax(logq.cpu().numpy(), fx["logq"]) <= 1e-5
    T = fx["x"].shape[0]
    e = eng.MetaMLLEngine(arch, fx["x"], fx["y"], DEV)
    mu, sigma = arch.hyper_prior(0.5, 3.0)
    tidx = torch.arange(T, dtype=torch.int32, device=DEV)
    _, packed, info = e.mll_fwd_bwd(theta, tidx)
    logp, g = eng.logprob_finalize(theta, mu.to(DEV), sigma.to(DEV), 0.01, eng.pre_factor([20] * T), packed)
    loss = -(logp - 0.01 * logq).mean()
    assert abs(loss.item() - float(fx["loss"])) <= RTOL * abs(float(fx["loss"]))
    dloc, dscale = eng.vi_grad(scale, eps, g, 0.01)
    assert relmax(dloc.cpu().numpy(), fx["dloc"]) <= RTOL
    assert relmax(dscale.cpu().numpy(), fx["dscale"]) <= RTOL


# ------------------------------------------------------------------------------------------ oracle, edge cases
def _synthetic(T, n, d=1, seed=0):
    rs = np.random.RandomState(seed)
    x = rs.uniform(-2, 2, size=(T, n, d)).astype(np.float32)
    y = (np.sin(2 * x[..., 0]) + 0.1 * rs.normal(size=(T, n))).astype(np.float32)
    return x, y


def _prior_particles(lay, P, seed):
    mu, sigma = orc.hyper_prior_params(lay, 0.5, 3.0)
    g = torch.Generator().manual_seed(seed)
    return (mu + sigma * torch.randn(P, lay.D, generator=g)).numpy()


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 20, 31, 32, 33, 40, 47, 48, 49, 50, 52, 63, 64,
                               65, 72, 80, 97, 100, 112, 127, 128])
def test_every_matrix_size_up_to_128(eng, n):
    """all NC instantiations of the warp-per-matrix kernel (n <= 32), the tensor-memory kernel with two matrices per CTA
    (n <= 64), the blocked Cholesky path with a single tile (64 < n <= 128), ragged last tile of the MLP kernels,
    repeated tasks."""
    x, y = _synthetic(5, n, seed=n)
    lay, arch = orc.Layout(1), eng.GPArch(1)
    theta = _prior_particles(lay, 3, 100 + n)
    idx = [4, 0, 0, 3, 1, 4, 2]
    mll, logp, score, info = run_engine(eng, arch, x, y, theta, idx)
    mll64, logp64, g64 = oracle64(lay, x, y, theta, idx)
    assert (info == 0).all()
    assert relmax(mll, mll64) <= RTOL and relmax(logp, logp64) <= RTOL
    assert_groups(arch, score, g64, ref32=oracle32(lay, x, y, theta, idx))


@pytest.mark.parametrize("kw", [
    dict(input_dim=1, mean_layers=(32,), kernel_layers=(32,)),
    dict(input_dim=1, mean_layers=(32, 32, 32), kernel_layers=(32, 32, 32)),
    dict(input_dim=1, mean_layers=(32,) * 4, kernel_layers=(32,) * 4),
    dict(input_dim=3, mean_layers=(32, 32), kernel_layers=(32, 32), feature_dim=4),
    dict(input_dim=2, mean_layers=(32, 32), kernel_layers=(16, 16), feature_dim=1),           # different depths / widths
    dict(input_dim=1, mean_layers=(64, 64), kernel_layers=(32, 32)),                         # generic + fast mixed
    dict(input_dim=6, mean_layers=(20, 12), kernel_layers=(40,), feature_dim=7),              # generic: d > 4, F > 4
    dict(input_dim=2, mean_kind="zero", covar_kind="NN"),
    dict(input_dim=3, mean_kind="NN", covar_kind="SE"),
    dict(input_dim=1, outputscale=True, noise_floor=1e-3),                                    # PACOH-MAP variant
])
def test_architectures_match_oracle(eng, kw):
    d = kw["input_dim"]
    x, y = _synthetic(6, 13, d=d, seed=3)
    lay, arch = orc.Layout(**kw), eng.GPArch(**kw)
    assert lay.D == arch.D
    theta = _prior_particles(lay, 4, 7)
    idx = [5, 1, 1, 0, 2, 3, 4, 4]
    mll, logp, score, info = run_engine(eng, arch, x, y, theta, idx)
    mll64, logp64, g64 = oracle64(lay, x, y, theta, idx)
    assert (info == 0).all()
    assert relmax(mll, mll64) <= RTOL and relmax(logp, logp64) <= RTOL
    assert_groups(arch, score, g64)


@pytest.mark.parametrize("n,kw", [
    (50, dict(input_dim=3, mean_layers=(32, 32), kernel_layers=(32, 32), feature_dim=4)),    # FT = 4 instantiation
    (33, dict(input_dim=3, mean_kind="NN", covar_kind="SE")),                                 # F = 3, raw inputs as features
    (64, dict(input_dim=2, mean_layers=(32, 32), kernel_layers=(16, 16), feature_dim=1)),    # full 64 rows, F = 1
    (41, dict(input_dim=2, mean_kind="zero", covar_kind="NN")),
    (57, dict(input_dim=1, outputscale=True, noise_floor=1e-3)),                              # PACOH-MAP variant
    (50, dict(input_dim=1, mean_kind="constant", covar_kind="SE")),
    (90, dict(input_dim=3, mean_layers=(32, 32), kernel_layers=(32, 32), feature_dim=4)),    # one matrix per CTA, FT = 4
    (128, dict(input_dim=1, outputscale=True, noise_floor=1e-3)),
])
def test_tensor_core_gp_kernel_shapes(eng, n, kw):
    """gp_tc.cu (32 < n <= 64): feature widths, mean kinds, output scale, odd task counts (the dummy second matrix)."""
    x, y = _synthetic(6, n, d=kw["input_dim"], seed=n)
    lay, arch = orc.Layout(**kw), eng.GPArch(**kw)
    theta = _prior_particles(lay, 3, 11 + n)
    for idx in ([5, 1, 1, 0, 2, 3, 4], [2], [0, 5, 3, 3]):
        mll, logp, score, info = run_engine(eng, arch, x, y, theta, idx)
        mll64, logp64, g64 = oracle64(lay, x, y, theta, idx)
        assert (info == 0).all()
        assert relmax(mll, mll64) <= RTOL and relmax(logp, logp64) <= RTOL
        assert_groups(arch, score, g64, tol=RTOL if n <= 64 else RTOL_N128)


@pytest.mark.parametrize("n_max,lo,kw", [
    (20, 3, dict(input_dim=1)),                                                            # register GP kernel
    (50, 33, dict(input_dim=1)),                                                           # tensor-memory GP kernel, all sizes > 32
    (64, 1, dict(input_dim=2, mean_kind="zero", covar_kind="NN")),                          # both kernels' size range in one batch
    (40, 7, dict(input_dim=3, mean_kind="constant", covar_kind="SE")),
    (57, 20, dict(input_dim=1, outputscale=True, noise_floor=1e-3)),                        # PACOH-MAP variant
    (100, 40, dict(input_dim=1)),                                                          # small kernels' sizes inside a large-n batch
    (300, 60, dict(input_dim=1)),                                                          # blocked path: 1 .. 3 tiles per matrix in one batch
    (260, 129, dict(input_dim=2, mean_kind="constant", covar_kind="SE")),
])
def test_ragged_task_sets_match_oracle(eng, n_max, lo, kw):
    """Tasks with different numbers of points (the reference's per-task loop handles them implicitly, random_gp.py:214-217;
    pre-factor with the harmonic mean, :209-212): padded arrays + task_n through pacoh_meta_mll_fwd_bwd_ragged."""
    T_total, d = 9, kw["input_dim"]
    rs = np.random.RandomState(n_max + lo)
    sizes = rs.randint(lo, n_max + 1, size=T_total)
    sizes[0], sizes[1] = n_max, lo
    x = rs.uniform(-2, 2, size=(T_total, n_max, d)).astype(np.float32)
    y = (np.sin(2 * x[..., 0]) + 0.1 * rs.normal(size=(T_total, n_max))).astype(np.float32)
    for t in range(T_total):                                  # poison the padding: it must not influence anything
        x[t, sizes[t]:] = 7.5
        y[t, sizes[t]:] = -3.0
    lay, arch = orc.Layout(**kw), eng.GPArch(**kw)
    theta = _prior_particles(lay, 3, 5 + n_max)
    idx = [8, 1, 1, 0, 2, 3, 4, 7, 6, 5, 0]
    e = eng.MetaMLLEngine(arch, x, y, DEV, task_n=sizes)
    th = torch.from_numpy(theta).to(DEV)
    mll, packed, info = e.mll_fwd_bwd(th, torch.tensor(idx, dtype=torch.int32, device=DEV))
    mu, sigma = arch.hyper_prior(0.5, 3.0)
    pre = eng.pre_factor(sizes[idx])
    logp, dth = eng.logprob_finalize(th, mu.to(DEV), sigma.to(DEV), 0.01, pre, packed)
    tasks = [(torch.from_numpy(x[t, :sizes[t]]).double(), torch.from_numpy(y[t, :sizes[t]]).double()) for t in range(T_total)]
    mu64, s64 = orc.hyper_prior_params(lay, 0.5, 3.0, torch.float64)
    logp64, g64, mll64 = orc.meta_log_prob_and_grad(torch.from_numpy(theta).double(), lay, [tasks[i] for i in idx], 0.01, mu64, s64)
    assert (info.cpu().numpy() == 0).all()
    assert relmax(mll.cpu().numpy(), mll64.numpy()) <= RTOL and relmax(logp.cpu().numpy(), logp64.numpy()) <= RTOL
    assert_groups(arch, dth.cpu().numpy(), g64.numpy(), tol=RTOL if n_max <= 64 else RTOL_N128)


def test_map_demo_first_iteration_matches_logged_loss(eng, golden_dir):
    """BASELINE config #1: the engine's P=1 ScaleRBF/noise-floor path reproduces demo.ipynb's 'Loss: 5.755850'."""
    train = orc.sinusoid_tasks(20, 5, seed=26)
    m = orc.MAPOracle(train, weight_decay=0.2, seed=30)
    kw = dict(input_dim=1, outputscale=True, noise_floor=1e-3)
    lay, arch = orc.Layout(**kw), eng.GPArch(**kw)
    theta = m.flat_parameters(lay)
    x = np.stack([t[0].numpy() for t in m.tasks]); y = np.stack([t[1].numpy() for t in m.tasks])
    idx = [18, 16, 2, 6, 10]
    e = eng.MetaMLLEngine(arch, x, y, DEV)
    mll, packed, info = e.mll_fwd_bwd(theta.to(DEV), torch.tensor(idx, dtype=torch.int32, device=DEV))
    loss = -float(packed[-1].item())
    assert "%.5f" % loss == "5.75585"
    # gradient of the MAP loss w.r.t. the torch modules, packed into the flat layout
    total = 0.0
    for i in idx:
        total = total - m.mll(*m.tasks[i])
    total.backward()
    g_ref = torch.zeros(lay.D)
    for prefix, net in (("mean_nn", m.mean_nn), ("kernel_nn", m.kernel_nn)):
        lin = [mod for mod in net if isinstance(mod, torch.nn.Linear)]
        names = ["fc_%d" % (k + 1) for k in range(len(lin) - 1)] + ["out"]
        for nm, l in zip(names, lin):
            a, b = lay.entries["%s.%s.bias" % (prefix, nm)]; g_ref[a:b] = l.bias.grad
            a, b = lay.entries["%s.%s.weight" % (prefix, nm)]; g_ref[a:b] = l.weight.grad.reshape(-1)
    a, b = lay.entries["lengthscale_raw"]; g_ref[a:b] = m.raw_lengthscale.grad.reshape(-1)
    a, b = lay.entries["noise_raw"]; g_ref[a:b] = m.raw_noise.grad
    a, b = lay.entries["outputscale_raw"]; g_ref[a:b] = m.raw_outputscale.grad.reshape(1)
    got = -packed[:lay.D].cpu().numpy().reshape(1, -1)
    assert_groups(arch, got, g_ref.numpy().reshape(1, -1))


@pytest.mark.parametrize("n", [8, 40, 96, 200])  # register kernel / tensor-memory kernel (two matrices per CTA retry together) / blocked path (retry list)
def test_jitter_ladder_and_not_psd_reporting(eng, n):
    """duplicate inputs + vanishing noise: singular K.  The kernel must climb the 1e-6/1e-5/1e-4 jitter ladder
    (gpytorch psd_safe_cholesky) or report failure; the host wrapper raises NotPSDError on failure."""
    arch = eng.GPArch(1, mean_kind="zero", covar_kind="SE")
    x = np.zeros((2, n, 1), np.float32)             # all points identical -> rank-1 Gram
    x[1] = np.linspace(-1, 1, n, dtype=np.float32).reshape(n, 1)
    y = np.ones((2, n), np.float32)
    theta = np.zeros((2, arch.D), np.float32)
    theta[0, arch.entries()["noise_raw"][0]] = -40.0      # softplus(-40) ~ 4e-18
    theta[1, arch.entries()["noise_raw"][0]] = 0.0
    e = eng.MetaMLLEngine(arch, x, y, DEV)
    mll, packed, info = e.mll_fwd_bwd(torch.from_numpy(theta).to(DEV), torch.tensor([0, 1], dtype=torch.int32, device=DEV))
    info = info.cpu().numpy()
    assert info[1, 0] == 0 and info[1, 1] == 0           # healthy particle untouched
    assert info[0, 0] != 0                               # singular matrix needed jitter (or failed)
    if (info < 0).any():
        with pytest.raises(eng.NotPSDError):
            eng.check_info(torch.from_numpy(info))
    else:
        assert np.isfinite(mll.cpu().numpy()).all()
    # the healthy particle's values do not depend on the retries of the matrices it shares a launch / CTA with
    mll1, packed1, info1 = e.mll_fwd_bwd(torch.from_numpy(theta[1:2]).to(DEV), torch.tensor([0, 1], dtype=torch.int32, device=DEV))
    assert torch.equal(mll1[0], mll[1]) and int(info1.abs().max()) == 0
    D = arch.D
    assert torch.equal(packed1[:D], packed[D:2 * D])


@pytest.mark.parametrize("n,P,T,kw", [
    (129, 2, 3, dict(input_dim=1)),                                                            # two tiles, the second almost empty
    (200, 2, 4, dict(input_dim=1)),
    (256, 1, 3, dict(input_dim=1, outputscale=True, noise_floor=1e-3)),                        # exactly two full tiles
    (300, 2, 3, dict(input_dim=3, mean_layers=(32, 32), kernel_layers=(32, 32), feature_dim=4)),
    (385, 1, 2, dict(input_dim=2, mean_kind="constant", covar_kind="SE")),
    (512, 2, 2, dict(input_dim=1)),
])
def test_blocked_cholesky_path_matches_oracle(eng, n, P, T, kw):
    """csrc/gp_big.cu (n > 64): left-looking blocked Cholesky, U = L^-T, Khat^-1 tiles on tcgen05 (3xTF32), prior particles."""
    x, y = _synthetic(4, n, d=kw["input_dim"], seed=n)
    lay, arch = orc.Layout(**kw), eng.GPArch(**kw)
    theta = _prior_particles(lay, P, 300 + n)
    idx = [3, 0, 0, 2][:T]
    mll, logp, score, info = run_engine(eng, arch, x, y, theta, idx)
    mll64, logp64, g64 = oracle64(lay, x, y, theta, idx)
    assert (info == 0).all()
    assert relmax(mll, mll64) <= RTOL and relmax(logp, logp64) <= RTOL
    assert_groups(arch, score, g64, ref32=oracle32(lay, x, y, theta, idx))


@pytest.mark.parametrize("n", [512, 1024, 2048])
def test_config5_sizes_match_fp64_oracle(eng, n):
    """BASELINE config #5 (PACOH-MAP, 512 - 2048 points per task: GPR_meta_mll.py:104-119 with dense Cholesky forced):
    sinusoid tasks, torch.nn.Linear default initialisation, raw hyper-parameters 0, spot-check of 3 tasks against fp64."""
    kw = dict(input_dim=1, outputscale=True, noise_floor=1e-3)
    lay, arch = orc.Layout(**kw), eng.GPArch(**kw)
    train = orc.sinusoid_tasks(4, n, seed=26)
    stats = orc.normalization_stats(train)
    tasks = [orc.prepare_task(xx, yy, stats, torch.float64) for xx, yy in train]
    x = np.stack([t[0].numpy() for t in tasks]).astype(np.float32)
    y = np.stack([t[1].numpy() for t in tasks]).astype(np.float32)
    theta = orc.MAPOracle(train, weight_decay=0.0, seed=30).flat_parameters(lay).numpy()
    idx = [2, 0, 3]
    e = eng.MetaMLLEngine(arch, x, y, DEV)
    mll, packed, info = e.mll_fwd_bwd(torch.from_numpy(theta).to(DEV), torch.tensor(idx, dtype=torch.int32, device=DEV))
    mu, sigma = orc.hyper_prior_params(lay, 0.5, 3.0, torch.float64)
    batch = [(torch.from_numpy(x[i]).double(), torch.from_numpy(y[i]).double()) for i in idx]
    _, g64, mll64 = orc.meta_log_prob_and_grad(torch.from_numpy(theta).double(), lay, batch, 0.0, mu, sigma)
    g64 = g64.numpy() / orc.pre_factor([n] * len(idx))                    # d sum_t mll_t / d theta
    mu32, sigma32 = orc.hyper_prior_params(lay, 0.5, 3.0, torch.float32)
    _, g32, _ = orc.meta_log_prob_and_grad(torch.from_numpy(theta).float(), lay, [(a.float(), b.float()) for a, b in batch], 0.0, mu32, sigma32)
    g32 = g32.double().numpy() / orc.pre_factor([n] * len(idx))
    assert (info.cpu().numpy() == 0).all()
    assert relmax(mll.cpu().numpy(), mll64.numpy()) <= RTOL
    assert abs(float(packed[-1]) - float(mll64.sum())) <= RTOL * abs(float(mll64.sum()))
    assert_groups(arch, packed[:lay.D].cpu().numpy()[None, :], g64, ref32=g32)


def test_config5_full_batch_properties(eng):
    """1024 tasks x 1024 points (config #5's middle size) in one call: duplicates count twice, the batch sum is the sum of
    the per-task values, repeated calls are bitwise identical."""
    n, T = 1024, 1024
    kw = dict(input_dim=1, outputscale=True, noise_floor=1e-3)
    lay, arch = orc.Layout(**kw), eng.GPArch(**kw)
    x, y = _synthetic(T, n, seed=5)
    theta = torch.from_numpy(orc.MAPOracle(orc.sinusoid_tasks(2, 8, seed=26), weight_decay=0.0, seed=30).flat_parameters(lay).numpy()).to(DEV)
    e = eng.MetaMLLEngine(arch, x, y, DEV)
    idx = torch.arange(T, dtype=torch.int32, device=DEV)
    mll, packed, info = e.mll_fwd_bwd(theta, idx)
    assert int(info.abs().max()) == 0 and bool(torch.isfinite(mll).all())
    assert abs(float(packed[-1]) - float(mll.double().sum())) <= 1e-5 * abs(float(mll.double().sum()))
    mll2, packed2, _ = e.mll_fwd_bwd(theta, idx)
    assert torch.equal(packed, packed2) and torch.equal(mll, mll2)
    sub = torch.tensor([5, 9, 9, 700], dtype=torch.int32, device=DEV)
    mll_s, packed_s, _ = e.mll_fwd_bwd(theta, sub)
    assert torch.equal(mll_s[0, 1], mll_s[0, 2])
    assert relmax(mll_s.cpu().numpy()[0], mll.cpu().numpy()[0, [5, 9, 9, 700]]) <= 1e-6
    one = [e.mll_fwd_bwd(theta, torch.tensor([i], dtype=torch.int32, device=DEV))[1][:lay.D].double() for i in (5, 9, 700)]
    want = one[0] + 2 * one[1] + one[2]
    assert float((packed_s[:lay.D].double() - want).abs().max()) <= 1e-5 * float(want.abs().max())


def test_unsupported_sizes_fail_loudly(eng):
    from meta_learning_pacoh_b200._lib import PacohError
    arch = eng.GPArch(1)
    x, y = _synthetic(2, 4100)
    with pytest.raises(PacohError):
        e = eng.MetaMLLEngine(arch, x, y, DEV)
        e.mll_fwd_bwd(torch.zeros(2, arch.D, device=DEV), torch.tensor([0, 1], dtype=torch.int32, device=DEV))
    wide = eng.GPArch(6, mean_layers=(20, 12), kernel_layers=(40,), feature_dim=7)       # F > 4 has no large-n kernel
    x, y = _synthetic(2, 130, d=6)
    with pytest.raises(PacohError):
        e = eng.MetaMLLEngine(wide, x, y, DEV)
        e.mll_fwd_bwd(torch.zeros(2, wide.D, device=DEV), torch.tensor([0, 1], dtype=torch.int32, device=DEV))
    with pytest.raises(NotImplementedError):
        eng.SVGDDirection(4, 10, DEV, kernel="Matern")


# ------------------------------------------------------------------------------------------ full-size properties
@pytest.fixture(scope="module")
def config4(eng):
    P, T, n = 64, 4096, 50
    train = orc.sinusoid_tasks(T, n, seed=26)
    stats = orc.normalization_stats(train)
    x = np.stack([(a - stats[0]) / stats[1] for a, _ in train]).astype(np.float32)
    y = np.stack([((b - stats[2]) / stats[3]).reshape(-1) for _, b in train]).astype(np.float32)
    lay, arch = orc.Layout(1), eng.GPArch(1)
    theta = torch.from_numpy(_prior_particles(lay, P, 30)).to(DEV)
    e = eng.MetaMLLEngine(arch, x, y, DEV)
    idx = np.random.RandomState(31).choice(T, size=T).astype(np.int32)
    return dict(P=P, T=T, n=n, x=x, y=y, lay=lay, arch=arch, theta=theta, e=e, idx=idx)


def test_full_size_deterministic_and_sharding_is_linear(eng, config4):
    c = config4
    tidx = torch.from_numpy(c["idx"]).to(DEV)
    mll_a, packed_a, info = c["e"].mll_fwd_bwd(c["theta"], tidx)
    mll_b, packed_b, _ = c["e"].mll_fwd_bwd(c["theta"], tidx)
    assert int(info.min()) == 0 and int(info.max()) == 0
    assert torch.equal(packed_a, packed_b) and torch.equal(mll_a, mll_b)          # bitwise repeatable
    assert torch.isfinite(packed_a).all()
    # task sharding (multi-GPU decomposition): the sum over 4 shards equals the full batch
    acc = torch.zeros_like(packed_a, dtype=torch.float64)
    for s in range(4):
        _, pk, _ = c["e"].mll_fwd_bwd(c["theta"], tidx[s * 1024:(s + 1) * 1024].contiguous())
        acc += pk.double()
    scale = packed_a.abs().max().item()
    # two fp32 evaluation orders of a 204800-point sum: half the parity bar (the fp64 check below is the real one)
    assert (acc - packed_a.double()).abs().max().item() <= 5e-5 * scale
    # per-task values do not depend on the batch they are evaluated in
    mll_s, _, _ = c["e"].mll_fwd_bwd(c["theta"], tidx[:1024].contiguous())
    assert torch.equal(mll_s, mll_a[:, :1024])


def test_full_size_spot_check_against_oracle(eng, config4):
    """64 x 4096 x 50 is too big for the CPU oracle; check a random subset of (particle, task) values in fp64 and
    the gradient on a 32-task sub-batch (size-independent: the kernels process a task identically in any batch)."""
    c = config4
    tidx = torch.from_numpy(c["idx"]).to(DEV)
    mll, _, _ = c["e"].mll_fwd_bwd(c["theta"], tidx)
    mll = mll.cpu().numpy()
    rs = np.random.RandomState(0)
    th64 = c["theta"].cpu().double()
    for t in rs.choice(c["T"], size=6, replace=False):
        src = c["idx"][t]
        ref = orc.task_mll(th64, c["lay"], torch.from_numpy(c["x"][src]).double(), torch.from_numpy(c["y"][src]).double()).numpy()
        assert np.abs(mll[:, t] - ref).max() <= RTOL * np.abs(ref).max()
    sub = c["idx"][:32]
    P8 = 8
    _, logp, score, _ = run_engine(eng, c["arch"], c["x"], c["y"], c["theta"][:P8].cpu().numpy(), sub)
    _, logp64, g64 = oracle64(c["lay"], c["x"], c["y"], c["theta"][:P8].cpu().numpy(), sub)
    assert relmax(logp, logp64) <= RTOL
    assert_groups(c["arch"], score, g64)


def test_full_size_gradient_against_fp64_oracle(eng, config4):
    """The whole 64 x 4096 x 50 batch through the kernels (so every accumulator sees its real length: the weight
    gradients are summed over 204800 points per particle) against the fp64 oracle for three of the particles."""
    c = config4
    P, D, K = c["P"], c["lay"].D, 3
    _, packed, info = c["e"].mll_fwd_bwd(c["theta"], torch.from_numpy(c["idx"]).to(DEV))
    assert int(info.max()) == 0
    g = packed[:P * D].view(P, D).cpu().double()
    msum = packed[P * D:].cpu().double()
    th64 = c["theta"][:K].cpu().double().requires_grad_(True)
    xs, ys = torch.from_numpy(c["x"]).double(), torch.from_numpy(c["y"]).double()
    tot = torch.zeros(K, dtype=torch.float64)
    for t in c["idx"]:
        tot = tot + orc.task_mll(th64, c["lay"], xs[t], ys[t])
    tot.sum().backward()
    assert ((msum[:K] - tot.detach()).abs() / tot.detach().abs()).max().item() <= RTOL
    for k in range(K):
        assert (g[k] - th64.grad[k]).abs().max().item() <= 0.5 * RTOL * th64.grad[k].abs().max().item()


def test_full_size_svgd_direction_properties(eng, config4):
    c = config4
    P, D = c["theta"].shape
    score = torch.randn(P, D, device=DEV)
    sv = eng.SVGDDirection(P, D, DEV)
    phi = sv(c["theta"], score)
    ref, gamma = orc.svgd_phi(c["theta"].cpu().double(), score.cpu().double())
    assert abs(float(sv.gamma.item()) - gamma) <= 1e-5 * gamma
    assert relmax(phi.cpu().numpy(), ref.numpy()) <= RTOL
    # permutation equivariance: permuting the particles permutes phi
    perm = torch.randperm(P, device=DEV)
    phi_p = sv(c["theta"][perm].contiguous(), score[perm].contiguous())
    assert (phi_p - phi[perm]).abs().max().item() <= 1e-5 * phi.abs().max().item()


def test_cuda_core_mlp_kernels_still_match_oracle():
    """The FFMA versions of the MLP kernels (mlp.cu) are selected by PACOH_MLP_FWD/BWD=ffma (read once per process):
    run them in a subprocess against the fp64 oracle so both implementations stay parity-green."""
    import subprocess
    import sys
    env = dict(os.environ, PACOH_MLP_FWD="ffma", PACOH_MLP_BWD="ffma")
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ffma_path_check.py")
    r = subprocess.run([sys.executable, script, "ffma"], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr


def test_register_gp_kernel_still_matches_oracle():
    """32 < n <= 64 runs on the tensor-memory GP kernel (gp_tc.cu); PACOH_GP=warp selects the register kernel of
    gp_mll.cu for those sizes (kept for A/B measurements): same subprocess check against the fp64 oracle."""
    import subprocess
    import sys
    env = dict(os.environ, PACOH_GP="warp")
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ffma_path_check.py")
    r = subprocess.run([sys.executable, script, "gpwarp"], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
