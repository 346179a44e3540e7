"""CPU-only tests of the host-side logic above the C ABI: RNG-parity initialisation, data handling, distributions,
quantile root finder, optimizer state compatibility.  No kernel is launched."""
import math
import os

import numpy as np
import pytest
import torch

from meta_learning_pacoh_b200 import engine as eng
from meta_learning_pacoh_b200.meta_learn import prior_init
from meta_learning_pacoh_b200.meta_learn.distributions import AffineTransformedDistribution, EqualWeightedMixtureDist
from meta_learning_pacoh_b200.meta_learn.util import _handle_input_dimensionality, find_root_by_bounding


def test_particle_init_reproduces_reference_rng_stream(golden_dir):
    """same seed => same particles as RandomGPMeta(...).sample_params_from_prior in the reference (GPR_meta_svgd.py:182)."""
    fx = np.load(os.path.join(golden_dir, "svgd_cfg2.npz"))
    torch.manual_seed(30)
    arch = eng.GPArch(1)
    prior_init.consume_vectorized_gp_init(arch)
    p = prior_init.sample_params_from_prior(arch, 10, 0.5, 3.0)
    assert np.array_equal(p.numpy(), fx["particles"])
    fx = np.load(os.path.join(golden_dir, "svgd_arch.npz"))
    torch.manual_seed(32)
    arch = eng.GPArch(2, mean_layers=(16,), kernel_layers=(8, 24, 16))
    prior_init.consume_vectorized_gp_init(arch)
    assert np.array_equal(prior_init.sample_params_from_prior(arch, 5, 0.5, 3.0).numpy(), fx["particles"])


def test_vi_posterior_init_and_eps_reproduce_reference_rng_stream(golden_dir):
    fx = np.load(os.path.join(golden_dir, "vi_cfg3.npz"))
    torch.manual_seed(34)
    arch = eng.GPArch(1)
    prior_init.consume_vectorized_gp_init(arch)
    loc, scale = prior_init.init_diag_posterior(arch.D)
    assert np.array_equal(loc.numpy(), fx["loc"]) and np.array_equal(scale.numpy(), fx["scale"])
    assert np.array_equal(torch.empty(8, arch.D).normal_().numpy(), fx["eps"])


def test_input_dimensionality_contract():
    x, y = _handle_input_dimensionality(np.zeros(5), np.zeros(5))
    assert x.shape == (5, 1) and y.shape == (5, 1)
    with pytest.raises(AssertionError):
        _handle_input_dimensionality(np.zeros((5, 1)), np.zeros(4))
    with pytest.raises(AssertionError):
        _handle_input_dimensionality(np.zeros((5, 1, 1)))


def test_mixture_distribution_moments_logprob_cdf_icdf():
    torch.manual_seed(22)
    means, stds = torch.randn(4, 7), torch.rand(4, 7) + 0.1
    mix = EqualWeightedMixtureDist(torch.distributions.Normal(means, stds), batched=True)
    lst = EqualWeightedMixtureDist([torch.distributions.Normal(means[i], stds[i]) for i in range(4)])
    assert torch.allclose(mix.mean, means.mean(0)) and torch.allclose(mix.mean, lst.mean)
    var = ((means - means.mean(0)) ** 2).mean(0) + (stds ** 2).mean(0)
    assert torch.allclose(mix.variance, var) and torch.allclose(lst.stddev, var.sqrt())
    v = torch.randn(7)
    lp = torch.logsumexp(torch.distributions.Normal(means, stds).log_prob(v), 0) - math.log(4.0)
    assert torch.allclose(mix.log_prob(v), lp) and torch.allclose(lst.log_prob(v), lp)
    q = torch.full((7,), 0.9)
    x = mix.icdf(q)
    assert torch.allclose(mix.cdf(x), q, atol=1e-4)


def test_affine_transformed_distribution():
    base = torch.distributions.MultivariateNormal(torch.tensor([0.5, -1.0]), covariance_matrix=torch.tensor([[2.0, 0.3], [0.3, 1.0]]))
    d = AffineTransformedDistribution(base, normalization_mean=np.array([3.0]), normalization_std=np.array([2.0]))
    assert torch.allclose(d.mean, torch.tensor([4.0, 1.0]))
    assert torch.allclose(d.variance, torch.tensor([8.0, 4.0])) and torch.allclose(d.stddev, torch.tensor([8.0, 4.0]).sqrt())
    y = torch.tensor([3.5, 0.0])
    ref = base.log_prob((y - 3.0) / 2.0) - 2 * math.log(2.0)
    assert torch.allclose(d.log_prob(y), ref, atol=1e-6)


def test_find_root_by_bounding_quantiles():
    n = torch.distributions.Normal(torch.tensor([0.0, 2.0]), torch.tensor([1.0, 0.5]))
    for q in (0.05, 0.5, 0.95):
        target = torch.full((2,), q)
        x = find_root_by_bounding(lambda v: n.cdf(v) - target, -1e8 * torch.ones(2), 1e8 * torch.ones(2))
        assert torch.allclose(x, n.icdf(target), atol=1e-4)


def test_learner_classes_keep_reference_signatures():
    import inspect
    from meta_learning_pacoh_b200.meta_learn import GPRegressionMetaLearned, GPRegressionMetaLearnedSVGD, GPRegressionMetaLearnedVI
    sig = inspect.signature(GPRegressionMetaLearnedSVGD.__init__)
    assert list(sig.parameters)[1:] == ["meta_train_data", "num_iter_fit", "feature_dim", "prior_factor", "weight_prior_std",
                                        "bias_prior_std", "covar_module", "mean_module", "mean_nn_layers", "kernel_nn_layers",
                                        "optimizer", "lr", "lr_decay", "kernel", "bandwidth", "num_particles", "task_batch_size",
                                        "normalize_data", "random_seed"]
    assert sig.parameters["num_particles"].default == 10 and sig.parameters["prior_factor"].default == 0.01
    sig = inspect.signature(GPRegressionMetaLearnedVI.__init__)
    assert sig.parameters["svi_batch_size"].default == 10 and sig.parameters["cov_type"].default == "diag"
    sig = inspect.signature(GPRegressionMetaLearned.__init__)
    assert sig.parameters["task_batch_size"].default == 5 and sig.parameters["weight_decay"].default == 0.0
    assert list(inspect.signature(GPRegressionMetaLearned.meta_fit).parameters)[1:] == ["valid_tuples", "verbose", "log_period", "n_iter"]
    for cls in (GPRegressionMetaLearned, GPRegressionMetaLearnedSVGD, GPRegressionMetaLearnedVI):
        for m in ("meta_fit", "predict", "eval", "eval_datasets", "confidence_intervals"):
            assert callable(getattr(cls, m))


def test_persistent_backward_schedule_covers_every_tile_once():
    """The tensor-core MLP backward walks the linearised (net, particle, tile) space in equal per-CTA ranges and writes one
    partial-gradient slot per (CTA, (net, particle)) it touches (mlp_tc_bwd.cu): restate the kernel's index arithmetic and
    check coverage, slot uniqueness, the slot bound the workspace is sized with, and the accumulation-length bound."""
    import ctypes
    from meta_learning_pacoh_b200 import _lib
    for P, nets, T, n in [(64, 2, 4096, 50), (64, 2, 512, 50), (10, 2, 20, 5), (1, 1, 5, 5), (3, 2, 7, 64), (8, 1, 256, 20),
                          (64, 2, 16384, 50), (5, 2, 11, 128), (128, 2, 33, 37)]:
        g, per, slots = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
        _lib.check(_lib.lib.pacoh_mlp_bwd_schedule(P, nets, T * n, ctypes.byref(g), ctypes.byref(per), ctypes.byref(slots)))
        grid, per_cta, slots = g.value, per.value, slots.value
        tiles = (T * n + 127) // 128
        total = nets * P * tiles
        assert (grid - 1) * per_cta < total <= grid * per_cta
        seen = np.zeros(total, dtype=np.int32)
        used = set()
        for cta in range(grid):
            g0, g1 = cta * per_cta, min(total, (cta + 1) * per_cta)
            while g0 < g1:
                pn = g0 // tiles
                t0, t1 = g0 - pn * tiles, min(g1, (pn + 1) * tiles) - pn * tiles
                slot = cta - (pn * tiles) // per_cta
                assert 0 <= slot < slots and (pn, slot) not in used
                used.add((pn, slot))
                seen[pn * tiles + t0:pn * tiles + t1] += 1
                g0 += t1 - t0
        assert (seen == 1).all()
        assert per_cta <= max(6, 3 * 120 + 2)           # <= 120 tiles per warpgroup accumulator (3 warpgroups per CTA)



def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the oracle port on the host cores, a bounded slice of the workload) runs without a GPU and
    prints ONE JSON line with the contract's keys; config 2 is the reference's own CPU-runnable shape."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--config", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "task-particle MLL+grad evals/s" and d["unit"] == "evals/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["config"]["config_id"] == 2 and d["config"]["evals_per_step"] == 200
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
