"""GPU bring-up diagnostics: runs the CUDA path on several shapes and prints per-parameter-group errors against the
fp64 oracle.  Never stops at the first mismatch (one gpurun call should tell as much as possible)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pacoh_oracle as orc  # noqa: E402
from meta_learning_pacoh_b200 import engine as eng  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
dev = torch.device("cuda:0")


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def run_case(name, lay_kw, arch, x, y, theta, idx, prior_factor=0.01):
    lay = orc.Layout(**lay_kw)
    assert lay.D == arch.D, (lay.D, arch.D)
    tasks64 = [(torch.from_numpy(x[i]).double(), torch.from_numpy(y[i]).double()) for i in range(x.shape[0])]
    batch = [tasks64[i] for i in idx]
    mu, sigma = orc.hyper_prior_params(lay, 0.5, 3.0, torch.float64)
    logp64, score64, mll64 = orc.meta_log_prob_and_grad(torch.from_numpy(theta).double(), lay, batch, prior_factor, mu, sigma)
    e = eng.MetaMLLEngine(arch, x, y, dev)
    th = torch.from_numpy(theta).to(dev)
    tidx = torch.from_numpy(np.asarray(idx, dtype=np.int32)).to(dev)
    mll, packed, info = e.mll_fwd_bwd(th, tidx)
    torch.cuda.synchronize()
    pmu, psig = arch.hyper_prior(0.5, 3.0)
    pre = eng.pre_factor([x.shape[1]] * len(idx))
    logp, dth = eng.logprob_finalize(th, pmu.to(dev), psig.to(dev), prior_factor, pre, packed)
    torch.cuda.synchronize()
    print("== %s: P=%d T=%d n=%d d=%d D=%d info[min,max]=%d,%d" % (name, theta.shape[0], len(idx), x.shape[1], x.shape[2], arch.D,
                                                                  int(info.min()), int(info.max())))
    print("   mll   rel err %.3e   (max |mll| %.3f)" % (rel(mll.cpu().numpy(), mll64.numpy()), float(mll64.abs().max())))
    print("   logp  rel err %.3e" % rel(logp.cpu().numpy(), logp64.numpy()))
    g, g64 = dth.cpu().numpy(), score64.numpy()
    print("   score rel err %.3e (global)" % rel(g, g64))
    for nm, (a, b) in arch.entries().items():
        print("      %-24s rel %.3e   |ref|max %.3e" % (nm, rel(g[:, a:b], g64[:, a:b]), np.abs(g64[:, a:b]).max()))
    return e, th, tidx


def main():
    torch.manual_seed(0)
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_properties(0).multi_processor_count, "SMs")
    fx = np.load(os.path.join(GOLD, "svgd_cfg2.npz"))
    run_case("cfg2 (golden, n=5)", dict(input_dim=1), eng.GPArch(1), fx["x"], fx["y"], fx["particles"], fx["idx"])
    fx = np.load(os.path.join(GOLD, "svgd_n20.npz"))
    run_case("n20 (golden)", dict(input_dim=1), eng.GPArch(1), fx["x"], fx["y"], fx["particles"], fx["idx"])
    fx = np.load(os.path.join(GOLD, "svgd_arch.npz"))
    run_case("arch d=2 (16,)/(8,24,16)", dict(input_dim=2, mean_layers=(16,), kernel_layers=(8, 24, 16)),
             eng.GPArch(2, mean_layers=(16,), kernel_layers=(8, 24, 16)), fx["x"], fx["y"], fx["particles"], fx["idx"])
    fx = np.load(os.path.join(GOLD, "const_se.npz"))
    run_case("constant mean / SE d=2", dict(input_dim=2, mean_kind="constant", covar_kind="SE"),
             eng.GPArch(2, mean_kind="constant", covar_kind="SE"), fx["x"], fx["y"], fx["particles"], fx["idx"])

    # n = 50 synthetic (config #4 shape, small T)
    train = orc.sinusoid_tasks(24, 50, seed=26)
    stats = orc.normalization_stats(train)
    prep = [orc.prepare_task(a, b, stats) for a, b in train]
    x = np.stack([a.numpy() for a, _ in prep]); y = np.stack([b.numpy() for _, b in prep])
    lay = orc.Layout(1)
    mu, sigma = orc.hyper_prior_params(lay, 0.5, 3.0)
    g = torch.Generator().manual_seed(30)
    theta = (mu + sigma * torch.randn(16, lay.D, generator=g)).numpy()
    idx = np.random.RandomState(31).choice(24, size=24)
    e, th, tidx = run_case("n50 P=16 T=24", dict(input_dim=1), eng.GPArch(1), x, y, theta, idx)

    # generic path: 4 x 128 MAP-like nets, outputscale + noise floor
    arch = eng.GPArch(1, mean_layers=(128,) * 4, kernel_layers=(128,) * 4, outputscale=True, noise_floor=1e-3)
    lay_kw = dict(input_dim=1, mean_layers=(128,) * 4, kernel_layers=(128,) * 4, outputscale=True, noise_floor=1e-3)
    lay = orc.Layout(**lay_kw)
    theta = (0.1 * torch.randn(2, lay.D, generator=g)).numpy()
    run_case("generic 4x128 MAP", lay_kw, arch, x[:6, :20].copy(), y[:6, :20].copy(), theta, np.arange(6))
    # 4 x 32 (experiment default), fast path L=4
    arch = eng.GPArch(1, mean_layers=(32,) * 4, kernel_layers=(32,) * 4)
    lay_kw = dict(input_dim=1, mean_layers=(32,) * 4, kernel_layers=(32,) * 4)
    lay = orc.Layout(**lay_kw)
    mu, sigma = orc.hyper_prior_params(lay, 0.5, 3.0)
    theta = (mu + sigma * torch.randn(4, lay.D, generator=g)).numpy()
    run_case("4x32", lay_kw, arch, x[:8, :20].copy(), y[:8, :20].copy(), theta, np.arange(8))

    # SVGD phi
    fx = np.load(os.path.join(GOLD, "svgd_cfg2.npz"))
    P, D = fx["particles"].shape
    sv = eng.SVGDDirection(P, D, dev)
    phi = sv(torch.from_numpy(fx["particles"]).to(dev), torch.from_numpy(fx["score"]).to(dev))
    print("== svgd phi rel err %.3e  gamma %.6e vs %.6e" % (rel(phi.cpu().numpy(), fx["phi"]), float(sv.gamma.item()), float(fx["gamma"])))
    # VI
    fx = np.load(os.path.join(GOLD, "vi_cfg3.npz"))
    loc, scale, eps = (torch.from_numpy(fx[k]).to(dev) for k in ("loc", "scale", "eps"))
    theta, logq = eng.vi_sample(loc, scale, eps)
    print("== vi theta err %.3e logq rel %.3e" % (rel(theta.cpu().numpy(), fx["theta"]), rel(logq.cpu().numpy(), fx["logq"])))

    # timing at config-4-like size (smaller T to stay quick)
    for (P, T, n) in ((64, 512, 50), (64, 4096, 50)):
        train = orc.sinusoid_tasks(T, n, seed=26)
        stats = orc.normalization_stats(train)
        x = np.stack([((a - stats[0]) / stats[1]) for a, _ in train]).astype(np.float32)
        y = np.stack([((b - stats[2]) / stats[3]).reshape(-1) for _, b in train]).astype(np.float32)
        lay = orc.Layout(1)
        mu, sigma = orc.hyper_prior_params(lay, 0.5, 3.0)
        theta = (mu + sigma * torch.randn(P, lay.D, generator=g)).to(dev)
        e = eng.MetaMLLEngine(eng.GPArch(1), x, y, dev)
        tidx = torch.from_numpy(np.random.RandomState(31).choice(T, size=T).astype(np.int32)).to(dev)
        for _ in range(3):
            mll, packed, info = e.mll_fwd_bwd(theta, tidx)
        torch.cuda.synchronize()
        t0 = time.time()
        reps = 5
        for _ in range(reps):
            e.mll_fwd_bwd(theta, tidx, want_mll=False, want_info=False)
        torch.cuda.synchronize()
        dt = (time.time() - t0) / reps
        print("== timing P=%d T=%d n=%d: %.3f ms/step  %.2f M evals/s   info min %d max %d  mll finite %s" % (
            P, T, n, dt * 1e3, P * T / dt / 1e6, int(info.min()), int(info.max()), bool(torch.isfinite(mll).all())))
    print("ffma peak TFLOP/s: %.1f" % eng.ffma_peak_tflops())


if __name__ == "__main__":
    main()
