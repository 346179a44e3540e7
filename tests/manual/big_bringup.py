"""Bring-up diagnostics of the large-n GP path (csrc/gp_big.cu): for several task sizes, compares the factors the
kernels leave in the workspace (L, U = L^-T, v, alphahat, via pacoh_debug_big_layout) with an fp64 Cholesky, then the
outputs (mll, logp, gradients per parameter group) with the fp64 oracle.  Never stops at the first mismatch.

    python tests/manual/big_bringup.py [n ...]
"""
import ctypes
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pacoh_oracle as orc  # noqa: E402
from meta_learning_pacoh_b200 import engine as eng, _lib  # noqa: E402

dev = torch.device("cuda:0")


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def synthetic(T, n, d=1, seed=0):
    rs = np.random.RandomState(seed)
    x = rs.uniform(-2, 2, size=(T, n, d)).astype(np.float32)
    y = (np.sin(2 * x[..., 0]) + 0.1 * rs.normal(size=(T, n))).astype(np.float32)
    return x, y


def prior_particles(lay, P, seed):
    mu, sigma = orc.hyper_prior_params(lay, 0.5, 3.0)
    g = torch.Generator().manual_seed(seed)
    return (mu + sigma * torch.randn(P, lay.D, generator=g)).numpy()


def factors(e, arch, lay, P, T, n, theta, x, y, idx):
    """Reads L, U, v, alphahat of every matrix back and compares with fp64."""
    out = (ctypes.c_int64 * 14)()
    a = arch.c_struct()
    _lib.check(_lib.lib.pacoh_debug_big_layout(ctypes.byref(a), P, T, n, out))
    off, nb, npad, batch = out[0], out[1], out[2], out[3]
    if nb == 0:
        print("   (small-matrix kernel: no factors to inspect)")
        return
    ws = e._workspace(P, T)
    base = ws[off:]

    def view(o, count, dtype=torch.float32):
        return base[o:o + 4 * count].view(dtype)

    Lb = view(out[4], batch * npad * npad).view(batch, npad, npad).cpu().double()
    SB = view(out[5], batch * nb * 2 * 128 * 128).view(batch, nb, 2, 128, 128).cpu().double()
    vb = view(out[8], batch * npad).view(batch, npad).cpu().double()
    ab = view(out[9], batch * npad).view(batch, npad).cpu().double()
    th64 = torch.from_numpy(theta).double()
    worst = dict(L=0.0, U=0.0, Linv=0.0, v=0.0, alpha=0.0)
    for p in range(P):
        for t in range(min(T, 3)):
            m = p * T + t
            if m >= batch:
                continue
            xt, yt = torch.from_numpy(x[idx[t]]).double(), torch.from_numpy(y[idx[t]]).double()
            mean, z, ls, noise, osc = orc.gp_components(th64[p:p + 1], lay, xt)
            K = orc.se_gram(z, ls, outputscale=osc)[0]
            s = float(osc[0]) if osc is not None else 1.0
            tot = s + float(noise[0])
            Kh = K / tot
            Kh[range(n), range(n)] = 1.0
            Kp = torch.eye(npad, dtype=torch.float64)
            Kp[:n, :n] = Kh
            L = torch.linalg.cholesky(Kp)
            U = torch.linalg.inv(L).T
            r = torch.zeros(npad, dtype=torch.float64)
            r[:n] = yt - mean[0]
            v = torch.linalg.solve_triangular(L, r[:, None], upper=False)[:, 0]
            alpha = U @ v
            Lg = torch.tril(Lb[m])
            # block-lower part only: diagonal tiles hold L (zeros above the diagonal), strictly upper block tiles hold U
            Lref = L.clone()
            blk = torch.arange(npad) // 128
            lower = blk[:, None] >= blk[None, :]
            eL = (Lb[m] * lower - Lref).abs().max().item() / Lref.abs().max().item()
            Ug = Lb[m] * (~lower)
            for k in range(nb):
                Ug[k * 128:(k + 1) * 128, k * 128:(k + 1) * 128] = SB[m, k, 1]
            eU = (Ug - U).abs().max().item() / U.abs().max().item()
            Linv = torch.linalg.inv(L)
            eI = max((SB[m, k, 0] - torch.linalg.inv(L[k * 128:(k + 1) * 128, k * 128:(k + 1) * 128])).abs().max().item() for k in range(nb)) / Linv.abs().max().item()
            ev = (vb[m] - v).abs().max().item() / v.abs().max().item()
            ea = (ab[m] - alpha).abs().max().item() / alpha.abs().max().item()
            for kname, val in zip(("L", "U", "Linv", "v", "alpha"), (eL, eU, eI, ev, ea)):
                worst[kname] = max(worst[kname], val if np.isfinite(val) else 1e30)
            if m == 0:
                print("   matrix 0: cond(Khat) %.2e  tot %.4f  per-tile L err:" % (float(torch.linalg.cond(Kp)), tot))
                for i in range(nb):
                    print("      " + " ".join("%.1e" % ((Lb[m] * lower - Lref)[i * 128:(i + 1) * 128, j * 128:(j + 1) * 128].abs().max().item())
                                              for j in range(i + 1)))
                if nb > 1:
                    print("   matrix 0: per-tile U err:")
                    for i in range(nb):
                        print("      " + " ".join("%.1e" % ((Ug - U)[i * 128:(i + 1) * 128, j * 128:(j + 1) * 128].abs().max().item())
                                                  for j in range(nb)))
    print("   factors (max rel err over inspected matrices): " + "  ".join("%s %.2e" % kv for kv in worst.items()))


def run_case(n, P=2, T=4, kw=None, seed=0, inspect=True):
    kw = kw or dict(input_dim=1)
    lay, arch = orc.Layout(**kw), eng.GPArch(**kw)
    x, y = synthetic(5, n, d=kw["input_dim"], seed=seed + n)
    theta = prior_particles(lay, P, 100 + n)
    idx = [4, 0, 0, 3, 1, 4, 2][:T] if T <= 7 else list(np.random.RandomState(1).randint(0, 5, size=T))
    e = eng.MetaMLLEngine(arch, x, y, dev)
    th = torch.from_numpy(theta).to(dev)
    tidx = torch.from_numpy(np.asarray(idx, dtype=np.int32)).to(dev)
    t0 = time.time()
    mll, packed, info = e.mll_fwd_bwd(th, tidx)
    torch.cuda.synchronize()
    t1 = time.time()
    mll2, packed2, _ = e.mll_fwd_bwd(th, tidx)
    torch.cuda.synchronize()
    t2 = time.time()
    print("== n=%d P=%d T=%d kw=%s: first call %.1f ms, second %.1f ms, info[min,max]=%d,%d, repeatable=%s" %
          (n, P, T, kw, 1e3 * (t1 - t0), 1e3 * (t2 - t1), int(info.min()), int(info.max()), bool(torch.equal(packed, packed2))))
    if inspect:
        try:
            factors(e, arch, lay, P, len(idx), n, theta, x, y, idx)
        except Exception as ex:  # noqa: BLE001
            print("   factor inspection failed:", repr(ex))
    tasks64 = [(torch.from_numpy(x[i]).double(), torch.from_numpy(y[i]).double()) for i in range(x.shape[0])]
    mu, sigma = orc.hyper_prior_params(lay, 0.5, 3.0, torch.float64)
    logp64, g64, mll64 = orc.meta_log_prob_and_grad(torch.from_numpy(theta).double(), lay, [tasks64[i] for i in idx], 0.01, mu, sigma)
    pmu, psig = arch.hyper_prior(0.5, 3.0)
    pre = eng.pre_factor([n] * len(idx))
    logp, dth = eng.logprob_finalize(th, pmu.to(dev), psig.to(dev), 0.01, pre, packed)
    g, g64 = dth.cpu().numpy(), g64.numpy()
    print("   mll rel err %.3e  logp rel err %.3e  score rel err %.3e (global)" %
          (rel(mll.cpu().numpy(), mll64.numpy()), rel(logp.cpu().numpy(), logp64.numpy()), rel(g, g64)))
    scale = np.abs(g64).max()
    worst = 0.0
    for nm, (a, b) in arch.entries().items():
        err = np.abs(g[:, a:b] - g64[:, a:b]).max() / max(np.abs(g64[:, a:b]).max(), 1e-3 * scale)
        worst = max(worst, err)
        print("      %-24s grp-rel %.3e   |ref|max %.3e" % (nm, err, np.abs(g64[:, a:b]).max()))
    print("   WORST group error %.3e  %s" % (worst, "OK" if worst <= 1e-4 else "** ABOVE 1e-4 **"))


def main():
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_properties(0).multi_processor_count, "SMs")
    ns = [int(a) for a in sys.argv[1:]] or [100, 128, 200, 300, 512]
    for n in ns:
        try:
            run_case(n)
        except Exception as ex:  # noqa: BLE001
            print("== n=%d FAILED: %r" % (n, ex))
    if not sys.argv[1:]:
        run_case(130, P=1, T=3, kw=dict(input_dim=1, outputscale=True, noise_floor=1e-3))
        run_case(260, P=2, T=3, kw=dict(input_dim=3, mean_layers=(32, 32), kernel_layers=(32, 32), feature_dim=4))




# ---------------------------------------------------------------------------------------------------------------------
# BASELINE config #5 shapes: PACOH-MAP (outputscale, noise floor 1e-3), torch nn.Linear default init, raw hypers 0.
def map_theta(lay, seed=30):
    g = torch.Generator().manual_seed(seed)
    th = torch.zeros(1, lay.D)
    for name, (a, b) in lay.entries.items():
        if "weight" in name or "bias" in name:
            # fan-in of the layer: weight (out, in) -> in; bias uses the same bound (torch.nn.Linear.reset_parameters)
            pref = name.rsplit(".", 1)[0]
            wa, wb = lay.entries[pref + ".weight"]
            ba, bb = lay.entries[pref + ".bias"]
            fan_in = (wb - wa) // (bb - ba)
            bound = 1.0 / np.sqrt(fan_in)
            th[0, a:b] = (torch.rand(b - a, generator=g) * 2 - 1) * bound
    return th.numpy()


def config5(n, T=4, T_time=0, fp32_ref=True):
    kw = dict(input_dim=1, outputscale=True, noise_floor=1e-3)
    lay, arch = orc.Layout(**kw), eng.GPArch(**kw)
    train = orc.sinusoid_tasks(max(T, 4), n, seed=26)
    stats = orc.normalization_stats(train)
    tasks64 = [orc.prepare_task(xx, yy, stats, torch.float64) for xx, yy in train]
    x = np.stack([t[0].numpy() for t in tasks64]).astype(np.float32)
    y = np.stack([t[1].numpy() for t in tasks64]).astype(np.float32)
    theta = map_theta(lay)
    idx = list(range(T))
    e = eng.MetaMLLEngine(arch, x, y, dev)
    th = torch.from_numpy(theta).to(dev)
    tidx = torch.from_numpy(np.asarray(idx, dtype=np.int32)).to(dev)
    mll, packed, info = e.mll_fwd_bwd(th, tidx)
    torch.cuda.synchronize()
    tasks = [(torch.from_numpy(x[i]).double(), torch.from_numpy(y[i]).double()) for i in idx]
    mu, sigma = orc.hyper_prior_params(lay, 0.5, 3.0, torch.float64)
    t0 = time.time()
    logp64, g64, mll64 = orc.meta_log_prob_and_grad(torch.from_numpy(theta).double(), lay, tasks, 0.0, mu, sigma)
    t_or = time.time() - t0
    g64 = g64.numpy() / eng.pre_factor([n] * T)        # d sum_t mll / d theta
    g = packed[:lay.D].cpu().numpy()[None, :]
    print("== config5 n=%d T=%d: info[min,max]=%d,%d  mll rel err %.3e  (oracle fp64 %.1f s)" %
          (n, T, int(info.min()), int(info.max()), rel(mll.cpu().numpy(), mll64.numpy()), t_or))
    scale = np.abs(g64).max()
    worst = 0.0
    for nm, (a, b) in arch.entries().items():
        err = np.abs(g[:, a:b] - g64[:, a:b]).max() / max(np.abs(g64[:, a:b]).max(), 1e-3 * scale)
        worst = max(worst, err)
        print("      %-24s grp-rel %.3e   |ref|max %.3e" % (nm, err, np.abs(g64[:, a:b]).max()))
    print("   WORST group error %.3e  %s" % (worst, "OK" if worst <= 1e-4 else "** ABOVE 1e-4 **"))
    if fp32_ref:
        tasks32 = [(a.float(), b.float()) for a, b in tasks]
        mu32, s32 = orc.hyper_prior_params(lay, 0.5, 3.0, torch.float32)
        _, g32, mll32 = orc.meta_log_prob_and_grad(torch.from_numpy(theta).float(), lay, tasks32, 0.0, mu32, s32)
        g32 = g32.numpy() / eng.pre_factor([n] * T)
        w32 = max(np.abs(g32[:, a:b] - g64[:, a:b]).max() / max(np.abs(g64[:, a:b]).max(), 1e-3 * scale) for a, b in arch.entries().values())
        print("   fp32 torch oracle (the reference's own precision) vs fp64: mll %.3e  worst group %.3e" % (rel(mll32.numpy(), mll64.numpy()), w32))
    if T_time:
        train = orc.sinusoid_tasks(T_time, n, seed=27)
        xs = np.stack([orc.prepare_task(xx, yy, stats, torch.float32)[0].numpy() for xx, yy in train])
        ys = np.stack([orc.prepare_task(xx, yy, stats, torch.float32)[1].numpy() for xx, yy in train])
        e2 = eng.MetaMLLEngine(arch, xs, ys, dev)
        tidx2 = torch.arange(T_time, dtype=torch.int32, device=dev)
        e2.mll_fwd_bwd(th, tidx2)
        torch.cuda.synchronize()
        with eng.StageTiming() as stg:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            reps = 3
            for _ in range(reps):
                _, _, inf2 = e2.mll_fwd_bwd(th, tidx2)
            ev1.record()
            ev1.synchronize()
            ms, calls = stg.read()
        tot = ev0.elapsed_time(ev1) / reps
        flop = T_time * (n ** 3 + 2.0 * n * n)
        print("   timing n=%d T=%d: %.2f ms per fwd+bwd (stages per call: %s)  GP stage %.1f TFLOP/s (n^3 + 2 n^2)  info max %d" %
              (n, T_time, tot, {k: round(v / max(calls, 1), 3) for k, v in ms.items()}, flop / (ms["gp_mll"] / max(calls, 1) * 1e-3) / 1e12,
               int(inf2.max())))


if len(sys.argv) > 1 and sys.argv[1] == "config5":
    for n, tt in ((512, 1024), (1024, 1024), (2048, 1024)):
        try:
            config5(n, T=3, T_time=tt)
        except Exception as ex:  # noqa: BLE001
            import traceback
            traceback.print_exc()
    sys.exit(0)


if __name__ == "__main__":
    main()
