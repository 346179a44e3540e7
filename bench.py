#!/usr/bin/env python
"""Benchmark of the PACOH meta-training hot path (BASELINE.json metric: task-particle MLL+grad evals/s).

    python bench.py --gpus N --steps K --warmup W            # this engine (one rank per GPU under torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's loop structure on the host CPU (oracle port)

Workload (config.workload): BASELINE configs[3], the configuration the metric and the north-star target are quoted
on -- PACOH-SVGD, 64 particles x 4096 synthetic sinusoid tasks x 50 points, (32,32) mean and kernel nets, F = 2; it fits
one GPU.  One "step" = one full SVGD meta-training step on one sampled batch of T = 4096 tasks (with replacement):
batched MLL forward+backward for all 64 x 4096 (particle, task) pairs, hyper-prior, SVGD direction with the median
heuristic, Adam update.  For N > 1 the same global batch is task-sharded over the ranks (strong scaling) with one cross-rank
all-reduce of the packed (P, D+1) gradient buffer per step.  One eval = one (particle, task) MLL value + its gradient.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

P, T, N_PTS, D_IN, HID, FEAT = 64, 4096, 50, 1, (32, 32), 2
CPU_SLICE_T = 64                      # tasks per CPU-baseline step (bounded sample of the same workload)
METRIC = "task-particle MLL+grad evals/s"
UNIT = "evals/s"


def mac(out_dim):
    m, prev = 0, D_IN
    for h in HID:
        m += prev * h
        prev = h
    return m + prev * out_dim


def flops_per_eval(n=N_PTS, F=FEAT):
    """F_eval(n) = 6n(MAC(1)+MAC(F)) + n^2(7F+6) + n^3 + 2n^2  (842.4 kFLOP at n=50)."""
    return 6 * n * (mac(1) + mac(F)) + n * n * (7 * F + 6) + n ** 3 + 2 * n * n


KERNEL_FLOPS = {   # algorithmic FLOPs per eval attributed to each kernel (sums to flops_per_eval)
    "mlp_fwd": lambda: 2 * N_PTS * (mac(1) + mac(FEAT)),
    "gp_mll": lambda: N_PTS ** 2 * (7 * FEAT + 6) + N_PTS ** 3 + 2 * N_PTS ** 2,
    "mlp_bwd": lambda: 4 * N_PTS * (mac(1) + mac(FEAT)),
}


def config_dict(n_gpus):
    return {"workload": "PACOH-SVGD meta-training step: 64 particles x 4096 sinusoid tasks x 50 points "
                        "(BASELINE configs[3]); mean/kernel MLP (32,32), F=2, D=2342; task batch sampled with replacement",
            "particles": P, "tasks_per_step": T, "points_per_task": N_PTS, "evals_per_step": P * T,
            "flops_per_eval": flops_per_eval(), "parallelism": "task-sharded x%d, 1 all-reduce/step" % n_gpus if n_gpus > 1 else "single GPU",
            "l2": "per-step working set (intermediate mean/feature/gradient buffers) is ~0.33 GB per GPU at N=1, larger than "
                  "the 126 MB L2; every step streams a freshly sampled batch (no flush needed)"}


def make_data():
    from meta_learning_pacoh_b200.data_sim import SinusoidDataset
    ds = SinusoidDataset(random_state=np.random.RandomState(26))
    return ds.generate_meta_train_data(n_tasks=T, n_samples=N_PTS)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU baseline (oracle port)
def cpu_reference_steps(steps, warmup, data):
    """The reference's loop structure on the host CPU (oracle/pacoh_oracle.py::SVGDOracle: per-task Python loop with
    P-batched torch ops, autograd score, SVGD phi, Adam -- random_gp.py:206-222, svgd.py:12-28), all host threads, on a
    bounded sample of the workload: the same 64 particles and 50-point tasks, CPU_SLICE_T tasks per step."""
    import torch
    from oracle import pacoh_oracle as orc
    torch.set_num_threads(os.cpu_count() or 1)
    stats = orc.normalization_stats(data)
    tasks = [orc.prepare_task(x, y, stats) for x, y in data[:CPU_SLICE_T]]
    lay = orc.Layout(D_IN)
    mu, sigma = orc.hyper_prior_params(lay, 0.5, 3.0)
    g = torch.Generator().manual_seed(30)
    particles = mu + sigma * torch.randn(P, lay.D, generator=g)
    s = orc.SVGDOracle(tasks, lay, particles, seed=30)
    for _ in range(warmup):
        s.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        s.step()
    dt = time.perf_counter() - t0
    return P * CPU_SLICE_T * steps / dt, dt / steps, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    data = make_data()[:CPU_SLICE_T]
    steps = max(1, args.steps)
    warm = min(args.warmup, 2)
    value, sec_per_step, cores = cpu_reference_steps(steps, warm, data)
    sample = "%d-task slice of the 4096-task batch per step (64 particles x 50 points), %d steps" % (CPU_SLICE_T, steps)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": sec_per_step * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config_dict(args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference loop structure (per-task Python loop, P-batched torch ops, autograd) restated in oracle/; "
                    "gpytorch/pyro are not installable offline. ms_per_step is for the %d-task slice." % CPU_SLICE_T}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ this engine
def run_ours(args):
    import torch
    import torch.distributed as dist
    from meta_learning_pacoh_b200 import engine as eng
    from meta_learning_pacoh_b200.meta_learn import GPRegressionMetaLearnedSVGD

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, "launch with torchrun --nproc-per-node %d (WORLD_SIZE=%d)" % (args.gpus, world)

    data = make_data()
    model = GPRegressionMetaLearnedSVGD(data, num_particles=P, random_seed=30)
    if world > 1:
        model.shard_tasks()
    collective = ("none (single GPU)" if world == 1 else
                  "all-reduce fused into the finalize kernel over NVLink peer memory (pacoh_peer_allreduce_finalize)" if model._peer is not None
                  else "NCCL all-reduce of the packed (P, D+1) buffer (peer memory unavailable: %s)" % getattr(model, "_peer_error", None))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    # ---- device-resident timing (value)
    step_dev = lambda: model.svgd_step(model._sample_task_indices())   # noqa: E731
    for _ in range(args.warmup):
        step_dev()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    with eng.StageTiming() as stg:
        ms_total = timed(step_dev, args.steps)
        stage_ms, stage_calls = stg.read()
    clock_info = clocks.stop() if rank == 0 else None
    eng.check_info(model._last_info)
    ms_per_step = ms_total / args.steps
    value = P * T / (ms_per_step * 1e-3)

    # ---- end-to-end through the public API with host buffers (e2e)
    Xh = model.engine.x.cpu().pin_memory()
    Yh = model.engine.y.cpu().pin_memory()
    # pipelined like a real input pipeline: two pinned staging sets; while step k runs on the device the host samples and
    # gathers batch k+1, then reads step k's logp (device->host copy issued every step, waited for one step later)
    lo_e, hi_e = eng.shard_bounds(T, rank, world)
    xb = [torch.empty((hi_e - lo_e,) + tuple(Xh.shape[1:]), dtype=Xh.dtype).pin_memory() for _ in range(2)]
    yb = [torch.empty((hi_e - lo_e,) + tuple(Yh.shape[1:]), dtype=Yh.dtype).pin_memory() for _ in range(2)]
    pending = [None, None]
    e2e_count = [0]
    logp_sink = [0.0]

    def step_e2e():
        k = e2e_count[0] & 1
        if pending[k] is not None:                      # staging set k was last used two steps ago: its step must be done
            out, ev = pending[k]
            ev.synchronize()
            logp_sink[0] += float(out[0])               # the host really reads the result
        idx = torch.from_numpy(model._sample_task_indices()[lo_e:hi_e])
        torch.index_select(Xh, 0, idx, out=xb[k])       # host-side gather of this rank's shard of the sampled batch (the
        torch.index_select(Yh, 0, idx, out=yb[k])       # reference passes the sampled task tensors themselves, GPR_meta_svgd.py:102-103)
        pending[k] = model.svgd_step_host(xb[k], yb[k], global_tasks=T, wait=False)   # H2D, full step, D2H of logp
        e2e_count[0] += 1

    def drain_e2e():
        for k in range(2):
            if pending[k] is not None:
                out, ev = pending[k]
                ev.synchronize()
                logp_sink[0] += float(out[0])
                pending[k] = None

    for _ in range(max(3, args.warmup // 2)):
        step_e2e()
    drain_e2e()
    e2e_steps = args.steps
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    drain_e2e()
    barrier()
    e2e_ms = torch.tensor([(time.perf_counter() - t0) * 1e3], device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = P * T / (e2e_ms.item() / e2e_steps * 1e-3)
    h2d = (T // world) * N_PTS * (D_IN + 1) * 4
    d2h = P * 4

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (live CUDA-event stage timing inside the timed region)
    ffma_peak = eng.ffma_peak_tflops()
    nominal = 148 * 128 * 2 * 1.965e9 / 1e12
    evals_rank = P * (T // world)
    kernels = {}
    for name, fl in KERNEL_FLOPS.items():
        ms = stage_ms[name] / max(stage_calls, 1)
        tf = fl() * evals_rank / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
        kernels[name] = {"ms_per_launch": ms, "algorithmic_gflop_per_launch": fl() * evals_rank / 1e9, "tflops": tf,
                         "frac_of_fp32_peak": tf / ffma_peak, "share_of_step": ms / ms_per_step}
    kernels["reduce"] = {"ms_per_launch": stage_ms["reduce"] / max(stage_calls, 1), "share_of_step": stage_ms["reduce"] / max(stage_calls, 1) / ms_per_step}
    dom = max(KERNEL_FLOPS, key=lambda k: kernels[k]["ms_per_launch"])
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(dom)
        except Exception:
            traffic = None
    roofline = {"bound": "fp32", "kernel": dom, "achieved": kernels[dom]["tflops"], "peak": ffma_peak, "unit": "TFLOP/s",
                "frac": kernels[dom]["tflops"] / ffma_peak, "traffic": traffic,
                "peak_source": "FFMA micro-benchmark run in this process (pacoh_ffma_peak_launch); MEASURED_PEAKS.json holds no FP32 "
                               "figure; nominal 148 SM x 128 lanes x 2 x 1.965 GHz = %.1f TFLOP/s" % nominal,
                "bound_note": "FP32-compute bound (arithmetic intensity ~5e4 FLOP/B, DRAM ~1% busy). All three kernels run their "
                              "matrix work on tcgen05 in 3xTF32 (fp32-accurate): the MLP hidden-layer and weight-gradient "
                              "contractions and the rank-4 Gauss-Jordan updates of the GP kernel; tanh / exp2 / the 4x4 pivot-block "
                              "inverses / row sums run on the CUDA cores. Fractions are ALGORITHMIC fp32 FLOPs over the measured FP32 "
                              "FFMA peak (the honest denominator for an fp32-parity path), so a kernel can exceed what FFMA alone "
                              "could reach",
                "step_achieved_tflops": value * flops_per_eval() / 1e12, "step_frac_of_fp32_peak": value * flops_per_eval() / 1e12 / ffma_peak,
                "kernels": kernels}
    # the driver-measured peaks (HBM copy bandwidth, dense bf16 tensor throughput) for the same kernel, for the record:
    # neither bounds this path (see bound_note), which is why the FP32 FFMA peak is the denominator above
    mp_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(mp_path):
        try:
            mp = json.load(open(mp_path))
            dom_ms = kernels[dom]["ms_per_launch"]
            vs = {"source": "MEASURED_PEAKS.json"}
            if traffic and mp.get("hbm_gbs"):
                gbs = traffic / (dom_ms * 1e-3) / 1e9
                vs["hbm"] = {"achieved_gbs": gbs, "peak_gbs": mp["hbm_gbs"], "frac": gbs / mp["hbm_gbs"]}
            tpk = mp.get("bf16_tflops_sustained") or mp.get("bf16_tflops")
            if tpk:
                vs["tensor_bf16"] = {"achieved_algorithmic_tflops": kernels[dom]["tflops"], "peak_tflops": tpk,
                                     "frac": kernels[dom]["tflops"] / tpk,
                                     "note": "fp32 parity needs 3xTF32: 3 passes at half the bf16 rate, i.e. 6 bf16-equivalent "
                                             "tensor FLOPs per algorithmic FLOP of the GEMM part"}
            roofline["vs_measured_peaks"] = vs
        except Exception:
            pass

    # ---- CPU baseline (oracle port) on the host cores, bounded sample, N = 1 only
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, sec, cores = cpu_reference_steps(steps=8, warmup=1, data=data[:CPU_SLICE_T])
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d-task slice of the 4096-task batch per step (64 particles x 50 points), 8 steps, %.2f s/step" % (CPU_SLICE_T, sec)}

    launches_per_step = 3 + 3 + 1 + 3 + 1      # mlp_fwd, gp_mll, mlp_bwd | 2 partial reductions + hyp reduction | finalize | svgd x3 | adam
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": dict(config_dict(world), collective=collective),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms.item() / e2e_steps, "timing": "host wall clock around K steps (2-deep pipelined: batch k+1 is gathered on the host while step k runs; every step copies its batch H2D from pinned memory and its logp D2H), barrier + synchronize on both sides"},
            "gpu_launches": launches_per_step * args.steps, "clocks": clock_info, "roofline": roofline, "cpu_baseline": cpu}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
