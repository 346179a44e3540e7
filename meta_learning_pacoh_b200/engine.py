"""Host side of the B200 engine: architecture descriptors, device workspaces and the autograd.Functions that
wrap the C-ABI entry points of libpacoh_b200 (include/pacoh_b200.h).

PyTorch is used here for device memory, streams and torch.distributed only; every arithmetic step of the hot path
runs in the hand-written kernels under csrc/.  There is no CPU / eager fallback.

Reference seams replaced (paths relative to the reference root):
  RandomGPMeta.log_prob + autograd.grad      meta_learn/random_gp.py:206-222, meta_learn/svgd.py:13-16
  SVGD.phi tail + RBF_Kernel                 meta_learn/svgd.py:18-21, 32-59
  RandomGPPosterior.rsample / log_prob       meta_learn/random_gp.py:253-263, meta_learn/GPR_meta_vi.py:216-224
"""
import ctypes
import math
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Tuple

import numpy as np
import torch

from . import _lib
from ._lib import lib, check


class NotPSDError(RuntimeError):
    """Raised when a task's kernel matrix is not positive definite even after the jitter ladder
    (same role as gpytorch.utils.errors.NotPSDError on the reference path)."""


_MEAN = {"zero": _lib.MEAN_ZERO, "constant": _lib.MEAN_CONSTANT, "NN": _lib.MEAN_NN}
_COVAR = {"SE": _lib.COVAR_SE, "NN": _lib.COVAR_NN}


@dataclass(frozen=True)
class GPArch:
    """Architecture of the vectorised GP prior (VectorizedGP.__init__, meta_learn/random_gp.py:22-52).

    ``outputscale`` / ``noise_floor`` select the PACOH-MAP variant (gpytorch ScaleKernel and
    GaussianLikelihood(noise_constraint=GreaterThan(1e-3)), meta_learn/GPR_meta_mll.py:54-55, 218).
    """
    input_dim: int
    mean_kind: str = "NN"
    covar_kind: str = "NN"
    mean_layers: Tuple[int, ...] = (32, 32)
    kernel_layers: Tuple[int, ...] = (32, 32)
    feature_dim: int = 2
    outputscale: bool = False
    noise_floor: float = 0.0
    _cache: dict = field(default_factory=dict, compare=False, repr=False, hash=False)

    def c_struct(self) -> _lib.PacohArch:
        a = _lib.PacohArch()
        a.input_dim = self.input_dim
        a.mean_kind = _MEAN[self.mean_kind]
        a.covar_kind = _COVAR[self.covar_kind]
        assert len(self.mean_layers) <= _lib.PACOH_MAX_LAYERS and len(self.kernel_layers) <= _lib.PACOH_MAX_LAYERS
        a.n_mean_layers = len(self.mean_layers)
        for i, w in enumerate(self.mean_layers):
            a.mean_layers[i] = int(w)
        a.n_kernel_layers = len(self.kernel_layers)
        for i, w in enumerate(self.kernel_layers):
            a.kernel_layers[i] = int(w)
        a.feature_dim = int(self.feature_dim)
        a.has_outputscale = 1 if self.outputscale else 0
        a.noise_floor = float(self.noise_floor)
        return a

    @property
    def num_features(self) -> int:
        return self.feature_dim if self.covar_kind == "NN" else self.input_dim

    @property
    def D(self) -> int:
        """Flat parameter count, computed by the library (pacoh_param_count)."""
        if "D" not in self._cache:
            a = self.c_struct()
            self._cache["D"] = int(check(lib.pacoh_param_count(ctypes.byref(a))))
        return self._cache["D"]

    def entries(self) -> "OrderedDict[str, Tuple[int, int]]":
        """name -> (start, end) in the flat vector; names and order follow RandomGPMeta.parameter_shapes()
        (meta_learn/random_gp.py:186-190, models.py:351-383, 319-323)."""
        out, off = OrderedDict(), 0

        def add(name, size):
            nonlocal off
            out[name] = (off, off + size)
            off += size

        def mlp(prefix, widths, out_dim):
            prev = self.input_dim
            for i, w in enumerate(widths):
                add("%s.fc_%d.bias" % (prefix, i + 1), w)
                add("%s.fc_%d.weight" % (prefix, i + 1), w * prev)
                prev = w
            add("%s.out.bias" % prefix, out_dim)
            add("%s.out.weight" % prefix, out_dim * prev)

        if self.mean_kind == "NN":
            mlp("mean_nn", self.mean_layers, 1)
        elif self.mean_kind == "constant":
            add("constant_mean", 1)
        if self.covar_kind == "NN":
            mlp("kernel_nn", self.kernel_layers, self.feature_dim)
        add("lengthscale_raw", self.num_features)
        add("noise_raw", 1)
        if self.outputscale:
            add("outputscale_raw", 1)
        assert off == self.D, (off, self.D)
        return out

    def hyper_prior(self, weight_prior_std=0.5, bias_prior_std=3.0):
        """(mu, sigma) float32 CPU tensors of the factorised Gaussian hyper-prior (random_gp.py:118-157)."""
        mu = np.empty(self.D, dtype=np.float32)
        sigma = np.empty(self.D, dtype=np.float32)
        a = self.c_struct()
        check(lib.pacoh_hyper_prior_params(ctypes.byref(a), float(weight_prior_std), float(bias_prior_std),
                                           mu.ctypes.data_as(ctypes.c_void_p), sigma.ctypes.data_as(ctypes.c_void_p)))
        return torch.from_numpy(mu), torch.from_numpy(sigma)


def pre_factor(task_sizes) -> float:
    """n_h / (n_h + T): harmonic-mean task size over the sampled batch (meta_learn/random_gp.py:209-212)."""
    sizes = np.asarray(task_sizes, dtype=np.float64)
    hm = 1.0 / np.mean(1.0 / sizes)
    return float(hm / (hm + len(sizes)))


def shard_bounds(T, rank, world):
    """Contiguous slice [lo, hi) of a length-T sampled task batch owned by `rank` (the batch index list,
    duplicates included, is split into `world` contiguous chunks)."""
    return (rank * T) // world, ((rank + 1) * T) // world


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32c(t, device):
    assert t.dtype == torch.float32 and t.is_contiguous() and t.device == device, "expected a contiguous fp32 tensor on %s" % device
    return t


class MetaMLLEngine:
    """Device-resident task set + workspaces for the batched marginal log-likelihood (pacoh_meta_mll_fwd_bwd)."""

    def __init__(self, arch: GPArch, x, y, device=None, task_n=None):
        """x (T_total, n, d), y (T_total, n): normalised task data (abstract.py:224-258), uploaded once.
        ``task_n`` (T_total,) ints: ragged task sets -- x / y are padded to n = max_t n_t rows and task t only uses its
        first task_n[t] rows (pacoh_meta_mll_fwd_bwd_ragged); None: every task has n points."""
        self.arch = arch
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.type != "cuda":
            raise RuntimeError("the PACOH engine runs on CUDA devices only (no CPU fallback)")
        x = torch.as_tensor(x, dtype=torch.float32)
        y = torch.as_tensor(y, dtype=torch.float32)
        assert x.ndim == 3 and y.ndim == 2 and x.shape[:2] == y.shape and x.shape[2] == arch.input_dim
        self.x = x.contiguous().to(self.device)
        self.y = y.contiguous().to(self.device)
        self.T_total, self.n, self.d = x.shape
        self.task_n = None
        if task_n is not None:
            tn = torch.as_tensor(task_n, dtype=torch.int32).reshape(-1)
            assert tn.numel() == self.T_total and int(tn.min()) >= 1 and int(tn.max()) <= self.n
            if int(tn.min()) < self.n:
                self.task_n = tn.contiguous().to(self.device)
        self._c_arch = arch.c_struct()
        self._ws = {}

    def _workspace(self, P, T):
        key = (P, T)
        if key not in self._ws:
            nbytes = check(lib.pacoh_workspace_bytes(ctypes.byref(self._c_arch), P, T, self.n))
            self._ws[key] = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=self.device)
        return self._ws[key]

    def mll_fwd_bwd(self, theta, task_idx, want_mll=True, want_info=True, out=None):
        """theta (P, D) fp32 cuda; task_idx (T,) int32 cuda (batch order, may repeat).

        Returns (mll (P, T) | None, packed, info | None) where ``packed`` is a flat fp32 buffer of P*D + P
        floats: d mll_sum / d theta (P, D) followed by mll_sum (P) -- one buffer so that a task-sharded run
        needs a single all-reduce.
        """
        theta = _f32c(theta, self.device)
        assert task_idx.dtype == torch.int32 and task_idx.is_contiguous() and task_idx.device == self.device
        P, D = theta.shape
        assert D == self.arch.D, "theta has %d columns, architecture needs %d" % (D, self.arch.D)
        T = task_idx.numel()
        ws = self._workspace(P, T)
        packed = out if out is not None else torch.empty(P * D + P, dtype=torch.float32, device=self.device)
        mll = torch.empty(P, T, dtype=torch.float32, device=self.device) if want_mll else None
        info = torch.empty(P, T, dtype=torch.int32, device=self.device) if want_info else None
        dth_ptr = ctypes.c_void_p(packed.data_ptr())
        sum_ptr = ctypes.c_void_p(packed.data_ptr() + 4 * P * D)
        check(lib.pacoh_meta_mll_fwd_bwd_ragged(ctypes.byref(self._c_arch), P, T, self.n, _ptr(theta), _ptr(self.x), _ptr(self.y),
                                                _ptr(self.task_n), _ptr(task_idx), _ptr(mll), sum_ptr, dth_ptr, _ptr(info),
                                                _ptr(ws), ws.numel(), _stream()))
        return mll, packed, info


class PinnedRing:
    """Host->device upload of small per-step arrays (the sampled task indices) without a host synchronisation and
    without the reuse race of a single pinned buffer: a ring of pinned staging buffers, each guarded by a CUDA event
    recorded after its copy.  A slot is only rewritten once its previous copy has executed (event.synchronize() is a no-op
    unless the host has run more than `slots` steps ahead of the device)."""

    def __init__(self, device, slots=8, dtype=torch.int32):
        self.device, self.dtype, self.slots = torch.device(device), dtype, slots
        self._buf = [None] * slots
        self._ev = [None] * slots
        self._next = 0

    def upload(self, array, out=None):
        """Returns a device tensor holding `array` (or fills the given device tensor `out`, e.g. the fixed buffer a
        captured CUDA graph reads); the copy is asynchronous on the current stream."""
        src = torch.from_numpy(np.ascontiguousarray(array)).reshape(-1)
        assert src.dtype == self.dtype
        k = self._next
        self._next = (k + 1) % self.slots
        if self._ev[k] is not None:
            self._ev[k].synchronize()
        if self._buf[k] is None or self._buf[k].numel() < src.numel():
            self._buf[k] = torch.empty(max(src.numel(), 16), dtype=self.dtype).pin_memory()
            self._ev[k] = torch.cuda.Event()
        self._buf[k][:src.numel()].copy_(src)
        if out is None:
            out = self._buf[k][:src.numel()].to(self.device, non_blocking=True)
        else:
            out.view(-1)[:src.numel()].copy_(self._buf[k][:src.numel()], non_blocking=True)
        self._ev[k].record(torch.cuda.current_stream(self.device))
        return out


class StepState:
    """Device-side state of a training loop (pacoh_step_prepare): number of completed steps, current learning rate
    (StepLR) and Adam's bias corrections.  Every kernel of a step reads these from the device, so the step sequence can
    be captured in a CUDA graph and replayed without the host touching it; `steps` mirrors the counter on the host."""

    def __init__(self, device, lr, lr_decay=1.0, decay_every=1000, betas=(0.9, 0.999)):
        self.device = torch.device(device)
        self.buf = torch.zeros(8, dtype=torch.int32, device=self.device)
        self.lr, self.gamma, self.decay_every, self.betas = float(lr), float(lr_decay), int(decay_every), betas
        self.steps = 0                                   # host mirror of buf[0]

    def prepare(self, K=0, T=0, idx_stream=None, idx_out=None, fstream=None, fout=None):
        F = fout.numel() if fout is not None else 0
        check(lib.pacoh_step_prepare(_ptr(self.buf), int(K), int(T), _ptr(idx_stream), _ptr(idx_out), int(F), _ptr(fstream), _ptr(fout),
                                     self.lr, self.gamma, self.decay_every if self.gamma < 1.0 else 0, float(self.betas[0]),
                                     float(self.betas[1]), _stream()))
        self.steps += 1


class StepGraph:
    """`k` consecutive training steps captured in one CUDA graph ("CUDA-graph capture of the whole step").
    `step_fn` must be capture-safe: no host synchronisation, everything step-dependent read from device memory
    (StepState, pre-uploaded index streams).  It is run eagerly once on a side stream before the capture (lazy
    initialisations, workspace allocations), which COUNTS as real steps -- callers account for `warm_steps`."""

    def __init__(self, step_fn, k, device, warm=True):
        self.k, self.device = int(k), torch.device(device)
        self.graph = torch.cuda.CUDAGraph()
        self.warm_steps = 0
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            with torch.cuda.graph(self.graph, stream=side, capture_error_mode="thread_local"):
                self.out = [step_fn() for _ in range(self.k)]
        torch.cuda.current_stream(self.device).wait_stream(side)

    def replay(self):
        self.graph.replay()
        return self.out[-1]


class FailureFlag:
    """Sticky device-side record of the worst per-matrix status seen since the last check (no host synchronisation per
    step).  The reference raises NotPSDError the moment a Cholesky fails (gpytorch psd_safe_cholesky); here the failing
    matrices write mll = NaN / zero gradients and the exception is raised at the next ``check()`` -- every log period, at
    the end of meta_fit and before any prediction."""

    def __init__(self, device):
        self.worst = torch.zeros((), dtype=torch.int32, device=device)

    def update(self, info):
        if info is not None:
            torch.minimum(self.worst, info.min(), out=self.worst)

    def check(self):
        if int(self.worst.item()) < 0:
            self.worst.zero_()
            raise NotPSDError("a task kernel matrix was not positive definite after jitter 1e-4 since the last check")


def check_info(info):
    """Host check of the per-matrix status (one device->host read): raises NotPSDError like the reference would."""
    if info is not None and int(info.min().item()) < 0:
        bad = int((info < 0).sum().item())
        raise NotPSDError("%d task kernel matrices are not positive definite after jitter 1e-4" % bad)


def logprob_finalize(theta, prior_mu, prior_sigma, prior_factor, pre, packed, want_grad=True):
    """logp (P,), dtheta (P, D) from the (all-reduced) packed likelihood buffer: pacoh_logprob_finalize."""
    P, D = theta.shape
    logp = torch.empty(P, dtype=torch.float32, device=theta.device)
    dtheta = torch.empty(P, D, dtype=torch.float32, device=theta.device) if want_grad else None
    check(lib.pacoh_logprob_finalize(P, D, _ptr(theta), _ptr(prior_mu), _ptr(prior_sigma), float(prior_factor), float(pre),
                                     ctypes.c_void_p(packed.data_ptr() + 4 * P * D), ctypes.c_void_p(packed.data_ptr()),
                                     _ptr(logp), _ptr(dtheta), _stream()))
    return logp, dtheta


class PeerAllReduce:
    """All-reduce of the packed (P*D + P) likelihood buffer over NVLink peer memory, fused into the finalize kernel
    (pacoh_peer_allreduce_finalize) -- the task-sharded replacement for ``all_reduce`` + ``logprob_finalize``.

    Every rank allocates [2 x packed | flags] in torch symmetric memory (CUDA IPC / fabric handles under the hood) and
    exchanges the mappings once (collective).  Per step the MLL kernels write this rank's partial sums straight into
    the buffer of the step's parity; one kernel then announces / waits through the flags and reads all ranks' buffers.
    Construction raises if symmetric memory is unavailable for the group (callers fall back to NCCL)."""

    def __init__(self, group, P, D, device):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.P, self.D, self.device = int(P), int(D), torch.device(device)
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > _lib.MAX_PEERS:
            raise RuntimeError("peer all-reduce supports at most %d ranks" % _lib.MAX_PEERS)
        self.n = self.P * self.D + self.P
        self.n_pad = (self.n + 63) // 64 * 64
        total = 2 * self.n_pad + 64                      # two parities, then the flag words (uint32 stored in float slots)
        self.local = symm.empty(total, dtype=torch.float32, device=self.device)
        self.local.zero_()
        self.handle = symm.rendezvous(self.local, group)
        bases = [self.handle.get_buffer(r, (total,), torch.float32).data_ptr() for r in range(self.world)]
        self._bufs = [(ctypes.c_void_p * self.world)(*[b + 4 * h * self.n_pad for b in bases]) for h in range(2)]
        self._flags = (ctypes.c_void_p * self.world)(*[b + 4 * 2 * self.n_pad for b in bases])
        self.token = 0
        self.err = torch.zeros(1, dtype=torch.int32, device=self.device)   # set by the kernel if a peer never announced (30 s)
        torch.cuda.synchronize(self.device)
        dist.barrier(group)                              # every rank's flags are zero before anyone announces

    def out_buffer(self):
        """The packed buffer the NEXT finalize() will sum (pass it as ``out=`` to MetaMLLEngine.mll_fwd_bwd)."""
        h = (self.token + 1) & 1
        return self.local[h * self.n_pad:h * self.n_pad + self.n]

    def finalize(self, theta, prior_mu, prior_sigma, prior_factor, pre, token_dev=None):
        """``token_dev``: int32 device tensor whose element 0 holds this call's token (the step counter of a StepState
        that the caller keeps in lock-step with ``self.token``): lets the call sit inside a captured CUDA graph.  The
        buffer parity is still chosen on the host, so a graph must hold an EVEN number of steps."""
        self.token += 1
        h = self.token & 1
        P, D = theta.shape
        assert (P, D) == (self.P, self.D)
        logp = torch.empty(P, dtype=torch.float32, device=self.device)
        dtheta = torch.empty(P, D, dtype=torch.float32, device=self.device)
        tok = 0 if token_dev is not None else self.token & 0xFFFFFFFF
        check(lib.pacoh_peer_allreduce_finalize_dev(self.world, self.rank, self._bufs[h], self._flags, tok, _ptr(token_dev),
                                                    _ptr(self.err), P, D, _ptr(theta), _ptr(prior_mu), _ptr(prior_sigma),
                                                    float(prior_factor), float(pre), _ptr(logp), _ptr(dtheta), _stream()))
        return logp, dtheta

    def check(self):
        if int(self.err.item()) != 0:
            raise RuntimeError("peer all-reduce: rank %d never announced its buffer within 30 s (dead or failed rank)" % (int(self.err.item()) - 1))


def meta_log_prob_and_score(theta, engine, task_idx, prior_mu, prior_sigma, prior_factor, pre, group=None, peer=None, token_dev=None):
    """(logp (P,), score = d logp / d theta (P, D), info) without autograd: what SVGD.phi needs (svgd.py:13-16).
    ``task_idx`` is this rank's shard when ``group`` is given; ``pre`` is computed from the GLOBAL batch.  With a
    ``PeerAllReduce`` the cross-rank sum runs over NVLink peer memory inside the finalize kernel, else as one NCCL
    all-reduce of the packed buffer."""
    if peer is not None:
        _, _, info = engine.mll_fwd_bwd(theta, task_idx, want_mll=False, want_info=True, out=peer.out_buffer())
        logp, dtheta = peer.finalize(theta, prior_mu, prior_sigma, prior_factor, pre, token_dev=token_dev)
        return logp, dtheta, info
    _, packed, info = engine.mll_fwd_bwd(theta, task_idx, want_mll=False, want_info=True)
    if group is not None:
        torch.distributed.all_reduce(packed, group=group)
    logp, dtheta = logprob_finalize(theta, prior_mu, prior_sigma, prior_factor, pre, packed)
    return logp, dtheta, info


class MetaLogProb(torch.autograd.Function):
    """logp_p = prior_factor * log p(theta_p) + pre_factor * sum_t mll_{p,t} with its analytic gradient.

    Drop-in for ``RandomGPMeta.log_prob(params, tuples)`` followed by ``autograd.grad`` / ``backward``
    (random_gp.py:221-222; svgd.py:15-16; GPR_meta_vi.py:221).  With ``group`` set, ``task_idx`` is this rank's
    shard of the sampled batch and the packed (P, D+1) likelihood buffer is summed over ranks with one NCCL
    all-reduce; ``pre`` must then be computed from the GLOBAL batch.
    """

    @staticmethod
    def forward(ctx, theta, engine, task_idx, prior_mu, prior_sigma, prior_factor, pre, group=None):
        logp, dtheta, info = meta_log_prob_and_score(theta.detach().contiguous(), engine, task_idx, prior_mu, prior_sigma,
                                                     prior_factor, pre, group)
        ctx.save_for_backward(dtheta)
        ctx.info = info
        ctx.mark_non_differentiable(info)
        return logp, info

    @staticmethod
    def backward(ctx, grad_logp, _grad_info):
        (dtheta,) = ctx.saved_tensors
        return grad_logp.reshape(-1, 1) * dtheta, None, None, None, None, None, None, None


def meta_log_prob(theta, engine, task_idx, prior_mu, prior_sigma, prior_factor, pre, group=None):
    return MetaLogProb.apply(theta, engine, task_idx, prior_mu, prior_sigma, prior_factor, pre, group)


class SVGDDirection:
    """phi = (K s + grad K) / P on device, median heuristic included (pacoh_svgd_phi; svgd.py:12-23, 32-99).

    ``prepare(theta)`` starts the score-independent half (pairwise distances, median bandwidth, K, row sums) on a side
    stream so that it overlaps the batched MLL forward/backward; the following ``__call__(theta, score)`` with the SAME
    particles then only waits for it and applies K.  Without ``prepare`` the call runs both halves in order."""

    def __init__(self, P, D, device, bandwidth=None, kernel="RBF"):
        if kernel not in ("RBF", "IMQ"):
            raise NotImplementedError("Stein kernel %r: only 'RBF' and 'IMQ' exist (GPR_meta_svgd.py:174-179)" % (kernel,))
        # 'IMQ': IMQSteinKernel(alpha=0.5, beta=-0.5) with the per-dimension median bandwidth and the gradient through it
        self.kind = _lib.SVGD_RBF if kernel == "RBF" else _lib.SVGD_IMQ
        self.P, self.D, self.device = P, D, torch.device(device)
        self.bandwidth = -1.0 if bandwidth is None else float(bandwidth)
        nbytes = check(lib.pacoh_svgd_workspace_bytes(P, D))
        self.ws = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        self.gamma = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._side = None            # side stream + event of a pending prepare()
        self._ready = None
        self._prepared_for = None

    def prepare(self, theta):
        theta = _f32c(theta, self.device)
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
            self._ready = torch.cuda.Event()
        cur = torch.cuda.current_stream(self.device)
        self._side.wait_stream(cur)          # particles of the previous update are final; the last phi_apply is done with K
        with torch.cuda.stream(self._side):
            check(lib.pacoh_svgd_kernel_matrix(self.P, self.D, _ptr(theta), self.bandwidth, self.kind, _ptr(self.gamma),
                                               _ptr(self.ws), self.ws.numel(), _stream()))
            self._ready.record(self._side)
        self._prepared_for = (theta.data_ptr(), theta._version)

    def __call__(self, theta, score, out=None):
        theta, score = _f32c(theta, self.device), _f32c(score, self.device)
        phi = out if out is not None else torch.empty_like(theta)
        if self._prepared_for is not None and self._prepared_for == (theta.data_ptr(), theta._version):
            torch.cuda.current_stream(self.device).wait_event(self._ready)
            check(lib.pacoh_svgd_phi_apply(self.P, self.D, _ptr(theta), _ptr(score), self.kind, _ptr(phi), _ptr(self.gamma),
                                           _ptr(self.ws), self.ws.numel(), _stream()))
        else:
            if self._prepared_for is not None:      # a stale prepare(): let it finish before the workspace is reused
                torch.cuda.current_stream(self.device).wait_event(self._ready)
            check(lib.pacoh_svgd_phi(self.P, self.D, _ptr(theta), _ptr(score), self.bandwidth, self.kind, _ptr(phi),
                                     _ptr(self.gamma), _ptr(self.ws), self.ws.numel(), _stream()))
        self._prepared_for = None
        return phi


def vi_sample(loc, scale, eps):
    """theta = loc + exp(scale) * eps, logq(theta)  (random_gp.py:248, 256-263)."""
    S, D = eps.shape
    theta = torch.empty_like(eps)
    logq = torch.empty(S, dtype=torch.float32, device=eps.device)
    check(lib.pacoh_vi_sample(S, D, _ptr(loc), _ptr(scale), _ptr(eps), _ptr(theta), _ptr(logq), _stream()))
    return theta, logq


def vi_grad(scale, eps, g, prior_factor):
    """Gradient of -(1/S) sum_s [logp(theta_s) - prior_factor * logq(theta_s)] w.r.t. (loc, scale), g = dlogp/dtheta."""
    S, D = eps.shape
    dloc = torch.empty(D, dtype=torch.float32, device=eps.device)
    dscale = torch.empty(D, dtype=torch.float32, device=eps.device)
    check(lib.pacoh_vi_grad(S, D, _ptr(scale), _ptr(eps), _ptr(g), float(prior_factor), _ptr(dloc), _ptr(dscale), _stream()))
    return dloc, dscale


def ffma_peak_tflops(iters=4096, reps=5, gemm_form=False):
    """Measured FP32 FFMA throughput of the device (TFLOP/s): the FP32 roofline denominator reported by bench.py.
    gemm_form=False: immediate-operand chains (peak); True: three-register acc += w * v form."""
    sink = torch.zeros(32, dtype=torch.float32, device="cuda")
    flops = ctypes.c_double(0.0)
    best = 0.0
    for _ in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(lib.pacoh_ffma_peak_launch(-iters if gemm_form else iters, _ptr(sink), ctypes.byref(flops), _stream()))
        e1.record()
        e1.synchronize()
        best = max(best, flops.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


def gp_forward(arch: GPArch, theta, x):
    """mean (P, npts) and features (P, npts, F) of the learned nets at x (npts, d): pacoh_gp_forward.
    For constant / zero mean and the plain SE kernel the trivial values are filled in here."""
    theta, x = theta.contiguous(), x.contiguous()
    P, npts = theta.shape[0], x.shape[0]
    dev = theta.device
    a = arch.c_struct()
    nbytes = check(lib.pacoh_gp_forward_workspace_bytes(ctypes.byref(a), P, npts))
    ws = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
    mean = torch.empty(P, npts, dtype=torch.float32, device=dev) if arch.mean_kind == "NN" else None
    feat = torch.empty(P, npts, arch.feature_dim, dtype=torch.float32, device=dev) if arch.covar_kind == "NN" else None
    check(lib.pacoh_gp_forward(ctypes.byref(a), P, npts, _ptr(theta), _ptr(x), _ptr(mean), _ptr(feat), _ptr(ws), ws.numel(), _stream()))
    ent = arch.entries()
    if mean is None:
        if arch.mean_kind == "constant":
            c0 = ent["constant_mean"][0]
            mean = theta[:, c0:c0 + 1].expand(P, npts).contiguous()
        else:
            mean = torch.zeros(P, npts, dtype=torch.float32, device=dev)
    if feat is None:
        feat = x.unsqueeze(0).expand(P, npts, x.shape[1]).contiguous()
    return mean, feat


class PosteriorBatch:
    """Eval-mode posterior of P parameter vectors on a batch of test tasks (pacoh_gp_posterior), normalised space.
    ``mu`` / ``var`` (P, Tt, ns_max) device tensors (var includes the observation noise), ``n_s`` per-task test sizes,
    ``joint_ll`` (P, Tt) when targets were given, ``cov`` (P, Tt, ns_max, ns_max) on request, ``info`` (P, Tt)."""

    def __init__(self, mu, var, n_s, joint_ll, cov, info, ys_dev):
        self.mu, self.var, self.n_s, self.joint_ll, self.cov, self.info, self.ys_dev = mu, var, n_s, joint_ll, cov, info, ys_dev


_post_ws = {}


def gp_posterior_batch(arch: GPArch, theta, contexts, tests, targets=None, want_cov=False):
    """Exact GP posterior for every row of ``theta`` (P, D) on Tt test tasks at once (get_pred_dist of the reference, per
    task and per particle: GPR_meta_svgd.py:203-212, GPR_meta_vi.py:229-252, GPR_meta_mll.py:174-183).

    contexts: list of (x_c (n_c, d), y_c (n_c,)) normalised float arrays / tensors; tests: list of x* (n*, d);
    targets: optional list of normalised y* (n*,) -> joint log-likelihoods.  All algebra runs in the CUDA kernels of
    csrc/gp_post.cu (no torch.linalg): returns a PosteriorBatch."""
    theta = _f32c(theta.detach().contiguous(), theta.device)
    dev = theta.device
    P, D = theta.shape
    assert D == arch.D
    Tt, d = len(contexts), arch.input_dim
    assert len(tests) == Tt and (targets is None or len(targets) == Tt)
    as_np = lambda v: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)).astype(np.float32)   # noqa: E731
    nc = np.asarray([len(c[0]) for c in contexts], dtype=np.int32)
    ns = np.asarray([len(x) for x in tests], dtype=np.int32)
    nc_max, ns_max = int(nc.max()), int(ns.max())
    xc = np.zeros((Tt, nc_max, d), np.float32); yc = np.zeros((Tt, nc_max), np.float32)
    xs = np.zeros((Tt, ns_max, d), np.float32); ys = np.zeros((Tt, ns_max), np.float32)
    for t in range(Tt):
        xc[t, :nc[t]] = as_np(contexts[t][0]).reshape(nc[t], d)
        yc[t, :nc[t]] = as_np(contexts[t][1]).reshape(nc[t])
        xs[t, :ns[t]] = as_np(tests[t]).reshape(ns[t], d)
        if targets is not None:
            ys[t, :ns[t]] = as_np(targets[t]).reshape(ns[t])
    up = lambda a_: torch.from_numpy(a_).to(dev)   # noqa: E731
    xc_d, yc_d, xs_d, nc_d, ns_d = up(xc), up(yc), up(xs), up(nc), up(ns)
    ys_d = up(ys) if targets is not None else None
    a = arch.c_struct()
    nbytes = check(lib.pacoh_gp_posterior_workspace_bytes(ctypes.byref(a), P, Tt, nc_max, ns_max, 1 if want_cov else 0))
    key = (dev, )
    if key not in _post_ws or _post_ws[key].numel() < nbytes:
        _post_ws[key] = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=dev)
    ws = _post_ws[key]
    mu = torch.empty(P, Tt, ns_max, dtype=torch.float32, device=dev)
    var = torch.empty_like(mu)
    cov = torch.empty(P, Tt, ns_max, ns_max, dtype=torch.float32, device=dev) if want_cov else None
    jll = torch.empty(P, Tt, dtype=torch.float32, device=dev) if targets is not None else None
    info = torch.empty(P, Tt, dtype=torch.int32, device=dev)
    check(lib.pacoh_gp_posterior(ctypes.byref(a), P, Tt, nc_max, ns_max, _ptr(theta), _ptr(xc_d), _ptr(yc_d), _ptr(nc_d), _ptr(xs_d),
                                 _ptr(ns_d), _ptr(ys_d), _ptr(mu), _ptr(var), _ptr(cov), _ptr(jll), _ptr(info), _ptr(ws), ws.numel(), _stream()))
    return PosteriorBatch(mu, var, ns_d, jll, cov, info, ys_d)


def pred_metrics(post: PosteriorBatch, y_std):
    """(Tt, 3) tensor [avg test log-likelihood, RMSE, calibration error] per test task of the equally weighted mixture over
    the parameter vectors (abstract.py:157-161, 260-272; models.py:90-126): pacoh_pred_metrics."""
    P, Tt, ns_max = post.mu.shape
    assert post.ys_dev is not None, "metrics need the test targets"
    out = torch.empty(Tt, 3, dtype=torch.float32, device=post.mu.device)
    check(lib.pacoh_pred_metrics(P, Tt, ns_max, _ptr(post.mu), _ptr(post.var), _ptr(post.n_s), _ptr(post.ys_dev), _ptr(post.joint_ll),
                                 float(y_std), _ptr(out), _stream()))
    return out


class GPPredictive(torch.distributions.Distribution):
    """The normalised-space predictive distribution of one test task: what gpytorch's MultivariateNormal is to the
    reference's get_pred_dist (batch = parameter vectors, event = test points).  ``mean`` / ``variance`` / ``stddev`` come
    from the posterior kernel; ``log_prob(y)`` is the JOINT density with the full predictive covariance, evaluated on the
    device through the marginal-likelihood kernels; ``covariance_matrix`` is built on first access; ``cdf`` / ``icdf`` are
    the marginal Normal ones.  ``squeeze=True`` drops the batch dimension (PACOH-MAP / VI 'MAP' mode: one parameter vector)."""
    arg_constraints = {}
    has_rsample = False

    def __init__(self, arch, theta, context, x_test, squeeze=False):
        self._arch, self._theta, self._context, self._x = arch, theta, context, x_test
        post = gp_posterior_batch(arch, theta, [context], [x_test])
        if int(post.info.min().item()) < 0:
            raise NotPSDError("context kernel matrix is not positive definite")
        self._squeeze = squeeze
        self._mean, self._var = post.mu[:, 0].cpu(), post.var[:, 0].cpu()
        self._cov = None
        P, ns = self._mean.shape
        super().__init__(batch_shape=torch.Size([]) if squeeze else torch.Size([P]), event_shape=torch.Size([ns]), validate_args=False)

    def _sq(self, v):
        return v[0] if self._squeeze else v

    @property
    def mean(self):
        return self._sq(self._mean)

    loc = mean

    @property
    def variance(self):
        return self._sq(self._var)

    @property
    def stddev(self):
        return self._sq(self._var.sqrt())

    @property
    def covariance_matrix(self):
        if self._cov is None:
            self._cov = gp_posterior_batch(self._arch, self._theta, [self._context], [self._x], want_cov=True).cov[:, 0].cpu()
        return self._sq(self._cov)

    def log_prob(self, value):
        value = torch.as_tensor(value, dtype=torch.float32).reshape(-1)
        post = gp_posterior_batch(self._arch, self._theta, [self._context], [self._x], targets=[value])
        return self._sq(post.joint_ll[:, 0].cpu())

    def cdf(self, value):
        return torch.distributions.Normal(self.mean, self.stddev).cdf(value)

    def icdf(self, value):
        return torch.distributions.Normal(self.mean, self.stddev).icdf(value)


class PacohAdam(torch.optim.Optimizer):
    """torch.optim.Adam-compatible optimizer (same state keys: step / exp_avg / exp_avg_sq, same update order) whose
    step is one fused CUDA kernel over the flat (P, D) particle matrix (pacoh_adam_step).  ``direction`` lets the
    SVGD step feed phi directly (grad = -phi, svgd.py:27) without materialising ``.grad``."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    @torch.no_grad()
    def step(self, direction=None, state=None):
        """``state``: a StepState already advanced for this step (pacoh_step_prepare): the learning rate and the bias
        corrections are then read on the device (pacoh_adam_step_dev) -- the form a captured CUDA graph needs.  The host
        copies of ``step`` / ``lr`` are kept in sync by the caller (sync_from)."""
        for group in self.param_groups:
            for p in group["params"]:
                g, sign = (direction, -1.0) if direction is not None else (p.grad, 1.0)
                if g is None:
                    continue
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
                b1, b2 = group["betas"]
                assert p.is_contiguous() and g.is_contiguous() and p.dtype == torch.float32
                if state is not None:
                    check(lib.pacoh_adam_step_dev(p.numel(), _ptr(p), _ptr(g), sign, _ptr(st["exp_avg"]), _ptr(st["exp_avg_sq"]),
                                                  float(b1), float(b2), float(group["eps"]), _ptr(state.buf), _stream()))
                    continue
                st["step"] += 1
                check(lib.pacoh_adam_step(p.numel(), _ptr(p), _ptr(g), sign, _ptr(st["exp_avg"]), _ptr(st["exp_avg_sq"]),
                                          float(group["lr"]), float(b1), float(b2), float(group["eps"]), int(st["step"]), _stream()))

    def sync_from(self, state):
        """Host bookkeeping after device-state steps: torch-compatible ``step`` counts (state_dict round trips)."""
        for group in self.param_groups:
            for p in group["params"]:
                if p in self.state and len(self.state[p]):
                    self.state[p]["step"] = state.steps


class StageTiming:
    """Context manager around pacoh_stage_timing_*: per-stage device milliseconds of pacoh_meta_mll_fwd_bwd."""
    NAMES = ("mlp_fwd", "gp_mll", "mlp_bwd", "reduce")

    def __enter__(self):
        check(lib.pacoh_stage_timing_enable(1))
        return self

    def read(self):
        ms = (ctypes.c_float * 4)()
        calls = ctypes.c_int32(0)
        check(lib.pacoh_stage_timing_read(ms, ctypes.byref(calls)))
        return {n: float(ms[i]) for i, n in enumerate(self.NAMES)}, int(calls.value)

    def __exit__(self, *exc):
        check(lib.pacoh_stage_timing_enable(0))
        return False
