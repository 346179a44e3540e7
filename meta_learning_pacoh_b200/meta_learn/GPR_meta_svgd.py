"""PACOH-SVGD on the B200 engine: same constructor, attributes and methods as the reference's
meta_learn/GPR_meta_svgd.py:14-235 (GPRegressionMetaLearnedSVGD), with the per-task / per-particle Python loops
(random_gp.py:214-217, svgd.py:12-28) replaced by three C-ABI calls per step:

    pacoh_meta_mll_fwd_bwd  ->  pacoh_logprob_finalize   [task-sharded: pacoh_peer_allreduce_finalize, or NCCL + finalize]
    pacoh_svgd_phi          ->  pacoh_adam_step
"""
import os
import time

import numpy as np
import torch

from .. import engine as eng
from . import prior_init
from .abstract import RegressionModelMetaLearned
from .distributions import AffineTransformedDistribution, EqualWeightedMixtureDist
from .util import DummyLRScheduler, _handle_input_dimensionality


class GPRegressionMetaLearnedSVGD(RegressionModelMetaLearned):

    def __init__(self, meta_train_data, num_iter_fit=10000, feature_dim=1,
                 prior_factor=0.01, weight_prior_std=0.5, bias_prior_std=3.0,
                 covar_module='NN', mean_module='NN', mean_nn_layers=(32, 32), kernel_nn_layers=(32, 32),
                 optimizer='Adam', lr=1e-3, lr_decay=1.0, kernel='RBF', bandwidth=None, num_particles=10,
                 task_batch_size=-1, normalize_data=True, random_seed=None):
        """PACOH-SVGD: Stein variational gradient descent on the PAC-optimal hyper-posterior over GP priors.

        Args: identical to the reference (GPR_meta_svgd.py:16-45).  ``feature_dim`` is accepted and ignored exactly
        as the reference does (its kernel net always has 2 outputs: GPR_meta_svgd.py:167-170, random_gp.py:24).
        """
        super().__init__(normalize_data, random_seed)
        assert mean_module in ['NN', 'constant', 'zero']
        assert covar_module in ['NN', 'SE']
        assert optimizer in ['Adam', 'SGD']

        self.num_iter_fit, self.prior_factor, self.feature_dim = num_iter_fit, prior_factor, feature_dim
        self.weight_prior_std, self.bias_prior_std = weight_prior_std, bias_prior_std
        self.num_particles = num_particles
        if task_batch_size < 1:
            self.task_batch_size = len(meta_train_data)
        else:
            self.task_batch_size = min(task_batch_size, len(meta_train_data))

        self._check_meta_data_shapes(meta_train_data)
        self._compute_normalization_stats(meta_train_data)

        self._setup_model_inference(mean_module, covar_module, mean_nn_layers, kernel_nn_layers,
                                    kernel, bandwidth, optimizer, lr, lr_decay)

        X, Y = self._build_task_dicts(meta_train_data)
        self.engine = eng.MetaMLLEngine(self.arch, X, Y, self.device, task_n=self.task_sizes)
        self._idx_ring = eng.PinnedRing(self.device)
        self._failures = eng.FailureFlag(self.device)
        self._group, self._peer = None, None
        self._rank, self._world = 0, 1
        self.fitted = False

    # ------------------------------------------------------------------ multi-GPU
    def shard_tasks(self, group=None):
        """Task-shard the sampled batch over the ranks of ``group`` (default: the world group).  Every rank must be
        constructed with the same data and seed: all ranks then draw the same batch indices, evaluate a contiguous
        slice of them, and one cross-rank sum of the packed (P, D+1) likelihood buffer restores the full sums."""
        import torch.distributed as dist
        assert dist.is_initialized()
        self._group = group if group is not None else dist.group.WORLD
        self._rank, self._world = dist.get_rank(self._group), dist.get_world_size(self._group)
        # cross-rank sum fused into the finalize kernel over NVLink peer memory (engine.PeerAllReduce); NCCL all-reduce
        # when symmetric memory is unavailable for this group or PACOH_ALLREDUCE=nccl asks for it (A/B measurements)
        self._peer, self._peer_error = None, None
        if self._world > 1 and os.environ.get("PACOH_ALLREDUCE", "peer") != "nccl":
            ok = torch.zeros(1, device=self.device)
            try:
                self._peer = eng.PeerAllReduce(self._group, self.num_particles, self.arch.D, self.device)
                ok += 1
            except Exception as e:   # noqa: BLE001  (any failure to map peer memory: fall back on every rank)
                self._peer_error = repr(e)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self._group)
            if ok.item() < 1:
                self._peer = None
        self._graph = None
        if self._peer is not None and self._state is not None:
            self._peer.token = self._state.steps       # the kernel's token is the device step count; the host copy picks the buffer parity
        return self

    # ------------------------------------------------------------------ training
    def meta_fit(self, valid_tuples=None, verbose=True, log_period=500, n_iter=None):
        """Fits the hyper-posterior particles with SVGD -- GPR_meta_svgd.py:82-121."""
        assert (valid_tuples is None) or (all([len(valid_tuple) == 4 for valid_tuple in valid_tuples]))
        t = time.time()
        if n_iter is None:
            n_iter = self.num_iter_fit
        itr = 0
        while itr < n_iter:
            # run up to the next logging point (iteration 1 and every multiple of log_period, GPR_meta_svgd.py:106)
            nxt = 1 if itr == 0 else min(n_iter, (itr // log_period + 1) * log_period)
            self.run_steps(nxt - itr)
            itr = nxt
            if itr == 1 or itr % log_period == 0:
                self._failures.check()
                duration = time.time() - t
                t = time.time()
                message = 'Iter %d/%d - Time %.2f sec' % (itr, self.num_iter_fit, duration)
                if valid_tuples is not None:
                    valid_ll, valid_rmse, calibr_err = self.eval_datasets(valid_tuples)
                    message += ' - Valid-LL: %.3f - Valid-RMSE: %.3f - Calib-Err %.3f' % (valid_ll, valid_rmse, calibr_err)
                if verbose:
                    self.logger.info(message)
        self._failures.check()
        self.fitted = True

    # One code path for eager and graph-captured steps: everything step-dependent (Adam's step count and bias
    # corrections, the StepLR learning rate, the cross-rank token, the sampled task indices) is read from DEVICE memory
    # (engine.StepState / pacoh_step_prepare), so the same call sequence can be replayed from a CUDA graph.
    def _device_step(self, engine, idx_cur, K, idx_stream, pre):
        st = self._state
        st.prepare(K, idx_cur.numel(), idx_stream, idx_cur)
        self._phi.prepare(self.particles)                              # K(theta) on a side stream, under the MLL kernels
        logp, score, info = eng.meta_log_prob_and_score(self.particles, engine, idx_cur, self._prior_mu, self._prior_sigma,
                                                        self.prior_factor, pre, self._group, self._peer,
                                                        token_dev=st.buf if self._peer is not None else None)
        phi = self._phi(self.particles, score)
        self.optimizer.step(direction=phi, state=st)                   # grad = -phi (svgd.py:27)
        self._failures.update(info)
        return logp, info

    def svgd_step(self, task_idx):
        """One SVGD update on the sampled batch (closure svgd_step, GPR_meta_svgd.py:190-199 -> svgd.py:25-28).
        ``task_idx``: numpy / sequence of task indices into the meta-training set (with repetitions); the reference's
        closure receives the sampled task dicts themselves -- ``svgd_step_host`` is the variant fed with the tensors."""
        idx = np.asarray(task_idx, dtype=np.int32)
        T = idx.shape[0]
        assert T >= self._world, "task batch (%d) smaller than the number of ranks (%d): every rank needs a task" % (T, self._world)
        lo, hi = eng.shard_bounds(T, self._rank, self._world)
        pre = eng.pre_factor(self.task_sizes[idx])                     # GLOBAL batch, harmonic mean of its n_t (random_gp.py:209-212)
        if self._state is None:                                        # optimizer='SGD': host-side step (no device state)
            idx_dev = self._idx_ring.upload(idx[lo:hi])
            self._phi.prepare(self.particles)
            logp, score, info = eng.meta_log_prob_and_score(self.particles, self.engine, idx_dev, self._prior_mu,
                                                            self._prior_sigma, self.prior_factor, pre, self._group, self._peer)
            phi = self._phi(self.particles, score)
            self.optimizer.zero_grad()
            self.particles.grad = -phi
            self.optimizer.step()
            self._failures.update(info)
        else:
            if self._idx_cur is None or self._idx_cur.numel() != hi - lo:
                self._idx_cur = torch.empty(hi - lo, dtype=torch.int32, device=self.device)
                self._graph = None
            stream = self._idx_ring.upload(idx[lo:hi])
            logp, info = self._device_step(self.engine, self._idx_cur, 1, stream, pre)
            self.optimizer.sync_from(self._state)
        self._last_info, self._last_logp = info, logp
        return logp

    GRAPH_STEPS = 10     # training steps per captured CUDA graph (even: the peer buffers alternate by step parity)

    def run_steps(self, n):
        """``n`` meta-training steps (sample a batch with replacement, svgd_step, lr schedule) -- the body of the reference's
        meta_fit loop (GPR_meta_svgd.py:100-104).  With the Adam optimizer and equally sized tasks, GRAPH_STEPS steps at a
        time are replayed from one CUDA graph whose kernels read the step state and the pre-uploaded index stream (the same
        numpy RandomState.choice draws, in the same order) from the device; PACOH_GRAPH=0 keeps every step eager.  Both
        forms launch the same kernels in the same order."""
        K = self.GRAPH_STEPS
        # graphs only where no NCCL collective would sit inside the capture (single rank, or the peer-memory cross-rank sum):
        # processes whose captured graphs hold NCCL kernels were seen to hang at process-group teardown
        use_graph = (self._state is not None and not self._ragged and os.environ.get("PACOH_GRAPH", "1") != "0"
                     and (self._world == 1 or self._peer is not None))
        while n > 0:
            if use_graph and n >= K and self._state.steps % 2 == 0 and self._state.steps > 0 and self._idx_cur is not None \
                    and self._idx_cur.numel() == self._local_batch():
                lo, hi = eng.shard_bounds(self.task_batch_size, self._rank, self._world)
                if self._graph is None:
                    self._idx_stream = torch.zeros(K, hi - lo, dtype=torch.int32, device=self.device)
                    pre = eng.pre_factor(self.task_sizes[:1].repeat(self.task_batch_size))
                    s0, t0 = self._state.steps, (self._peer.token if self._peer is not None else 0)
                    self._graph = eng.StepGraph(lambda: self._device_step(self.engine, self._idx_cur, K, self._idx_stream, pre), K, self.device)
                    self._state.steps = s0                          # the capture only recorded: nothing ran
                    if self._peer is not None:
                        self._peer.token = t0
                s0 = self._state.steps
                rows = np.empty((K, hi - lo), dtype=np.int32)
                for j in range(K):                                   # row (step mod K) = the batch of that step
                    rows[(s0 + j) % K] = self._sample_task_indices()[lo:hi]
                self._idx_ring.upload(rows, out=self._idx_stream)
                self._last_logp, self._last_info = self._graph.replay()
                self._state.steps += K
                if self._peer is not None:
                    self._peer.token += K
                self.optimizer.sync_from(self._state)
                for _ in range(K):
                    self.lr_scheduler.step()
                n -= K
            else:
                self.svgd_step(self._sample_task_indices())
                self.lr_scheduler.step()
                n -= 1

    def _local_batch(self):
        lo, hi = eng.shard_bounds(self.task_batch_size, self._rank, self._world)
        return hi - lo

    def svgd_step_host(self, x_batch, y_batch, global_tasks=None, wait=True):
        """Same update as svgd_step, fed like the reference's closure is (GPR_meta_svgd.py:190-199 receives the sampled
        task tensors themselves): ``x_batch`` (T, n, d) / ``y_batch`` (T, n) are HOST tensors (pinned for async copies)
        holding the sampled, normalised batch in order.  They are copied to the device, the step runs on them with
        identity task indices, and logp (P,) is read back to the host.  Used for the end-to-end measurement.

        ``global_tasks``: when given, ``x_batch`` / ``y_batch`` are already THIS rank's shard of a ``global_tasks``-task
        batch (the caller gathered only rows ``shard_bounds(global_tasks, rank, world)``); otherwise the full batch is
        passed and sliced here.  ``wait=False`` returns ``(logp_pinned, event)`` without synchronising: the device->host
        copy of logp is in flight and ``event.synchronize()`` must precede any read -- lets the caller prepare the next
        batch on the host while this step runs (the copies and the read-back still happen every step)."""
        if self._ragged:
            raise NotImplementedError("svgd_step_host takes dense (T, n, d) host batches: use svgd_step for ragged task sets")
        if global_tasks is None:
            T = x_batch.shape[0]
            lo, hi = eng.shard_bounds(T, self._rank, self._world)
            x_batch, y_batch = x_batch[lo:hi], y_batch[lo:hi]
        else:
            T = int(global_tasks)
            lo, hi = eng.shard_bounds(T, self._rank, self._world)
            assert x_batch.shape[0] == hi - lo, "pre-sharded batch does not match shard_bounds"
        Ts = hi - lo
        if getattr(self, "_stage_engine", None) is None or self._stage_engine.T_total != 2 * Ts:
            # two device staging halves: the batch of step k+1 is copied (on a copy stream) while step k computes
            xs = torch.empty((2 * Ts,) + tuple(x_batch.shape[1:]), dtype=torch.float32)
            ys = torch.empty((2 * Ts,) + tuple(y_batch.shape[1:]), dtype=torch.float32)
            self._stage_engine = eng.MetaMLLEngine(self.arch, xs, ys, self.device)
            self._stage_idx = [torch.arange(h * Ts, (h + 1) * Ts, dtype=torch.int32, device=self.device) for h in range(2)]
            self._logp_host = [torch.empty(self.num_particles, dtype=torch.float32).pin_memory() for _ in range(2)]
            self._logp_event = [torch.cuda.Event() for _ in range(2)]
            self._copy_event = [torch.cuda.Event() for _ in range(2)]
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._logp_slot = 0
            self._stage_used = [False, False]
        se = self._stage_engine
        slot = self._logp_slot
        self._logp_slot ^= 1
        if self._stage_used[slot]:
            self._logp_event[slot].synchronize()        # the step that last used this half (two calls ago) is done
        self._stage_used[slot] = True
        with torch.cuda.stream(self._copy_stream):
            se.x[slot * Ts:(slot + 1) * Ts].copy_(x_batch, non_blocking=True)
            se.y[slot * Ts:(slot + 1) * Ts].copy_(y_batch, non_blocking=True)
            self._copy_event[slot].record(self._copy_stream)
        torch.cuda.current_stream(self.device).wait_event(self._copy_event[slot])
        pre = eng.pre_factor([se.n] * T)
        if self._state is not None:
            logp, info = self._device_step(se, self._stage_idx[slot], 0, None, pre)
            self.optimizer.sync_from(self._state)
        else:
            self._phi.prepare(self.particles)
            logp, score, info = eng.meta_log_prob_and_score(self.particles, se, self._stage_idx[slot], self._prior_mu,
                                                            self._prior_sigma, self.prior_factor, pre, self._group, self._peer)
            phi = self._phi(self.particles, score)
            self.optimizer.zero_grad()
            self.particles.grad = -phi
            self.optimizer.step()
            self._failures.update(info)
        self._last_info = info
        out, ev = self._logp_host[slot], self._logp_event[slot]
        out.copy_(logp, non_blocking=True)
        ev.record(torch.cuda.current_stream(self.device))
        if not wait:
            return out, ev
        ev.synchronize()
        return out.clone()

    # ------------------------------------------------------------------ prediction
    def predict(self, context_x, context_y, test_x, return_density=False):
        """Posterior inference on (context_x, context_y), predictive mixture over particles at test_x --
        GPR_meta_svgd.py:123-159."""
        self._failures.check()
        base = self._predictive(context_x, context_y, test_x)
        pred_dist = AffineTransformedDistribution(base, normalization_mean=self.y_mean, normalization_std=self.y_std)
        pred_dist = EqualWeightedMixtureDist(pred_dist, batched=True)
        if return_density:
            return pred_dist
        return pred_dist.mean.numpy(), pred_dist.stddev.numpy()

    def _predict_params(self):
        return self.particles

    # ------------------------------------------------------------------ setup
    def _setup_model_inference(self, mean_module_str, covar_module_str, mean_nn_layers, kernel_nn_layers,
                               kernel, bandwidth, optimizer, lr, lr_decay):
        assert mean_module_str in ['NN', 'constant']
        assert covar_module_str in ['NN', 'SE']
        # random GP model: RandomGPMeta(size_in, prior_factor, ..., covar_module_str, mean_module_str, layers)
        # (GPR_meta_svgd.py:166-170); feature_dim is NOT forwarded there -> always 2 (random_gp.py:24)
        self.arch = eng.GPArch(self.input_dim, mean_kind=mean_module_str, covar_kind=covar_module_str,
                               mean_layers=tuple(mean_nn_layers), kernel_layers=tuple(kernel_nn_layers), feature_dim=2)
        prior_init.consume_vectorized_gp_init(self.arch)
        mu, sigma = self.arch.hyper_prior(self.weight_prior_std, self.bias_prior_std)
        self._prior_mu, self._prior_sigma = mu.to(self.device), sigma.to(self.device)

        if kernel not in ('RBF', 'IMQ'):
            raise NotImplementedError
        # initial particle locations from the hyper-prior (GPR_meta_svgd.py:182)
        self.particles = prior_init.sample_params_from_prior(self.arch, self.num_particles, self.weight_prior_std,
                                                             self.bias_prior_std).contiguous().to(self.device)
        self._phi = eng.SVGDDirection(self.num_particles, self.arch.D, self.device, bandwidth=bandwidth, kernel=kernel)
        self._setup_optimizer(optimizer, lr, lr_decay)
        self._last_info = None

    def _setup_optimizer(self, optimizer, lr, lr_decay):
        assert hasattr(self, 'particles'), "SVGD must be initialized before setting up optimizer"
        self._state, self._graph, self._idx_cur = None, None, None
        if optimizer == 'Adam':
            self.optimizer = eng.PacohAdam([self.particles], lr=lr)
            self._state = eng.StepState(self.device, lr, lr_decay)     # lr, StepLR(1000, lr_decay) and Adam's step count on the device
        elif optimizer == 'SGD':
            self.optimizer = torch.optim.SGD([self.particles], lr=lr)
        else:
            raise NotImplementedError('Optimizer must be Adam or SGD')
        if lr_decay < 1.0:
            self.lr_scheduler = torch.optim.lr_scheduler.StepLR(self.optimizer, 1000, gamma=lr_decay)
        else:
            self.lr_scheduler = DummyLRScheduler()

    def _vectorize_pred_dist(self, pred_dist):
        mvn = pred_dist.dists
        normal = torch.distributions.Normal(mvn.mean, mvn.stddev)
        return EqualWeightedMixtureDist(normal, batched=True, num_dists=mvn.batch_shape[0])
