"""PACOH-MAP on the B200 engine: call-compatible with the reference's meta_learn/GPR_meta_mll.py:12-264
(GPRegressionMetaLearned).  The parameters stay ordinary torch modules (same names and state_dict keys as the
reference's LearnedGPRegressionModel, so checkpoints interchange); every iteration packs them into the engine's flat
layout, evaluates loss = -sum_{t in batch} mll_t and its gradient with ONE pacoh_meta_mll_fwd_bwd call (P = 1,
ScaleKernel outputscale, noise floor 1e-3) instead of the per-task loop (GPR_meta_mll.py:109-113), and scatters the
gradient back into ``.grad`` for torch's AdamW / SGD.
"""
import os
import time
from collections import OrderedDict

import numpy as np
import torch

from .. import engine as eng
from .abstract import RegressionModelMetaLearned
from .distributions import AffineTransformedDistribution
from .util import DummyLRScheduler, _handle_input_dimensionality


class NeuralNetwork(torch.nn.Sequential):
    """tanh MLP with submodules fc_1..fc_L, out (same names / init order as meta_learn/models.py:190-217)."""

    def __init__(self, input_dim=2, output_dim=2, layer_sizes=(64, 64)):
        super().__init__()
        self.n_layers = len(layer_sizes)
        prev = input_dim
        for i, size in enumerate(layer_sizes):
            setattr(self, 'fc_%i' % (i + 1), torch.nn.Linear(prev, size))
            prev = size
        setattr(self, 'out', torch.nn.Linear(prev, output_dim))

    def linears(self):
        return [getattr(self, 'fc_%i' % i) for i in range(1, self.n_layers + 1)] + [self.out]

    def forward(self, x):
        for lin in self.linears()[:-1]:
            x = torch.tanh(lin(x))
        return self.out(x)


class _SyncedAdamW(torch.optim.AdamW):
    """torch.optim.AdamW whose moment buffers are views of the learner's flat (D,) buffers and whose step count mirrors the
    learner's device-side step state: the fused device step (pacoh_adamw_step_dev) and torch's own ``step()`` advance the SAME
    optimizer state, and ``state_dict()`` keeps torch's format (GPR_meta_mll.py:192-205 checkpoints)."""

    owner = None

    def step(self, closure=None):
        o = self.owner
        if o is not None:
            o._sync_step_tensors()
        out = super().step(closure)
        if o is not None:
            o._state.steps += 1
            o._state.buf[0:1].fill_(o._state.steps)
        return out


class GPRegressionMetaLearned(RegressionModelMetaLearned):
    NOISE_FLOOR = 1e-3     # gpytorch GaussianLikelihood(noise_constraint=GreaterThan(1e-3)), GPR_meta_mll.py:54-55

    def __init__(self, meta_train_data, learning_mode='both', lr_params=1e-3, weight_decay=0.0, feature_dim=2,
                 num_iter_fit=10000, covar_module='NN', mean_module='NN', mean_nn_layers=(32, 32), kernel_nn_layers=(32, 32),
                 task_batch_size=5, normalize_data=True, optimizer='Adam', lr_decay=1.0, random_seed=None):
        """Meta-learning GP priors (mean and kernel function) via PACOH-MAP.  Args identical to the reference
        (GPR_meta_mll.py:14-37); gpytorch Kernel/Mean objects as covar_module / mean_module are not supported."""
        super().__init__(normalize_data, random_seed)
        assert learning_mode in ['learn_mean', 'learn_kernel', 'both', 'vanilla']
        assert mean_module in ['NN', 'constant', 'zero'], "gpytorch Mean objects are outside the CUDA path"
        assert covar_module in ['NN', 'SE'], "gpytorch Kernel objects are outside the CUDA path"
        assert optimizer in ['Adam', 'SGD']

        self.lr_params, self.weight_decay, self.feature_dim = lr_params, weight_decay, feature_dim
        self.num_iter_fit, self.task_batch_size, self.normalize_data = num_iter_fit, task_batch_size, normalize_data
        self.learning_mode = learning_mode

        self._check_meta_data_shapes(meta_train_data)
        self._compute_normalization_stats(meta_train_data)
        self._setup_gp_prior(mean_module, covar_module, learning_mode, feature_dim, mean_nn_layers, kernel_nn_layers)

        # likelihood: noise = 1e-3 + softplus(raw_noise), raw_noise (1,) = 0  (GPR_meta_mll.py:54-56)
        self.raw_noise = torch.nn.Parameter(torch.zeros(1, device=self.device))
        self.shared_parameters.append({'params': [self.raw_noise], 'lr': self.lr_params})

        X, Y = self._build_task_dicts(meta_train_data)
        self.engine = eng.MetaMLLEngine(self.arch, X, Y, self.device, task_n=self.task_sizes)
        self._flatten_parameters()
        self._setup_optimizer(optimizer, lr_params, lr_decay)
        self._idx_ring = eng.PinnedRing(self.device)
        self._failures = eng.FailureFlag(self.device)
        self.fitted = False

    # ------------------------------------------------------------------ flat-layout packing
    def _named_flat(self):
        """[(flat-layout name, tensor)] in the engine's order."""
        out = []
        for prefix, net in (("mean_nn", self.nn_mean_fn), ("kernel_nn", self.nn_kernel_map)):
            if net is None:
                continue
            lins = net.linears()
            names = ["fc_%d" % (i + 1) for i in range(len(lins) - 1)] + ["out"]
            for nm, lin in zip(names, lins):
                out.append(("%s.%s.bias" % (prefix, nm), lin.bias))
                out.append(("%s.%s.weight" % (prefix, nm), lin.weight))
        if self.arch.mean_kind == "constant":
            out.insert(0, ("constant_mean", self.constant_mean))
        out.append(("lengthscale_raw", self.raw_lengthscale))
        out.append(("noise_raw", self.raw_noise))
        out.append(("outputscale_raw", self.raw_outputscale))
        assert [n for n, _ in out] == list(self.arch.entries().keys())
        return out

    def _flatten_parameters(self):
        """Make every module parameter a VIEW into one flat (1, D) buffer in the engine's layout (and its gradient a view into
        a flat gradient buffer): packing the parameters for the kernels is then free and scattering the gradient is one
        copy, instead of a cat over 15 tensors and 15 clones per step.  The nn.Parameter objects (names, state_dict keys,
        optimizer groups) are unchanged; in-place updates by the optimizer / load_state_dict keep the views valid."""
        named = self._named_flat()
        flat = torch.cat([t.detach().reshape(-1) for _, t in named]).contiguous()
        self._flat, self._gradflat = flat.view(1, -1), torch.zeros_like(flat)
        learn_kernel = self.learning_mode in ("learn_kernel", "both")
        learn_mean = self.learning_mode in ("learn_mean", "both")
        self._grad_views = []
        for (name, t), (a, b) in zip(named, self.arch.entries().values()):
            t.data = flat[a:b].view(t.shape)
            trainable = {"mean_nn": learn_mean, "constant_mean": learn_mean, "kernel_nn": learn_kernel,
                         "lengthscale_raw": learn_kernel, "outputscale_raw": learn_kernel, "noise_raw": True}[name.split(".")[0]]
            if trainable:
                self._grad_views.append((t, self._gradflat[a:b].view(t.shape)))

    def _link_optimizer_state(self):
        """Moment buffers of every trainable parameter = views of the flat (D,) buffers the fused AdamW kernel updates."""
        D = self.arch.D
        if not hasattr(self, "_mflat"):
            self._mflat = torch.zeros(D, dtype=torch.float32, device=self.device)
            self._vflat = torch.zeros(D, dtype=torch.float32, device=self.device)
            self._mask = torch.zeros(D, dtype=torch.uint8, device=self.device)
        trainable = {id(t) for t, _ in self._grad_views}
        for (name, t), (a, b) in zip(self._named_flat(), self.arch.entries().values()):
            if id(t) not in trainable:
                continue
            self._mask[a:b] = 1
            st = self.optimizer.state[t]
            if "exp_avg" in st:                       # loaded from a checkpoint: adopt the values, then alias
                self._mflat[a:b].copy_(st["exp_avg"].reshape(-1))
                self._vflat[a:b].copy_(st["exp_avg_sq"].reshape(-1))
                self._state.steps = int(float(st["step"]))
            st["step"] = torch.tensor(float(self._state.steps))
            st["exp_avg"] = self._mflat[a:b].view(t.shape)
            st["exp_avg_sq"] = self._vflat[a:b].view(t.shape)
        self._state.buf[0:1].fill_(self._state.steps)

    def _sync_step_tensors(self):
        for st in self.optimizer.state.values():
            if "step" in st:
                st["step"].fill_(float(self._state.steps))

    def _pack(self):
        return self._flat

    def _scatter_grad(self, flat_grad):
        self._gradflat.copy_(flat_grad)
        for t, g in self._grad_views:
            t.grad = g

    # ------------------------------------------------------------------ training
    def meta_fit(self, valid_tuples=None, verbose=True, log_period=500, n_iter=None):
        """Meta-learns the GP prior parameters -- GPR_meta_mll.py:82-147."""
        assert (valid_tuples is None) or (all([len(valid_tuple) == 4 for valid_tuple in valid_tuples]))
        loss = torch.zeros((), device=self.device)
        # the reference gates on len(self.shared_parameters) > 0, which always holds (the likelihood's noise is always a
        # shared parameter, GPR_meta_mll.py:56): 'vanilla' mode still trains the noise
        if len(self.shared_parameters) > 0:
            t = time.time()
            if n_iter is None:
                n_iter = self.num_iter_fit
            itr = 0
            self._cum_loss.zero_()
            while itr < n_iter:
                nxt = 1 if itr == 0 else min(n_iter, (itr // log_period + 1) * log_period)
                loss = self.run_steps(nxt - itr)
                itr = nxt
                if itr == 1 or itr % log_period == 0:
                    self._failures.check()
                    duration = time.time() - t
                    avg_loss = self._cum_loss / (log_period if itr > 1 else 1.0)     # GPR_meta_mll.py:119-126
                    self._cum_loss.zero_()                                           # in place: a captured graph accumulates into it
                    t = time.time()
                    message = 'Iter %d/%d - Loss: %.6f - Time %.2f sec' % (itr, self.num_iter_fit, avg_loss.item(), duration)
                    if valid_tuples is not None:
                        valid_ll, valid_rmse, calibr_err = self.eval_datasets(valid_tuples)
                        message += ' - Valid-LL: %.3f - Valid-RMSE: %.3f - Calib-Err %.3f' % (valid_ll, valid_rmse, calibr_err)
                    if verbose:
                        self.logger.info(message)
        else:
            self.logger.info('Vanilla mode - nothing to fit')
        self._failures.check()
        self.fitted = True
        return loss.item()

    GRAPH_STEPS = 10

    def _device_step(self, K, idx_stream):
        """One iteration with every step-dependent scalar on the device (capture-safe): batch hand-over + step state,
        batched MLL forward + backward, [cross-rank sum], fused AdamW on the flat parameter buffer (the modules are views of it)."""
        st, D = self._state, self.arch.D
        st.prepare(K, self._idx_cur.numel(), idx_stream, self._idx_cur)
        _, packed, info = self.engine.mll_fwd_bwd(self._flat, self._idx_cur, want_mll=False, want_info=True)
        if getattr(self, "_world", 1) > 1:
            torch.distributed.all_reduce(packed, group=self._group)
        g = self.optimizer.param_groups[0]
        eng.check(eng.lib.pacoh_adamw_step_dev(D, eng._ptr(self._flat), eng._ptr(packed), -1.0, eng._ptr(self._mflat), eng._ptr(self._vflat),
                                               float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]), float(self.weight_decay),
                                               eng._ptr(self._mask), eng._ptr(st.buf), eng._stream()))      # grad = -d sum mll
        self._failures.update(info)
        loss = -packed[D]
        self._cum_loss += loss
        return loss, info

    def map_step(self, task_idx):
        """One iteration of the meta_fit loop body on the given batch (GPR_meta_mll.py:107-117): zero_grad, loss = - sum of the
        batch's MLLs, backward, optimizer step, lr schedule.  Returns the loss (device scalar)."""
        if self._state is None:                    # optimizer='SGD': torch's optimizer on the scattered gradient
            self.optimizer.zero_grad()
            loss = self._loss_and_grad(task_idx)
            self.optimizer.step()
            self.lr_scheduler.step()
            self._cum_loss += loss
            return loss
        idx = np.asarray(task_idx, dtype=np.int32)
        world, rank = getattr(self, "_world", 1), getattr(self, "_rank", 0)
        assert len(idx) >= world, "task batch smaller than the number of ranks"
        lo, hi = eng.shard_bounds(len(idx), rank, world)
        if self._idx_cur is None or self._idx_cur.numel() != hi - lo:
            self._idx_cur = torch.empty(hi - lo, dtype=torch.int32, device=self.device)
            self._graph = None
        loss, self._last_info = self._device_step(1, self._idx_ring.upload(idx[lo:hi]))
        self.lr_scheduler.step()
        return loss

    def run_steps(self, n):
        """``n`` iterations of the meta_fit loop (GPR_meta_mll.py:104-117); with AdamW and equally sized tasks GRAPH_STEPS of them
        replay from one CUDA graph (PACOH_GRAPH=0: always eager), bitwise identical to the eager sequence."""
        loss = None
        K = self.GRAPH_STEPS
        world, rank = getattr(self, "_world", 1), getattr(self, "_rank", 0)
        # single rank only: a sharded run has an NCCL all-reduce in the step, and processes whose captured graphs hold NCCL
        # kernels were seen to hang at process-group teardown (config #5 steps are tens of ms: launch overhead is irrelevant there)
        use_graph = self._state is not None and not self._ragged and world == 1 and os.environ.get("PACOH_GRAPH", "1") != "0"
        while n > 0:
            lo, hi = eng.shard_bounds(self.task_batch_size, rank, world)
            if use_graph and n >= K and self._state.steps > 0 and self._idx_cur is not None and self._idx_cur.numel() == hi - lo:
                if self._graph is None:
                    self._idx_stream = torch.zeros(K, hi - lo, dtype=torch.int32, device=self.device)
                    s0 = self._state.steps
                    self._graph = eng.StepGraph(lambda: self._device_step(K, self._idx_stream), K, self.device)
                    self._state.steps = s0          # the capture only recorded: nothing ran
                s0 = self._state.steps
                rows = np.empty((K, hi - lo), dtype=np.int32)
                for j in range(K):
                    rows[(s0 + j) % K] = self.rds_numpy.choice(len(self.task_dicts), size=self.task_batch_size)[lo:hi]
                self._idx_ring.upload(rows, out=self._idx_stream)
                loss, self._last_info = self._graph.replay()
                self._state.steps += K
                for _ in range(K):
                    self.lr_scheduler.step()
                n -= K
            else:
                loss = self.map_step(self.rds_numpy.choice(len(self.task_dicts), size=self.task_batch_size))
                n -= 1
        return loss

    def shard_tasks(self, group=None):
        """Task-shard every sampled batch over the ranks of ``group``: all ranks hold the same model and draw
        the same batch; each evaluates a contiguous slice and the packed (D + 1) buffer (gradient | sum of MLLs) is summed
        with one NCCL all-reduce."""
        import torch.distributed as dist
        assert dist.is_initialized()
        self._group = group if group is not None else dist.group.WORLD
        self._rank, self._world = dist.get_rank(self._group), dist.get_world_size(self._group)
        return self

    def _loss_and_grad(self, task_idx):
        task_idx = np.asarray(task_idx, dtype=np.int32)
        world, rank = getattr(self, "_world", 1), getattr(self, "_rank", 0)
        if world > 1:
            assert len(task_idx) >= world, "task batch smaller than the number of ranks"
            lo, hi = eng.shard_bounds(len(task_idx), rank, world)
            task_idx = task_idx[lo:hi]
        idx = self._idx_ring.upload(task_idx)
        theta = self._pack()
        _, packed, info = self.engine.mll_fwd_bwd(theta, idx, want_mll=False, want_info=True)
        if world > 1:
            torch.distributed.all_reduce(packed, group=self._group)
        D = self.arch.D
        self._scatter_grad(-packed[:D])                 # loss = - sum_t mll_t
        self._last_info = info
        self._failures.update(info)
        return -packed[D]

    # ------------------------------------------------------------------ prediction
    def predict(self, context_x, context_y, test_x, return_density=False):
        """Posterior inference on the context set, predictive distribution at test_x -- GPR_meta_mll.py:149-190."""
        base = self._predictive(context_x, context_y, test_x, squeeze=True)
        pred = AffineTransformedDistribution(base, normalization_mean=self.y_mean, normalization_std=self.y_std)
        if return_density:
            return pred
        return pred.mean.numpy(), pred.stddev.numpy()

    def _predict_params(self):
        return self._pack()

    # ------------------------------------------------------------------ checkpointing (GPR_meta_mll.py:192-205)
    def _model_state(self):
        sd = OrderedDict()
        sd['likelihood.noise_covar.raw_noise'] = self.raw_noise.detach().clone()
        if self.arch.mean_kind == "constant":
            sd['mean_module.constant'] = self.constant_mean.detach().clone()
        sd['covar_module.raw_outputscale'] = self.raw_outputscale.detach().clone()
        sd['covar_module.base_kernel.raw_lengthscale'] = self.raw_lengthscale.detach().clone()
        for prefix, net in (('learned_kernel', self.nn_kernel_map), ('learned_mean', self.nn_mean_fn)):
            if net is not None:
                for k, v in net.state_dict().items():
                    sd['%s.%s' % (prefix, k)] = v.detach().clone()
        return sd

    def state_dict(self):
        # deep copy: torch >= 2 no longer copies optimizer state tensors in load_state_dict, so a snapshot that aliases
        # the live exp_avg buffers would be shared between two learners
        import copy
        if self._state is not None:
            self._sync_step_tensors()
        return {'optimizer': copy.deepcopy(self.optimizer.state_dict()), 'model': self._model_state()}

    def load_state_dict(self, state_dict):
        sd = state_dict['model']
        with torch.no_grad():
            self.raw_noise.copy_(sd['likelihood.noise_covar.raw_noise'])
            if self.arch.mean_kind == "constant":
                self.constant_mean.copy_(sd['mean_module.constant'])
            self.raw_outputscale.copy_(sd['covar_module.raw_outputscale'])
            self.raw_lengthscale.copy_(sd['covar_module.base_kernel.raw_lengthscale'])
        for prefix, net in (('learned_kernel', self.nn_kernel_map), ('learned_mean', self.nn_mean_fn)):
            if net is not None:
                net.load_state_dict({k[len(prefix) + 1:]: v for k, v in sd.items() if k.startswith(prefix + '.')})
        self.optimizer.load_state_dict(state_dict['optimizer'])
        if self._state is not None:
            self._link_optimizer_state()
            self._graph = None

    # ------------------------------------------------------------------ setup
    def _setup_gp_prior(self, mean_module, covar_module, learning_mode, feature_dim, mean_nn_layers, kernel_nn_layers):
        """Module construction order = RNG consumption order of the reference (GPR_meta_mll.py:207-251): kernel net
        first, then mean net; raw lengthscale (1, F), raw outputscale (), constant mean start at 0."""
        self.shared_parameters = []
        if covar_module == 'NN':
            assert learning_mode in ['learn_kernel', 'both'], 'neural network parameters must be learned'
            self.nn_kernel_map = NeuralNetwork(input_dim=self.input_dim, output_dim=feature_dim,
                                               layer_sizes=kernel_nn_layers).to(self.device)
            self.shared_parameters.append({'params': self.nn_kernel_map.parameters(), 'lr': self.lr_params,
                                           'weight_decay': self.weight_decay})
            n_feat = feature_dim
        else:
            self.nn_kernel_map = None
            n_feat = self.input_dim
        if mean_module == 'NN':
            assert learning_mode in ['learn_mean', 'both'], 'neural network parameters must be learned'
            self.nn_mean_fn = NeuralNetwork(input_dim=self.input_dim, output_dim=1, layer_sizes=mean_nn_layers).to(self.device)
            self.shared_parameters.append({'params': self.nn_mean_fn.parameters(), 'lr': self.lr_params,
                                           'weight_decay': self.weight_decay})
        else:
            self.nn_mean_fn = None
        self.raw_lengthscale = torch.nn.Parameter(torch.zeros(1, n_feat, device=self.device))
        self.raw_outputscale = torch.nn.Parameter(torch.zeros((), device=self.device))
        if mean_module == 'constant':
            self.constant_mean = torch.nn.Parameter(torch.zeros(1, device=self.device))
        if learning_mode in ["learn_kernel", "both"]:
            self.shared_parameters.append({'params': [self.raw_lengthscale, self.raw_outputscale], 'lr': self.lr_params})
        if learning_mode in ["learn_mean", "both"] and mean_module == 'constant':
            self.shared_parameters.append({'params': [self.constant_mean], 'lr': self.lr_params})
        self.arch = eng.GPArch(self.input_dim, mean_kind=mean_module, covar_kind=covar_module,
                               mean_layers=tuple(mean_nn_layers), kernel_layers=tuple(kernel_nn_layers),
                               feature_dim=feature_dim, outputscale=True, noise_floor=self.NOISE_FLOOR)
        self._last_info = None

    def _setup_optimizer(self, optimizer, lr, lr_decay):
        # AdamW's constructor-level weight_decay also applies to the groups that did not set one (likelihood, covar,
        # mean hypers) -- a reference quirk pinned by the demo.ipynb trajectory (GPR_meta_mll.py:56, 248-255)
        self._state, self._graph, self._idx_cur = None, None, None
        self._cum_loss = torch.zeros((), device=self.device)
        if optimizer == 'Adam':
            self.optimizer = _SyncedAdamW(self.shared_parameters, lr=lr, weight_decay=self.weight_decay)
            self.optimizer.owner = self
            self._state = eng.StepState(self.device, lr, lr_decay)       # lr, StepLR(1000, lr_decay), AdamW step count on the device
            self._link_optimizer_state()
        elif optimizer == 'SGD':
            self.optimizer = torch.optim.SGD(self.shared_parameters, lr=lr)
        else:
            raise NotImplementedError('Optimizer must be Adam or SGD')
        if lr_decay < 1.0:
            self.lr_scheduler = torch.optim.lr_scheduler.StepLR(self.optimizer, 1000, gamma=lr_decay)
        else:
            self.lr_scheduler = DummyLRScheduler()

    def _vectorize_pred_dist(self, pred_dist):
        return torch.distributions.Normal(pred_dist.mean, pred_dist.stddev)
