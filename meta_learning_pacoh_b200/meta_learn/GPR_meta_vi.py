"""PACOH-VI on the B200 engine: call-compatible with the reference's meta_learn/GPR_meta_vi.py:14-275
(GPRegressionMetaLearnedVI).  Per step (get_neg_elbo, GPR_meta_vi.py:216-224):

    eps ~ N(0, I) (S, D)  ->  pacoh_vi_sample  ->  pacoh_meta_mll_fwd_bwd / pacoh_logprob_finalize  ->  pacoh_vi_grad
    ->  Adam on (loc, scale)

``cov_type='full'`` keeps the D x D Cholesky factor in torch (autograd through the MetaLogProb Function), as the
boundary table in SURVEY 8(b) specifies.
"""
import math
import os
import time

import numpy as np
import torch

from .. import engine as eng
from . import prior_init
from .abstract import RegressionModelMetaLearned
from .distributions import AffineTransformedDistribution, EqualWeightedMixtureDist
from .util import DummyLRScheduler, _handle_input_dimensionality


class _DiagPosterior(torch.nn.Module):
    """Gaussian VI posterior on the GP-prior parameters (RandomGPPosterior, random_gp.py:224-286), diagonal case:
    ``scale`` holds the LOG standard deviation."""

    def __init__(self, loc, scale, entries):
        super().__init__()
        self.loc = torch.nn.Parameter(loc)
        self.scale = torch.nn.Parameter(scale)
        self.param_idx_ranges = entries

    @property
    def mean(self):
        return self.loc.detach()

    mode = mean

    @property
    def stddev(self):
        return self.scale.detach().exp()


class _FullPosterior(torch.nn.Module):
    def __init__(self, loc, tril_cov, entries):
        super().__init__()
        self.loc = torch.nn.Parameter(loc)
        self.tril_cov = torch.nn.Parameter(tril_cov)
        self.param_idx_ranges = entries

    def dist(self):
        return torch.distributions.MultivariateNormal(loc=self.loc, scale_tril=torch.tril(self.tril_cov))

    @property
    def mean(self):
        return self.loc.detach()

    mode = mean

    @property
    def stddev(self):
        return self.dist().stddev.detach()


class GPRegressionMetaLearnedVI(RegressionModelMetaLearned):

    def __init__(self, meta_train_data, num_iter_fit=10000, feature_dim=1,
                 prior_factor=0.01, weight_prior_std=0.5, bias_prior_std=3.0,
                 covar_module='NN', mean_module='NN', mean_nn_layers=(32, 32), kernel_nn_layers=(32, 32),
                 optimizer='Adam', lr=1e-3, lr_decay=1.0, svi_batch_size=10, cov_type='diag',
                 task_batch_size=-1, normalize_data=True, random_seed=None):
        """PACOH-VI: variational inference on the PAC-optimal hyper-posterior with a Gaussian family.
        Args identical to the reference (GPR_meta_vi.py:16-46)."""
        super().__init__(normalize_data, random_seed)
        assert mean_module in ['NN', 'constant', 'zero']
        assert covar_module in ['NN', 'SE']
        assert optimizer in ['Adam', 'SGD']
        assert cov_type in ['diag', 'full']

        self.num_iter_fit, self.prior_factor, self.feature_dim = num_iter_fit, prior_factor, feature_dim
        self.weight_prior_std, self.bias_prior_std = weight_prior_std, bias_prior_std
        self.svi_batch_size = svi_batch_size
        self.cov_type = cov_type
        if task_batch_size < 1:
            self.task_batch_size = len(meta_train_data)
        else:
            self.task_batch_size = min(task_batch_size, len(meta_train_data))

        self._check_meta_data_shapes(meta_train_data)
        self._compute_normalization_stats(meta_train_data)
        self._setup_model_inference(mean_module, covar_module, mean_nn_layers, kernel_nn_layers, cov_type)
        self._setup_optimizer(optimizer, lr, lr_decay)

        X, Y = self._build_task_dicts(meta_train_data)
        self.engine = eng.MetaMLLEngine(self.arch, X, Y, self.device, task_n=self.task_sizes)
        self._group, self._rank, self._world = None, 0, 1
        self._last_info = None
        self._idx_ring = eng.PinnedRing(self.device)
        self._failures = eng.FailureFlag(self.device)
        self.fitted = False

    def shard_tasks(self, group=None):
        """Task-shard the sampled batch over ranks (see GPRegressionMetaLearnedSVGD.shard_tasks)."""
        import torch.distributed as dist
        assert dist.is_initialized()
        self._group = group if group is not None else dist.group.WORLD
        self._rank, self._world = dist.get_rank(self._group), dist.get_world_size(self._group)
        return self

    # ------------------------------------------------------------------ training
    def meta_fit(self, valid_tuples=None, verbose=True, log_period=500, n_iter=None):
        """Fits the variational hyper-posterior by minimising the negative ELBO -- GPR_meta_vi.py:84-128."""
        assert (valid_tuples is None) or (all([len(valid_tuple) == 4 for valid_tuple in valid_tuples]))
        t = time.time()
        if n_iter is None:
            n_iter = self.num_iter_fit
        loss = None
        itr = 0
        while itr < n_iter:
            nxt = 1 if itr == 0 else min(n_iter, (itr // log_period + 1) * log_period)
            loss = self.run_steps(nxt - itr)
            itr = nxt
            if itr == 1 or itr % log_period == 0:
                self._failures.check()
                duration = time.time() - t
                t = time.time()
                message = 'Iter %d/%d - Loss: %.6f - Time %.2f sec' % (itr, self.num_iter_fit, loss.item(), duration)
                if valid_tuples is not None:
                    valid_ll, valid_rmse, calibr_err = self.eval_datasets(valid_tuples)
                    message += ' - Valid-LL: %.3f - Valid-RMSE: %.3f - Calib-Err %.3f' % (valid_ll, valid_rmse, calibr_err)
                if verbose:
                    self.logger.info(message)
        self._failures.check()
        self.fitted = True
        return loss.item()

    # ------------------------------------------------------------------ device-state / graph-captured steps (diag, Adam)
    GRAPH_STEPS = 10

    def _device_step(self, K, idx_stream, eps_stream, pre):
        """One optimisation step whose step-dependent inputs (Adam state, learning rate, task indices, normal draws) are all
        read from device memory: the same call sequence eager or replayed from a CUDA graph."""
        st, post = self._state, self.posterior
        st.prepare(K, self._idx_cur.numel(), idx_stream, self._idx_cur, eps_stream, self._eps_cur)
        theta, logq = eng.vi_sample(post.loc.detach(), post.scale.detach(), self._eps_cur)
        logp, g, info = eng.meta_log_prob_and_score(theta, self.engine, self._idx_cur, self._prior_mu, self._prior_sigma,
                                                    self.prior_factor, pre, self._group)
        dloc, dscale = eng.vi_grad(post.scale.detach(), self._eps_cur, g, self.prior_factor)
        post.loc.grad, post.scale.grad = dloc, dscale
        self.optimizer.step(state=st)
        self._failures.update(info)
        return -(logp - self.prior_factor * logq).mean(), info

    def run_steps(self, n):
        """``n`` iterations of the meta_fit loop body (GPR_meta_vi.py:103-110): sample a batch, -ELBO and its gradient,
        optimizer step, lr schedule.  Diagonal posterior + Adam: GRAPH_STEPS steps at a time replay from one CUDA graph
        (task indices from the numpy stream and normal draws from torch's CPU generator are pre-drawn in the reference's
        order and uploaded per graph); otherwise the reference's eager sequence.  Returns the last loss."""
        loss = None
        dev_path = self._state is not None and self.cov_type == 'diag'
        K = self.GRAPH_STEPS
        use_graph = dev_path and not self._ragged and self._world == 1 and os.environ.get("PACOH_GRAPH", "1") != "0"
        S, D = self.svi_batch_size, self.arch.D
        while n > 0:
            if not dev_path:
                loss = self.vi_step(self._sample_task_indices())
                n -= 1
                continue
            lo, hi = eng.shard_bounds(self.task_batch_size, self._rank, self._world)
            if use_graph and n >= K and self._state.steps > 0 and self._idx_cur is not None and self._idx_cur.numel() == hi - lo:
                if self._graph is None:
                    self._idx_stream = torch.zeros(K, hi - lo, dtype=torch.int32, device=self.device)
                    self._eps_stream = torch.zeros(K, S * D, dtype=torch.float32, device=self.device)
                    pre = eng.pre_factor(self.task_sizes[:1].repeat(self.task_batch_size))
                    s0 = self._state.steps
                    self._graph = eng.StepGraph(lambda: self._device_step(K, self._idx_stream, self._eps_stream, pre), K, self.device)
                    self._state.steps = s0
                s0 = self._state.steps
                rows = np.empty((K, hi - lo), dtype=np.int32)
                eps = np.empty((K, S * D), dtype=np.float32)
                for j in range(K):
                    rows[(s0 + j) % K] = self._sample_task_indices()[lo:hi]
                    eps[(s0 + j) % K] = torch.empty(S, D).normal_().reshape(-1).numpy()   # Normal.rsample -> _standard_normal
                self._idx_ring.upload(rows, out=self._idx_stream)
                self._eps_ring.upload(eps, out=self._eps_stream)
                loss, self._last_info = self._graph.replay()
                self._state.steps += K
                for _ in range(K):
                    self.lr_scheduler.step()
                self.optimizer.sync_from(self._state)
                n -= K
            else:
                loss = self.vi_step(self._sample_task_indices())
                n -= 1
        return loss

    def vi_step(self, task_idx):
        """One iteration of the meta_fit loop body on the given batch (GPR_meta_vi.py:106-110): zero_grad, -ELBO and its
        gradient, optimizer step, lr schedule.  Returns the loss (device scalar)."""
        if self._state is None or self.cov_type != 'diag':
            self.optimizer.zero_grad()
            loss = self.get_neg_elbo(task_idx)
            self.optimizer.step()
            self.lr_scheduler.step()
            return loss
        S, D = self.svi_batch_size, self.arch.D
        idx = np.asarray(task_idx, dtype=np.int32)
        assert idx.shape[0] >= self._world, "task batch smaller than the number of ranks"
        lo, hi = eng.shard_bounds(idx.shape[0], self._rank, self._world)
        if self._idx_cur is None or self._idx_cur.numel() != hi - lo:
            self._idx_cur = torch.empty(hi - lo, dtype=torch.int32, device=self.device)
            self._eps_cur = torch.empty(S, D, dtype=torch.float32, device=self.device)
            self._eps_ring = eng.PinnedRing(self.device, dtype=torch.float32)
            self._graph = None
        pre = eng.pre_factor(self.task_sizes[idx])
        istream = self._idx_ring.upload(idx[lo:hi])
        estream = self._eps_ring.upload(torch.empty(S, D).normal_().reshape(-1).numpy())      # Normal.rsample -> _standard_normal
        loss, self._last_info = self._device_step(1, istream, estream, pre)
        self.lr_scheduler.step()
        self.optimizer.sync_from(self._state)
        return loss

    def _shard(self, task_idx):
        idx = np.asarray(task_idx, dtype=np.int32)
        T = idx.shape[0]
        assert T >= self._world, "task batch (%d) smaller than the number of ranks (%d): every rank needs a task" % (T, self._world)
        lo, hi = eng.shard_bounds(T, self._rank, self._world)
        return self._idx_ring.upload(idx[lo:hi]), eng.pre_factor(self.task_sizes[idx])

    def get_neg_elbo(self, task_idx, eps=None):
        """-mean_s [ log p(theta_s | data) - prior_factor * log q(theta_s) ] and its gradient w.r.t. the posterior
        parameters (written to ``.grad``) -- the closure get_neg_elbo + loss.backward(), GPR_meta_vi.py:106-108, 216-224.
        ``eps`` (S, D): optional caller-supplied standard-normal draws (parity tests); by default they come from
        torch's global CPU generator exactly like ``posterior.rsample`` does in the reference."""
        S, D = self.svi_batch_size, self.arch.D
        idx_dev, pre = self._shard(task_idx)
        if eps is None:
            eps = torch.empty(S, D).normal_()                                   # Normal.rsample -> _standard_normal
        eps = eps.to(self.device)
        if self.cov_type == 'diag':
            post = self.posterior
            theta, logq = eng.vi_sample(post.loc.detach(), post.scale.detach(), eps)
            logp, g, info = eng.meta_log_prob_and_score(theta, self.engine, idx_dev, self._prior_mu, self._prior_sigma,
                                                        self.prior_factor, pre, self._group)
            dloc, dscale = eng.vi_grad(post.scale.detach(), eps, g, self.prior_factor)
            post.loc.grad, post.scale.grad = dloc, dscale
            loss = -(logp - self.prior_factor * logq).mean()
        else:
            q = self.posterior.dist()
            theta = self.posterior.loc + eps @ torch.tril(self.posterior.tril_cov).T      # MultivariateNormal.rsample
            logp, info = eng.meta_log_prob(theta, self.engine, idx_dev, self._prior_mu, self._prior_sigma,
                                           self.prior_factor, pre, self._group)
            loss = -(logp - self.prior_factor * q.log_prob(theta)).mean()
            loss.backward()
            loss = loss.detach()
        self._last_info = info
        self._failures.update(info)
        return loss

    # ------------------------------------------------------------------ prediction
    def predict(self, context_x, context_y, test_x, n_posterior_samples=100, mode='Bayes', return_density=False):
        """Predictive distribution p(y | test_x, context) -- GPR_meta_vi.py:130-174."""
        assert mode in ['bayes', 'Bayes', 'MAP', 'map']
        bayes = mode in ('Bayes', 'bayes')
        base = self._predictive(context_x, context_y, test_x, squeeze=not bayes, n_posterior_samples=n_posterior_samples, mode=mode)
        pred_dist = AffineTransformedDistribution(base, normalization_mean=self.y_mean, normalization_std=self.y_std)
        if bayes:
            pred_dist = EqualWeightedMixtureDist(pred_dist, batched=True)
        if return_density:
            return pred_dist
        return pred_dist.mean.numpy(), pred_dist.stddev.numpy()

    def _sample_posterior(self, n):
        """posterior.sample((n,)) from the CPU generator (GPR_meta_vi.py:236)."""
        if self.cov_type == 'diag':
            loc, std = self.posterior.loc.detach().cpu(), self.posterior.scale.detach().exp().cpu()
            return torch.normal(loc.expand(n, -1), std.expand(n, -1)).to(self.device)
        with torch.no_grad():
            eps = torch.empty(n, self.arch.D).normal_().to(self.device)
            return (self.posterior.loc + eps @ torch.tril(self.posterior.tril_cov).T).contiguous()

    def _predict_params(self, n_posterior_samples=100, mode='Bayes'):
        """'Bayes': n_posterior_samples draws from the variational posterior (GPR_meta_vi.py:229-239); 'MAP': its mode (:241-252)."""
        assert mode in ['bayes', 'Bayes', 'MAP', 'map']
        with torch.no_grad():
            if mode in ('Bayes', 'bayes'):
                return self._sample_posterior(n_posterior_samples).contiguous()
            return self.posterior.mode.view(1, -1).contiguous()

    def _eval_per_task(self, n_posterior_samples=100, mode='Bayes'):
        return mode in ('Bayes', 'bayes')      # the reference draws fresh posterior samples for every test task (predict per task)

    # ------------------------------------------------------------------ setup
    def _setup_model_inference(self, mean_module_str, covar_module_str, mean_nn_layers, kernel_nn_layers, cov_type):
        assert mean_module_str in ['NN', 'constant']
        assert covar_module_str in ['NN', 'SE']
        self.arch = eng.GPArch(self.input_dim, mean_kind=mean_module_str, covar_kind=covar_module_str,
                               mean_layers=tuple(mean_nn_layers), kernel_layers=tuple(kernel_nn_layers), feature_dim=2)
        prior_init.consume_vectorized_gp_init(self.arch)                          # RandomGPMeta.__init__ RNG draws
        mu, sigma = self.arch.hyper_prior(self.weight_prior_std, self.bias_prior_std)
        self._prior_mu, self._prior_sigma = mu.to(self.device), sigma.to(self.device)
        if cov_type == 'diag':
            loc, scale = prior_init.init_diag_posterior(self.arch.D)
            self.posterior = _DiagPosterior(loc.to(self.device), scale.to(self.device), self.arch.entries())
        else:
            loc, tril = prior_init.init_full_posterior(self.arch.D)
            self.posterior = _FullPosterior(loc.to(self.device), tril.to(self.device), self.arch.entries())

    def _setup_optimizer(self, optimizer, lr, lr_decay):
        self._state, self._graph, self._idx_cur = None, None, None
        if optimizer == 'Adam':
            self.optimizer = eng.PacohAdam(self.posterior.parameters(), lr=lr)
            self._state = eng.StepState(self.device, lr, lr_decay)
        elif optimizer == 'SGD':
            self.optimizer = torch.optim.SGD(self.posterior.parameters(), lr=lr)
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
