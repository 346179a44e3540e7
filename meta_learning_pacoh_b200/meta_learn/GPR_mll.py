"""Single-task GP regression with a learned mean / kernel on the B200 engine: call-compatible with the reference's
meta_learn/GPR_mll.py:11-216 (GPRegressionLearned, the "no meta-learning" sibling used as a baseline in
experiments/compuational_comparison.py:12 and tests/test_GPR.py::TestGPR_mll).

It is the PACOH-MAP learner with ONE task that serves both as training set and as context set: every fit iteration is one
pacoh_meta_mll_fwd_bwd call with P = T = 1 (GPR_mll.py:139-144), prediction is the same posterior kernel with the training
data as context (GPR_mll.py:172-196).  Differences to PACOH-MAP that the reference has and this class keeps: the likelihood
is gpytorch's default GaussianLikelihood() (noise floor 1e-4, not 1e-3), the optimizer is AdamW with torch's DEFAULT lr and
weight decay on the groups that set none (GPR_mll.py:101: likelihood / kernel / mean hyper-parameters get weight_decay 0.01),
and the scheduler is ReduceLROnPlateau on the validation log-likelihood (GPR_mll.py:107-110, 158).
"""
import time

import numpy as np
import torch

from .GPR_meta_mll import GPRegressionMetaLearned
from .util import _handle_input_dimensionality


class GPRegressionLearned(GPRegressionMetaLearned):
    NOISE_FLOOR = 1e-4     # gpytorch GaussianLikelihood() default noise constraint GreaterThan(1e-4)

    def __init__(self, train_x, train_t, learning_mode='both', lr=1e-3, weight_decay=0.0, feature_dim=2,
                 num_iter_fit=1000, covar_module='NN', mean_module='NN', mean_nn_layers=(32, 32), kernel_nn_layers=(32, 32),
                 optimizer='Adam', normalize_data=True, lr_scheduler=True, random_seed=None):
        """Args identical to the reference (GPR_mll.py:13-36); gpytorch Kernel / Mean objects are not supported."""
        self._use_plateau_scheduler = lr_scheduler
        self.lr = lr
        train_x, train_t = _handle_input_dimensionality(np.asarray(train_x), np.asarray(train_t))
        self.train_x, self.train_t = train_x, train_t
        super().__init__([(train_x, train_t)], learning_mode=learning_mode, lr_params=lr, weight_decay=weight_decay,
                         feature_dim=feature_dim, num_iter_fit=num_iter_fit, covar_module=covar_module, mean_module=mean_module,
                         mean_nn_layers=mean_nn_layers, kernel_nn_layers=kernel_nn_layers, task_batch_size=1,
                         normalize_data=normalize_data, optimizer=optimizer, lr_decay=1.0, random_seed=random_seed)
        if random_seed is not None:
            np.random.seed(random_seed + 1)           # abstract.py:18-20 of the reference (RegressionModel seeds numpy's global state)
        self.parameters = self.shared_parameters

    def _setup_optimizer(self, optimizer, lr, lr_decay):
        self._state, self._graph, self._idx_cur = None, None, None       # torch's optimizer owns the state here (plateau scheduler)
        self._cum_loss = torch.zeros((), device=self.device)
        if optimizer == 'Adam':
            self.optimizer = torch.optim.AdamW(self.shared_parameters)        # torch defaults for groups that set nothing (GPR_mll.py:101)
        elif optimizer == 'SGD':
            self.optimizer = torch.optim.SGD(self.shared_parameters)
        else:
            raise NotImplementedError('Optimizer must be Adam or SGD')
        self.lr_scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(self.optimizer, mode='max',
                                                                      factor=0.2 if self._use_plateau_scheduler else 1.0)

    def fit(self, valid_x=None, valid_t=None, verbose=True, log_period=500, n_iter=None):
        """Maximises the marginal log-likelihood of the training data -- GPR_mll.py:116-170."""
        assert (valid_x is None and valid_t is None) or (isinstance(valid_x, np.ndarray) and isinstance(valid_t, np.ndarray))
        loss = torch.zeros((), device=self.device)
        if len(self.parameters) > 0:
            t = time.time()
            if n_iter is None:
                n_iter = self.num_iter_fit
            idx = np.zeros(1, dtype=np.int32)
            for itr in range(1, n_iter + 1):
                self.optimizer.zero_grad()
                loss = self._loss_and_grad(idx)
                self.optimizer.step()
                if itr == 1 or itr % log_period == 0:
                    self._failures.check()
                    duration = time.time() - t
                    t = time.time()
                    message = 'Iter %d/%d - Loss: %.3f - Time %.3f sec' % (itr, self.num_iter_fit, loss.item(), duration)
                    if valid_x is not None:
                        valid_ll, valid_rmse, calibr_err = self.eval(valid_x, valid_t)
                        self.lr_scheduler.step(valid_ll)
                        message += ' - Valid-LL: %.3f - Valid-RMSE: %.3f - Calib-Err %.3f' % (valid_ll, valid_rmse, calibr_err)
                    if verbose:
                        self.logger.info(message)
        else:
            self.logger.info('Vanilla mode - nothing to fit')
        self._failures.check()
        self.fitted = True
        return loss.item()

    def meta_fit(self, *args, **kwargs):
        raise AttributeError("GPRegressionLearned has fit(), not meta_fit()")

    # the training data is the context set (GPR_mll.py:172-196, abstract.py:25-57 of the reference)
    def predict(self, test_x, return_density=False, **kwargs):
        test_x = np.asarray(test_x)
        if test_x.ndim == 1:
            test_x = np.expand_dims(test_x, axis=-1)
        return GPRegressionMetaLearned.predict(self, self.train_x, self.train_t, test_x, return_density=return_density)

    def eval(self, test_x, test_t, **kwargs):
        return GPRegressionMetaLearned.eval(self, self.train_x, self.train_t, test_x, test_t)

    def confidence_intervals(self, test_x, confidence=0.9, **kwargs):
        """(ucb, lcb) of the marginal predictive distribution -- abstract.py:50-57 of the reference."""
        pred_dist = self._vectorize_pred_dist(self.predict(test_x, return_density=True))
        alpha = (1 - confidence) / 2
        size = np.asarray(test_x).size
        return pred_dist.icdf(torch.ones(size) * (1 - alpha)), pred_dist.icdf(torch.ones(size) * alpha)
