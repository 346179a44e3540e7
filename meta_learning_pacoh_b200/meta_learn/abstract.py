"""Base class of the meta-learners: seeding, normalisation, per-task tensors and the evaluation API
(eval / eval_datasets / confidence_intervals), call-compatible with meta_learn/abstract.py:117-272 of the reference.

The evaluation path runs in the CUDA kernels of csrc/gp_post.cu: one batched posterior call for all test tasks
(engine.gp_posterior_batch -> pacoh_gp_posterior) and one metrics kernel (pacoh_pred_metrics); the reference builds
gpytorch / torch.distributions objects per task on the CPU and the same numbers fall out of them (abstract.py:157-161,
models.py:15-43, 90-126).
"""
import math
import os

import numpy as np
import torch

from .. import engine as eng
from .distributions import AffineTransformedDistribution, EqualWeightedMixtureDist  # noqa: F401
from .util import _handle_input_dimensionality, get_logger


def default_device():
    """CUDA device of this process: PACOH_DEVICE, else cuda:LOCAL_RANK, else cuda:0 (the reference's config.py:4
    hard-codes cpu; this engine has no CPU path)."""
    name = os.environ.get("PACOH_DEVICE")
    if name:
        return torch.device(name)
    return torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))


class RegressionModelMetaLearned:

    def __init__(self, normalize_data=True, random_seed=None):
        self.normalize_data = normalize_data
        self.logger = get_logger()
        self.input_dim = None
        self.output_dim = None
        self.device = default_device()
        if random_seed is not None:           # abstract.py:125-129
            torch.manual_seed(random_seed)
            self.rds_numpy = np.random.RandomState(random_seed + 1)
        else:
            self.rds_numpy = np.random

    # ------------------------------------------------------------------ public evaluation API
    def predict(self, context_x, context_y, test_x, **kwargs):
        raise NotImplementedError

    def eval(self, context_x, context_y, test_x, test_y, flatten_y=True, **kwargs):
        """(avg test log-likelihood, rmse, calibration error) on one task -- abstract.py:134-163."""
        assert flatten_y, "only scalar targets are supported"
        return tuple(float(v) for v in self._eval_batch([(context_x, context_y, test_x, test_y)], **kwargs)[0])

    def eval_datasets(self, test_tuples, flatten_y=True, **kwargs):
        """Average of eval() over test tasks -- abstract.py:165-181.  All test tasks go through ONE batched posterior call
        (the reference loops over them) unless the learner draws fresh parameter samples per task."""
        assert all(len(t) == 4 for t in test_tuples) and flatten_y
        if self._eval_per_task(**kwargs):
            res = np.concatenate([self._eval_batch([t], **kwargs) for t in test_tuples], axis=0)
        else:
            res = self._eval_batch(list(test_tuples), **kwargs)
        return float(np.mean(res[:, 0])), float(np.mean(res[:, 1])), float(np.mean(res[:, 2]))

    def _eval_per_task(self, **kwargs):
        return False

    def _eval_batch(self, tuples, **kwargs):
        """(len(tuples), 3) array of [ll, rmse, calib]: posterior + metrics kernels (csrc/gp_post.cu)."""
        ctxs, xss, yns = [], [], []
        for cx, cy, tx, ty in tuples:
            cx, cy = _handle_input_dimensionality(cx, cy)
            tx, ty = _handle_input_dimensionality(tx, ty)
            assert tx.shape[1] == cx.shape[1]
            cxn, cyn = self._normalize_data(cx, cy)
            ctxs.append((cxn, cyn.flatten()))
            xss.append(self._normalize_data(X=tx, Y=None))
            yns.append(((ty - self.y_mean[None, :]) / self.y_std[None, :]).flatten())
        with torch.no_grad():
            post = eng.gp_posterior_batch(self.arch, self._predict_params(**kwargs), ctxs, xss, targets=yns)
            out = eng.pred_metrics(post, float(self.y_std[0]))
            if int(post.info.min().item()) < 0:
                raise eng.NotPSDError("a context / predictive kernel matrix is not positive definite")
        return out.cpu().numpy().astype(np.float64)

    def _predict_params(self, **kwargs):
        """(P, D) parameter vectors the predictive distribution mixes over."""
        raise NotImplementedError

    def _predictive(self, context_x, context_y, test_x, squeeze=False, **kwargs):
        """engine.GPPredictive for one task in normalised space."""
        context_x, context_y = _handle_input_dimensionality(context_x, context_y)
        test_x = _handle_input_dimensionality(test_x)
        assert test_x.shape[1] == context_x.shape[1]
        cxn, cyn = self._normalize_data(context_x, context_y)
        with torch.no_grad():
            return eng.GPPredictive(self.arch, self._predict_params(**kwargs), (cxn, cyn.flatten()), self._normalize_data(X=test_x, Y=None), squeeze=squeeze)

    def confidence_intervals(self, context_x, context_y, test_x, confidence=0.9, **kwargs):
        """(ucb, lcb) of the marginal predictive distribution -- abstract.py:183-204."""
        pred_dist = self.predict(context_x, context_y, test_x, return_density=True, **kwargs)
        pred_dist = self._vectorize_pred_dist(pred_dist)
        alpha = (1 - confidence) / 2
        shape = np.asarray(test_x).shape
        ucb = pred_dist.icdf(torch.ones(shape) * (1 - alpha))
        lcb = pred_dist.icdf(torch.ones(shape) * alpha)
        return ucb, lcb

    # ------------------------------------------------------------------ data handling (abstract.py:212-258)
    def _compute_normalization_stats(self, meta_train_tuples):
        xs, ys = zip(*[_handle_input_dimensionality(x, y) for x, y in meta_train_tuples])
        X, Y = np.concatenate(xs, axis=0), np.concatenate(ys, axis=0)
        if self.normalize_data:
            self.x_mean, self.y_mean = np.mean(X, axis=0), np.mean(Y, axis=0)
            self.x_std, self.y_std = np.std(X, axis=0) + 1e-8, np.std(Y, axis=0) + 1e-8
        else:
            self.x_mean, self.y_mean = np.zeros(X.shape[1]), np.zeros(Y.shape[1])
            self.x_std, self.y_std = np.ones(X.shape[1]), np.ones(Y.shape[1])

    def _normalize_data(self, X, Y=None):
        assert hasattr(self, "x_mean") and hasattr(self, "y_std"), "requires computing normalization stats beforehand"
        Xn = (X - self.x_mean[None, :]) / self.x_std[None, :]
        if Y is None:
            return Xn
        return Xn, (Y - self.y_mean[None, :]) / self.y_std[None, :]

    def _check_meta_data_shapes(self, meta_train_data):
        for i in range(len(meta_train_data)):
            meta_train_data[i] = _handle_input_dimensionality(*meta_train_data[i])
        self.input_dim = meta_train_data[0][0].shape[-1]
        self.output_dim = meta_train_data[0][1].shape[-1]
        assert all(self.input_dim == x.shape[-1] and self.output_dim == t.shape[-1] for x, t in meta_train_data)

    def _prepare_data_per_task(self, x_data, y_data, flatten_y=True):
        x_data, y_data = _handle_input_dimensionality(x_data, y_data)
        x_data, y_data = self._normalize_data(x_data, y_data)
        if flatten_y:
            assert y_data.shape[1] == 1
            y_data = y_data.flatten()
        return torch.from_numpy(x_data).float().to(self.device), torch.from_numpy(y_data).float().to(self.device)

    def _build_task_dicts(self, meta_train_data):
        """Per-task tensors (task_dict['train_x'], ['train_y'] as in the reference) plus the stacked, device-resident
        (T, n, d) / (T, n) arrays the kernels index with the sampled task ids.  Tasks may have different numbers of
        points (the reference loops over tasks, random_gp.py:214-217): the arrays are zero-padded to the largest task
        and ``self.task_sizes`` holds every task's own n_t for the kernels (ragged entry point) and for the
        harmonic-mean pre-factor (random_gp.py:209-212)."""
        self.task_dicts = []
        for train_x, train_y in meta_train_data:
            x_t, y_t = self._prepare_data_per_task(train_x, train_y)
            self.task_dicts.append({"train_x": x_t, "train_y": y_t})
        self.task_sizes = np.asarray([td["train_x"].shape[0] for td in self.task_dicts], dtype=np.int64)
        n_max = int(self.task_sizes.max())
        d = self.task_dicts[0]["train_x"].shape[1]
        X = torch.zeros(len(self.task_dicts), n_max, d, dtype=torch.float32)
        Y = torch.zeros(len(self.task_dicts), n_max, dtype=torch.float32)
        for t, td in enumerate(self.task_dicts):
            X[t, :self.task_sizes[t]] = td["train_x"].cpu()
            Y[t, :self.task_sizes[t]] = td["train_y"].cpu()
        return X, Y

    @property
    def _ragged(self):
        return bool(self.task_sizes.min() != self.task_sizes.max())

    def _sample_task_indices(self):
        """rds_numpy.choice(self.task_dicts, size=B) of the reference (GPR_meta_svgd.py:102) draws the same index
        stream as choice(len(task_dicts), size=B): sampling WITH replacement, duplicates count twice."""
        return self.rds_numpy.choice(len(self.task_dicts), size=self.task_batch_size)

    def _calib_error(self, pred_dist_vectorized, test_t_tensor):
        cdf_vals = pred_dist_vectorized.cdf(test_t_tensor)
        if test_t_tensor.shape[0] == 1:
            test_t_tensor, cdf_vals = test_t_tensor.flatten(), cdf_vals.flatten()
        conf = torch.linspace(0.05, 0.95, 20)
        emp = torch.sum(cdf_vals[:, None] <= conf, dim=0).float() / test_t_tensor.shape[0]
        return torch.sqrt(torch.mean((emp - conf) ** 2))

    def _vectorize_pred_dist(self, pred_dist):
        raise NotImplementedError


