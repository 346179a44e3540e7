"""Predictive distribution objects returned by ``predict(..., return_density=True)``.

Same public surface as the reference's meta_learn/models.py:15-43 (affine un-normalisation of the GP predictive) and
models.py:74-140 (equally weighted mixture over particles / posterior samples): ``mean``, ``stddev``, ``variance``,
``log_prob``, ``cdf``, ``icdf``.  They wrap small CPU tensors copied back from the device.
"""
import math

import torch
from torch.distributions import AffineTransform, Distribution, TransformedDistribution

from .util import find_root_by_bounding


class AffineTransformedDistribution(TransformedDistribution):
    """y = normalization_mean + normalization_std * z,  z ~ base_dist."""

    def __init__(self, base_dist, normalization_mean, normalization_std):
        self.loc_tensor = torch.as_tensor(normalization_mean, dtype=torch.float32).reshape((1,))
        self.scale_tensor = torch.as_tensor(normalization_std, dtype=torch.float32).reshape((1,))
        super().__init__(base_dist, AffineTransform(loc=self.loc_tensor, scale=self.scale_tensor))

    @property
    def mean(self):
        return self.loc_tensor + self.scale_tensor * self.base_dist.mean

    @property
    def stddev(self):
        return self.base_dist.stddev * self.scale_tensor

    @property
    def variance(self):
        return self.base_dist.variance * self.scale_tensor ** 2


class EqualWeightedMixtureDist(Distribution):
    """Mixture with weights 1/K over a batch (``batched=True``: leading batch dim) or a list of distributions."""

    def __init__(self, dists, batched=False, num_dists=None):
        self.batched = batched
        if batched:
            assert isinstance(dists, Distribution)
            self.num_dists = dists.batch_shape[0] if num_dists is None else num_dists
            event_shape = dists.event_shape
        else:
            assert all(isinstance(d, Distribution) for d in dists)
            self.num_dists = len(dists)
            event_shape = dists[0].event_shape
        self.dists = dists
        super().__init__(event_shape=event_shape, validate_args=False)

    def _stack(self, attr):
        if self.batched:
            return getattr(self.dists, attr)
        return torch.stack([getattr(d, attr) for d in self.dists], dim=0)

    @property
    def mean(self):
        return self._stack("mean").mean(dim=0)

    @property
    def variance(self):
        means, variances = self._stack("mean"), self._stack("variance")
        return ((means - means.mean(dim=0)) ** 2).mean(dim=0) + variances.mean(dim=0)

    @property
    def stddev(self):
        return torch.sqrt(self.variance)

    @property
    def arg_constraints(self):
        return {}

    def log_prob(self, value):
        lp = self.dists.log_prob(value) if self.batched else torch.stack([d.log_prob(value) for d in self.dists])
        return torch.logsumexp(lp, dim=0) - math.log(float(self.num_dists))

    def cdf(self, value):
        c = self.dists.cdf(value) if self.batched else torch.stack([d.cdf(value) for d in self.dists])
        assert c.shape[0] == self.num_dists
        return c.mean(dim=0)

    def icdf(self, quantile):
        left = -1e8 * torch.ones(quantile.shape)
        right = 1e8 * torch.ones(quantile.shape)
        return find_root_by_bounding(lambda v: self.cdf(v) - quantile, left, right)
