"""Host helpers with the reference's semantics (meta_learn/util.py:9-58, 94-100); pure tensor/numpy glue, no kernels."""
import logging
import warnings

import numpy as np
import torch


def find_root_by_bounding(fun, left, right, eps=1e-6, max_iter=1e4):
    """Vectorised bisection on a monotone function (meta_learn/util.py:9-42): used for mixture quantiles."""
    assert callable(fun)
    n_iter, err = 0, 1e12
    middle = (left + right) / 2
    while err > eps:
        middle = (right + left) / 2
        f = fun(middle)
        below = (f < 0).flatten()
        left[below] = middle[below]
        right[~below] = middle[~below]
        assert torch.all(left <= right).item()
        err = torch.max(torch.abs(right - left)) / 2
        n_iter += 1
        if n_iter > max_iter:
            warnings.warn("Max_iter has been reached - stopping bisection for determining quantiles")
            return torch.full_like(left, float("nan"))
    return middle


def _handle_input_dimensionality(x, y=None):
    """meta_learn/util.py:44-58: x -> (n, d), y -> (n, 1)."""
    x = np.asarray(x)
    if x.ndim == 1:
        x = np.expand_dims(x, -1)
    assert x.ndim == 2
    if y is None:
        return x
    y = np.asarray(y)
    if y.ndim == 1:
        y = np.expand_dims(y, -1)
    assert x.shape[0] == y.shape[0]
    assert y.ndim == 2
    return x, y


def get_logger():
    logger = logging.getLogger("gp-priors")
    logger.setLevel(logging.INFO)
    if len(logger.handlers) == 0:
        sh = logging.StreamHandler()
        sh.setFormatter(logging.Formatter("[%(asctime)s -%(levelname)s]  %(message)s"))
        sh.setLevel(logging.INFO)
        logger.addHandler(sh)
        logger.propagate = False
    return logger


class DummyLRScheduler:
    """meta_learn/util.py:94-100."""

    def __init__(self, *args, **kwargs):
        pass

    def step(self, *args, **kwargs):
        pass
