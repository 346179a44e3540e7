"""Synthetic task environments used by the reference's configs (host-side input generators, numpy only).

SinusoidDataset restates experiments/data_sim.py:203-248 of the reference: the same sequence of draws from the same
numpy RandomState, so ``SinusoidDataset(np.random.RandomState(26)).generate_meta_train_data(20, 5)`` returns the
reference's demo data bit for bit (checked against tests/golden/sinusoid.npz).
"""
import numpy as np


class MetaDataset:
    def __init__(self, random_state=None):
        self.random_state = np.random if random_state is None else random_state

    def generate_meta_train_data(self, n_tasks, n_samples):
        raise NotImplementedError

    def generate_meta_test_data(self, n_tasks, n_samples_context, n_samples_test):
        raise NotImplementedError


class SinusoidDataset(MetaDataset):
    """f(x) = slope * x + amp * sin(period * (x - x_shift)) + y_shift, with noisy observations at x ~ U[x_low, x_high]."""

    def __init__(self, amp_low=0.7, amp_high=1.3, period_low=1.5, period_high=1.5, x_shift_mean=0.0, x_shift_std=0.1,
                 y_shift_mean=5.0, y_shift_std=0.1, slope_mean=0.5, slope_std=0.2, noise_std=0.1, x_low=-5, x_high=5,
                 random_state=None):
        super().__init__(random_state)
        assert y_shift_std >= 0 and noise_std >= 0, "std must be non-negative"
        self.amp, self.period = (amp_low, amp_high), (period_low, period_high)
        self.x_shift, self.y_shift, self.slope = (x_shift_mean, x_shift_std), (y_shift_mean, y_shift_std), (slope_mean, slope_std)
        self.noise_std, self.x_low, self.x_high = noise_std, x_low, x_high

    def _sample_sinusoid(self):
        rs = self.random_state
        amplitude = rs.uniform(*self.amp)
        x_shift = rs.normal(loc=self.x_shift[0], scale=self.x_shift[1])
        y_shift = rs.normal(loc=self.y_shift[0], scale=self.y_shift[1])
        slope = rs.normal(loc=self.slope[0], scale=self.slope[1])
        period = rs.uniform(*self.period)
        return lambda x: slope * x + amplitude * np.sin(period * (x - x_shift)) + y_shift

    def _draw(self, n):
        f = self._sample_sinusoid()
        X = self.random_state.uniform(self.x_low, self.x_high, size=(n, 1))
        Y = f(X) + self.noise_std * self.random_state.normal(size=f(X).shape)
        return X, Y

    def generate_meta_train_data(self, n_tasks, n_samples):
        return [self._draw(n_samples) for _ in range(n_tasks)]

    def generate_meta_test_data(self, n_tasks, n_samples_context, n_samples_test):
        assert n_samples_test > 0
        out = []
        for _ in range(n_tasks):
            X, Y = self._draw(n_samples_context + n_samples_test)
            out.append((X[:n_samples_context], Y[:n_samples_context], X[n_samples_context:], Y[n_samples_context:]))
        return out
