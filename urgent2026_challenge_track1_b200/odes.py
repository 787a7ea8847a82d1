"""Flow-matching ODE — mirrors ``baseline_code/models/odes.py::FLOWMATCHING`` (reference odes.py:52-98):
mean_t = (1-t) x0 + t y, std_t = (1-t) sigma_min + t sigma_max, prior x_T = y + std(1) z."""
from __future__ import annotations

import warnings

import torch


class FLOWMATCHING:
    def __init__(self, sigma_min=0.00, sigma_max=0.5, **ignored_kwargs):
        self.sigma_min, self.sigma_max = sigma_min, sigma_max

    def copy(self):
        return FLOWMATCHING(self.sigma_min, self.sigma_max)

    def _mean(self, x0, t, y):
        return (1 - t)[:, None, None, None] * x0 + t[:, None, None, None] * y

    def _std(self, t):
        return (1 - t) * self.sigma_min + t * self.sigma_max

    def marginal_prob(self, x0, t, y):
        return self._mean(x0, t, y), self._std(t)

    def prior_sampling(self, shape, y):
        if shape != y.shape:
            warnings.warn(f"Target shape {shape} does not match shape of y {y.shape}! Ignoring target shape.")
        std = self._std(torch.ones((y.shape[0],), device=y.device))
        z = torch.randn_like(y)
        return y + z * std[:, None, None, None], z

    def der_mean(self, x0, t, y):
        return y - x0

    def der_std(self, t):
        return self.sigma_max - self.sigma_min
