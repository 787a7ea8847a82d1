"""ExponentialMovingAverage with the torch_ema==0.3 behaviour the reference relies on (flow_model.py:53,84,87-112):
shadow copy, warm-up decay min(decay, (1+n)/(10+n)), store / copy_to / restore, state_dict with the same keys so
reference checkpoints' ``ema`` entry loads unchanged."""
from __future__ import annotations

import torch


def _invalidate():
    from .runtime import invalidate_packed
    invalidate_packed()


class ExponentialMovingAverage:
    def __init__(self, parameters, decay, use_num_updates=True):
        if decay < 0.0 or decay > 1.0:
            raise ValueError("Decay must be between 0 and 1")
        self.decay = decay
        self.num_updates = 0 if use_num_updates else None
        self.shadow_params = [p.clone().detach() for p in parameters]
        self.collected_params = None

    @torch.no_grad()
    def update(self, parameters):
        d = self.decay
        if self.num_updates is not None:
            self.num_updates += 1
            d = min(d, (1 + self.num_updates) / (10 + self.num_updates))
        params = list(parameters)
        torch._foreach_lerp_(self.shadow_params, [p.detach() for p in params], 1.0 - d)

    @torch.no_grad()
    def copy_to(self, parameters):
        for s, p in zip(self.shadow_params, parameters):
            p.copy_(s)                       # (not p.data.copy_: the version counter must see the swap)
        _invalidate()

    def store(self, parameters):
        self.collected_params = [p.clone() for p in parameters]

    @torch.no_grad()
    def restore(self, parameters):
        if self.collected_params is None:
            raise RuntimeError("This ExponentialMovingAverage has no `store()`ed weights to `restore()`")
        for c, p in zip(self.collected_params, parameters):
            p.copy_(c)
        _invalidate()

    def to(self, device=None, dtype=None):
        self.shadow_params = [s.to(device=device, dtype=dtype) if s.is_floating_point() else s.to(device=device)
                              for s in self.shadow_params]
        if self.collected_params is not None:
            self.collected_params = [c.to(device=device) for c in self.collected_params]

    def state_dict(self):
        return {"decay": self.decay, "num_updates": self.num_updates, "shadow_params": self.shadow_params,
                "collected_params": self.collected_params}

    def load_state_dict(self, sd):
        self.decay, self.num_updates = sd["decay"], sd["num_updates"]
        self.shadow_params = [t.clone() for t in sd["shadow_params"]]
        self.collected_params = sd["collected_params"]
