"""Oracle shim (TEST INFRASTRUCTURE): restatement of torch_ema==0.3 ``ExponentialMovingAverage``
as used by baseline_code/flow_model.py:53,84,87-109."""
import torch


class ExponentialMovingAverage:
    def __init__(self, parameters, decay, use_num_updates=True):
        if not 0.0 <= decay <= 1.0:
            raise ValueError("Decay must be between 0 and 1")
        self.decay = decay
        self.num_updates = 0 if use_num_updates else None
        params = [p for p in parameters]
        self.shadow_params = [p.clone().detach() for p in params]
        self.collected_params = None
        self._n = len(params)

    def update(self, parameters):
        params = list(parameters)
        d = self.decay
        if self.num_updates is not None:
            self.num_updates += 1
            d = min(d, (1 + self.num_updates) / (10 + self.num_updates))
        with torch.no_grad():
            for s, p in zip(self.shadow_params, params):
                s.sub_((1.0 - d) * (s - p))

    def copy_to(self, parameters):
        for s, p in zip(self.shadow_params, parameters):
            p.data.copy_(s.data)

    def store(self, parameters):
        self.collected_params = [p.clone() for p in parameters]

    def restore(self, parameters):
        if self.collected_params is None:
            raise RuntimeError("This ExponentialMovingAverage has no `store()`ed weights to `restore()`")
        for c, p in zip(self.collected_params, parameters):
            p.data.copy_(c.data)

    def to(self, device=None, dtype=None):
        self.shadow_params = [s.to(device=device, dtype=dtype) if s.is_floating_point() else s.to(device=device)
                              for s in self.shadow_params]
        if self.collected_params is not None:
            self.collected_params = [c.to(device=device) for c in self.collected_params]

    def state_dict(self):
        return {"decay": self.decay, "num_updates": self.num_updates,
                "shadow_params": self.shadow_params, "collected_params": self.collected_params}

    def load_state_dict(self, sd):
        self.decay, self.num_updates = sd["decay"], sd["num_updates"]
        self.shadow_params = [t.clone() for t in sd["shadow_params"]]
        self.collected_params = sd["collected_params"]
