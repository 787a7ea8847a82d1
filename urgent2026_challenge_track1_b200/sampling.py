"""ODE samplers — mirrors ``baseline_code/sampling`` (reference sampling/__init__.py:30-65, odesolvers.py:11-107):
the string-keyed ``ODEsolverRegistry`` (euler / midpoint / heun) and ``get_white_box_solver``.

``euler`` through ``get_white_box_solver`` reproduces the reference loop literally (N calls of VF_fn on (B,1,F,T)
complex tensors).  ``FlowSEModel.enhance`` uses the fused path instead (flow_model.py): the state stays in the
(B,T,F,2) kernel layout, ``band_split_y`` is hoisted out of the loop and the Euler update
x <- x + step*(m*x + r) is one kernel fused with the network output.
"""
from __future__ import annotations

import abc
import warnings

import torch


class Registry:
    def __init__(self, managed_thing: str):
        self.managed_thing = managed_thing
        self._registry = {}

    def register(self, name: str):
        def inner_wrapper(wrapped_class):
            if name in self._registry:
                warnings.warn(f"{self.managed_thing} with name '{name}' doubly registered, old class will be replaced.")
            self._registry[name] = wrapped_class
            return wrapped_class
        return inner_wrapper

    def get_by_name(self, name: str):
        if name in self._registry:
            return self._registry[name]
        raise ValueError(f"{self.managed_thing} with name '{name}' unknown.")      # reference odesolvers.py:36

    def get_all_names(self):
        return list(self._registry.keys())


ODEsolverRegistry = Registry("ODEsolver")


class ODEsolver(abc.ABC):
    def __init__(self, ode, VF_fn):
        self.ode, self.VF_fn = ode, VF_fn

    @abc.abstractmethod
    def update_fn(self, x, t, *args):
        ...


@ODEsolverRegistry.register("euler")
class EulerODEsolver(ODEsolver):
    def update_fn(self, x, t, y, stepsize, *args):
        return x + self.VF_fn(x, t, y) * (-stepsize)


@ODEsolverRegistry.register("midpoint")
class MidpointODEsolver(ODEsolver):
    def update_fn(self, x, t, y, stepsize, *args):
        dt = -stepsize
        return x + dt * self.VF_fn(x + dt / 2 * self.VF_fn(x, t, y), t + dt / 2, y)


@ODEsolverRegistry.register("heun")
class HeunODEsolver(ODEsolver):
    def update_fn(self, x, t, y, stepsize, *args):
        dt = -stepsize
        v0 = self.VF_fn(x, t, y)
        return x + dt / 2 * (v0 + self.VF_fn(x + dt * v0, t + dt, y))


def euler_schedule(T_rev, t_eps, N, device=None):
    """t_i = linspace(T_rev, t_eps, N); step_i = t_i - t_{i+1}; the LAST step is t_{N-1} itself
    (integrates down to 0; reference sampling/__init__.py:48-56)."""
    ts = torch.linspace(T_rev, t_eps, N, device=device)
    steps = torch.cat([ts[:-1] - ts[1:], ts[-1:]])
    return ts, steps


def get_white_box_solver(odesolver_name, ode, VF_fn, Y, Y_prior=None, T_rev=1.0, t_eps=0.03, N=30, **kwargs):
    odesolver = ODEsolverRegistry.get_by_name(odesolver_name)(ode, VF_fn)

    def ode_solver(Y_prior=Y_prior):
        with torch.no_grad():
            if Y_prior is None:
                Y_prior = Y
            xt, _ = ode.prior_sampling(Y_prior.shape, Y_prior)
            ts, steps = euler_schedule(T_rev, t_eps, N, device=Y.device)
            xt = xt.to(Y_prior.device)
            for i in range(N):
                vec_t = torch.ones(Y.shape[0], device=Y.device) * ts[i]
                xt = odesolver.update_fn(xt, vec_t, Y, steps[i])
            return xt, N
    return ode_solver
