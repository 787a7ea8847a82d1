"""TEST INFRASTRUCTURE: import the reference's own files *verbatim* from /root/reference on top of the
third-party shim in oracle/shim (espnet2 subset, pytorch_lightning, torch_ema, matplotlib).

Only usable in the build container (``/root/reference`` does not exist on the GPU box); used to pin
``oracle/restated.py`` and to generate ``tests/golden/*.npz`` (tests/golden/make_golden.py).
"""
import importlib
import os
import sys

REFERENCE_ROOT = os.environ.get("URGENT_REFERENCE_ROOT", "/root/reference")
SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "baseline_code"))


def load():
    """Returns a namespace with the reference classes (BSRNN_SE, FlowBSRNN, FlowSEModel, SEModel, Config, ...)."""
    if not available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    for p in (REFERENCE_ROOT, SHIM):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    import torch
    if not torch.cuda.is_available():
        # flow_model.py:194 hard-codes Y.cuda(); identity on a CPU-only box.
        torch.Tensor.cuda = lambda self, *a, **k: self
    ns = type("ref", (), {})()
    ns.bsrnn = importlib.import_module("baseline_code.models.bsrnn")
    ns.bsrnn_flowse = importlib.import_module("baseline_code.models.bsrnn_flowse")
    ns.odes = importlib.import_module("baseline_code.models.odes")
    ns.sampling = importlib.import_module("baseline_code.sampling")
    ns.config = importlib.import_module("baseline_code.config")
    ns.d_model = importlib.import_module("baseline_code.d_model")
    ns.flow_model = importlib.import_module("baseline_code.flow_model")
    ns.BSRNN_SE = ns.bsrnn.BSRNN_SE
    ns.FlowBSRNN = ns.bsrnn_flowse.BSRNN
    ns.SEModel = ns.d_model.SEModel
    ns.FlowSEModel = ns.flow_model.FlowSEModel
    ns.Config = ns.config.Config
    return ns


def flowse_config(ns, **over):
    """Config carrying conf/models/BSRNN_flowse.yaml:31-53 values."""
    cfg = ns.Config(model_type="flowse", ema_decay=0.999, theta=1.5, sigma_max=0.5, sigma_min=0.05, t_eps=0.03,
                    T_rev=1.0, loss_type="mse", loss_abs_exponent=0.5, n_fft=1536, hop_length=384,
                    spec_transform_type="exponent", spec_abs_exponent=0.667, spec_factor=0.065,
                    bsrnn_hidden=384, num_layer=6, learning_rate=1e-4)
    for k, v in over.items():
        setattr(cfg, k, v)
    return cfg
