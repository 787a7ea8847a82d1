"""TEST INFRASTRUCTURE: import the reference's own files *verbatim* on top of the third-party shim in oracle/shim
(espnet2 subset, pytorch_lightning, torch_ema, matplotlib).

Source of the files: ``/root/reference`` in the build container; on the GPU box (where that path does not exist) the
git-ignored snapshot ``oracle/_ref/`` written by ``oracle/make_ref.sh`` (it travels with the gpurun snapshot, it is
never committed).  Used to pin ``oracle/restated.py``, to generate ``tests/golden/*.npz``
(tests/golden/make_golden.py), as the full-width checker of the GPU parity tests and as bench.py's reference arm.
"""
import importlib
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
SHIM = os.path.join(_HERE, "shim")
SNAPSHOT = os.path.join(_HERE, "_ref")


def _pick_root():
    for cand in (os.environ.get("URGENT_REFERENCE_ROOT"), "/root/reference", SNAPSHOT):
        if cand and os.path.isfile(os.path.join(cand, "baseline_code", "models", "bsrnn_flowse.py")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _pick_root()


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "baseline_code", "models", "bsrnn_flowse.py"))


def kind():
    """'reference' when the verbatim files are importable (tree or snapshot), else 'port' (oracle/restated.py only)."""
    return "reference" if available() else "port"


def load():
    """Returns a namespace with the reference classes (BSRNN_SE, FlowBSRNN, FlowSEModel, SEModel, Config, ...)."""
    if not available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    # the product's checkpoint module may have registered an ALIAS package ``baseline_code`` (empty __path__) so that
    # pickled reference Configs resolve without the reference; the real package must win here
    pkg = sys.modules.get("baseline_code")
    if pkg is not None and list(getattr(pkg, "__path__", [])) == []:
        for name in [n for n in sys.modules if n == "baseline_code" or n.startswith("baseline_code.")]:
            del sys.modules[name]
    for p in (REFERENCE_ROOT, SHIM):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    import torch
    if not torch.cuda.is_available():
        # flow_model.py:194 hard-codes Y.cuda(); identity on a CPU-only box.
        torch.Tensor.cuda = lambda self, *a, **k: self
    if "torchaudio" not in sys.modules:
        try:                                   # d_model.py / flow_model.py import torchaudio but never use it
            import torchaudio  # noqa: F401
        except Exception:                      # a CUDA-mismatched torchaudio build must not take the oracle down
            import types
            sys.modules["torchaudio"] = types.ModuleType("torchaudio")
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


class on_cpu:
    """Context manager: run the reference on the CPU of a box that HAS a GPU.  flow_model.py:194 hard-codes
    ``Y.cuda()`` inside ``enhance``; while the context is active ``torch.Tensor.cuda`` is the identity, so the CPU-resident
    reference model keeps its tensors where its weights are.  (Never active while product code runs.)"""

    def __enter__(self):
        import torch
        self._saved = torch.Tensor.cuda
        torch.Tensor.cuda = lambda t, *a, **k: t
        return self

    def __exit__(self, *exc):
        import torch
        torch.Tensor.cuda = self._saved


def flowse_config(ns, **over):
    """Config carrying conf/models/BSRNN_flowse.yaml:31-53 values."""
    cfg = ns.Config(model_type="flowse", ema_decay=0.999, theta=1.5, sigma_max=0.5, sigma_min=0.05, t_eps=0.03,
                    T_rev=1.0, loss_type="mse", loss_abs_exponent=0.5, n_fft=1536, hop_length=384,
                    spec_transform_type="exponent", spec_abs_exponent=0.667, spec_factor=0.065,
                    bsrnn_hidden=384, num_layer=6, learning_rate=1e-4)
    for k, v in over.items():
        setattr(cfg, k, v)
    return cfg
