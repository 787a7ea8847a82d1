"""SEModel — inference-side mirror of ``baseline_code/d_model.py::SEModel`` (reference d_model.py:12-113): selector
``cfg.se_model == "bsrnn"`` -> ``BSRNN_SE(**cfg.model_configs)`` under the attribute ``se_model`` (so checkpoint keys are
``se_model.bsrnn.bsrnn.*``), TypeError for anything else (reference :19-23)."""
from __future__ import annotations

import torch
import torch.nn as nn

from .bsrnn import BSRNN_SE


class SEModel(nn.Module):
    def __init__(self, cfg, precision=None):
        super().__init__()
        self.cfg = cfg
        if cfg.se_model == "bsrnn":
            self.se_model = BSRNN_SE(**cfg.model_configs, precision=precision)
        else:
            self.se_model = None
            raise TypeError

    @torch.no_grad()
    def forward(self, noisy_speech, speech_length, fs):
        return self.se_model(noisy_speech, speech_length, fs)

    @classmethod
    def load_from_checkpoint(cls, ckpt_path, map_location="cuda", precision=None):
        """Reads a Lightning-style .ckpt (``state_dict`` + ``hyper_parameters['cfg']``) or a raw state_dict
        (reference train_se.py:55-60 accepts both for init_from)."""
        ckpt = torch.load(ckpt_path, map_location="cpu", weights_only=False)
        sd = ckpt["state_dict"] if "state_dict" in ckpt else ckpt
        cfg = ckpt.get("hyper_parameters", {}).get("cfg") if isinstance(ckpt, dict) else None
        if cfg is None:
            from .config import Config
            n = sd["se_model.bsrnn.bsrnn.fc_time.0.bias"].numel()
            layers = 1 + max(int(k.split(".")[4]) for k in sd if k.startswith("se_model.bsrnn.bsrnn.fc_time."))
            cfg = Config(se_model="bsrnn", model_configs={"num_channel": n, "num_layer": layers})
        model = cls(cfg, precision=precision)
        model.load_state_dict(sd)
        return model.to(map_location)
