"""Oracle shim (TEST INFRASTRUCTURE): the thinnest stand-in for pytorch_lightning==2.5.2 that lets the
reference's LightningModules (baseline_code/d_model.py:12, flow_model.py:17) be constructed and called."""
import random
import numpy as np
import torch


class LightningModule(torch.nn.Module):
    def save_hyperparameters(self, *a, **k):
        pass

    def log(self, *a, **k):
        pass

    def optimizer_step(self, epoch, batch_idx, optimizer, optimizer_closure=None):
        optimizer.step(closure=optimizer_closure)


class LightningDataModule:
    pass


def seed_everything(seed=0, workers=False):
    random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
    return seed
