"""Training criteria of the reference's SEModel (baseline_code/d_model.py:24-25,74,80): espnet2's MultiResL1SpecLoss and
SISNRLoss (espnet==202412, un-vendored; behaviour per SURVEY.md Appendix A / §8a row a15), written with differentiable
torch ops so the same code runs on CPU (parity tests against the oracle) and on the GPU (train step).

PARITY UNPINNED against a real espnet install (none available offline): `get_magnitude` is taken to be |.| and the STFTs
use a rectangular window (window=None), hop = w/2, center=True, reflect padding.
"""
from __future__ import annotations

import torch


def si_snr_loss(ref: torch.Tensor, est: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """SISNRLoss(zero_mean=True): negative SI-SNR in dB per item, (B,)   [d_model.py:25,80]."""
    assert ref.shape == est.shape
    ref = ref - ref.mean(dim=-1, keepdim=True)
    est = est - est.mean(dim=-1, keepdim=True)
    energy = torch.sum(ref ** 2, dim=-1, keepdim=True) + eps
    proj = torch.sum(ref * est, dim=-1, keepdim=True) * ref / energy
    noise = est - proj
    ratio = torch.sum(proj ** 2, dim=-1) / (torch.sum(noise ** 2, dim=-1) + eps)
    return -10 * torch.log10(ratio + eps)


def _stft_mag(x: torch.Tensor, w: int) -> torch.Tensor:
    spec = torch.stft(x, n_fft=w, hop_length=w // 2, win_length=w, window=torch.ones(w, dtype=x.dtype, device=x.device),
                      center=True, pad_mode="reflect", normalized=False, onesided=True, return_complex=True)
    return spec.abs()


def multires_l1_spec_loss(target: torch.Tensor, estimate: torch.Tensor, window_sz=(256, 512, 768, 1024), eps: float = 1e-6,
                          time_domain_weight: float = 0.5, normalize_variance: bool = True) -> torch.Tensor:
    """MultiResL1SpecLoss(window_sz=[256,512,768,1024], eps=1e-6, normalize_variance=True, time_domain_weight=0.5),
    reduction "sum" -> (B,)   [d_model.py:24,74].  target / estimate: (B, L)."""
    if normalize_variance:
        target = target / torch.std(target, dim=1, keepdim=True)
        estimate = estimate / torch.std(estimate, dim=1, keepdim=True)
    alpha = torch.sum(estimate * target, -1, keepdim=True) / (torch.sum(estimate ** 2, -1, keepdim=True) + eps)
    scaled = estimate * alpha
    td = (scaled - target).abs().sum(dim=-1)
    sp = torch.zeros_like(td)
    for w in window_sz:
        sp = sp + (_stft_mag(scaled, w) - _stft_mag(target, w)).abs().sum(dim=(1, 2))
    return td * time_domain_weight + (1 - time_domain_weight) * sp / len(window_sz)
