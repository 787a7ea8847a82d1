"""Training criteria of the reference's SEModel (baseline_code/d_model.py:24-25,74,80): espnet2's MultiResL1SpecLoss and
SISNRLoss (espnet==202412, un-vendored; behaviour per SURVEY.md Appendix A / §8a row a15), On CUDA tensors the L1 / multi-resolution STFT-magnitude core (value and gradient in one pass) is the hand-written
kernel pair of csrc/loss.cu behind ``MultiResL1Core``; the (B,)-sized scale / variance reductions around it stay in the
autograd graph.  ``multires_l1_spec_loss_reference`` is the same criterion in plain differentiable torch ops: it is what
the CPU parity tests check against the oracle and what the GPU tests check the kernels against -- the training step never
calls it.

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


class MultiResL1Core(torch.autograd.Function):
    """(scaled estimate e, target t) (B, L) CUDA f32 ->  w_td * sum|e - t| + (1 - w_td)/len(W) * sum_w sum| |STFT_w e| -
    |STFT_w t| |  per item (B,).  Forward launches bsrnn_l1_time_fwd_bwd + one bsrnn_mrl1_spec_fwd_bwd per window size,
    which also produce d loss / d e; backward is a broadcast multiply.  No gradient flows into the target."""

    @staticmethod
    def forward(ctx, est, tgt, window_sz, time_domain_weight):
        from . import _lib as L
        from .runtime import twiddle
        L.require_device()
        est, tgt = est.contiguous().float(), tgt.contiguous().float()
        B, n = est.shape
        loss = torch.zeros(B, dtype=torch.float64, device=est.device)
        grad = torch.zeros_like(est)
        st = L.stream_ptr()
        L.call("bsrnn_l1_time_fwd_bwd", est.data_ptr(), tgt.data_ptr(), loss.data_ptr(), grad.data_ptr(), B, n,
               float(time_domain_weight), st)
        w_sp = (1.0 - float(time_domain_weight)) / len(window_sz)
        for w in window_sz:
            L.call("bsrnn_mrl1_spec_fwd_bwd", est.data_ptr(), tgt.data_ptr(), loss.data_ptr(), grad.data_ptr(),
                   twiddle(int(w), est.device).data_ptr(), B, n, int(w), w_sp, st)
        ctx.save_for_backward(grad)
        return loss.float()

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g[:, None], None, None, None


def multires_l1_spec_loss(target: torch.Tensor, estimate: torch.Tensor, window_sz=(256, 512, 768, 1024), eps: float = 1e-6,
                          time_domain_weight: float = 0.5, normalize_variance: bool = True) -> torch.Tensor:
    """MultiResL1SpecLoss(window_sz=[256,512,768,1024], eps=1e-6, normalize_variance=True, time_domain_weight=0.5),
    reduction "sum" -> (B,)   [d_model.py:24,74].  target / estimate: (B, L) CUDA tensors (no CPU fallback: the CPU
    checker is multires_l1_spec_loss_reference)."""
    if not (target.is_cuda and estimate.is_cuda):
        from ._lib import NativeLibraryError
        raise NativeLibraryError("multires_l1_spec_loss runs on CUDA tensors only (csrc/loss.cu); the torch restatement "
                                 "for CPU checks is multires_l1_spec_loss_reference")
    if normalize_variance:
        target = target / torch.std(target, dim=1, keepdim=True)
        estimate = estimate / torch.std(estimate, dim=1, keepdim=True)
    alpha = torch.sum(estimate * target, -1, keepdim=True) / (torch.sum(estimate ** 2, -1, keepdim=True) + eps)
    return MultiResL1Core.apply(estimate * alpha, target, tuple(window_sz), time_domain_weight)


def multires_l1_spec_loss_reference(target: torch.Tensor, estimate: torch.Tensor, window_sz=(256, 512, 768, 1024),
                                    eps: float = 1e-6, time_domain_weight: float = 0.5,
                                    normalize_variance: bool = True) -> torch.Tensor:
    """The same criterion in differentiable torch ops (CHECKER for the kernels and the CPU oracle tests; not on the
    training path)."""
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
