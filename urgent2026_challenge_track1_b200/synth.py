"""Deterministic synthetic inputs of the benchmark / smoke workloads: a speech-like harmonic glide with
a 4 Hz syllabic envelope plus 1/f noise at 5 dB SNR, peak <= 0.9 like the simulator's normalisation
(reference simulation/simulate_data_from_param.py:576-581).  Host-side tensor generation only (no model arithmetic)."""
from __future__ import annotations

import math

import torch


def synth_pair(batch, n_samples, fs, seed=1):
    """-> (clean (B, n), noisy (B, n)) float32 on the CPU; utterance b is a function of (seed, b) only."""
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(n_samples, dtype=torch.float64) / fs
    cleans, noisys = [], []
    for b in range(batch):
        f0 = 100 + 150 * (0.5 + 0.5 * torch.sin(2 * math.pi * (0.3 + 0.05 * b) * t))
        phase = 2 * math.pi * torch.cumsum(f0, 0) / fs
        clean = sum(torch.sin(h * phase) / h for h in range(1, 9) if h * 250 < fs / 2)
        env = 0.5 - 0.5 * torch.cos(2 * math.pi * 4 * t + b)
        clean = clean * env
        clean = clean / clean.pow(2).mean().sqrt() * 0.05
        white = torch.randn(n_samples, generator=g, dtype=torch.float64)
        spec = torch.fft.rfft(white)
        k = torch.arange(spec.numel(), dtype=torch.float64).clamp(min=1.0)
        noise = torch.fft.irfft(spec / k.sqrt(), n=n_samples)
        noise = noise / noise.pow(2).mean().sqrt() * 0.05 * 10 ** (-5 / 20)
        x = clean + noise
        scale = 1.0 / max(1.0, float(x.abs().max()) / 0.9)
        cleans.append(clean * scale)
        noisys.append(x * scale)
    return torch.stack(cleans).float(), torch.stack(noisys).float()


def synth_noisy(batch, n_samples, fs, seed=1):
    return synth_pair(batch, n_samples, fs, seed)[1]


def synth_batch(batch, n_samples, fs, seed=1, distinct=4):
    """A (batch, n) noisy batch built from `distinct` generated utterances (generation is the slow part at 64 x 10 s),
    each copy scaled slightly differently so that no two rows are identical."""
    base = synth_noisy(min(batch, distinct), n_samples, fs, seed)
    x = base.repeat((batch + base.size(0) - 1) // base.size(0), 1)[:batch].contiguous()
    return x * (1.0 + 0.01 * torch.arange(batch)[:, None])
