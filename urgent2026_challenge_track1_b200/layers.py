"""Parameter containers with the reference's module names and shapes, so that
``load_state_dict`` of a reference checkpoint works unchanged.  They hold nn.Parameters only — the arithmetic is in
the CUDA kernels sequenced by runtime.py — and use torch's stock initialisers so random-init statistics match
nn.LSTM / nn.Linear / nn.Conv1d / nn.GroupNorm defaults of the reference.
"""
from __future__ import annotations

import torch.nn as nn

from .runtime import subbands_for

# GroupNorm eps per construction site (espnet==202412): the 1-D norms of BandSplit and the decoders come from
# ``espnet2.enh.layers.tcn.choose_norm`` (imported as choose_norm1d; "GN" = nn.GroupNorm(1, C, eps=1e-8),
# reference bsrnn_flowse.py:48,121,128); norm_time / norm_freq come from ``espnet2.enh.layers.bsrnn.choose_norm``
# ("GN" = nn.GroupNorm(1, C), torch default 1e-5, reference bsrnn_flowse.py:229,239).  The runtime reads
# ``module.eps`` from these containers, so a different value only needs changing here.
EPS_NORM1D = 1e-8
EPS_NORM = 1e-5


class BandSplitParams(nn.Module):
    """norm.{k} = GroupNorm(1, 2 s_k), fc.{k} = Conv1d(2 s_k, N, 1)   [reference bsrnn_flowse.py:16-50]."""

    def __init__(self, input_dim, target_fs=48000, channels=128):
        super().__init__()
        assert input_dim % 2 == 1, input_dim                               # bsrnn_flowse.py:19
        self.subbands = subbands_for(input_dim, target_fs)
        assert sum(self.subbands) == input_dim, (self.subbands, input_dim)  # bsrnn_flowse.py:42
        self.norm = nn.ModuleList()
        self.fc = nn.ModuleList()
        for s in self.subbands:
            self.norm.append(nn.GroupNorm(1, 2 * s, eps=EPS_NORM1D))
            self.fc.append(nn.Conv1d(2 * s, channels, 1))


class MaskDecoderParams(nn.Module):
    """mlp_mask.{k} / mlp_residual.{k} = Sequential(GN(1,N), Conv1d(N,4N,1), Tanh, Conv1d(4N,4 s_k,1), GLU(dim=1))
    (espnet2 MaskDecoder; indices 0,1,3 carry parameters)."""

    def __init__(self, freq_dim, subbands, channels=128, num_spk=1):
        super().__init__()
        assert freq_dim == sum(subbands), (freq_dim, subbands)
        self.subbands, self.freq_dim, self.num_spk = subbands, freq_dim, num_spk
        self.mlp_mask = nn.ModuleList()
        self.mlp_residual = nn.ModuleList()
        for s in subbands:
            for lst in (self.mlp_mask, self.mlp_residual):
                lst.append(nn.Sequential(nn.GroupNorm(1, channels, eps=EPS_NORM1D), nn.Conv1d(channels, 4 * channels, 1), nn.Tanh(),
                                         nn.Conv1d(4 * channels, int(s * 4 * num_spk), 1), nn.GLU(dim=1)))


class GradDecoderParams(nn.Module):
    """GradDecoder parameters [reference bsrnn_flowse.py:103-134]: per band Sequential(GN, Conv1d(N,16 s), Tanh) for
    mask and residual, plus conv_after_{mask,residual} = Sequential(Conv2d(16,4,5,1,2), GLU(dim=1))."""

    def __init__(self, freq_dim, subbands, channels=128, num_spk=1, sub_channel=16):
        super().__init__()
        assert freq_dim == sum(subbands), (freq_dim, subbands)            # bsrnn_flowse.py:106
        assert num_spk == 1                                               # bsrnn_flowse.py:110
        self.subbands, self.freq_dim, self.num_spk, self.sub_channel = subbands, freq_dim, num_spk, sub_channel
        self.mlp_mask = nn.ModuleList()
        self.mlp_residual = nn.ModuleList()
        self.conv_after_mask = nn.Sequential(nn.Conv2d(sub_channel, 4, 5, 1, 2), nn.GLU(dim=1))
        self.conv_after_residual = nn.Sequential(nn.Conv2d(sub_channel, 4, 5, 1, 2), nn.GLU(dim=1))
        for s in subbands:
            self.mlp_mask.append(nn.Sequential(nn.GroupNorm(1, channels, eps=EPS_NORM1D), nn.Conv1d(channels, s * sub_channel, 1), nn.Tanh()))
            self.mlp_residual.append(nn.Sequential(nn.GroupNorm(1, channels, eps=EPS_NORM1D), nn.Conv1d(channels, s * sub_channel, 1), nn.Tanh()))


def add_dual_path(mod, num_channel, num_layer, with_t_cond=False, t_cond_cls=None):
    """Registers norm_time/rnn_time/fc_time/norm_freq/rnn_freq/fc_freq (+ t_cond) on `mod`
    in the reference's construction order [bsrnn_flowse.py:216-238]."""
    for name in ("norm_time", "rnn_time", "fc_time", "norm_freq", "rnn_freq", "fc_freq"):
        setattr(mod, name, nn.ModuleList())
    if with_t_cond:
        mod.t_cond = nn.ModuleList()
    hdim = 2 * num_channel
    for _ in range(num_layer):
        if with_t_cond:
            mod.t_cond.append(t_cond_cls(num_channel // 2, scale=1))
        mod.norm_time.append(nn.GroupNorm(1, num_channel, eps=EPS_NORM))
        mod.rnn_time.append(nn.LSTM(num_channel, hdim, batch_first=True, bidirectional=True))
        mod.fc_time.append(nn.Linear(2 * hdim, num_channel))
        mod.norm_freq.append(nn.GroupNorm(1, num_channel, eps=EPS_NORM))
        mod.rnn_freq.append(nn.LSTM(num_channel, hdim, batch_first=True, bidirectional=True))
        mod.fc_freq.append(nn.Linear(4 * num_channel, num_channel))
