"""Oracle shim (TEST INFRASTRUCTURE): restatement of ``espnet2.enh.layers.bsrnn`` (espnet==202412,
un-vendored dependency of the reference, setup.py:18).  Behaviour per SURVEY.md Appendix A and
§8(a) rows a4-a8; validated by the reference's only known-answer (conf/models/BSRNN_baseline.yaml:30-32,
parameter counts, see tests/test_oracle.py).  Module/parameter names are the ones the published
checkpoints use."""
from itertools import accumulate

import torch
import torch.nn as nn
import torch.nn.functional as F


def choose_norm(norm_type, channel_size, shape="BDTF"):
    if norm_type == "GN":
        return nn.GroupNorm(1, channel_size)
    if norm_type == "BN":
        return nn.BatchNorm2d(channel_size)
    raise ValueError(f"oracle shim only restates GN/BN, got {norm_type}")


def choose_norm1d(norm_type, channel_size):
    # espnet2/enh/layers/bsrnn.py imports this name as ``from espnet2.enh.layers.tcn import choose_norm as
    # choose_norm1d``; tcn.choose_norm builds "GN" as nn.GroupNorm(1, C, eps=1e-8), unlike the 4-D choose_norm
    # above (torch default 1e-5).  Used by BandSplit and the Mask/Grad decoders (reference bsrnn_flowse.py:48,121,128).
    if norm_type == "GN":
        return nn.GroupNorm(1, channel_size, eps=1e-8)
    if norm_type == "BN":
        return nn.BatchNorm1d(channel_size)
    raise ValueError(f"oracle shim only restates GN/BN, got {norm_type}")


def _subbands_for(input_dim, target_fs):
    if input_dim == 481 and target_fs == 48000:
        return (5,) + (4,) * 19 + (10,) * 6 + (40,) * 7 + (60,)
    raise NotImplementedError(f"Please define your own subbands for input_dim={input_dim} and target_fs={target_fs}")


class BandSplit(nn.Module):
    def __init__(self, input_dim, target_fs=48000, channels=128, norm_type="GN"):
        super().__init__()
        assert input_dim % 2 == 1, input_dim
        self.subbands = _subbands_for(input_dim, target_fs)
        assert sum(self.subbands) == input_dim
        freqs = torch.fft.rfftfreq((input_dim - 1) * 2, 1.0 / target_fs)
        self.subband_freqs = freqs[[e - 1 for e in accumulate(self.subbands)]]
        self.norm = nn.ModuleList(choose_norm1d(norm_type, 2 * s) for s in self.subbands)
        self.fc = nn.ModuleList(nn.Conv1d(2 * s, channels, 1) for s in self.subbands)

    def forward(self, x, fs=None):
        # x (B,T,F,2) -> (B,N,T,K')
        outs, lo, nbins = [], 0, x.size(2)
        for i, s in enumerate(self.subbands):
            xb = x[:, :, lo:lo + s, :]
            if xb.size(2) < s:
                xb = F.pad(xb, (0, 0, 0, s - xb.size(2)))
            xb = xb.reshape(xb.size(0), xb.size(1), -1).transpose(1, 2)  # (B, 2s, T), channel = 2*bin+ri
            outs.append(self.fc[i](self.norm[i](xb)))
            lo += s
            if lo >= nbins:
                break
            if fs is not None and self.subband_freqs[i] >= fs / 2:
                break
        return torch.stack(outs, dim=-1)


class MaskDecoder(nn.Module):
    def __init__(self, freq_dim, subbands, channels=128, num_spk=1, norm_type="GN"):
        super().__init__()
        assert freq_dim == sum(subbands), (freq_dim, subbands)
        self.subbands = subbands
        self.freq_dim = freq_dim
        self.num_spk = num_spk

        def mlp(s):
            return nn.Sequential(choose_norm1d(norm_type, channels), nn.Conv1d(channels, 4 * channels, 1),
                                 nn.Tanh(), nn.Conv1d(4 * channels, int(s * 4 * num_spk), 1), nn.GLU(dim=1))
        self.mlp_mask = nn.ModuleList(mlp(s) for s in subbands)
        self.mlp_residual = nn.ModuleList(mlp(s) for s in subbands)

    def forward(self, x):
        # x (B,N,T,K') -> m, r (B,num_spk,T,F,2)
        B, _, T, K = x.shape
        ms, rs = [], []
        for i in range(min(K, len(self.subbands))):
            xb = x[..., i]
            ms.append(self.mlp_mask[i](xb).transpose(1, 2).reshape(B, T, self.num_spk, -1, 2))
            rs.append(self.mlp_residual[i](xb).transpose(1, 2).reshape(B, T, self.num_spk, -1, 2))
        m, r = torch.cat(ms, dim=3), torch.cat(rs, dim=3)
        pad = self.freq_dim - m.size(3)
        m, r = F.pad(m, (0, 0, 0, pad)), F.pad(r, (0, 0, 0, pad))
        return m.moveaxis(1, 2), r.moveaxis(1, 2)


class BSRNN(nn.Module):
    def __init__(self, input_dim=481, num_channel=16, num_layer=6, target_fs=48000, causal=True,
                 num_spk=1, norm_type="GN"):
        super().__init__()
        self.num_layer = num_layer
        self.band_split = BandSplit(input_dim, target_fs=target_fs, channels=num_channel, norm_type=norm_type)
        self.target_fs, self.causal, self.num_spk = target_fs, causal, num_spk
        hdim = 2 * num_channel
        mods = {k: nn.ModuleList() for k in ("norm_time", "rnn_time", "fc_time", "norm_freq", "rnn_freq", "fc_freq")}
        for _ in range(num_layer):
            mods["norm_time"].append(choose_norm(norm_type, num_channel))
            mods["rnn_time"].append(nn.LSTM(num_channel, hdim, batch_first=True, bidirectional=not causal))
            mods["fc_time"].append(nn.Linear(hdim if causal else 2 * hdim, num_channel))
            mods["norm_freq"].append(choose_norm(norm_type, num_channel))
            mods["rnn_freq"].append(nn.LSTM(num_channel, hdim, batch_first=True, bidirectional=True))
            mods["fc_freq"].append(nn.Linear(4 * num_channel, num_channel))
        for k in ("norm_time", "rnn_time", "fc_time", "norm_freq", "rnn_freq", "fc_freq"):
            setattr(self, k, mods[k])
        self.mask_decoder = MaskDecoder(input_dim, self.band_split.subbands, channels=num_channel,
                                        num_spk=num_spk, norm_type=norm_type)

    def forward(self, x, fs=None):
        # x (B,T,F,2) -> (B,num_spk,T,F,2)
        z = self.band_split(x, fs=fs)
        B, N, T, K = z.shape
        skip = z
        for i in range(self.num_layer):
            h = self.norm_time[i](skip).permute(0, 3, 2, 1).reshape(B * K, T, N)
            h = self.fc_time[i](self.rnn_time[i](h)[0])
            skip = skip + h.reshape(B, K, T, N).permute(0, 3, 2, 1)
            h = self.norm_freq[i](skip).permute(0, 2, 3, 1).reshape(B * T, K, N)
            h = self.fc_freq[i](self.rnn_freq[i](h)[0])
            skip = skip + h.reshape(B, T, K, N).permute(0, 3, 1, 2)
        m, r = self.mask_decoder(skip)
        m = torch.view_as_complex(m.contiguous())[..., : x.size(2)]
        r = torch.view_as_complex(r.contiguous())[..., : x.size(2)]
        xc = torch.view_as_complex(x.contiguous())
        return torch.view_as_real(m * xc.unsqueeze(1) + r)
