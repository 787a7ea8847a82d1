"""Oracle shim (TEST INFRASTRUCTURE, never on the product path).

Restatement of ``espnet2.layers.stft.Stft`` (espnet==202412, pinned by the reference's
setup.py:18; source NOT vendored under /root/reference).  Behaviour follows SURVEY.md
Appendix A; call sites in the reference: baseline_code/models/bsrnn.py:14-25,37,40 and
baseline_code/flow_model.py:26-42,136,145.
"""
import torch


class Stft(torch.nn.Module):
    def __init__(self, n_fft=512, win_length=None, hop_length=128, window="hann",
                 center=True, normalized=False, onesided=True):
        super().__init__()
        self.n_fft = n_fft
        self.win_length = n_fft if win_length is None else win_length
        self.hop_length = hop_length
        self.center = center
        self.normalized = normalized
        self.onesided = onesided
        if window is not None and not hasattr(torch, f"{window}_window"):
            raise ValueError(f"{window} window is not implemented")
        self.window = window

    def _win(self, ref):
        if self.window is None:
            return None
        return getattr(torch, f"{self.window}_window")(self.win_length, dtype=ref.dtype, device=ref.device)

    def forward(self, input, ilens=None):
        spec = torch.stft(input, n_fft=self.n_fft, win_length=self.win_length, hop_length=self.hop_length,
                          center=self.center, window=self._win(input), normalized=self.normalized,
                          onesided=self.onesided, return_complex=True)
        out = torch.view_as_real(spec).transpose(1, 2)  # (B, T, F, 2)
        if ilens is None:
            return out, None
        if self.center:
            ilens = ilens + 2 * (self.n_fft // 2)
        olens = torch.div(ilens - self.n_fft, self.hop_length, rounding_mode="trunc") + 1
        frame_idx = torch.arange(out.size(1), device=out.device)
        dead = frame_idx[None, :] >= olens.to(out.device)[:, None]
        out = out.masked_fill(dead[:, :, None, None], 0.0)
        return out, olens

    def inverse(self, input, ilens=None):
        if not torch.is_complex(input):
            input = torch.complex(input[..., 0], input[..., 1])
        real_dtype = input.real.dtype
        win = None
        if self.window is not None:
            win = getattr(torch, f"{self.window}_window")(self.win_length, dtype=real_dtype, device=input.device)
        wav = torch.istft(input.transpose(1, 2), n_fft=self.n_fft, hop_length=self.hop_length,
                          win_length=self.win_length, window=win, center=self.center,
                          normalized=self.normalized, onesided=self.onesided,
                          length=int(ilens.max()) if ilens is not None else None)
        return wav, ilens
