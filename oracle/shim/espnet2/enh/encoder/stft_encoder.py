"""Oracle shim (TEST INFRASTRUCTURE): restatement of espnet2 ``STFTEncoder`` (espnet==202412),
behaviour per SURVEY.md Appendix A.  Reference call sites: baseline_code/models/bsrnn.py:14-19,37;
baseline_code/flow_model.py:26-34,136."""
import torch
from espnet2.layers.stft import Stft


def _reconfig(stft, base, fs, default_fs):
    n_fft, win, hop = base
    stft.n_fft = n_fft * fs // default_fs
    stft.win_length = win * fs // default_fs
    stft.hop_length = hop * fs // default_fs


class STFTEncoder(torch.nn.Module):
    def __init__(self, n_fft=512, win_length=None, hop_length=128, window="hann", center=True,
                 normalized=False, onesided=True, use_builtin_complex=True, default_fs=16000,
                 spec_transform_type=None, spec_factor=0.15, spec_abs_exponent=0.5):
        super().__init__()
        self.stft = Stft(n_fft=n_fft, win_length=win_length, hop_length=hop_length, window=window,
                         center=center, normalized=normalized, onesided=onesided)
        self._output_dim = n_fft // 2 + 1 if onesided else n_fft
        self.use_builtin_complex = use_builtin_complex
        self.win_length = win_length if win_length else n_fft
        self.hop_length = hop_length
        self.window = window
        self.n_fft = n_fft
        self.center = center
        self.default_fs = default_fs
        self.spec_transform_type = spec_transform_type
        self.spec_factor = spec_factor
        self.spec_abs_exponent = spec_abs_exponent

    @property
    def output_dim(self):
        return self._output_dim

    def spec_transform_func(self, spec):
        if self.spec_transform_type == "exponent":
            if self.spec_abs_exponent != 1:
                e = self.spec_abs_exponent
                spec = spec.abs() ** e * torch.exp(1j * spec.angle())
            spec = spec * self.spec_factor
        elif self.spec_transform_type == "log":
            spec = torch.log(1 + spec.abs()) * torch.exp(1j * spec.angle())
            spec = spec * self.spec_factor
        elif self.spec_transform_type in (None, "none"):
            pass
        return spec

    def forward(self, input, ilens, fs=None):
        with torch.autocast(device_type=input.device.type, enabled=False):
            if fs is not None:
                _reconfig(self.stft, (self.n_fft, self.win_length, self.hop_length), int(fs), self.default_fs)
            if input.dtype in (torch.float16, torch.bfloat16):
                input = input.float()
            spec, flens = self.stft(input, ilens)
            spec = torch.complex(spec[..., 0], spec[..., 1])
            if fs is not None:
                _reconfig(self.stft, (self.n_fft, self.win_length, self.hop_length), self.default_fs, self.default_fs)
            spec = self.spec_transform_func(spec)
        return spec, flens
