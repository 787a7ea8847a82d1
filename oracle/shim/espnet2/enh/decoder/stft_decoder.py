"""Oracle shim (TEST INFRASTRUCTURE): restatement of espnet2 ``STFTDecoder`` (espnet==202412),
behaviour per SURVEY.md Appendix A.  Reference call sites: baseline_code/models/bsrnn.py:21-25,40;
baseline_code/flow_model.py:35-42,145."""
import torch
from espnet2.layers.stft import Stft
from espnet2.enh.encoder.stft_encoder import _reconfig


class STFTDecoder(torch.nn.Module):
    def __init__(self, n_fft=512, win_length=None, hop_length=128, window="hann", center=True,
                 normalized=False, onesided=True, default_fs=16000, spec_transform_type=None,
                 spec_factor=0.15, spec_abs_exponent=0.5):
        super().__init__()
        self.stft = Stft(n_fft=n_fft, win_length=win_length, hop_length=hop_length, window=window,
                         center=center, normalized=normalized, onesided=onesided)
        self.win_length = win_length if win_length else n_fft
        self.n_fft = n_fft
        self.hop_length = hop_length
        self.window = window
        self.center = center
        self.default_fs = default_fs
        self.spec_transform_type = spec_transform_type
        self.spec_factor = spec_factor
        self.spec_abs_exponent = spec_abs_exponent

    def spec_back(self, spec):
        if self.spec_transform_type == "exponent":
            spec = spec / self.spec_factor
            if self.spec_abs_exponent != 1:
                e = self.spec_abs_exponent
                spec = spec.abs() ** (1 / e) * torch.exp(1j * spec.angle())
        elif self.spec_transform_type == "log":
            spec = spec / self.spec_factor
            spec = (torch.exp(spec.abs()) - 1) * torch.exp(1j * spec.angle())
        return spec

    def forward(self, input, ilens, fs=None):
        with torch.autocast(device_type=input.device.type, enabled=False):
            if fs is not None:
                _reconfig(self.stft, (self.n_fft, self.win_length, self.hop_length), int(fs), self.default_fs)
            input = self.spec_back(input)
            wav, wav_lens = self.stft.inverse(input, ilens)
            if fs is not None:
                _reconfig(self.stft, (self.n_fft, self.win_length, self.hop_length), self.default_fs, self.default_fs)
        return wav, wav_lens
