"""Oracle shim (TEST INFRASTRUCTURE): restatement of espnet2 ``BSRNNSeparator`` (espnet==202412),
SURVEY.md Appendix A / §8a row a3.  Reference call site: baseline_code/models/bsrnn.py:27-34,38."""
import torch
from espnet2.enh.layers.bsrnn import BSRNN


class BSRNNSeparator(torch.nn.Module):
    def __init__(self, input_dim, num_spk=2, num_channels=16, num_layers=6, target_fs=48000,
                 causal=True, norm_type="GN", ref_channel=None):
        super().__init__()
        self._num_spk = num_spk
        self.ref_channel = ref_channel
        self.bsrnn = BSRNN(input_dim=input_dim, num_channel=num_channels, num_layer=num_layers,
                           target_fs=target_fs, causal=causal, num_spk=num_spk, norm_type=norm_type)

    @property
    def num_spk(self):
        return self._num_spk

    def forward(self, input, ilens, additional=None):
        assert torch.is_complex(input)
        feature = torch.stack([input.real, input.imag], dim=-1)
        assert feature.ndim == 4, "oracle shim restates the single-channel path only"
        processed = self.bsrnn(feature)
        processed = torch.complex(processed[..., 0], processed[..., 1])
        return list(processed.unbind(1)), ilens, {}
