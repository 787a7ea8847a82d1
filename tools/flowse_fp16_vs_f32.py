"""FlowSE at the full width (N=384, H=768, L=6): tensor-core dual path against the f32 CUDA-core path, same weights,
same prior noise (relative L2 of the vector field and of the enhanced waveform)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import restated as R
from urgent2026_challenge_track1_b200.config import Config
from urgent2026_challenge_track1_b200.flow_model import FlowSEModel
torch.manual_seed(0)
cfg = Config(model_type="flowse", ema_decay=0.999, sigma_max=0.5, sigma_min=0.05, t_eps=0.03, T_rev=1.0, loss_type="mse",
             loss_abs_exponent=0.5, n_fft=1536, hop_length=384, spec_transform_type="exponent", spec_abs_exponent=0.667,
             spec_factor=0.065, bsrnn_hidden=384, num_layer=6, learning_rate=1e-4)
m = FlowSEModel(cfg).cuda().eval(no_ema=True)
fs, n, B = 48000, 96000, 2
y = R.synth_noisy(B, n, fs, seed=5).cuda()
lens = torch.tensor([n, n - 5000], dtype=torch.int32)
rel = lambda a, b: float((a - b).norm() / b.norm())
Y = m.speech_to_feature(y, fs, lens)
torch.manual_seed(3); z = torch.randn_like(Y)
t = torch.tensor([0.7, 0.3], device="cuda")
outs = {}
for prec in ("fp32", "fp16"):
    m.dnn.precision = prec
    vf = m(Y + 0.5 * z, t, Y)
    enh = m.enhance(y, fs, lens, N=5, z=z)
    outs[prec] = (vf, enh)
print(f"FlowSE N=384 L=6, {B}x2s@48kHz: fp16-vs-f32 rel_l2 vector field {rel(outs['fp16'][0], outs['fp32'][0]):.3e}, "
      f"enhanced (NFE=5) {rel(outs['fp16'][1], outs['fp32'][1]):.3e}")
