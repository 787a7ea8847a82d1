"""One small tensor-core-mode forward (2 layers, 2 x 0.5 s @ 16 kHz, ragged) for compute-sanitizer runs:
  compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import restated as R
from urgent2026_challenge_track1_b200 import BSRNN_SE
torch.manual_seed(0)
m = BSRNN_SE(num_channel=196, num_layer=2, precision="fp16")
sd = {k: v.clone() for k, v in m.state_dict().items()}
m.cuda()
fs, n = 16000, 8000
x = R.synth_noisy(2, n, fs)
lens = torch.tensor([n, n - 1234])
wav, _ = m(x, lens, fs)
torch.cuda.synchronize()
ref, _ = R.bsrnn_se_forward(sd, x, lens, fs, num_layer=2)
print("rel_l2", float((wav.cpu() - ref).norm() / ref.norm()))
