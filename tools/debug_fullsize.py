"""Bisect helper: rows 0 / 2 of a 64 x SECONDS s batch against the same utterances in a batch of 2 (tensor-core mode), under
whatever BSRNN_* switches the environment sets.  usage: python tools/debug_fullsize.py [seconds] [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from urgent2026_challenge_track1_b200 import BSRNN_SE, synth

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
FS = 48000
torch.manual_seed(0)
m = BSRNN_SE(num_channel=196, num_layer=2, precision="fp16").cuda()
n = int(FS * secs)
base = synth.synth_noisy(8, n, FS, seed=3)
idx = torch.tensor([(i // 2) % 8 for i in range(B)])
x = base[idx].contiguous()
lens = torch.full((B,), n, dtype=torch.int32)
wav, _ = m(x, lens, FS)
small, _ = m(base[:2].contiguous(), lens[:2], FS)
rel = lambda a, b: float((a - b).norm() / b.norm())
print("switches:", {k: v for k, v in os.environ.items() if k.startswith("BSRNN_")}, "B", B, "secs", secs,
      "rows vs batch of 2: %.3e %.3e" % (rel(wav[0].cpu(), small[0].cpu()), rel(wav[2].cpu(), small[1].cpu())), flush=True)
