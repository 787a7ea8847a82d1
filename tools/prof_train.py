"""Where a tensor-core training step spends its time: torch profiler summary (host + device) of one SETrainer.step at
BASELINE config 5 (B=4 x 96000 samples @48 kHz).   python tools/prof_train.py [--precision fp16]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from urgent2026_challenge_track1_b200 import BSRNN_SE
from urgent2026_challenge_track1_b200.synth import synth_pair
from urgent2026_challenge_track1_b200.training import SETrainer

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="fp16"); ap.add_argument("--batch", type=int, default=4)
a = ap.parse_args()
torch.manual_seed(0)
m = BSRNN_SE(196, 6, precision="fp32").cuda()
tr = SETrainer(m, lr=1e-3, precision=a.precision)
clean, noisy = synth_pair(a.batch, 96000, 48000, seed=1)
clean, noisy = clean.view(a.batch, 1, -1).cuda(), noisy.view(a.batch, 1, -1).cuda()
lens = torch.full((a.batch,), 96000, dtype=torch.int32); fs = torch.tensor(48000, dtype=torch.int32)
for _ in range(2):
    tr.step(noisy, clean, lens, fs)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    tr.step(noisy, clean, lens, fs)
torch.cuda.synchronize()
print(f"wall per step: {(time.perf_counter() - t0) / 3 * 1e3:.1f} ms")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    tr.step(noisy, clean, lens, fs)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=60))
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=22, max_name_column_width=60))
