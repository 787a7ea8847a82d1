import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import restated as R
from urgent2026_challenge_track1_b200 import BSRNN_SE

def rel(a, b): return float((a - b).norm() / b.norm())
FS, n = 48000, 480000
torch.manual_seed(0)
prec = sys.argv[1] if len(sys.argv) > 1 else "fp16"
Bs = [int(v) for v in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["4", "64"])]
m = BSRNN_SE(num_channel=196, num_layer=6, precision=prec).cuda()
base = R.synth_noisy(8, n, FS, seed=3)
for B in Bs:
    idx = torch.tensor([(i // 2) % 8 for i in range(B)])
    x = base[idx].contiguous()
    lens = torch.full((B,), n, dtype=torch.int32)
    w1 = m(x, lens, FS)[0].cpu()
    w2 = m(x, lens, FS)[0].cpu()
    adj = [rel(w1[2 * j + 1], w1[2 * j]) for j in range(B // 2)]
    print(f"[{prec}] B={B}: run-to-run {rel(w2, w1):.2e}; duplicate rows adjacent max {max(adj):.2e} min {min(adj):.2e}", flush=True)
    if B > 16:
        far = [rel(w1[i + 16], w1[i]) for i in range(B - 16)]
        print(f"          far (16 rows apart) max {max(far):.2e} min {min(far):.2e}", flush=True)
    if B >= 4:
        small = m(base[:2].contiguous(), torch.full((2,), n, dtype=torch.int32), FS)[0].cpu()
        print(f"          vs batch of 2: {rel(w1[0], small[0]):.2e} {rel(w1[2], small[1]):.2e}", flush=True)
