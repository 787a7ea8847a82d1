import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import restated as R
from urgent2026_challenge_track1_b200 import BSRNN_SE
from urgent2026_challenge_track1_b200.pipeline import StreamedEnhancer

def rel(a, b): return float((a - b).norm() / b.norm())
torch.manual_seed(0)
for graph in (False, True):
    m = BSRNN_SE(num_channel=32, num_layer=1, precision="fp32", cuda_graph=graph).cuda()
    fs = 16000
    batches = []
    for i, n in enumerate([9000, 9000, 7000, 9000, 9000]):
        x = R.synth_noisy(2, n, fs, seed=10 + i).pin_memory()
        batches.append((x, torch.tensor([n, n - 500 * (i + 1)]), fs))
    direct = [m(x, lens, fs)[0].cpu().clone() for x, lens, fs in batches]
    direct2 = [m(x, lens, fs)[0].cpu().clone() for x, lens, fs in batches]
    print("graph", graph, "direct repeat:", [f"{rel(a, b):.1e}" for a, b in zip(direct2, direct)])
    dev_in = [m(x.cuda(), lens, fs)[0].cpu().clone() for x, lens, fs in batches]
    print("graph", graph, "device input:", [f"{rel(a, b):.1e}" for a, b in zip(dev_in, direct)])
    enh = StreamedEnhancer(m)
    got = [out.clone() for out, _, _ in enh.run(iter(batches))]
    print("graph", graph, "streamed    :", [f"{rel(a, b):.1e}" for a, b in zip(got, direct)])
    for i, g in enumerate(got):
        for j, d in enumerate(direct):
            if g.shape == d.shape and rel(g, d) < 1e-3 and i != j: print("   streamed", i, "equals direct", j)
