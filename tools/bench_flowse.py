"""FlowSE generative inference (BASELINE config 4 shape at a reduced batch): BSRNN_flowse (N=384) random-init weights,
B x 10 s @ 48 kHz, NFE Euler steps, through FlowSEModel.enhance.  The H=768 recurrence runs the f32 CUDA-core kernels
(no tensor-core kernel for this width yet), so this records the gap, not a target.

  python tools/bench_flowse.py --batch 2 --nfe 15
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import restated as R
from urgent2026_challenge_track1_b200 import _lib
from urgent2026_challenge_track1_b200.config import Config
from urgent2026_challenge_track1_b200.flow_model import FlowSEModel

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=2); ap.add_argument("--seconds", type=float, default=10.0)
ap.add_argument("--nfe", type=int, default=15); ap.add_argument("--hidden", type=int, default=384)
ap.add_argument("--layers", type=int, default=6); ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--precision", default="fp16", choices=["fp32", "fp16"]); ap.add_argument("--graph", action="store_true")
a = ap.parse_args()
_lib.require_device()
torch.manual_seed(0)
cfg = Config(model_type="flowse", ema_decay=0.999, sigma_max=0.5, sigma_min=0.05, t_eps=0.03, T_rev=1.0, loss_type="mse",
             loss_abs_exponent=0.5, n_fft=1536, hop_length=384, spec_transform_type="exponent", spec_abs_exponent=0.667,
             spec_factor=0.065, bsrnn_hidden=a.hidden, num_layer=a.layers, learning_rate=1e-4)
m = FlowSEModel(cfg).cuda().eval(no_ema=True)
m.dnn.precision = a.precision
m.dnn.cuda_graph = a.graph
fs, n = 48000, int(48000 * a.seconds)
y = R.synth_noisy(a.batch, n, fs, seed=5).cuda()
lens = torch.full((a.batch,), n, dtype=torch.int32)
torch.manual_seed(2)
out = m.enhance(y, fs, lens, N=a.nfe); torch.cuda.synchronize()
ts = []
for _ in range(a.reps):
    torch.manual_seed(2)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = m.enhance(y, fs, lens, N=a.nfe); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = min(ts)
if os.environ.get("BSRNN_FLOWSE_REGIONS", "0") == "1":
    from urgent2026_challenge_track1_b200 import runtime
    m.dnn.cuda_graph = False
    with runtime.Profile() as prof:
        out = m.enhance(y, fs, lens, N=1)
        torch.cuda.synchronize()
        print("regions of ONE network evaluation (ms):", {k: round(v[0], 1) for k, v in prof.totals_ms().items()}, file=sys.stderr)
print(json.dumps({"workload": f"BSRNN_flowse N={a.hidden} L={a.layers}, {a.batch}x{a.seconds:g}s@48kHz, NFE={a.nfe}, dual path {a.precision}{', CUDA graph' if a.graph else ''}",
                  "ms": ms, "audio_s_per_s": a.batch * a.seconds / (ms / 1e3), "finite": bool(torch.isfinite(out).all()),
                  "out_shape": list(out.shape)}))
