"""Runs the fused BLSTM layer kernel at the FlowSE width (H = 768, N = 384) alone: timing, optional per-role probe
(-DBSRNN_FUSED_PROBE build).

  python tools/prof_lstm768.py --R 1536 --steps 1251 --slots 0 [--trace]        # time axis of BASELINE config 4
  python tools/prof_lstm768.py --R 40032 --steps 48                             # band axis
"""
import argparse, ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from urgent2026_challenge_track1_b200 import runtime_tc_steps as S, _lib as L

ap = argparse.ArgumentParser()
ap.add_argument("--R", type=int, default=1536); ap.add_argument("--steps", type=int, default=1251)
ap.add_argument("--slots", type=int, default=0); ap.add_argument("--maxgroups", type=int, default=0)
ap.add_argument("--reps", type=int, default=3); ap.add_argument("--trace", action="store_true")
a = ap.parse_args()
torch.manual_seed(0)
N, H = 384, 768
rnn = torch.nn.LSTM(N, H, batch_first=True, bidirectional=True).cuda()
p = S.pack_lstm_fused768(rnn)
tiles = (a.R + 127) // 128
ws = S.StepsWorkspace(1, 1, H, "cuda")
y = [torch.empty(a.steps * tiles * (H // 8) * 1024, dtype=torch.float16, device="cuda") for _ in range(2)]
zero = torch.zeros((H // 8) * 1024, dtype=torch.float16, device="cuda")
xhat = torch.empty(a.steps * tiles * p["kc_fused"] * 1024, dtype=torch.float16, device="cuda")
chunk = 1 << 26
for i in range(0, xhat.numel(), chunk):
    n = min(chunk, xhat.numel() - i)
    xhat[i:i + n] = (torch.randn(n, device="cuda") * 0.5).half()
st = L.stream_ptr()


def run():
    L.call("bsrnn_blstm_fused768_tc", xhat.data_ptr(), p["wfused"].data_ptr(), zero.data_ptr(), y[0].data_ptr(), y[1].data_ptr(),
           (H // 8) * 1024, a.R, a.steps, tiles, a.maxgroups, a.slots, ws.sync.data_ptr(), st)


tag = f"FUSED768 slots={a.slots} maxgroups={a.maxgroups}"
print(f"[{tag}] co-resident groups (24 CTA pairs each): {L.lib().bsrnn_blstm_fused768_max_groups()}", flush=True)
flops = 2.0 * a.R * a.steps * 2 * 4 * H * (H + N)
for _ in range(a.reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"[{tag}] R={a.R} steps={a.steps}: {ms:.3f} ms, {1e3 * ms / a.steps:.2f} us per step, TFLOP/s={flops / ms / 1e9:.1f}", flush=True)
if a.trace:
    tr = torch.zeros(32, dtype=torch.int64, device="cuda")
    L.lib().bsrnn_debug_set_fused_probe.argtypes = [ctypes.c_void_p]
    L.lib().bsrnn_debug_set_fused_probe(tr.data_ptr())
    run(); torch.cuda.synchronize()
    L.lib().bsrnn_debug_set_fused_probe(None)
    acc = tr.cpu().tolist()
    print(f"[{tag}] pair 0 whole-launch cycles (needs a -DBSRNN_FUSED_PROBE build):")
    for e in (0, 1):
        o = 16 * e
        print(f"  CTA {e} producer : wait flag {acc[o]:>12d}  wait ring-empty {acc[o+1]:>12d}  other {acc[o+2]:>12d}   total {acc[o]+acc[o+1]+acc[o+2]}")
        print(f"  CTA {e} epilogue (quarter 0, quadrant 0): wait acc_full {acc[o+12]:>11d}  busy {acc[o+13]:>12d}  acc_empty arrive {acc[o+14]:>10d}")
    print(f"  leader mma : wait acc_empty {acc[4]:>10d}  wait full {acc[5]:>12d}  wait pfull {acc[6]:>12d}  other {acc[7]:>12d}")
