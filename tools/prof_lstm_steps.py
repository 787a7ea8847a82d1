"""Step-wise tensor-core BLSTM (runtime_tc_steps) alone: parity against torch.nn.LSTM on the CPU and per-step timing.
  python tools/prof_lstm_steps.py --R 96 --steps 40 --N 384 [--check]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from urgent2026_challenge_track1_b200 import runtime_tc_steps as S, _lib as L

ap = argparse.ArgumentParser()
ap.add_argument("--R", type=int, default=96); ap.add_argument("--steps", type=int, default=40)
ap.add_argument("--N", type=int, default=384); ap.add_argument("--check", action="store_true"); ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--bn", type=int, default=0, help="N-tile width of the step GEMM (0 = default choice)")
ap.add_argument("--graph", action="store_true", help="time a CUDA-graph replay of the whole sequence (no host launch cost)")
a = ap.parse_args()
if a.bn:
    S._bn_for = lambda cols, kcores=None: a.bn
torch.manual_seed(0)
N, H, R, steps = a.N, 2 * a.N, a.R, a.steps
rnn = torch.nn.LSTM(N, H, batch_first=True, bidirectional=True)
x = torch.randn(R, steps, N) * 0.7
ref = None
if a.check:
    with torch.no_grad():
        ref = rnn(x)[0]                                             # (R, steps, 2H)
rnn = rnn.cuda()
p = S.pack_lstm_steps_tc(rnn)
tiles = (R + 127) // 128
ws = S.StepsWorkspace(steps, tiles, H, "cuda")
# operand tiles: m = step*tiles + j, row r = sequence j*128 + r  -> token (seq, step) of x laid out (R, steps, N)
xg = x.cuda().contiguous()
xhat = torch.empty(steps * tiles * p["kc_in"] * 1024, dtype=torch.float16, device="cuda")
st = L.stream_ptr()
L.call("bsrnn_norm_cast_kb8", xg.data_ptr(), None, None, xhat.data_ptr(), N, 0, N, p["kc_in"], steps * tiles, tiles, R,
       1 << 40, 0, steps, 1, R * steps, 1, st)                      # token = seq*steps + step
run = lambda: S.blstm_steps_tc(xhat, p, steps, tiles, ws)
run(); torch.cuda.synchronize()
if a.graph:
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        run()
    run = g.replay
    run(); torch.cuda.synchronize()
for _ in range(a.reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"[steps-tc N={N} H={H} BN={p['BN']}{' graph' if a.graph else ''}] R={R} steps={steps}: {ms:.3f} ms, {1e3 * ms / steps:.1f} us per step (both directions), "
          f"{2 * 2.0 * R * steps * H * 4 * H / ms / 1e9:.1f} TFLOP/s recurrent", flush=True)
if a.check:
    outs = []
    for d in (0, 1):
        y = ws.y[d].view(steps, tiles, H // 8, 128, 8).permute(0, 1, 3, 2, 4).reshape(steps, tiles * 128, H)[:, :R]
        outs.append(y.permute(1, 0, 2).float().cpu())
    mine = torch.cat(outs, 2)
    err = float((mine.double() - ref.double()).norm() / ref.double().norm())
    print(f"[steps-tc N={N} H={H}] CHECK R={R} steps={steps}: rel_l2={err:.3e}  {'OK' if err < 3e-3 else 'FAIL'}", flush=True)
