"""Runs the tensor-core BLSTM recurrence alone (for ncu / timing).  python tools/prof_lstm.py [B T K axis reps maxcl]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from urgent2026_challenge_track1_b200 import runtime_tc as tc, _lib as L

B, T, K = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (16, 401, 34)
axis = sys.argv[4] if len(sys.argv) > 4 else "time"
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
maxcl = int(sys.argv[6]) if len(sys.argv) > 6 else 0
torch.manual_seed(0)
N, H = 196, 392
rnn = torch.nn.LSTM(N, H, batch_first=True, bidirectional=True).cuda()
p = tc.pack_lstm_tc(rnn)
M = B * T * K
if axis == "time":
    R, steps, addr = B * K, T, (K, T * K, 1, K)
else:
    R, steps, addr = B * T, K, (1, K, 0, 1)
tiles = (R + 127) // 128
gates = (torch.randn(M, 3328, device="cuda") * 0.5).half()
y = torch.zeros(steps * tiles * 2 * 50 * 1024, dtype=torch.float16, device="cuda")
st = L.stream_ptr()
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    L.call("bsrnn_blstm_recurrence_tc", gates.data_ptr(), p["whh"].data_ptr(), y.data_ptr(), R, steps, tiles, *addr, maxcl, st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    ncl = min(2 * tiles, 15 if maxcl <= 0 else maxcl)
    waves = -(-2 * tiles // ncl)
    print(f"B={B} T={T} K={K} {axis}: {ms:.3f} ms, units={2*tiles}, clusters={ncl}, waves={waves}, "
          f"us/step={1e3*ms/(waves*steps):.2f}, TFLOP/s={2*R*steps*2*H*4*H/ms/1e9:.1f}")

if os.environ.get("LSTM_TRACE"):
    import ctypes
    tr = torch.zeros(64 * 8, dtype=torch.int64, device="cuda")
    L.lib().bsrnn_debug_set_lstm_trace.argtypes = [ctypes.c_void_p]
    L.lib().bsrnn_debug_set_lstm_trace(tr.data_ptr())
    L.call("bsrnn_blstm_recurrence_tc", gates.data_ptr(), p["whh"].data_ptr(), y.data_ptr(), R, steps, tiles, *addr, maxcl, st)
    torch.cuda.synchronize()
    L.lib().bsrnn_debug_set_lstm_trace(None)
    t = tr.view(64, 8).cpu()
    names = ["hready_seen", "loads_issued", "first_full", "last_full", "mma_committed", "acc_full_seen", "epi_done", "arrived"]
    print("step " + " ".join(f"{n:>14s}" for n in names) + "   (cycles relative to hready_seen of the step)")
    for s_ in range(2, 12):
        base = int(t[s_, 0])
        print(f"{s_:4d} " + " ".join(f"{int(t[s_, i]) - base:14d}" for i in range(8)) + f"   step period {int(t[s_+1,0]) - base}")
