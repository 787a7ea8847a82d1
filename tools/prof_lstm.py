"""Runs the tensor-core BLSTM recurrence alone (timing / ncu / parity vs torch CPU).

  python tools/prof_lstm.py --B 64 --T 1001 --K 34 --axis time --slots 3 --variant 0 [--check] [--trace] [--v2]
"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from urgent2026_challenge_track1_b200 import runtime_tc as tc, _lib as L

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=16); ap.add_argument("--T", type=int, default=401); ap.add_argument("--K", type=int, default=34)
ap.add_argument("--axis", default="time"); ap.add_argument("--reps", type=int, default=3); ap.add_argument("--maxcl", type=int, default=0)
ap.add_argument("--slots", type=int, default=0); ap.add_argument("--variant", type=int, default=0)
ap.add_argument("--trace-cid", type=int, default=0); ap.add_argument("--flags", type=int, default=0)
ap.add_argument("--ver", type=int, default=0, help="recurrence schedule 4..9 (0 = BSRNN_LSTM_VER / default)")
ap.add_argument("--flag", action="store_true", help="flag-group schedule (bsrnn_blstm_recurrence_tc_flag)")
ap.add_argument("--fused", action="store_true", help="fused layer (bsrnn_blstm_fused_tc): x * W_ih inside the recurrence")
ap.add_argument("--geo", type=int, default=8, help="fused kernel geometry: 8 pairs x 49 units or 7 pairs x 56 units")
ap.add_argument("--v2", action="store_true"); ap.add_argument("--check", action="store_true"); ap.add_argument("--trace", action="store_true")
a = ap.parse_args()
B, T, K, axis = a.B, a.T, a.K, a.axis
torch.manual_seed(0)
N, H = 196, 392
rnn = torch.nn.LSTM(N, H, batch_first=True, bidirectional=True)
if axis == "time":
    R, steps, addr = B * K, T, (K, T * K, 1, K)
else:
    R, steps, addr = B * T, K, (1, K, 0, 1)
M = B * T * K
tiles = (R + 127) // 128
ref = None
if a.check:
    x = torch.randn(B, T, K, N) * 0.7
    with torch.no_grad():
        if axis == "time":
            ref = rnn(x.permute(0, 2, 1, 3).reshape(B * K, T, N))[0].reshape(B, K, T, 2 * H).permute(0, 2, 1, 3)
        else:
            ref = rnn(x.reshape(B * T, K, N))[0].reshape(B, T, K, 2 * H)
rnn = rnn.cuda()
p = tc.pack_lstm_tc(rnn)
st = L.stream_ptr()
ntile = steps * tiles
gates = torch.empty((1 if a.fused else ntile) * 416 * 1024, dtype=torch.float16, device="cuda")
zero_tile = torch.zeros(50 * 1024, dtype=torch.float16, device="cuda")
xhat = torch.empty(ntile * p["kc_in"] * 1024, dtype=torch.float16, device="cuda")
if a.check or a.fused:
    xg = (x.cuda() if a.check else torch.randn(B, T, K, N, device="cuda") * 0.7).reshape(M, N).contiguous()
    L.call("bsrnn_norm_cast_kb8_ones", xg.data_ptr(), None, None, xhat.data_ptr(), N, 0, N, p["kc_in"], ntile, tiles, R,
           *addr, M, 1, p["one_col"], st)
    if not a.fused:
        L.call("bsrnn_gemm_tc", xhat.data_ptr(), p["wih"].data_ptr(), None, gates.data_ptr(), None, ntile, 16,
               p["kc_in"], 208, L.TC_F16_KB8, 0, 3328, 416, M, tiles, R, *addr, st)
else:
    chunk = 1 << 26
    for i in range(0, gates.numel(), chunk):
        n = min(chunk, gates.numel() - i)
        gates[i:i + n] = (torch.randn(n, device="cuda") * 0.5).half()
y = torch.zeros(steps * tiles * 2 * 50 * 1024, dtype=torch.float16, device="cuda")


sync = torch.zeros(max(L.lib().bsrnn_blstm_tc_sync_bytes(), L.lib().bsrnn_blstm_fused_sync_bytes()) // 4, dtype=torch.int32,
                   device="cuda")


def run():
    if a.fused:
        L.call(f"bsrnn_blstm_fused{a.geo}_tc" if a.geo in (7, 14) else "bsrnn_blstm_fused_tc", xhat.data_ptr(),
               tc.fused_weights(p, a.geo).data_ptr(), zero_tile.data_ptr(), y.data_ptr(), R, steps,
               tiles, a.maxcl, a.slots, sync.data_ptr(), st)
        return
    if a.flag:
        L.call("bsrnn_blstm_recurrence_tc_flag", gates.data_ptr(), p["whh"].data_ptr(), zero_tile.data_ptr(), y.data_ptr(), R,
               steps, tiles, a.maxcl, a.slots, sync.data_ptr(), st)
        return
    L.call("bsrnn_blstm_recurrence_tc_ex", gates.data_ptr(), p["whh"].data_ptr(), zero_tile.data_ptr(), y.data_ptr(), R, steps,
           tiles, a.maxcl, a.slots, st)


if a.ver:
    L.lib().bsrnn_debug_set_lstm_schedule(a.ver)
tag = f"{(f'FUSED{a.geo}' if a.geo in (7, 14) else 'FUSED') if a.fused else 'FLAG' if a.flag else 'v' + str(a.ver or os.environ.get('BSRNN_LSTM_VER', '8'))} slots={a.slots} maxcl={a.maxcl}"
if (a.ver or int(os.environ.get('BSRNN_LSTM_VER', '8'))) == 7:
    print(f"[{tag}] co-resident 16-CTA clusters: {L.lib().bsrnn_blstm_tc_max_pair_clusters()}  (8-CTA: {L.lib().bsrnn_blstm_tc_max_clusters()})", flush=True)
if a.fused:
    print(f"[{tag}] co-resident fused groups (8 CTA pairs each): {getattr(L.lib(), f'bsrnn_blstm_fused{a.geo}_max_groups')() if a.geo in (7, 14) else L.lib().bsrnn_blstm_fused_max_groups()}", flush=True)
for _ in range(a.reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"[{tag}] B={B} T={T} K={K} {axis}: {ms:.3f} ms, units={2*tiles}, us per unit-step={1e3*ms*15/(2*tiles*steps):.2f} (x15 clusters), "
          f"TFLOP/s={2*R*steps*2*H*4*H/ms/1e9:.1f}", flush=True)

if a.check:
    yv = y.view(steps, tiles, 2, 50, 128, 8).permute(0, 1, 4, 2, 3, 5).reshape(steps, tiles * 128, 2, 400)[:, :R, :, :H]
    yv = yv.reshape(steps, R, 2 * H).float().cpu()
    mine = yv.reshape(T, B, K, 2 * H).permute(1, 0, 2, 3) if axis == "time" else yv.reshape(K, B, T, 2 * H).permute(1, 2, 0, 3)
    err = float((mine.double() - ref.double()).norm() / ref.double().norm())
    pad = y.view(steps, tiles, 2, 50, 128, 8)[:, :, :, 49].abs().max().item()
    print(f"[{tag}] CHECK {axis} B={B} T={T} K={K}: rel_l2={err:.3e} pad_core_max={pad}  {'OK' if err < 3e-3 and pad == 0 else 'FAIL'}", flush=True)

if a.trace and a.fused:
    import ctypes
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
        print(f"  CTA {e} epilogue (third 0, quadrant 0): wait acc_full {acc[o+12]:>11d}  busy {acc[o+13]:>12d}  acc_empty arrive {acc[o+14]:>10d}")
    print(f"  leader mma : wait acc_empty {acc[4]:>10d}  wait full {acc[5]:>12d}  wait pfull {acc[6]:>12d}  other {acc[7]:>12d}")
elif a.trace:
    import ctypes
    tr = torch.zeros(32, dtype=torch.int64, device="cuda")
    L.lib().bsrnn_debug_set_lstm_probe.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.lib().bsrnn_debug_set_lstm_probe(tr.data_ptr(), a.trace_cid)
    run(); torch.cuda.synchronize()
    L.lib().bsrnn_debug_set_lstm_probe(None, 0)
    acc = tr.cpu().tolist()
    print(f"[{tag}] cluster {a.trace_cid} CTA 0 whole-launch cycles:")
    print(f"  producer : wait h_ready {acc[0]:>12d}  wait ring-empty {acc[1]:>12d}  other {acc[2]:>12d}   total {acc[0]+acc[1]+acc[2]}")
    print(f"  mma      : wait acc_empty {acc[4]:>10d}  wait ring-full {acc[5]:>13d}  other {acc[6]:>12d}")
    print(f"  epilogue (slot 0, quadrant 0): wait acc_full {acc[12]:>11d}  busy {acc[13]:>12d}")
