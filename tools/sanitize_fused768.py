"""One small bsrnn_blstm_fused768_tc launch (H = 768 fused layer kernel, 3 tiles, 5 steps) for compute-sanitizer runs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from urgent2026_challenge_track1_b200 import runtime_tc_steps as S, _lib as L
torch.manual_seed(0)
N, H, Rr, steps = 384, 768, 300, 5
rnn = torch.nn.LSTM(N, H, batch_first=True, bidirectional=True)
x = torch.randn(Rr, steps, N) * 0.7
with torch.no_grad():
    ref = rnn(x)[0]
p = S.pack_lstm_fused768(rnn.cuda())
tiles = (Rr + 127) // 128
ws = S.StepsWorkspace(steps, tiles, H, "cuda")
xhat = torch.empty(steps * tiles * p["kc_fused"] * 1024, dtype=torch.float16, device="cuda")
xg = x.cuda().contiguous()
st = L.stream_ptr()
L.call("bsrnn_norm_cast_kb8_ones", xg.data_ptr(), None, None, xhat.data_ptr(), N, 0, N, p["kc_fused"], steps * tiles, tiles, Rr,
       1 << 40, 0, steps, 1, Rr * steps, 1, p["one_col"], st)
for d in (0, 1):
    ws.y[d].zero_()
L.call("bsrnn_blstm_fused768_tc", xhat.data_ptr(), p["wfused"].data_ptr(), ws.zero.data_ptr(), ws.y[0].data_ptr(), ws.y[1].data_ptr(),
       (H // 8) * 1024, Rr, steps, tiles, 0, 0, ws.sync.data_ptr(), st)
torch.cuda.synchronize()
outs = []
for d in (0, 1):
    yd = ws.y[d].view(steps, tiles, H // 8, 128, 8).permute(0, 1, 3, 2, 4).reshape(steps, tiles * 128, H)[:, :Rr]
    outs.append(yd.permute(1, 0, 2).float().cpu())
o = torch.cat(outs, 2)
print("rel_l2", float((o - ref).norm() / ref.norm()))
