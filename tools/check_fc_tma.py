"""Linear + skip GEMM: the TMA residual epilogue (bsrnn_gemm_tc epilogue 8) against the register-staged one (epilogue 1) on
the same operands: outputs and GroupNorm statistics must agree to f32 rounding of the same sums."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from urgent2026_challenge_track1_b200 import _lib as L

torch.manual_seed(0)
for (B, T, K, N, kc, BN, nt, axis) in [(3, 50, 34, 196, 100, 208, 1, "time"), (3, 50, 34, 196, 100, 208, 1, "freq"),
                                       (2, 40, 48, 384, 192, 192, 2, "time"), (2, 40, 48, 384, 192, 192, 2, "freq")]:
    if axis == "time":
        R, steps, addr = B * K, T, (K, T * K, 1, K)
    else:
        R, steps, addr = B * T, K, (1, K, 0, 1)
    tiles = (R + 127) // 128
    ntile = steps * tiles
    st = L.stream_ptr()
    A = (torch.randn(ntile * kc * 1024, device="cuda") * 0.3).half()
    W = (torch.randn(nt * kc * BN * 8, device="cuda") * 0.03).half()
    bias = torch.randn(nt * BN, device="cuda")
    base = torch.randn(B, T, K, N, device="cuda")
    outs, stats = [], []
    for epi in (L.TC_RESID_F32, L.TC_RESID_TMA):
        out = base.clone()
        s = torch.zeros(B, 2, dtype=torch.float64, device="cuda")
        L.call("bsrnn_gemm_tc", A.data_ptr(), W.data_ptr(), bias.data_ptr(), out.data_ptr(), s.data_ptr(), ntile, nt, kc, BN, epi, N, N, 0,
               T * K, tiles, R, *addr, st)
        torch.cuda.synchronize()
        outs.append(out); stats.append(s)
    d = float((outs[0] - outs[1]).abs().max()); ds = float(((stats[0] - stats[1]).abs() / stats[0].abs().clamp_min(1e-9)).max())
    moved = float((outs[0] - base).abs().max())
    print(f"B={B} T={T} K={K} N={N} {axis}: max|ldst - tma| = {d:.3e} (update magnitude {moved:.2f}), stats rel diff {ds:.2e}  {'OK' if d < 1e-5 and ds < 1e-6 else 'FAIL'}")
