"""Unit checks of the tcgen05 kernels against torch / the oracle.  Run on the GPU box: python tools/gpu_check_tc.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import restated as R
from urgent2026_challenge_track1_b200 import runtime as rt, runtime_tc as tc, _lib as L, BSRNN_SE


def rel(a, b):
    a, b = a.detach().float().cpu().double(), b.detach().float().cpu().double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def check_gemm(M, K, Nout, BN, epi="rows"):
    g = torch.Generator(device="cuda").manual_seed(M + K + Nout)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(Nout, K, device="cuda", generator=g) / K ** 0.5
    bias = torch.randn(Nout, device="cuda", generator=g)
    kc = (K + 15) // 16 * 2
    A8, W8 = tc.to_kb8(A, 128, kc), tc.to_kb8(W, BN, kc)
    n_tiles = W8.shape[0]
    bias_p = torch.zeros(n_tiles * BN, device="cuda"); bias_p[:Nout] = bias
    ref = A.half().float() @ W.half().float().t() + bias
    m_tiles = A8.shape[0]
    st = L.stream_ptr()
    if epi == "rows":
        out = torch.zeros(M, n_tiles * BN, dtype=torch.float16, device="cuda")
        L.call("bsrnn_gemm_tc", A8.data_ptr(), W8.data_ptr(), bias_p.data_ptr(), out.data_ptr(), None, m_tiles, n_tiles, kc, BN,
               L.TC_F16_ROWS, n_tiles * BN, Nout, 0, M, m_tiles, M, tc.BIG, 0, 1, 0, st)
        torch.cuda.synchronize()
        return rel(out[:, :Nout], ref)
    if epi == "resid":
        out = torch.randn(M, Nout, device="cuda", generator=g)
        base = out.clone()
        stats = torch.zeros(4, 2, dtype=torch.float64, device="cuda")
        L.call("bsrnn_gemm_tc", A8.data_ptr(), W8.data_ptr(), bias_p.data_ptr(), out.data_ptr(), stats.data_ptr(), m_tiles, n_tiles,
               kc, BN, L.TC_RESID_F32, Nout, Nout, 0, (M + 3) // 4, m_tiles, M, tc.BIG, 0, 1, 0, st)
        torch.cuda.synchronize()
        want = base + ref
        tps = (M + 3) // 4
        s_ref = torch.stack([torch.stack([want[i * tps:(i + 1) * tps].double().sum(), (want[i * tps:(i + 1) * tps].double() ** 2).sum()]) for i in range(4)])
        return rel(out, want), rel(stats, s_ref)
    if epi == "tanh":
        okc = (Nout + 15) // 16 * 2
        out = torch.zeros(m_tiles, okc, 128, 8, dtype=torch.float16, device="cuda")
        L.call("bsrnn_gemm_tc", A8.data_ptr(), W8.data_ptr(), bias_p.data_ptr(), out.data_ptr(), None, m_tiles, n_tiles, kc, BN,
               L.TC_TANH_KB8, 0, Nout, okc, M, m_tiles, M, tc.BIG, 0, 1, 0, st)
        torch.cuda.synchronize()
        return rel(tc.from_kb8(out, M, Nout), torch.tanh(ref))
    if epi == "glu":
        # interleave (value, gate) rows: ref GLU over halves of the un-interleaved output
        half = Nout // 2
        order = torch.stack([torch.arange(half), torch.arange(half) + half], 1).reshape(-1).cuda()
        W8i = tc.to_kb8(W[order], BN, kc)
        bias_i = torch.zeros(n_tiles * BN, device="cuda"); bias_i[:Nout] = bias[order]
        out = torch.zeros(M, half, device="cuda")
        L.call("bsrnn_gemm_tc", A8.data_ptr(), W8i.data_ptr(), bias_i.data_ptr(), out.data_ptr(), None, m_tiles, n_tiles, kc, BN,
               L.TC_GLU_F32, half, half, 0, M, m_tiles, M, tc.BIG, 0, 1, 0, st)
        torch.cuda.synchronize()
        return rel(out, ref[:, :half] * torch.sigmoid(ref[:, half:]))


def check_lstm(B, T, K, axis, slots=0):
    """recurrence + input projection vs torch: delegated to tools/prof_lstm.py --check (same layouts as the runtime)."""
    import subprocess
    r = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "prof_lstm.py"), "--B", str(B),
                        "--T", str(T), "--K", str(K), "--axis", axis, "--slots", str(slots), "--check", "--reps", "1"],
                       capture_output=True, text=True, timeout=300)
    return [l for l in (r.stdout + r.stderr).splitlines() if "CHECK" in l or "rror" in l]


def main():
    L.require_device()
    print("max co-resident clusters:", L.lib().bsrnn_blstm_tc_max_clusters())
    for (M, K, Nn, BN) in [(128, 64, 208, 208), (300, 208, 3328, 208), (1000, 800, 196, 208), (513, 784, 240, 240), (257, 196, 784, 256), (5120, 196, 3328, 208), (9000, 196, 784, 208)]:
        print(f"gemm rows   M={M} K={K} N={Nn} BN={BN}:", check_gemm(M, K, Nn, BN, "rows"))
    print("gemm resid :", check_gemm(1000, 800, 196, 208, "resid"))
    print("gemm tanh  :", check_gemm(700, 196, 784, 256, "tanh"))
    print("gemm glu   :", check_gemm(700, 784, 240, 240, "glu"))
    sys.stdout.flush()
    for (B, T, K, axis) in [(1, 9, 20, "time"), (2, 33, 34, "time"), (2, 33, 34, "freq"), (12, 40, 34, "time"), (3, 300, 34, "freq")]:
        for slots in (1, 3):
            print(f"lstm B={B} T={T} K={K} {axis} slots={slots}:", check_lstm(B, T, K, axis, slots))
            sys.stdout.flush()


if __name__ == "__main__":
    main()
