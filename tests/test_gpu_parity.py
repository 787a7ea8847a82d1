"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI (ctypes -> libbsrnn_b200.so), against the CPU
oracle (oracle/restated.py) and the committed golden vectors of the verbatim reference.  Tolerances: relative L2
<= 1e-3 for the f32 mode, <= 1e-2 for the bf16 mode on enhanced waveforms (BASELINE.json north_star); individual f32
kernels are held to much tighter bounds."""
import pytest
import torch

from conftest import golden, golden_sd, rel_l2
from oracle import restated as R

pytestmark = pytest.mark.gpu

RATES = (8000, 16000, 22050, 24000, 32000, 44100, 48000)


@pytest.fixture(scope="module")
def rt():
    from urgent2026_challenge_track1_b200 import runtime, _lib
    _lib.require_device()
    return runtime


@pytest.mark.parametrize("fs", RATES)
def test_stft_istft_all_rates(rt, fs):
    n_fft, hop = R.stft_dims(fs, 960, 480)
    n = fs // 2 + 37
    x = R.synth_noisy(3, n, fs, seed=fs)
    lens = torch.tensor([n, n - 1000, n // 2])
    spec_ref, _ = R.stft_encode(x, lens, fs)
    spec = rt.stft(x.cuda(), lens.int().cuda(), n_fft, hop)
    got = torch.view_as_complex(spec).cpu()
    assert got.shape == spec_ref.shape
    assert rel_l2(got, spec_ref) < 5e-6
    assert float(got[2, int(R.frame_lengths(lens, n_fft, hop)[2]):].abs().max()) == 0.0      # masked frames are exact zeros
    wav_ref = R.stft_decode(spec_ref, lens, fs)
    wav, _ = rt.istft(torch.view_as_real(spec_ref).contiguous().cuda(), None, None, int(lens.max()), n_fft, hop)
    assert rel_l2(wav.cpu(), wav_ref) < 5e-6


@pytest.mark.parametrize("n_fft,hop,fs", [(1536, 384, 48000), (705, 176, 22050), (1411, 352, 44100), (512, 128, 16000)])
def test_stft_exponent_transform_flowse_sizes(rt, n_fft, hop, fs):
    x = R.synth_noisy(2, fs // 3, fs, seed=4)
    lens = torch.tensor([fs // 3, fs // 4])
    ref, _ = R.stft_encode(x, lens, fs, 1536, 384, 48000, "exponent", 0.667, 0.065)
    spec = rt.stft(x.cuda(), lens.int().cuda(), n_fft, hop, 1, 0.667, 0.065)
    assert rel_l2(torch.view_as_complex(spec).cpu(), ref) < 2e-5
    wav_ref = R.stft_decode(ref, lens, fs, 1536, 384, 48000, "exponent", 0.667, 0.065)
    wav, _ = rt.istft(torch.view_as_real(ref).contiguous().cuda(), None, None, int(lens.max()), n_fft, hop,
                      transform=1, exponent=0.667, factor=0.065)
    assert rel_l2(wav.cpu(), wav_ref) < 2e-5


def test_istft_fused_mask(rt):
    fs, n = 48000, 9600
    g = torch.Generator().manual_seed(0)
    x = R.synth_noisy(2, n, fs)
    lens = torch.tensor([n, n])
    spec, _ = R.stft_encode(x, lens, fs)
    m = torch.randn(spec.shape, generator=g, dtype=torch.complex64)
    r = torch.randn(spec.shape, generator=g, dtype=torch.complex64) * 0.1
    ref_spec = m * spec + r
    ref_wav = R.stft_decode(ref_spec, lens, fs)
    wav, est = rt.istft(*(torch.view_as_real(t).contiguous().cuda() for t in (spec, m, r)), n, 960, 480)
    assert rel_l2(torch.view_as_complex(est).cpu(), ref_spec) < 1e-6
    assert rel_l2(wav.cpu(), ref_wav) < 5e-6


@pytest.mark.parametrize("fs", RATES)
def test_bsrnn_se_f32_vs_golden(fs):
    """Small-width model with the verbatim reference's weights and outputs (tests/golden)."""
    from urgent2026_challenge_track1_b200 import BSRNN_SE
    g = golden("bsrnn_se_n16_l2.npz")
    m = BSRNN_SE(num_channel=16, num_layer=2, precision="fp32")
    m.load_state_dict(golden_sd(g))
    m.cuda()
    wav, lens = torch.from_numpy(g[f"in/{fs}/wav"]), torch.from_numpy(g[f"in/{fs}/lens"])
    out, spec = m(wav, lens, fs)
    assert out.shape == g[f"out/{fs}/wav"].shape and spec.dtype == torch.complex64
    assert rel_l2(spec.cpu(), g[f"out/{fs}/spec"]) < 1e-4
    assert rel_l2(out.cpu(), g[f"out/{fs}/wav"]) < 1e-4          # bar is 1e-3 (north_star, f32 mode)


@pytest.mark.parametrize("fs,secs,B", [(16000, 1.0, 2), (48000, 0.5, 2)])
def test_bsrnn_se_f32_fullwidth_vs_oracle(fs, secs, B):
    """BSRNN_baseline.yaml width (N=196, 6 layers), random init, ragged lengths, against the oracle on CPU."""
    from urgent2026_challenge_track1_b200 import BSRNN_SE
    torch.manual_seed(0)
    m = BSRNN_SE(num_channel=196, num_layer=6, precision="fp32")
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    m.cuda()
    n = int(fs * secs)
    x = R.synth_noisy(B, n, fs, seed=1)
    lens = torch.tensor([n] + [n - 997 * (i + 1) for i in range(B - 1)])
    with torch.no_grad():
        ref_wav, ref_spec = R.bsrnn_se_forward(sd, x, lens, fs, num_layer=6)
    out, spec = m(x, lens, fs)
    e_w, e_s = rel_l2(out.cpu(), ref_wav), rel_l2(spec.cpu(), ref_spec)
    print(f"fs={fs} rel_l2 wav={e_w:.3e} spec={e_s:.3e}")
    assert e_w < 1e-3 and e_s < 1e-3


@pytest.mark.parametrize("fs,secs,B", [(16000, 1.0, 2), (48000, 1.0, 3), (22050, 0.7, 2)])
def test_bsrnn_se_tensorcore_vs_oracle(fs, secs, B):
    """Tensor-core mode (fp16 operands, f32 accumulation): bar is rel L2 <= 1e-2 on the enhanced waveform."""
    from urgent2026_challenge_track1_b200 import BSRNN_SE
    torch.manual_seed(0)
    m = BSRNN_SE(num_channel=196, num_layer=6, precision="fp16")
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    m.cuda()
    n = int(fs * secs)
    x = R.synth_noisy(B, n, fs, seed=1)
    lens = torch.tensor([n] + [n - 997 * (i + 1) for i in range(B - 1)])
    with torch.no_grad():
        ref_wav, ref_spec = R.bsrnn_se_forward(sd, x, lens, fs, num_layer=6)
    out, spec = m(x, lens, fs)
    e_w, e_s = rel_l2(out.cpu(), ref_wav), rel_l2(spec.cpu(), ref_spec)
    print(f"fs={fs} tensor-core rel_l2 wav={e_w:.3e} spec={e_s:.3e}")
    assert e_w < 1e-2 and e_s < 1e-2
    out2, _ = m(x, lens, fs)                       # workspaces are reused across calls: result must not drift
    assert rel_l2(out2.cpu(), out.cpu()) < 1e-6


@pytest.mark.parametrize("precision,width,layers", [("fp16", 196, 2), ("fp32", 32, 2)])
def test_bsrnn_se_cuda_graph_replay_matches_eager(precision, width, layers):
    """cuda_graph=True replays the captured launch sequence: same result as host launches, new inputs are honoured,
    and an in-place parameter update invalidates the graph."""
    from urgent2026_challenge_track1_b200 import BSRNN_SE
    torch.manual_seed(0)
    m = BSRNN_SE(num_channel=width, num_layer=layers, precision=precision).cuda()
    fs, n = 16000, 12000
    lens = torch.tensor([n, n - 777])
    x1, x2 = R.synth_noisy(2, n, fs, seed=1), R.synth_noisy(2, n, fs, seed=2)
    e1, e2 = m(x1, lens, fs)[0].clone(), m(x2, lens, fs)[0].clone()
    m.cuda_graph = True
    g1 = m(x1, lens, fs)[0].clone()
    g2 = m(x2.cuda(), lens, fs)[0].clone()          # second call = pure replay with a device-resident input
    g1b = m(x1, lens, fs)[0].clone()
    assert len(m._graphs) == 1
    assert rel_l2(g1.cpu(), e1.cpu()) < 1e-5 and rel_l2(g2.cpu(), e2.cpu()) < 1e-5 and rel_l2(g1b.cpu(), g1.cpu()) < 1e-6
    with torch.no_grad():
        m.bsrnn.bsrnn.fc_time[0].weight.mul_(0.5)  # in-place update (optimizer / EMA swap): graph must be rebuilt
    g3 = m(x1, lens, fs)[0].clone()
    m.cuda_graph = False
    e3 = m(x1, lens, fs)[0]
    assert rel_l2(g3.cpu(), e3.cpu()) < 1e-5 and rel_l2(g3.cpu(), g1.cpu()) > 1e-4


@pytest.mark.parametrize("graph", [False, True])
def test_streamed_enhancer_matches_direct_calls(graph):
    """pipeline.StreamedEnhancer (H2D / D2H on their own streams around the forward) returns, batch by batch and in
    order, what direct calls return -- including ragged batches, changing shapes and reused staging buffers."""
    from urgent2026_challenge_track1_b200 import BSRNN_SE
    from urgent2026_challenge_track1_b200.pipeline import StreamedEnhancer
    torch.manual_seed(0)
    m = BSRNN_SE(num_channel=32, num_layer=1, precision="fp32", cuda_graph=graph).cuda()
    fs = 16000
    batches = []
    for i, n in enumerate([9000, 9000, 7000, 9000, 9000]):
        x = R.synth_noisy(2, n, fs, seed=10 + i).pin_memory()
        batches.append((x, torch.tensor([n, n - 500 * (i + 1)]), fs))
    direct = [m(x, lens, fs)[0].cpu().clone() for x, lens, fs in batches]
    enh = StreamedEnhancer(m)
    got = [out.clone() for out, _, _ in enh.run(iter(batches))]
    assert len(got) == len(direct)
    for g, d in zip(got, direct):
        assert g.shape == d.shape and rel_l2(g, d) < 1e-6
    assert enh.h2d_bytes == sum(x.numel() * 4 for x, _, _ in batches)


# ------------------------------------------------------------------------------------------------ FlowSE
def _flow_model(g):
    from urgent2026_challenge_track1_b200.config import Config
    from urgent2026_challenge_track1_b200.flow_model import FlowSEModel
    cfg = Config(model_type="flowse", ema_decay=0.999, sigma_max=0.5, sigma_min=0.05, t_eps=0.03, T_rev=1.0,
                 loss_type="mse", loss_abs_exponent=0.5, n_fft=1536, hop_length=384, spec_transform_type="exponent",
                 spec_abs_exponent=0.667, spec_factor=0.065, bsrnn_hidden=16, num_layer=1, learning_rate=1e-4)
    m = FlowSEModel(cfg)
    m.load_state_dict(golden_sd(g))
    m.dnn.precision = "fp32"                    # the f32-bar tests below; the tensor-core tests switch it themselves
    return m.cuda().eval(no_ema=True)


@pytest.mark.parametrize("fs", (16000, 22050, 48000))
def test_flowse_vs_golden(fs):
    """Vector field, fused Euler sampler and the literal registry loop against the verbatim reference's outputs."""
    g = golden("flowse_n16_l1.npz")
    m = _flow_model(g)
    y, lens = torch.from_numpy(g[f"in/{fs}/wav"]), torch.from_numpy(g[f"in/{fs}/lens"])
    z, t = torch.from_numpy(g[f"in/{fs}/z"]), torch.from_numpy(g[f"in/{fs}/t"])
    Y = m.speech_to_feature(y, fs, lens)
    assert rel_l2(Y.cpu(), g[f"out/{fs}/feature"]) < 2e-5
    vf = m(Y + 0.5 * z.cuda(), t.cuda(), Y)
    assert vf.shape == Y.shape and rel_l2(vf.cpu(), g[f"out/{fs}/vf"]) < 1e-4
    enh = m.enhance(y, fs, lens, N=3, z=z)
    assert rel_l2(enh.cpu(), g[f"out/{fs}/enhanced"]) < 1e-3           # f32 bar (north_star)
    # the literal registry loop draws its own prior noise on the GPU: replay that draw for the fused path
    torch.manual_seed(11)
    z_gpu = torch.randn_like(Y)
    torch.manual_seed(11)
    enh2 = m.enhance(y, fs, lens, N=3, solver="euler")
    enh3 = m.enhance(y, fs, lens, N=3, z=z_gpu)
    assert rel_l2(enh2.cpu(), enh3.cpu()) < 1e-4


@pytest.mark.parametrize("fs,graph", [(16000, False), (48000, False), (48000, True)])
def test_flowse_tensorcore_steps_vs_golden(fs, graph):
    """FlowSE with the dual path on fp16 tensor-core GEMMs + the step-wise tensor-core BLSTM (runtime_tc_steps, any
    H % 16 == 0): vector field and sampled waveform against the verbatim reference's outputs; bar 1e-2 (16-bit mode)."""
    g = golden("flowse_n16_l1.npz")
    m = _flow_model(g)
    m.dnn.precision = "fp16"
    m.dnn.cuda_graph = graph                    # graph: one captured network evaluation replayed per Euler step
    y, lens = torch.from_numpy(g[f"in/{fs}/wav"]), torch.from_numpy(g[f"in/{fs}/lens"])
    z, t = torch.from_numpy(g[f"in/{fs}/z"]), torch.from_numpy(g[f"in/{fs}/t"])
    Y = m.speech_to_feature(y, fs, lens)
    vf = m(Y + 0.5 * z.cuda(), t.cuda(), Y)
    e_vf = rel_l2(vf.cpu(), g[f"out/{fs}/vf"])
    enh = m.enhance(y, fs, lens, N=3, z=z)
    e_enh = rel_l2(enh.cpu(), g[f"out/{fs}/enhanced"])
    enh_again = m.enhance(y, fs, lens, N=3, z=z)   # second call: pure replays (graph) / reused workspaces
    print(f"fs={fs} graph={graph} FlowSE tensor-core steps: vf rel_l2={e_vf:.3e} enhanced rel_l2={e_enh:.3e}")
    assert e_vf < 1e-2 and e_enh < 1e-2 and rel_l2(enh_again.cpu(), enh.cpu()) < 1e-6
    if graph:
        assert len(m.dnn._graphs) == 1


@pytest.mark.parametrize("geo", [8, 7, 14])
@pytest.mark.parametrize("axis,B,T,slots", [("time", 12, 40, 1), ("time", 10, 24, 2), ("freq", 3, 300, 3), ("freq", 5, 77, 0),
                                             ("time", 40, 30, 0), ("time", 2, 50, 0)])
def test_blstm_fused_vs_torch(axis, B, T, slots, geo):
    """bsrnn_blstm_fused_tc (input projection inside the persistent recurrence, CTA pairs + flag groups) against
    torch.nn.LSTM on the CPU; even / odd tile counts (the odd CTA of the last pair idles), 1..3 interleaved tile pairs,
    single-tile inputs."""
    from urgent2026_challenge_track1_b200 import runtime_tc as tc, _lib as L
    torch.manual_seed(0)
    N, H, K = 196, 392, 34
    rnn = torch.nn.LSTM(N, H, batch_first=True, bidirectional=True)
    x = torch.randn(B, T, K, N) * 0.7
    with torch.no_grad():
        if axis == "time":
            ref = rnn(x.permute(0, 2, 1, 3).reshape(B * K, T, N))[0].reshape(B, K, T, 2 * H).permute(0, 2, 1, 3)
            R_, steps, addr = B * K, T, (K, T * K, 1, K)
        else:
            ref = rnn(x.reshape(B * T, K, N))[0].reshape(B, T, K, 2 * H)
            R_, steps, addr = B * T, K, (1, K, 0, 1)
    p = tc.pack_lstm_tc(rnn.cuda())
    st = L.stream_ptr()
    M, tiles = B * T * K, (R_ + 127) // 128
    ntile = steps * tiles
    xhat = torch.empty(ntile * p["kc_in"] * 1024, dtype=torch.float16, device="cuda")
    xg = x.cuda().reshape(M, N).contiguous()
    L.call("bsrnn_norm_cast_kb8_ones", xg.data_ptr(), None, None, xhat.data_ptr(), N, 0, N, p["kc_in"], ntile, tiles, R_,
           *addr, M, 1, p["one_col"], st)
    y = torch.zeros(ntile * 2 * 50 * 1024, dtype=torch.float16, device="cuda")
    zero_tile = torch.zeros(50 * 1024, dtype=torch.float16, device="cuda")
    sync = torch.zeros(L.lib().bsrnn_blstm_fused_sync_bytes() // 4, dtype=torch.int32, device="cuda")
    for _ in range(2):                                  # the second launch reuses y and the counters
        L.call(f"bsrnn_blstm_fused{geo}_tc" if geo in (7, 14) else "bsrnn_blstm_fused_tc", xhat.data_ptr(),
               tc.fused_weights(p, geo).data_ptr(), zero_tile.data_ptr(), y.data_ptr(), R_, steps,
               tiles, 0, slots, sync.data_ptr(), st)
    yv = y.view(steps, tiles, 2, 50, 128, 8).permute(0, 1, 4, 2, 3, 5).reshape(steps, tiles * 128, 2, 400)[:, :R_, :, :H]
    yv = yv.reshape(steps, R_, 2 * H).float().cpu()
    mine = yv.reshape(T, B, K, 2 * H).permute(1, 0, 2, 3) if axis == "time" else yv.reshape(K, B, T, 2 * H).permute(1, 2, 0, 3)
    assert y.view(steps, tiles, 2, 50, 128, 8)[:, :, :, 49].abs().max().item() == 0      # K padding of y stays zero
    e = rel_l2(mine, ref)
    print(f"fused blstm {axis} B={B} T={T} slots={slots}: rel_l2={e:.3e}")
    assert e < 3e-3


def test_bsrnn_se_fused_and_unfused_schedules_agree(monkeypatch):
    """The default schedule (fused layer kernel on both axes) and the separate input-projection GEMM + flag-group
    recurrence compute the same network: outputs agree to the fp16 noise floor of the gates_x rounding."""
    from urgent2026_challenge_track1_b200 import BSRNN_SE, runtime_tc as tc
    torch.manual_seed(0)
    m = BSRNN_SE(num_channel=196, num_layer=2, precision="fp16").cuda()
    fs, n = 16000, 16000
    x = R.synth_noisy(3, n, fs, seed=1)
    lens = torch.tensor([n, n - 555, n - 1999])
    monkeypatch.setattr(tc, "FUSED_AXES", ("time", "freq"))          # the default (BSRNN_LSTM_FUSED may override it)
    a = m(x, lens, fs)[0].clone()
    monkeypatch.setattr(tc, "FUSED_AXES", ())
    b = m(x, lens, fs)[0].clone()
    monkeypatch.setattr(tc, "FUSED_AXES", ("freq",))
    c = m(x, lens, fs)[0].clone()
    e_ab, e_ac = rel_l2(a.cpu(), b.cpu()), rel_l2(a.cpu(), c.cpu())
    print(f"fused vs unfused: {e_ab:.3e}, fused vs freq-only fused: {e_ac:.3e}")
    assert 0 < e_ab < 2e-3 and 0 < e_ac < 2e-3


@pytest.mark.parametrize("fs,B,n", [(48000, 3, 24000), (16000, 2, 9000), (44100, 2, 15000)])
def test_band_split_tc_matches_f32_kernel(rt, fs, B, n):
    """Tensor-core BandSplit (operand builder + store-only tcgen05 GEMM per band) against the f32 CUDA-core kernel on the
    same spectrum, incl. the truncated last band (16 k / 44.1 k); its epilogue statistics against bsrnn_gn_stats."""
    from urgent2026_challenge_track1_b200 import BSRNN_SE, runtime_tc as tc, _lib as L
    torch.manual_seed(0)
    m = BSRNN_SE(num_channel=196, num_layer=1, precision="fp16").cuda()
    core = m.bsrnn.bsrnn
    n_fft, hop = rt.stft_dims(fs, m.N_FFT, m.HOP, m.DEFAULT_FS)
    plan = rt.BandPlan.make(core.band_split.subbands, n_fft // 2 + 1)
    x = R.synth_noisy(B, n, fs, seed=2).cuda()
    lens = torch.tensor([n - 311 * i for i in range(B)], dtype=torch.int32).cuda()
    spec, bstats = rt.stft(x, lens, n_fft, hop, plan=plan)
    ref = rt.band_split_f32(spec, plan, m._bs.get(), 196, stats=bstats.clone())
    got = tc.band_split_tc(spec, plan, m._bs.get(), m._bs_tc.get(), 196, bstats)
    torch.cuda.synchronize()
    e = rel_l2(got.cpu(), ref.cpu())
    print(f"fs={fs} band split tc vs f32: {e:.3e}")
    assert got.shape == ref.shape and 0 < e < 1e-3
    Bq, T, K, N = got.shape
    want = torch.zeros(Bq, 2, dtype=torch.float64, device="cuda")
    L.call("bsrnn_gn_stats", got.data_ptr(), want.data_ptr(), Bq, T * K, N, N, L.stream_ptr())
    have = tc.workspace(Bq, T, K, N, got.device).stats
    torch.cuda.synchronize()
    assert torch.allclose(have[:Bq].cpu(), want.cpu(), rtol=1e-5, atol=1e-3)


def test_decoder_statistics_from_last_linear_epilogue(monkeypatch):
    """Per-(sample, band) sums taken by the last Linear + skip epilogue (stats_inner = K) equal a separate pass of
    bsrnn_band_stats over the final stream, and the enhanced output agrees with the separate-pass schedule."""
    from urgent2026_challenge_track1_b200 import BSRNN_SE, runtime as rtm, runtime_tc as tc, _lib as L
    torch.manual_seed(0)
    m = BSRNN_SE(num_channel=196, num_layer=2, precision="fp16").cuda()
    fs, n = 24000, 12000
    x = R.synth_noisy(3, n, fs, seed=5)
    lens = torch.tensor([n, n - 700, n - 1301])
    seen = {}
    orig = rtm._decoder_norm_tables

    def spy(skip, packs, stats=None):
        if stats is not None:
            Bq, T, K, N = skip.shape
            ref = torch.empty(Bq, K, 2, dtype=torch.float64, device=skip.device)
            off = rtm._i32([k * N for k in range(K)], skip.device)
            wid = rtm._i32([N] * K, skip.device)
            L.call("bsrnn_band_stats", skip.data_ptr(), ref.data_ptr(), Bq, T, K * N, off.data_ptr(), wid.data_ptr(), K, L.stream_ptr())
            torch.cuda.synchronize()
            seen["ref"], seen["got"] = ref.cpu(), stats.cpu().clone()
        return orig(skip, packs, stats=stats)

    monkeypatch.setattr(rtm, "_decoder_norm_tables", spy)
    a = m(x, lens, fs)[0].clone()
    assert "got" in seen and torch.allclose(seen["got"], seen["ref"], rtol=1e-5, atol=1e-3)
    monkeypatch.setattr(tc, "FC_EPI", L.TC_RESID_F32)               # register-staged epilogue: separate statistics pass
    monkeypatch.setattr(tc, "BAND_SPLIT_TC", False)
    b = m(x, lens, fs)[0].clone()
    e = rel_l2(a.cpu(), b.cpu())
    print(f"fused statistics + tc band split vs separate passes: {e:.3e}")
    assert 0 < e < 2e-3


def test_weight_resident_gemm_paths_match_a_small_batch():
    """The weight-resident / specialised GEMM schedules only start at >= 74 row tiles (9 472 tokens per band), which none of
    the oracle-sized tests reach: rows of a 12 x 8 s batch must equal the same utterances in a batch of 2 (generic schedules),
    and the batch of 2 must match the oracle."""
    from urgent2026_challenge_track1_b200 import BSRNN_SE
    torch.manual_seed(0)
    m = BSRNN_SE(num_channel=196, num_layer=1, precision="fp16")
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    m.cuda()
    fs, n, B = 48000, 8 * 48000, 12
    base = R.synth_noisy(3, n, fs, seed=7)
    x = base[torch.tensor([i % 3 for i in range(B)])].contiguous()
    lens = torch.full((B,), n, dtype=torch.int32)
    big = m(x, lens, fs)[0].cpu()
    small = m(base[:2].contiguous(), lens[:2], fs)[0].cpu()
    e0, e1 = rel_l2(big[0], small[0]), rel_l2(big[1], small[1])
    print(f"12 x 8 s batch vs batch of 2: rel_l2 {e0:.2e} {e1:.2e}")
    assert e0 < 2e-3 and e1 < 2e-3
    with torch.no_grad():
        ref, _ = R.bsrnn_se_forward(sd, base[:1], lens[:1].long(), fs, num_layer=1)
    e = rel_l2(big[:1], ref)
    print(f"12 x 8 s batch row 0 vs oracle: {e:.3e}")
    assert e < 1e-2


@pytest.mark.parametrize("width,fs", [(192, 16000), (32, 48000), (64, 22050)])
def test_bsrnn_se_tensorcore_other_widths_vs_oracle(width, fs):
    """Tensor-core mode at widths other than the published 196 (the reference constructor's default is num_channel=192): the
    step-wise tensor-core BLSTM kernels serve any num_channel % 8 == 0.  Bar 1e-2 (16-bit mode)."""
    from urgent2026_challenge_track1_b200 import BSRNN_SE
    torch.manual_seed(0)
    m = BSRNN_SE(num_channel=width, num_layer=2, precision="fp16")
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    m.cuda()
    n = fs // 2
    x = R.synth_noisy(2, n, fs, seed=1)
    lens = torch.tensor([n, n - 801])
    with torch.no_grad():
        ref_wav, ref_spec = R.bsrnn_se_forward(sd, x, lens, fs, num_layer=2)
    out, spec = m(x, lens, fs)
    e_w, e_s = rel_l2(out.cpu(), ref_wav), rel_l2(spec.cpu(), ref_spec)
    print(f"width={width} fs={fs} tensor-core rel_l2 wav={e_w:.3e} spec={e_s:.3e}")
    assert e_w < 1e-2 and e_s < 1e-2
    with pytest.raises(NotImplementedError):
        BSRNN_SE(num_channel=20, num_layer=1, precision="fp16").cuda()(x, lens, fs)


def test_lstm_step_tc_vs_torch_h768():
    """bsrnn_lstm_step_tc at the FlowSE width (N = 384, H = 768), ragged last tile: against torch.nn.LSTM on the CPU."""
    from urgent2026_challenge_track1_b200 import runtime_tc_steps as S, _lib as L
    torch.manual_seed(0)
    N, H, Rr, steps = 384, 768, 150, 9
    rnn = torch.nn.LSTM(N, H, batch_first=True, bidirectional=True)
    x = torch.randn(Rr, steps, N) * 0.7
    with torch.no_grad():
        ref = rnn(x)[0]
    rnn = rnn.cuda()
    p = S.pack_lstm_steps_tc(rnn)
    tiles = (Rr + 127) // 128
    ws = S.StepsWorkspace(steps, tiles, H, "cuda")
    xhat = torch.empty(steps * tiles * p["kc_in"] * 1024, dtype=torch.float16, device="cuda")
    xg = x.cuda().contiguous()
    L.call("bsrnn_norm_cast_kb8", xg.data_ptr(), None, None, xhat.data_ptr(), N, 0, N, p["kc_in"],
           steps * tiles, tiles, Rr, 1 << 40, 0, steps, 1, Rr * steps, 1, L.stream_ptr())
    S.blstm_steps_tc(xhat, p, steps, tiles, ws)
    outs = []
    for d in (0, 1):
        yd = ws.y[d].view(steps, tiles, H // 8, 128, 8).permute(0, 1, 3, 2, 4).reshape(steps, tiles * 128, H)[:, :Rr]
        outs.append(yd.permute(1, 0, 2).float().cpu())
    assert rel_l2(torch.cat(outs, 2), ref) < 3e-3


@pytest.mark.parametrize("Rr,steps,slots", [(150, 9, 0), (700, 5, 2), (130, 33, 1), (1000, 4, 3)])
def test_blstm_fused768_vs_torch(Rr, steps, slots):
    """bsrnn_blstm_fused768_tc (FlowSE width N = 384 / H = 768: input projection inside the persistent recurrence, groups
    of 24 CTA pairs) against torch.nn.LSTM on the CPU; ragged last tile, odd / even tile counts, 1..3 interleaved pairs."""
    from urgent2026_challenge_track1_b200 import runtime_tc_steps as S, _lib as L
    torch.manual_seed(0)
    N, H = 384, 768
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
    L.call("bsrnn_norm_cast_kb8_ones", xg.data_ptr(), None, None, xhat.data_ptr(), N, 0, N, p["kc_fused"],
           steps * tiles, tiles, Rr, 1 << 40, 0, steps, 1, Rr * steps, 1, p["one_col"], st)
    for _ in range(2):
        L.call("bsrnn_blstm_fused768_tc", xhat.data_ptr(), p["wfused"].data_ptr(), ws.zero.data_ptr(), ws.y[0].data_ptr(),
               ws.y[1].data_ptr(), (H // 8) * 1024, Rr, steps, tiles, 0, slots, ws.sync.data_ptr(), st)
    outs = []
    for d in (0, 1):
        yd = ws.y[d].view(steps, tiles, H // 8, 128, 8).permute(0, 1, 3, 2, 4).reshape(steps, tiles * 128, H)[:, :Rr]
        outs.append(yd.permute(1, 0, 2).float().cpu())
    e = rel_l2(torch.cat(outs, 2), ref)
    print(f"fused768 R={Rr} steps={steps} slots={slots}: rel_l2={e:.3e}")
    assert e < 3e-3


def test_flowse_solver_registry_errors():
    from urgent2026_challenge_track1_b200.sampling import ODEsolverRegistry
    assert set(ODEsolverRegistry.get_all_names()) >= {"euler", "midpoint", "heun"}
    with pytest.raises(ValueError):
        ODEsolverRegistry.get_by_name("rk45")


def test_flowse_midpoint_heun_vs_oracle():
    g = golden("flowse_n16_l1.npz")
    m = _flow_model(g)
    sd = golden_sd(g)
    fs = 16000
    y, lens = torch.from_numpy(g[f"in/{fs}/wav"]), torch.from_numpy(g[f"in/{fs}/lens"])
    Y = m.speech_to_feature(y, fs, lens)
    for solver in ("midpoint", "heun"):
        torch.manual_seed(11)
        z = torch.randn_like(Y).cpu()                   # the draw prior_sampling will make on the GPU
        with torch.no_grad():
            ref = R.flowse_enhance(sd, y, fs, lens, N=2, z=z, num_layer=1, solver=solver)
        torch.manual_seed(11)
        out = m.enhance(y, fs, lens, N=2, solver=solver)
        assert rel_l2(out.cpu(), ref) < 1e-3
