"""Stage-by-stage numerical check of the CUDA path against the oracle; prints one line per stage.
Run on the GPU box:  python tools/gpu_check.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from oracle import restated as R
from urgent2026_challenge_track1_b200 import runtime as rt, BSRNN_SE, _lib


def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def main():
    _lib.require_device()
    torch.manual_seed(0)
    fs, N, nl = 16000, 196, 2
    m = BSRNN_SE(N, nl, precision="fp32")
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    m.cuda()
    n = fs
    x = R.synth_noisy(2, n, fs)
    lens = torch.tensor([n, n - 3000])
    n_fft, hop = R.stft_dims(fs, 960, 480)
    spec_ref, _ = R.stft_encode(x, lens, fs)
    spec = rt.stft(x.cuda(), lens.int().cuda(), n_fft, hop)
    print("stft", rel(torch.view_as_complex(spec), spec_ref))
    core = m.bsrnn.bsrnn
    plan = rt.BandPlan.make(core.band_split.subbands, n_fft // 2 + 1)
    p = "bsrnn.bsrnn."
    z_ref = R.band_split(sd, p + "band_split.", torch.view_as_real(spec_ref), R.SUBBANDS_481)
    z = rt.band_split_f32(spec, plan, m._bs.get(), N)
    print("band_split", rel(z, z_ref), tuple(z.shape))
    # one time-axis block by hand
    B, T, K, _ = z_ref.shape
    h = R._gn(z_ref.permute(0, 3, 1, 2), sd[p + "norm_time.0.weight"], sd[p + "norm_time.0.bias"]).permute(0, 2, 3, 1)
    hh = R.blstm(sd, p + "rnn_time.0.", h.permute(0, 2, 1, 3).reshape(B * K, T, N))
    y_ref = hh.reshape(B, K, T, 4 * N).permute(0, 2, 1, 3)
    lay = m._dual.get()[0]["time"]
    scale, shift = rt._layer_norm_tables(z, lay["gamma"], lay["beta"])
    hn = z * scale[:, None, None, :] + shift[:, None, None, :]
    print("gn_apply", rel(hn, h))
    skip_ref = R.dual_path_layers(sd, p, z_ref, nl)
    skip = rt.dual_path_f32(z.clone(), m._dual.get())
    print("dual_path", rel(skip, skip_ref))
    m_ref, r_ref = R.mask_decoder(sd, p + "mask_decoder.", skip_ref, R.SUBBANDS_481, 481)
    mk, rs = rt.mask_decoder_f32(skip, plan, m._md.get())
    Fb = spec_ref.size(2)
    print("mask", rel(torch.view_as_complex(mk), m_ref[..., :Fb]), "resid", rel(torch.view_as_complex(rs), r_ref[..., :Fb]))
    wav_ref, est_ref = R.bsrnn_se_forward(sd, x, lens, fs, nl)
    t0 = time.time()
    wav, est = m(x, lens, fs)
    torch.cuda.synchronize()
    print("forward wav", rel(wav, wav_ref), "spec", rel(est, est_ref), f"{time.time()-t0:.3f}s")


if __name__ == "__main__":
    main()
