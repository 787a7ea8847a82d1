"""CPU tests pinning oracle/restated.py: (a) against the committed golden vectors produced by the reference's own
files run verbatim (tests/golden/make_golden.py), (b) against that verbatim import itself when /root/reference is
present, (c) against the reference's only known-answer: the parameter counts in conf/models/BSRNN_baseline.yaml:30-32."""
import numpy as np
import pytest
import torch

from conftest import golden, golden_sd, rel_l2
from oracle import restated as R

RATES = (8000, 16000, 22050, 24000, 32000, 44100, 48000)
KPRIME = {8000: 20, 16000: 27, 22050: 28, 24000: 29, 32000: 31, 44100: 34, 48000: 34}


@pytest.mark.parametrize("fs", RATES)
def test_bsrnn_se_matches_golden(fs):
    g = golden("bsrnn_se_n16_l2.npz")
    sd = golden_sd(g)
    wav, lens = torch.from_numpy(g[f"in/{fs}/wav"]), torch.from_numpy(g[f"in/{fs}/lens"])
    with torch.no_grad():
        out, spec = R.bsrnn_se_forward(sd, wav, lens, fs, num_layer=2)
    assert out.shape == g[f"out/{fs}/wav"].shape
    assert rel_l2(out, g[f"out/{fs}/wav"]) < 2e-6          # f32 re-association only
    assert rel_l2(spec, g[f"out/{fs}/spec"]) < 2e-6


@pytest.mark.parametrize("fs", (16000, 22050, 48000))
def test_flowse_matches_golden(fs):
    g = golden("flowse_n16_l1.npz")
    sd = golden_sd(g)
    y, lens = torch.from_numpy(g[f"in/{fs}/wav"]), torch.from_numpy(g[f"in/{fs}/lens"])
    z, t = torch.from_numpy(g[f"in/{fs}/z"]), torch.from_numpy(g[f"in/{fs}/t"])
    with torch.no_grad():
        spec, _ = R.stft_encode(y, lens, fs, 1536, 384, 48000, "exponent", 0.667, 0.065)
        Y = spec.permute(0, 2, 1).unsqueeze(1)
        assert rel_l2(Y, g[f"out/{fs}/feature"]) < 2e-6
        vf = -R.flow_bsrnn_forward(sd, torch.cat([Y + 0.5 * z, Y], dim=1), t, 769, num_layer=1)
        assert rel_l2(vf, g[f"out/{fs}/vf"]) < 5e-6
        enh = R.flowse_enhance(sd, y, fs, lens, N=3, z=z, num_layer=1)
    assert rel_l2(enh, g[f"out/{fs}/enhanced"]) < 1e-5


def test_losses_match_golden():
    g = golden("losses.npz")
    a, b = torch.from_numpy(g["target"]), torch.from_numpy(g["estimate"])
    assert np.allclose(R.multires_l1_spec_loss(a, b).numpy(), g["mr_l1"], rtol=1e-5)
    assert np.allclose(R.si_snr_loss(a, b).numpy(), g["sisnr"], rtol=1e-5)


def test_euler_schedule_quirk():
    # sampling/__init__.py:48-56 — last step size is t_{N-1} itself (integrate to 0), not a difference
    ts, steps = R.euler_schedule(1.0, 0.03, 15)
    assert torch.allclose(ts, torch.linspace(1.0, 0.03, 15))
    assert torch.allclose(steps[:-1], ts[:-1] - ts[1:]) and float(steps[-1]) == pytest.approx(0.03)
    assert float(steps.sum()) == pytest.approx(1.0)


def test_frame_lengths_and_dims():
    # SURVEY.md §8a per-rate table: n_fft / hop per fs
    table = {8000: (160, 80), 16000: (320, 160), 22050: (441, 220), 24000: (480, 240), 32000: (640, 320),
             44100: (882, 441), 48000: (960, 480)}
    for fs, (nf, hp) in table.items():
        assert R.stft_dims(fs, 960, 480) == (nf, hp)
    assert R.stft_dims(22050, 1536, 384) == (705, 176) and R.stft_dims(44100, 1536, 384) == (1411, 352)
    assert int(R.frame_lengths(torch.tensor([480000]), 960, 480)) == 1001


# ---------------------------------------------------------------- checks that need /root/reference (build container)
@pytest.mark.reference
def test_known_answer_param_counts(ref_ns):
    """conf/models/BSRNN_baseline.yaml:30-32: 'Parameters: 36.01795196533203 M' @48k, '32.0456657409668 M' @16k
    (unit 2**20, GroupNorm affine parameters not counted)."""
    m = ref_ns.BSRNN_SE(num_channel=196, num_layer=6)
    total = sum(p.numel() for p in m.parameters())
    assert total == 37_800_844
    gn = sum(p.numel() for mod in m.modules() if isinstance(mod, torch.nn.GroupNorm) for p in mod.parameters())
    assert (total - gn) / 2 ** 20 == pytest.approx(36.01795196533203, abs=1e-9)
    # @16 kHz only K'=27 bands of band_split / mask_decoder are touched
    b = m.bsrnn.bsrnn
    used = 0
    for name, p in b.named_parameters():
        parts = name.split(".")
        if parts[0] == "band_split" and int(parts[2]) >= 27:
            continue
        if parts[0] == "mask_decoder" and int(parts[2]) >= 27:
            continue
        used += p.numel()
    gn16 = 0
    for name, mod in b.named_modules():
        if isinstance(mod, torch.nn.GroupNorm):
            parts = name.split(".")
            if parts[0] in ("band_split", "mask_decoder") and int(parts[2]) >= 27:
                continue
            gn16 += sum(p.numel() for p in mod.parameters())
    assert (used - gn16) / 2 ** 20 == pytest.approx(32.0456657409668, abs=1e-9)
    cfg = __import__("oracle.ref_loader", fromlist=["x"]).flowse_config(ref_ns)
    assert sum(p.numel() for p in ref_ns.FlowSEModel(cfg).parameters()) == 103_245_488


@pytest.mark.reference
@pytest.mark.parametrize("fs", RATES)
def test_restated_vs_verbatim_bsrnn(ref_ns, fs):
    torch.manual_seed(fs)
    m = ref_ns.BSRNN_SE(num_channel=24, num_layer=1).eval()
    n = fs // 4
    x = R.synth_noisy(2, n, fs, seed=3)
    lens = torch.tensor([n, n - 1234])
    with torch.no_grad():
        w, s = m(x, lens, fs)
        w2, s2 = R.bsrnn_se_forward(m.state_dict(), x, lens, fs, num_layer=1)
        z = R.band_split(m.state_dict(), "bsrnn.bsrnn.band_split.", torch.view_as_real(s2), R.SUBBANDS_481)
    assert z.shape[2] == KPRIME[fs]
    assert rel_l2(w2, w) < 2e-6 and rel_l2(s2, s) < 2e-6


@pytest.mark.reference
@pytest.mark.parametrize("solver", ("euler", "midpoint", "heun"))
def test_restated_vs_verbatim_solvers(ref_ns, solver):
    from oracle import ref_loader
    cfg = ref_loader.flowse_config(ref_ns, bsrnn_hidden=16, num_layer=1)
    torch.manual_seed(1)
    fm = ref_ns.FlowSEModel(cfg).eval(no_ema=True)
    fs, n = 16000, 3200
    y = R.synth_noisy(1, n, fs, seed=9)
    lens = torch.tensor([n])
    with torch.no_grad():
        Y = fm.speech_to_feature(y, fs, lens)
        torch.manual_seed(5)
        # euler is the only name get_white_box_solver defines a schedule for (sampling/__init__.py:47-48); drive the
        # registered update_fn classes with that same schedule for the other two.
        cls = ref_ns.sampling.ODEsolverRegistry.get_by_name(solver)(fm.ode, fm)
        x, _ = fm.ode.prior_sampling(Y.shape, Y)
        ts = torch.linspace(cfg.T_rev, cfg.t_eps, 3)
        for i in range(3):
            step = ts[i] - ts[i + 1] if i != 2 else ts[-1]
            x = cls.update_fn(x, torch.ones(1) * ts[i], Y, step)
        torch.manual_seed(5)
        z = torch.randn_like(Y)
        x0 = R.fm_prior(Y, z, 0.05, 0.5)
        vf = lambda a, t, b: -R.flow_bsrnn_forward(fm.state_dict(), torch.cat([a, b], 1), t, 769, 1)
        mine = R.ode_sample(vf, Y, x0, 1.0, 0.03, 3, solver)
    assert rel_l2(mine, x) < 5e-6
