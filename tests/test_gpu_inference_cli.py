"""GPU test of the inference entry point (mirror of baseline_code/inference.py:26-112): scp in, inf.scp + wav/<uid>.wav out,
SEModel first with the FlowSEModel fallback, peak normalisation, PCM-16 — and batched results equal to batch-1 results."""
import os

import numpy as np
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def _write_inputs(tmp_path, specs):
    from urgent2026_challenge_track1_b200.inference import _write_wav
    from urgent2026_challenge_track1_b200.synth import synth_noisy
    lines = []
    for i, (fs, n) in enumerate(specs):
        x = synth_noisy(1, n, fs, seed=100 + i)[0].numpy()
        p = str(tmp_path / f"in_{i}.wav")
        _write_wav(p, x, fs)
        lines.append(f"utt{i} {p}")
    scp = tmp_path / "wav.scp"
    scp.write_text("\n".join(lines) + "\n")
    return str(scp)


def test_inference_cli_semodel_batches_equal_batch1(tmp_path):
    from urgent2026_challenge_track1_b200 import inference as I
    from urgent2026_challenge_track1_b200.checkpoint import save_checkpoint
    from urgent2026_challenge_track1_b200.config import Config
    from urgent2026_challenge_track1_b200.d_model import SEModel
    torch.manual_seed(0)
    cfg = Config(se_model="bsrnn", model_configs={"num_channel": 32, "num_layer": 1})
    m = SEModel(cfg, precision="fp32")
    ckpt = save_checkpoint(str(tmp_path / "se.ckpt"), m, cfg)
    specs = [(16000, 9000), (16000, 9000), (16000, 7000), (8000, 5000), (16000, 9000)]
    scp = _write_inputs(tmp_path, specs)
    out_dir = str(tmp_path / "out")
    args = I.build_parser().parse_args(["--input_scp", scp, "--output_dir", out_dir, "--ckpt_path", ckpt, "--precision", "fp32",
                                        "--max_batch", "4"])
    I.main(args)
    listed = dict(l.split() for l in open(os.path.join(out_dir, "inf.scp")).read().strip().splitlines())
    assert set(listed) == {f"utt{i}" for i in range(len(specs))}
    model = I.load_model(ckpt, torch.device("cuda"), "fp32")
    for i, (fs, n) in enumerate(specs):
        x, sr = I._read_wav(str(tmp_path / f"in_{i}.wav"))
        assert sr == fs and len(x) == n
        one, _ = model.se_model(torch.from_numpy(x).view(1, -1), torch.tensor([n]), fs)      # the reference's batch-1 call
        one = (one[0] / one[0].abs().max() * 0.9).cpu().numpy()
        got, sr2 = I._read_wav(listed[f"utt{i}"])
        assert sr2 == fs and got.shape == one.shape
        assert np.abs(got - one).max() < 2.0 / 32768                                          # PCM-16 quantisation only


def test_inference_cli_falls_back_to_flowse(tmp_path):
    from urgent2026_challenge_track1_b200 import inference as I
    from urgent2026_challenge_track1_b200.checkpoint import save_checkpoint
    from urgent2026_challenge_track1_b200.config import Config
    from urgent2026_challenge_track1_b200.flow_model import FlowSEModel
    torch.manual_seed(0)
    cfg = Config(model_type="flowse", ema_decay=0.999, sigma_max=0.5, sigma_min=0.05, t_eps=0.03, T_rev=1.0, loss_type="mse",
                 loss_abs_exponent=0.5, n_fft=1536, hop_length=384, spec_transform_type="exponent", spec_abs_exponent=0.667,
                 spec_factor=0.065, bsrnn_hidden=16, num_layer=1, learning_rate=1e-4)
    fm = FlowSEModel(cfg)
    with torch.no_grad():
        fm.dnn.condition_fc.bias.add_(0.3)
    fm.ema.update(fm.parameters())
    ckpt = save_checkpoint(str(tmp_path / "flow.ckpt"), fm, cfg)
    scp = _write_inputs(tmp_path, [(16000, 6000), (16000, 6000)])
    out_dir = str(tmp_path / "out")
    args = I.build_parser().parse_args(["--input_scp", scp, "--output_dir", out_dir, "--ckpt_path", ckpt, "--nfe", "2"])
    I.main(args)
    model = I.load_model(ckpt, torch.device("cuda"))
    assert isinstance(model, FlowSEModel) and model.ema.num_updates == 1      # EMA restored, eval() swapped it in
    for i in range(2):
        y, sr = I._read_wav(os.path.join(out_dir, "wav", f"utt{i}.wav"))
        assert sr == 16000 and len(y) == 6000 and np.isfinite(y).all() and abs(np.abs(y).max() - 0.9) < 1e-3
