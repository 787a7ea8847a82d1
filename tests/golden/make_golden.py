"""Generate tests/golden/*.npz from the reference's own files imported VERBATIM (oracle/ref_loader.py) — run in
the build container only:  python tests/golden/make_golden.py
Each fixture stores the random-init state_dict (small model sizes so the files stay small), the inputs and the
reference outputs.  torch 2.11.0 CPU, seeds fixed below."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader, restated as R  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def sd_np(sd, prefix=""):
    return {"sd/" + prefix + k: v.detach().cpu().numpy() for k, v in sd.items()}


def main():
    ns = ref_loader.load()
    torch.set_num_threads(4)
    # ---- BSRNN_SE, small width, every sample rate of the reference's rate set, ragged lengths
    torch.manual_seed(0)
    m = ns.BSRNN_SE(num_channel=16, num_layer=2).eval()
    case = sd_np(m.state_dict())
    for fs in (8000, 16000, 22050, 24000, 32000, 44100, 48000):
        n = fs // 5
        x = R.synth_noisy(2, n, fs, seed=fs)
        lens = torch.tensor([n, n - fs // 16])
        with torch.no_grad():
            wav, spec = m(x, lens, fs)
        case[f"in/{fs}/wav"] = x.numpy()
        case[f"in/{fs}/lens"] = lens.numpy()
        case[f"out/{fs}/wav"] = wav.numpy()
        case[f"out/{fs}/spec"] = spec.numpy()
    np.savez_compressed(os.path.join(OUT, "bsrnn_se_n16_l2.npz"), **case)

    # ---- FlowSE enhance (Euler N=3), small width; z drawn exactly as odes.py:88 does after manual_seed
    cfg = ref_loader.flowse_config(ns, bsrnn_hidden=16, num_layer=1)
    torch.manual_seed(0)
    fm = ns.FlowSEModel(cfg).eval(no_ema=True)
    case = sd_np(fm.state_dict())
    for fs in (16000, 22050, 48000):
        n = fs // 5
        y = R.synth_noisy(2, n, fs, seed=fs + 1)
        lens = torch.tensor([n, n - 301])
        torch.manual_seed(11)
        with torch.no_grad():
            enh = fm.enhance(y, fs, lens, N=3)
            Y = fm.speech_to_feature(y, fs, lens)
        torch.manual_seed(11)
        z = torch.randn_like(Y)
        t = torch.tensor([0.7, 0.31])
        with torch.no_grad():
            vf = fm(Y + 0.5 * z, t, Y)                       # one vector-field evaluation, flow_model.py:203-209
        case[f"in/{fs}/wav"] = y.numpy()
        case[f"in/{fs}/lens"] = lens.numpy()
        case[f"in/{fs}/z"] = z.numpy()
        case[f"in/{fs}/t"] = t.numpy()
        case[f"out/{fs}/feature"] = Y.numpy()
        case[f"out/{fs}/vf"] = vf.numpy()
        case[f"out/{fs}/enhanced"] = enh.numpy()
    np.savez_compressed(os.path.join(OUT, "flowse_n16_l1.npz"), **case)

    # ---- losses (d_model.py:24-25) on a fixed pair
    sem_cfg = ns.Config(se_model="bsrnn", model_configs={"num_channel": 16, "num_layer": 1})
    torch.manual_seed(0)
    sem = ns.SEModel(sem_cfg)
    a = R.synth_noisy(2, 9600, 48000, seed=5)
    b = a + 0.01 * torch.randn(a.shape, generator=torch.Generator().manual_seed(6))
    np.savez_compressed(os.path.join(OUT, "losses.npz"), target=a.numpy(), estimate=b.numpy(),
                        mr_l1=sem.mr_l1_loss(a, b).numpy(), sisnr=sem.sisnr_loss(a, b).numpy())
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
