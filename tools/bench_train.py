"""Training-step timing (BASELINE config 5: BSRNN_baseline, B=4 x 96000 samples @48 kHz per GPU; f32 kernels this round).
  python tools/bench_train.py [--batch 4] [--steps 5]          (or under torchrun for N GPUs: one allreduce per step)
Prints one JSON line: step time, audio-seconds trained per second, split forward / backward / optimizer tail."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from urgent2026_challenge_track1_b200 import BSRNN_SE
from urgent2026_challenge_track1_b200.training import SETrainer

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4); ap.add_argument("--samples", type=int, default=96000)
ap.add_argument("--steps", type=int, default=5); ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--channels", type=int, default=196); ap.add_argument("--layers", type=int, default=6)
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
m = BSRNN_SE(a.channels, a.layers, precision="fp32").to(dev)
tr = SETrainer(m, lr=1e-3)
g = torch.Generator().manual_seed(1 + rank)
clean = (0.05 * torch.randn(a.batch, 1, a.samples, generator=g)).to(dev)
noisy = clean + (0.03 * torch.randn(a.batch, 1, a.samples, generator=g)).to(dev)
lens = torch.full((a.batch,), a.samples, dtype=torch.int32)
fs = torch.tensor(48000, dtype=torch.int32)
ev = lambda: torch.cuda.Event(enable_timing=True)
for _ in range(a.warmup):
    tr.step(noisy, clean, lens, fs)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t_f = t_b = t_o = 0.0
e0, e3 = ev(), ev()
e0.record()
for _ in range(a.steps):
    a0, a1, a2, a3 = ev(), ev(), ev(), ev()
    tr.flat.zero_grad()
    a0.record(); loss, _ = tr.loss(noisy, clean, lens, fs); a1.record()
    loss.backward(); tr.flat.gather_grads(); a2.record()
    tr.apply_gradients(); a3.record()
    torch.cuda.synchronize()
    t_f += a0.elapsed_time(a1); t_b += a1.elapsed_time(a2); t_o += a2.elapsed_time(a3)
e3.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e3) / a.steps
t = torch.tensor([ms], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"metric": "BSRNN_baseline train step (f32 kernels)", "n_gpus": world, "ms_per_step": float(t[0]),
                      "audio_s_per_s": world * a.batch * a.samples / 48000 / (float(t[0]) / 1e3),
                      "fwd_ms": t_f / a.steps, "bwd_ms": t_b / a.steps, "allreduce_clip_adamw_ms": t_o / a.steps,
                      "loss": float(loss), "params": tr.flat.numel,
                      "config": {"batch_per_gpu": a.batch, "samples": a.samples, "fs": 48000, "channels": a.channels,
                                 "layers": a.layers}}))
if world > 1:
    dist.destroy_process_group()
