"""2+ rank NCCL check of the training step (run under torchrun): every rank trains on its own batch with its own
sample rate; after the step all ranks must hold identical parameters, and the averaged flat gradient must equal the
mean of the per-rank gradients gathered explicitly."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from urgent2026_challenge_track1_b200 import BSRNN_SE
from urgent2026_challenge_track1_b200.training import SETrainer

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)                                              # identical initial weights on every rank
m = BSRNN_SE(32, 2, precision="fp32").to(dev)
tr = SETrainer(m, lr=1e-3)
fs = (16000, 48000, 8000, 44100)[rank % 4]                        # rank-dependent K': unused bands contribute zeros
n = fs // 2
g = torch.Generator().manual_seed(10 + rank)
clean = (0.05 * torch.randn(2, 1, n, generator=g)).to(dev)
noisy = clean + (0.03 * torch.randn(2, 1, n, generator=g)).to(dev)
lens = torch.tensor([n, n - 300], dtype=torch.int32)
tr.flat.zero_grad()
loss, _ = tr.loss(noisy, clean, lens, torch.tensor(fs, dtype=torch.int32))
loss.backward()
tr.flat.gather_grads()
mine = tr.flat.grad.clone()
allg = [torch.zeros_like(mine) for _ in range(world)]
dist.all_gather(allg, mine)
want = torch.stack(allg).mean(0)
tr.apply_gradients()
got = tr.flat.grad / world
e_g = float((got - want).norm() / want.norm())
ps = [torch.zeros_like(tr.flat.flat) for _ in range(world)]
dist.all_gather(ps, tr.flat.flat)
same = all(torch.equal(ps[0], p) for p in ps)
moved = float((ps[0] - ps[0].new_tensor(0)).abs().sum()) > 0
print(f"rank {rank}/{world} fs={fs} loss={float(loss):.4f} grad_err={e_g:.2e} params_identical={same} "
      f"{'OK' if (e_g < 1e-6 and same and moved) else 'FAIL'}", flush=True)
dist.destroy_process_group()
sys.exit(0 if (e_g < 1e-6 and same) else 1)
