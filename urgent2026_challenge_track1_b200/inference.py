"""Batched inference front-end — mirrors ``baseline_code/inference.py`` (reference inference.py:26-112: same flags,
same scp in / ``inf.scp`` + ``wav/<uid>.wav`` out, peak-normalise to 0.9 (:60), PCM-16 WAV (:62)) but groups
utterances into (fs, length) buckets and runs them as batches, sharded over ranks when launched under torchrun
(SURVEY.md §8f.1).  WAV I/O uses scipy.io.wavfile (soundfile is not a dependency of this package)."""
from __future__ import annotations

import argparse
import os

import numpy as np
import torch

from .d_model import SEModel
from .pipeline import StreamedEnhancer
from .sharding import shard_utterances


def _read_wav(path):
    from scipy.io import wavfile
    sr, x = wavfile.read(path)
    if x.dtype.kind == "i":
        x = x.astype(np.float32) / float(np.iinfo(x.dtype).max + 1)
    elif x.dtype.kind == "u":
        x = (x.astype(np.float32) - 128.0) / 128.0
    if x.ndim > 1:
        x = x[:, 0]
    return x.astype(np.float32), int(sr)


def _write_wav(path, x, sr):
    from scipy.io import wavfile
    wavfile.write(path, sr, np.clip(np.round(x * 32767.0), -32768, 32767).astype(np.int16))


def main(args):
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dev = torch.device(args.device, int(os.environ.get("LOCAL_RANK", 0))) if args.device == "cuda" else torch.device(args.device)
    model = SEModel.load_from_checkpoint(args.ckpt_path, map_location=dev, precision=args.precision).eval()
    utts = []
    with open(args.input_scp) as f:
        for line in f:
            uid, wav = line.strip().split()
            utts.append((uid, wav))
    os.makedirs(os.path.join(args.output_dir, "wav"), exist_ok=True)
    audio = [_read_wav(p) for _, p in utts]
    batches = shard_utterances([len(a) for a, _ in audio], [sr for _, sr in audio], rank, world, args.max_batch)
    with open(os.path.join(args.output_dir, f"inf.{rank}.scp" if world > 1 else "inf.scp"), "w") as f:
        def host_batches():
            for fs, idx in batches:
                lens = torch.tensor([len(audio[i][0]) for i in idx], dtype=torch.int32)
                batch = torch.zeros(len(idx), int(lens.max()), dtype=torch.float32).pin_memory()
                for row, i in enumerate(idx):
                    batch[row, : lens[row]] = torch.from_numpy(audio[i][0])      # right zero-pad (dataset.py:404-441)
                yield batch, lens, fs

        # H2D of the next batch and D2H of the previous one overlap the current batch's kernels
        for (fs, idx), (enhanced, lens, _) in zip(batches, StreamedEnhancer(model.se_model).run(host_batches())):
            for row, i in enumerate(idx):
                y = enhanced[row, : lens[row]].clone()
                y = y / y.abs().max().clamp_min(1e-12) * 0.9                    # inference.py:60
                out = os.path.join(args.output_dir, "wav", f"{utts[i][0]}.wav")
                _write_wav(out, y.numpy(), fs)
                print(f"{utts[i][0]} {out}", file=f)
    print("done")


if __name__ == "__main__":
    p = argparse.ArgumentParser()
    p.add_argument("--input_scp", type=str, required=True)
    p.add_argument("--output_dir", type=str, default="./tmp/se")
    p.add_argument("--ckpt_path", type=str, default="./tmp/se")
    p.add_argument("--device", type=str, default="cuda")
    p.add_argument("--precision", type=str, default=None)
    p.add_argument("--max_batch", type=int, default=64)
    main(p.parse_args())
