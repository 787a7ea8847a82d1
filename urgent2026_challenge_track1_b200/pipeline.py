"""Host <-> device pipelining of the batched inference loop (the reference's inference.py:36-64 is
batch-1 and synchronous).

``StreamedEnhancer`` pushes successive HOST batches through ``BSRNN_SE.forward``: the H2D copy of batch i+1 and the
D2H copy of batch i-1 run on their own CUDA streams while batch i computes, so a step costs max(compute, copies)
instead of their sum (at BASELINE config 2 a batch is 123 MB each way).  Inputs should be pinned; outputs land in
pinned buffers owned by the enhancer (one flat buffer per slot, grown to the largest batch seen and viewed per batch,
reused round-robin: consume a result before asking for the one after the next).  Memory is therefore bounded by
`depth` x the largest batch, however many different (B, L) shapes a run produces."""
from __future__ import annotations

import torch

from . import _lib as L


class StreamedEnhancer:
    def __init__(self, model, depth=2):
        L.require_device()
        self.model = model
        self.dev = model._device()
        if self.dev.type != "cuda":
            raise L.NativeLibraryError("StreamedEnhancer needs the model on a CUDA device (no CPU fallback)")
        self.depth = int(depth)
        self.h2d = torch.cuda.Stream(self.dev)
        self.d2h = torch.cuda.Stream(self.dev)
        self._in = [None] * self.depth         # slot -> flat device staging buffer (grown on demand)
        self._out = [None] * self.depth        # slot -> flat pinned host buffer (grown on demand)
        self._free = [None] * self.depth      # event: compute that read staging slot s has finished
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    @staticmethod
    def _numel(shape):
        n = 1
        for d in shape:
            n *= int(d)
        return n

    def _stage(self, slot, shape):
        n = self._numel(shape)
        buf = self._in[slot]
        if buf is None or buf.numel() < n:
            if buf is not None and self._free[slot] is not None:
                self._free[slot].synchronize()     # the forward still reading the old buffer must finish before it goes
            buf = self._in[slot] = torch.empty(n, dtype=torch.float32, device=self.dev)
            # the caching allocator may hand out a block that work still queued on the compute stream writes (a
            # tensor freed a moment ago): the copy stream must not touch it before that work has drained
            self.h2d.wait_stream(torch.cuda.current_stream(self.dev))
        return buf[:n].view(tuple(shape))

    def _host_out(self, slot, shape):
        n = self._numel(shape)
        buf = self._out[slot]
        if buf is None or buf.numel() < n:
            self.d2h.synchronize()                 # a copy into the old pinned buffer may still be in flight
            buf = self._out[slot] = torch.empty(n, dtype=torch.float32).pin_memory()
        return buf[:n].view(tuple(shape))

    def _submit(self, i, wav_host, lens, fs):
        """Enqueue H2D (copy stream), forward (current stream) and D2H (copy-back stream) of batch i."""
        slot = i % self.depth
        cur = torch.cuda.current_stream(self.dev)
        x = self._stage(slot, wav_host.shape)
        with torch.cuda.stream(self.h2d):
            if self._free[slot] is not None:
                self.h2d.wait_event(self._free[slot])          # the forward that read this staging buffer is done
            x.copy_(wav_host, non_blocking=True)
            up = torch.cuda.Event()
            up.record(self.h2d)
        self.h2d_bytes += wav_host.numel() * 4
        cur.wait_event(up)
        wav, _ = self.model(x, lens, fs)
        done = torch.cuda.Event()
        done.record(cur)
        self._free[slot] = done
        out = self._host_out(slot, wav.shape)
        wav.record_stream(self.d2h)
        with torch.cuda.stream(self.d2h):
            self.d2h.wait_event(done)
            out.copy_(wav, non_blocking=True)
            back = torch.cuda.Event()
            back.record(self.d2h)
        self.d2h_bytes += wav.numel() * 4
        return out, back

    def run(self, batches):
        """batches: iterable of (wav_host (B,L) float32 [pinned], lengths (B,), fs).  Yields (enhanced_host (B, max
        len) pinned float32, lengths, fs) in order; a yielded buffer is overwritten `depth` batches later."""
        pending = None
        for i, (wav_host, lens, fs) in enumerate(batches):
            nxt = (self._submit(i, wav_host, lens, fs), lens, fs)
            if pending is not None:
                (out, ev), pl, pf = pending
                ev.synchronize()
                yield out, pl, pf
            pending = nxt
        if pending is not None:
            (out, ev), pl, pf = pending
            ev.synchronize()
            yield out, pl, pf
