"""BSRNN_SE — drop-in for ``baseline_code/models/bsrnn.py::BSRNN_SE`` (reference bsrnn.py:9-41).

Same constructor (``num_channel=192, num_layer=6``), same forward signature and return value, same state_dict
keys (``bsrnn.bsrnn.*``); the STFT encoder / BSRNN separator / iSTFT decoder all run in the sm_100a
kernels of libbsrnn_b200 (include/bsrnn_b200.h).  Non-causal (bidirectional) only, num_spk=1, as the reference
instantiates it (bsrnn.py:27-34).
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from . import _lib as L
from . import runtime as R
from . import runtime_tc as TC
from .layers import BandSplitParams, MaskDecoderParams, add_dual_path


class _BSRNNCore(nn.Module):
    """Parameter tree of espnet2.enh.layers.bsrnn.BSRNN (attribute names per SURVEY.md Appendix A)."""

    def __init__(self, input_dim=481, num_channel=16, num_layer=6, target_fs=48000, num_spk=1):
        super().__init__()
        self.num_layer, self.num_channel, self.input_dim = num_layer, num_channel, input_dim
        self.band_split = BandSplitParams(input_dim, target_fs=target_fs, channels=num_channel)
        add_dual_path(self, num_channel, num_layer)
        self.mask_decoder = MaskDecoderParams(input_dim, self.band_split.subbands, channels=num_channel, num_spk=num_spk)


class _Separator(nn.Module):
    """espnet2 BSRNNSeparator holds its network under ``.bsrnn`` — hence the ``bsrnn.bsrnn.`` key prefix."""

    def __init__(self, input_dim, num_channels, num_layers, target_fs):
        super().__init__()
        self.bsrnn = _BSRNNCore(input_dim, num_channels, num_layers, target_fs)


class BSRNN_SE(nn.Module):
    N_FFT, HOP, DEFAULT_FS = 960, 480, 48000          # reference bsrnn.py:14-25
    MAX_GRAPHS = 16                                   # captured (shape, fs, lengths) signatures kept alive

    def __init__(self, num_channel=192, num_layer=6, precision=None, cuda_graph=None):
        super().__init__()
        # cuda_graph: replay the fixed-shape launch sequence from a captured CUDA graph (one graph per
        # (B, L, fs, lengths) signature, invalidated when parameters change).  Default from BSRNN_B200_GRAPH (off).
        self.cuda_graph = (os.environ.get("BSRNN_B200_GRAPH", "0") == "1") if cuda_graph is None else bool(cuda_graph)
        self._graphs = {}
        self.bsrnn = _Separator(self.N_FFT // 2 + 1, num_channel, num_layer, self.DEFAULT_FS)
        self.num_channel, self.num_layer = num_channel, num_layer
        # "fp32": CUDA-core kernels, parity bar 1e-3.  "fp16" (alias "bf16"): tcgen05 tensor cores with 16-bit
        # operands and f32 accumulation, parity bar 1e-2 (BASELINE.json north_star).
        self.precision = precision or os.environ.get("BSRNN_B200_PRECISION", "fp16")
        if self.precision == "bf16":
            self.precision = "fp16"
        core = self.bsrnn.bsrnn
        self._dual = R.PackedCache(core, R.pack_dual_path)
        self._bs = R.PackedCache(core.band_split, R.pack_band_split)
        self._md = R.PackedCache(core.mask_decoder, R.pack_mask_decoder)
        self._dual_tc = R.PackedCache(core, TC.pack_dual_path_tc)
        self._md_tc = R.PackedCache(core.mask_decoder, TC.pack_mask_decoder_tc)
        self._bs_tc = R.PackedCache(core.band_split, TC.pack_band_split_tc)
        self._dual_steps = None                          # step-wise tensor-core packs, built on first use (widths != 196)

    # ------------------------------------------------------------------------------------------------
    def _device(self):
        return self.bsrnn.bsrnn.fc_time[0].weight.device

    @torch.no_grad()
    def forward(self, speech_mix, speech_lengths, fs):
        """speech_mix (B,L) float, speech_lengths (B,) int, fs int or 0-dim tensor ->
        (enhanced_wav (B, max(lengths)) f32, enhanced_feature (B,T,F) complex64)   [reference bsrnn.py:36-41]"""
        L.require_device()
        dev = self._device()
        if dev.type != "cuda":
            raise L.NativeLibraryError("BSRNN_SE parameters must live on a CUDA device (no CPU fallback)")
        fs = int(fs)
        if speech_mix.dim() != 2:
            raise ValueError(f"speech_mix must be (B, L), got {tuple(speech_mix.shape)}")
        lens_host = speech_lengths.detach().to("cpu") if torch.is_tensor(speech_lengths) else torch.as_tensor(speech_lengths)
        L_out = int(lens_host.max())
        if not self.cuda_graph:
            wav = speech_mix.to(device=dev, dtype=torch.float32, non_blocking=True).contiguous()
            lens = lens_host.to(device=dev, dtype=torch.int32, non_blocking=True)
            return self._forward_device(wav, lens, L_out, fs)
        # ---- CUDA-graph path: one captured launch sequence per input signature and parameter version
        params = R.param_signature(self.parameters())
        key = (tuple(speech_mix.shape), fs, tuple(int(v) for v in lens_host.tolist()), self.precision)
        entry = self._graphs.get(key)
        if entry is None or entry[1] != params:
            # A replay may still be running (callers such as pipeline.StreamedEnhancer do not synchronise between
            # batches): destroying or replacing its graph would hand its private memory pool to the next capture, so
            # drain the device before touching the cache (a capture is a slow path anyway).
            torch.cuda.synchronize(dev)
            while len(self._graphs) >= self.MAX_GRAPHS:
                self._graphs.pop(next(iter(self._graphs)))   # graphs pin their workspaces: oldest signature goes first
            x_static = torch.empty(tuple(speech_mix.shape), dtype=torch.float32, device=dev)
            x_static.copy_(speech_mix)
            lens = lens_host.to(device=dev, dtype=torch.int32)
            g = R.GraphedForward(lambda x: self._forward_device(x, lens, L_out, fs), [x_static], keep=[lens])
            entry = self._graphs[key] = (g, params)
        wav_out, est = entry[0].run(speech_mix)
        return wav_out.clone(), est.clone()               # the graph's own output buffers are reused by the next replay

    def _forward_device(self, wav, lens, L_out, fs):
        """The launch sequence on device-resident inputs (everything stream-ordered on the current stream)."""
        n_fft, hop = R.stft_dims(fs, self.N_FFT, self.HOP, self.DEFAULT_FS)
        core = self.bsrnn.bsrnn
        plan = R.BandPlan.make(core.band_split.subbands, n_fft // 2 + 1)

        spec, bstats = R.stft(wav, lens, n_fft, hop, plan=plan)        # band statistics fused into the STFT epilogue
        if self.precision == "fp32":
            skip = R.band_split_f32(spec, plan, self._bs.get(), self.num_channel, stats=bstats)
            R.dual_path_f32(skip, self._dual.get())
            mask, resid = R.mask_decoder_f32(skip, plan, self._md.get())
        elif self.precision == "fp16":
            if 2 * self.num_channel != TC.CL * TC.LU:
                # other widths (e.g. the constructor default num_channel=192): the step-wise tensor-core BLSTM kernels of the
                # FlowSE / training paths serve any H = 2N with H % 16 == 0; the persistent fused layer kernel is H = 392 / 768 only
                if self.num_channel % 8:
                    raise NotImplementedError(
                        f"tensor-core mode needs num_channel % 8 == 0 (or 196, the published width); got {self.num_channel}. "
                        "Use precision='fp32' for other widths.")
                from . import runtime_tc_steps as TS
                if self._dual_steps is None:
                    self._dual_steps = R.PackedCache(core, TS.pack_dual_path_steps)
                skip = R.band_split_f32(spec, plan, self._bs.get(), self.num_channel, stats=bstats)
                TS.dual_path_tc_steps(skip, self._dual_steps.get())
                mask, resid = TC.mask_decoder_tc(skip, plan, self._md_tc.get())
                wav_out, est = R.istft(spec, mask, resid, L_out, n_fft, hop, want_spec=True)
                return wav_out, torch.view_as_complex(est)
            B, T = spec.shape[0], spec.shape[1]
            tc_bs = TC.BAND_SPLIT_TC and self.num_channel % 4 == 0
            if tc_bs:                                  # Conv1d(2 s_k -> N) on tcgen05, first GroupNorm's sums in its epilogue
                skip = TC.band_split_tc(spec, plan, self._bs.get(), self._bs_tc.get(), self.num_channel, bstats)
            else:
                skip = R.band_split_f32(spec, plan, self._bs.get(), self.num_channel, stats=bstats)
            dec_stats = torch.zeros(B, plan.K, 2, dtype=torch.float64, device=spec.device)
            have = TC.dual_path_tc(skip, self._dual_tc.get(), stats_ready=tc_bs, band_stats=dec_stats)
            mask, resid = TC.mask_decoder_tc(skip, plan, self._md_tc.get(), band_stats=dec_stats if have else None)
        else:
            raise NotImplementedError(f"precision {self.precision!r}")
        wav_out, est = R.istft(spec, mask, resid, L_out, n_fft, hop, want_spec=True)
        return wav_out, torch.view_as_complex(est)
