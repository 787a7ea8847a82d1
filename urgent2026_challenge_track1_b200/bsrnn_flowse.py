"""Flow-matching BSRNN backbone — drop-in for ``baseline_code/models/bsrnn_flowse.py::BSRNN`` (reference
bsrnn_flowse.py:171-318) with the same constructor, ``forward(dnn_input, t, fs=None)`` signature, ``current_fs``
attribute and state_dict keys (band_split_x/y, condition_fc, norm/rnn/fc_{time,freq}, t_cond.{i}.W, grad_decoder.*).

All arithmetic runs in libbsrnn_b200 kernels.  ``precision = "fp16"`` (default) runs the dual path on fp16 tensor-core
GEMMs with the step-wise tensor-core BLSTM of runtime_tc_steps (any H % 16 == 0, so H = 768 too); ``"fp32"`` keeps
everything on the CUDA-core kernels.  Besides
the reference API the class exposes ``mask_resid(x, y_embed, t)`` working on the (B,T,F,2) layout so the sampler can
fuse the Euler update with the network output and hoist the loop-invariant ``band_split_y``.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from . import _lib as L
from . import runtime as R
from . import runtime_tc_steps as TS
from .layers import BandSplitParams, GradDecoderParams, add_dual_path


class GaussianFourierProjection(nn.Module):
    """Parameter container of the Gaussian Fourier time embedding (reference bsrnn_flowse.py:90-99);
    W is in the state_dict with requires_grad=False, exactly as in the reference."""

    def __init__(self, embedding_size=256, scale=1.0):
        super().__init__()
        self.W = nn.Parameter(torch.randn(embedding_size) * scale, requires_grad=False)

    def forward(self, x):
        out = torch.empty(x.shape[0], 2 * self.W.numel(), dtype=torch.float32, device=self.W.device)
        t = x.to(device=self.W.device, dtype=torch.float32).contiguous()
        L.call("bsrnn_time_embed", t.data_ptr(), self.W.data_ptr(), out.data_ptr(), t.shape[0], self.W.numel(), L.stream_ptr())
        return out


BandSplit = BandSplitParams          # reference names kept importable
GradDecoder = GradDecoderParams


class BSRNN(nn.Module):
    def __init__(self, input_dim=481, num_channel=16, num_layer=6, target_fs=48000, causal=True, num_spk=1,
                 norm_type="GN"):
        super().__init__()
        if causal:
            raise NotImplementedError("the B200 path implements the non-causal (BLSTM) configuration the reference "
                                      "instantiates (flow_model.py:44-49)")
        if norm_type != "GN":
            raise NotImplementedError("only norm_type='GN' (the reference default) is implemented")
        self.num_layer, self.num_channel, self.input_dim = num_layer, num_channel, input_dim
        self.band_split_y = BandSplitParams(input_dim, target_fs=target_fs, channels=num_channel)
        self.band_split_x = BandSplitParams(input_dim, target_fs=target_fs, channels=num_channel)
        self.condition_fc = nn.Linear(2 * num_channel, num_channel)
        self.target_fs, self.causal, self.num_spk = target_fs, causal, num_spk
        add_dual_path(self, num_channel, num_layer, with_t_cond=True, t_cond_cls=GaussianFourierProjection)
        self.grad_decoder = GradDecoderParams(input_dim, self.band_split_x.subbands, channels=num_channel, num_spk=1)
        self.current_fs = None
        self._dual = R.PackedCache(self, R.pack_dual_path)
        # precision "fp16" (default when H = 2*num_channel is a multiple of 16): fp16 tensor-core GEMMs + the step-wise
        # tensor-core BLSTM (runtime_tc_steps) for the dual path and condition_fc -- 99 % of the FLOPs; band split and
        # GradDecoder stay f32.  Against the f32 path at the full width (N=384, L=6): vector field 4.2e-4, enhanced
        # waveform 1.3e-4 relative L2 (tools/flowse_fp16_vs_f32.py), i.e. inside the f32 bar of 1e-3 as well.
        # "fp32": CUDA-core kernels everywhere.  BSRNN_FLOWSE_PRECISION overrides the default.
        self.precision = os.environ.get("BSRNN_FLOWSE_PRECISION", "fp16" if (2 * num_channel) % 16 == 0 else "fp32")
        self._dual_steps = R.PackedCache(self, TS.pack_dual_path_steps)
        # cuda_graph: replay one network evaluation (tens of thousands of launches in the step-wise mode) from a captured
        # CUDA graph, one graph per input shape, rebuilt when parameters change.  Default from BSRNN_B200_GRAPH (off).
        self.cuda_graph = os.environ.get("BSRNN_B200_GRAPH", "0") == "1"
        self._graphs = {}
        self._zz = {}
        self._cond_tc = R.PackedCache(self.condition_fc, TS.pack_linear_tc)
        self._cond_ws = None
        self._bsx = R.PackedCache(self.band_split_x, R.pack_band_split)
        self._bsy = R.PackedCache(self.band_split_y, R.pack_band_split)
        self._gd = R.PackedCache(self.grad_decoder, R.pack_grad_decoder)
        self._gd_tc = R.PackedCache(self.grad_decoder, TS.pack_grad_decoder_tc)

    # ------------------------------------------------------------------------------------------------ (B,T,F,2) core
    def embed_y(self, y_btf):
        """band_split_y(y) -> (B,T,K',2N) buffer with columns [N,2N) filled; loop-invariant across ODE steps."""
        B, T, F, _ = y_btf.shape
        plan = R.BandPlan.make(self.band_split_x.subbands, F)
        N = self.num_channel
        zkey = (B, T, plan.K, 2 * N, str(y_btf.device))
        zz = self._zz.get(zkey)                 # persistent per shape: a captured graph keeps pointing at it
        if zz is None:
            self._zz.clear()
            zz = self._zz[zkey] = torch.empty(B, T, plan.K, 2 * N, dtype=torch.float32, device=y_btf.device)
        R.band_split_f32(y_btf, plan, self._bsy.get(), N, out=zz, out_col=N, out_width=2 * N)
        return zz, plan

    @torch.no_grad()
    def mask_resid(self, x_btf, zz, plan, t):
        """One network evaluation on the (B,T,F,2) layout -> (mask, resid) with g = mask*x + resid.  With cuda_graph the
        returned tensors are the graph's own outputs: valid until the next call (the Euler update consumes them at once)."""
        if not self.cuda_graph:
            return self._mask_resid_eager(x_btf, zz, plan, t)
        dev = x_btf.device
        t = t.to(device=dev, dtype=torch.float32)
        params = R.param_signature(self.parameters())
        key = (tuple(x_btf.shape), zz.data_ptr(), self.precision)
        entry = self._graphs.get(key)
        if entry is None or entry[1] != params:
            torch.cuda.synchronize(dev)         # a replay of the graph being replaced may still be running
            self._graphs.clear()
            xs, ts = x_btf.clone(), t.clone()
            g = R.GraphedForward(lambda a, b: self._mask_resid_eager(a, zz, plan, b), [xs, ts], keep=[zz])
            entry = self._graphs[key] = (g, params)
        return entry[0].run(x_btf, t)

    @torch.no_grad()
    def _mask_resid_eager(self, x_btf, zz, plan, t):
        B, T, F, _ = x_btf.shape
        N = self.num_channel
        dev = x_btf.device
        st = L.stream_ptr()
        R.band_split_f32(x_btf, plan, self._bsx.get(), N, out=zz, out_col=0, out_width=2 * N)
        skip = torch.empty(B, T, plan.K, N, dtype=torch.float32, device=dev)
        M = B * T * plan.K
        if self.precision in ("fp16", "bf16") and (2 * N) % 8 == 0 and N % 4 == 0:
            # condition_fc Linear(2N -> N) [bsrnn_flowse.py:284-285] on tensor cores: zz -> fp16 KB8 operand tiles (token
            # order), then the residual-epilogue GEMM into a zeroed skip (1.1 TFLOP per evaluation at config 4: 45 ms on
            # the f32 CUDA-core GEMM)
            cp = self._cond_tc.get()
            m_tiles = (M + 127) // 128
            ckey = (M, 2 * N, str(dev))
            if self._cond_ws is None or self._cond_ws[0] != ckey:
                self._cond_ws = (ckey, torch.empty(m_tiles * (2 * N // 8) * 1024, dtype=torch.float16, device=dev))
            xz = self._cond_ws[1]
            L.call("bsrnn_norm_cast_kb8", zz.data_ptr(), None, None, xz.data_ptr(), 2 * N, 0, 2 * N, 2 * N // 8, m_tiles,
                   m_tiles, M, 1 << 40, 0, 1, 0, M, 1, st)
            skip.zero_()
            L.call("bsrnn_gemm_tc", xz.data_ptr(), cp["w"].data_ptr(), cp["b"].data_ptr(), skip.data_ptr(), None, m_tiles,
                   cp["nt"], 2 * N // 8, cp["bn"], TS.FC_EPI, N, N, 0, M, m_tiles, M, 1 << 40, 0, 1, 0, st)
        else:
            dl = R.DescList()
            w, b = self.condition_fc.weight, self.condition_fc.bias               # bsrnn_flowse.py:284-285
            dl.add(**R._rows_desc(zz.data_ptr(), w.data_ptr(), b.data_ptr(), skip.data_ptr(), M, N, 2 * N,
                                  a_stride=2 * N, c_stride=N))
            dl.upload(dev)
            L.call("bsrnn_gemm_f32", dl.ptr(0), 1, M, N, st)
        t = t.to(device=dev, dtype=torch.float32)
        t_emb = [self.t_cond[i](t) for i in range(self.num_layer)]          # bsrnn_flowse.py:293
        if self.precision in ("fp16", "bf16"):
            TS.dual_path_tc_steps(skip, self._dual_steps.get(), t_emb=t_emb)
        else:
            R.dual_path_f32(skip, self._dual.get(), t_emb=t_emb)
        if self.precision in ("fp16", "bf16") and N % 16 == 0:
            return TS.grad_decoder_tc(skip, plan, self._gd_tc.get(), self.grad_decoder.sub_channel)
        return R.grad_decoder_f32(skip, plan, self._gd.get(), self.grad_decoder.sub_channel)

    # ------------------------------------------------------------------------------------------------ reference API
    @torch.no_grad()
    def forward(self, dnn_input, t=None, fs=None):
        """dnn_input complex (B,2,F,T), t (B,) -> g complex (B,1,F,T)   [reference bsrnn_flowse.py:255-318]."""
        L.require_device()
        assert t is not None                                                  # bsrnn_flowse.py:280
        dev = self.condition_fc.weight.device
        d = dnn_input.to(dev)
        x = torch.view_as_real(d[:, 0].permute(0, 2, 1).contiguous()).contiguous()
        y = torch.view_as_real(d[:, 1].permute(0, 2, 1).contiguous()).contiguous()
        zz, plan = self.embed_y(y)
        m, r = self.mask_resid(x, zz, plan, t)
        g = torch.empty_like(x)
        L.call("bsrnn_complex_mask", g.data_ptr(), x.data_ptr(), m.data_ptr(), r.data_ptr(), 1.0, x.numel() // 2, L.stream_ptr())
        return torch.view_as_complex(g).permute(0, 2, 1).unsqueeze(1)
