"""Mirror of the reference's axial_model.py: `AxialAttention` (axial_model.py:19-40), the Decoder variant
that applies it to the token embeddings in forward() (axial_model.py:56-158) and the two-argument
`EncoderDecoder` (axial_model.py:161-174).  All arithmetic runs in libmdc_b200.so.

Reference quirks kept on purpose: the softmax scale is 64**-0.5 = 0.125 regardless of the
real head width, the attention is un-masked, `to_qkv` has no bias, forward() does NOT prepend BOS,
and predict() does not use the axial attention at all.
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn

from . import _lib as L
from .config import CFG
from .model import Decoder as _BaseDecoder
from .model import Encoder, EncoderDecoder as _BaseEncoderDecoder, Engine, _EngineOwner, _Holder  # noqa: F401


class AxialAttention(nn.Module, _EngineOwner):
    def __init__(self, dim, heads=8, dim_head=64):
        super().__init__()
        if heads != 8 or dim_head != 64:
            raise NotImplementedError("the B200 kernel path implements the reference's only configuration "
                                      "(heads=8, dim_head=64 -> scale 0.125)")
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.dim = dim
        self.to_qkv = nn.Linear(dim, dim * 3, bias=False)
        self.to_out = nn.Linear(dim, dim)
        self.precision = None

    def forward(self, x, axis=-1):
        """x (b, n, dim) -> (b, n, dim); axis in {-1, 3} (keys, default) or {-2, 2} (queries)."""
        if axis in (-1, 3):
            soq = 0
        elif axis in (-2, 2):
            soq = 1
        else:
            raise ValueError("softmax axis must be one of the two attention-matrix axes (-1 or -2)")
        holder = _AxialOnly(self)
        eng = self._engine_for(None, holder, x.device if x.is_cuda else None)
        x = x.to(eng.device, torch.float32).contiguous()
        B, n, dim = x.shape
        out = torch.empty_like(x)
        wsb = eng.lib.mdc_axial_workspace_bytes(eng.handle, B, n)
        ws = torch.empty(wsb, dtype=torch.uint8, device=eng.device)
        with torch.cuda.device(eng.device):
            L.check(eng.lib.mdc_axial_attention(eng.handle, L.ptr(x), B, n, soq, L.ptr(out), L.ptr(ws), wsb, L.stream_ptr(eng.device)))
        return out


class _AxialOnly:
    """Adapter presenting a lone AxialAttention as a zero-layer decoder to the Engine."""

    class _NoLayers:
        layers = []

    def __init__(self, ax):
        self.axial_attention = ax
        self.dim, self.vocab_size, self.encoder_length, self.num_heads = ax.dim, 1, 1, ax.heads
        dev = ax.to_qkv.weight.device
        self.embedding = nn.Embedding(1, ax.dim).to(dev)
        self.decoder_pos_embed = torch.zeros(1, 1, ax.dim, device=dev)
        self.encoder_pos_embed = torch.zeros(1, 1, ax.dim, device=dev)
        self.output = nn.Linear(ax.dim, 1).to(dev)
        self.decoder = self._NoLayers()
        self._ax = ax

    def parameters(self):
        return self._ax.parameters()


class Decoder(_BaseDecoder):
    """axial_model.py:56-158."""

    def _extra_init(self):
        self.axial_attention = AxialAttention(self.dim)

    def _forward_with(self, eng, memory, tgt):
        # axial_model.py:88-118: no BOS prepend; axial attention on raw embeddings, THEN + pos
        B, n = tgt.shape
        tokens = tgt.to(eng.device, torch.int32).contiguous()
        pos = eng.interp_pos(self.decoder_pos_embed, n)
        x = torch.empty((B, n, self.dim), dtype=torch.float32, device=eng.device)
        wsb = eng.lib.mdc_axial_embed_workspace_bytes(eng.handle, B, n)
        ws = torch.empty(wsb, dtype=torch.uint8, device=eng.device)
        with torch.cuda.device(eng.device):
            L.check(eng.lib.mdc_axial_embed(eng.handle, L.ptr(tokens), n, B, n, L.ptr(pos), 0, L.ptr(x), L.ptr(ws), wsb, L.stream_ptr(eng.device)))
        return self._run_forced(eng, memory, tokens, n, n, 0, x_override=x)


class EncoderDecoder(_BaseEncoderDecoder):
    def __init__(self, encoder, decoder):
        super().__init__(encoder, decoder, patch_dropout_rate=0.0)
