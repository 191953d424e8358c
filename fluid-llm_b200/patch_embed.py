"""Patch-embedding projection on the B200 tensor cores (tcgen05), inference path.

Mirror of `/root/reference/src/models/layers/input_embeddings.py:36-52` (`InputEmbeddings.forward`) with the
configuration the reference ships (`configs/training1.yaml:40-51`: MLP encoder 768 -> 512 -> llm_dim with
LeakyReLU, learned x/y/t positional embeddings, no LayerNorm) as it runs under bf16 autocast
(`src/utils.py:53-62`).  Dropout is a training-time op and is not applied (eval mode).
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, load, ptr, stream_ptr


class PatchEmbedder:
    """Holds bf16 copies of the MLP weights and the fp32 positional tables on the device."""

    def __init__(self, w1, b1, w2, b2, x_emb=None, y_emb=None, t_emb=None, device=None):
        _lib.require_cuda()
        dev = torch.device(device or "cuda")
        self.device = dev
        self.w1 = w1.detach().to(dev, torch.bfloat16).contiguous()          # [hid, in]   (nn.Linear layout)
        self.w2 = w2.detach().to(dev, torch.bfloat16).contiguous()          # [out, hid]
        # autocast casts the bias to bf16 before the GEMM epilogue adds it
        self.b1 = b1.detach().to(dev, torch.bfloat16).float().contiguous()
        self.b2 = b2.detach().to(dev, torch.bfloat16).float().contiguous()
        self.hid_dim, self.in_dim = self.w1.shape
        self.out_dim = self.w2.shape[0]
        if self.w2.shape[1] != self.hid_dim or self.b1.numel() != self.hid_dim or self.b2.numel() != self.out_dim:
            raise ValueError("PatchEmbedder: weight shapes do not chain (in -> hid -> out)")
        self.pos = None
        if x_emb is not None:
            tabs = [t.detach().to(dev, torch.float32).contiguous() for t in (x_emb, y_emb, t_emb)]
            if any(t.dim() != 2 or t.shape[1] != self.out_dim for t in tabs):
                raise ValueError("PatchEmbedder: positional tables must be [max, llm_dim]")
            self.pos = tabs

    @classmethod
    def from_module(cls, input_embeddings, device=None):
        """Build from a reference `InputEmbeddings` module (MLP encoder with two layers, 'pos' embeddings)."""
        layers = input_embeddings.patch_embeddings.encoder.layers
        if len(layers) != 2:
            raise ValueError("only the two-layer MLP encoder of configs/training1.yaml is supported")
        pe = input_embeddings.position_embeddings
        return cls(layers[0].weight, layers[0].bias, layers[1].weight, layers[1].bias,
                   pe.x_embeddings.weight, pe.y_embeddings.weight, pe.time_embeddings.weight, device)

    def __call__(self, x, position_ids=None):
        """x (..., N_patch, 3, H, W) fp32 CUDA, position_ids (..., N_patch, 3) int64 -> (..., N_patch, llm_dim) fp32."""
        if not x.is_cuda:
            raise _lib.FluidGridError("PatchEmbedder: x must be a CUDA tensor")
        lead = x.shape[:-1] if x.shape[-1] == self.in_dim else x.shape[:-3]     # already-flattened patches are accepted too
        xf = x.reshape(-1, self.in_dim)                 # patch_encoder.py:25: flatten (C, H, W) -> c*H*W + i*W + j
        if xf.shape[1] != self.in_dim:
            raise ValueError(f"PatchEmbedder: patches flatten to {xf.shape[1]} values, weights expect {self.in_dim}")
        n = xf.shape[0]
        lib = load()
        with torch.cuda.device(x.device):
            if xf.dtype == torch.bfloat16:
                xb = xf.contiguous()
            else:
                xf = xf.float().contiguous()
                xb = torch.empty((n, self.in_dim), dtype=torch.bfloat16, device=x.device)
                check(lib.fl_cast_bf16(ptr(xf), ptr(xb), n * self.in_dim, stream_ptr()), "fl_cast_bf16")
            hidden = torch.empty((n, self.hid_dim), dtype=torch.bfloat16, device=x.device)
            out = torch.empty((n, self.out_dim), dtype=torch.float32, device=x.device)
            ids = None
            if position_ids is not None:
                if self.pos is None:
                    raise ValueError("PatchEmbedder: position_ids given but no positional tables")
                ids = position_ids.reshape(-1, 3).to(torch.int64).contiguous()
                if ids.shape[0] != n:
                    raise ValueError("PatchEmbedder: position_ids do not match the number of patches")
            xe, ye, te = self.pos if self.pos is not None else (None, None, None)
            check(lib.fl_patch_embed(ptr(xb), ptr(self.w1), ptr(self.b1), ptr(self.w2), ptr(self.b2), ptr(xe), ptr(ye), ptr(te),
                                     ptr(ids), xe.shape[0] if xe is not None else 0, ye.shape[0] if ye is not None else 0,
                                     te.shape[0] if te is not None else 0, ptr(hidden), ptr(out), n, self.in_dim, self.hid_dim,
                                     self.out_dim, stream_ptr()), "fl_patch_embed")
        return out.view(*lead, self.out_dim)


    def graphed(self, n_tokens: int, with_position_ids: bool = True, in_dtype=torch.float32):
        """Fixed-shape variant captured in a CUDA graph (cast + two GEMM launches replayed as one submission): for the
        rollout loop, where the same number of tokens is embedded every step and launch gaps are a third of the time."""
        return GraphedPatchEmbed(self, int(n_tokens), with_position_ids, in_dtype)


class GraphedPatchEmbed:
    """`x` (n_tokens, in_dim) and `ids` (n_tokens, 3) are static input buffers, `out` (n_tokens, llm_dim) the static output:
    fill the inputs (or let the producing kernel write into them) and call `replay()`; `__call__(x, ids)` copies first."""

    def __init__(self, emb: PatchEmbedder, n_tokens, with_position_ids, in_dtype):
        if with_position_ids and emb.pos is None:
            raise ValueError("PatchEmbedder: position ids requested but no positional tables")
        dev = emb.device if emb.device.index is not None else torch.device("cuda", torch.cuda.current_device())
        self.emb = emb
        with torch.cuda.device(dev):
            self.x = torch.zeros((n_tokens, emb.in_dim), dtype=in_dtype, device=dev)
            self.ids = torch.zeros((n_tokens, 3), dtype=torch.int64, device=dev) if with_position_ids else None
            cur = torch.cuda.current_stream(dev)
            side = torch.cuda.Stream(dev)
            side.wait_stream(cur)
            with torch.cuda.stream(side):                 # warm-up outside the capture (function attributes, allocator)
                emb(self.x, self.ids)
            cur.wait_stream(side)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.out = emb(self.x, self.ids)

    def replay(self):
        self.graph.replay()
        return self.out

    def __call__(self, x, position_ids=None):
        self.x.copy_(x.reshape(self.x.shape))
        if self.ids is not None:
            if position_ids is None:
                raise ValueError("GraphedPatchEmbed: captured with position ids")
            self.ids.copy_(position_ids.reshape(-1, 3))
        return self.replay()
