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
        # the GEMM steps through K in blocks of 64: a token length that is not a multiple of it (patches other than 16 x 16,
        # e.g. 3 * 10 * 7 = 210) is padded with zeros, weights and tokens alike -- the products are exact zeros
        self.k_dim = (self.in_dim + 63) // 64 * 64
        if self.k_dim != self.in_dim:
            self.w1 = torch.nn.functional.pad(self.w1, (0, self.k_dim - self.in_dim)).contiguous()
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

    def check_ids(self, position_ids):
        """nn.Embedding raises IndexError for an id outside its table (input_embeddings.py:18-26 looks the three ids up in
        nn.Embedding modules); the kernels clamp instead, so the eager call validates first (one device -> host read)."""
        if self.pos is None:
            raise ValueError("PatchEmbedder: position_ids given but no positional tables")
        ids = position_ids.reshape(-1, 3)
        lo, hi = ids.amin(dim=0).tolist(), ids.amax(dim=0).tolist()
        for k, name in enumerate(("x", "y", "time")):
            if lo[k] < 0 or hi[k] >= self.pos[k].shape[0]:
                raise IndexError(f"PatchEmbedder: {name} position id out of range [0, {self.pos[k].shape[0]}) (found {lo[k]}..{hi[k]})")

    def __call__(self, x, position_ids=None, validate_ids=True, out=None):
        """x (..., N_patch, 3, H, W) fp32 CUDA, position_ids (..., N_patch, 3) int64 -> (..., N_patch, llm_dim) fp32.
        validate_ids=False skips the range check of the ids (and its host synchronisation): out-of-range ids are then clamped.
        `out`: a contiguous fp32 (n_tokens, llm_dim) CUDA tensor to write into instead of a fresh one."""
        if not x.is_cuda:
            raise _lib.FluidGridError("PatchEmbedder: x must be a CUDA tensor")
        lead = x.shape[:-1] if x.shape[-1] == self.in_dim else x.shape[:-3]     # already-flattened patches are accepted too
        xf = x.reshape(-1, self.in_dim)                 # patch_encoder.py:25: flatten (C, H, W) -> c*H*W + i*W + j
        if xf.shape[1] != self.in_dim:
            raise ValueError(f"PatchEmbedder: patches flatten to {xf.shape[1]} values, weights expect {self.in_dim}")
        n = xf.shape[0]
        lib = load()
        with torch.cuda.device(x.device):
            if xf.dtype == torch.bfloat16 and self.k_dim == self.in_dim:
                xb = xf.contiguous()
            elif self.k_dim == self.in_dim and (n * self.in_dim) % 4 == 0:
                xf = xf.float().contiguous()
                xb = torch.empty((n, self.in_dim), dtype=torch.bfloat16, device=x.device)
                check(lib.fl_cast_bf16(ptr(xf), ptr(xb), n * self.in_dim, stream_ptr()), "fl_cast_bf16")
            else:                                       # zero-padded rows (bf16 -> fp32 -> bf16 is exact)
                xf = xf.float().contiguous()
                xb = torch.empty((n, self.k_dim), dtype=torch.bfloat16, device=x.device)
                check(lib.fl_cast_bf16_rows(ptr(xf), ptr(xb), n, self.in_dim, self.k_dim, stream_ptr()), "fl_cast_bf16_rows")
            hidden = torch.empty((n, self.hid_dim), dtype=torch.bfloat16, device=x.device)
            if out is None:
                out = torch.empty((n, self.out_dim), dtype=torch.float32, device=x.device)
            elif out.shape != (n, self.out_dim) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != x.device:
                raise ValueError("PatchEmbedder: out must be a contiguous float32 (n_tokens, llm_dim) tensor on x's device")
            ids = None
            if position_ids is not None:
                if self.pos is None:
                    raise ValueError("PatchEmbedder: position_ids given but no positional tables")
                ids = position_ids.reshape(-1, 3).to(torch.int64).contiguous()
                if ids.shape[0] != n:
                    raise ValueError("PatchEmbedder: position_ids do not match the number of patches")
                if validate_ids and not torch.cuda.is_current_stream_capturing():
                    self.check_ids(ids)
            xe, ye, te = self.pos if self.pos is not None else (None, None, None)
            check(lib.fl_patch_embed(ptr(xb), ptr(self.w1), ptr(self.b1), ptr(self.w2), ptr(self.b2), ptr(xe), ptr(ye), ptr(te),
                                     ptr(ids), xe.shape[0] if xe is not None else 0, ye.shape[0] if ye is not None else 0,
                                     te.shape[0] if te is not None else 0, ptr(hidden), ptr(out), n, self.k_dim, self.hid_dim,
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
                emb(self.x, self.ids, validate_ids=False)
            cur.wait_stream(side)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.out = emb(self.x, self.ids, validate_ids=False)

    def replay(self):
        self.graph.replay()
        return self.out

    def __call__(self, x, position_ids=None, validate_ids=False):
        self.x.copy_(x.reshape(self.x.shape))
        if self.ids is not None:
            if position_ids is None:
                raise ValueError("GraphedPatchEmbed: captured with position ids")
            if validate_ids:
                self.emb.check_ids(position_ids)
            self.ids.copy_(position_ids.reshape(-1, 3))
        return self.replay()


class RolloutEmbedCache:
    """The embedding side of the autoregressive rollout without re-embedding the whole context every step.

    `src/models/model.py:187-204` feeds the last `max_ctx_len` states through `InputEmbeddings` at every predicted step, with
    time ids re-based so that the oldest state of the context is t = 0 (`:196-199`).  Of those states only ONE is new; the
    other rows of bf16(h W2^T + b2) are what they were a step ago and only the positional term moves.  The cache keeps the
    pre-positional embedding of every state of the context in a ring [ctx, B, L, llm_dim]; a step embeds the new state alone
    (bf16 tokens straight from `rollout_step(..., tokens_bf16=True)`: no cast kernel) and re-applies the positional add over
    the ring (`fl_pos_add_ring`).  Results are bit-identical to embedding the concatenated context (tests/test_gpu_embed.py)."""

    def __init__(self, emb: PatchEmbedder, batch: int, n_patch: int, ctx: int, graphs: bool = False):
        """graphs=True: once the ring is full, `step(tokens, ids)` replays one CUDA graph per ring position (the two GEMMs of
        the new state + the positional add: one submission instead of three launches and four tensor-map encodes)."""
        if emb.pos is None:
            raise ValueError("RolloutEmbedCache needs the positional tables")
        self.emb, self.B, self.L, self.ctx = emb, int(batch), int(n_patch), int(ctx)
        dev = emb.device if emb.device.index is not None else torch.device("cuda", torch.cuda.current_device())
        self.ring = torch.zeros((self.ctx, self.B, self.L, emb.out_dim), dtype=torch.float32, device=dev)
        self.n = 0            # states in the ring
        self.head = 0         # slot the next state goes to
        self.graphs = {} if graphs else None
        if graphs:
            self.g_tok = torch.zeros((self.B * self.L, emb.in_dim), dtype=torch.bfloat16, device=dev)
            self.g_ids = torch.zeros((self.B, self.ctx, self.L, 3), dtype=torch.int64, device=dev)
            self.g_out = None

    @property
    def token_buffer(self):
        """graphs=True: the static (B * L, in_dim) bf16 input of the graphs -- let `rollout_step(..., tokens_out=...)` write
        the new state's tokens straight into it and call `step(None, ...)`."""
        return self.g_tok

    def step(self, state, position_ids):
        """append(state) + tokens(position_ids) in one call; with graphs=True and a full ring, one graph replay.
        `state=None`: the tokens are already in `token_buffer`; `position_ids=None`: the ids of the previous call (in steady
        state they are the same every step: time ids re-based to 0 .. ctx-1).  The returned tensor is reused by the next replay."""
        if state is None:
            state = self.g_tok
        if self.graphs is None or self.n < self.ctx - 1 or state.dtype != torch.bfloat16:
            if position_ids is None:
                raise ValueError("RolloutEmbedCache.step: position_ids=None needs the graphed steady state")
            self.append(state)
            return self.tokens(position_ids)
        if state is not self.g_tok:
            self.g_tok.copy_(state.reshape(self.g_tok.shape))
        if position_ids is not None:
            self.g_ids.copy_(position_ids.reshape(self.g_ids.shape))
        g = self.graphs.get(self.head)
        if g is None:
            dev = self.ring.device
            cur = torch.cuda.current_stream(dev)
            side = torch.cuda.Stream(dev)
            side.wait_stream(cur)
            n0, h0 = self.n, self.head
            with torch.cuda.stream(side):                 # warm-up outside the capture
                self.append(self.g_tok)
                self.tokens(self.g_ids)
            cur.wait_stream(side)
            self.n, self.head = n0, h0
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.append(self.g_tok)
                out = self.tokens(self.g_ids, out=self.g_out)
            if self.g_out is None:
                self.g_out = out
            self.n, self.head = n0, h0
            self.graphs[h0] = g
        g.replay()
        self.head = (self.head + 1) % self.ctx
        self.n = min(self.n + 1, self.ctx)
        return self.g_out.view(self.B, self.ctx * self.L, self.emb.out_dim)

    def append(self, state):
        """state (B, 1, L, 3, px, py) or (B, L, in_dim), fp32 or bf16: embed it (no positional term) into the ring."""
        x = state.reshape(self.B * self.L, self.emb.in_dim)
        self.emb(x, None, out=self.ring[self.head].view(self.B * self.L, self.emb.out_dim))
        self.head = (self.head + 1) % self.ctx
        self.n = min(self.n + 1, self.ctx)

    def tokens(self, position_ids, out=None):
        """position_ids (B, c, L, 3) int64 with c = states in the ring -> (B, c * L, llm_dim) fp32, oldest state first."""
        c = self.n
        ids = position_ids.reshape(self.B, c, self.L, 3).to(torch.int64).contiguous()
        if out is None:
            out = torch.empty((self.B, c, self.L, self.emb.out_dim), dtype=torch.float32, device=self.ring.device)
        xe, ye, te = self.emb.pos
        start = (self.head - c) % self.ctx
        with torch.cuda.device(self.ring.device):
            check(load().fl_pos_add_ring(ptr(self.ring), ptr(xe), ptr(ye), ptr(te), ptr(ids), xe.shape[0], ye.shape[0], te.shape[0],
                                         ptr(out), self.B, c, self.L, self.ctx, start, self.emb.out_dim, stream_ptr()),
                  "fl_pos_add_ring")
        return out.view(self.B, c * self.L, self.emb.out_dim)
