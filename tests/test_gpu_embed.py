"""-m gpu: the tcgen05 patch-embedding GEMMs against PyTorch's own bf16-autocast evaluation of the reference
layers (src/models/layers/patch_encoder.py, MLP.py, input_embeddings.py, positional_embeddings.py)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _reference(x, w1, b1, w2, b2, tabs, ids):
    with torch.autocast("cuda", dtype=torch.bfloat16):
        h = F.leaky_relu(F.linear(x, w1, b1), 0.01)          # MLP.py:48-54, nn.LeakyReLU() default slope
        y = F.linear(h, w2, b2)
    if tabs is None:
        return y.float()
    xe, ye, te = tabs
    return y + (xe[ids[..., 0]] + ye[ids[..., 1]] + te[ids[..., 2]])   # positional_embeddings.py:32-37


@pytest.mark.parametrize("n_tokens", [128, 4800, 1000, 20000])    # 20000: more tiles than SMs -> the 2-stage, 2-CTA/SM variant
def test_patch_embed_matches_autocast_reference(n_tokens):
    from fluid_llm_b200.patch_embed import PatchEmbedder
    g = torch.Generator(device="cuda").manual_seed(n_tokens)
    x = torch.randn(n_tokens, 3, 16, 16, device="cuda", generator=g)
    w1 = torch.randn(512, 768, device="cuda", generator=g) * 768 ** -0.5
    b1 = torch.randn(512, device="cuda", generator=g) * 0.1
    w2 = torch.randn(768, 512, device="cuda", generator=g) * 512 ** -0.5
    b2 = torch.randn(768, device="cuda", generator=g) * 0.1
    tabs = [torch.randn(m, 768, device="cuda", generator=g) * 768 ** -0.5 for m in (20, 10, 30)]
    ids = torch.stack([torch.randint(0, m, (n_tokens,), device="cuda", generator=g) for m in (20, 10, 30)], dim=1)
    emb = PatchEmbedder(w1, b1, w2, b2, *tabs)
    out = emb(x, ids)
    ref = _reference(x.reshape(n_tokens, -1), w1, b1, w2, b2, tabs, ids)
    assert out.shape == (n_tokens, 768) and out.dtype == torch.float32
    # both sides round h and y to bf16 (8 mantissa bits); fp32 accumulation order differs -> at most a few bf16 ulps
    torch.testing.assert_close(out, ref, rtol=2e-2, atol=2e-2)
    assert float((out - ref).abs().mean()) < 2e-3
    # exact-arithmetic check of the same bf16-rounded operands (catches layout / swizzle bugs a loose tolerance hides)
    xb, w1b, w2b = x.reshape(n_tokens, -1).bfloat16().double(), w1.bfloat16().double(), w2.bfloat16().double()
    h = (xb @ w1b.T + b1.bfloat16().double()).float().bfloat16().float()
    h = torch.where(h > 0, h, 0.01 * h).bfloat16().double()
    y = (h @ w2b.T + b2.bfloat16().double()).float().bfloat16().float()
    exact = y + (tabs[0][ids[:, 0]] + tabs[1][ids[:, 1]] + tabs[2][ids[:, 2]])
    assert float((out - exact).abs().max()) < 6e-2 and float(((out - exact).abs() > 1e-6).float().mean()) < 0.05


@pytest.mark.parametrize("patch", [(10, 7), (5, 5), (8, 8), (32, 32)])
def test_patch_embed_for_other_patch_sizes(patch):
    """Token lengths 3 * px * py that are not a multiple of the GEMM's K step (210, 75) are zero-padded, weights and tokens alike."""
    from fluid_llm_b200.patch_embed import PatchEmbedder
    n, k = 777, 3 * patch[0] * patch[1]
    g = torch.Generator(device="cuda").manual_seed(k)
    x = torch.randn(n, 3, *patch, device="cuda", generator=g)
    w1 = torch.randn(256, k, device="cuda", generator=g) * k ** -0.5
    b1 = torch.randn(256, device="cuda", generator=g) * 0.1
    w2 = torch.randn(512, 256, device="cuda", generator=g) * 256 ** -0.5
    b2 = torch.randn(512, device="cuda", generator=g) * 0.1
    tabs = [torch.randn(m, 512, device="cuda", generator=g) * 0.05 for m in (9, 7, 5)]
    ids = torch.stack([torch.randint(0, m, (n,), device="cuda", generator=g) for m in (9, 7, 5)], dim=1)
    emb = PatchEmbedder(w1, b1, w2, b2, *tabs)
    assert emb.in_dim == k and emb.k_dim % 64 == 0 and emb.k_dim - k < 64
    out = emb(x, ids)
    ref = _reference(x.reshape(n, -1), w1, b1, w2, b2, tabs, ids)
    torch.testing.assert_close(out, ref, rtol=2e-2, atol=2e-2)
    xb, w1b, w2b = x.reshape(n, -1).bfloat16().double(), w1.bfloat16().double(), w2.bfloat16().double()
    h = (xb @ w1b.T + b1.bfloat16().double()).float().bfloat16().float()
    h = torch.where(h > 0, h, 0.01 * h).bfloat16().double()
    y = (h @ w2b.T + b2.bfloat16().double()).float().bfloat16().float()
    exact = y + (tabs[0][ids[:, 0]] + tabs[1][ids[:, 1]] + tabs[2][ids[:, 2]])
    assert float((out - exact).abs().max()) < 6e-2 and float(((out - exact).abs() > 1e-6).float().mean()) < 0.05
    assert torch.equal(emb(x.bfloat16(), ids), emb(x.bfloat16().float(), ids))          # bf16 tokens take the same path


def test_patch_embed_shapes_and_errors():
    from fluid_llm_b200.patch_embed import PatchEmbedder
    w1, b1, w2, b2 = torch.randn(512, 768), torch.zeros(512), torch.randn(768, 512), torch.zeros(768)
    emb = PatchEmbedder(w1, b1, w2, b2)
    x = torch.randn(2, 3, 60, 3, 16, 16, device="cuda")
    out = emb(x)
    assert out.shape == (2, 3, 60, 768)
    ref = _reference(x.reshape(-1, 768), w1.cuda(), b1.cuda(), w2.cuda(), b2.cuda(), None, None).view(2, 3, 60, 768)
    torch.testing.assert_close(out, ref, rtol=2e-2, atol=2e-2)
    with pytest.raises(ValueError):
        emb(x, torch.zeros(2, 3, 60, 3, dtype=torch.int64, device="cuda"))      # no tables
    with pytest.raises(ValueError):
        PatchEmbedder(w1, b1, torch.randn(768, 256), b2)


def test_graphed_patch_embed_replays_the_same_bits():
    from fluid_llm_b200.patch_embed import PatchEmbedder
    g = torch.Generator(device="cuda").manual_seed(5)
    w1, b1 = torch.randn(512, 768, device="cuda", generator=g) * 0.03, torch.randn(512, device="cuda", generator=g) * 0.1
    w2, b2 = torch.randn(768, 512, device="cuda", generator=g) * 0.04, torch.randn(768, device="cuda", generator=g) * 0.1
    tabs = [torch.randn(m, 768, device="cuda", generator=g) * 0.03 for m in (20, 10, 30)]
    emb = PatchEmbedder(w1, b1, w2, b2, *tabs)
    ge = emb.graphed(600, with_position_ids=True)
    for seed in (1, 2):
        x = torch.randn(10, 60, 3, 16, 16, device="cuda", generator=g)
        ids = torch.stack([torch.randint(0, m, (10, 60), device="cuda", generator=g) for m in (20, 10, 30)], dim=-1)
        assert torch.equal(ge(x, ids), emb(x, ids).reshape(600, 768))
    with pytest.raises(ValueError):
        ge(x)


@pytest.mark.parametrize("n_tokens", [130, 19000])          # a ragged last row block; more tiles than SMs
def test_patch_embed_writes_stay_inside_the_output_buffers(n_tokens):
    from fluid_llm_b200._lib import check, load, ptr, stream_ptr
    g = torch.Generator(device="cuda").manual_seed(11)
    G = 8192
    x = (torch.randn(n_tokens, 768, device="cuda", generator=g)).bfloat16()
    w1 = (torch.randn(512, 768, device="cuda", generator=g) * 0.03).bfloat16()
    w2 = (torch.randn(768, 512, device="cuda", generator=g) * 0.04).bfloat16()
    b1, b2 = torch.zeros(512, device="cuda"), torch.zeros(768, device="cuda")
    hidden = torch.full((G + n_tokens * 512 + G,), 3.0, dtype=torch.bfloat16, device="cuda")
    out = torch.full((G + n_tokens * 768 + G,), 777.0, dtype=torch.float32, device="cuda")
    check(load().fl_patch_embed(ptr(x), ptr(w1), ptr(b1), ptr(w2), ptr(b2), None, None, None, None, 0, 0, 0, ptr(hidden[G:]), ptr(out[G:]),
                                n_tokens, 768, 512, 768, stream_ptr()), "fl_patch_embed")
    torch.cuda.synchronize()
    assert bool((out[:G] == 777.0).all()) and bool((out[-G:] == 777.0).all())
    assert bool((hidden[:G] == 3.0).all()) and bool((hidden[-G:] == 3.0).all())
    ref = torch.nn.functional.leaky_relu((x.float() @ w1.float().T).bfloat16().float(), 0.01).bfloat16()
    y = (ref.float() @ w2.float().T).bfloat16().float()
    torch.testing.assert_close(out[G:-G].view(n_tokens, 768), y, rtol=2e-2, atol=2e-2)


def test_rollout_embed_cache_equals_reembedding_the_whole_context():
    """model.py:187-204: every step re-embeds the last ctx states with time ids re-based to 0.  RolloutEmbedCache embeds only
    the new state and re-applies the positional add over its ring: same bits as embedding the concatenated context."""
    from fluid_llm_b200.patch_embed import PatchEmbedder, RolloutEmbedCache
    g = torch.Generator(device="cuda").manual_seed(9)
    B, L, ctx, d = 3, 60, 4, 768
    w1, b1 = torch.randn(512, 768, device="cuda", generator=g) * 0.03, torch.randn(512, device="cuda", generator=g) * 0.1
    w2, b2 = torch.randn(d, 512, device="cuda", generator=g) * 0.04, torch.randn(d, device="cuda", generator=g) * 0.1
    tabs = [torch.randn(m, d, device="cuda", generator=g) * 0.03 for m in (15, 4, 12)]
    emb = PatchEmbedder(w1, b1, w2, b2, *tabs)
    cache = RolloutEmbedCache(emb, B, L, ctx)
    xs = torch.arange(L, device="cuda") // 4
    ys = torch.arange(L, device="cuda") % 4
    states = []
    for step in range(7):                      # the ring fills up, then wraps
        s = torch.randn(B, 1, L, 3, 16, 16, device="cuda", generator=g)
        states.append(s)
        cache.append(s.bfloat16() if step % 2 else s)            # bf16 tokens and fp32 states give the same embedding
        c = min(len(states), ctx)
        ids = torch.stack([xs.expand(B, c, L), ys.expand(B, c, L), torch.arange(c, device="cuda").view(1, c, 1).expand(B, c, L)], dim=-1)
        got = cache.tokens(ids)
        want = emb(torch.cat(states[-c:], dim=1), ids).view(B, c * L, d)
        assert torch.equal(got, want), step
    # the graphed form: one replay per step once the ring is full, a graph per ring position
    gc = RolloutEmbedCache(emb, B, L, ctx, graphs=True)
    hist = []
    for step in range(11):
        s = torch.randn(B, 1, L, 3, 16, 16, device="cuda", generator=g).bfloat16()
        hist.append(s)
        c = min(len(hist), ctx)
        ids = torch.stack([xs.expand(B, c, L), ys.expand(B, c, L), torch.arange(c, device="cuda").view(1, c, 1).expand(B, c, L)], dim=-1)
        if step >= ctx + 2:            # steady state: tokens written straight into the graphs' input, ids as in the previous call
            gc.token_buffer.copy_(s.reshape(gc.token_buffer.shape))
            got = gc.step(None, None)
        else:
            got = gc.step(s, ids)
        want = emb(torch.cat(hist[-c:], dim=1), ids).view(B, c * L, d)
        assert torch.equal(got, want), step
    assert len(gc.graphs) == ctx


def test_position_ids_outside_the_tables_raise_like_nn_embedding():
    from fluid_llm_b200.patch_embed import PatchEmbedder
    tabs = [torch.zeros(m, 768) for m in (5, 4, 3)]
    emb = PatchEmbedder(torch.randn(512, 768), torch.zeros(512), torch.randn(768, 512), torch.zeros(768), *tabs)
    x = torch.randn(6, 3, 16, 16, device="cuda")
    ids = torch.zeros(6, 3, dtype=torch.int64, device="cuda")
    emb(x, ids)
    ids[2, 1] = 4
    with pytest.raises(IndexError):
        emb(x, ids)
    assert emb(x, ids, validate_ids=False).shape == (6, 768)      # clamped, as the kernels always did
