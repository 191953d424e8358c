"""Time the tcgen05 patch-embedding projection (fl_patch_embed) next to PyTorch/cuBLAS bf16 autocast.
usage: python tools/bench_embed.py [n_tokens ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from fluid_llm_b200.patch_embed import PatchEmbedder


def timeit(fn, reps=50):
    for _ in range(5):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [4800, 38400, 153600]
    w1, b1 = torch.randn(512, 768, device="cuda") * 768 ** -0.5, torch.randn(512, device="cuda")
    w2, b2 = torch.randn(768, 512, device="cuda") * 512 ** -0.5, torch.randn(768, device="cuda")
    tabs = [torch.randn(m, 768, device="cuda") for m in (20, 10, 30)]
    emb = PatchEmbedder(w1, b1, w2, b2, *tabs)
    for n in sizes:
        x = torch.randn(n, 768, device="cuda")
        xb = x.bfloat16()
        if os.environ.get("EMBED_RANDOM_IDS"):
            ids = torch.stack([torch.randint(0, m, (n,), device="cuda") for m in (20, 10, 30)], dim=1)   # worst case: no table-row reuse
        else:
            # the ids the data path produces (simple_dataloader.py:218-226): 60 patches per frame, frames in order
            a = torch.arange(n, device="cuda")
            ids = torch.stack([a % 15, (a // 15) % 4, (a // 60) % 30], dim=1)

        def ref():
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y = F.linear(F.leaky_relu(F.linear(x, w1, b1), 0.01), w2, b2)
            return y + (tabs[0][ids[:, 0]] + tabs[1][ids[:, 1]] + tabs[2][ids[:, 2]])

        flops = 2.0 * n * (768 * 512 + 512 * 768)
        t_ours, t_ours_bf16, t_ref = timeit(lambda: emb(x, ids)), timeit(lambda: emb(xb, ids)), timeit(ref)
        g32, g16 = emb.graphed(n, True, torch.float32), emb.graphed(n, True, torch.bfloat16)
        g32.x.copy_(x); g32.ids.copy_(ids); g16.x.copy_(xb); g16.ids.copy_(ids)
        assert torch.equal(g32.replay(), emb(x, ids)) and torch.equal(g16.replay(), emb(xb, ids))
        t_g32, t_g16 = timeit(g32.replay), timeit(g16.replay)
        print(f"tokens {n:7d}: ours {t_ours * 1e3:8.1f} us ({flops / t_ours / 1e9:7.1f} TFLOP/s)  bf16-in {t_ours_bf16 * 1e3:8.1f} us "
              f"({flops / t_ours_bf16 / 1e9:7.1f} TFLOP/s)  graphed {t_g32 * 1e3:8.1f} / bf16-in {t_g16 * 1e3:8.1f} us "
              f"({flops / t_g16 / 1e9:7.1f} TFLOP/s)  torch autocast {t_ref * 1e3:8.1f} us ({flops / t_ref / 1e9:7.1f} TFLOP/s)")


if __name__ == "__main__":
    main()
