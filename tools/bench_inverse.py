"""GB/s of the inverse-path and statistics kernels against the measured HBM copy peak.

    python tools/bench_inverse.py

Sizes are chosen so that every call moves more than the 126 MB L2 holds.  Algorithmic bytes per call as in DESIGN.md 4.3-4.6."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps / 1e3


def main():
    from fluid_llm_b200.compute_ds_stats import ds_stats
    from fluid_llm_b200.ds_props import DSProps
    from fluid_llm_b200.img_eagle import grid2mesh
    from fluid_llm_b200.utils_model import img_to_patch, patch_to_img, rollout_step
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    dev = "cuda"
    props = DSProps(15, 4, (16, 16), 10)                       # cylinder: 240 x 64 image, 60 patches
    P = 240 * 64
    rows = []

    bs, T = 64, 32                                            # 2048 frames: 377 MB each way
    patches = torch.randn(bs, T, 60, 3, 16, 16, device=dev)
    img = torch.randn(bs, T, 3, 240, 64, device=dev)
    with torch.no_grad():
        t = timeit(lambda: patch_to_img(patches, props))
        rows.append(("patch_to_img (a16)", bs * T * 24 * P, t))
        t = timeit(lambda: img_to_patch(img, props))
        rows.append(("img_to_patch (a17)", bs * T * 24 * P, t))
        B = 2048                                              # rollout step over 2048 sequences at once
        last = torch.randn(B, 1, 60, 3, 16, 16, device=dev)
        pred = torch.randn(B, 1, 3, 240, 64, device=dev)
        mask = torch.rand(B, 1, 60, 3, 16, 16, device=dev) < 0.1
        t = timeit(lambda: rollout_step(last, pred, mask, props))
        rows.append(("rollout_step (a18): read last + pred + mask, write diffs + next", B * (4 * 12 * P + 3 * P), t))
        Tg, N = 990, 3388                                     # grid2mesh: EAGLE constants, 256 x 128 grid
        vg, pg = torch.randn(Tg, 128, 256, 2, device=dev), torch.randn(Tg, 128, 256, 2, device=dev)
        pos = torch.stack([torch.rand(Tg, N, device=dev) * 4.9 - 2.45, torch.rand(Tg, N, device=dev) * 3.0 - 1.6], dim=-1)
        t = timeit(lambda: grid2mesh(vg, pg, pos, check_bounds=False))
        rows.append(("grid2mesh (a19): positions + gathered cells in, node values out", Tg * N * (8 + 2 * (8 + 8)), t))
        Ts = 2048
        states = torch.randn(Ts, 60, 3, 16, 16, device=dev)
        m = (torch.rand(Ts, 60, 16, 16, device=dev) < 0.1).to(torch.uint8)
        t = timeit(lambda: ds_stats(states, m))
        rows.append(("ds_stats (a20): states + mask read once", Ts * (12 * P + P), t))
    for name, nbytes, t in rows:
        print(f"{name:70s} {t * 1e6:9.1f} us  {nbytes / 1e6:9.1f} MB  {nbytes / t / 1e9:8.1f} GB/s  {nbytes / t / 1e9 / peak:6.1%} of {peak:.0f} GB/s")


if __name__ == "__main__":
    main()
