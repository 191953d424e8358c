"""Per-frame dynamic meshes: frames/s of fl_dyn_interp_patchify on an EAGLE-shaped window,
beside the oracle (trapezoid map + plane interpolation per frame, one host thread).

    python tools/bench_dynamic.py [--frames 990] [--steps 10] [--warmup 3] [--cpu-frames 24]

Prints one JSON line.  Algorithmic bytes per frame: 8 N (positions) + 12 F (triangles) + 12 N (u, v, p) + 12 P (states);
the u8 mask and the binning workspace traffic are not counted."""
import argparse
import json
import os
import sys
import time


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=990)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--cpu-frames", type=int, default=24)
    ap.add_argument("--kind", default="eagle")
    args = ap.parse_args()
    import torch
    from fluid_llm_b200 import synth
    from fluid_llm_b200.dynamic_mesh import DynamicTrajectory
    from fluid_llm_b200.field_path import CYLINDER
    tr = synth.make_dynamic_trajectory(args.kind, args.frames, mesh_seed=0, field_seed=1, flip_frac=0.0)
    dt = DynamicTrajectory(tr["mesh_pos"], tr["cells"], tr["velocity"], tr["pressure"])
    T, N, F = args.frames, dt.n_nodes, dt.n_cells
    n_bx, n_by = dt.patch_grid((16, 16))
    P_px = n_bx * n_by * 256

    def step():
        return dt.interp_patchify(0, T, 1, (16, 16), CYLINDER)

    for _ in range(max(3, args.warmup)):
        step()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    torch.cuda.synchronize()
    sec = ev0.elapsed_time(ev1) / 1e3
    value = T * args.steps / sec
    algo = T * (8 * N + 12 * F + 12 * N + 12 * P_px)
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak = float(json.load(f)["hbm_gbs"])
    except Exception:
        peak = 6650.0
    out = {"metric": "grid-frames/sec, per-frame dynamic meshes (locate+interp+normalise+patchify)", "value": value,
           "unit": "frames/s", "n_gpus": 1, "steps": args.steps, "ms_per_step": sec / args.steps * 1e3,
           "config": {"workload": f"{args.kind}-shaped dynamic window: T={T} frames, each its own mesh of {N} nodes / {F} triangles, "
                                  f"grid {dt.nx}x{dt.ny}, {n_bx}x{n_by} patches of 16x16"},
           "roofline": {"bound": "hbm", "achieved": algo * args.steps / sec / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": algo * args.steps / sec / 1e9 / peak, "algorithmic_bytes_per_step": algo}}
    if args.cpu_frames > 0:
        from oracle import pipeline as P
        n = min(args.cpu_frames, T)
        t0 = time.perf_counter()
        P.dynamic_ds_get(tr["mesh_pos"], tr["cells"], tr["velocity"], tr["pressure"], 0, n, 1)
        cpu = n / (time.perf_counter() - t0)
        out["cpu_baseline"] = {"value": cpu, "unit": "frames/s", "cores": 1, "kind": "port",
                               "sample": f"{n} frames, trapezoid map built per frame, one thread"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
