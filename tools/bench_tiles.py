#!/usr/bin/env python
"""Sweep of the tiled per-step kernel (csrc/fl_tiled.cu) over patches-per-tile, next to the whole-mesh staged kernel.

    python tools/bench_tiles.py [--workload airfoil] [--tiles 0,8,16,24] [--steps 20]

tile size 0 = no tile plan (k_interp_patchify_staged / gather).  Prints one line per variant: ms per launch, frames/s and
the fraction of the measured HBM roofline (same algorithmic bytes as bench.py).  Every variant's output is compared
bit for bit with the first one's.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="airfoil")
    ap.add_argument("--tiles", default="0,8,16,24,32,48")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--tf", default="")
    ap.add_argument("--prod", default="", help="comma-separated FLUIDGRID_PROD_WARPS values (2 or 4)")
    ap.add_argument("--ring", default="0", help="comma-separated FLUIDGRID_RING values (1: the experimental frame-ring kernel where it fits -- tiles of <= 6 patches --, 0: node-list kernel)")
    ap.add_argument("--env", default="", help="extra environment for every variant, e.g. FLUIDGRID_RING_D=3,FLUIDGRID_RING_NT=4")
    ap.add_argument("--dbg", default="", help="comma-separated FLUIDGRID_DBG values (development ablations; results are wrong)")
    args = ap.parse_args()
    import torch
    import fluid_llm_b200
    from fluid_llm_b200.field_path import AIRFOIL, CYLINDER, DeviceTrajectory, TrajBatch
    from fluid_llm_b200.mesh_utils import MeshPlan
    fluid_llm_b200.load()
    w = bench.WORKLOADS[args.workload]
    pers = AIRFOIL if w["personality"] == "airfoil" else CYLINDER
    meshes, trajs = bench.make_inputs(w, 0)
    dev = torch.device("cuda", 0)
    plans = [MeshPlan(pos, cells, w.get("res", bench.RES), device=dev) for (pos, cells, _) in meshes]
    tables = [p.patch_table(bench.PATCH, pers.crop_patches, pers.flip_y) for p in plans]
    dtrajs = [DeviceTrajectory(vel, prs, plans[mi]) for (mi, vel, prs) in trajs]
    tab = tables[0]
    P_px = tab.n_patches * 256
    algo = sum(w["T"] * (12 * plans[mi].n_nodes + 12 * P_px) for (mi, _, _) in trajs) + len(meshes) * 32 * P_px
    peak, _ = bench.measured_peak()
    ref = None
    tfs = [int(x) for x in args.tf.split(",")] if args.tf else [0]
    dbgs = [int(x) for x in args.dbg.split(",")] if args.dbg else [0]
    prods = [int(x) for x in args.prod.split(",")] if args.prod else [0]
    for kv in filter(None, args.env.split(",")):
        os.environ[kv.split("=")[0]] = kv.split("=")[1]
    rings = [int(x) for x in args.ring.split(",")]
    for tp, dbg, prod, ring in [(int(x), d, pw, rg) for x in args.tiles.split(",") for d in dbgs for pw in prods for rg in rings]:
        os.environ["FLUIDGRID_DBG"] = str(dbg)
        os.environ["FLUIDGRID_RING"] = str(ring)
        if prod:
            os.environ["FLUIDGRID_PROD_WARPS"] = str(prod)
        else:
            os.environ.pop("FLUIDGRID_PROD_WARPS", None)
        for tf in tfs:
            if tf:
                os.environ["FLUIDGRID_TF"] = str(tf)
            else:
                os.environ.pop("FLUIDGRID_TF", None)
            batch = TrajBatch(dtrajs, [tables[mi] for (mi, _, _) in trajs], [0] * len(trajs), 1, w["T"], want_mask=True,
                              tile_patches=tp)
            for _ in range(3):
                batch.run(pers)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                batch.run(pers)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            chk = (batch.states.view(torch.int32).sum(dtype=torch.int64).item(), int(batch.mask.sum().item()))
            if ref is None:
                ref = chk
            tpl = batch.tile_plans[0] if batch.tile_plans else None
            print(json.dumps({"workload": args.workload, "tile_patches": tp, "tf_cap": tf, "dbg": dbg, "prod_warps": prod, "kernel": fluid_llm_b200.load().fl_last_interp_kernel().decode(), "n_tiles": tpl.n_tiles if tpl else 0,
                              "max_tile_nodes": tpl.max_tile_nodes if tpl else 0, "ms": round(ms, 4),
                              "frames_per_s": round(len(trajs) * w["T"] / ms * 1e3), "roofline_frac": round(algo / ms / 1e6 / peak, 4),
                              "same_bits_as_first": chk == ref}), flush=True)
            del batch


if __name__ == "__main__":
    main()
