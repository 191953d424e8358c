#!/usr/bin/env python
"""bench.py -- grid-frames/s of the per-timestep field data path (interp + normalise + patchify).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload airfoil|cylinder|eagle]

One "step" = one pass of the hot path over one batch of synthetic trajectories (BASELINE.json
configs[1] by default: Airfoil-shaped, ~5k-node mesh cropped as airfoil_ds.py:164-183 does,
238 x 142 grid, 13 x 7 patches of 16 x 16, T = 600 frames per trajectory).  Prints ONE JSON line.

  value      frames/s, inputs resident in HBM, CUDA-event time over exactly K steps, max over ranks
  e2e        same metric through the public API with HOST buffers: pinned host -> device copy of the
             node fields, the kernel, device -> host copy of states + mask, all inside the timed region
  roofline   algorithmic bytes per launch / average launch duration against the measured HBM peak
  cpu_baseline  the oracle (CPU restatement of the reference path) timed on this box's host cores

--impl reference times the reference's CPU path (the oracle port: the reference's own arithmetic is
in matplotlib's C++ which is not installed, see oracle/tri_oracle.cpp) on all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

PATCH = (16, 16)
RES = 238
WORKLOADS = {
    # kind, personality, trajectories per GPU per step, frames per trajectory
    "airfoil": dict(kind="airfoil", personality="airfoil", n_traj=16, T=600),
    "cylinder": dict(kind="cylinder", personality="cylinder", n_traj=24, T=600),
    "eagle": dict(kind="eagle", personality="cylinder", n_traj=12, T=990),
    # BASELINE config 5: ~1M-triangle mesh, grid_res 2048 (2048 x 1024 cells, 8192 patches), T = 64
    "big": dict(kind="big", personality="cylinder", n_traj=4, T=64, res=2048, n_meshes=1),
}
N_MESHES = 4   # distinct meshes per GPU (trajectories cycle through them)
# padded grid in patches every rank must agree on (so per-GPU work is identical at every N)
WORKLOAD_GRIDS = {"airfoil": (15, 9), "cylinder": (15, 4), "eagle": (15, 10), "big": (128, 64)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def sample(self):
        if self.nv is None:
            return
        nv = self.nv
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                     0x80: "hw_power_brake_slowdown"}
            for bit, name in names.items():
                if r & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def run(self):
        while not self.stop_flag:
            self.sample()
            time.sleep(0.005)

    def result(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def make_inputs(w, rank):
    """Seeded synthetic trajectories of the workload (host arrays), cropped as the dataset would."""
    from fluid_llm_b200 import synth
    from fluid_llm_b200.airfoil_ds import crop_airfoil_mesh
    from fluid_llm_b200.mesh_utils import _grid_shape
    meshes, want, seed = [], None, 100 * rank
    while len(meshes) < w.get("n_meshes", N_MESHES):
        pos, cells = synth.make_mesh(w["kind"], seed=seed)
        seed += 1
        sel = None
        if w["personality"] == "airfoil":
            sel, pos, cells = crop_airfoil_mesh(pos, cells)
        # one batch = one patch grid (the reference's DataLoader collation needs that too): skip the rare synthetic mesh
        # whose cropped bounding box lands on the other side of a multiple of the patch size
        lo, hi = pos.min(axis=0), pos.max(axis=0)
        nx, ny = _grid_shape(lo[0], hi[0], lo[1], hi[1], w.get("res", RES), "1.26")
        grid = (-(-nx // PATCH[0]), -(-ny // PATCH[1]))
        if want is None:
            want = WORKLOAD_GRIDS.get(w["kind"], grid)
        if grid != want:
            continue
        meshes.append((pos, cells, sel))
    trajs = []
    for i in range(w["n_traj"]):
        pos, cells, sel = meshes[i % len(meshes)]
        n_full = len(sel) if sel is not None else len(pos)
        full_pos = np.zeros((n_full, 2), np.float32)
        if sel is not None:
            full_pos[sel] = pos
        else:
            full_pos = pos
        vel, prs = synth.make_fields(w["kind"], full_pos, w["T"], seed=1000 * rank + i)
        if sel is not None:
            vel, prs = np.ascontiguousarray(vel[:, sel]), np.ascontiguousarray(prs[:, sel])
        trajs.append((i % len(meshes), vel, prs))
    return meshes, trajs


def workload_text(name, w, meshes):
    """Same text for both arms: what one step of the full workload is."""
    n = int(np.mean([len(m[0]) for m in meshes]))
    return (f"{name}-shaped (BASELINE.json configs): {WORKLOADS[name]['n_traj']} trajectories/GPU x T={WORKLOADS[name]['T']} frames per step, "
            f"meshes of ~{n} nodes (after the dataset's crop), grid_res {w.get('res', RES)}, patch 16x16")


def cpu_frames_per_s(w, meshes, trajs, budget_s, n_frames_cap=None):
    """The oracle's per-frame path (3x to_grid + pad + unfold + normalise), one thread, trifinder
    built outside the timed loop (the most favourable reading of the reference's CPU path)."""
    from oracle import pipeline as P
    mi, vel, prs = trajs[0]
    pos, cells, _ = meshes[mi]
    triang, tri_index, gx, gy = P.get_mesh_interpolation(pos, cells, w.get("res", RES))
    done, t0 = 0, time.perf_counter()
    cap = n_frames_cap or 10 ** 9          # no cap: keep cycling through the trajectory until the time budget is spent
    T = vel.shape[0]
    while done < cap:
        chunk = []
        for i in range(done, min(cap, done + 10)):
            state, mask = P.get_step(triang, tri_index, gx, gy, vel, prs, i % T, PATCH)
            chunk.append(np.concatenate([state, mask[None].astype(state.dtype)], axis=0))
        seq = np.stack(chunk).astype(np.float32)
        if w["personality"] == "airfoil":
            seq = np.ascontiguousarray(seq[:, :, :, ::-1])[:, :, PATCH[0]:-PATCH[0], PATCH[1]:-PATCH[1]]
        patches = P.unfold_patches(seq, PATCH)
        states = np.ascontiguousarray(patches[:, :-1].transpose(0, 4, 1, 2, 3))
        masks = np.ascontiguousarray(patches[:, -1].transpose(0, 3, 1, 2))
        P.normalize(states, masks, w["personality"])
        done += len(chunk)
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return done / dt, done, dt


def _cpu_worker(args):
    w, meshes, traj, n_frames = args
    fps, done, dt = cpu_frames_per_s(w, meshes, [traj], 1e9, n_frames)
    return done


def run_reference(args, w):
    """--impl reference: the reference's CPU path (oracle port) on all host cores; each step is a
    bounded sample of the workload (frames spread over one process per core, like DataLoader workers)."""
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    meshes, trajs = make_inputs(dict(w, n_traj=min(w["n_traj"], N_MESHES), T=64), 0)
    fps1, _, _ = cpu_frames_per_s(w, meshes, trajs, 2.0, 64)          # calibrate: frames per core-second
    total_steps = args.steps + args.warmup
    per_worker = max(4, min(64, int(fps1 * 60.0 / max(total_steps, 1))))   # whole run ~ a minute or two
    jobs = [(w, meshes, trajs[i % len(trajs)], per_worker) for i in range(cores)]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for _ in range(args.warmup):
            pool.map(_cpu_worker, jobs)
        t0 = time.perf_counter()
        frames = 0
        for _ in range(args.steps):
            frames += sum(pool.map(_cpu_worker, jobs))
        dt = time.perf_counter() - t0
    value = frames / dt
    line = {"impl": "reference", "metric": "grid-frames/sec (interp+normalise+patchify)", "value": value, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_text(args.workload, w, meshes),
                       "sample": f"each step = {per_worker} frames x {cores} worker processes of that workload", "frames_per_step": frames // args.steps},
            "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port",
                             "sample": f"{per_worker} frames x {cores} processes per step; trifinder build outside the timed loop"},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_ours(args, w):
    import torch
    import torch.distributed as dist
    import fluid_llm_b200
    from fluid_llm_b200 import compute_ds_stats
    from fluid_llm_b200.field_path import AIRFOIL, CYLINDER, DeviceTrajectory, HostPipeline, TrajBatch
    from fluid_llm_b200.mesh_utils import MeshPlan

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    try:        # run (and first-touch the pinned host buffers) on the CPUs next to this rank's GPU
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
    except Exception:
        pass
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    fluid_llm_b200.load()
    pers = AIRFOIL if w["personality"] == "airfoil" else CYLINDER

    meshes, trajs = make_inputs(w, rank)
    plans = [MeshPlan(pos, cells, w.get("res", RES), device=dev) for (pos, cells, _) in meshes]
    tables = [p.patch_table(PATCH, pers.crop_patches, pers.flip_y) for p in plans]
    dtrajs = [DeviceTrajectory(vel, prs, plans[mi]) for (mi, vel, prs) in trajs]
    batch = TrajBatch(dtrajs, [tables[mi] for (mi, _, _) in trajs], [0] * len(trajs), 1, w["T"], want_mask=True)
    tab = tables[0]
    n_frames_step = len(trajs) * w["T"]
    P_px = tab.n_patches * tab.px * tab.py
    n_nodes = [plans[mi].n_nodes for (mi, _, _) in trajs]
    # algorithmic bytes: read u,v,p once (12 N), write 3-channel fp32 patches once (12 P)
    # [+ the u8 mask this run also writes is NOT counted]; static table 32 B per output pixel per mesh
    bytes_frames = sum(w["T"] * (12 * n + 12 * P_px) for n in n_nodes)
    bytes_static = len(meshes) * 32 * P_px
    algo_bytes = bytes_frames + bytes_static

    def step():
        batch.run(pers)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    sampler.sample()
    torch.cuda.synchronize()
    sampler.stop_flag = True
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    sec = ms / 1e3
    value = world * n_frames_step * args.steps / sec
    peak, peak_src = measured_peak()
    achieved = algo_bytes * args.steps / sec / 1e9       # per GPU: every rank launches the same kernel on its own shard
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "r1_final_traffic.json")) as f:
            tj = json.load(f)
        if tj.get("workload") == args.workload:
            traffic = tj["traffic_bytes_per_launch"]
    except Exception:
        pass

    # ---- e2e: host buffers in, host buffers out, through the same public API ----------------
    e2e_traj = min(4, len(trajs))
    e2e_T = w["T"]
    h_vel = [torch.from_numpy(trajs[i][1]).pin_memory() for i in range(e2e_traj)]
    h_prs = [torch.from_numpy(trajs[i][2]).pin_memory() for i in range(e2e_traj)]
    # three streams, two device slots: upload of trajectory i+1 / kernel of i / download of i-1 overlap
    pipe = HostPipeline([plans[trajs[i][0]] for i in range(e2e_traj)], [tables[trajs[i][0]] for i in range(e2e_traj)], pers,
                        n_steps=e2e_T, t0=0, interval=1, n_frames=e2e_T, depth=2)
    tab0 = tables[trajs[0][0]]
    h_states = [torch.empty((e2e_T, tab0.n_patches, 3, tab0.px, tab0.py), dtype=torch.float32).pin_memory() for _ in range(e2e_traj)]
    h_mask = [torch.empty((e2e_T, tab0.n_patches, tab0.px, tab0.py), dtype=torch.uint8).pin_memory() for _ in range(e2e_traj)]
    h2d = sum(t.numel() * 4 for t in h_vel) + sum(t.numel() * 4 for t in h_prs)
    d2h = sum(t.numel() * 4 for t in h_states) + sum(t.numel() for t in h_mask)

    def e2e_step():
        pipe.run(h_vel, h_prs, h_states, h_mask)

    e2e_steps = max(2, min(args.steps, 20))
    e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_sec = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_sec], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_sec = float(t.item())
    e2e_value = world * e2e_traj * e2e_T * e2e_steps / e2e_sec
    # the same call with a consumer on the GPU (the training / rollout case): inputs still come from the host every step,
    # outputs stay in HBM and only a per-trajectory checksum (3 floats) is read back
    sums = torch.zeros((e2e_traj, 3), dtype=torch.float32, device=dev)
    h_sums = torch.empty((e2e_traj, 3), dtype=torch.float32).pin_memory()

    def consume(i, states, mask):
        sums[i] = states.sum(dim=(0, 1, 3, 4))

    def e2e_dev_step():
        pipe.run(h_vel, h_prs, on_device=consume)
        with torch.cuda.stream(pipe.s_run):
            h_sums.copy_(sums, non_blocking=True)

    e2e_dev_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_dev_step()
    torch.cuda.synchronize()
    dev_sec = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dev_sec], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_sec = float(t.item())
    e2e_dev_value = world * e2e_traj * e2e_T * e2e_steps / dev_sec

    # ---- the one-off per-mesh step, reported separately (cells located per second) ----
    plans[0].locate()
    torch.cuda.synchronize()
    reps = []
    for _ in range(10):
        t0 = time.perf_counter()
        plans[0].locate()            # four kernels + its own stream sync (the bin-item count is read back)
        reps.append((time.perf_counter() - t0) * 1e3)
    loc_ms = float(np.median(reps))
    locate = {"ms_per_mesh": loc_ms, "cells_per_s": plans[0].nx * plans[0].ny / (loc_ms / 1e3),
              "mesh": f"{plans[0].n_nodes} nodes / {plans[0].n_cells} triangles -> {plans[0].nx}x{plans[0].ny} cells"}

    # ---- the path's only collective: dataset statistics merged over ranks (functional check) ----
    agg = compute_ds_stats.ds_stats(batch.states[0], batch.mask[0])
    agg = compute_ds_stats.all_reduce_stats(agg)
    torch.cuda.synchronize()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        fps, done, dt = cpu_frames_per_s(w, meshes, trajs, 12.0)
        cpu = {"value": fps, "unit": "frames/s", "cores": 1, "kind": "port",
               "sample": f"{done} frames of one trajectory of the workload in {dt:.1f} s, one thread (of {os.cpu_count()} host cores); "
                         "trifinder build outside the timed loop"}

    if rank == 0:
        line = {"metric": "grid-frames/sec (interp+normalise+patchify)", "value": value, "unit": "frames/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_text(args.workload, w, meshes),
                           "grid": f"{plans[0].nx}x{plans[0].ny} cells, {tab.n_bx}x{tab.n_by} patches of 16x16",
                           "frames_per_step_per_gpu": n_frames_step, "parallelism": f"trajectory-sharded x{world}",
                           "l2_policy": "inputs+outputs per step (%.0f MB) exceed the 126 MB L2" % (algo_bytes / 1e6),
                           "mesh_seed": "100*rank + 0,1,2,.. (meshes with another patch grid skipped)", "field_seed": "1000*rank+i", "stats_check_n": float(agg[0, 0].item())},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": algo_bytes, "kernel": "k_interp_patchify_staged", "traffic_source": "ncu --set full, profiles/r1_final_traffic.json" if traffic else None},
                "cpu_baseline": cpu,
                "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "note": f"{e2e_traj} trajectories/step/GPU, pinned host buffers both ways, HostPipeline (3 streams, 2 device slots)",
                        "outputs_consumed_on_gpu": {"value": e2e_dev_value, "unit": "frames/s", "d2h_bytes_per_step": int(h_sums.numel() * 4),
                                                    "note": "same host inputs every step; states stay in HBM for a GPU consumer (per-channel sums read back)"}},
                "gpu_launches": args.steps,
                "locate_one_off": locate,
                "clocks": sampler.result()}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="airfoil", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_ours(args, w)


if __name__ == "__main__":
    main()
