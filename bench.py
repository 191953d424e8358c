#!/usr/bin/env python
"""bench.py -- grid-frames/s of the per-timestep field data path (interp + normalise + patchify).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload airfoil|cylinder|eagle|big]
                    [--no-extra] [--no-cpu-baseline]

One "step" = one pass of the hot path over one batch of synthetic trajectories (BASELINE.json configs[1] by default:
Airfoil-shaped, ~5k-node mesh cropped as airfoil_ds.py:164-183 does, 238 x 142 grid, 13 x 7 patches of 16 x 16, T = 600
frames per trajectory).  A pass is `launches_per_step` back-to-back launches over the resident batch, chosen during the
warm-up so that the K timed steps last at least ~0.6 s (clock sampling needs that).  Prints ONE JSON line.

  value        frames/s, inputs resident in HBM, CUDA-event time over exactly K steps, max over ranks
  e2e          same metric through the public API with HOST buffers: pinned host -> device copy of the node fields, the
               kernel, device -> host copy of states + mask, all inside the timed region; next to it the same call with a
               consumer on the GPU, and the box's plain pinned-copy ceilings measured on all ranks at once
  roofline     algorithmic bytes per launch / average launch duration against the measured HBM peak; `kernel` is the kernel
               the library actually launched (fl_last_interp_kernel)
  workloads    the other BASELINE configs from the same process: cylinder, eagle (mean/std from the statistics kernel,
               merged over ranks), big (1M-triangle mesh, 256 trajectories sharded over the ranks, ending in the
               statistics merge; `stats_allgather_us` is the collective alone, >= 100 repetitions)
  cpu_baseline the oracle (CPU restatement of the reference path) timed on this box's host cores: one thread, min(6, cores)
               worker processes (configs/training1.yaml:73), and the per-sample trifinder build the reference pays on top

--impl reference times the reference's CPU path (the oracle port: the reference's own arithmetic is in matplotlib's C++
which is not installed, see oracle/tri_oracle.cpp) on all host cores.
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
METRIC = "grid-frames/sec (interp+normalise+patchify)"
WORKLOADS = {
    # kind, personality, trajectories per GPU per launch, frames per trajectory
    "airfoil": dict(kind="airfoil", personality="airfoil", n_traj=16, T=600),
    "cylinder": dict(kind="cylinder", personality="cylinder", n_traj=24, T=600),
    "eagle": dict(kind="eagle", personality="cylinder", n_traj=12, T=990),
    # BASELINE config 5: ~1M-triangle mesh, grid_res 2048 (2048 x 1024 cells, 8192 patches), T = 64; 256 trajectories in all,
    # sharded over the ranks and processed 4 per launch
    "big": dict(kind="big", personality="cylinder", n_traj=4, T=64, res=2048, n_meshes=1, total_traj=256),
}
N_MESHES = 4   # distinct meshes per GPU (trajectories cycle through them)
# padded grid in patches every rank must agree on (so per-GPU work is identical at every N)
WORKLOAD_GRIDS = {"airfoil": (15, 9), "cylinder": (15, 4), "eagle": (15, 10), "big": (128, 64)}
MIN_TIMED_S = 0.6


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
    """Seeded synthetic trajectories of the workload (host arrays), cropped as the dataset would.  One field set is
    generated per distinct mesh; further trajectories on the same mesh are that set rolled in time (different frames at
    every step number, same cost to process)."""
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
    fields = []
    for mi, (pos, cells, sel) in enumerate(meshes):
        n_full = len(sel) if sel is not None else len(pos)
        full_pos = np.zeros((n_full, 2), np.float32)
        if sel is not None:
            full_pos[sel] = pos
        else:
            full_pos = pos
        vel, prs = synth.make_fields(w["kind"], full_pos, w["T"], seed=1000 * rank + mi)
        if sel is not None:
            vel, prs = np.ascontiguousarray(vel[:, sel]), np.ascontiguousarray(prs[:, sel])
        fields.append((vel, prs))
    trajs = []
    for i in range(w["n_traj"]):
        mi, lap = i % len(meshes), i // len(meshes)
        vel, prs = fields[mi]
        if lap:
            vel, prs = np.roll(vel, 37 * lap, axis=0), np.roll(prs, 37 * lap, axis=0)
        trajs.append((mi, vel, prs))
    return meshes, trajs


def workload_text(name, w, meshes):
    """Same text for both arms: what one launch of the full workload is."""
    n = int(np.mean([len(m[0]) for m in meshes]))
    return (f"{name}-shaped (BASELINE.json configs): {WORKLOADS[name]['n_traj']} trajectories/GPU x T={WORKLOADS[name]['T']} frames per launch, "
            f"meshes of ~{n} nodes (after the dataset's crop), grid_res {w.get('res', RES)}, patch 16x16")


# ---------------------------------------------------------------------------------------------------------------------
# CPU side (oracle port)
# ---------------------------------------------------------------------------------------------------------------------
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


def cpu_pool_frames_per_s(w, meshes, trajs, workers, per_worker, rounds):
    """`workers` processes (the reference's DataLoader workers), each running the single-thread path on its own frames."""
    import multiprocessing as mp
    jobs = [(w, meshes, trajs[i % len(trajs)], per_worker) for i in range(workers)]
    ctx = mp.get_context("fork")
    with ctx.Pool(workers) as pool:
        pool.map(_cpu_worker, jobs)                  # warm-up: fork, first touches
        t0 = time.perf_counter()
        frames = 0
        for _ in range(rounds):
            frames += sum(pool.map(_cpu_worker, jobs))
        dt = time.perf_counter() - t0
    return frames / dt, frames, dt


def cpu_trifinder_ms(w, meshes, reps=3):
    """The one-off the reference pays on EVERY __getitem__ (simple_dataloader.py:181 -> mesh_utils.py:94-106): triangulation +
    trapezoid map + find_many over the grid."""
    from oracle import pipeline as P
    pos, cells, _ = meshes[0]
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        P.get_mesh_interpolation(pos, cells, w.get("res", RES))
        ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ts))


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
    sample = f"each step = {per_worker} frames x {cores} worker processes of that workload; trifinder build outside the timed loop"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_text(args.workload, w, meshes)},
            "details": {"sample": sample, "frames_per_step": frames // args.steps},
            "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# GPU side
# ---------------------------------------------------------------------------------------------------------------------
class Workload:
    """One workload resident on this rank's GPU: meshes located, trajectories uploaded, launch descriptors built."""

    def __init__(self, name, rank, dev, want_host=False):
        import torch
        from fluid_llm_b200.field_path import AIRFOIL, CYLINDER, DeviceTrajectory, TrajBatch
        from fluid_llm_b200.mesh_utils import MeshPlan
        self.name, self.w = name, WORKLOADS[name]
        w = self.w
        self.pers = AIRFOIL if w["personality"] == "airfoil" else CYLINDER
        self.meshes, self.trajs = make_inputs(w, rank)
        self.plans = [MeshPlan(pos, cells, w.get("res", RES), device=dev) for (pos, cells, _) in self.meshes]
        self.tables = [p.patch_table(PATCH, self.pers.crop_patches, self.pers.flip_y) for p in self.plans]
        self.dtrajs = [DeviceTrajectory(vel, prs, self.plans[mi]) for (mi, vel, prs) in self.trajs]
        self.batch = TrajBatch(self.dtrajs, [self.tables[mi] for (mi, _, _) in self.trajs], [0] * len(self.trajs), 1, w["T"],
                               want_mask=True)
        tab = self.tables[0]
        self.tab = tab
        self.P_px = tab.n_patches * tab.px * tab.py
        self.frames_per_launch = len(self.trajs) * w["T"]
        n_nodes = [self.plans[mi].n_nodes for (mi, _, _) in self.trajs]
        # algorithmic bytes: read u,v,p once (12 N), write 3-channel fp32 patches once (12 P)
        # [+ the u8 mask this run also writes is NOT counted]; static table 32 B per output pixel per mesh
        self.algo_bytes = sum(w["T"] * (12 * n + 12 * self.P_px) for n in n_nodes) + len(self.meshes) * 32 * self.P_px
        if not want_host:
            self.trajs = [(mi, None, None) for (mi, _, _) in self.trajs]      # the host copies are only needed by e2e / cpu legs
        torch.cuda.synchronize()

    def launch(self, means=None, stds=None):
        self.batch.run(self.pers, means=means, stds=stds)

    def kernel_name(self):
        import fluid_llm_b200
        return fluid_llm_b200.load().fl_last_interp_kernel().decode()


def time_launches(fn, n, dev):
    """n back-to-back calls of fn between two CUDA events on the current stream -> milliseconds"""
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1)


def max_over_ranks(x, dev, world):
    import torch
    import torch.distributed as dist
    if world == 1:
        return float(x)
    t = torch.tensor([x], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(world):
    import torch.distributed as dist
    if world > 1:
        dist.barrier()


def copy_ceilings(dev, world, nbytes=256 << 20, reps=6):
    """Plain pinned cudaMemcpyAsync in each direction, all ranks at once: what the box's host links give N GPUs together."""
    import torch
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    out = {}
    for name, (dst, src) in (("d2h", (h, d)), ("h2d", (d, h))):
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize(dev)
        barrier(world)
        t0 = time.perf_counter()
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize(dev)
        sec = max_over_ranks(time.perf_counter() - t0, dev, world)
        out[name + "_gbs_all_ranks"] = world * nbytes * reps / sec / 1e9
    return out


def run_extra_workload(name, rank, world, dev, peak):
    """cylinder / eagle / big from the same process: value, ms per launch, roofline fraction, kernel actually launched."""
    import torch
    from fluid_llm_b200 import compute_ds_stats
    wl = Workload(name, rank, dev)
    w = wl.w
    means = stds = None
    info = {}
    if name == "eagle":
        # BASELINE config 3: mean / std come from the statistics kernel (max/compute_ds_stats.py:52-75), merged over ranks,
        # and feed the normalisation of the timed launches
        wl.batch.run(wl.pers, normalize=False)
        parts = torch.stack([compute_ds_stats.ds_stats(wl.batch.states[i], wl.batch.mask[i]) for i in range(len(wl.dtrajs))])
        agg = compute_ds_stats.all_reduce_stats(compute_ds_stats.merge_stats(parts))
        m, s = compute_ds_stats.mean_std(agg)
        means, stds = tuple(float(x) for x in m[:3]), tuple(float(x) for x in s[:3])
        info["normalisation"] = {"means": means, "stds": stds, "source": "fl_ds_stats -> fl_stats_merge -> all-gather over ranks"}
    for _ in range(3):
        wl.launch(means, stds)
    torch.cuda.synchronize(dev)
    one = time_launches(lambda: wl.launch(means, stds), 2, dev) / 2
    passes = 1
    launches_per_pass = 1
    if name == "big":
        # 256 trajectories sharded over the ranks (field_path.shard_range), 4 resident trajectories re-run per launch
        from fluid_llm_b200.field_path import shard_range
        lo, hi = shard_range(w["total_traj"], rank, world)
        launches_per_pass = -(-(hi - lo) // w["n_traj"])
        info["sharding"] = f"{w['total_traj']} trajectories over {world} rank(s): {hi - lo} on rank {rank} = {launches_per_pass} launches of {w['n_traj']}"
    else:
        launches_per_pass = max(1, int(np.ceil(0.25e3 / max(one * 10, 1e-3))))     # ~0.25 s over the 10 timed passes
    n_pass = 10 if name != "big" else 2
    barrier(world)
    ms = time_launches(lambda: wl.launch(means, stds), n_pass * launches_per_pass, dev)
    ms = max_over_ranks(ms, dev, world)
    n_launch = n_pass * launches_per_pass
    value = world * wl.frames_per_launch * n_launch / (ms / 1e3)
    achieved = wl.algo_bytes * n_launch / (ms / 1e3) / 1e9
    res = {"value": value, "unit": "frames/s", "ms_per_launch": ms / n_launch, "launches_timed": n_launch,
           "frames_per_launch_per_gpu": wl.frames_per_launch,
           "roofline": {"achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "kernel": wl.kernel_name(),
                        "algorithmic_bytes_per_launch": wl.algo_bytes},
           "workload": workload_text(name, w, wl.meshes),
           "grid": f"{wl.plans[0].nx}x{wl.plans[0].ny} cells, {wl.tab.n_bx}x{wl.tab.n_by} patches"}
    res.update(info)
    if name == "big":
        # the pass ends with the statistics of what it produced: per-launch aggregates, merged on the rank, all-gathered
        t0 = time_launches(lambda: [compute_ds_stats.ds_stats(wl.batch.states[i], wl.batch.mask[i]) for i in range(len(wl.dtrajs))], 1, dev)
        parts = torch.stack([compute_ds_stats.ds_stats(wl.batch.states[i], wl.batch.mask[i]) for i in range(len(wl.dtrajs))])
        agg = compute_ds_stats.merge_stats(parts)
        res["stats_ms_per_launch_outputs"] = t0
        if world > 1:
            compute_ds_stats.all_reduce_stats(agg)
            torch.cuda.synchronize(dev)
            barrier(world)
            reps = 200
            us = time_launches(lambda: compute_ds_stats.all_reduce_stats(agg), reps, dev) / reps * 1e3
            res["stats_allgather_us"] = max_over_ranks(us, dev, world)
            res["stats_allgather_reps"] = reps
        merged = compute_ds_stats.all_reduce_stats(agg)
        res["stats_n_state_ch0"] = float(merged[0, 0].item())
    del wl
    torch.cuda.empty_cache()
    return res


def dataset_e2e(dev, n_files=8, n_samples=64, seq_len=10):
    """Samples/s through the reference-facing data set API (`MGNDataset.ds_get` / `ds_get_many`, the calls a DataLoader makes)
    on the reference's own pickle format: cold = every sample reads its 18 MB pickle (ingest pool unpickles ahead), warm =
    trajectories resident in HBM.  A few seconds on rank 0."""
    import pickle
    import shutil
    import tempfile
    import torch
    from fluid_llm_b200 import synth
    from fluid_llm_b200.simple_dataloader import MGNDataset
    tmp = tempfile.mkdtemp(prefix="fluidgrid_bench_ds_")
    try:
        for i in range(n_files):
            with open(os.path.join(tmp, f"save_{i:03d}.pkl"), "wb") as f:
                pickle.dump(synth.make_trajectory("cylinder", 600, mesh_seed=i, field_seed=100 + i), f)
        rng = np.random.default_rng(0)
        order = [(int(rng.integers(n_files)), int(rng.integers(0, 500))) for _ in range(n_samples)]
        out = {"sample": f"{n_samples} samples of seq_len {seq_len} from {n_files} cylinder-shaped pickles (T=600, 17.9 MB each, page cache warm); "
                         "cold: every sample reads a file of its own (hard links), second pass over the samples",
               "unit": "samples/s"}
        cold_dir = os.path.join(tmp, "cold")                # one path per sample, as in a data set of many files
        os.makedirs(cold_dir)
        for k, (fi, _) in enumerate(order):
            os.link(os.path.join(tmp, f"save_{fi:03d}.pkl"), os.path.join(cold_dir, f"save_{k:03d}.pkl"))
        cold_order = [(k, step) for k, (_, step) in enumerate(order)]

        def timed(fn):
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize(dev)
            return n_samples / (time.perf_counter() - t0)
        ds = MGNDataset(cold_dir, RES, PATCH, PATCH, seq_len, mode="valid", device=dev)
        ds.cache_size = 0                                   # cold: nothing stays resident

        def cold():
            for i, (fi, step) in enumerate(cold_order):
                ds.prefetch(cold_order[i:i + 24])            # what num_workers x prefetch_factor does (utils_model.LookaheadBatchSampler)
                ds.ds_get(fi, step)
        cold()                                              # (first pass: the allocators and the pool's interpreters warm up)
        out["pickle_cold_ingest_pool"] = timed(cold)
        out["ingest_workers"] = ds._ingest.workers if ds._ingest is not None else 0
        ds._ingest.close()
        ds = MGNDataset(tmp, RES, PATCH, PATCH, seq_len, mode="valid", device=dev)
        for i in range(n_files):
            ds.ds_get(i, 0)
        def one_by_one():                   # (samples are dropped as a consumer would: keeping all of them alive times cudaMalloc)
            for fi, step in order:
                ds.ds_get(fi, step)

        def in_batches():
            for i in range(0, n_samples, 8):
                ds.ds_get_many(order[i:i + 8])
        one_by_one()
        out["resident_one_sample_per_call"] = timed(one_by_one)
        in_batches()
        out["resident_batches_of_8"] = timed(in_batches)
        if ds._ingest is not None:
            ds._ingest.close()
        return out
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def run_ours(args, w):
    import torch
    import torch.distributed as dist
    import fluid_llm_b200
    from fluid_llm_b200 import compute_ds_stats
    from fluid_llm_b200.field_path import HostPipeline

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
    peak, peak_src = measured_peak()
    warmup = max(args.warmup, 3)

    wl = Workload(args.workload, rank, dev, want_host=True)
    pers, plans, tables, trajs, batch, tab = wl.pers, wl.plans, wl.tables, wl.trajs, wl.batch, wl.tab

    # ---- warm-up; size a step so that the K timed steps last >= MIN_TIMED_S ----
    for _ in range(warmup):
        wl.launch()
    torch.cuda.synchronize(dev)
    one_ms = time_launches(wl.launch, 3, dev) / 3
    lps = max(1, int(np.ceil(MIN_TIMED_S * 1e3 / (one_ms * args.steps))))
    if world > 1:
        t = torch.tensor([lps], device=dev, dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        lps = int(t.item())

    def step():
        for _ in range(lps):
            wl.launch()

    for _ in range(warmup):
        step()
    torch.cuda.synchronize(dev)
    barrier(world)
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    sampler.sample()
    torch.cuda.synchronize(dev)
    sampler.stop_flag = True
    ms = max_over_ranks(ev0.elapsed_time(ev1), dev, world)
    sec = ms / 1e3
    n_launches = args.steps * lps
    value = world * wl.frames_per_launch * n_launches / sec
    algo_bytes = wl.algo_bytes
    achieved = algo_bytes * n_launches / sec / 1e9          # per GPU: every rank launches the same kernel on its own shard
    kernel = wl.kernel_name()
    # the same kernel as a BURST (20 launches after the GPU has idled): the timed region above runs for >= 0.6 s under the
    # 1000 W power cap (see `clocks`), the copy that measured the roofline's denominator (MEASURED_PEAKS.json: best of 10
    # copies of 4 GB) did not.  Information only: `frac` stays the sustained number.
    time.sleep(1.0)
    burst_ms = time_launches(wl.launch, 20, dev) / 20
    burst = {"ms_per_launch": burst_ms, "achieved": algo_bytes / burst_ms / 1e6, "frac": algo_bytes / burst_ms / 1e6 / peak,
             "how": "20 back-to-back launches after 1 s idle, CUDA events: no power capping yet"}
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
            tj = json.load(f)
        if tj.get("workload") == args.workload and tj.get("kernel") == kernel:
            traffic, traffic_src = tj["traffic_bytes_per_launch"], tj.get("source")
    except Exception:
        pass

    # ---- e2e: host buffers in, host buffers out, through the same public API, the SAME trajectories per step ----
    e2e_traj = len(trajs)
    e2e_T = w["T"]
    h_vel = [torch.from_numpy(trajs[i][1]).pin_memory() for i in range(e2e_traj)]
    h_prs = [torch.from_numpy(trajs[i][2]).pin_memory() for i in range(e2e_traj)]
    # three streams, two device slots: upload of trajectory i+1 / kernel of i / download of i-1 overlap
    pipe = HostPipeline([plans[trajs[i][0]] for i in range(e2e_traj)], [tables[trajs[i][0]] for i in range(e2e_traj)], pers,
                        n_steps=e2e_T, t0=0, interval=1, n_frames=e2e_T, depth=2)
    h_states = [torch.empty((e2e_T, tab.n_patches, 3, tab.px, tab.py), dtype=torch.float32).pin_memory() for _ in range(e2e_traj)]
    h_mask = [torch.empty((e2e_T, tab.n_patches, tab.px, tab.py), dtype=torch.uint8).pin_memory() for _ in range(e2e_traj)]
    h2d = sum(t.numel() * 4 for t in h_vel) + sum(t.numel() * 4 for t in h_prs)
    d2h = sum(t.numel() * 4 for t in h_states) + sum(t.numel() for t in h_mask)

    def e2e_step():
        pipe.run(h_vel, h_prs, h_states, h_mask)

    def timed_wall(fn, n):
        fn()
        torch.cuda.synchronize(dev)
        barrier(world)
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize(dev)
        return max_over_ranks(time.perf_counter() - t0, dev, world)

    e2e_steps = max(2, min(args.steps, 6))
    e2e_sec = timed_wall(e2e_step, e2e_steps)
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

    dev_sec = timed_wall(e2e_dev_step, e2e_steps)
    e2e_dev_value = world * e2e_traj * e2e_T * e2e_steps / dev_sec
    ceil = copy_ceilings(dev, world)
    # what the links alone allow for these byte counts: outputs dominate host-out, inputs host-in
    frames_step = e2e_traj * e2e_T
    ceil_host_out = ceil["d2h_gbs_all_ranks"] * 1e9 / (d2h / frames_step)
    ceil_gpu_consumer = ceil["h2d_gbs_all_ranks"] * 1e9 / (h2d / frames_step)
    del pipe, h_states, h_mask, h_vel, h_prs
    torch.cuda.empty_cache()

    # ---- the one-off per-mesh step, reported separately (cells located per second) ----
    plans[0].locate()
    torch.cuda.synchronize(dev)
    reps = []
    for _ in range(10):
        t0 = time.perf_counter()
        plans[0].locate()            # four kernels + its own stream sync (the bin-item count is read back)
        reps.append((time.perf_counter() - t0) * 1e3)
    loc_ms = float(np.median(reps))
    locate = {"ms_per_mesh": loc_ms, "cells_per_s": plans[0].nx * plans[0].ny / (loc_ms / 1e3),
              "mesh": f"{plans[0].n_nodes} nodes / {plans[0].n_cells} triangles -> {plans[0].nx}x{plans[0].ny} cells"}

    # ---- the path's only collective: dataset statistics merged over ranks ----
    agg = compute_ds_stats.ds_stats(batch.states[0], batch.mask[0])
    merged = compute_ds_stats.all_reduce_stats(agg)
    torch.cuda.synchronize(dev)
    stats = {"n_state_ch0_all_ranks": float(merged[0, 0].item())}
    if world > 1:
        barrier(world)
        reps_c = 200
        us = time_launches(lambda: compute_ds_stats.all_reduce_stats(agg), reps_c, dev) / reps_c * 1e3
        stats["allgather_us"] = max_over_ranks(us, dev, world)
        stats["allgather_reps"] = reps_c

    config = {"workload": workload_text(args.workload, w, wl.meshes)}
    details = {"grid": f"{plans[0].nx}x{plans[0].ny} cells, {tab.n_bx}x{tab.n_by} patches of 16x16",
               "frames_per_launch_per_gpu": wl.frames_per_launch, "launches_per_step": lps,
               "frames_per_step_per_gpu": wl.frames_per_launch * lps, "parallelism": f"trajectory-sharded x{world}",
               "l2_policy": "inputs+outputs per launch (%.0f MB) exceed the 126 MB L2" % (wl.algo_bytes / 1e6),
               "mesh_seed": "100*rank + 0,1,2,.. (meshes with another patch grid skipped)",
               "field_seed": "1000*rank + mesh index; further trajectories on a mesh = that field set rolled in time"}
    host_trajs, host_meshes = wl.trajs, wl.meshes
    del wl, batch
    torch.cuda.empty_cache()

    extra = {}
    if not args.no_extra:
        for name in ("cylinder", "eagle", "big"):
            if name == args.workload:
                continue
            try:
                extra[name] = run_extra_workload(name, rank, world, dev, peak)
            except Exception as e:      # an extra workload must never cost the headline line
                extra[name] = {"error": f"{type(e).__name__}: {e}"}
                torch.cuda.empty_cache()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        fps, done, dt = cpu_frames_per_s(w, host_meshes, host_trajs, 8.0)
        cores = os.cpu_count() or 1
        nw = min(6, cores)
        per_worker = max(8, int(fps * 2.0))
        fps6, done6, dt6 = cpu_pool_frames_per_s(w, host_meshes, host_trajs, nw, per_worker, 3)
        tri_ms = cpu_trifinder_ms(w, host_meshes)
        cpu = {"value": fps, "unit": "frames/s", "cores": 1, "kind": "port",
               "sample": f"{done} frames of one trajectory of the workload in {dt:.1f} s, one thread (of {cores} host cores); "
                         "trifinder build outside the timed loop",
               "dataloader_workers": {"value": fps6, "unit": "frames/s", "cores": nw,
                                      "sample": f"{done6} frames in {dt6:.1f} s over {nw} worker processes (configs/training1.yaml:73 num_workers: 6)"},
               "trifinder_ms_per_sample": tri_ms,
               "trifinder_note": "the reference rebuilds triangulation + trapezoid map + find_many on every __getitem__ "
                                 "(simple_dataloader.py:181); for a 10-frame training sample that is this many ms on top of 10 frame-times"}

    ds_e2e = None
    if rank == 0 and world == 1 and not args.no_extra:
        try:
            ds_e2e = dataset_e2e(dev)
        except Exception as e:      # never at the cost of the headline line
            ds_e2e = {"error": f"{type(e).__name__}: {e}"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world,
                "steps": args.steps, "warmup": warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config, "details": details,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": algo_bytes, "kernel": kernel,
                             "ms_per_launch": ms / n_launches, "traffic_source": traffic_src, "burst": burst},
                "cpu_baseline": cpu,
                "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "note": f"{e2e_traj} trajectories/step/GPU (the batch of `value`), pinned host buffers both ways, HostPipeline (3 streams, 2 device slots)",
                        "link_ceiling": dict(ceil, host_out_frames_per_s=ceil_host_out, frac_of_ceiling=e2e_value / ceil_host_out,
                                             how="plain pinned cudaMemcpyAsync of 256 MiB, all ranks at once, each direction alone"),
                        "outputs_consumed_on_gpu": {"value": e2e_dev_value, "unit": "frames/s", "d2h_bytes_per_step": int(h_sums.numel() * 4),
                                                    "h2d_ceiling_frames_per_s": ceil_gpu_consumer,
                                                    "frac_of_ceiling": e2e_dev_value / ceil_gpu_consumer,
                                                    "note": "same host inputs every step; states stay in HBM for a GPU consumer (per-channel sums read back)"}},
                "gpu_launches": n_launches,
                "locate_one_off": locate,
                "stats": stats,
                "workloads": extra,
                "dataset_e2e": ds_e2e,
                "clocks": sampler.result()}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    # stdout carries ONE JSON line: NCCL's own chatter (its version line at init, anything NCCL_DEBUG asks for) goes to stderr
    # (NCCL honours NCCL_DEBUG_FILE only above the VERSION level, which is what prints that line)
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="airfoil", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the `workloads` object (cylinder / eagle / big)")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_ours(args, w)


if __name__ == "__main__":
    main()
