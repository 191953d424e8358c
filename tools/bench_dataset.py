"""Dataset-level throughput: samples/s through `MGNDataset.ds_get` -- file read, mesh plan, upload,
fused kernel, 5-tuple -- from the reference's pickles and from the flat `.fgt` files, cold (every sample loads its file)
and warm (trajectory resident on the device), beside the oracle's `ds_get` on one host core.

    python tools/bench_dataset.py [--files 12] [--samples 48] [--seq-len 10]
"""
import argparse
import os
import pickle
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--files", type=int, default=12)
    ap.add_argument("--samples", type=int, default=96)
    ap.add_argument("--seq-len", type=int, default=10)
    args = ap.parse_args()
    from fluid_llm_b200 import synth
    from fluid_llm_b200.simple_dataloader import MGNDataset
    from fluid_llm_b200.traj_store import convert_pickle
    from oracle import pipeline as P
    tmp = tempfile.mkdtemp(prefix="fluidgrid_ds_")
    d_pkl, d_fgt = os.path.join(tmp, "pkl"), os.path.join(tmp, "fgt")
    os.makedirs(d_pkl), os.makedirs(d_fgt)
    size = 0
    for i in range(args.files):
        tr = synth.make_trajectory("cylinder", 600, mesh_seed=i, field_seed=100 + i)
        p = os.path.join(d_pkl, f"save_{i:03d}.pkl")
        with open(p, "wb") as f:
            pickle.dump(tr, f)
        size = os.path.getsize(p)
        convert_pickle(p, os.path.join(d_fgt, f"save_{i:03d}.fgt"))
    print(f"{args.files} cylinder-shaped trajectories, T=600, {size / 1e6:.1f} MB per pickle; sample = seq_len {args.seq_len} frames "
          f"at a random start (page cache warm: the files were just written)")
    rng = np.random.default_rng(0)
    order = [(int(rng.integers(args.files)), int(rng.integers(0, 500))) for _ in range(args.samples)]

    def run(ds, label):
        ds.ds_get(0, 0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for (fi, step) in order:
            out = ds.ds_get(fi, step)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"{label:58s} {args.samples / dt:9.1f} samples/s  {args.samples * args.seq_len / dt:10.1f} frames/s  {dt / args.samples * 1e3:8.2f} ms/sample")
        return out

    def fresh(d, **attrs):                      # one data set object per mode: no allocator / cache state carried over
        ds = MGNDataset(d, 238, (16, 16), (16, 16), args.seq_len, mode="valid")
        for k, v in attrs.items():
            setattr(ds, k, v)
        ds._cache.clear()
        return ds

    for d, name in ((d_pkl, "pickle"), (d_fgt, ".fgt  ")):
        run(fresh(d, cache_size=0, window_loads=False, ingest_workers=0), f"{name} cold (file -> plan -> upload -> kernel per sample)")
        if name == "pickle":
            # the ingest pool: worker processes unpickle ahead into page-locked slots, the parent only copies and launches
            d_cold = os.path.join(tmp, "cold")       # one path per sample (hard links), as in a data set of many files: the pool
            os.makedirs(d_cold)                      # loads a path once at a time, and 96 samples over 12 paths would queue up
            for k, (fi, _) in enumerate(order):
                os.link(os.path.join(d, f"save_{fi:03d}.pkl"), os.path.join(d_cold, f"save_{k:03d}.pkl"))
            cold_order = [(k, step) for k, (_, step) in enumerate(order)]
            dsp = fresh(d_cold, cache_size=0, window_loads=False)
            dsp.ds_get(0, 0)

            def cold_pool(batch):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for i in range(0, len(cold_order), batch):
                    dsp.prefetch(cold_order[i:i + max(3 * batch, 2 * dsp._ingest.workers)])      # what num_workers x prefetch_factor does (utils_model.LookaheadBatchSampler)
                    if batch == 1:
                        dsp.ds_get(*cold_order[i])
                    else:
                        dsp.ds_get_many(cold_order[i:i + batch])
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                lab = f"{name} cold, ingest pool ({dsp._ingest.workers} processes), " + ("one sample per call" if batch == 1 else f"batches of {batch}")
                print(f"{lab:58s} {args.samples / dt:9.1f} samples/s  {args.samples * args.seq_len / dt:10.1f} frames/s  {dt / args.samples * 1e3:8.2f} ms/sample")
                tt = dsp.ingest_times
                print(f"    parent, per sample: waiting for the pool {tt['wait'] / tt['n'] * 1e3:.2f} ms, mesh plan {tt['plan'] / tt['n'] * 1e3:.2f} ms, "
                      f"upload calls {tt['upload'] / tt['n'] * 1e3:.2f} ms (of {dt / args.samples * 1e3:.2f} ms)")
                for k in tt:
                    tt[k] = 0
            for batch in (1, 8, 1, 8):               # (the first pass of each form also warms the allocators up)
                cold_pool(batch)
            dsp._ingest.close()
        if name.strip() == ".fgt":
            dsw = fresh(d, cache_size=0, window_loads=True)
            run(dsw, f"{name} window loads, first pass (12 mesh plans built on the way)")
            run(dsw, f"{name} window loads, plans cached ({args.seq_len} frames read per sample)")
        ds = fresh(d, cache_size=args.files)
        for i in range(args.files):
            ds.ds_get(i, 0)
        run(ds, f"{name} warm (trajectory and plan resident on the device)")
        if name.strip() == ".fgt":
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(0, len(order), 8):
                ds.ds_get_many(order[i:i + 8])
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            print(f"{name + ' warm, batches of 8 in one launch (ds_get_many)':58s} {args.samples / dt:9.1f} samples/s  "
                  f"{args.samples * args.seq_len / dt:10.1f} frames/s  {dt / args.samples * 1e3:8.2f} ms/sample")
    # the oracle: what one DataLoader worker of the reference does per sample
    n = max(2, args.samples // 8)
    t0 = time.perf_counter()
    for (fi, step) in order[:n]:
        with open(os.path.join(d_pkl, f"save_{fi:03d}.pkl"), "rb") as f:
            tr = pickle.load(f)
        P.ds_get(tr, step, args.seq_len, 1)
    dt = time.perf_counter() - t0
    print(f"{'oracle (unpickle + trapezoid map + 3 x to_grid per frame), 1 core':58s} {n / dt:9.1f} samples/s  {n * args.seq_len / dt:10.1f} frames/s  "
          f"{dt / n * 1e3:8.2f} ms/sample")


if __name__ == "__main__":
    main()
