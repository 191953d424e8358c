"""Parallel ingest of the reference's trajectory pickles: unpickle in worker processes, straight into pinned memory.

The reference reads a whole ~14-18 MB pickle per sample inside a DataLoader worker
(`/root/reference/src/dataloader/simple_dataloader.py:154-164`); its six workers (`configs/training1.yaml:73`) hide that
behind the training step.  The GPU data sets run in the calling process, so an in-process `pickle.load` (8 ms) would bound
them at ~125 samples/s whatever the kernels do.  Here a small pool of processes does the unpickling (and the airfoil node
crop, `airfoil_ds.py:164-183`) and writes the node fields, already in the device pitch, into shared-memory slots that the
parent has page-locked once (`cudaHostRegister`): the parent only issues the asynchronous host -> device copies.

Every worker owns its own two slots and keeps them mapped for its lifetime: a process that maps a 20 MB slot for the first
time pays ~20 ms of page faults, four times the unpickling itself, so slots are never handed from one worker to another.

    pool = PickleIngest(workers=8)
    pool.submit(paths)                       # runs ahead; order of consumption is free
    pos, cells, vel, prs, release = pool.take(path)     # vel [T, vel_stride], prs [T, prs_stride] pinned views
    ... upload ...; release()

Workers never touch CUDA.
"""
from __future__ import annotations

import atexit
import os
import pickle
from collections import OrderedDict
from multiprocessing import get_context, shared_memory

import numpy as np
import torch

SLOTS_PER_WORKER = 2


def _strides(n_nodes):
    ps = (n_nodes + 3) // 4 * 4
    return 2 * ps, ps


def load_trajectory(path, airfoil_crop):
    """Unpickle `path` and crop if asked -> (mesh_pos f32 [N,2], cells i32 [F,3], velocity [T,N,2], pressure [T,N,1])."""
    with open(path, "rb") as f:
        d = pickle.load(f)
    pos, cells = np.asarray(d["mesh_pos"]), np.asarray(d["cells"])
    vel, prs = np.asarray(d["velocity"]), np.asarray(d["pressure"])
    if airfoil_crop:
        mask = (pos[:, 0] > -.5) & (pos[:, 0] < 2) & (pos[:, 1] > -.75) & (pos[:, 1] < 0.75)      # airfoil_ds.py:166-168
        wanted = np.nonzero(mask)[0]
        renum = np.zeros(len(mask), dtype=np.int64)
        renum[mask] = np.arange(len(wanted), dtype=np.int64)
        cells = renum[cells[np.isin(cells, wanted).all(axis=1)]]
        pos, vel, prs = pos[mask], vel[:, mask], prs[:, mask]
    return np.ascontiguousarray(pos, dtype=np.float32), np.ascontiguousarray(cells, dtype=np.int32), vel, prs


def _fill_slot(buf, nbytes, vel, prs):
    """Write the node fields in the device pitch (frames padded to 4 nodes, pad = 0) into `buf`; -> None, or the bytes needed."""
    T, N = vel.shape[0], vel.shape[1]
    vs, ps = _strides(N)
    need = 4 * T * (vs + ps)
    if need > nbytes:
        return need
    v = np.ndarray((T, vs), dtype=np.float32, buffer=buf, offset=0)
    p = np.ndarray((T, ps), dtype=np.float32, buffer=buf, offset=4 * T * vs)
    v[:, :2 * N] = vel.reshape(T, 2 * N)
    v[:, 2 * N:] = 0
    p[:, :N] = prs.reshape(T, N)
    p[:, N:] = 0
    return None


def _worker_main(in_q, out_q, slot_names, slot_bytes, airfoil_crop):
    shms = [shared_memory.SharedMemory(name=n) for n in slot_names]       # mapped once, for the life of the worker
    try:
        while True:
            job = in_q.get()
            if job is None:
                break
            ticket, path, si = job
            try:
                pos, cells, vel, prs = load_trajectory(path, airfoil_crop)
                need = _fill_slot(shms[si].buf, slot_bytes, vel, prs)
                r = {"too_small": need} if need else {"mesh_pos": pos, "cells": cells, "T": vel.shape[0], "N": pos.shape[0]}
            except Exception as e:      # noqa: BLE001 -- reported to the parent, which re-raises
                r = {"error": f"{type(e).__name__}: {e}"}
            out_q.put((ticket, r))
    finally:
        for s in shms:
            s.close()


class _Slot:
    def __init__(self, nbytes):
        self.shm = shared_memory.SharedMemory(create=True, size=nbytes)
        self.nbytes = nbytes
        self.host = torch.frombuffer(self.shm.buf, dtype=torch.uint8)
        self.registered = False
        if torch.cuda.is_available():
            rc = torch.cuda.cudart().cudaHostRegister(self.host.data_ptr(), nbytes, 0)
            self.registered = int(rc) == 0

    def close(self):
        if self.registered:
            torch.cuda.cudart().cudaHostUnregister(self.host.data_ptr())
        self.host = None
        self.shm.close()
        self.shm.unlink()


class PickleIngest:
    def __init__(self, workers=None, slot_bytes=24 << 20, airfoil_crop=False):
        self.workers = workers or max(2, min(12, (os.cpu_count() or 4) - 2))
        self.slot_bytes = slot_bytes
        self.airfoil_crop = airfoil_crop
        # fork, like torch's DataLoader workers: the children only unpickle and copy with NumPy, they never touch CUDA
        ctx = get_context("fork")
        self._out_q = ctx.Queue()
        self._slots = [[_Slot(slot_bytes) for _ in range(SLOTS_PER_WORKER)] for _ in range(self.workers)]
        self._in_qs, self._procs = [], []
        for w in range(self.workers):
            q = ctx.Queue()
            p = ctx.Process(target=_worker_main, daemon=True,
                            args=(q, self._out_q, [s.shm.name for s in self._slots[w]], slot_bytes, airfoil_crop))
            p.start()
            self._in_qs.append(q)
            self._procs.append(p)
        self._free = [(w, si) for si in range(SLOTS_PER_WORKER) for w in range(self.workers)]     # round-robin over the workers
        self._pending = OrderedDict()          # path -> ticket
        self._where = {}                       # ticket -> (worker, slot index)
        self._done = {}                        # ticket -> result (arrived, not yet taken)
        self._queue = []                       # paths waiting for a free slot
        self._ticket = 0
        atexit.register(self.close)

    def _pump(self):
        while self._queue and self._free:
            path = self._queue.pop(0)
            w, si = self._free.pop(0)
            self._ticket += 1
            self._pending[path] = self._ticket
            self._where[self._ticket] = (w, si)
            self._in_qs[w].put((self._ticket, path, si))

    # -- API -----------------------------------------------------------------------------------
    def submit(self, paths):
        """Start loading these files (those not already on their way); returns at once."""
        for p in paths:
            if p not in self._pending and p not in self._queue:
                self._queue.append(p)
        self._pump()

    def take(self, path):
        """-> (mesh_pos, cells, vel [T, vel_stride], prs [T, prs_stride], release): the two field tensors are views of a
        page-locked slot; call release() once the copies out of them have completed."""
        if path not in self._pending:
            if path in self._queue:            # jump the queue
                self._queue.remove(path)
            self._queue.insert(0, path)
            self._pump()
            if path not in self._pending:      # every slot holds a file nobody took yet (or whose upload is still in flight)
                self._queue.remove(path)
                return self._load_here(path)
        ticket = self._pending.pop(path)
        while ticket not in self._done:
            t, r = self._out_q.get()
            self._done[t] = r
        r = self._done.pop(ticket)
        w, si = self._where.pop(ticket)
        slot = self._slots[w][si]
        if "error" in r or "too_small" in r:
            self._release(w, si)
            if "error" in r:
                raise RuntimeError(f"ingest worker failed on {path}: {r['error']}")
            return self._load_here(path)       # a trajectory larger than the slots
        T, N = r["T"], r["N"]
        vs, ps = _strides(N)
        vel = slot.host[: 4 * T * vs].view(torch.float32).view(T, vs)
        prs = slot.host[4 * T * vs: 4 * T * (vs + ps)].view(torch.float32).view(T, ps)
        return r["mesh_pos"], r["cells"], vel, prs, (lambda: self._release(w, si))

    def _load_here(self, path):
        """In the calling process, into fresh pinned memory (no slot was free, or the file does not fit one)."""
        pos, cells, vel, prs = load_trajectory(path, self.airfoil_crop)
        T, N = vel.shape[0], pos.shape[0]
        vs, ps = _strides(N)
        host = torch.empty(4 * T * (vs + ps), dtype=torch.uint8, pin_memory=torch.cuda.is_available())
        _fill_slot(memoryview(host.numpy()), host.numel(), vel, prs)
        v = host[: 4 * T * vs].view(torch.float32).view(T, vs)
        p = host[4 * T * vs:].view(torch.float32).view(T, ps)
        return pos, cells, v, p, (lambda: None)

    def _release(self, w, si):
        self._free.append((w, si))
        self._pump()

    def close(self):
        procs, self._procs = getattr(self, "_procs", None), None
        if not procs:
            return
        for q in self._in_qs:
            try:
                q.put(None)
            except Exception:
                pass
        for p in procs:
            p.join(timeout=2)
            if p.is_alive():
                p.terminate()
        for ws in self._slots:
            for s in ws:
                try:
                    s.close()
                except Exception:
                    pass
        self._slots, self._free, self._pending, self._queue = [], [], OrderedDict(), []
