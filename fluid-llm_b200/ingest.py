"""Parallel ingest of the reference's trajectory pickles: unpickle in worker processes, straight into pinned memory.

The reference reads a whole ~14-18 MB pickle per sample inside a DataLoader worker
(`/root/reference/src/dataloader/simple_dataloader.py:154-164`); its six workers (`configs/training1.yaml:73`) hide that
behind the training step.  The GPU data sets run in the calling process, so an in-process `pickle.load` (8 ms) would bound
them at ~125 samples/s whatever the kernels do.  Here a small pool of processes does the unpickling (and the airfoil node
crop, `airfoil_ds.py:164-183`) and writes the node fields, already in the device pitch, into shared-memory slots that the
parent has page-locked once (`cudaHostRegister`): the parent only issues the asynchronous host -> device copies.

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
from concurrent.futures import ProcessPoolExecutor
from multiprocessing import get_context, shared_memory

import numpy as np
import torch


def _strides(n_nodes):
    ps = (n_nodes + 3) // 4 * 4
    return 2 * ps, ps


def _worker_load(path, shm_name, shm_bytes, airfoil_crop):
    """Runs in a worker process: unpickle `path`, crop if asked, write the padded node fields into the shared-memory slot."""
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
    T, N = vel.shape[0], pos.shape[0]
    vs, ps = _strides(N)
    need = 4 * T * (vs + ps)
    if need > shm_bytes:
        return {"too_small": need}
    shm = shared_memory.SharedMemory(name=shm_name)
    try:
        v = np.ndarray((T, vs), dtype=np.float32, buffer=shm.buf, offset=0)
        p = np.ndarray((T, ps), dtype=np.float32, buffer=shm.buf, offset=4 * T * vs)
        v[:, :2 * N] = vel.reshape(T, 2 * N)
        v[:, 2 * N:] = 0
        p[:, :N] = prs.reshape(T, N)
        p[:, N:] = 0
        del v, p
    finally:
        shm.close()
    return {"mesh_pos": np.ascontiguousarray(pos, dtype=np.float32), "cells": np.ascontiguousarray(cells, dtype=np.int32),
            "T": T, "N": N}


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
    def __init__(self, workers=None, slots=None, slot_bytes=24 << 20, airfoil_crop=False):
        self.workers = workers or max(2, min(8, (os.cpu_count() or 4) - 1))
        self.n_slots = slots or 2 * self.workers
        self.slot_bytes = slot_bytes
        self.airfoil_crop = airfoil_crop
        # fork, like torch's DataLoader workers: the children only unpickle and copy with NumPy, they never touch CUDA
        self._pool = ProcessPoolExecutor(self.workers, mp_context=get_context("fork"))
        self._free, self._all = [], []
        self._pending = OrderedDict()          # path -> (future, slot)
        self._queue = []                       # paths waiting for a free slot
        atexit.register(self.close)

    # -- slots ---------------------------------------------------------------------------------
    def _get_slot(self):
        if self._free:
            return self._free.pop()
        if len(self._all) < self.n_slots:
            s = _Slot(self.slot_bytes)
            self._all.append(s)
            return s
        return None

    def _pump(self):
        while self._queue:
            slot = self._get_slot()
            if slot is None:
                return
            path = self._queue.pop(0)
            fut = self._pool.submit(_worker_load, path, slot.shm.name, slot.nbytes, self.airfoil_crop)
            self._pending[path] = (fut, slot)

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
            if path not in self._queue:
                self._queue.insert(0, path)
            else:                              # jump the queue
                self._queue.remove(path)
                self._queue.insert(0, path)
            self._pump()
            while path not in self._pending:   # every slot is busy with files nobody took yet: wait for none, grow instead
                s = _Slot(self.slot_bytes)
                self._all.append(s)
                self._free.append(s)
                self._pump()
        fut, slot = self._pending.pop(path)
        r = fut.result()
        if "too_small" in r:                   # a trajectory larger than the slots: enlarge them and retry this file
            self._release(slot)
            self.slot_bytes = int(r["too_small"] * 1.25)
            for s in self._free:
                s.close()
                self._all.remove(s)
            self._free = []
            return self.take(path)
        T, N = r["T"], r["N"]
        vs, ps = _strides(N)
        vel = slot.host[: 4 * T * vs].view(torch.float32).view(T, vs)
        prs = slot.host[4 * T * vs: 4 * T * (vs + ps)].view(torch.float32).view(T, ps)
        return r["mesh_pos"], r["cells"], vel, prs, (lambda: self._release(slot))

    def _release(self, slot):
        if slot.nbytes != self.slot_bytes:     # an old, smaller slot
            slot.close()
            if slot in self._all:
                self._all.remove(slot)
        else:
            self._free.append(slot)
        self._pump()

    def close(self):
        pool, self._pool = getattr(self, "_pool", None), None
        if pool is None:
            return
        for fut, _ in self._pending.values():
            fut.cancel()
        pool.shutdown(wait=True, cancel_futures=True)
        for s in self._all:
            try:
                s.close()
            except Exception:
                pass
        self._all, self._free, self._pending, self._queue = [], [], OrderedDict(), []
