"""Parallel ingest of the reference's trajectory pickles: unpickle in worker processes, straight into pinned memory.

The reference reads a whole ~14-18 MB pickle per sample inside a DataLoader worker
(`/root/reference/src/dataloader/simple_dataloader.py:154-164`); its six workers (`configs/training1.yaml:73`) hide that
behind the training step.  The GPU data sets run in the calling process, so an in-process `pickle.load` (8 ms) would bound
them at ~125 samples/s whatever the kernels do.  Here a small pool of processes does the unpickling (and the airfoil node
crop, `airfoil_ds.py:164-183`) and writes the node fields, already in the device pitch, into shared-memory slots that the
parent has page-locked once (`cudaHostRegister`): the parent only issues the asynchronous host -> device copies.

Every worker owns its own two slots and keeps them mapped for its lifetime: a process that maps a 20 MB slot for the first
time pays ~20 ms of page faults, four times the unpickling itself, so slots are never handed from one worker to another.
The workers are separate interpreters (`_ingest_worker.py`, started with fork+exec), not forks of the calling process.

    pool = PickleIngest(workers=8)
    pool.submit(paths)                       # runs ahead; order of consumption is free
    pos, cells, vel, prs, release, plan = pool.take(path)     # vel [T, vel_stride], prs [T, prs_stride] pinned views
    ... upload ...; release()

Workers never touch CUDA.
"""
from __future__ import annotations

import atexit
import os
from collections import OrderedDict
import subprocess
import sys
from multiprocessing import Pipe, shared_memory
from multiprocessing.connection import wait as conn_wait

import torch

from ._ingest_worker import fill_slot as _fill_slot, load_trajectory, strides as _strides

SLOTS_PER_WORKER = 2
WORKER_SCRIPT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ingest_worker.py")


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
    def __init__(self, workers=None, slot_bytes=24 << 20, airfoil_crop=False, plan_resolution=0, numpy_semantics="1.26"):
        """`plan_resolution` > 0: the workers also prepare the host side of each trajectory's mesh plan for that grid resolution
        (`_plan_host.prepare_plan`); `take` then returns it as its sixth value (else None)."""
        self.workers = workers or max(2, min(12, (os.cpu_count() or 4) - 2))
        self.slot_bytes = slot_bytes
        self.airfoil_crop = airfoil_crop
        self._slots = [[_Slot(slot_bytes) for _ in range(SLOTS_PER_WORKER)] for _ in range(self.workers)]
        # workers are separate interpreters started with fork+exec (see _ingest_worker.py for why they are not forks of this
        # process); each talks to the parent over its own duplex pipe
        self._conns, self._procs = [], []
        for w in range(self.workers):
            parent_c, child_c = Pipe(duplex=True)
            fd = child_c.fileno()
            os.set_inheritable(fd, True)
            p = subprocess.Popen([sys.executable, WORKER_SCRIPT, str(fd), str(slot_bytes), "1" if airfoil_crop else "0",
                                  str(int(plan_resolution)), str(numpy_semantics), *[s.shm.name for s in self._slots[w]]],
                                 pass_fds=(fd,), close_fds=True)
            child_c.close()
            self._conns.append(parent_c)
            self._procs.append(p)
        self._free = [(w, si) for si in range(SLOTS_PER_WORKER) for w in range(self.workers)]     # round-robin over the workers
        self._pending = OrderedDict()          # path -> ticket
        self._where = {}                       # ticket -> (worker, slot index)
        self._done = {}                        # ticket -> result (arrived, not yet taken)
        self._queue = []                       # paths waiting for a free slot
        self._ticket = 0
        self.grow_to = 0                       # > 0: a file did not fit a slot; the owner should rebuild the pool with slots this large
        atexit.register(self.close)

    def _pump(self):
        while self._queue and self._free:
            path = self._queue.pop(0)
            w, si = self._free.pop(0)
            self._ticket += 1
            self._pending[path] = self._ticket
            self._where[self._ticket] = (w, si)
            self._conns[w].send((self._ticket, path, si))

    # -- API -----------------------------------------------------------------------------------
    def submit(self, paths):
        """Start loading these files (those not already on their way); returns at once."""
        for p in paths:
            if p not in self._pending and p not in self._queue:
                self._queue.append(p)
        self._pump()

    def take(self, path):
        """-> (mesh_pos, cells, vel [T, vel_stride], prs [T, prs_stride], release, plan): the two field tensors are views of a
        page-locked slot; call release() once the copies out of them have completed.  `plan`: the prepared host side of the
        mesh plan (see __init__), or None."""
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
            for c in conn_wait(self._conns):
                try:
                    t, r = c.recv()
                except EOFError:
                    raise RuntimeError("an ingest worker process died") from None
                self._done[t] = r
        r = self._done.pop(ticket)
        w, si = self._where.pop(ticket)
        slot = self._slots[w][si]
        if "error" in r or "too_small" in r:
            self._release(w, si)
            if "error" in r:
                raise RuntimeError(f"ingest worker failed on {path}: {r['error']}")
            self.grow_to = max(self.grow_to, (int(r["too_small"] * 1.25) + (1 << 20) - 1) >> 20 << 20)
            return self._load_here(path)       # a trajectory larger than the slots
        T, N = r["T"], r["N"]
        vs, ps = _strides(N)
        vel = slot.host[: 4 * T * vs].view(torch.float32).view(T, vs)
        prs = slot.host[4 * T * vs: 4 * T * (vs + ps)].view(torch.float32).view(T, ps)
        return r["mesh_pos"], r["cells"], vel, prs, (lambda: self._release(w, si)), r.get("plan")

    def _load_here(self, path):
        """In the calling process, into fresh pinned memory (no slot was free, or the file does not fit one)."""
        pos, cells, vel, prs = load_trajectory(path, self.airfoil_crop)
        T, N = vel.shape[0], pos.shape[0]
        vs, ps = _strides(N)
        host = torch.empty(4 * T * (vs + ps), dtype=torch.uint8, pin_memory=torch.cuda.is_available())
        _fill_slot(memoryview(host.numpy()), host.numel(), vel, prs)
        v = host[: 4 * T * vs].view(torch.float32).view(T, vs)
        p = host[4 * T * vs:].view(torch.float32).view(T, ps)
        return pos, cells, v, p, (lambda: None), None

    def _release(self, w, si):
        self._free.append((w, si))
        self._pump()

    def close(self):
        procs, self._procs = getattr(self, "_procs", None), None
        if not procs:
            return
        for c in self._conns:
            try:
                c.send(None)
            except Exception:
                pass
        for p in procs:
            try:
                p.wait(timeout=2)
            except Exception:
                p.kill()
        for c in self._conns:
            c.close()
        for ws in self._slots:
            for s in ws:
                try:
                    s.close()
                except Exception:
                    pass
        self._slots, self._free, self._pending, self._queue = [], [], OrderedDict(), []
