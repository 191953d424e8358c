"""Worker process of ingest.PickleIngest: unpickles trajectory files into shared-memory slots.

Run as a script by `subprocess.Popen` (NOT forked from the caller: a fork of a process with tens of GB of mapped memory
leaves the parent paying copy-on-write faults for as long as the children live), so it imports nothing but NumPy:

    python _ingest_worker.py <fd of the duplex pipe to the parent> <slot bytes> <airfoil crop 0/1> <grid resolution, 0 = no plan>
                             <numpy semantics> <slot name> [<slot name> ...]

Protocol (pickled tuples over a multiprocessing Connection): parent -> (ticket, path, slot index) or None to stop;
worker -> (ticket, result dict).  With a grid resolution the result also carries the host side of the trajectory's mesh plan
(`_plan_host.prepare_plan`: validation, grid axes, Morton slots), so the parent only uploads it.
"""
import pickle
import sys
from multiprocessing import shared_memory
from multiprocessing.connection import Connection

import numpy as np

try:
    from . import _plan_host
except ImportError:          # run as a script (by ingest.PickleIngest): the module lies next to this file
    import _plan_host


def strides(n_nodes):
    ps = (n_nodes + 3) // 4 * 4
    return 2 * ps, ps


_declined = 0             # consecutive files the lazy reader gave up on
LAZY_BYTES = 1 << 18      # payloads of at least this many bytes are not copied out of the file while unpickling


class _LazyFile:
    """File object over a memory-mapped pickle for `pickle.Unpickler` that SKIPS large byte-string payloads.

    The C unpickler reads a BINBYTES payload straight into the bytes object it has just allocated (`readinto`), and NumPy's
    `__setstate__` keeps pointing into that bytes object when it is large and aligned.  For payloads >= LAZY_BYTES this file
    only notes (address of the target, offset in the file, length) and copies nothing: the unpickled array then has the right
    shape and dtype over memory nobody ever touched (no page is faulted in), and `lazy_arrays` swaps it for a view of the
    mapped file.  A 600-step trajectory is ~18 MB of such payloads, a third of which (density) nobody reads."""

    def __init__(self, mm):
        self.mm, self.pos, self.lazy = mm, 0, {}

    def read(self, n=-1):
        end = len(self.mm) if n is None or n < 0 else min(self.pos + n, len(self.mm))
        out = self.mm[self.pos:end]
        self.pos = end
        return out

    def readline(self):
        end = self.mm.find(b"\n", self.pos)
        end = len(self.mm) if end < 0 else end + 1
        out = self.mm[self.pos:end]
        self.pos = end
        return out

    def readinto(self, b):
        n = min(len(b), len(self.mm) - self.pos)
        if n >= LAZY_BYTES and n == len(b):
            self.lazy[np.frombuffer(b, dtype=np.uint8).ctypes.data] = (self.pos, n)
        else:
            b[:n] = self.mm[self.pos:self.pos + n]
        self.pos += n
        return n


class _ArraysOnly(pickle.Unpickler):
    """Only NumPy's own array / dtype / scalar reconstruction may run over a _LazyFile: nothing of it looks INTO a large payload
    while unpickling.  Any other class (a tensor rebuilt from its storage bytes, say) would be handed untouched memory, so its
    pickle is refused before that happens and goes to the plain unpickler."""
    _ALLOWED = {("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"), ("numpy", "ndarray"),
                ("numpy", "dtype"), ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar"),
                ("numpy.core.numeric", "_frombuffer"), ("numpy._core.numeric", "_frombuffer")}

    def find_class(self, module, name):
        if (module, name) not in self._ALLOWED:
            raise TypeError(f"{module}.{name} in the pickle")
        return super().find_class(module, name)


def unpickle_lazy(path, keys):
    """-> {key: array} with the large arrays of `keys` as read-only views of the mapped file (which lives as long as they do), or
    None when the file is not a plain dict of NumPy arrays unpickled the way _LazyFile expects (the caller then unpickles
    normally).  Arrays of other keys may point at untouched memory and are dropped unread."""
    import mmap
    global _declined
    if _declined >= 2:          # these files are not written the way the lazy reader needs: stop paying for the attempt
        return None
    try:
        with open(path, "rb") as f:
            mm = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
        lf = _LazyFile(mm)
        d = _ArraysOnly(lf).load()
        if not isinstance(d, dict):
            raise TypeError("not a dict")
        out = {}
        for k in keys:
            a = d[k]
            if not isinstance(a, np.ndarray) or a.dtype.hasobject:
                raise TypeError("not a plain array")
            if a.nbytes >= LAZY_BYTES:
                where = lf.lazy.get(a.ctypes.data)
                if where is None or where[1] != a.nbytes or not a.flags.c_contiguous:
                    raise TypeError("payload not where expected")       # (NumPy copied it, or a view of a larger buffer)
                a = np.ndarray(a.shape, dtype=a.dtype, buffer=mm, offset=where[0])
            out[k] = a
        _declined = 0
        return out
    except TypeError:       # not the kind of pickle this reader handles
        _declined += 1
        return None
    except Exception:       # noqa: BLE001 -- a missing / truncated / foreign file: the plain unpickler decides (and raises)
        return None


def load_trajectory(path, airfoil_crop):
    """Unpickle `path` and crop if asked -> (mesh_pos f32 [N,2], cells i32 [F,3], velocity [T,N,2], pressure [T,N,1]); the two
    field arrays may be views of the mapped file (they are only read once more, by fill_slot)."""
    got = unpickle_lazy(path, ("mesh_pos", "cells", "velocity", "pressure"))
    if got is None:
        with open(path, "rb") as f:
            d = pickle.load(f)
    else:
        d = got
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


def fill_slot(buf, nbytes, vel, prs):
    """Write the node fields in the device pitch (frames padded to 4 nodes, pad = 0) into `buf`; -> None, or the bytes needed."""
    T, N = vel.shape[0], vel.shape[1]
    vs, ps = strides(N)
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


def main(argv):
    conn = Connection(int(argv[1]))
    slot_bytes, airfoil_crop, plan_res, plan_sem = int(argv[2]), argv[3] == "1", int(argv[4]), argv[5]
    # mapped once, for the life of the worker: a fresh mapping of a 20 MB slot costs ~20 ms of page faults.  The parent owns
    # (and unlinks) the segments: keep this process's resource tracker out of it.
    shms = []
    for name in argv[6:]:
        s = shared_memory.SharedMemory(name=name)
        try:
            from multiprocessing import resource_tracker
            resource_tracker.unregister(s._name, "shared_memory")
        except Exception:
            pass
        shms.append(s)
    try:
        while True:
            try:
                job = conn.recv()
            except EOFError:
                break
            if job is None:
                break
            ticket, path, si = job
            try:
                pos, cells, vel, prs = load_trajectory(path, airfoil_crop)
                need = fill_slot(shms[si].buf, slot_bytes, vel, prs)
                r = {"too_small": need} if need else {"mesh_pos": pos, "cells": cells, "T": vel.shape[0], "N": pos.shape[0]}
                if plan_res > 0 and not need:
                    try:
                        r["plan"] = _plan_host.prepare_plan(pos, cells, plan_res, plan_sem)
                    except ValueError:       # an invalid triangulation: the parent validates again and raises it properly
                        pass
            except Exception as e:      # noqa: BLE001 -- reported to the parent, which re-raises
                r = {"error": f"{type(e).__name__}: {e}"}
            conn.send((ticket, r))
    finally:
        for s in shms:
            s.close()


if __name__ == "__main__":
    main(sys.argv)
