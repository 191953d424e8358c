"""Worker process of ingest.PickleIngest: unpickles trajectory files into shared-memory slots.

Run as a script by `subprocess.Popen` (NOT forked from the caller: a fork of a process with tens of GB of mapped memory
leaves the parent paying copy-on-write faults for as long as the children live), so it imports nothing but NumPy:

    python _ingest_worker.py <fd of the duplex pipe to the parent> <slot bytes> <airfoil crop 0/1> <slot name> [<slot name> ...]

Protocol (pickled tuples over a multiprocessing Connection): parent -> (ticket, path, slot index) or None to stop;
worker -> (ticket, result dict).
"""
import pickle
import sys
from multiprocessing import shared_memory
from multiprocessing.connection import Connection

import numpy as np


def strides(n_nodes):
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
    slot_bytes, airfoil_crop = int(argv[2]), argv[3] == "1"
    # mapped once, for the life of the worker: a fresh mapping of a 20 MB slot costs ~20 ms of page faults.  The parent owns
    # (and unlinks) the segments: keep this process's resource tracker out of it.
    shms = []
    for name in argv[4:]:
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
            except Exception as e:      # noqa: BLE001 -- reported to the parent, which re-raises
                r = {"error": f"{type(e).__name__}: {e}"}
            conn.send((ticket, r))
    finally:
        for s in shms:
            s.close()


if __name__ == "__main__":
    main(sys.argv)
