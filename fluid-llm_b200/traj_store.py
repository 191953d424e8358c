"""Flat on-disk trajectory format (`.fgt`) and loader: SURVEY.md 8(f) rank 1.

Once the kernels run near memory speed the reference's input side dominates: every `__getitem__` unpickles a whole
~14 MB trajectory (`/root/reference/src/dataloader/simple_dataloader.py:154-164`, pickles written by
`max/ds_download/torch_MGN.py:68-95`).  A `.fgt` file holds the same arrays as that pickle, already in the layout the
device wants: node fields padded to the kernels' frame pitch (prs_stride = roundup(N, 4), vel_stride = 2 * prs_stride),
every array at a 4096-byte aligned offset, so loading is `np.memmap` -> one pinned staging copy -> async H2D, with no
parsing, no per-load padding and (for the airfoil set) no per-load node crop.

Layout: 16-byte magic/version, 4-byte little-endian header length, JSON header (array name -> dtype, shape, offset),
then the arrays.  Arrays: mesh_pos f32[N,2], cells i32[F,3], velocity f32[T, vel_stride], pressure f32[T, prs_stride].
"""
from __future__ import annotations

import json
import os
import pickle

import numpy as np
import torch

MAGIC = b"FLUIDGRID-FGT-1\n"
ALIGN = 4096


def _strides(n_nodes):
    ps = (n_nodes + 3) // 4 * 4
    return 2 * ps, ps


def write_fgt(path, mesh_pos, cells, velocity, pressure, meta=None):
    """Write one trajectory.  velocity [T,N,2], pressure [T,N,1] or [T,N] (float32-convertible)."""
    mesh_pos = np.ascontiguousarray(mesh_pos, dtype=np.float32)
    cells = np.ascontiguousarray(cells, dtype=np.int32)
    velocity = np.asarray(velocity, dtype=np.float32)
    pressure = np.asarray(pressure, dtype=np.float32).reshape(velocity.shape[0], -1)
    T, N = velocity.shape[0], mesh_pos.shape[0]
    if velocity.shape != (T, N, 2) or pressure.shape != (T, N) or cells.ndim != 2 or cells.shape[1] != 3:
        raise ValueError("write_fgt: expected mesh_pos [N,2], cells [F,3], velocity [T,N,2], pressure [T,N(,1)]")
    vs, ps = _strides(N)
    vel = np.zeros((T, vs), dtype=np.float32)
    vel[:, :2 * N] = velocity.reshape(T, 2 * N)
    prs = np.zeros((T, ps), dtype=np.float32)
    prs[:, :N] = pressure
    arrays = {"mesh_pos": mesh_pos, "cells": cells, "velocity": vel, "pressure": prs}
    header = {"n_nodes": N, "n_cells": int(cells.shape[0]), "n_steps": T, "vel_stride": vs, "prs_stride": ps,
              "meta": meta or {}, "arrays": {}}
    # two passes: the header length moves the first offset
    off = 0
    for _ in range(2):
        hdr_bytes = json.dumps(header).encode()
        off = (len(MAGIC) + 4 + len(hdr_bytes) + ALIGN - 1) // ALIGN * ALIGN
        for name, a in arrays.items():
            header["arrays"][name] = {"dtype": str(a.dtype), "shape": list(a.shape), "offset": off}
            off = (off + a.nbytes + ALIGN - 1) // ALIGN * ALIGN
    hdr_bytes = json.dumps(header).encode()
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(MAGIC)
        f.write(len(hdr_bytes).to_bytes(4, "little"))
        f.write(hdr_bytes)
        for name, a in arrays.items():
            f.seek(header["arrays"][name]["offset"])
            f.write(a.tobytes())
        f.truncate(off)
    os.replace(tmp, path)
    return header


def convert_pickle(pkl_path, out_path, airfoil_crop=False):
    """One reference pickle -> one .fgt file; `airfoil_crop` applies airfoil_ds.py:164-183 once, at conversion time."""
    with open(pkl_path, "rb") as f:
        d = pickle.load(f)
    pos, cells, vel, prs = d["mesh_pos"], d["cells"], d["velocity"], d["pressure"]
    if airfoil_crop:
        from .airfoil_ds import crop_airfoil_mesh
        m, pos, cells = crop_airfoil_mesh(pos, cells)
        vel, prs = vel[:, m], prs[:, m]
    return write_fgt(out_path, pos, cells, vel, prs, {"source": os.path.basename(pkl_path), "airfoil_crop": bool(airfoil_crop)})


class TrajectoryFile:
    """Memory-mapped view of a .fgt file (nothing is read until an array is touched)."""

    def __init__(self, path):
        self.path = path
        with open(path, "rb") as f:
            if f.read(len(MAGIC)) != MAGIC:
                raise ValueError(f"{path}: not a fluidgrid .fgt file")
            n = int.from_bytes(f.read(4), "little")
            self.header = json.loads(f.read(n).decode())
        h = self.header
        self.n_nodes, self.n_cells, self.n_steps = h["n_nodes"], h["n_cells"], h["n_steps"]
        self.vel_stride, self.prs_stride = h["vel_stride"], h["prs_stride"]

    def array(self, name):
        a = self.header["arrays"][name]
        return np.memmap(self.path, mode="r", dtype=np.dtype(a["dtype"]), shape=tuple(a["shape"]), offset=a["offset"])

    @property
    def mesh_pos(self):
        return np.asarray(self.array("mesh_pos"))

    @property
    def cells(self):
        return np.asarray(self.array("cells"))

    def to_device(self, plan, stream=None, pinned=None, first=0, last=None):
        """-> DeviceTrajectory with the node fields copied file -> pinned host -> HBM (async on `stream`).
        `first` / `last` select a window of time steps [first, last): only those rows are read from the file."""
        from .field_path import DeviceTrajectory
        if plan.n_nodes != self.n_nodes:
            raise ValueError(f"{self.path}: {self.n_nodes} nodes, plan has {plan.n_nodes}")
        last = self.n_steps if last is None else last
        if not (0 <= first < last <= self.n_steps):
            raise ValueError(f"{self.path}: window [{first}, {last}) outside 0..{self.n_steps}")
        return DeviceTrajectory.from_padded(self.array("velocity")[first:last], self.array("pressure")[first:last], plan, stream,
                                            pinned)


class PinnedStage:
    """Reusable pinned host buffers for the file -> device hop."""

    def __init__(self):
        self._buf = {}

    def get(self, key, shape):
        n = int(np.prod(shape))
        b = self._buf.get(key)
        if b is None or b.numel() < n:
            b = torch.empty(n, dtype=torch.float32).pin_memory()
            self._buf[key] = b
        return b[:n].view(*shape)
