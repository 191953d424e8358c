"""Per-frame dynamic meshes: a trajectory whose mesh changes in every frame.

True EAGLE simulations are stored as `pointcloud[T,N,2]`, `triangles[T,F,3]`, `VX/VY/PS[T,N]`
(`/root/reference/max/ds_download/eagle.py:123-144`, `get_data`).  Feeding them to the reference's data path means one
`get_mesh_interpolation` (matplotlib trapezoid map) plus three `to_grid` calls per FRAME
(`src/dataloader/mesh_utils.py:82-106`), then `_pad`, `_patch` and `_normalize` as for the static datasets
(`src/dataloader/simple_dataloader.py:104-152,193-216`).  `fl_dyn_interp_patchify` (csrc/fl_dynamic.cu) does the
whole window in a handful of launches: point location is part of the per-frame loop.

All frames of a window share ONE regular grid.  The reference would compute the grid from each frame's bounding box;
the two agree whenever the frames share their bounding box (EAGLE's domain is fixed), which is checked unless
`extents` is given explicitly.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import FL_FLIP_Y, FL_MASK_AWARE_NORM, FL_NO_NORM, check, load, ptr, stream_ptr
from .field_path import Personality
from .mesh_utils import _grid_axes, default_numpy_semantics
from .simple_dataloader import position_ids

F32 = np.float32


def _dev_tensor(a, dtype, dev):
    t = a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a))
    return t.to(device=dev, dtype=dtype, non_blocking=True).contiguous()


class DynamicTrajectory:
    """Node positions, triangles and node fields of every frame, resident in HBM.

    mesh_pos (T, N, 2) float, cells (T, F, 3) int (any winding), velocity (T, N, 2), pressure (T, N) or (T, N, 1)."""

    def __init__(self, mesh_pos, cells, velocity, pressure, grid_res=238, extents=None, numpy_semantics=None, device=None):
        _lib.require_cuda()
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        with torch.cuda.device(dev):
            self.pos = _dev_tensor(mesh_pos, torch.float32, dev)
            self.cells = _dev_tensor(cells, torch.int32, dev)
            self.velocity = _dev_tensor(velocity, torch.float32, dev)
            prs = _dev_tensor(pressure, torch.float32, dev)
        if self.pos.dim() != 3 or self.pos.shape[2] != 2:
            raise ValueError(f"mesh_pos must be (T, N, 2), got {tuple(self.pos.shape)}")
        T, N = int(self.pos.shape[0]), int(self.pos.shape[1])
        if self.cells.dim() != 3 or self.cells.shape[0] != T or self.cells.shape[2] != 3 or self.cells.shape[1] < 1:
            raise ValueError(f"triangles must be a ({T}, F, 3) int array, but found shape {tuple(self.cells.shape)}")
        if tuple(self.velocity.shape) != (T, N, 2):
            raise ValueError(f"velocity must be ({T}, {N}, 2), got {tuple(self.velocity.shape)}")
        if prs.dim() == 3 and prs.shape[2] == 1:
            prs = prs[:, :, 0].contiguous()
        if tuple(prs.shape) != (T, N):
            raise ValueError(f"pressure must be ({T}, {N}) or ({T}, {N}, 1), got {tuple(prs.shape)}")
        self.pressure = prs
        self.n_steps, self.n_nodes, self.n_cells = T, N, int(self.cells.shape[1])
        self.numpy_semantics = numpy_semantics or default_numpy_semantics()
        if extents is None:
            lo, hi = self.pos.amin(dim=1), self.pos.amax(dim=1)                   # (T, 2) each
            if not (bool((lo == lo[0]).all()) and bool((hi == hi[0]).all())):
                raise ValueError("the frames do not share one bounding box: pass extents=(x_min, x_max, y_min, y_max) "
                                 "to choose the grid")
            lo0, hi0 = lo[0].cpu().numpy(), hi[0].cpu().numpy()
            extents = (float(lo0[0]), float(hi0[0]), float(lo0[1]), float(hi0[1]))   # mesh_utils.py:99-100
        self.extents = tuple(float(F32(e)) for e in extents)
        self.ax, self.ay = _grid_axes(*self.extents, int(grid_res), self.numpy_semantics)
        self.nx, self.ny = len(self.ax), len(self.ay)
        with torch.cuda.device(dev):
            self.ax_d = torch.from_numpy(self.ax).to(dev)
            self.ay_d = torch.from_numpy(self.ay).to(dev)
        self._ws = None

    def patch_grid(self, patch_size, crop_patches=0):
        px, py = int(patch_size[0]), int(patch_size[1])
        n_bx = (self.nx + (-self.nx) % px) // px - 2 * crop_patches
        n_by = (self.ny + (-self.ny) % py) // py - 2 * crop_patches
        if n_bx < 1 or n_by < 1:
            raise ValueError(f"no patches left: grid {self.nx}x{self.ny}, patch {px}x{py}, crop {crop_patches}")
        return n_bx, n_by

    def interp_patchify(self, step_num, seq_len, seq_interval, patch_size, personality: Personality, normalize=True,
                        means=None, stds=None, want_tri=False, force_binned=False):
        """Frames step_num, step_num + interval, ... -> (states (T, L, 3, px, py) f32, mask (T, L, px, py) u8,
        tri (T, L, px, py) i32 or None).  One host synchronisation at the end (the status words)."""
        last = int(step_num) + (int(seq_len) - 1) * int(seq_interval)
        if step_num < 0 or seq_len < 1 or seq_interval < 1 or last >= self.n_steps:
            raise ValueError(f"frames {step_num}..{last} step {seq_interval} outside trajectory of {self.n_steps} steps")
        px, py = int(patch_size[0]), int(patch_size[1])
        n_bx, n_by = self.patch_grid(patch_size, personality.crop_patches)
        L, dev, lib = n_bx * n_by, self.device, load()
        sel = slice(int(step_num), last + 1, int(seq_interval))
        if seq_interval == 1:       # contiguous views, no copy
            pos, cells, vel, prs = self.pos[sel], self.cells[sel], self.velocity[sel], self.pressure[sel]
        else:
            pos, cells, vel, prs = (t[sel].contiguous() for t in (self.pos, self.cells, self.velocity, self.pressure))
        T = int(pos.shape[0])
        flags = (FL_FLIP_Y if personality.flip_y else 0) | (FL_MASK_AWARE_NORM if personality.mask_aware_norm else 0) \
            | (0 if normalize else FL_NO_NORM) | (_lib.FL_FORCE_GATHER if force_binned else 0)
        m = (ctypes.c_float * 3)(*(means if means is not None else personality.means))
        s = (ctypes.c_float * 3)(*(stds if stds is not None else personality.stds))
        with torch.cuda.device(dev):
            states = torch.empty((T, L, 3, px, py), dtype=torch.float32, device=dev)
            mask = torch.empty((T, L, px, py), dtype=torch.uint8, device=dev)
            tri = torch.empty((T, L, px, py), dtype=torch.int32, device=dev) if want_tri else None
            status = torch.empty(2, dtype=torch.int32, device=dev)
            ws_bytes = int(lib.fl_dyn_workspace_bytes(T, self.n_cells, self.nx, self.ny))
            for _ in range(3):
                if self._ws is None or self._ws.numel() < ws_bytes:
                    self._ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
                ws = self._ws
                check(lib.fl_dyn_interp_patchify(ptr(pos), ptr(cells), ptr(vel), ptr(prs), T, self.n_nodes, self.n_cells,
                                                 ptr(self.ax_d), ptr(self.ay_d), self.nx, self.ny, px, py,
                                                 personality.crop_patches, m, s, flags, ptr(states), ptr(mask),
                                                 ptr(tri) if tri is not None else None, ptr(status), ptr(ws), ws.numel(),
                                                 stream_ptr()), "fl_dyn_interp_patchify")
                bad, need = (int(v) for v in status.cpu())
                if bad:
                    raise ValueError("triangles are indices into the points and must be in the range "
                                     f"0 <= i < {self.n_nodes} ({bad} triangles are not)")
                if need <= int(lib.fl_dyn_capacity(T, self.n_cells, self.nx, self.ny, ws.numel())):
                    break
                ws_bytes = ws.numel() * 4      # very uneven meshes: more room for the bins' triangle lists
            else:
                raise MemoryError("fl_dyn_interp_patchify: bin-item store still too small after two retries")
        return states, mask, tri

    def ds_get(self, step_num, seq_len, seq_interval, patch_size, personality: Personality, normalize=True):
        """The 5-tuple of simple_dataloader.py:72-102 for a window of this trajectory."""
        states, mask, _ = self.interp_patchify(step_num, seq_len, seq_interval, patch_size, personality, normalize)
        n_bx, n_by = self.patch_grid(patch_size, personality.crop_patches)
        diffs = states[1:] - states[:-1]
        masks = mask[1:].unsqueeze(2).repeat(1, 1, 3, 1, 1).bool()
        return states[:-1], states[1:], diffs, masks, position_ids(seq_len, n_bx, n_by).to(states.device)


def load_eagle_sim(path, t0=0, window_length=None):
    """`get_data` of max/ds_download/eagle.py:123-144 without the random window: <path>/sim.npz (pointcloud, VX, VY, PS, PG,
    mask) + <path>/triangles.npy -> host arrays (mesh_pos, cells, velocity (T,N,2), pressure (T,N) = PS)."""
    import os
    data = np.load(os.path.join(path, "sim.npz"), mmap_mode="r")
    T = data["pointcloud"].shape[0]
    t1 = T if window_length is None else min(T, t0 + window_length)
    mesh_pos = np.asarray(data["pointcloud"][t0:t1], dtype=F32)
    cells = np.asarray(np.load(os.path.join(path, "triangles.npy"), mmap_mode="r")[t0:t1], dtype=np.int32)
    velocity = np.stack([data["VX"][t0:t1], data["VY"][t0:t1]], axis=-1).astype(F32)
    pressure = np.asarray(data["PS"][t0:t1], dtype=F32)
    return mesh_pos, cells, velocity, pressure
