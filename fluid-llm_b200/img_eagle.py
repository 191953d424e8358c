"""`grid2mesh`: grid -> node nearest-cell resample on the GPU.

Mirror of `/root/reference/eagle/Dataloader/IMG_Eagle.py:93-123` (same constants, same float32
index arithmetic as NumPy 1.26 evaluates it, same row flip and negative-index wrap).  Returns CPU
tensors like the reference when given NumPy/CPU inputs, device tensors for device inputs.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import check, load, ptr, stream_ptr

XMIN, XMAX, YMIN, YMAX, LENGTH, HEIGHT = -2.5, 2.5, -1.7, 1.5, 256, 128   # IMG_Eagle.py:95-99


def _resample(grid, pos, step_x, step_y, x_min, y_min):
    T, H, W, C = grid.shape
    out = torch.empty((T, pos.shape[1], C), dtype=torch.float32, device=grid.device)
    with torch.cuda.device(grid.device):
        check(load().fl_grid2mesh(ptr(grid), ptr(pos), ptr(out), T, pos.shape[1], H, W, C, float(x_min), float(y_min),
                                  float(step_x), float(step_y), stream_ptr()), "fl_grid2mesh")
    return out


def grid2mesh(velocity_grid, pressure_grid, mesh_pos, device=None):
    """Project the grid back onto the mesh nodes (IMG_Eagle.py:93-123).

    velocity_grid (T, H, W, Cv), pressure_grid (T, H, W, Cp), mesh_pos (T, N, 2)
    -> velocity_mesh (T, N, Cv), pressure_mesh (T, N, Cp)."""
    _lib.require_cuda()
    on_device = torch.is_tensor(velocity_grid) and velocity_grid.is_cuda
    dev = velocity_grid.device if on_device else torch.device(device or "cuda")

    def dev_f32(a):
        t = a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a))
        return t.to(device=dev, dtype=torch.float32).contiguous()

    vg, pg, mp = dev_f32(velocity_grid), dev_f32(pressure_grid), dev_f32(mesh_pos)
    if vg.dim() != 4 or pg.dim() != 4 or mp.dim() != 3 or mp.shape[2] != 2:
        raise ValueError("grid2mesh expects (T,H,W,C) grids and (T,N,2) mesh positions")
    if vg.shape[:3] != pg.shape[:3] or vg.shape[0] != mp.shape[0]:
        raise ValueError("grid2mesh: velocity, pressure and mesh_pos disagree on T/H/W")
    x = np.linspace(XMIN, XMAX, LENGTH)              # IMG_Eagle.py:101-102
    y = np.linspace(YMAX, YMIN, HEIGHT)
    step_x, step_y = x[1] - x[0], y[1] - y[0]
    vm = _resample(vg, mp, step_x, step_y, XMIN, YMIN)
    pm = _resample(pg, mp, step_x, step_y, XMIN, YMIN)
    if on_device:
        return vm, pm
    return vm.cpu(), pm.cpu()
