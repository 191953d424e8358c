"""The pre-gridded EAGLE consumer on the GPU: `EagleDataset` over `states.npy` and `grid2mesh`.

Mirror of `/root/reference/eagle/Dataloader/IMG_Eagle.py`: the data set class (`:8-90`: window of the pre-gridded 4-channel
states, normalised with the fixed mean / std of `:76-77`, pixel-type mask, optional irregular mesh) and `grid2mesh` (`:93-123`:
same constants, same float32 index arithmetic as NumPy 1.26 evaluates it, same row flip and negative-index wrap).  `grid2mesh`
returns CPU tensors like the reference when given NumPy/CPU inputs, device tensors for device inputs.
"""
from __future__ import annotations

import ctypes
import os
import random

import numpy as np
import torch
from torch.utils.data import Dataset

from . import _lib
from ._lib import check, load, ptr, stream_ptr

XMIN, XMAX, YMIN, YMAX, LENGTH, HEIGHT = -2.5, 2.5, -1.7, 1.5, 256, 128   # IMG_Eagle.py:95-99
STATE_MEAN = (-0.0147, 0.2125, -0.5327, 3.7694)                          # IMG_Eagle.py:76,86
STATE_STD = (1.5943, 1.8824, 6.3553, 9.0565)                             # IMG_Eagle.py:77,87


def _affine(state, denormalize):
    """(state - mean) / std or state * std + mean over the last (4-channel) axis; CUDA tensors go through
    fl_affine_channels, CPU tensors through the same two fp32 torch operations the reference uses."""
    if state.shape[-1] != 4:
        raise RuntimeError(f"shape '[-1, 4]' is invalid for input of size {state.numel()}")      # the reference's reshape(-1, 4)
    if not state.is_cuda:
        mean, std = torch.tensor(STATE_MEAN).to(state.device), torch.tensor(STATE_STD).to(state.device)
        flat = state.reshape(-1, 4)
        return ((flat * std + mean) if denormalize else ((flat - mean) / std)).reshape(state.shape)
    x = state.float().contiguous()
    out = torch.empty_like(x)
    m, s = (ctypes.c_float * 4)(*STATE_MEAN), (ctypes.c_float * 4)(*STATE_STD)
    with torch.cuda.device(x.device):
        check(load().fl_affine_channels(ptr(x), ptr(out), x.numel(), 4, m, s, 1 if denormalize else 0, stream_ptr()),
              "fl_affine_channels")
    return out


class EagleDataset(Dataset):
    """IMG_Eagle.py:8-90 with the window's float conversion and normalisation on the GPU: the window of `states.npy` goes from
    the memory map through pinned memory to the device, one kernel normalises it.  `output_device=None` (default) leaves
    `'states'` on the GPU for a GPU consumer; `"cpu"` returns the reference's host tensor.  `splits_dir` is where
    `{mode}.txt` lives (the reference reads `Splits/{mode}.txt` relative to the working directory)."""

    def __init__(self, data_path, mode="test", window_length=990, with_mesh=False, device=None, output_device=None,
                 splits_dir="Splits"):
        super().__init__()
        if mode not in ("train", "test", "valid"):
            raise AssertionError(f"unknown mode {mode!r}")
        if window_length > 990:
            raise AssertionError("window length must be smaller than 990")
        if not os.path.exists(data_path):
            raise AssertionError(f"Path {data_path} does not exist")
        _lib.require_cuda()
        self.fn, self.mode, self.window_length, self.with_mesh = data_path, mode, window_length, with_mesh
        self.length = 990
        with open(os.path.join(splits_dir, f"{mode}.txt")) as f:
            self.dataloc = [os.path.join(data_path, ln.strip()) for ln in f]
        self.device = torch.device(device or "cuda")
        self.output_device = output_device
        self._pin = None

    def __len__(self):
        return len(self.dataloc)

    def _window_start(self):
        """IMG_Eagle.py:39-41: the full trajectory starts at 1; shorter windows start at 550 in test / valid, at random in training."""
        if self.window_length == 990:
            return 1
        if self.mode != "train":
            return 550
        return random.randint(1, 990 - self.window_length)

    def _upload(self, win):
        """A window of the memory-mapped states -> device float32: through a reused pinned buffer when it is float32 already."""
        if win.dtype != np.float32:
            return torch.from_numpy(win.copy()).to(self.device).float()           # the reference's `.float()`, on the device
        if self._pin is not None and self._pin.shape == win.shape:
            torch.cuda.current_stream(self.device).synchronize()                  # the previous window's upload has left the buffer
        else:
            self._pin = torch.empty(win.shape, dtype=torch.float32, pin_memory=True)
        self._pin.numpy()[...] = win                                              # page cache -> pinned memory
        return self._pin.to(self.device, non_blocking=True)

    def __getitem__(self, item):
        t = self._window_start()
        sim_dir = self.dataloc[item]
        frames = slice(t, t + self.window_length)
        states = np.load(os.path.join(sim_dir, "states.npy"), mmap_mode='r')
        pixel_type = np.load(os.path.join(sim_dir, "pixel_type.npy"), mmap_mode='r')
        normalised = self.normalize(self._upload(states[frames]))
        if self.output_device is not None:
            normalised = normalised.to(self.output_device)
        output = {'states': normalised, 'mask': pixel_type.copy(),
                  'example': torch.tensor((int(sim_dir.split("/")[-2]),))}
        if self.with_mesh:                                                        # IMG_Eagle.py:51-68: the irregular mesh of the same window
            mesh_dir = sim_dir.replace("_img", "")
            if not os.path.exists(mesh_dir):
                raise AssertionError(f"Can not find mesh files in {mesh_dir}, please check the path in the dataloader")
            sim = np.load(os.path.join(mesh_dir, 'sim.npz'), mmap_mode='r')
            take = lambda key: sim[key][frames].copy()
            output.update(mesh_pos=take("pointcloud"), mesh_velocity=np.stack([take("VX"), take("VY")], axis=-1),
                          mesh_pressure=np.stack([take("PS"), take("PG")], axis=-1), mesh_node_type=take("mask"))
        return output

    def normalize(self, state):
        return _affine(state, False)

    def denormalize(self, state):
        return _affine(state, True)


def _resample(grid, pos, step_x, step_y, x_min, y_min, oob=None):
    T, H, W, C = grid.shape
    out = torch.empty((T, pos.shape[1], C), dtype=torch.float32, device=grid.device)
    with torch.cuda.device(grid.device):
        check(load().fl_grid2mesh(ptr(grid), ptr(pos), ptr(out), T, pos.shape[1], H, W, C, float(x_min), float(y_min),
                                  float(step_x), float(step_y), ptr(oob), stream_ptr()), "fl_grid2mesh")
    return out


def grid2mesh(velocity_grid, pressure_grid, mesh_pos, device=None, check_bounds=True):
    """Project the grid back onto the mesh nodes (IMG_Eagle.py:93-123).

    velocity_grid (T, H, W, Cv), pressure_grid (T, H, W, Cp), mesh_pos (T, N, 2)
    -> velocity_mesh (T, N, Cv), pressure_mesh (T, N, Cp).  A node outside the grid raises IndexError as the reference's
    fancy indexing does (`check_bounds=False` skips the check and its device -> host read: such nodes are then clamped)."""
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
    oob = torch.zeros(1, dtype=torch.int32, device=dev) if check_bounds else None
    vm = _resample(vg, mp, step_x, step_y, XMIN, YMIN, oob)
    pm = _resample(pg, mp, step_x, step_y, XMIN, YMIN, None)
    if check_bounds and int(oob.item()):
        raise IndexError(f"grid2mesh: {int(oob.item())} node position(s) index outside the {vg.shape[1]} x {vg.shape[2]} grid")
    if on_device:
        return vm, pm
    return vm.cpu(), pm.cpu()
