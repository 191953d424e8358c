"""DilResNet image dataset on the GPU data path (a second consumer of the same kernels).

Mirror of `/root/reference/eagle/Dataloader/IMG_MGN.py:12-174` (`EagleDataset`): walks `data_path/mode` for
`.pkl` trajectories, interpolates `window_length` frames to the 238-point grid, crops 16 pixels per side for the
airfoil set (:91-95), normalises EVERY pixel with the per-dataset constants (:141-157) and returns channel-last
frames `{'states': (T, H, W, 3) float32, 'mask': (T, H, W) bool}`.  No padding, no patchify.
"""
from __future__ import annotations

import ctypes
import os
from collections import OrderedDict
import pickle
import random
import re

import numpy as np
import torch
from torch.utils.data import Dataset

from ._ingest_worker import unpickle_lazy
from ._lib import FL_NO_NORM, check, load, ptr, stream_ptr
from .airfoil_ds import crop_airfoil_mesh
from .field_path import DeviceTrajectory
from .mesh_utils import MeshPlan

AIRFOIL_STATS = ((170.1, -1.183, 9.935e+04), (71.06, 46.73, 8964.0))          # IMG_MGN.py:143-145
CYLINDER_STATS = ((0.823, 0.0005865, 0.04763), (0.275, 0.275, 0.275))          # IMG_MGN.py:147-149


def _natsorted(seq):
    return sorted(seq, key=lambda s: [int(t) if t.isdigit() else t for t in re.split(r"(\d+)", s)])


def interp_frames(traj: DeviceTrajectory, t0, n_frames, interval=1, crop=0, means=None, stds=None):
    """-> states (T, H, W, 3) f32 and mask (T, H, W) u8 on the device (one launch of fl_interp_frames)."""
    plan = traj.plan
    H, W = plan.nx - 2 * crop, plan.ny - 2 * crop
    if t0 < 0 or t0 + (n_frames - 1) * interval >= traj.n_steps:
        raise ValueError(f"frames {t0}..{t0 + (n_frames - 1) * interval} outside trajectory of {traj.n_steps} steps")
    with torch.cuda.device(plan.device):
        states = torch.empty((n_frames, H, W, 3), dtype=torch.float32, device=plan.device)
        mask = torch.empty((n_frames, H, W), dtype=torch.uint8, device=plan.device)
        m = (ctypes.c_float * 3)(*(means or (0, 0, 0)))
        s = (ctypes.c_float * 3)(*(stds or (1, 1, 1)))
        check(load().fl_interp_frames(ptr(plan.cell_idx_d), ptr(plan.cell_w_d), plan.nx, plan.ny, crop, ptr(traj.vel_buf),
                                      ptr(traj.prs_buf), plan.n_nodes, traj.vel_stride, traj.prs_stride, t0, interval, n_frames,
                                      m, s, 0 if means is not None else FL_NO_NORM, ptr(states), ptr(mask), stream_ptr()),
              "fl_interp_frames")
    return states, mask


class EagleDataset(Dataset):
    """IMG_MGN.py:12 -- the name is the reference's (it loads MGN pickles, not EAGLE).  Trajectories (node fields + mesh plan) stay
    resident on the device, the `cache_size` most recently used ones (a 600-step trajectory is ~15 MB of HBM)."""
    cache_size = 256

    def __init__(self, data_path, mode="test", window_length=990, with_mesh=False, device=None, output_device="cpu"):
        super().__init__()
        assert mode in ["train", "test", "valid"]
        self.window_length = window_length
        assert window_length <= 990, "window length must be smaller than 990"
        self.fn = os.path.join(data_path, mode)
        assert os.path.exists(self.fn), f"Path {self.fn} does not exist"
        self.dataloc = []
        for root, _, files in os.walk(self.fn):
            for filename in files:
                if filename.endswith(".pkl"):
                    self.dataloc.append(os.path.join(root, filename))
        self.dataloc = _natsorted(self.dataloc)
        self.mode, self.length, self.with_mesh = mode, 990, with_mesh
        self.device, self.output_device = device, output_device
        self._cache = OrderedDict()

    def __len__(self):
        return len(self.dataloc)

    def _stats(self):
        if "airfoil" in self.fn:
            return AIRFOIL_STATS
        if "cylinder" in self.fn:
            return CYLINDER_STATS
        raise ValueError(f"Unknown dataset {self.fn}")

    def _load_step(self, save_file) -> DeviceTrajectory:
        """IMG_MGN.py:46-76 (cached per file)."""
        key = (save_file, os.path.getmtime(save_file))
        if key in self._cache:
            self._cache.move_to_end(key)
        else:
            save_data = unpickle_lazy(save_file, ("mesh_pos", "cells", "velocity", "pressure"))   # field arrays as views of the file
            if save_data is None:
                with open(save_file, "rb") as f:
                    save_data = pickle.load(f)
            pos, faces, vel, prs = save_data["mesh_pos"], save_data["cells"], save_data["velocity"], save_data["pressure"]
            if "airfoil" in self.fn:
                m, pos, faces = crop_airfoil_mesh(pos, faces)
                vel, prs = np.ascontiguousarray(vel[:, m]), np.ascontiguousarray(prs[:, m])
            self._cache[key] = DeviceTrajectory(vel, prs, MeshPlan(pos, faces, 238, None, self.device))
            while len(self._cache) > max(1, self.cache_size):
                self._cache.popitem(last=False)
        return self._cache[key]

    def __getitem__(self, item):
        # IMG_MGN.py:99-101
        t = 0 if self.window_length == 600 else random.randint(0, 600 - self.window_length)
        t = 100 if self.mode != "train" and self.window_length != 600 else t
        traj = self._load_step(self.dataloc[item])
        means, stds = self._stats()
        states, mask = interp_frames(traj, t, self.window_length, 1, 16 if "airfoil" in self.fn else 0, means, stds)
        mask = mask.bool()
        if self.output_device == "cpu":
            return {"states": states.cpu(), "mask": mask.cpu().numpy()}
        return {"states": states, "mask": mask}

    def normalize(self, state):
        """IMG_MGN.py:141-157: channel-first (.., 3, H, W) states."""
        means, stds = self._stats()
        means = torch.tensor(means, device=state.device).reshape(1, 3, 1, 1)
        stds = torch.tensor(stds, device=state.device).reshape(1, 3, 1, 1)
        return (state - means) / stds

    def denormalize(self, state):
        """IMG_MGN.py:159-174: channel-last states."""
        means, stds = self._stats()
        return state * torch.tensor(stds, device=state.device) + torch.tensor(means, device=state.device)
