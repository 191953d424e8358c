"""`MGNDataset` (Cylinder flow) on the GPU data path.

Host-side mirror of `/root/reference/src/dataloader/simple_dataloader.py:23-229`: same constructor,
same attributes (`N_x_patch`, `N_y_patch`, `N_patch`, `patch_size`, `seq_len`, `ds_min_max`,
`save_files`), same 5-tuple from `__getitem__` / `ds_get`.  The per-frame work (3x to_grid, pad,
mask concat, unfold, permute, normalise) is one launch of csrc/fl_interp.cu; the triangle finder
is built once per trajectory file and cached (the reference rebuilds it on every `__getitem__`,
simple_dataloader.py:181), together with the node fields resident in HBM.
"""
from __future__ import annotations

import os
import pickle
import random
from collections import OrderedDict

import numpy as np
import torch
from torch.utils.data import Dataset

from ._ingest_worker import unpickle_lazy
from ._lib import check, load, ptr, stream_ptr
from .field_path import TrajBatch, CYLINDER, DeviceTrajectory, Personality, interp_patchify
from .mesh_utils import MeshPlan, to_grid
from .traj_store import PinnedStage, TrajectoryFile


def num_patches(dim_size, kern_size, stride, padding=0):
    """simple_dataloader.py:16-20."""
    return (dim_size + 2 * padding - kern_size) // stride + 1


def position_ids(seq_len, n_x_patch, n_y_patch):
    """simple_dataloader.py:218-226 -> int64 (seq_len-1, N_patch, 3).  Labels patch l as (l % N_x, l // N_x)
    although F.unfold orders patches l = bx * N_y + by; reproduced as is (model weights depend on it)."""
    n_patch = n_x_patch * n_y_patch
    arange = np.arange((seq_len - 1) * n_patch)
    x_idx = arange % n_x_patch
    y_idx = (arange // n_x_patch) % n_y_patch
    t_idx = arange // n_patch
    ids = np.stack([x_idx, y_idx, t_idx], axis=1).reshape(seq_len - 1, n_patch, 3)
    return torch.from_numpy(ids.astype(np.int64))


def sample_assemble(states, mask):
    """states (B, T, L, 3, px, py) f32, mask (B, T, L, px, py) u8 -> diffs = states[:, 1:] - states[:, :-1] and
    masks = mask[:, 1:] repeated over the channels as torch.bool (simple_dataloader.py:93,100), one kernel launch."""
    B, T, L, C, px, py = states.shape
    diffs = torch.empty((B, T - 1, L, C, px, py), dtype=torch.float32, device=states.device)
    mask3 = torch.empty((B, T - 1, L, C, px, py), dtype=torch.bool, device=states.device)
    if T > 1:
        with torch.cuda.device(states.device):
            check(load().fl_sample_assemble(ptr(states), ptr(mask), B, T, L, px, py, ptr(diffs), ptr(mask3), stream_ptr()),
                  "fl_sample_assemble")
    return diffs, mask3


class _GpuFieldDataset(Dataset):
    """Shared machinery of MGNDataset / AirfoilDataset."""

    personality: Personality = CYLINDER
    cache_size = 4096          # trajectories kept resident on the device (plan + node fields), at most ...
    cache_fraction = 0.25      # ... this fraction of the device's memory (a 1000-trajectory MGN data set is ~18 GB of 180)
    ingest_workers = None      # processes that unpickle ahead of the GPU (None: min(12, cores - 2); 0: unpickle in-process)
    ingest_slot_bytes = 24 << 20   # page-locked bytes per ingest slot (two per worker); grows when a trajectory does not fit
    ingest_airfoil_crop = False
    plan_cache_size = 256   # mesh plans kept resident (static tables, ~1 MB each) for window loads from .fgt files
    window_loads = True     # .fgt file not resident: read and upload only the time steps the sample needs

    def __init__(self, load_dir, resolution: int, patch_size: tuple, stride: tuple, seq_len: int, seq_interval=1,
                 pad=True, mode="train", normalize=True, noise=None, device=None, output_device=None,
                 numpy_semantics=None):
        super().__init__()
        assert mode in ["train", "valid", "test"]
        if len(patch_size) != 2 or len(stride) != 2 or min(*patch_size, *stride) < 1:
            raise ValueError(f"patch_size and stride must be pairs of positive ints, got {patch_size}, {stride}")
        self.mode = mode
        self.load_dir = load_dir
        self.resolution = resolution
        self.patch_size = patch_size
        self.stride = stride
        self.pad = pad
        self.seq_len = seq_len
        self.seq_interval = seq_interval
        self.max_step_num = 600 - self.seq_len * self.seq_interval
        self.normalize = normalize
        self.noise = noise
        self.device = device
        self.output_device = output_device     # None: tensors stay on the GPU; "cpu": reference-style host tensors
        self.numpy_semantics = numpy_semantics
        self._cache = OrderedDict()
        self._cache_bytes = 0
        self._unverified = []
        self.ingest_times = {"wait": 0.0, "plan": 0.0, "upload": 0.0, "n": 0}
        self._ingest = None
        self._uploads = []         # (event, release) of ingest slots whose host -> device copies are still in flight
        self._plans = OrderedDict()
        self._pos_ids = None
        self._pinned = PinnedStage()

        self.save_files = self._list_files()
        # simple_dataloader.py:45-56: probe file [1], step 20, for value ranges and the patch grid
        traj = self._load_step(self.save_files[1])
        # ... min / max over the whole padded frame, BEFORE the airfoil's ring crop (airfoil_ds.py:46-50): no crop here
        full = Personality(self.personality.name, self.personality.flip_y, 0, self.personality.mask_aware_norm,
                           self.personality.means, self.personality.stds)
        # (the frame tiled without overlap or gaps -- stride = patch -- so that every pixel of it is seen exactly once)
        states, _, _ = interp_patchify(traj, 20, 1, 1, self.patch_size, full, normalize=False, pad=self.pad)
        lo, hi = states[0].amin(dim=(0, 2, 3)).cpu().numpy(), states[0].amax(dim=(0, 2, 3)).cpu().numpy()
        if not self.pad:                        # unfold dropped the remainder columns / rows: the plain gridded frame instead
            plan = traj.plan
            fields = torch.stack([traj.velocity[20, :, 0], traj.velocity[20, :, 1], traj.pressure[20, :, 0]])
            grid, _ = to_grid(fields, plan.grid_x, plan.grid_y, plan, plan.tri_index)
            lo, hi = grid.amin(dim=(1, 2)).cpu().numpy(), grid.amax(dim=(1, 2)).cpu().numpy()
        self.ds_min_max = [(lo[0], hi[0]), (lo[1], hi[1]), (lo[2], hi[2])]
        self._table_of(traj.plan)               # refuses a patch / stride / pad combination that leaves no patches
        # simple_dataloader.py:52-56 / airfoil_ds.py:52-55: patches per axis of the frame BEFORE the airfoil's ring crop, minus
        # two for the airfoil -- which is the number of patches unfold produces only when stride == patch_size; the reference's
        # attributes (and the position ids made from them) are reproduced as they are
        nx_full = traj.plan.nx + ((-traj.plan.nx) % self.patch_size[0] if self.pad else 0)
        ny_full = traj.plan.ny + ((-traj.plan.ny) % self.patch_size[1] if self.pad else 0)
        ring = 2 * self.personality.crop_patches
        self.N_x_patch = num_patches(nx_full, self.patch_size[0], self.stride[0]) - ring
        self.N_y_patch = num_patches(ny_full, self.patch_size[1], self.stride[1]) - ring
        self.N_patch = self.N_x_patch * self.N_y_patch

    def _table_of(self, plan):
        return plan.patch_table(self.patch_size, self.personality.crop_patches, self.personality.flip_y, self.stride, self.pad)

    # -- file handling ----------------------------------------------------------------------
    def _list_files(self):
        # the reference's pickles, or the flat .fgt files of traj_store.py (same stems, converted once)
        return sorted(self._one_per_stem())

    def _one_per_stem(self):
        """A directory converted in place holds x.pkl AND x.fgt: one entry per trajectory (the .fgt), so __len__ and the
        index -> trajectory map stay those of the reference's listing of the pickles."""
        names = [f for f in os.listdir(f"{self.load_dir}/") if f.endswith(('.pkl', '.fgt'))]
        fgt = {os.path.splitext(f)[0] for f in names if f.endswith('.fgt')}
        return [f for f in names if f.endswith('.fgt') or os.path.splitext(f)[0] not in fgt]

    def _prepare_mesh(self, save_data):
        """-> (pos, faces, velocity, pressure) as the locate step should see them."""
        return save_data['mesh_pos'], save_data['cells'], save_data['velocity'], save_data['pressure']

    def _load_step(self, save_file) -> DeviceTrajectory:
        """simple_dataloader.py:154-164: unpickle + mesh interpolation set-up (cached per file)."""
        path = f"{self.load_dir}/{save_file}"
        key = (path, os.path.getmtime(path))
        traj = self._cache.get(key)
        if traj is not None:
            self._cache.move_to_end(key)
            return traj
        if path.endswith('.fgt'):
            # flat format: crop (if any) was applied at conversion time, node fields are already in the device pitch
            tf = TrajectoryFile(path)
            if self.personality.crop_patches and not tf.header["meta"].get("airfoil_crop"):
                # converted without the airfoil node crop: apply it now, like the pickle route does
                n = tf.n_nodes
                raw = {"mesh_pos": tf.mesh_pos, "cells": tf.cells,
                       "velocity": np.asarray(tf.array("velocity"))[:, :2 * n].reshape(tf.n_steps, n, 2),
                       "pressure": np.asarray(tf.array("pressure"))[:, :n, None]}
                pos, faces, vel, prs = self._prepare_mesh(raw)
                plan = self._new_plan(pos, faces)
                traj = DeviceTrajectory(vel, prs, plan)
            else:
                plan = self._new_plan(tf.mesh_pos, tf.cells)
                traj = tf.to_device(plan, pinned=self._pinned)
        elif self._ingest_pool() is not None:
            # a worker process unpickled (and cropped) it into a page-locked slot: only the copies are issued here
            import time as _time
            t0 = _time.perf_counter()
            pos, faces, vel, prs, release, prepared = self._ingest.take(path)
            t1 = _time.perf_counter()
            plan = self._new_plan(pos, faces, prepared)
            t2 = _time.perf_counter()
            traj = DeviceTrajectory.from_padded(vel, prs, plan)
            t3 = _time.perf_counter()
            tt = self.ingest_times                 # seconds spent waiting for the pool / building the mesh plan / issuing the upload
            tt["wait"] += t1 - t0; tt["plan"] += t2 - t1; tt["upload"] += t3 - t2; tt["n"] += 1
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(plan.device))
            self._uploads.append((ev, release))
            self._reap_uploads()
        else:
            save_data = unpickle_lazy(path, ("mesh_pos", "cells", "velocity", "pressure"))      # field arrays as views of the file
            if save_data is None:
                with open(path, 'rb') as f:
                    save_data = pickle.load(f)
            pos, faces, vel, prs = self._prepare_mesh(save_data)
            plan = self._new_plan(pos, faces)
            # the only host copy: file mapping (or unpickled arrays) -> reused pinned buffers in the device pitch -> one upload
            T, N = int(np.shape(vel)[0]), plan.n_nodes
            hv, hp = self._pinned.get("vel", (T, 2 * plan.n_padded)), self._pinned.get("prs", (T, plan.n_padded))
            hv_np, hp_np = hv.numpy(), hp.numpy()
            hv_np[:, :2 * N] = np.asarray(vel).reshape(T, 2 * N)
            hv_np[:, 2 * N:] = 0
            hp_np[:, :N] = np.asarray(prs).reshape(T, N)
            hp_np[:, N:] = 0
            traj = DeviceTrajectory.from_padded(hv, hp, plan)
            torch.cuda.current_stream(plan.device).synchronize()       # the pinned buffers are reused by the next load
        self._cache[key] = traj
        self._cache_bytes += self._traj_bytes(traj)
        budget = self._cache_budget(traj.plan.device)
        while self._cache and (len(self._cache) > self.cache_size or self._cache_bytes > budget):
            if len(self._cache) == 1 and self.cache_size >= 1:
                break                                          # never evict the trajectory just loaded
            _, old = self._cache.popitem(last=False)
            self._cache_bytes -= self._traj_bytes(old)
        return traj

    def _new_plan(self, pos, faces, prepared=None):
        """A mesh plan built without waiting for the GPU (MeshPlan(sync=False)); `_verify_plans` checks it before the sample
        that used it is handed out.  `prepared`: its host side, when an ingest worker has done it already."""
        plan = MeshPlan(pos, faces, self.resolution, self.numpy_semantics, self.device, sync=False, prepared=prepared)
        self._unverified.append(plan)
        return plan

    def _verify_plans(self):
        """-> True if every plan built since the last call was fine; False if one had to be rebuilt (bin store overflow on a
        very uneven mesh): the caller recomputes its sample from the rebuilt tables."""
        ok = True
        for plan in self._unverified:
            ok = plan.ready() and ok
        self._unverified = []
        return ok

    @staticmethod
    def _traj_bytes(traj):
        return traj.vel_buf.numel() * 4 + traj.prs_buf.numel() * 4

    def _cache_budget(self, device):
        if not hasattr(self, "_budget"):
            self._budget = int(self.cache_fraction * torch.cuda.get_device_properties(device).total_memory)
        return self._budget

    def _ingest_pool(self):
        """The unpickling pool (created on first use), or None when disabled."""
        if self.ingest_workers == 0:
            return None
        if self._ingest is not None and self._ingest.grow_to > self.ingest_slot_bytes:
            # a trajectory did not fit the page-locked slots (it was unpickled in this process instead): larger slots from here on
            self._reap_uploads(wait=True)                  # no copy out of the old slots is still in flight
            self.ingest_slot_bytes = self._ingest.grow_to
            self._ingest.close()
            self._ingest = None
        if self._ingest is None:
            from .ingest import PickleIngest
            from .mesh_utils import default_numpy_semantics
            self._ingest = PickleIngest(workers=self.ingest_workers, slot_bytes=self.ingest_slot_bytes,
                                        airfoil_crop=self.ingest_airfoil_crop,
                                        plan_resolution=self.resolution,
                                        numpy_semantics=self.numpy_semantics or default_numpy_semantics())
        return self._ingest

    def _reap_uploads(self, wait=False):
        keep = []
        for ev, release in self._uploads:
            if wait:
                ev.synchronize()
            if ev.query():
                release()
            else:
                keep.append((ev, release))
        self._uploads = keep

    def prefetch(self, requests):
        """Start loading the pickles of upcoming samples -- [(save_file or index, step_num), ...] or file names / indices --
        in the ingest pool; resident trajectories and .fgt files need nothing.  Returns at once."""
        if self.ingest_workers == 0:
            return
        paths = []
        for r in requests:
            f = r[0] if isinstance(r, (tuple, list)) else r
            if isinstance(f, (int, np.integer)):
                f = self.save_files[int(f)]
            path = f"{self.load_dir}/{f}"
            if path.endswith('.pkl') and (path, os.path.getmtime(path)) not in self._cache and path not in paths:
                paths.append(path)
        if paths and self._ingest_pool() is not None:          # (the pool is only started once a pickle has to be read)
            self._ingest.submit(paths)

    def _load_window(self, save_file, step_num):
        """-> (trajectory, step inside it).  A resident trajectory is used as is; a `.fgt` file that is not resident is read
        through its memory map for the sample's time steps only (the mesh plan is cached per file), instead of the
        reference's whole-trajectory unpickle per sample (simple_dataloader.py:154-164)."""
        path = f"{self.load_dir}/{save_file}"
        key = (path, os.path.getmtime(path))
        # a data set that fits the resident cache is simply kept on the device (3-4x faster per sample than a window load)
        if key in self._cache or len(self.save_files) <= self.cache_size or not (self.window_loads and path.endswith('.fgt')):
            return self._load_step(save_file), step_num
        tf = TrajectoryFile(path)
        if self.personality.crop_patches and not tf.header["meta"].get("airfoil_crop"):
            return self._load_step(save_file), step_num            # needs the node crop: whole-trajectory route
        plan = self._plans.get(key)
        if plan is None:
            plan = self._new_plan(tf.mesh_pos, tf.cells)
            self._plans[key] = plan
            while len(self._plans) > self.plan_cache_size:
                self._plans.popitem(last=False)
        else:
            self._plans.move_to_end(key)
        last = step_num + (self.seq_len - 1) * self.seq_interval + 1
        return tf.to_device(plan, pinned=self._pinned, first=step_num, last=last), 0

    # -- reference API ------------------------------------------------------------------------
    def __getitem__(self, idx):
        # simple_dataloader.py:67-70: random start in train, fixed 100 otherwise
        step_num = random.randint(0, self.max_step_num)
        step_num = 100 if self.mode in ["test", "valid"] else step_num
        return self.ds_get(save_file=self.save_files[idx], step_num=step_num)

    def ds_get(self, save_file=None, step_num=None):
        """simple_dataloader.py:72-102 -> (input_states, next_state, diffs, masks, position_ids)."""
        if save_file is None:
            save_file = random.choice(self.save_files)
        elif isinstance(save_file, int):
            save_file = self.save_files[save_file]
        if step_num is None:
            step_num = np.random.randint(0, self.max_step_num)
        if step_num > self.max_step_num:          # simple_dataloader.py:177-179
            step_num = self.max_step_num
        traj, local_step = self._load_window(save_file, step_num)
        states, mask, _ = interp_patchify(traj, local_step, self.seq_len, self.seq_interval, self.patch_size,
                                          self.personality, normalize=self.normalize, stride=self.stride, pad=self.pad)
        diffs, masks = sample_assemble(states.unsqueeze(0), mask.unsqueeze(0))       # :93, :100 in one launch
        if not self._verify_plans():            # (everything above is enqueued by now: this wait overlaps the GPU work)
            return self.ds_get(save_file, step_num)
        out = (states[:-1], states[1:], diffs[0], masks[0], self._get_pos_id().to(states.device))
        if self.output_device is not None:
            out = tuple(t.to(self.output_device) for t in out)
        return out

    def ds_get_many(self, requests):
        """[(save_file or index, step_num), ...] -> list of 5-tuples, all samples in ONE kernel launch.

        The per-sample call is bound by host work (descriptor set-up, one launch, five small tensor ops); a DataLoader batch
        of B samples costs little more than one.  Differences and mask expansion are done once on the stacked tensors."""
        self.prefetch(requests)             # every pickle of the batch is unpickled at the same time, in the pool
        trajs, steps = [], []
        for save_file, step_num in requests:
            if isinstance(save_file, int):
                save_file = self.save_files[save_file]
            step_num = min(int(step_num), self.max_step_num)          # simple_dataloader.py:177-179
            traj, local_step = self._load_window(save_file, step_num)
            trajs.append(traj)
            steps.append(local_step)
        tabs = [self._table_of(t.plan) for t in trajs]
        batch = TrajBatch(trajs, tabs, steps, self.seq_interval, self.seq_len)
        states, mask = batch.run(self.personality, self.normalize)     # (B, T, L, 3, px, py), (B, T, L, px, py)
        diffs, masks = sample_assemble(states, mask)
        if not self._verify_plans():
            return self.ds_get_many(requests)
        pos = self._get_pos_id().to(states.device)
        out = [(states[b, :-1], states[b, 1:], diffs[b], masks[b], pos) for b in range(len(trajs))]
        if self.output_device is not None:
            out = [tuple(t.to(self.output_device) for t in o) for o in out]
        return out

    def __getitems__(self, indices):
        """torch's DataLoader fetches a whole batch through this when it exists: one launch per batch."""
        reqs = []
        for idx in indices:
            step_num = random.randint(0, self.max_step_num)
            reqs.append((self.save_files[idx], 100 if self.mode in ["test", "valid"] else step_num))
        return self.ds_get_many(reqs)

    def _get_pos_id(self):
        """simple_dataloader.py:218-226 (the labelling quirk is reproduced as is)."""
        if self._pos_ids is None:
            self._pos_ids = position_ids(self.seq_len, self.N_x_patch, self.N_y_patch)
        return self._pos_ids

    def __len__(self):
        return len(self.save_files)


class MGNDataset(_GpuFieldDataset):
    """Load a sequence of timesteps of one Cylinder-flow trajectory (simple_dataloader.py:23)."""
    personality = CYLINDER
