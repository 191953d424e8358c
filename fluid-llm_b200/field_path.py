"""The per-step field data path as one batched call: node fields of many trajectories ->
normalised, patchified states (+ mask) on the device.

This is the engine behind `MGNDataset.ds_get` / `AirfoilDataset.ds_get` (the host-side mirrors of
`/root/reference/src/dataloader/simple_dataloader.py:72-102` and `airfoil_ds.py:71-103`) and what
bench.py times.  One launch of csrc/fl_interp.cu processes every requested frame of every
trajectory in the batch; trajectories may have different meshes as long as they share the patch
grid (which the reference's DataLoader collation requires anyway).
"""
from __future__ import annotations

import ctypes
import os
import warnings
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from ._lib import FL_MASK_AWARE_NORM, FL_NO_NORM, FlTraj, check, load, stream_ptr
from .mesh_utils import MeshPlan


@dataclass(frozen=True)
class Personality:
    """Dataset-specific geometry/normalisation rules (SURVEY.md 'two dataset personalities')."""
    name: str
    flip_y: bool            # airfoil_ds.py:80
    crop_patches: int       # airfoil_ds.py:132-133 (outer ring of patches dropped)
    mask_aware_norm: bool   # airfoil_ds.py:236-242
    means: tuple
    stds: tuple


CYLINDER = Personality("cylinder", False, 0, False, (0.823, 0.0005865, 0.04763), (0.275, 0.275, 0.275))
AIRFOIL = Personality("airfoil", True, 1, True, (170.1, -1.183, 9.935e+04), (50.0, 50.0, 6197.0))


STAGED_MAX_NODES = 4800     # 3 frames x 16 bytes x nodes = the 227 KB of shared memory of one SM


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous block of trajectories owned by `rank` (sizes differ by at most one).  Trajectories are
    independent, so this is the whole multi-GPU story of the per-frame path: no data-path collective."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _as_tensor(a):
    """NumPy -> float32 CPU tensor without a copy where possible (read-only views of a mapped file included: only read here)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.flags.writeable:
        return torch.from_numpy(a)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", UserWarning)
        return torch.from_numpy(a)


class DeviceTrajectory:
    """Node fields of one trajectory resident in HBM, in the pickle's own layout
    (velocity f32[T,N,2], pressure f32[T,N,1]; max/ds_download/torch_MGN.py:68-95)."""

    def __init__(self, velocity, pressure, plan: MeshPlan):
        dev = plan.device
        v = velocity if torch.is_tensor(velocity) else _as_tensor(velocity)
        p = pressure if torch.is_tensor(pressure) else _as_tensor(pressure)
        if v.dim() != 3 or v.shape[2] != 2 or v.shape[1] != plan.n_nodes:
            raise ValueError(f"velocity must be (T, {plan.n_nodes}, 2), got {tuple(v.shape)}")
        if p.dim() == 2:
            p = p.unsqueeze(-1)
        if p.shape != (v.shape[0], plan.n_nodes, 1):
            raise ValueError(f"pressure must be ({v.shape[0]}, {plan.n_nodes}, 1), got {tuple(p.shape)}")
        # frame pitch padded to 16 bytes so whole frames can be bulk-copied (TMA) into shared memory
        T, N = int(v.shape[0]), plan.n_nodes
        self.prs_stride = (N + 3) // 4 * 4          # nodes padded to a multiple of 4 (pad values are zero)
        self.vel_stride = 2 * self.prs_stride
        self.vel_buf = torch.zeros((T, self.vel_stride), dtype=torch.float32, device=dev)
        self.prs_buf = torch.zeros((T, self.prs_stride), dtype=torch.float32, device=dev)
        self.velocity = self.vel_buf[:, :2 * N].view(T, N, 2)     # reference-shaped views of the padded buffers
        self.pressure = self.prs_buf[:, :N].view(T, N, 1)
        self.velocity.copy_(v.to(dtype=torch.float32), non_blocking=True)
        self.pressure.copy_(p.to(dtype=torch.float32), non_blocking=True)
        self.plan = plan
        self.n_steps = int(v.shape[0])

    @classmethod
    def from_padded(cls, vel_padded, prs_padded, plan: MeshPlan, stream=None, pinned=None):
        """Node fields already in the device pitch (traj_store .fgt files): host [T, vel_stride] / [T, prs_stride]
        -> pinned staging -> HBM, asynchronously on `stream` (default: the current stream)."""
        self = cls.__new__(cls)
        dev = plan.device
        T, N = int(vel_padded.shape[0]), plan.n_nodes
        self.prs_stride = (N + 3) // 4 * 4
        self.vel_stride = 2 * self.prs_stride
        if tuple(vel_padded.shape) != (T, self.vel_stride) or tuple(prs_padded.shape) != (T, self.prs_stride):
            raise ValueError(f"padded node fields must be ({T}, {self.vel_stride}) and ({T}, {self.prs_stride})")
        st = stream if stream is not None else torch.cuda.current_stream(dev)
        with torch.cuda.device(dev), torch.cuda.stream(st):
            self.vel_buf = torch.empty((T, self.vel_stride), dtype=torch.float32, device=dev)
            self.prs_buf = torch.empty((T, self.prs_stride), dtype=torch.float32, device=dev)
            for name, src, dst in (("vel", vel_padded, self.vel_buf), ("prs", prs_padded, self.prs_buf)):
                if torch.is_tensor(src):                          # already page-locked (ingest.PickleIngest slots)
                    dst.copy_(src, non_blocking=True)
                    continue
                if pinned is not None:
                    h = pinned.get(name, src.shape)
                    h.numpy()[...] = src                      # page cache / disk -> pinned memory
                else:
                    h = torch.from_numpy(np.ascontiguousarray(src)).pin_memory()
                dst.copy_(h, non_blocking=True)
            if pinned is not None:
                st.synchronize()                               # the staging buffers are reused by the next load
        self.velocity = self.vel_buf[:, :2 * N].view(T, N, 2)
        self.pressure = self.prs_buf[:, :N].view(T, N, 1)
        self.plan = plan
        self.n_steps = T
        return self


class TrajBatch:
    """Pre-built launch descriptors for a fixed batch (descriptor array already on the device), so
    the steady-state call is exactly one kernel launch and no host->device traffic."""

    def __init__(self, trajs, tables, t0s, interval, n_frames, want_mask=True, use_slots=None, out=None, tile_patches=None,
                 compact_idx=None):
        if use_slots is None:
            use_slots = os.environ.get("FLUIDGRID_SLOTS", "1") != "0"
        if compact_idx is None:        # 8-byte table records (FlTraj::idx_slot_format = 1) wherever the staged kernel may run
            compact_idx = os.environ.get("FLUIDGRID_IDX16", "1") != "0"
        if not (len(trajs) == len(tables) == len(t0s)) or not trajs:
            raise ValueError("trajs, tables and t0s must be non-empty and of equal length")
        tab0 = tables[0]
        for tab in tables:
            if (tab.n_bx, tab.n_by, tab.px, tab.py) != (tab0.n_bx, tab0.n_by, tab0.px, tab0.py):
                raise ValueError("all trajectories of a batch must share the patch grid")
        self.n_traj, self.n_frames, self.tab0 = len(trajs), int(n_frames), tab0
        dev = trajs[0].plan.device
        self.device = dev
        L, px, py = tab0.n_patches, tab0.px, tab0.py
        for tr, t0 in zip(trajs, t0s):
            last = int(t0) + (self.n_frames - 1) * int(interval)
            if t0 < 0 or last >= tr.n_steps or interval < 1 or n_frames < 1:
                raise ValueError(f"frames {t0}..{last} step {interval} outside trajectory of {tr.n_steps} steps")
        if out is not None:           # caller-owned output buffers (HostPipeline's device slots)
            self.states, self.mask = out
            if tuple(self.states.shape) != (self.n_traj, self.n_frames, L, 3, px, py) or self.states.dtype != torch.float32 \
                    or not self.states.is_contiguous() or self.states.device != dev:
                raise ValueError("out[0] must be a contiguous float32 (n_traj, n_frames, L, 3, px, py) tensor on the plan's device")
            if want_mask and (self.mask is None or tuple(self.mask.shape) != (self.n_traj, self.n_frames, L, px, py)
                              or self.mask.dtype != torch.uint8 or not self.mask.is_contiguous()):
                raise ValueError("out[1] must be a contiguous uint8 (n_traj, n_frames, L, px, py) tensor")
        else:
            self.states = torch.empty((self.n_traj, self.n_frames, L, 3, px, py), dtype=torch.float32, device=dev)
            self.mask = torch.empty((self.n_traj, self.n_frames, L, px, py), dtype=torch.uint8, device=dev) if want_mask else None
        # tile plans (csrc/fl_tiled.cu): one split of the patch grid for the whole batch, sized for its largest mesh
        ppx = px * py
        self.tile_plans = None
        # (tile_patches: None = tiles only where whole frames of the mesh do not fit shared memory at >= 3 frames per work
        # item -- measured on B200 the whole-mesh staged kernel is the faster one below that, profiles/README.md; 0 = never;
        # n > 0 = tiles of n patches)
        biggest = max(tables, key=lambda t: t.n_nodes)
        if tile_patches is None and os.environ.get("FLUIDGRID_TILE_PATCHES"):
            tile_patches = int(os.environ["FLUIDGRID_TILE_PATCHES"])
        want_tiles = tile_patches is not None and tile_patches > 0 or (tile_patches is None and biggest.n_nodes > STAGED_MAX_NODES)
        if ppx in (128, 256) and want_tiles:
            tp = int(tile_patches) if tile_patches else biggest.default_tile_patches()
            self.tile_plans = [tab.tile_plan(tp) for tab in tables]
        arr = (FlTraj * self.n_traj)()
        colour = os.environ.get("FLUIDGRID_COLOUR_STAGED", "0") == "1" and ppx % 32 == 0 and py == 16
        for i, (tr, tab, t0) in enumerate(zip(trajs, tables, t0s)):
            tpl = self.tile_plans[i] if self.tile_plans else None
            idx_slot, node_slot = tab.idx_slot, tr.plan.node_slot_d
            if colour and use_slots and idx_slot is not None and tpl is None:
                node_slot, idx_slot = tab.coloured_slots(tr.prs_stride)
            fmt = 0
            if compact_idx and not colour and use_slots and idx_slot is not None and tpl is None and tr.prs_stride <= 65536 and ppx % 128 == 0:
                idx_slot, fmt = tab.idx_slot16(tr.prs_stride), 1        # 8-byte table records for the staged kernel
            arr[i] = FlTraj(tr.vel_buf.data_ptr(), tr.prs_buf.data_ptr(), tab.idx.data_ptr(), tab.w.data_ptr(),
                            idx_slot.data_ptr() if use_slots and idx_slot is not None else 0,
                            node_slot.data_ptr() if use_slots and idx_slot is not None else 0,
                            self.states[i].data_ptr(), self.mask[i].data_ptr() if want_mask else 0,
                            tr.plan.n_nodes, int(t0), int(interval), self.n_frames, tr.vel_stride, tr.prs_stride,
                            tpl.idx_tile.data_ptr() if tpl else 0, tpl.tile_nodes.data_ptr() if tpl else 0,
                            tpl.tile_desc.data_ptr() if tpl else 0, tpl.tile_patches.data_ptr() if tpl else 0,
                            tpl.tile_quads.data_ptr() if tpl else 0, tpl.tile_qslots.data_ptr() if tpl else 0,
                            tpl.n_tiles if tpl else 0, tpl.max_tile_nodes if tpl else 0, min(tpl.tp, tab.n_patches) if tpl else 0, fmt)
        self.host_desc = arr
        raw = np.frombuffer(bytes(arr), dtype=np.uint8).copy()
        self.desc = torch.from_numpy(raw).to(dev)
        self._keep = (trajs, tables)   # keep the device buffers alive

    def run(self, personality: Personality, normalize=True, means=None, stds=None, force_gather=False, force_staged=False,
            force_ring=False):
        """Enqueue the fused kernel on the current stream; returns (states, mask) device tensors."""
        flags = (FL_MASK_AWARE_NORM if personality.mask_aware_norm else 0) | (0 if normalize else FL_NO_NORM)
        flags |= _lib.FL_FORCE_GATHER if force_gather else 0
        flags |= _lib.FL_FORCE_STAGED if force_staged else 0
        flags |= _lib.FL_FORCE_RING if force_ring else 0          # experimental frame-ring kernel (csrc/fl_ring.cu)
        m = (ctypes.c_float * 3)(*(means if means is not None else personality.means))
        s = (ctypes.c_float * 3)(*(stds if stds is not None else personality.stds))
        tab = self.tab0
        with torch.cuda.device(self.device):
            check(load().fl_interp_patchify_dev(ctypes.c_void_p(self.desc.data_ptr()), self.host_desc, self.n_traj,
                                                tab.n_patches, tab.px, tab.py, m, s, flags, stream_ptr()),
                  "fl_interp_patchify")
        return self.states, self.mask


def interp_patchify(traj: DeviceTrajectory, step_num: int, seq_len: int, seq_interval: int, patch_size,
                    personality: Personality, normalize=True, means=None, stds=None, force_gather=False, force_staged=False,
                    tile_patches=None, force_ring=False, stride=None, pad=True, compact_idx=None):
    """One trajectory, frames step_num, step_num+interval, ... -> (states (T,L,3,px,py) f32,
    mask (T,L,px,py) u8, table).  `stride` / `pad`: the data sets' unfold stride (default: the patch size) and padding switch."""
    _lib.require_cuda()
    tab = traj.plan.patch_table(patch_size, personality.crop_patches, personality.flip_y, stride, pad)
    batch = TrajBatch([traj], [tab], [step_num], seq_interval, seq_len, tile_patches=tile_patches, compact_idx=compact_idx)
    states, mask = batch.run(personality, normalize, means, stds, force_gather, force_staged, force_ring)
    return states[0], mask[0], tab


class HostPipeline:
    """Host buffers in, host buffers out: the call a caller without device-resident data makes (the reference's
    DataLoader workers hand CPU tensors to the trainer, `src/dataloader/simple_dataloader.py:72-102`).

    `depth` device slots and three streams: while the kernel works on trajectory i, the node fields of trajectory i+1
    are on their way up and the states of trajectory i-1 on their way down, so a steady stream of trajectories runs at
    the speed of the slower PCIe direction (the output: 13 bytes per pixel-frame) instead of the sum of all three.
    Host tensors must be pinned for the copies to be asynchronous; `run` does not synchronise -- call `wait()` (or
    synchronise the device) before reading the host outputs.
    """

    def __init__(self, plans, tables, personality: Personality, n_steps: int, t0: int = 0, interval: int = 1,
                 n_frames=None, depth: int = 2, want_mask: bool = True):
        _lib.require_cuda()
        if len(plans) != len(tables) or not plans:
            raise ValueError("plans and tables must be non-empty and of equal length")
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.personality, self.n = personality, len(plans)
        self.n_frames = int(n_frames) if n_frames is not None else (int(n_steps) - int(t0) + int(interval) - 1) // int(interval)
        dev = plans[0].device
        self.device, self.depth, self.want_mask = dev, min(depth, self.n), want_mask
        tab0 = tables[0]
        L, px, py = tab0.n_patches, tab0.px, tab0.py
        max_stride = max((pl.n_nodes + 3) // 4 * 4 for pl in plans)
        self.s_in, self.s_run, self.s_out = (torch.cuda.Stream(dev) for _ in range(3))
        self.slots = []
        for _ in range(self.depth):
            self.slots.append({
                "vel": torch.zeros((n_steps, 2 * max_stride), dtype=torch.float32, device=dev),
                "prs": torch.zeros((n_steps, max_stride), dtype=torch.float32, device=dev),
                "states": torch.empty((1, self.n_frames, L, 3, px, py), dtype=torch.float32, device=dev),
                "mask": torch.empty((1, self.n_frames, L, px, py), dtype=torch.uint8, device=dev) if want_mask else None,
                "in_done": torch.cuda.Event(), "run_done": torch.cuda.Event(), "out_done": torch.cuda.Event(),
            })
        # one descriptor per (trajectory, slot it will use): the device views of that slot in the trajectory's own pitch
        self.items = []
        for i, (pl, tab) in enumerate(zip(plans, tables)):
            sl = self.slots[i % self.depth]
            tr = DeviceTrajectory.__new__(DeviceTrajectory)
            N = pl.n_nodes
            tr.prs_stride = (N + 3) // 4 * 4
            tr.vel_stride = 2 * tr.prs_stride
            tr.vel_buf = sl["vel"].view(-1)[: n_steps * tr.vel_stride].view(n_steps, tr.vel_stride)
            tr.prs_buf = sl["prs"].view(-1)[: n_steps * tr.prs_stride].view(n_steps, tr.prs_stride)
            tr.velocity = tr.vel_buf[:, :2 * N].view(n_steps, N, 2)
            tr.pressure = tr.prs_buf[:, :N].unsqueeze(-1)
            tr.plan, tr.n_steps = pl, int(n_steps)
            batch = TrajBatch([tr], [tab], [t0], interval, self.n_frames, want_mask=want_mask, out=(sl["states"], sl["mask"]))
            self.items.append((tr, batch, sl))
        torch.cuda.current_stream(dev).synchronize()      # slot buffers were zero-filled on the current stream
        self._first = True

    def run(self, h_velocity, h_pressure, h_states=None, h_mask=None, normalize=True, on_device=None):
        """h_velocity[i] (T, N_i, 2) / h_pressure[i] (T, N_i, 1) pinned host tensors -> h_states[i] (n_frames, L, 3, px, py),
        h_mask[i] (n_frames, L, px, py) pinned host tensors.  Everything is enqueued asynchronously.

        A consumer that lives on the GPU passes `on_device(i, states, mask)` instead of host outputs: it is called with
        the device slot of trajectory i while the compute stream is current (enqueue the consumer's work there; the slot
        is reused `depth` trajectories later)."""
        if len(h_velocity) != self.n or len(h_pressure) != self.n:
            raise ValueError(f"expected {self.n} trajectories")
        if on_device is None:
            if h_states is None or len(h_states) != self.n:
                raise ValueError(f"expected {self.n} host output tensors")
            if self.want_mask and (h_mask is None or len(h_mask) != self.n):
                raise ValueError("h_mask missing")
        with torch.cuda.device(self.device):
            for i, (tr, batch, sl) in enumerate(self.items):
                with torch.cuda.stream(self.s_in):
                    if not self._first or i >= self.depth:
                        self.s_in.wait_event(sl["run_done"])          # the slot's previous kernel has consumed its inputs
                    tr.velocity.copy_(h_velocity[i], non_blocking=True)
                    tr.pressure.copy_(h_pressure[i], non_blocking=True)
                    sl["in_done"].record(self.s_in)
                with torch.cuda.stream(self.s_run):
                    self.s_run.wait_event(sl["in_done"])
                    if not self._first or i >= self.depth:
                        self.s_run.wait_event(sl["out_done"])         # the slot's previous outputs have left
                    batch.run(self.personality, normalize)
                    if on_device is not None:
                        on_device(i, sl["states"][0], sl["mask"][0] if self.want_mask else None)
                    sl["run_done"].record(self.s_run)
                with torch.cuda.stream(self.s_out):
                    self.s_out.wait_event(sl["run_done"])
                    if on_device is None:
                        h_states[i].copy_(sl["states"][0], non_blocking=True)
                        if self.want_mask:
                            h_mask[i].copy_(sl["mask"][0], non_blocking=True)
                    sl["out_done"].record(self.s_out)
            self._first = False

    def wait(self):
        self.s_out.synchronize()
