"""Dataset statistics on the GPU, merged across ranks with one small collective.

Mirror of `/root/reference/max/compute_ds_stats.py:20-34,52-62`: per-channel running
(count, mean, M2) of states and diffs over unmasked pixels, population std = sqrt(M2 / n).
The reference's mixed float32/float64 accumulation depends on the NumPy version; here the
aggregates are defined in float64 (csrc/fl_stats.cu), reduced with warp shuffles and merged in a
fixed order, so every rank ends up with bit-identical numbers.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import check, load, ptr, stream_ptr

N_ACC = 6   # state ch0..2, diff ch0..2


def update_variance_batch(existingAggregate, newValues):
    """compute_ds_stats.py:20-30 (host helper kept for API parity; float64)."""
    (count, mean, M2) = existingAggregate
    newValues = np.asarray(newValues.detach().cpu() if torch.is_tensor(newValues) else newValues, dtype=np.float64)
    newCount = count + len(newValues)
    delta = newValues - mean
    mean = mean + np.sum(delta) / newCount
    delta2 = newValues - mean
    M2 = M2 + np.sum(delta * delta2)
    return (newCount, mean, M2)


def get_std(existingAggregate):
    """compute_ds_stats.py:33-34."""
    return np.sqrt(existingAggregate[2] / existingAggregate[0])


def ds_stats(states, mask):
    """states f32 (T, L, 3, px, py), mask u8/bool (T, L, px, py) on the device -> float64 (6, 3) device
    tensor of (n, mean, M2) for {state ch0..2, diff ch0..2}; sample t uses states[t], states[t+1]-states[t]
    where mask[t+1] is clear (the 5-tuple's input_states / diffs / masks, simple_dataloader.py:93-100)."""
    _lib.require_cuda()
    if not states.is_cuda or states.dtype != torch.float32 or states.dim() != 5 or states.shape[2] != 3:
        raise ValueError("ds_stats: states must be a CUDA float32 tensor of shape (T, L, 3, px, py)")
    T, L, _, px, py = states.shape
    if tuple(mask.shape) != (T, L, px, py):
        raise ValueError(f"ds_stats: mask shape {tuple(mask.shape)} != {(T, L, px, py)}")
    m = mask.view(torch.uint8) if mask.dtype == torch.bool else mask.to(torch.uint8)
    lib = load()
    ws_bytes = int(lib.fl_stats_workspace_bytes())
    with torch.cuda.device(states.device):
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=states.device)
        agg = torch.empty((N_ACC, 3), dtype=torch.float64, device=states.device)
        check(lib.fl_ds_stats(ptr(states.contiguous()), ptr(m.contiguous()), T, L, px, py, ptr(agg), ptr(ws), ws_bytes,
                              stream_ptr()), "fl_ds_stats")
    return agg


def merge_stats(parts):
    """Fixed-order Chan merge of aggregates (P, 6, 3) -> (6, 3), on the device."""
    parts = parts.contiguous()
    out = torch.empty((N_ACC, 3), dtype=torch.float64, device=parts.device)
    with torch.cuda.device(parts.device):
        check(load().fl_stats_merge(ptr(parts), parts.shape[0], ptr(out), stream_ptr()), "fl_stats_merge")
    return out


def gather_stats(agg, group=None):
    """All-gather the (6, 3) aggregates of every rank -> (world, 6, 3) in rank order (works on any backend)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return agg.unsqueeze(0)
    world = dist.get_world_size(group)
    parts = [torch.empty_like(agg) for _ in range(world)]
    dist.all_gather(parts, agg.contiguous(), group=group)
    return torch.stack(parts)


def all_reduce_stats(agg, group=None):
    """The path's only collective: all-gather the 144-byte aggregates of every rank (NCCL over
    NVLink; latency-bound) and merge them in rank order on every rank -> identical results."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return agg
    return merge_stats(gather_stats(agg, group))


def mean_std(agg):
    """(6, 3) aggregate -> (means[6], stds[6]) host float64 arrays (population std)."""
    a = agg.detach().cpu().numpy()
    return a[:, 1].copy(), np.sqrt(a[:, 2] / np.maximum(a[:, 0], 1.0))
