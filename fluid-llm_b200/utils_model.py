"""Patch <-> image ops and the rollout-step glue on the GPU.

Mirror of `/root/reference/src/utils_model.py:77-109` (`patch_to_img`, `img_to_patch`) and of the
per-step glue of `/root/reference/src/models/model.py:164,206,210`.  Both permutations are
dtype-preserving and differentiable (the reference uses them inside the training graph,
src/trainer.py:95-98): the backward of one is the other.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, load, ptr, stream_ptr
from .ds_props import DSProps


def _perm(fn_name, src, dst, B, n_bx, n_by, C, px, py):
    with torch.cuda.device(src.device):
        check(getattr(load(), fn_name)(ptr(src), ptr(dst), B, n_bx, n_by, C, px, py, src.element_size(), stream_ptr()),
              fn_name)


def _check_tensor(t, name):
    if not t.is_cuda:
        raise _lib.FluidGridError(f"{name}: expected a CUDA tensor (the fluidgrid data path has no CPU fallback)")
    if t.element_size() not in (2, 4):
        raise ValueError(f"{name}: only 2- and 4-byte element types are supported, got {t.dtype}")


class _PatchToImg(torch.autograd.Function):
    @staticmethod
    def forward(ctx, patches, n_bx, n_by):
        B, L, C, px, py = patches.shape
        ctx.geom = (n_bx, n_by)
        img = torch.empty((B, C, n_bx * px, n_by * py), dtype=patches.dtype, device=patches.device)
        _perm("fl_patch_to_img", patches.contiguous(), img, B, n_bx, n_by, C, px, py)
        return img

    @staticmethod
    def backward(ctx, g):
        n_bx, n_by = ctx.geom
        B, C, X, Y = g.shape
        px, py = X // n_bx, Y // n_by
        out = torch.empty((B, n_bx * n_by, C, px, py), dtype=g.dtype, device=g.device)
        _perm("fl_img_to_patch", g.contiguous(), out, B, n_bx, n_by, C, px, py)
        return out, None, None


class _ImgToPatch(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, n_bx, n_by):
        B, C, X, Y = img.shape
        px, py = X // n_bx, Y // n_by
        ctx.geom = (n_bx, n_by)
        out = torch.empty((B, n_bx * n_by, C, px, py), dtype=img.dtype, device=img.device)
        _perm("fl_img_to_patch", img.contiguous(), out, B, n_bx, n_by, C, px, py)
        return out

    @staticmethod
    def backward(ctx, g):
        n_bx, n_by = ctx.geom
        B, L, C, px, py = g.shape
        img = torch.empty((B, C, n_bx * px, n_by * py), dtype=g.dtype, device=g.device)
        _perm("fl_patch_to_img", g.contiguous(), img, B, n_bx, n_by, C, px, py)
        return img, None, None


def patch_to_img(patches, ds_props: DSProps):
    """utils_model.py:77-92: (bs, seq_len, N_patch, C, px, py) -> (bs, seq_len, C, tot_px, tot_py)."""
    _check_tensor(patches, "patch_to_img")
    bs, seq_len, N_patch, channel, px, py = patches.shape
    px_patch, py_patch = ds_props.patch_size
    channel = ds_props.channel
    tot_px, tot_py = ds_props.input_tot_size
    if N_patch != ds_props.N_patch or (px, py) != (px_patch, py_patch) or patches.shape[3] != channel:
        raise ValueError(f"patches of shape {tuple(patches.shape)} do not match ds_props "
                         f"(N_patch={ds_props.N_patch}, channel={channel}, patch_size={ds_props.patch_size})")
    img = _PatchToImg.apply(patches.reshape(bs * seq_len, N_patch, channel, px, py), ds_props.Nx_patch, ds_props.Ny_patch)
    return img.view(bs, seq_len, channel, tot_px, tot_py)


def img_to_patch(img, ds_props: DSProps):
    """utils_model.py:95-109: (bs, seq_len, C, tot_px, tot_py) -> (bs, seq_len, N_patch, C, px, py)."""
    _check_tensor(img, "img_to_patch")
    bs, seq_len, channel, tot_px, tot_py = img.shape
    px_patch, py_patch = ds_props.patch_size
    if (tot_px, tot_py) != (ds_props.Nx_patch * px_patch, ds_props.Ny_patch * py_patch) or channel != ds_props.channel:
        raise ValueError(f"image of shape {tuple(img.shape)} does not match ds_props "
                         f"(input_tot_size={ds_props.input_tot_size}, channel={ds_props.channel})")
    out = _ImgToPatch.apply(img.reshape(bs * seq_len, channel, tot_px, tot_py), ds_props.Nx_patch, ds_props.Ny_patch)
    return out.view(bs, seq_len, ds_props.N_patch, channel, px_patch, py_patch)


def rollout_step(last_state, pred_diff_img, mask, ds_props: DSProps, tokens_bf16=False, tokens_out=None):
    """model.py:164,206,210 fused into one kernel (fp32, inference only):

        diffs = img_to_patch(pred_diff_img);  diffs[mask] = 0.;  next_state = last_state + diffs

    last_state (bs, 1, N_patch, C, px, py), pred_diff_img (bs, 1, C, tot_px, tot_py),
    mask bool (bs, 1, N_patch, C, px, py) -> (next_state, diffs), or (next_state, diffs, next_state as bf16) with
    tokens_bf16=True: the tokens the patch embedding takes, written by the same kernel (into `tokens_out`, a contiguous
    bf16 tensor of last_state's element count, if given -- e.g. `RolloutEmbedCache.token_buffer`)."""
    for t, n in ((last_state, "last_state"), (pred_diff_img, "pred_diff_img"), (mask, "mask")):
        if not t.is_cuda:
            raise _lib.FluidGridError(f"rollout_step: {n} must be a CUDA tensor")
    if last_state.dtype != torch.float32 or pred_diff_img.dtype != torch.float32:
        raise ValueError("rollout_step computes in float32 (model.py keeps states in fp32)")
    bs, T, N_patch, C, px, py = last_state.shape
    if mask.shape != last_state.shape or pred_diff_img.shape != (bs, T, C, ds_props.Nx_patch * px, ds_props.Ny_patch * py):
        raise ValueError("rollout_step: shapes of last_state, mask and pred_diff_img do not agree with ds_props")
    last_c, img_c = last_state.contiguous(), pred_diff_img.contiguous()
    m = mask.contiguous().view(torch.uint8) if mask.dtype == torch.bool else mask.to(torch.uint8).contiguous()
    diffs, nxt = torch.empty_like(last_c), torch.empty_like(last_c)
    tok = None
    if tokens_out is not None:
        if tokens_out.dtype != torch.bfloat16 or tokens_out.numel() != last_c.numel() or not tokens_out.is_contiguous() \
                or tokens_out.device != last_c.device:
            raise ValueError("rollout_step: tokens_out must be a contiguous bf16 CUDA tensor with last_state's element count")
        tok, tokens_bf16 = tokens_out, True
    elif tokens_bf16:
        tok = torch.empty(last_c.shape, dtype=torch.bfloat16, device=last_c.device)
    with torch.cuda.device(last_c.device):
        check(load().fl_rollout_step(ptr(img_c), ptr(m), ptr(last_c), ptr(diffs), ptr(nxt), ptr(tok), bs * T, ds_props.Nx_patch,
                                     ds_props.Ny_patch, C, px, py, stream_ptr()), "fl_rollout_step")
    return (nxt, diffs, tok) if tokens_bf16 else (nxt, diffs)


class LookaheadBatchSampler(torch.utils.data.Sampler):
    """A batch sampler that tells the data set which files the NEXT batches will need, so that its ingest pool
    (`ingest.PickleIngest`) unpickles them while the current batch is on the GPU -- what the reference gets from
    `num_workers` processes with `prefetch_factor` batches each (src/utils_model.py:35-36), without moving the CUDA work out of
    the calling process.  Same index order as `BatchSampler(sampler, batch_size, drop_last)`."""

    def __init__(self, sampler, batch_size, drop_last, dataset, lookahead=2):
        self.sampler, self.batch_size, self.drop_last = sampler, int(batch_size), bool(drop_last)
        self.dataset, self.lookahead = dataset, int(lookahead)

    def __len__(self):
        n = len(self.sampler)
        return n // self.batch_size if self.drop_last else -(-n // self.batch_size)

    def __iter__(self):
        idx = list(self.sampler)
        batches = [idx[i:i + self.batch_size] for i in range(0, len(idx), self.batch_size)]
        if batches and self.drop_last and len(batches[-1]) < self.batch_size:
            batches.pop()
        for i, b in enumerate(batches):
            ahead = [j for nb in batches[i:i + 1 + self.lookahead] for j in nb]
            if hasattr(self.dataset, "prefetch"):
                self.dataset.prefetch(ahead)
            yield b


def get_data_loader(config, mode="train", device=None, numpy_semantics=None):
    """utils_model.py:9-45: config -> (DataLoader, DSProps), same dataset selection, same DSProps.

    The datasets produce their samples on the GPU (one kernel launch per DataLoader batch through `__getitems__`), so the
    loader runs in the calling process (`num_workers=0`, no pinning: nothing is on the host to pin) and the default
    collate stacks device tensors.  `config['num_workers']` is accepted and ignored."""
    from torch.utils.data import DataLoader
    from .airfoil_ds import AirfoilDataset
    from .simple_dataloader import MGNDataset
    ds_name = config["load_dir"]
    kw = dict(load_dir=f'{ds_name}/{mode}', resolution=config['resolution'], patch_size=config['patch_size'],
              stride=config['stride'], seq_len=config['seq_len'], seq_interval=config['seq_interval'], mode=mode,
              normalize=config['normalize_ds'], device=device, numpy_semantics=numpy_semantics)
    if (ds_name == "./ds/MGN/cylinder_dataset") | (ds_name == "cylinder"):
        ds = MGNDataset(**kw)
    elif (ds_name == "./ds/MGN/airfoil_dataset") | (ds_name == "airfoil"):
        ds = AirfoilDataset(**kw)
    else:
        raise ValueError(f"Invalid dataset {ds_name}")
    from torch.utils.data import RandomSampler, SequentialSampler
    sampler = RandomSampler(ds) if mode == 'train' else SequentialSampler(ds)          # shuffle=(mode == 'train'), :36
    dl = DataLoader(ds, num_workers=0,
                    batch_sampler=LookaheadBatchSampler(sampler, config['batch_size'], False, ds, config.get('prefetch_batches', 2)))
    ds_props = DSProps(Nx_patch=ds.N_x_patch, Ny_patch=ds.N_y_patch, patch_size=ds.patch_size, seq_len=ds.seq_len - 1)
    return dl, ds_props
