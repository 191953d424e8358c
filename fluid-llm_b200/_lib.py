"""ctypes binding of libfluidgrid.so (include/fluidgrid.h).  There is no CPU fallback: if the
library is missing or no CUDA device is visible, compute calls raise."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_double, c_float, c_int, c_int32, c_size_t, c_uint, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libfluidgrid.so")

FL_FLIP_Y = 1
FL_MASK_AWARE_NORM = 2
FL_NO_NORM = 4
FL_FORCE_GATHER = 8
FL_FORCE_STAGED = 16
FL_FORCE_TILED = 32
FL_FORCE_RING = 64
FL_NO_PAD = 128


class FlTraj(ctypes.Structure):
    _fields_ = [("d_velocity", c_void_p), ("d_pressure", c_void_p), ("d_idx", c_void_p), ("d_w", c_void_p),
                ("d_idx_slot", c_void_p), ("d_node_slot", c_void_p),
                ("d_states", c_void_p), ("d_mask", c_void_p),
                ("n_nodes", c_int32), ("t0", c_int32), ("interval", c_int32), ("n_frames", c_int32),
                ("vel_stride", c_int32), ("prs_stride", c_int32),
                ("d_idx_tile", c_void_p), ("d_tile_nodes", c_void_p), ("d_tile_desc", c_void_p), ("d_tile_patches", c_void_p),
                ("d_tile_quads", c_void_p), ("d_tile_qslots", c_void_p), ("n_tiles", c_int32), ("max_tile_nodes", c_int32), ("max_tile_patches", c_int32), ("idx_slot_format", c_int32)]


class FluidGridError(RuntimeError):
    pass


# name -> (restype, argtypes); every symbol include/fluidgrid.h declares
SIGNATURES = {
    "fl_abi_version": (c_int, []),
    "fl_last_error": (ctypes.c_char_p, []),
    "fl_device_count": (c_int, []),
    "fl_locate_workspace_bytes": (c_size_t, [c_int, c_int]),
    "fl_locate": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                          c_void_p, c_void_p, c_size_t, c_void_p]),
    "fl_locate_async": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "fl_plan_patch_table": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_uint, c_void_p, c_void_p,
                                    POINTER(c_int), POINTER(c_int), c_void_p, c_void_p, c_void_p]),
    "fl_pack_idx16": (c_int, [c_void_p, ctypes.c_long, c_int, c_void_p, c_void_p]),
    "fl_interp_patchify": (c_int, [POINTER(FlTraj), c_int, c_int, c_int, c_int, POINTER(c_float), POINTER(c_float),
                                   c_uint, c_void_p]),
    "fl_interp_patchify_dev": (c_int, [c_void_p, POINTER(FlTraj), c_int, c_int, c_int, c_int, POINTER(c_float),
                                       POINTER(c_float), c_uint, c_void_p]),
    "fl_last_interp_kernel": (ctypes.c_char_p, []),
    "fl_to_grid": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "fl_interp_frames": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                 c_int, POINTER(c_float), POINTER(c_float), c_uint, c_void_p, c_void_p, c_void_p]),
    "fl_patch_to_img": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "fl_img_to_patch": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "fl_rollout_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                c_int, c_void_p]),
    "fl_sample_assemble": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "fl_pos_add_ring": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int,
                                c_int, c_int, c_int, c_void_p]),
    "fl_affine_channels": (c_int, [c_void_p, c_void_p, ctypes.c_long, c_int, POINTER(c_float), POINTER(c_float), c_int, c_void_p]),
    "fl_grid2mesh": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_float,
                             c_double, c_double, c_void_p, c_void_p]),
    "fl_dyn_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "fl_dyn_capacity": (c_int, [c_int, c_int, c_int, c_int, c_size_t]),
    "fl_dyn_interp_patchify": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p,
                                       c_int, c_int, c_int, c_int, c_int, POINTER(c_float), POINTER(c_float), c_uint,
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "fl_stats_workspace_bytes": (c_size_t, []),
    "fl_ds_stats": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "fl_stats_merge": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "fl_cast_bf16": (c_int, [c_void_p, c_void_p, ctypes.c_long, c_void_p]),
    "fl_cast_bf16_rows": (c_int, [c_void_p, c_void_p, ctypes.c_long, c_int, c_int, c_void_p]),
    "fl_patch_embed": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
}

_lib = None


def load():
    """Load libfluidgrid.so (built in-tree by build.py) and declare every signature."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FluidGridError(f"{LIB_PATH} is missing: run `python fluid-llm_b200/build.py` (nvcc, sm_100a). "
                             "There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header and library out of sync
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    """Map a non-zero return code onto the exception type the reference raises for that condition."""
    if rc == 0:
        return
    msg = load().fl_last_error().decode("utf-8", "replace")
    if rc in (-1, -2, -4):
        raise ValueError(f"{what}: {msg}" if what else msg)
    if rc == -3:
        raise MemoryError(f"{what}: {msg}" if what else msg)
    raise FluidGridError(f"{what}: CUDA error {rc}: {msg}" if what else f"CUDA error {rc}: {msg}")


def require_cuda():
    import torch
    if not torch.cuda.is_available() or load().fl_device_count() < 1:
        raise FluidGridError("no CUDA device visible: the fluidgrid data path runs on B200 only (no CPU fallback)")


def stream_ptr():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """device pointer of a torch tensor (None -> NULL)"""
    return c_void_p(0 if t is None else t.data_ptr())
