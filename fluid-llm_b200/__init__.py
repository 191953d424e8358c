"""fluid-llm_b200: FLUID-LLM's per-timestep field data path, B200-native.

Python host side above the C ABI of libfluidgrid.so (include/fluidgrid.h).  Module names follow the
reference files they mirror (`src/dataloader/mesh_utils.py`, `simple_dataloader.py`, `airfoil_ds.py`,
`ds_props.py`, `src/utils_model.py`, `eagle/Dataloader/IMG_Eagle.py`, `max/compute_ds_stats.py`).
Import name: `fluid_llm_b200` (the directory is `fluid-llm_b200/`; `fluid_llm_b200/` is a shim).
"""
from .ds_props import DSProps  # noqa: F401
from ._lib import FluidGridError, load  # noqa: F401

__all__ = ["DSProps", "FluidGridError", "load"]
