"""CPU: the C-ABI library loads, exports every symbol include/fluidgrid.h declares, and rejects bad
arguments before touching the device (no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "fluidgrid.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fl_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    from fluid_llm_b200 import _lib
    names = _header_functions()
    assert len(names) >= 16
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/fluidgrid.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes SIGNATURES and the header disagree"
    assert lib.fl_abi_version() == 11


def test_struct_layout_matches_header():
    from fluid_llm_b200._lib import FlTraj
    assert ctypes.sizeof(FlTraj) == 8 * 8 + 6 * 4 + 6 * 8 + 4 * 4 == 152
    assert FlTraj.n_nodes.offset == 64 and FlTraj.prs_stride.offset == 84
    assert FlTraj.d_idx_tile.offset == 88 and FlTraj.n_tiles.offset == 136 and FlTraj.max_tile_nodes.offset == 140


def test_argument_errors_without_a_device(lib):
    from fluid_llm_b200 import _lib
    assert lib.fl_locate(None, None, 0, 0, None, None, 0, 0, None, None, None, None, 0, None) == -1
    assert b"null" in lib.fl_last_error()
    assert lib.fl_patch_to_img(None, None, 1, 1, 1, 1, 1, 1, 4, None) == -1
    assert lib.fl_ds_stats(None, None, 2, 1, 1, 1, None, None, 0, None) == -1
    assert lib.fl_grid2mesh(None, None, None, 1, 1, 1, 1, 1, 0.0, 0.0, 1.0, 1.0, None, None) == -1
    assert lib.fl_interp_patchify(None, 1, 1, 16, 16, None, None, 0, None) == -1
    nbx, nby = ctypes.c_int(), ctypes.c_int()
    assert lib.fl_plan_patch_table(None, None, 238, 60, 16, 16, 0, 0, 0, 0, None, None, ctypes.byref(nbx), ctypes.byref(nby), None, None, None) == 0
    assert (nbx.value, nby.value) == (15, 4)
    assert lib.fl_plan_patch_table(None, None, 238, 142, 16, 16, 0, 0, 1, 1, None, None, ctypes.byref(nbx), ctypes.byref(nby), None, None, None) == 0
    assert (nbx.value, nby.value) == (13, 7)
    with pytest.raises(ValueError):
        _lib.check(-1, "x")
    with pytest.raises(MemoryError):
        _lib.check(-3, "x")
    with pytest.raises(_lib.FluidGridError):
        _lib.check(700, "x")
    assert lib.fl_locate_workspace_bytes(1000, 2000) > 0 and lib.fl_stats_workspace_bytes() > 0


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from fluid_llm_b200 import FluidGridError, synth
    from fluid_llm_b200.mesh_utils import MeshPlan, get_mesh_interpolation
    pos, cells = synth.make_mesh("cylinder")
    with pytest.raises(FluidGridError):
        MeshPlan(pos, cells)
    with pytest.raises(FluidGridError):
        get_mesh_interpolation(pos, cells)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "fluid-llm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports the oracle"
                assert "tri_oracle" not in txt
