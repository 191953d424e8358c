"""ORACLE test scaffolding: run the reference's own CALLERS of the data path -- `get_data_loader`
(`/root/reference/src/utils_model.py:9-45`) and the rollout methods `_gen_step` / `_generate` / `gen_seq`
(`/root/reference/src/models/model.py:154-233`) -- unmodified, in the development container.

`utils_model` imports as is behind oracle/stubs.  `models/model.py` imports peft / accelerate / transformers at module level
(none of which the three methods use), so the methods' source is taken from the file with `ast` and compiled as it stands into
a bare class whose module globals are exactly the names the methods reference: torch, deque and the patch ops.  Nothing is
copied into the repository; the text is read from /root/reference at run time.
"""
import ast
import os
from collections import deque

import torch

from . import ref_import

METHODS = ("_gen_step", "_generate", "gen_seq")


def reference_rollout_class(img_to_patch, patch_to_img):
    """-> a class carrying the reference's unmodified rollout methods, bound to the given patch ops."""
    path = os.path.join(ref_import.REF_ROOT, "src", "models", "model.py")
    with open(path) as f:
        tree = ast.parse(f.read(), path)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "MultivariateTimeLLM")
    keep = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in METHODS]
    assert len(keep) == len(METHODS)
    mod = ast.Module(body=[ast.ClassDef(name="RefRollout", bases=[], keywords=[], body=keep, decorator_list=[], type_params=[])],
                     type_ignores=[])
    ast.fix_missing_locations(mod)
    glob = {"torch": torch, "deque": deque, "img_to_patch": img_to_patch, "patch_to_img": patch_to_img}
    exec(compile(mod, path, "exec"), glob)
    return glob["RefRollout"]


def stub_forward(ds_props):
    """A deterministic stand-in for the backbone: element-wise fp32 arithmetic only (the same bits on CPU and GPU), built
    from views -- not from the patch ops under test."""
    def forward_see_init(states, position_ids):
        bs, T, L, C, px, py = states.shape
        nbx, nby = ds_props.Nx_patch, ds_props.Ny_patch
        img = states.view(bs, T, nbx, nby, C, px, py).permute(0, 1, 4, 2, 5, 3, 6).reshape(bs, T, C, nbx * px, nby * py)
        X = torch.arange(nbx * px, device=states.device).view(1, 1, 1, -1, 1)
        Y = torch.arange(nby * py, device=states.device).view(1, 1, 1, 1, -1)
        t = position_ids[:, :, 0, 2].to(torch.int64).view(bs, T, 1, 1, 1)
        bump = ((X * 7 + Y * 3 + t) % 13).to(torch.float32) * 0.001
        return img * 0.25 + bump
    return forward_see_init
