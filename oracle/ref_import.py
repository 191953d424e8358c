"""ORACLE test scaffolding: import the reference's UNMODIFIED Python from /root/reference.

Only usable in the development container (the GPU box has no /root/reference).  Puts
oracle/stubs (matplotlib -> our restatement, cprint, natsort) and /root/reference/src on
sys.path, exactly the layout `run_training.sh` uses (`src/` is the import root), and hands back
the reference's own modules.  Nothing is copied; the reference code runs where it lies.
Because the container has NumPy 2.x, what runs is the reference under NEP-50 promotion
("2.x" numpy_semantics of oracle/pipeline.py), not under its pinned NumPy 1.26.3.
"""
from __future__ import annotations

import os
import pickle
import sys
import tempfile

REF_ROOT = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "src", "dataloader"))


def modules():
    """-> dict of the reference's modules on the path (imported once)."""
    if not available():
        raise RuntimeError("/root/reference is not present")
    repo = os.path.dirname(HERE)
    for p in (os.path.join(REF_ROOT, "src"), os.path.join(HERE, "stubs"), repo):
        if p in sys.path:
            sys.path.remove(p)
    sys.path[:0] = [os.path.join(HERE, "stubs"), os.path.join(REF_ROOT, "src"), repo]
    import importlib
    out = {}
    for name in ("dataloader.mesh_utils", "dataloader.simple_dataloader", "dataloader.airfoil_ds",
                 "dataloader.ds_props", "utils_model", "_triinterpolate"):
        out[name.split(".")[-1]] = importlib.import_module(name)
    # eagle/Dataloader/IMG_Eagle.py needs no stubs; load it by path (eagle/ is not a package root)
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_IMG_Eagle", os.path.join(REF_ROOT, "eagle/Dataloader/IMG_Eagle.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out["IMG_Eagle"] = mod
    return out


def write_pickles(trajs, directory=None):
    """Write trajectory dicts as 0.pkl, 1.pkl, ... (the loaders read `save_files[1]` at init)."""
    directory = directory or tempfile.mkdtemp(prefix="fl_ref_ds_")
    for i, tr in enumerate(trajs):
        with open(os.path.join(directory, f"{i}.pkl"), "wb") as f:
            pickle.dump(tr, f)
    return directory
