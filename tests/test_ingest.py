"""CPU: the pickle ingest path (fluid_llm_b200/ingest.py, _ingest_worker.py) against a plain `pickle.load`."""
import copy
import pickle

import numpy as np
import pytest
import torch

from oracle import pipeline as P

from helpers import trajectory

from fluid_llm_b200 import _ingest_worker as W
from fluid_llm_b200._plan_host import prepare_plan
from fluid_llm_b200.ingest import PickleIngest


@pytest.fixture(autouse=True)
def _fresh_lazy_reader():
    W._declined = 0          # (a process stops trying the lazy reader after files it had to decline)
    yield
    W._declined = 0


def _reference_pickle(tr, path, protocol=None, cells_dtype=np.int16):
    """The reference's writer (max/ds_download/torch_MGN.py:66-93): velocity, pressure, density, then the static entries."""
    d = {"velocity": tr["velocity"], "pressure": tr["pressure"], "density": np.ones_like(tr["pressure"]),
         "cells": tr["cells"].astype(cells_dtype), "mesh_pos": tr["mesh_pos"],
         "node_type": np.zeros((tr["mesh_pos"].shape[0], 1), dtype=np.int32)}
    with open(path, "wb") as f:
        pickle.dump(d, f, protocol=protocol)
    return d


@pytest.mark.parametrize("protocol", [2, 4, 5])
def test_lazy_unpickle_equals_pickle_load(tmp_path, protocol):
    tr = trajectory("cylinder", 60, 0, 3)
    path = str(tmp_path / "a.pkl")
    d = _reference_pickle(tr, path, protocol)
    got = W.unpickle_lazy(path, ("mesh_pos", "cells", "velocity", "pressure"))
    if protocol == 2:        # array data travels as a latin-1 string there: NumPy copies it, the lazy reader declines
        assert got is None
    else:
        for k in ("velocity", "pressure"):
            assert d[k].nbytes >= W.LAZY_BYTES and not got[k].flags.owndata and not got[k].flags.writeable   # a view of the file
        for k, a in got.items():
            assert a.dtype == d[k].dtype and a.shape == d[k].shape and np.array_equal(a, d[k])
    pos, cells, vel, prs = W.load_trajectory(path, False)
    assert cells.dtype == np.int32 and np.array_equal(cells, tr["cells"]) and np.array_equal(pos, tr["mesh_pos"])
    assert np.array_equal(vel, tr["velocity"]) and np.array_equal(prs, tr["pressure"])


def test_lazy_unpickle_declines_what_it_does_not_know(tmp_path):
    tr = trajectory("cylinder", 60, 0, 3)
    for i, d in enumerate(({"velocity": torch.from_numpy(tr["velocity"]), "pressure": tr["pressure"], "cells": tr["cells"],
                            "mesh_pos": tr["mesh_pos"]},                                            # a tensor, not an array
                           {"velocity": tr["velocity"][:, :, ::-1], "pressure": tr["pressure"], "cells": tr["cells"],
                            "mesh_pos": tr["mesh_pos"]},                                            # pickled as a copy: fine either way
                           [tr["velocity"]])):                                                     # not a dict
        path = str(tmp_path / f"{i}.pkl")
        with open(path, "wb") as f:
            pickle.dump(d, f)
        W._declined = 0
        got = W.unpickle_lazy(path, ("mesh_pos", "cells", "velocity", "pressure"))
        if i == 1:
            assert got is not None and np.array_equal(got["velocity"], tr["velocity"][:, :, ::-1])
        else:
            assert got is None
    # ... and load_trajectory falls back to the plain unpickler
    pos, cells, vel, prs = W.load_trajectory(str(tmp_path / "0.pkl"), False)
    assert np.array_equal(vel, tr["velocity"])
    with open(tmp_path / "short.pkl", "wb") as f:
        f.write((tmp_path / "1.pkl").read_bytes()[:500000])          # truncated inside a payload
    W._declined = 0
    assert W.unpickle_lazy(str(tmp_path / "short.pkl"), ("velocity",)) is None        # a broken file ...
    assert W.unpickle_lazy(str(tmp_path / "nowhere.pkl"), ("velocity",)) is None and W._declined == 0     # ... says nothing about the format
    assert W.unpickle_lazy(str(tmp_path / "0.pkl"), ("velocity",)) is None
    assert W.unpickle_lazy(str(tmp_path / "2.pkl"), ("velocity",)) is None and W._declined == 2
    assert W.unpickle_lazy(str(tmp_path / "1.pkl"), ("velocity",)) is None        # two files of another kind in a row: it stops trying
    with pytest.raises(Exception):
        W.load_trajectory(str(tmp_path / "short.pkl"), False)


def test_airfoil_crop_in_the_worker_equals_the_oracle(tmp_path):
    tr = trajectory("airfoil", 40, 1, 5)
    path = str(tmp_path / "a.pkl")
    _reference_pickle(tr, path, cells_dtype=np.int32)
    pos, cells, vel, prs = W.load_trajectory(path, True)
    nmask, pos_o, faces_o = P.airfoil_crop(tr["mesh_pos"], tr["cells"])
    assert np.array_equal(pos, pos_o) and np.array_equal(cells, faces_o)
    assert np.array_equal(vel, tr["velocity"][:, nmask]) and np.array_equal(prs, tr["pressure"][:, nmask])


def test_pool_delivers_every_file_in_the_device_pitch(tmp_path):
    trajs = [trajectory("cylinder", 30, s, 7 + s) for s in range(5)]
    paths = []
    for i, tr in enumerate(trajs):
        paths.append(str(tmp_path / f"{i}.pkl"))
        _reference_pickle(copy.deepcopy(tr), paths[-1])
    pool = PickleIngest(workers=2, slot_bytes=4 << 20, plan_resolution=238, numpy_semantics="2.x")
    try:
        pool.submit(paths)
        for i in (3, 0, 4, 1, 2):                       # consumption order is free; only 4 slots for 5 files
            pos, cells, vel, prs, release, plan = pool.take(paths[i])
            tr = trajs[i]
            want = prepare_plan(tr["mesh_pos"], tr["cells"], 238, "2.x")          # the workers prepared the host side of the mesh plan
            assert set(plan) == set(want) and plan["numpy_semantics"] == "2.x" and plan["n_degenerate"] == 0
            for k in ("pos32", "tri", "ax", "ay", "slots"):
                assert plan[k].dtype == want[k].dtype and np.array_equal(plan[k], want[k])
            T, N = tr["velocity"].shape[:2]
            vs, ps = W.strides(N)
            assert tuple(vel.shape) == (T, vs) and tuple(prs.shape) == (T, ps)
            assert np.array_equal(vel[:, :2 * N].numpy().reshape(T, N, 2), tr["velocity"]) and not vel[:, 2 * N:].any()
            assert np.array_equal(prs[:, :N].numpy().reshape(T, N, 1), tr["pressure"]) and not prs[:, N:].any()
            assert np.array_equal(pos, tr["mesh_pos"]) and np.array_equal(cells, tr["cells"])
            release()
        # a file larger than a slot is loaded in the calling process
        big = trajectory("cylinder", 400, 0, 1)
        _reference_pickle(big, str(tmp_path / "big.pkl"))
        assert pool.grow_to == 0
        pos, cells, vel, prs, release, plan = pool.take(str(tmp_path / "big.pkl"))
        assert plan is None and pool.grow_to >= 400 * 12 * pos.shape[0]      # ... and the pool asks for larger slots
        assert np.array_equal(vel[:, :2 * pos.shape[0]].numpy().reshape(400, -1, 2), big["velocity"])
        release()
        with pytest.raises(RuntimeError, match="ingest worker failed"):
            pool.take(str(tmp_path / "missing.pkl"))
    finally:
        pool.close()


def test_lazy_reader_on_arrays_of_many_kinds(tmp_path):
    """Whatever the arrays look like -- dtypes, byte order, Fortran order, views, zero-size, below and above the lazy threshold --
    the lazy reader either returns exactly what `pickle.load` returns or declines."""
    rng = np.random.default_rng(0)
    big = rng.standard_normal((300, 700))                    # 1.7 MB as float64
    cases = {
        "f32": big.astype(np.float32), "f64": big, "i16": (big * 100).astype(np.int16), "i64": (big * 100).astype(np.int64),
        "u8": (np.abs(big) * 50).astype(np.uint8), "bool": big > 0, "c64": (big + 1j * big).astype(np.complex64),
        "swapped": big.astype(np.float32).astype(">f4"), "fortran": np.asfortranarray(big.astype(np.float32)),
        "view": big.astype(np.float32)[::2, 5:-5], "small": big[:3, :3].astype(np.float32), "empty": np.zeros((0, 4), np.float32),
        "one_d": big.astype(np.float32).reshape(-1), "five_d": big.astype(np.float32).reshape(3, 10, 10, 7, 100),
    }
    keys = tuple(cases)
    for protocol in (3, 4, 5):
        path = str(tmp_path / f"p{protocol}.pkl")
        with open(path, "wb") as f:
            pickle.dump(cases, f, protocol=protocol)
        W._declined = 0
        got = W.unpickle_lazy(path, keys)
        with open(path, "rb") as f:
            ref = pickle.load(f)
        if got is None:
            continue
        for k in keys:
            assert got[k].dtype == ref[k].dtype and got[k].shape == ref[k].shape and np.array_equal(got[k], ref[k]), (protocol, k)
    # every array alone: the plain ones are served lazily, the awkward ones are declined or copied -- never wrong
    served = 0
    for k, a in cases.items():
        path = str(tmp_path / f"{k}.pkl")
        with open(path, "wb") as f:
            pickle.dump({"x": a}, f)
        W._declined = 0
        got = W.unpickle_lazy(path, ("x",))
        if got is not None:
            assert got["x"].dtype == a.dtype and got["x"].shape == a.shape and np.array_equal(got["x"], a), k
            served += not got["x"].flags.owndata and a.nbytes >= W.LAZY_BYTES
    assert served >= 6
