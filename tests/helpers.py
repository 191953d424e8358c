"""Shared builders for the tests: seeded synthetic trajectories and the oracle's view of them."""
import functools

import numpy as np

from fluid_llm_b200 import synth
from oracle import pipeline as P

PATCH = (16, 16)


@functools.lru_cache(maxsize=None)
def trajectory(kind, T=8, mesh_seed=0, field_seed=1):
    return synth.make_trajectory(kind, T, mesh_seed=mesh_seed, field_seed=field_seed)


def personality_of(kind):
    return "airfoil" if kind == "airfoil" else "cylinder"


def oracle_ds_get(kind, step, seq_len, interval=1, sem="1.26", normalize=True, T=8, mesh_seed=0, field_seed=1,
                  means=None, stds=None):
    tr = trajectory(kind, T, mesh_seed, field_seed)
    return P.ds_get(tr, step, seq_len, interval, 238, PATCH, personality_of(kind), normalize_ds=normalize,
                    numpy_semantics=sem, return_all=True, means=means, stds=stds)


def tie_mesh():
    """Hand-built mesh whose grid points hit vertices, horizontal / vertical / oblique interior
    edges, boundary edges and a hole; some triangles clockwise."""
    # 5 x 5 lattice on [0,4]^2, each square split along alternating diagonals, centre square removed
    xs, ys = np.meshgrid(np.arange(5.0), np.arange(5.0), indexing="ij")
    pos = np.stack([xs.ravel(), ys.ravel()], axis=1).astype(np.float32)
    idx = np.arange(25).reshape(5, 5)
    tris = []
    for i in range(4):
        for j in range(4):
            if (i, j) in ((1, 1), (2, 2)):
                continue   # holes
            a, b, c, d = idx[i, j], idx[i + 1, j], idx[i + 1, j + 1], idx[i, j + 1]
            if (i + j) % 2 == 0:
                tris += [(a, b, c), (a, d, c)]      # second one clockwise
            else:
                tris += [(a, b, d), (b, d, c)]      # second one clockwise
    return pos, np.array(tris, dtype=np.int32)
