"""Host side of a mesh plan, NumPy only: what `mesh_utils.MeshPlan` needs before anything touches the GPU.

Kept free of torch so that the ingest workers (`_ingest_worker.py`, separate interpreters that import nothing but NumPy) can
prepare the plan of a trajectory they have just unpickled, ahead of the process that owns the GPU.
"""
import os
from functools import lru_cache

import numpy as np

F32, F64 = np.float32, np.float64


def default_numpy_semantics() -> str:
    """The reference pins NumPy 1.26.3 (environemnt.yml:177); its scalar promotion decides the last
    bit of the grid coordinates.  "1.26" reproduces the pinned environment, "2.x" reproduces the
    reference's code run under NumPy >= 2 (NEP 50)."""
    return os.environ.get("FLUIDGRID_NUMPY_SEMANTICS", "1.26")


def _grid_shape(x_min, x_max, y_min, y_max, grid_res, sem):
    x_min, x_max, y_min, y_max = F32(x_min), F32(x_max), F32(y_min), F32(y_max)
    dx, dy = F32(x_max - x_min), F32(y_max - y_min)
    ratio = F32(min(dx, dy) / max(dx, dy))                       # mesh_utils.py:67-69, float32 scalars
    n_short = int(F64(grid_res) * F64(ratio)) if sem == "1.26" else int(F32(F32(grid_res) * ratio))
    return (int(grid_res), n_short) if dx > dy else (n_short, int(grid_res))   # :71-76


def _grid_axis(start, stop, n, sem):
    # np.mgrid[start:stop:n*1j]: indices * step + start with step = (stop - start) / (n - 1);
    # float64 arithmetic under NumPy 1.26 (legacy promotion of the float32 scalars), float32 under 2.x
    start, stop = F32(start), F32(stop)
    i = np.arange(n)
    if sem == "1.26":
        step = F64(F32(stop - start)) / F64(n - 1) if n != 1 else F64(n)
        return (i.astype(F64) * step + F64(start)).astype(F32)
    step = F32(F32(stop - start) / F32(n - 1)) if n != 1 else F32(n)
    return (i.astype(F32) * step + start).astype(F32)


@lru_cache(maxsize=64)
def _grid_axes(x_min, x_max, y_min, y_max, grid_res, sem):
    nx, ny = _grid_shape(x_min, x_max, y_min, y_max, grid_res, sem)
    if nx < 1 or ny < 1:
        raise ValueError(f"degenerate grid {nx} x {ny} for extents ({x_min}, {x_max}) x ({y_min}, {y_max})")
    return _grid_axis(x_min, x_max, nx, sem), _grid_axis(y_min, y_max, ny, sem)


def morton_slots(pos32, n_padded):
    """Slot of every node in a Z-order (Morton) sort of the positions: spatial neighbours get neighbouring
    slots, so the 32 adjacent pixels one gather instruction serves read neighbouring 16-byte records
    (distinct bank groups) instead of colliding at random.  Pad nodes keep their own index."""
    n = len(pos32)
    lo, hi = pos32.min(axis=0).astype(np.float64), pos32.max(axis=0).astype(np.float64)
    q = ((pos32.astype(np.float64) - lo) / np.maximum(hi - lo, 1e-300) * 65535.0).astype(np.uint64)

    def spread(v):
        v = (v | (v << 8)) & np.uint64(0x00FF00FF)
        v = (v | (v << 4)) & np.uint64(0x0F0F0F0F)
        v = (v | (v << 2)) & np.uint64(0x33333333)
        v = (v | (v << 1)) & np.uint64(0x55555555)
        return v
    code = spread(q[:, 0]) | (spread(q[:, 1]) << np.uint64(1))
    order = np.argsort(code, kind="stable")
    slot = np.arange(n_padded, dtype=np.int32)
    slot[order] = np.arange(n, dtype=np.int32)
    return slot



def prepare_plan(pos, faces, grid_res=238, numpy_semantics=None, allow_degenerate=False):
    """Validation as `matplotlib.tri.Triangulation` does it (mesh_utils.py:103), the grid axes (mesh_utils.py:64-79, 99-100) and
    the Morton slots of the nodes -> dict(pos32 f32 [N,2], tri i32 [F,3], ax, ay f32, slots i32 [roundup(N,4)], n_degenerate,
    numpy_semantics).  Raises ValueError where the reference's Triangulation / trifinder would."""
    pos = np.asarray(pos)
    if pos.ndim != 2 or pos.shape[1] != 2:
        raise ValueError(f"x and y must be equal-length 1D arrays, but found pos of shape {pos.shape!r}")
    try:
        tri = np.array(faces, dtype=np.int32, order="C")   # matplotlib: int32 C-contiguous copy
    except (ValueError, TypeError) as e:
        raise ValueError(f"triangles must be a (N, 3) int array, not {faces!r}") from e
    if tri.ndim != 2 or tri.shape[1] != 3:
        raise ValueError(f"triangles must be a (N, 3) int array, but found shape {tri.shape!r}")
    if tri.shape[0] == 0:
        raise ValueError("triangles must be a (N, 3) int array with N >= 1")
    t_max, t_min = tri.max(), tri.min()
    if t_max >= len(pos):
        raise ValueError("triangles are indices into the points and must be in the range "
                         f"0 <= i < {len(pos)} but found value {t_max}")
    if t_min < 0:
        raise ValueError("triangles are indices into the points and must be in the range "
                         f"0 <= i < {len(pos)} but found value {t_min}")
    pos32 = np.ascontiguousarray(pos, dtype=F32)
    p = pos32.astype(F64)[tri]                              # the same fp64 cross product as correct_triangles
    area2 = (p[:, 1, 0] - p[:, 0, 0]) * (p[:, 2, 1] - p[:, 0, 1]) - (p[:, 1, 1] - p[:, 0, 1]) * (p[:, 2, 0] - p[:, 0, 0])
    n_degenerate = int((area2 == 0).sum())
    if n_degenerate and not allow_degenerate:
        raise ValueError(f"{n_degenerate} triangle(s) of zero area (first: triangle {int(np.argmax(area2 == 0))}): "
                         "the triangulation is invalid for the trapezoid-map trifinder; pass allow_degenerate=True to "
                         "locate with the stated rule anyway")
    sem = numpy_semantics or default_numpy_semantics()
    x_min, y_min = np.min(pos32, axis=0)                    # mesh_utils.py:99-100
    x_max, y_max = np.max(pos32, axis=0)
    ax, ay = _grid_axes(float(x_min), float(x_max), float(y_min), float(y_max), int(grid_res), sem)
    return {"pos32": pos32, "tri": tri, "ax": ax, "ay": ay, "slots": morton_slots(pos32, (len(pos32) + 3) // 4 * 4),
            "n_degenerate": n_degenerate, "numpy_semantics": sem}
