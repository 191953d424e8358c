"""ORACLE -- test infrastructure only (nothing under fluid-llm_b200/ imports this).

A third, arithmetic-independent witness for the one link of the parity chain that cannot be pinned to a real
matplotlib (matplotlib is not installed here; see oracle/tri_oracle.cpp): the STATED tie-break rule
(include/fluidgrid.h) evaluated in EXACT rational arithmetic over the float32 coordinates.

    (i)   a query equal to a mesh vertex -> the lowest-index triangle listing it
    (ii)  a query on an edge (orientation exactly 0) -> the triangle ABOVE the edge (left of the edge directed from its
          lexicographically smaller to larger end point), else, on a boundary edge, the triangle below
    (iii) otherwise the triangle whose three half-plane tests pass
    (iv)  -1 if none

The fp64 locators (the trapezoid map restated from matplotlib `_tri.cpp`, the brute-force rule in tri_oracle.cpp and the
CUDA kernels) evaluate the orientation expression (p - l) x (r - l) with separately rounded fp64 products, as matplotlib
does on x86-64; wherever that sign equals the exact sign -- every query of the test meshes -- all four must agree.

Pure Python (`fractions.Fraction`), a few thousand queries per second: used on the hand-built tie mesh and on the
boundary rows and columns of the data-set shaped meshes, where the grid points sit exactly on mesh edges.
"""
from fractions import Fraction

import numpy as np


def _orient_sign(px, py, lx, ly, rx, ry):
    s = (px - lx) * (ry - ly) - (py - ly) * (rx - lx)
    return (s > 0) - (s < 0)


def _rule(q, v):
    """-1 rejected, 0 accepted with full priority, 1 accepted only if nothing lies above the edge the query is on."""
    if any(q == p for p in v):
        return 0
    prio = 0
    for k in range(3):
        a, b = v[k], v[(k + 1) % 3]
        end_right = (b[1] > a[1]) if b[0] == a[0] else (b[0] > a[0])
        s = _orient_sign(q[0], q[1], a[0], a[1], b[0], b[1]) if end_right else _orient_sign(q[0], q[1], b[0], b[1], a[0], a[1])
        if end_right:
            if s > 0:
                return -1
        else:
            if s < 0:
                return -1
            if s == 0:
                prio = 1
    return prio


def exact_find_many(pos, triangles, qx, qy):
    """pos float32 (N, 2), triangles int (F, 3) (any winding), qx / qy float32 arrays -> int32 triangle ids, same shape.

    The winding fix of matplotlib's correct_triangles ((p1 - p0) x (p2 - p0) < 0 -> swap vertices 1 and 2) is applied in
    exact arithmetic too."""
    pos = np.asarray(pos, dtype=np.float32)
    tri = np.asarray(triangles, dtype=np.int64)
    P = [(Fraction(float(x)), Fraction(float(y))) for x, y in pos]
    ct = []
    for a, b, c in tri:
        if _orient_sign(P[b][0], P[b][1], P[a][0], P[a][1], P[c][0], P[c][1]) < 0:     # (b - a) x (c - a) < 0: clockwise
            b, c = c, b
        ct.append((a, b, c))
    p64 = pos.astype(np.float64)
    lo = p64[tri].min(axis=1)
    hi = p64[tri].max(axis=1)
    shape = np.shape(qx)
    qx = np.asarray(qx, dtype=np.float32).ravel()
    qy = np.asarray(qy, dtype=np.float32).ravel()
    out = np.full(len(qx), -1, dtype=np.int32)
    for i, (x, y) in enumerate(zip(qx.astype(np.float64), qy.astype(np.float64))):
        cand = np.nonzero((lo[:, 0] <= x) & (x <= hi[:, 0]) & (lo[:, 1] <= y) & (y <= hi[:, 1]))[0]
        q = (Fraction(float(x)), Fraction(float(y)))
        best = None
        for t in cand:
            a, b, c = ct[t]
            p = _rule(q, (P[a], P[b], P[c]))
            if p >= 0 and (best is None or (p, t) < best):
                best = (p, int(t))
        if best is not None:
            out[i] = best[1]
    return out.reshape(shape)
