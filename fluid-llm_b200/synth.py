"""Seeded synthetic meshes and trajectories shaped like the reference's datasets.

The reference ships no data; its loaders read per-trajectory pickles with the keys written by
`/root/reference/max/ds_download/torch_MGN.py:68-95` / `MGN_unload.py:84-99`:
`mesh_pos f32[N,2]`, `cells i32[F,3]`, `velocity f32[T,N,2]`, `pressure f32[T,N,1]`
(+ `density`, `node_type`).  The generators below produce exactly those arrays:

  cylinder   [0,1.6]x[0,0.41], circular hole r=0.05 at (0.33,0.2), ~1.9k nodes
  airfoil    graded O-mesh on [-20,20]^2 around an airfoil-shaped hole, ~5.2k nodes
  eagle      [-2.5,2.5]x[-1.7,1.5] with two box obstacles, ~3.4k nodes
  big        structured, jittered, ~0.5M nodes / ~1M triangles on [0,2]x[0,1]

Triangles come from scipy's Delaunay (or a structured split for `big`), holes are cut out, and
half of the triangles get their winding flipped so the orientation fix of the locate step is
exercised.  Rectangle-boundary nodes sit exactly on the bounding box, so the first/last grid
rows and columns of `grid_pos` lie exactly on boundary edges (tie cases).
"""
from __future__ import annotations

import numpy as np

__all__ = ["make_mesh", "make_fields", "make_trajectory", "make_dynamic_trajectory", "MESH_KINDS"]

MESH_KINDS = ("cylinder", "airfoil", "eagle", "big")


def _delaunay(pos64):
    from scipy.spatial import Delaunay
    tri = Delaunay(pos64).simplices.astype(np.int32)
    p = pos64[tri]
    area2 = ((p[:, 1, 0] - p[:, 0, 0]) * (p[:, 2, 1] - p[:, 0, 1])
             - (p[:, 1, 1] - p[:, 0, 1]) * (p[:, 2, 0] - p[:, 0, 0]))
    scale = np.ptp(pos64[:, 0]) * np.ptp(pos64[:, 1])
    return tri[np.abs(area2) > 1e-12 * scale]


def _flip_half(tri, rng):
    flip = rng.random(len(tri)) < 0.5
    out = tri.copy()
    out[flip, 1], out[flip, 2] = tri[flip, 2], tri[flip, 1]
    return out


def _compact(pos, tri):
    used = np.zeros(len(pos), dtype=bool)
    used[tri.ravel()] = True
    remap = np.cumsum(used) - 1
    return pos[used], remap[tri].astype(np.int32)


def _rect_points(rng, x0, x1, y0, y1, nx, ny, jitter):
    xs = np.linspace(x0, x1, nx)
    ys = np.linspace(y0, y1, ny)
    gx, gy = np.meshgrid(xs, ys, indexing="ij")
    dx, dy = (x1 - x0) / (nx - 1), (y1 - y0) / (ny - 1)
    jx = (rng.random(gx.shape) - 0.5) * jitter * dx
    jy = (rng.random(gy.shape) - 0.5) * jitter * dy
    jx[0, :] = jx[-1, :] = 0.0   # keep the rectangle boundary exact
    jy[:, 0] = jy[:, -1] = 0.0
    return np.stack([(gx + jx).ravel(), (gy + jy).ravel()], axis=1)


def _cylinder(rng):
    pts = _rect_points(rng, 0.0, 1.6, 0.0, 0.41, 84, 22, 0.6)
    c, r = np.array([0.33, 0.2]), 0.05
    keep = np.linalg.norm(pts - c, axis=1) > r * 1.25
    th = np.linspace(0, 2 * np.pi, 40, endpoint=False)
    ring = c + r * np.stack([np.cos(th), np.sin(th)], axis=1)
    pos = np.concatenate([pts[keep], ring]).astype(np.float32)
    p64 = pos.astype(np.float64)
    tri = _delaunay(p64)
    cen = p64[tri].mean(axis=1)
    tri = tri[np.linalg.norm(cen - c, axis=1) > r * 0.999]
    return pos, tri


def _naca(n):
    # closed NACA-0012-like outline, chord [0,1]
    beta = np.linspace(0, np.pi, n // 2 + 1)
    xc = 0.5 * (1 - np.cos(beta))
    yt = 0.6 * (0.2969 * np.sqrt(xc) - 0.1260 * xc - 0.3516 * xc ** 2 + 0.2843 * xc ** 3 - 0.1036 * xc ** 4)
    up = np.stack([xc, yt], axis=1)
    lo = np.stack([xc[-2:0:-1], -yt[-2:0:-1]], axis=1)
    return np.concatenate([up, lo])


def _point_in_poly(pts, poly):
    x, y = pts[:, 0], pts[:, 1]
    inside = np.zeros(len(pts), dtype=bool)
    n = len(poly)
    for i in range(n):
        x0, y0 = poly[i]
        x1, y1 = poly[(i + 1) % n]
        cond = (y0 > y) != (y1 > y)
        with np.errstate(divide="ignore", invalid="ignore"):
            xin = (x1 - x0) * (y - y0) / (y1 - y0) + x0
        inside ^= cond & (x < xin)
    return inside


def _airfoil(rng):
    outline = _naca(120)
    # graded cloud: radius log-uniform from the airfoil outwards, plus the far-field box
    n_cloud = 5830
    rad = np.exp(rng.uniform(np.log(0.01), np.log(28.0), n_cloud))
    ang = rng.uniform(0, 2 * np.pi, n_cloud)
    anchor = outline[rng.integers(0, len(outline), n_cloud)]
    cloud = anchor + np.stack([rad * np.cos(ang), rad * np.sin(ang)], axis=1)
    cloud = cloud[(np.abs(cloud[:, 0]) < 19.5) & (np.abs(cloud[:, 1]) < 19.5)]
    cloud = cloud[~_point_in_poly(cloud, outline)]
    d = np.min(np.linalg.norm(cloud[:, None, :] - outline[None, :, :], axis=2), axis=1)
    cloud = cloud[d > 0.004]
    edge = np.linspace(-20, 20, 41)
    box = np.concatenate([np.stack([edge, np.full_like(edge, -20)], 1), np.stack([edge, np.full_like(edge, 20)], 1),
                          np.stack([np.full_like(edge, -20), edge], 1)[1:-1], np.stack([np.full_like(edge, 20), edge], 1)[1:-1]])
    pos = np.concatenate([outline, cloud, box]).astype(np.float32)
    pos = np.unique(pos, axis=0)
    perm = rng.permutation(len(pos))
    pos = pos[perm]
    p64 = pos.astype(np.float64)
    tri = _delaunay(p64)
    cen = p64[tri].mean(axis=1)
    tri = tri[~_point_in_poly(cen, outline.astype(np.float32).astype(np.float64))]
    return pos, tri


def _eagle(rng):
    pts = _rect_points(rng, -2.5, 2.5, -1.7, 1.5, 74, 47, 0.7)
    boxes = [(-1.2, -0.6, -0.5, 0.1), (0.7, 1.1, 0.2, 0.9)]
    keep = np.ones(len(pts), dtype=bool)
    extra = []
    for (bx0, bx1, by0, by1) in boxes:
        m = 0.04
        keep &= ~((pts[:, 0] > bx0 - m) & (pts[:, 0] < bx1 + m) & (pts[:, 1] > by0 - m) & (pts[:, 1] < by1 + m))
        ex = np.linspace(bx0, bx1, 10)
        ey = np.linspace(by0, by1, 10)
        extra += [np.stack([ex, np.full_like(ex, by0)], 1), np.stack([ex, np.full_like(ex, by1)], 1),
                  np.stack([np.full_like(ey, bx0), ey], 1)[1:-1], np.stack([np.full_like(ey, bx1), ey], 1)[1:-1]]
    pos = np.concatenate([pts[keep]] + extra).astype(np.float32)
    p64 = pos.astype(np.float64)
    tri = _delaunay(p64)
    cen = p64[tri].mean(axis=1)
    for (bx0, bx1, by0, by1) in boxes:
        inside = (cen[:, 0] > bx0) & (cen[:, 0] < bx1) & (cen[:, 1] > by0) & (cen[:, 1] < by1)
        tri = tri[~inside]
        cen = cen[~inside]
    return pos, tri


def _big(rng, nx=1001, ny=501):
    pts = _rect_points(rng, 0.0, 2.0, 0.0, 1.0, nx, ny, 0.5).astype(np.float32)
    idx = np.arange(nx * ny, dtype=np.int32).reshape(nx, ny)
    a, b, c, d = idx[:-1, :-1].ravel(), idx[1:, :-1].ravel(), idx[1:, 1:].ravel(), idx[:-1, 1:].ravel()
    diag = rng.random(len(a)) < 0.5
    t1 = np.where(diag[:, None], np.stack([a, b, c], 1), np.stack([a, b, d], 1))
    t2 = np.where(diag[:, None], np.stack([a, c, d], 1), np.stack([b, c, d], 1))
    tri = np.concatenate([t1, t2]).astype(np.int32)
    tri = tri[rng.permutation(len(tri))]
    return pts, tri


def make_mesh(kind: str, seed: int = 0, **kw):
    """-> (mesh_pos f32[N,2], cells i32[F,3]); deterministic for a given (kind, seed)."""
    rng = np.random.default_rng(seed)
    if kind == "cylinder":
        pos, tri = _cylinder(rng)
    elif kind == "airfoil":
        pos, tri = _airfoil(rng)
    elif kind == "eagle":
        pos, tri = _eagle(rng)
    elif kind == "big":
        pos, tri = _big(rng, **kw)
    else:
        raise ValueError(f"unknown mesh kind {kind!r}; expected one of {MESH_KINDS}")
    pos, tri = _compact(pos, tri)
    tri = _flip_half(tri, rng)
    return np.ascontiguousarray(pos, dtype=np.float32), np.ascontiguousarray(tri, dtype=np.int32)


_FIELD_STATS = {  # (mean, std) of u, v, p -- SURVEY.md section 6
    "cylinder": ((0.82, 0.33), (0.0, 0.12), (0.05, 0.27)),
    "airfoil": ((170.1, 71.06), (-1.183, 46.73), (9.935e4, 8964.0)),
    "eagle": ((0.2, 1.6), (0.0, 1.9), (-0.5, 6.4)),
    "big": ((0.82, 0.33), (0.0, 0.12), (0.05, 0.27)),
}


def make_fields(kind: str, pos: np.ndarray, T: int, seed: int = 1):
    """Smooth-in-space-and-time fields -> velocity f32[T,N,2], pressure f32[T,N,1]."""
    rng = np.random.default_rng(seed)
    N = len(pos)
    x = pos[:, 0].astype(np.float64)
    y = pos[:, 1].astype(np.float64)
    lx, ly = max(np.ptp(x), 1e-9), max(np.ptp(y), 1e-9)
    xs, ys = (x - x.min()) / lx, (y - y.min()) / ly
    t = np.arange(T, dtype=np.float64)[:, None] * 0.05
    out = []
    for (mean, std) in _FIELD_STATS[kind]:
        f = np.zeros((T, N))
        for _ in range(4):
            kx, ky, w, ph = rng.uniform(1, 9), rng.uniform(1, 9), rng.uniform(0.5, 3), rng.uniform(0, 6.28)
            f += np.sin(kx * xs[None, :] + ky * ys[None, :] * 0.7 + w * t + ph)
        f = f / 2.0 + 0.05 * rng.standard_normal((1, N))
        out.append((mean + std * f).astype(np.float32))
    velocity = np.ascontiguousarray(np.stack([out[0], out[1]], axis=2))
    pressure = np.ascontiguousarray(out[2][:, :, None])
    return velocity, pressure


def make_trajectory(kind: str, T: int, mesh_seed: int = 0, field_seed: int = 1, **kw):
    """One pickle-shaped dict, the layout `simple_dataloader.py:154-164` unpickles."""
    pos, tri = make_mesh(kind, mesh_seed, **kw)
    vel, prs = make_fields(kind, pos, T, field_seed)
    return {"mesh_pos": pos, "cells": tri, "velocity": vel, "pressure": prs,
            "density": np.ones((T, len(pos), 1), dtype=np.float32),
            "node_type": np.zeros((len(pos), 1), dtype=np.int32)}


def _flip_edges(pos64, tri, rng, frac):
    """Flip the shared diagonal of a random set of adjacent triangle pairs whose union is a strictly convex quad:
    same nodes, same triangle count, different connectivity."""
    tri = tri.copy()
    edge_owner = {}
    pairs = []
    for t, (a, b, c) in enumerate(tri):
        for (u, v, w) in ((a, b, c), (b, c, a), (c, a, b)):
            key = (min(u, v), max(u, v))
            if key in edge_owner:
                pairs.append((edge_owner[key], (t, w), key))
            else:
                edge_owner[key] = (t, w)
    if not pairs:
        return tri
    order = rng.permutation(len(pairs))[: max(1, int(frac * len(pairs)))]
    touched = set()

    def area2(p, q, r):
        return (q[0] - p[0]) * (r[1] - p[1]) - (q[1] - p[1]) * (r[0] - p[0])

    for i in order:
        (t1, w1), (t2, w2), (u, v) = pairs[i]
        if t1 in touched or t2 in touched:
            continue
        pu, pv, p1, p2 = pos64[u], pos64[v], pos64[w1], pos64[w2]
        # the new diagonal w1-w2 must separate u and v strictly, and both new triangles must keep some area
        s1, s2 = area2(p1, p2, pu), area2(p1, p2, pv)
        scale = abs(area2(pu, pv, p1)) + abs(area2(pu, pv, p2))
        if s1 * s2 >= 0 or min(abs(s1), abs(s2)) < 0.05 * scale:
            continue
        tri[t1] = (w1, w2, u)
        tri[t2] = (w2, w1, v)
        touched.update((t1, t2))
    return tri


def make_dynamic_trajectory(kind: str, T: int, mesh_seed: int = 0, field_seed: int = 1, flip_frac: float = 0.05, **kw):
    """A trajectory whose mesh changes every frame, in the layout of EAGLE's sim.npz + triangles.npy
    (max/ds_download/eagle.py:123-144): mesh_pos f32[T,N,2], cells i32[T,F,3], velocity f32[T,N,2], pressure f32[T,N,1].
    Interior nodes drift smoothly (the bounding box stays put), a few diagonals are flipped per frame, and every frame
    lists its triangles in a different order with random winding."""
    pos, tri = make_mesh(kind, mesh_seed, **kw)
    rng = np.random.default_rng(7919 * mesh_seed + 17)
    p64 = pos.astype(np.float64)
    lo, hi = p64.min(axis=0), p64.max(axis=0)
    ext = hi - lo
    u = (p64 - lo) / ext
    bump = np.sin(np.pi * u[:, 0]) * np.sin(np.pi * u[:, 1])          # zero on the bounding box
    amp = 0.2 * np.sqrt(ext[0] * ext[1] / len(pos))                      # a fraction of the mean node spacing
    ph = rng.uniform(0, 2 * np.pi, size=4)
    mesh_pos = np.empty((T, len(pos), 2), dtype=np.float32)
    cells = np.empty((T, len(tri), 3), dtype=np.int32)
    for t in range(T):
        a = 0.31 * t
        d = np.stack([np.sin(2 * np.pi * u[:, 1] + a + ph[0]) * np.cos(a + ph[1]),
                      np.cos(2 * np.pi * u[:, 0] - a + ph[2]) * np.sin(a + ph[3])], axis=1)
        pt = (p64 + amp * bump[:, None] * d).astype(np.float32)
        mesh_pos[t] = pt
        tt = _flip_edges(pt.astype(np.float64), tri, rng, flip_frac) if flip_frac > 0 else tri
        tt = _flip_half(tt[rng.permutation(len(tt))], rng)
        cells[t] = tt
    vel, prs = make_fields(kind, pos, T, field_seed)
    return {"mesh_pos": mesh_pos, "cells": cells, "velocity": vel, "pressure": prs}
