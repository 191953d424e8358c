"""ORACLE -- test infrastructure only (see oracle/tri_oracle.cpp for the header and the
"parity unpinned" statement).

matplotlib-shaped Python front of the C++ restatement: the same three entry points the
reference touches (`/root/reference/src/dataloader/mesh_utils.py:103-104`,
`/root/reference/src/_triinterpolate.py:262-263`):

    triang = Triangulation(x, y, triangles)        # matplotlib.tri.Triangulation
    tri_index = triang.get_trifinder()(gx, gy)     # TrapezoidMapTriFinder.__call__ -> find_many
    planes = triang.calculate_plane_coefficients(z)

Restated from matplotlib 3.8.2 `lib/matplotlib/tri/_triangulation.py` / `_trifinder.py`:
x, y become float64; triangles become a C-contiguous int32 copy; shapes and index ranges are
validated with ValueError; the C++ triangulation (orientation fix, neighbours) is created on
construction here (matplotlib: lazily, same result).
"""
import ctypes

import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = _build.build()
        L = ctypes.CDLL(path)
        c_dp = ctypes.POINTER(ctypes.c_double)
        c_ip = ctypes.POINTER(ctypes.c_int)
        L.fo_tri_create.restype = ctypes.c_void_p
        L.fo_tri_create.argtypes = [c_dp, c_dp, ctypes.c_int, c_ip, ctypes.c_int]
        L.fo_tri_destroy.argtypes = [ctypes.c_void_p]
        L.fo_tri_get_triangles.argtypes = [ctypes.c_void_p, c_ip]
        L.fo_tri_get_neighbors.argtypes = [ctypes.c_void_p, c_ip]
        L.fo_tri_plane_coefficients.argtypes = [ctypes.c_void_p, c_dp, c_dp]
        L.fo_trifinder_init.restype = ctypes.c_int
        L.fo_trifinder_init.argtypes = [ctypes.c_void_p]
        L.fo_trifinder_find_many.argtypes = [ctypes.c_void_p, c_dp, c_dp, ctypes.c_long, c_ip]
        L.fo_trifinder_stats.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_long), ctypes.POINTER(ctypes.c_long)]
        L.fo_rule_find_many.argtypes = [ctypes.c_void_p, c_dp, c_dp, ctypes.c_long, ctypes.c_int, c_ip]
        L.fo_to_grid.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float),
                                 ctypes.POINTER(ctypes.c_float), c_ip, ctypes.c_long,
                                 ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_ubyte), c_dp, c_dp]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


class Triangulation:
    """matplotlib.tri.Triangulation with explicit triangles (is_delaunay = False)."""

    def __init__(self, x, y, triangles=None, mask=None):
        self.x = np.asarray(x, dtype=np.float64)
        self.y = np.asarray(y, dtype=np.float64)
        if self.x.shape != self.y.shape or self.x.ndim != 1:
            raise ValueError("x and y must be equal-length 1D arrays, but found shapes "
                             f"{self.x.shape!r} and {self.y.shape!r}")
        if triangles is None:
            raise ValueError("the oracle restates only the explicit-triangles path used by the reference")
        if mask is not None:
            raise ValueError("masked triangulations are not used by the reference path")
        try:
            self.triangles = np.array(triangles, dtype=np.int32, order='C')
        except ValueError as e:
            raise ValueError('triangles must be a (N, 3) int array, not '
                             f'{triangles!r}') from e
        if self.triangles.ndim != 2 or self.triangles.shape[1] != 3:
            raise ValueError('triangles must be a (N, 3) int array, but found shape '
                             f'{self.triangles.shape!r}')
        if self.triangles.size and self.triangles.max() >= len(self.x):
            raise ValueError('triangles are indices into the points and must be in the range 0 <= i < '
                             f'{len(self.x)} but found value {self.triangles.max()}')
        if self.triangles.size and self.triangles.min() < 0:
            raise ValueError('triangles are indices into the points and must be in the range 0 <= i < '
                             f'{len(self.x)} but found value {self.triangles.min()}')
        self.mask = None
        self.is_delaunay = False
        self._x = np.ascontiguousarray(self.x)
        self._y = np.ascontiguousarray(self.y)
        self._h = lib().fo_tri_create(_dp(self._x), _dp(self._y), len(self._x), _ip(self.triangles),
                                      len(self.triangles))
        if not self._h:
            raise ValueError("bad triangle indices")
        self._trifinder = None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _lib is not None:
            _lib.fo_tri_destroy(h)
            self._h = None

    # -- C++ side views ------------------------------------------------------------------
    @property
    def corrected_triangles(self):
        out = np.empty_like(self.triangles)
        lib().fo_tri_get_triangles(self._h, _ip(out))
        return out

    @property
    def neighbors(self):
        out = np.empty_like(self.triangles)
        lib().fo_tri_get_neighbors(self._h, _ip(out))
        return out

    def calculate_plane_coefficients(self, z):
        z = np.ascontiguousarray(z, dtype=np.float64)
        if z.shape != self.x.shape:
            raise ValueError("z array must have same length as triangulation x and y arrays")
        out = np.empty((len(self.triangles), 3), dtype=np.float64)
        lib().fo_tri_plane_coefficients(self._h, _dp(z), _dp(out))
        return out

    def get_trifinder(self):
        if self._trifinder is None:
            self._trifinder = TrapezoidMapTriFinder(self)
        return self._trifinder

    def get_masked_triangles(self):
        return self.triangles


class TriFinder:
    def __init__(self, triangulation):
        self._triangulation = triangulation


class TrapezoidMapTriFinder(TriFinder):
    def __init__(self, triangulation):
        super().__init__(triangulation)
        if lib().fo_trifinder_init(triangulation._h) != 0:
            raise RuntimeError("Triangulation is invalid")

    def __call__(self, x, y):
        x = np.asarray(x, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        if x.shape != y.shape:
            raise ValueError("x and y must be array-like with the same shape")
        xr = np.ascontiguousarray(x.ravel())
        yr = np.ascontiguousarray(y.ravel())
        out = np.empty(xr.shape, dtype=np.int32)
        lib().fo_trifinder_find_many(self._triangulation._h, _dp(xr), _dp(yr), xr.size, _ip(out))
        return out.reshape(x.shape)

    def tree_stats(self):
        a, b = ctypes.c_long(0), ctypes.c_long(0)
        lib().fo_trifinder_stats(self._triangulation._h, ctypes.byref(a), ctypes.byref(b))
        return a.value, b.value


def rule_find_many(triangulation, x, y, bucketed=True):
    """The stated tie-break rule evaluated directly (brute force / bucket grid)."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    xr = np.ascontiguousarray(x.ravel())
    yr = np.ascontiguousarray(y.ravel())
    out = np.empty(xr.shape, dtype=np.int32)
    lib().fo_rule_find_many(triangulation._h, _dp(xr), _dp(yr), xr.size, int(bool(bucketed)), _ip(out))
    return out.reshape(x.shape)


def c_to_grid(triangulation, val, grid_x, grid_y, tri_index):
    """C restatement of mesh_utils.to_grid for one scalar field (CPU-baseline timing helper)."""
    val = np.ascontiguousarray(val, dtype=np.float32)
    gx = np.ascontiguousarray(grid_x, dtype=np.float32)
    gy = np.ascontiguousarray(grid_y, dtype=np.float32)
    ti = np.ascontiguousarray(tri_index, dtype=np.int32)
    data = np.empty(gx.shape, dtype=np.float32)
    mask = np.empty(gx.shape, dtype=np.uint8)
    plane = np.empty((len(triangulation.triangles), 3), dtype=np.float64)
    zs = np.empty(len(triangulation.x), dtype=np.float64)
    lib().fo_to_grid(triangulation._h, _fp(val), _fp(gx), _fp(gy), _ip(ti), gx.size, _fp(data),
                     mask.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)), _dp(plane), _dp(zs))
    return data, mask.astype(bool)
