from oracle.mpl_tri import Triangulation, TriFinder, TrapezoidMapTriFinder  # noqa: F401

TriInterpolator = None          # monkey-patched by the reference (mesh_utils.py:14-15)
LinearTriInterpolator = None
