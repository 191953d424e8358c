"""ORACLE test scaffolding: a matplotlib-shaped stub so the reference's UNMODIFIED Python
(/root/reference/src/dataloader/*.py, src/_triinterpolate.py) can be imported in a container
without matplotlib.  Only the names the reference touches exist; the arithmetic is
oracle/tri_oracle.cpp (a restatement of matplotlib._tri, see its header)."""
from . import _api, tri  # noqa: F401

__version__ = "0+oracle-stub"
