from oracle.mpl_tri import TriFinder, TrapezoidMapTriFinder  # noqa: F401
