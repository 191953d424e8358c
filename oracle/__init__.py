"""CPU oracle of the field data path -- TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs may import this package;
nothing under `fluid-llm_b200/` does (tests/test_abi.py::test_product_never_imports_the_oracle).

    tri_oracle.cpp   C++ restatement of matplotlib 3.8.2 `_tri` (trapezoid map, plane coefficients) -- parity unpinned,
                     see the file header -- plus an independent rule-based locator
    mpl_tri.py       ctypes front with matplotlib's class names
    pipeline.py      NumPy restatement of the reference's Python path, pinned against the reference run in-container
    ref_import.py    imports the UNMODIFIED reference over stub modules (only where /root/reference exists)
    make_golden.py   writes tests/golden/*.npz from that import
"""
