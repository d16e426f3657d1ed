"""hifiasm_b200 — B200-native all-vs-all overlap engine behind hifiasm's
overlap/error-correction stage boundary (see DESIGN.md / INTEGRATION.md).

The compute lives in csrc/ (hand-written sm_100a CUDA behind the C-ABI of
include/hifiasm_b200.h, built in-tree as libhifiasm_b200.so); this package is
the thin host-side mirror used by tests and bench.py.  There is no CPU path:
importing works anywhere, creating an Engine needs a CUDA device.
"""
from .engine import Engine, HBError, lib_path  # noqa: F401
