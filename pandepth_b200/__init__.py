"""
pandepth_b200 -- B200-native per-base depth engine, drop-in for PanDepth's coverage hot path.

The product is the C-ABI library `libpandepth_b200.so` (pandepth_b200/csrc, declared in include/pandepth_b200.h)
and the C++ host around it (pandepth_b200/host).  This Python package is only a thin ctypes binding used by the
tests and bench.py; it has no CPU fallback: importing `Engine` without the built CUDA library raises.
"""
from .engine import Engine, PDError, lib_path, load_library, PD_SORTED, PD_DEVICE, PD_ASYNC, PD_READY, PD_CRC  # noqa: F401
