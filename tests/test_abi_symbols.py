"""CPU: the C-ABI library builds, loads, and exports every function include/pandepth_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "pandepth_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pd_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_the_boundary():
    names = _declared()
    for must in ("pd_create", "pd_create_ranges", "pd_push", "pd_push_contig", "pd_push_bgzf", "pd_finalize", "pd_reduce", "pd_reduce_windows",
                 "pd_depth_read", "pd_destroy", "pd_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    import pandepth_b200 as pd
    lib = pd.load_library()
    for n in _declared():
        assert hasattr(lib, n), f"libpandepth_b200.so does not export {n}"


def test_host_library_exports_every_declared_symbol():
    """include/pandepth_host.h (the feeder's C entry points) against libpandepth_host.so."""
    src = open(os.path.join(ROOT, "include", "pandepth_host.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = sorted(set(re.findall(r"\b(pdh_[a-z_0-9]+)\s*\(", src)))
    assert "pdh_feed_bam" in names and "pdh_feed_bam_range" in names and "pdh_bai_block_offset" in names
    path = os.path.join(ROOT, "pandepth_b200", "libpandepth_host.so")
    if not os.path.exists(path):
        pytest.skip("host library not built")
    lib = ctypes.CDLL(path)
    for n in names:
        assert hasattr(lib, n), f"libpandepth_host.so does not export {n}"


def test_no_cpu_fallback_without_gpu():
    """Without a GPU pd_create must fail loudly (no silent CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import pandepth_b200 as pd
    with pytest.raises(pd.PDError):
        pd.Engine([1000])


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under pandepth_b200/ may reference it."""
    bad = []
    for dp, _, fns in os.walk(os.path.join(ROOT, "pandepth_b200")):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")) or fn == "Makefile":
                txt = open(os.path.join(dp, fn), errors="ignore").read()
                if re.search(r"\boracle\b|pd_oracle|pdo_", txt):
                    bad.append(os.path.join(dp, fn))
    assert not bad, bad
