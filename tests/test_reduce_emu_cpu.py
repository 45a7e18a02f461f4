"""
CPU: the finalize and reduce kernels (pandepth_b200/csrc/pd_kernels.cuh: k_delta_scatter, k_scan_add, k_fill_untouched,
k_reduce_units_part, k_reduce_units, k_reduce_windows) run unchanged through the warp emulator (tests/cpp/warp_emu.h) and are compared
with the reference's plain loops restated in tests/cpp/reduce_emu_check.cpp (PanDepth.cpp:732-760, :295-348, :4366-4381):

  scatter   unsorted multi-op batches on two contigs -> +-1 deltas (one launch / one launch per stripe / the device-side read count of
            range sharding) -> the decoupled look-back scan, over segments that already hold depth and segments never written
  reduce    random intervals (whole contigs, chunk-aligned, single bases, over both ends, empty) over depth whose segments are a mix of
            never written / written once with partials / written more than once, with and without the reduce-hint partials
  windows   every width class of k_reduce_windows (lane per window < 32; the 4-window groups of 32..160 with their 128-bit and
            5-load forms; 161..512; wider), 32-bit and 18-bit masks, -d 1 and -d 3, aligned and unaligned first windows

With tests/test_bgzf_emu_cpu.py, test_bucket_emu_cpu.py and test_tile_emu_cpu.py every kernel of the library has a CPU run of its source.
"""
import os
import subprocess

import pytest

from tests.helpers import ROOT


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("remu") / "reduce_emu_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-I" + os.path.join(ROOT, "pandepth_b200", "csrc"), "-I" + os.path.join(ROOT, "tests", "cpp"),
                    os.path.join(ROOT, "tests", "cpp", "reduce_emu_check.cpp"), "-o", exe], check=True)
    return exe


def _ok(checker, *args):
    r = subprocess.run([checker, *map(str, args)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    lines = r.stdout.strip().splitlines()
    assert r.returncode == 0 and lines and all(l.endswith(": ok") for l in lines), (args, r.stdout[-2000:], r.stderr[-2000:])


@pytest.mark.parametrize("clen,reads,seed,shift,style", [
    (50_000, 3_000, 1, 0, 0),
    (50_000, 3_000, 2, 7_000, 1),        # striped scatter of a range-sharded contig
    (50_000, 3_000, 3, 14_000, 2),       # the device-side count (reads left of a sorted batch's first cut)
    (8_191, 500, 4, 0, 0),               # contig + padding element fill exactly one pad unit
    (140_000, 200, 5, 0, 1),             # sparse: look-back over many all-zero tiles
    (300, 4_000, 6, 0, 0),               # everything piles into one tile
])
def test_delta_scatter_and_scan(clen, reads, seed, shift, style, checker):
    _ok(checker, "scatter", clen, reads, seed, shift, style)


@pytest.mark.parametrize("clen,n_iv,seed,min_dep,mask", [
    (70_000, 120, 1, 1, "0xffffffff"),
    (70_000, 120, 2, 3, "0x3ffff"),      # the 18-bit SiteInfo view (DataClass.h:84-87) with -d 3
    (8_192, 60, 3, 1, "0x3ffff"),
    (200_000, 40, 4, 2, "0xffffffff"),
    (100, 30, 5, 1, "0xffffffff"),
])
def test_interval_reduce_with_and_without_partials(clen, n_iv, seed, min_dep, mask, checker):
    _ok(checker, "reduce", clen, n_iv, seed, min_dep, mask)


WIDTHS = (1, 7, 31, 32, 33, 64, 100, 128, 129, 149, 150, 160, 161, 200, 512, 513, 1000, 20000)


@pytest.mark.parametrize("mask", ["0xffffffff", "0x3ffff"])
@pytest.mark.parametrize("min_dep", [1, 3])
@pytest.mark.parametrize("start0", [0, 37, 64])
def test_fixed_windows_every_width_class(mask, min_dep, start0, checker):
    for k, win in enumerate(WIDTHS):
        _ok(checker, "windows", 50_001 if k % 2 else 49_152, win, start0, min_dep, mask, 7 + k)


@pytest.mark.parametrize("first,last,mask", [(32, 170, "0x3ffff"), (1, 40, "0xffffffff"), (500, 520, "0x3ffff")])
def test_fixed_windows_every_width_of_a_range(first, last, mask, checker):
    """every width of the 4-window-group branch (multiples of 4 on aligned starts take its 128-bit form, the others the 5-load form; -d 3 on
    widths = 2 mod 4), the lane-per-window branch and the edge of the 32-bit lane-strided branch"""
    _ok(checker, "sweep", 12_001, first, last, mask, 5)
