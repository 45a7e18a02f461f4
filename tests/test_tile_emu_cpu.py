"""
CPU: the accumulate kernels of the sorted path (pandepth_b200/csrc/pd_kernels.cuh: k_cut_coarse, k_cut, k_filter_bits and the four
instantiations of k_tile_accumulate) run through the warp emulator (tests/cpp/warp_emu.h) on random coordinate-sorted batches --
single-op and mixed CIGARs (S/H clips, I, D, N skips of up to 6 kb, =/X, P, "*"), reads up to 3 kb (far runs through the delta
array), pile-ups, flags and mapq, reads over the contig end, an owned range that starts inside the contig -- and are compared base
by base with a plain host walk of PanDepth.cpp:732-760.  Each batch is accumulated twice (first write, then the add into segments
that already hold depth); the per-segment CoveredSite / TotalDepth partials of the reduce hint are checked against what was written,
and an out-of-order batch must be reported.  The device run of the same kernels against the oracle is tests/test_engine_gpu.py.
"""
import os
import subprocess

import pytest

from tests.helpers import ROOT


def _build(tmp, name, *defs):
    exe = str(tmp / name)
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", *defs, "-I" + os.path.join(ROOT, "pandepth_b200", "csrc"), "-I" + os.path.join(ROOT, "tests", "cpp"),
                    os.path.join(ROOT, "tests", "cpp", "tile_emu_check.cpp"), "-o", exe], check=True)
    return exe


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    return _build(tmp_path_factory.mktemp("temu"), "tile_emu_check")


@pytest.mark.parametrize("clen,reads,seed,has_off,has_filter,shift", [
    (60_000, 20_000, 1, 0, 0, 0),           # k_tile_accumulate<false,false>: the kernel bench.py's `value` times
    (60_000, 20_000, 1, 0, 1, 0),           # <false,true>: the filter bitmap
    (60_000, 20_000, 1, 1, 0, 0),           # <true,false>: the queue of multi-op reads
    (60_000, 20_000, 1, 1, 1, 0),           # <true,true>
    (60_000, 20_000, 2, 1, 1, 1_000_000),   # range sharding: segment 0 owns the reads of its halo
    (10_239, 3_000, 3, 0, 0, 4_096),        # contig end one base short of a segment boundary
    (400_000, 300, 5, 1, 1, 0),             # sparse: most segments see no read (never written, no partials)
    (4_000, 30_000, 6, 1, 0, 0),            # 15,000 reads per segment: the ring wraps many times
    (2_047, 50, 7, 0, 1, 0),                # one partial segment
    (2_048, 0, 7, 0, 1, 0),                 # an empty batch
    (257, 31_289, 8, 0, 1, 255),            # a range that starts 255 bases into the contig: the halo of segment 0 is cut at the contig start
    (1, 512, 9, 1, 0, 0),                   # a one-base contig
])
def test_tile_kernels_match_a_host_walk(clen, reads, seed, has_off, has_filter, shift, checker):
    r = subprocess.run([checker, str(clen), str(reads), str(seed), str(has_off), str(has_filter), str(shift)], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith(": ok"), (r.stdout[-2000:], r.stderr[-2000:])


@pytest.mark.parametrize("defs", [("-DPD_TILE_UNROLL=4",), ("-DPD_TILE_BULK=0",), ("-DPD_TILE_AGG=1",), ("-DPD_PREFETCH=0", "-DPD_TILE_BULK=1")])
def test_compile_time_variants(defs, tmp_path):
    """the build knobs DESIGN.md's measurements switch (unroll depth, flush variant, duplicate aggregation) compute the same depth"""
    exe = _build(tmp_path, "tile_emu_variant", *defs)
    for has_off in (0, 1):
        r = subprocess.run([exe, "30000", "12000", "9", str(has_off), "1", "0"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
        assert r.returncode == 0 and r.stdout.strip().endswith(": ok"), (defs, r.stdout[-2000:], r.stderr[-2000:])
