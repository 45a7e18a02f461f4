"""
CPU: the kernels of the bucketed unsorted path (pandepth_b200/csrc/pd_bucket.cuh: k_expand_runs, k_bucket_hist / scan / scatter,
k_bucket_fine, k_bucket_cuts) run through the warp emulator (tests/cpp/warp_emu.h: one host thread per GPU thread, barriers for
__syncthreads, a per-CTA buffer for dynamic shared memory) on random unsorted batches and are compared with a plain host walk:
the bucketed runs are exactly the expected runs, in 256-base bucket order, the fine offsets point at every bucket's first run and
the segment cuts equal the lower bounds k_cut computes for sorted input.  The same path is checked end to end on the device against
the oracle's depth in tests/test_engine_gpu.py::test_bucketed_unsorted_path_matches_oracle (-m gpu).
"""
import os
import subprocess

import pytest

from tests.helpers import ROOT


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("bemu") / "bucket_emu_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-I" + os.path.join(ROOT, "pandepth_b200", "csrc"), "-I" + os.path.join(ROOT, "tests", "cpp"),
                    os.path.join(ROOT, "tests", "cpp", "bucket_emu_check.cpp"), "-o", exe], check=True)
    return exe


@pytest.mark.parametrize("clen,reads,seed,shift", [
    (1_300_000, 40_000, 1, 0),          # three coarse buckets, several staging tiles
    (1_300_000, 40_000, 2, 700_000),    # an owned range that starts inside the contig (range sharding): runs are clipped to it
    (300, 500, 3, 0),                   # a contig smaller than one fine bucket
    (524_288, 90_000, 4, 12_345),       # exactly one coarse bucket, more slots than one histogram tile
    (2_100_000, 150_000, 5, 0),         # five coarse buckets
    (5_000, 20_000, 6, 100),            # a deep pile: thousands of runs per fine bucket
])
def test_bucket_kernels_match_a_host_walk(clen, reads, seed, shift, checker):
    r = subprocess.run([checker, str(clen), str(reads), str(seed), str(shift)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith(": ok"), (r.stdout[-2000:], r.stderr[-2000:])
