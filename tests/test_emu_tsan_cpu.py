"""
CPU race check of the kernels: the warp-emulator checkers (tests/cpp/*_emu_check.cpp: one host thread per GPU thread, barriers built
on acquire/release atomics for __syncwarp / __syncthreads / the *_sync collectives) compiled with -fsanitize=thread.  Shared-memory
tables, the multi-op ring of k_tile_accumulate, the staging tiles of the bucket passes, the look-back descriptors of k_scan_add and
the literal / match copies of the inflate kernels are then ordinary memory shared by real threads, and ThreadSanitizer reports any
pair of conflicting accesses that no barrier or atomic orders -- the host-side counterpart of compute-sanitizer's racecheck
(profiles/r02_*san*), run here on every kernel without a GPU.  The one intended same-value store from many threads (the
"deltas were written" flag) is an atomic store in the emulated build (pd_kernels.cuh: raise_flag).
"""
import os
import subprocess

import pytest

from tests.helpers import ROOT

BAMGEN = os.path.join(ROOT, "oracle", "_ref", "bamgen")


def _build(tmp, name, *libs):
    exe = str(tmp / (name + "_tsan"))
    r = subprocess.run(["g++", "-O1", "-g", "-fsanitize=thread", "-std=c++17", "-pthread", "-I" + os.path.join(ROOT, "pandepth_b200", "csrc"),
                        "-I" + os.path.join(ROOT, "tests", "cpp"), os.path.join(ROOT, "tests", "cpp", name + "_emu_check.cpp"), *libs, "-o", exe],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    if r.returncode != 0:
        if "tsan" in r.stderr.lower() or "sanitize" in r.stderr.lower():
            pytest.skip("this toolchain has no ThreadSanitizer runtime")
        raise AssertionError(r.stderr[-3000:])
    return exe


def _clean(exe, *args):
    r = subprocess.run([exe, *map(str, args)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=1200,
                       env=dict(os.environ, TSAN_OPTIONS="halt_on_error=0 exitcode=66"))
    if "FATAL: ThreadSanitizer" in r.stderr:
        pytest.skip("ThreadSanitizer cannot map its shadow memory in this container")
    assert "WARNING: ThreadSanitizer" not in r.stderr, r.stderr[:6000]
    lines = [l for l in r.stdout.strip().splitlines() if "checking" not in l]
    assert r.returncode == 0 and lines and all(l.endswith(": ok") for l in lines), (args, r.stdout[-2000:], r.stderr[-2000:])


def test_cut_filter_tile_kernels_race_free(tmp_path):
    exe = _build(tmp_path, "tile")
    _clean(exe, 20000, 4000, 1, 1, 1, 0)        # multi-op ring + filter bitmap
    _clean(exe, 20000, 4000, 2, 0, 0, 0)
    _clean(exe, 9000, 2000, 3, 1, 0, 5000)      # range-sharded


def test_bucket_kernels_race_free(tmp_path):
    _clean(_build(tmp_path, "bucket"), 600000, 8000, 1, 0)


def test_scatter_scan_reduce_window_kernels_race_free(tmp_path):
    exe = _build(tmp_path, "reduce")
    _clean(exe, "scatter", 30000, 1500, 1, 0, 0)
    _clean(exe, "reduce", 40000, 40, 1, 1, "0xffffffff")
    for win in (7, 64, 149, 200, 1000):
        _clean(exe, "windows", 20001, win, 0, 1, "0x3ffff", 3)


@pytest.fixture(scope="module")
def bgzf_case(tmp_path_factory):
    if not os.path.exists(BAMGEN):
        pytest.skip("oracle/_ref/bamgen not built")
    d = tmp_path_factory.mktemp("tsanbam")
    exe = _build(d, "bgzf", "-lz")
    m, b = str(d / "M.bam"), str(d / "B.bam")      # one small block each: the emulated threads run ~50x slower under the sanitizer
    subprocess.run([BAMGEN, "-o", m, "-c", "c1:200000,c2:1,c3:150000,c4:900,c5:30000", "-n", "500", "-s", "3", "-l", "100", "-m", "mixed"], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    subprocess.run([BAMGEN, "-o", b, "-c", "c1:300000", "-n", "80", "-s", "7", "-l", "150", "-q"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return exe, m, b


@pytest.mark.parametrize("K", [1, 2])
def test_bgzf_kernels_race_free(K, bgzf_case):
    exe, m, b = bgzf_case
    _clean(exe, m, 8, 1796, -1, 2, 0, K)        # match-dominated, mixed CIGARs, five contigs
    _clean(exe, b, 8, 1796, -1, 2, 0, K)        # literal-dominated (SEQ + QUAL)
