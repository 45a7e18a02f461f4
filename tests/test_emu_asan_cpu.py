"""
CPU memory check of the kernels: the warp-emulator checkers (tests/cpp/*_emu_check.cpp) compiled with -fsanitize=address,undefined.
Every buffer a kernel touches is then a heap block or a static object with red zones, sized exactly as the engine sizes it
(pd_engine.cu: bz_ensure, run_batch), so an index that leaves its array -- in a table build, a match copy, a staging tile, a ring --
is reported instead of landing in a neighbour.  Beside valid inputs the inflate / record-walk kernels get CORRUPTED BAMs (bit flips in
the compressed payloads and in the records): whatever they decode to, the kernels must stay inside their buffers and terminate; the
checker's verdict on such a file (differs / corrupt) is not the point, the sanitizer's silence is.  The device counterpart is
compute-sanitizer memcheck over the GPU suite (profiles/r02_*san*).
"""
import os
import random
import subprocess
from concurrent.futures import ThreadPoolExecutor

import pytest

from tests.helpers import ROOT

BAMGEN = os.path.join(ROOT, "oracle", "_ref", "bamgen")
ENV = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1")


def _build(tmp, name, *libs):
    exe = str(tmp / (name + "_asan"))
    r = subprocess.run(["g++", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-std=c++17", "-pthread",
                        "-I" + os.path.join(ROOT, "pandepth_b200", "csrc"), "-I" + os.path.join(ROOT, "tests", "cpp"),
                        os.path.join(ROOT, "tests", "cpp", name + "_emu_check.cpp"), *libs, "-o", exe], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    if r.returncode != 0:
        if "asan" in r.stderr.lower() or "sanitize" in r.stderr.lower():
            pytest.skip("this toolchain has no AddressSanitizer runtime")
        raise AssertionError(r.stderr[-3000:])
    return exe


def _run(exe, *args, must_pass=True):
    r = subprocess.run([exe, *map(str, args)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900, env=ENV)
    assert "Sanitizer" not in r.stderr and "runtime error" not in r.stderr and r.returncode >= 0, (args, r.stderr[:6000])
    if must_pass:
        lines = [l for l in r.stdout.strip().splitlines() if "checking" not in l]
        assert r.returncode == 0 and lines and all(l.endswith(": ok") for l in lines), (args, r.stdout[-2000:], r.stderr[-2000:])


def test_accumulate_and_reduce_kernels_stay_inside_their_buffers(tmp_path):
    with ThreadPoolExecutor(3) as ex:                 # the three compilations side by side
        tile, bucket, red = ex.map(lambda n: _build(tmp_path, n), ("tile", "bucket", "reduce"))
    _run(tile, 20000, 4000, 1, 1, 1, 0)
    _run(tile, 2047, 3000, 2, 0, 1, 0)              # the last segment is cut by the contig end
    _run(tile, 9000, 2000, 3, 1, 0, 5000)
    _run(bucket, 600000, 8000, 1, 0)
    _run(bucket, 300, 500, 3, 0)
    _run(red, "scatter", 30000, 1500, 1, 0, 1)
    _run(red, "scatter", 8191, 500, 4, 0, 2)
    _run(red, "reduce", 40000, 40, 1, 1, "0xffffffff")
    _run(red, "reduce", 100, 30, 5, 3, "0x3ffff")
    _run(red, "sweep", 3001, 124, 134, "0x3ffff", 5)      # the 128-bit and 5-load forms of the 4-window-group branch
    _run(red, "sweep", 3001, 156, 164, "0x3ffff", 5)      # ... up to its edge
    _run(red, "sweep", 777, 28, 36, "0xffffffff", 6)
    _run(red, "windows", 5000, 513, 3, 1, "0xffffffff", 2)


@pytest.mark.skipif(not os.path.exists(BAMGEN), reason="oracle/_ref/bamgen not built")
def test_inflate_and_record_walk_stay_inside_their_buffers_on_corrupt_files(tmp_path):
    exe = _build(tmp_path, "bgzf", "-lz")
    m, b = str(tmp_path / "M.bam"), str(tmp_path / "B.bam")
    subprocess.run([BAMGEN, "-o", m, "-c", "c1:200000,c2:1,c3:150000,c4:900,c5:30000", "-n", "500", "-s", "3", "-l", "100", "-m", "mixed"], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    subprocess.run([BAMGEN, "-o", b, "-c", "c1:300000", "-n", "80", "-s", "7", "-l", "150", "-q"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    for K in (1, 4):
        _run(exe, m, 8, 1796, -1, 2, 0, K)
        _run(exe, b, 8, 1796, -1, 2, 0, K)
    _run(exe, m, 8, 1796, 300, 2, 0, 1)           # a filter nothing passes: empty dense arrays
    rnd = random.Random(20)
    c = str(tmp_path / "corrupt.bam")
    for k in range(8):
        d = bytearray(open((m, b)[k & 1], "rb").read())
        for _ in range(rnd.choice([1, 1, 2, 5])):
            d[rnd.randrange(200, len(d) - 28)] ^= 1 << rnd.randrange(8)      # behind the BAM header block, in front of the EOF block
        open(c, "wb").write(d)
        _run(exe, c, rnd.choice([1, 2, 3]), 1796, -1, rnd.choice([1, 2]), 0, (1, 2, 4, 8)[k % 4], must_pass=False)
