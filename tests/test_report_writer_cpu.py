"""
CPU: the report writers of the C++ host (pandepth_b200/host/fastfmt.hpp row formatters, the parallel gzip-member writer of
main.cpp) without a GPU.
  * tests/cpp/fastfmt_check.cpp compares the hand-written formatters with printf("%.2f") / "%d" on exhaustive small inputs,
    millions of random ones, exact ties, range limits and whole rows (the reference prints the same doubles through
    ostream fixed/setprecision(2), PanDepth.cpp:4382-4389; glibc and libstdc++ agree);
  * `PANDEPTH_DRY=1 PANDEPTH_SELFTEST_ROWS=<prefix>` makes the CLI write a fixed synthetic window table and site dump through
    its real writer; the files must gunzip to exactly the text Python formats from the same table, for the libdeflate path,
    the zlib path and several thread counts.
"""
import gzip
import os
import subprocess

import pytest

from tests.helpers import GOLD, ROOT

CLI = os.path.join(ROOT, "pandepth_b200", "bin", "pandepth")
HOST = os.path.join(ROOT, "pandepth_b200", "host")


def test_row_formatters_equal_printf(tmp_path):
    exe = str(tmp_path / "fastfmt_check")
    subprocess.run(["g++", "-O2", "-std=c++14", "-I" + HOST, os.path.join(ROOT, "tests", "cpp", "fastfmt_check.cpp"), "-o", exe], check=True)
    r = subprocess.run([exe, "2"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode == 0 and "mismatches 0" in r.stdout, (r.stdout, r.stderr)


def _expected_tables():
    W, CC = 100, 10000003
    n_win = (CC - 1 + W - 1) // W
    rows = ["#Chr\tStart\tEnd\tLength\tCoveredSite\tTotalDepth\tCoverage(%)\tMeanDepth\n"]
    for k in range(n_win):
        ln = min(W, CC - k * W)
        cov = (k * 7919) % (ln + 1)
        tot = cov * ((k * 31) % 500)
        j = 1 + k * W
        end = min(j - 1 + W, CC)
        L = end - j + 1
        rows.append("selftest\t%d\t%d\t%d\t%d\t%d\t%.2f\t%.2f\n" % (j, end, L, cov, tot, cov * 100.0 / L, tot * 1.0 / L))
    site = "".join("selftest\t%d\t%d\n" % (5 + j, ((j * 2654435761) & 0xFFFFFFFF) & 0x3FFFF) for j in range(600000))
    return "".join(rows), site


@pytest.mark.skipif(not os.path.exists(CLI), reason="host binary not built")
def test_writer_files_gunzip_to_the_expected_text(tmp_path):
    win, site = _expected_tables()
    for tag, env_extra, threads in (("ld", {}, "4"), ("ld1", {"PANDEPTH_GZ_LEVEL": "1"}, "1"), ("z", {"PANDEPTH_GZ_ZLIB": "1"}, "7")):
        pre = str(tmp_path / tag)
        env = dict(os.environ, PANDEPTH_DRY="1", PANDEPTH_SELFTEST_ROWS=pre, **env_extra)
        r = subprocess.run([CLI, "-i", os.path.join(GOLD, "k.sam"), "-o", pre + ".o", "-t", threads], env=env,
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
        assert r.returncode == 0, (r.stdout, r.stderr)
        with gzip.open(pre + ".win.gz", "rt") as fh:
            assert fh.read() == win, tag
        with gzip.open(pre + ".site.gz", "rt") as fh:
            assert fh.read() == site, tag
        with gzip.open(pre + ".empty.gz", "rt") as fh:
            assert fh.read() == ""
        # zcat (what a user would run) agrees with Python's gzip on the concatenated members
        z = subprocess.run(["zcat", pre + ".win.gz"], stdout=subprocess.PIPE, check=True).stdout.decode()
        assert z == win
