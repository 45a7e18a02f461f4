"""
The seam, proven inside the reference program: oracle/_ref/pandepth_ref_seam is /root/reference/src/PanDepth.cpp with the hot
loop of ProDealChrBambai (`depth[p]++`, PanDepth.cpp:730-761) replaced by pd_create / pd_push_contig / pd_depth_read exactly
as INTEGRATION.md section 2 describes (built by oracle/make_seam.py in the build container; the binary travels with the
snapshot).  Everything else in that binary is the reference's own code, so its outputs must equal the golden outputs the
stock reference binary printed.  Indexed BAM inputs in the modes that go through ProDealChrBambai (per-chr, -w >= 150, -g, -b).
"""
import gzip
import json
import os
import subprocess

import pytest

from tests.helpers import GOLD, ROOT, assert_matches

SEAM = os.path.join(ROOT, "oracle", "_ref", "pandepth_ref_seam")
BAMGEN = os.path.join(ROOT, "oracle", "_ref", "bamgen")

with open(os.path.join(GOLD, "expected.json")) as _fh:
    _CASES = json.load(_fh)

# ProDealChrBambai serves the indexed default path: not -a, not -w < 150 (mode 6), not -s, not lists / PAF
SEAM_CASES = sorted(n for n, c in _CASES.items()
                    if c["kind"] == "bam" and not c["fixture"].startswith("list:") and "-a" not in c["args"] and "-s" not in c["args"]
                    and not ("-w" in c["args"] and int(c["args"][c["args"].index("-w") + 1]) < 150))


def _run_seam(case, tmp_path, env=None):
    c = _CASES[case]
    args = [os.path.join(GOLD, a[1:]) if a.startswith("@") else a for a in c["args"]]
    if c["fixture"] == "@stack":
        sam = str(tmp_path / "stack.sam")
        with open(sam, "w") as fh:
            fh.write("@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:c\tLN:1000\n")
            fh.write("r\t0\tc\t100\t60\t10M\t*\t0\t0\t*\t*\n" * 262149)
    else:
        sam = os.path.join(GOLD, c["fixture"])
    bam = str(tmp_path / "in.bam")
    subprocess.run([BAMGEN, "-I", sam, "-o", bam], check=True)
    pre = str(tmp_path / "out")
    r = subprocess.run([SEAM, "-i", bam, "-o", pre] + args + ["-t", "3"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, env=env)
    assert r.returncode == 0, (r.stdout, r.stderr)
    for k, want in c["outputs"].items():
        with gzip.open(pre + k, "rt") as fh:
            assert_matches(want, fh.read(), f"{case}{k} (reference program + C ABI)")


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(SEAM), reason="oracle/_ref/pandepth_ref_seam not built (needs /root/reference at build time)")
@pytest.mark.parametrize("case", SEAM_CASES)
def test_reference_program_with_the_engine_behind_its_hot_loop(case, tmp_path):
    _run_seam(case, tmp_path)
