"""
CPU: a short, fixed-seed leg of the randomised differential campaign (tools/stress_cli_cpu.py): the host program with the
test stand-in for the CUDA library against the reference binary on random BAM / CRAM / SAM / list / PAF inputs, random region
files and random option sets.  Longer legs are run by hand (`python tools/stress_cli_cpu.py 600`).
"""
import os
import subprocess
import sys

import pytest

from tests.helpers import ROOT

NEEDED = [os.path.join(ROOT, "pandepth_b200", "bin", "pandepth"), os.path.join(ROOT, "oracle", "_ref", "pandepth_ref"),
          os.path.join(ROOT, "oracle", "_ref", "bamgen"), os.path.join(ROOT, "oracle", "_ref", "libpd_oracle.so")]


@pytest.mark.skipif(not all(os.path.exists(p) for p in NEEDED), reason="host binary / oracle checkers not built")
def test_fixed_seed_leg_of_the_differential_campaign():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "stress_cli_cpu.py"), "n=8", "3000"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0 and " 0 differences" in r.stdout, r.stdout[-3000:]
