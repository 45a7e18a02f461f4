"""CPU: the patched reference program (oracle/_ref/pandepth_ref_seam, see tests/test_seam_gpu.py) with the C ABI answered by the
test stand-in -- checks the patch recipe of oracle/make_seam.py and the argument meaning at the boundary without a GPU."""
import os
import subprocess

import pytest

from tests.helpers import ROOT
from tests.test_seam_gpu import SEAM, SEAM_CASES, _run_seam

ORACLE_DIR = os.path.join(ROOT, "oracle", "_ref")

pytestmark = pytest.mark.skipif(not os.path.exists(SEAM), reason="oracle/_ref/pandepth_ref_seam not built (needs /root/reference at build time)")


@pytest.fixture(scope="module")
def standin_env(tmp_path_factory):
    d = tmp_path_factory.mktemp("standin_seam")
    so = str(d / "libpandepth_b200.so")
    subprocess.run(["g++", "-O2", "-std=c++14", "-fPIC", "-shared", os.path.join(ROOT, "tests", "stand_in", "pd_standin.cpp"),
                    "-L" + ORACLE_DIR, "-lpd_oracle", "-Wl,-rpath," + ORACLE_DIR, "-Wl,-soname,libpandepth_b200.so", "-o", so], check=True)
    return dict(os.environ, LD_LIBRARY_PATH=str(d) + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))


@pytest.mark.parametrize("case", [c for c in SEAM_CASES if not c.startswith("stack")])
def test_patched_reference_program_on_the_stand_in(case, tmp_path, standin_env):
    _run_seam(case, tmp_path, standin_env)
