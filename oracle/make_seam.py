#!/usr/bin/env python
"""Prove the seam (TEST INFRASTRUCTURE): build the REFERENCE PROGRAM with its hot loop replaced by calls into the C ABI.

INTEGRATION.md section 2 says which lines of /root/reference/src/PanDepth.cpp a maintainer would replace.  This script does
exactly that on a scratch copy (nothing of the reference is stored in this repository: the edit is located by two short
anchors -- the `sam_itr_next` loop header and the statement that follows the loop -- inside `ProDealChrBambai`), compiles the
result with the reference's own flags and links it against pandepth_b200/libpandepth_b200.so:

    oracle/_ref/pandepth_ref_seam     =  PanDepth with `depth[p]++` (PanDepth.cpp:730-761) done by pd_push_contig + pd_depth_read

Everything else in that binary -- argument parsing, htslib iterators, the 10 Mb windowing, StatChrDepthWin, the report -- is the
reference's own code, so byte-identical outputs on the golden fixtures (tests/test_seam_gpu.py) show that the boundary of
include/pandepth_b200.h sits where INTEGRATION.md says it does.  Needs /root/reference (build container only); the binary
travels to the GPU box with the snapshot like the other oracle/_ref files.
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("REF", "/root/reference")

LOOP = "while ((ret = sam_itr_next(fphts, iter, aln)) >= 0)"
AFTER = "auto GeneDataIT=GeneData.find(ChrNum);"

REPLACEMENT = r'''
				// ---- pandepth_b200 seam: the filter + CIGAR walk of this window runs on the GPU (include/pandepth_b200.h) ----
				{
					std::vector<int32_t> pd_pos; std::vector<uint32_t> pd_cig, pd_off(1, 0u);
					while ((ret = sam_itr_next(fphts, iter, aln)) >= 0)
					{
						cigar = bam_get_cigar(aln);
						pd_pos.push_back((int32_t)((aln->core).pos) - ShiftPosition);
						pd_flag.push_back((uint16_t)aln->core.flag); pd_mapq.push_back((uint8_t)(aln->core).qual);
						pd_cig.insert(pd_cig.end(), cigar, cigar + aln->core.n_cigar);
						pd_off.push_back((uint32_t)pd_cig.size());
					}
					uint32_t pd_len = (uint32_t)ThisWinDowLen;
					pd_engine *pd_e = pd_create(0, 1, &pd_len, NULL, flags, (paraFA04->InInt));      // -x / -q: the engine applies :732-733
					if (!pd_e
					    || pd_push_contig(pd_e, 0, (int64_t)pd_pos.size(), pd_pos.data(), pd_flag.data(), pd_mapq.data(), pd_off.data(), pd_cig.data(), PD_SORTED) != PD_OK
					    || pd_depth_read(pd_e, 0, 0, (int32_t)pd_len, depth, 0) != PD_OK)
					{ cerr << "pandepth_b200: " << pd_last_error(pd_e) << endl; exit(3); }
					pd_destroy(pd_e);
					pd_flag.clear(); pd_mapq.clear();
				}

				'''

PROLOGUE = '''
// ---- pandepth_b200 seam (oracle/make_seam.py) ----
#include <vector>
#include <stdint.h>
extern "C" {
#include "pandepth_b200.h"
}
static thread_local std::vector<uint16_t> pd_flag;
static thread_local std::vector<uint8_t> pd_mapq;
'''


def main():
    src_path = os.path.join(REF, "src", "PanDepth.cpp")
    if not os.path.exists(src_path):
        print("make_seam: %s absent -- keeping a prebuilt oracle/_ref/pandepth_ref_seam if there is one" % src_path)
        return 0
    src = open(src_path, encoding="utf-8", errors="surrogateescape").read()
    f0 = src.index("void ProDealChrBambai (")
    a = src.index(LOOP, f0)
    b = src.index(AFTER, a)
    # the statement before the loop header starts the edit at the beginning of the loop's own line
    a_line = src.rfind("\n", 0, a) + 1
    b_line = src.rfind("\n", 0, b) + 1
    patched = src[:a_line] + REPLACEMENT.lstrip("\n") + src[b_line:]
    # prologue behind the last #include of the file's head
    last_inc = [m for m in re.finditer(r"^#include[^\n]*\n", patched[:patched.index("void ProDealChrBambai (")], re.M)][-1]
    patched = patched[:last_inc.end()] + PROLOGUE + patched[last_inc.end():]
    out_dir = os.path.join(HERE, "_ref", "seam")
    os.makedirs(out_dir, exist_ok=True)
    cpp = os.path.join(out_dir, "PanDepth_seam.cpp")
    open(cpp, "w", encoding="utf-8", errors="surrogateescape").write(patched)
    exe = os.path.join(HERE, "_ref", "pandepth_ref_seam")
    cmd = ["g++", "-L" + os.path.join(REF, "lib"), "--std=c++11", "-g", "-O3", cpp, "-lhts", "-ldeflate", "-lz", "-pthread",
           "-I" + os.path.join(REF, "include"), "-I" + os.path.join(REF, "src"), "-I" + os.path.join(ROOT, "include"),
           "-L" + os.path.join(ROOT, "pandepth_b200"), "-lpandepth_b200", "-Wl,-rpath,$ORIGIN/../../pandepth_b200", "-o", exe]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout[-4000:])
        return 1
    os.remove(cpp)                               # the scratch copy holds reference source: it does not stay around
    print("make_seam: built", exe)
    return 0


if __name__ == "__main__":
    sys.exit(main())
