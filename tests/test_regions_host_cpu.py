"""
CPU: the C++ host's region model (pandepth_b200/host/regions.cpp: GFF/GTF/BED parsers, pseudo-gene tiling, region merge,
the indexed-window hand-over count, output mode and depth mask) must equal the oracle's model (oracle/model.py, itself pinned
byte-for-byte to the reference binary by tests/test_oracle_golden.py) for every golden option set and for seeded random region
files.  `PANDEPTH_DRY=1 PANDEPTH_DUMP_REGIONS=<file>` makes the CLI write the model as text without touching a GPU.
"""
import json
import os
import subprocess

import numpy as np
import pytest

from oracle import model
from tests.helpers import GOLD, ROOT, load_fixture, options_from_args

CLI = os.path.join(ROOT, "pandepth_b200", "bin", "pandepth")
BAMGEN = os.path.join(ROOT, "oracle", "_ref", "bamgen")

pytestmark = pytest.mark.skipif(not (os.path.exists(CLI) and os.path.exists(BAMGEN)), reason="host binary / bamgen not built")

with open(os.path.join(GOLD, "expected.json")) as _fh:
    _CASES = {k: v for k, v in json.load(_fh).items()
              if v["kind"] in ("bam", "sam") and not v["fixture"].startswith("list:") and not v["fixture"].startswith("@")}


def _dump(inp, args, tmp):
    out = os.path.join(tmp, "regions.txt")
    env = dict(os.environ, PANDEPTH_DRY="1", PANDEPTH_DUMP_REGIONS=out)
    r = subprocess.run([CLI, "-i", inp, "-o", os.path.join(tmp, "o")] + args, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    assert r.returncode == 0, (r.stdout, r.stderr)
    return open(out).read()


def _expected(targets, opts):
    """The same text from the oracle's model."""
    mode, genes = model.build_region_model(targets, opts)
    whole = (not opts.indexed) or opts.site_output or mode == 6
    lines = [f"mode {mode} mask {0x3FFFF if whole else 0xFFFFFFFF:x}"]
    for t in sorted(genes):
        g = genes[t]
        for gid in sorted(g, key=lambda s: s.encode()):                    # std::map<string> order
            x = g[gid]
            lines.append("\t".join(["G", str(t), gid, str(x.start), str(x.end), str(x.length)] + [f"{a},{b}" for a, b in x.cds]))
        merged = model.merge_regions(g)
        lines.append("\t".join(["M", str(t)] + [f"{a},{b}" for a, b in merged]))
        cnt = model.indexed_window_skips(merged, g, targets[t][1])
        for gid in sorted(cnt, key=lambda s: s.encode()):
            lines.append(f"W\t{t}\t{gid}\t{cnt[gid]}")
    return "\n".join(lines) + "\n"


@pytest.fixture(scope="module")
def bams(tmp_path_factory):
    d = tmp_path_factory.mktemp("regions")
    out = {}
    for fx in ("k.sam", "mix.sam", "big.sam"):
        out[fx] = str(d / (fx + ".bam"))
        subprocess.run([BAMGEN, "-I", os.path.join(GOLD, fx), "-o", out[fx]], check=True)
    return out


@pytest.mark.parametrize("case", sorted(_CASES))
def test_region_model_matches_oracle_on_golden_option_sets(case, bams, tmp_path):
    c = _CASES[case]
    targets, _, _ = load_fixture(c["fixture"])
    opts = options_from_args(c["args"], c["kind"], c["fixture"])
    args = [os.path.join(GOLD, a[1:]) if a.startswith("@") else a for a in c["args"]]
    inp = bams[c["fixture"]] if c["kind"] == "bam" else os.path.join(GOLD, c["fixture"])
    assert _dump(inp, args, str(tmp_path)) == _expected(targets, opts), case


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_region_model_matches_oracle_on_random_region_files(seed, tmp_path):
    """Seeded BED3 / BED4 / GFF3 files with overlapping, nested, touching, reversed, out-of-range and unknown-contig records,
    contig lengths around the tiling quirks (len < 2, len % W == 1), on an empty-read SAM header."""
    rng = np.random.default_rng(seed)
    lens = [int(x) for x in (1, 2, 151, 1001, 20_000_001, 10_000_001, 10_000_000, 300, int(rng.integers(1000, 90_000)))]
    targets = [(f"c{i}", ln) for i, ln in enumerate(lens)]
    sam = str(tmp_path / "h.sam")
    with open(sam, "w") as fh:
        fh.write("@HD\tVN:1.6\tSO:coordinate\n" + "".join(f"@SQ\tSN:{n}\tLN:{l}\n" for n, l in targets))
        fh.write("r\t0\tc3\t5\t60\t10M\t*\t0\t0\t*\t*\n")
    bam = str(tmp_path / "h.bam")
    subprocess.run([BAMGEN, "-I", sam, "-o", bam], check=True)

    def intervals(n):
        for _ in range(n):
            t = int(rng.integers(0, len(lens) + 1))
            nm = f"c{t}" if t < len(lens) else "nowhere"
            ln = lens[t] if t < len(lens) else 1000
            s = int(rng.integers(1, max(2, ln)))
            e = s + int(rng.integers(-3, max(2, min(ln, 5000))))
            if rng.integers(0, 10) == 0:
                e = ln + int(rng.integers(0, 50))                            # runs past the contig end
            yield nm, s, e

    bed3 = str(tmp_path / "r.bed3"); bed4 = str(tmp_path / "r.bed4"); gff = str(tmp_path / "r.gff")
    with open(bed3, "w") as f3, open(bed4, "w") as f4, open(gff, "w") as fg:
        f3.write("#comment\n"); fg.write("##gff-version 3\n")
        for k, (nm, s, e) in enumerate(intervals(400)):
            f3.write(f"{nm}\t{s}\t{e}\n")
            f4.write(f"{nm}\t{s}\t{e}\tg{k // 3}\n")
            if e >= s:
                fg.write(f"{nm}\tsrc\t{'CDS' if k % 4 else 'exon'}\t{s}\t{e}\t.\t+\t0\tID=x{k};Parent=t{k // 5}\n")
    for args, kw in ((["-b", bed3], {"bed": bed3}), (["-b", bed4], {"bed": bed4}), (["-g", gff], {"gff": gff}),
                     (["-g", gff, "-f", "exon"], {"gff": gff, "feature": "exon"}),
                     ([], {}), (["-w", "149"], {"win": 149}), (["-w", "150"], {"win": 150}), (["-w", "999999"], {"win": 999999}),
                     (["-a"], {"site_output": True}), (["-s", "-b", bed4], {"bed": bed4, "indexed": False})):
        opts = model.Options(**kw)
        assert _dump(bam, args, str(tmp_path)) == _expected(targets, opts), (seed, args)
