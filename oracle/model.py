"""
oracle/model.py -- whole-output CPU model of the PanDepth coverage path (TEST INFRASTRUCTURE ONLY).

Restates, around the C loops in pd_oracle.c, the reference's region model and report writer so that a
complete `.stat` text can be produced from (header, records, options) and compared byte-for-byte with
the reference binary's decompressed output.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.

Parity status: PINNED -- tests/test_oracle_golden.py compares this model against outputs of the
reference binary (oracle/_ref/pandepth_ref) stored under tests/golden/.

All citations are /root/reference/src/PanDepth.cpp unless noted.
"""
from __future__ import annotations

import ctypes
import gzip
import os
import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libpd_oracle.so")
_lib = None


def lib() -> ctypes.CDLL:
    """Load oracle/_ref/libpd_oracle.so (built by oracle/Makefile)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError(f"{_LIB_PATH} missing -- run `make -C oracle` (or __graft_entry__.build())")
        L = ctypes.CDLL(_LIB_PATH)
        p = ctypes.c_void_p
        L.pdo_accumulate.restype = ctypes.c_int64
        L.pdo_accumulate.argtypes = [ctypes.c_int, p, p, ctypes.c_int64, p, p, p, p, p, p, ctypes.c_uint32, ctypes.c_int]
        L.pdo_reduce.restype = None
        L.pdo_reduce.argtypes = [p, ctypes.c_uint32, ctypes.c_int64, p, p, ctypes.c_uint32, ctypes.c_uint32, p, p]
        L.pdo_tile_contig.restype = ctypes.c_int64
        L.pdo_tile_contig.argtypes = [ctypes.c_int64, ctypes.c_int64, p, p]
        L.pdo_mode6_windows.restype = ctypes.c_int64
        L.pdo_mode6_windows.argtypes = [ctypes.c_int64, ctypes.c_int64, p, p]
        L.pdo_format_ratio.restype = None
        L.pdo_format_ratio.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_char_p, ctypes.c_char_p]
        _lib = L
    return _lib


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


# --------------------------------------------------------------------------------------------------
# Records
# --------------------------------------------------------------------------------------------------
_CIG_OPS = "MIDNSHP=XB"
_CIG_RE = re.compile(r"(\d+)([MIDNSHP=XB])")


@dataclass
class Records:
    """SoA batch in the BAM core's own terms (include/sam.h:204-217, cigar packing :104-115)."""
    tid: np.ndarray        # int32
    pos: np.ndarray        # int32, 0-based
    flag: np.ndarray       # uint16
    mapq: np.ndarray       # uint8
    cigar_off: np.ndarray  # uint32, n+1
    cigar: np.ndarray      # uint32, len<<4|op

    @property
    def n(self) -> int:
        return int(self.tid.shape[0])

    def select(self, idx: np.ndarray) -> "Records":
        idx = np.asarray(idx, dtype=np.int64)
        lens = (self.cigar_off[1:] - self.cigar_off[:-1]).astype(np.int64)[idx]
        off = np.zeros(idx.size + 1, dtype=np.uint32)
        off[1:] = np.cumsum(lens)
        cig = np.empty(int(off[-1]), dtype=np.uint32)
        for j, i in enumerate(idx):
            cig[off[j]:off[j + 1]] = self.cigar[self.cigar_off[i]:self.cigar_off[i + 1]]
        return Records(self.tid[idx].copy(), self.pos[idx].copy(), self.flag[idx].copy(), self.mapq[idx].copy(), off, cig)


def encode_cigar(s: str) -> List[int]:
    if s == "*":
        return []
    return [(int(n) << 4) | _CIG_OPS.index(op) for n, op in _CIG_RE.findall(s)]


def records_from_tuples(rows, names: List[str]) -> Records:
    """rows: iterable of (flag, contig_name_or_tid, pos1, mapq, cigar_string)."""
    tid, pos, flag, mapq, off, cig = [], [], [], [], [0], []
    for fl, c, p1, mq, cs in rows:
        t = c if isinstance(c, int) else (names.index(c) if c in names else -1)
        tid.append(t); pos.append(int(p1) - 1); flag.append(int(fl)); mapq.append(int(mq))
        cig.extend(encode_cigar(cs)); off.append(len(cig))
    return Records(np.array(tid, np.int32), np.array(pos, np.int32), np.array(flag, np.uint16),
                   np.array(mapq, np.uint8), np.array(off, np.uint32), np.array(cig, np.uint32))


def parse_sam(path: str) -> Tuple[List[Tuple[str, int]], Records, bool]:
    """Minimal SAM text reader -> (targets, records, sorted_by_coordinate)."""
    targets: List[Tuple[str, int]] = []
    rows = []
    so = False
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rt") as fh:
        for line in fh:
            if line.startswith("@"):
                if line.startswith("@SQ"):
                    f = dict(x.split(":", 1) for x in line.rstrip("\n").split("\t")[1:])
                    targets.append((f["SN"], int(f["LN"])))
                elif line.startswith("@HD") and "\tSO:coordinate" in line:
                    so = True
                continue
            f = line.rstrip("\n").split("\t")
            if len(f) < 6:
                continue
            rows.append((int(f[1]), f[2], int(f[3]), int(f[4]), f[5]))
    names = [t[0] for t in targets]
    return targets, records_from_tuples(rows, names), so


def parse_paf(paths: List[str], xflags: int = 1796, min_mapq: int = -1) -> Tuple[List[Tuple[str, int]], Records, bool]:
    """PAF text -> (targets, records, False), following paf_main (:852-2024).

    Targets: column 6 / 7 of the FIRST file in first-seen order (:918-941).  Per line of every file (:1536-1623): dropped if
    it contains "tp:A:S" and -x has bit 0x100, or if column 12 < -q; tid = Chr2IntMap[col 6] (unknown names default-insert
    0); start/end = columns 8/9, swapped if start > end; a cg:Z: token at index > 1 is walked from the 0-based start (M = X
    add depth, D N advance, anything else nothing), otherwise the line covers [start-1, end).  The filters are applied here,
    so the records carry flag 0 / mapq 255."""
    chr2int: Dict[str, int] = {}
    targets: List[Tuple[str, int]] = []
    opener = lambda p: gzip.open(p, "rt") if p.endswith(".gz") else open(p, "rt")
    with opener(paths[0]) as fh:
        for line in fh:
            f = line.split()
            if len(f) < 7:
                continue
            if f[5] not in chr2int:
                chr2int[f[5]] = len(targets)
                targets.append((f[5], _atoi(f[6])))
    tid, pos, off, cig = [], [], [0], []
    for path in paths:
        with opener(path) as fh:
            for line in fh:
                line = line.rstrip("\n")
                if not line:
                    continue
                if (xflags & 0x100) and "tp:A:S" in line:
                    continue
                f = _cpp_split(line, " \t")
                if len(f) < 12:
                    continue
                t = chr2int.get(f[5], 0)
                if _atoi(f[11]) < min_mapq:
                    continue
                a, z = _atoi(f[7]), _atoi(f[8])
                if a > z:
                    a, z = z, a
                cg = next((i for i, x in enumerate(f) if x[:5] == "cg:Z:"), -1)
                if cg > 1:
                    ops = []
                    for n, op in re.findall(r"(\d*)(\D|$)", f[cg][5:]):
                        if not n and not op:
                            continue
                        code = {"M": 0, "=": 7, "X": 8, "D": 2, "N": 3}.get(op)
                        if code is not None:
                            ops.append((int(n or 0) << 4) | code)
                    pos.append(a)
                else:
                    s0 = a - 1 if a > 0 else 0          # the reference's write at index -1 is dropped
                    ops = [((z - s0) << 4) | 0] if z > s0 else []
                    pos.append(s0)
                tid.append(t); cig.extend(ops); off.append(len(cig))
    n = len(tid)
    rec = Records(np.array(tid, np.int32), np.array(pos, np.int32), np.zeros(n, np.uint16), np.full(n, 255, np.uint8),
                  np.array(off, np.uint32), np.array(cig, np.uint32))
    return targets, rec, False


# --------------------------------------------------------------------------------------------------
# L1 / L0 through the C restatement
# --------------------------------------------------------------------------------------------------
def accumulate(target_len: np.ndarray, rec: Records, xflags: int = 1796, min_mapq: int = -1,
               use_filter: bool = True) -> List[np.ndarray]:
    """Per-contig uint32 depth arrays (pd_oracle.c:pdo_accumulate; reference :732-760)."""
    L = lib()
    target_len = np.ascontiguousarray(target_len, dtype=np.uint32)
    depth = [np.zeros(int(n), dtype=np.uint32) for n in target_len]
    ptrs = (ctypes.c_void_p * len(depth))(*[d.ctypes.data for d in depth])
    L.pdo_accumulate(len(depth), _ptr(target_len), ctypes.cast(ptrs, ctypes.c_void_p), rec.n,
                     _ptr(rec.tid), _ptr(rec.pos), _ptr(rec.flag) if use_filter else None,
                     _ptr(rec.mapq) if use_filter else None, _ptr(rec.cigar_off), _ptr(rec.cigar),
                     xflags, min_mapq)
    return depth


def reduce_intervals(depth: np.ndarray, start0: np.ndarray, end0: np.ndarray, min_dep: int = 1,
                     mask: int = 0xFFFFFFFF) -> Tuple[np.ndarray, np.ndarray]:
    """(cov, sum) per half-open 0-based interval (pd_oracle.c:pdo_reduce; reference :295-348)."""
    L = lib()
    start0 = np.ascontiguousarray(start0, dtype=np.int32)
    end0 = np.ascontiguousarray(end0, dtype=np.int32)
    cov = np.zeros(start0.size, dtype=np.uint64)
    s = np.zeros(start0.size, dtype=np.uint64)
    L.pdo_reduce(_ptr(depth), int(depth.size), int(start0.size), _ptr(start0), _ptr(end0), min_dep, mask, _ptr(cov), _ptr(s))
    return cov, s


def tile_contig(length: int, W: int) -> Tuple[np.ndarray, np.ndarray]:
    L = lib()
    n = L.pdo_tile_contig(length, W, None, None)
    a = np.zeros(n, np.int32); b = np.zeros(n, np.int32)
    L.pdo_tile_contig(length, W, _ptr(a), _ptr(b))
    return a, b


def mode6_windows(length: int, W: int) -> Tuple[np.ndarray, np.ndarray]:
    L = lib()
    n = L.pdo_mode6_windows(length, W, None, None)
    a = np.zeros(n, np.int32); b = np.zeros(n, np.int32)
    L.pdo_mode6_windows(length, W, _ptr(a), _ptr(b))
    return a, b


def fmt_ratio(cov: int, total: int, length: int) -> Tuple[str, str]:
    if length == 0:  # 0/0 -> the reference prints "-nan" through iostream; glibc printf agrees
        length = 0
    a = ctypes.create_string_buffer(32); b = ctypes.create_string_buffer(32)
    lib().pdo_format_ratio(int(cov), int(total), int(length), a, b)
    return a.value.decode(), b.value.decode()


# --------------------------------------------------------------------------------------------------
# Region model (host text parsing; restated because it defines the interval table)
# --------------------------------------------------------------------------------------------------
@dataclass
class Gene:
    """include/DataClass.h:62-80 GeneInfo."""
    start: int
    end: int
    length: int
    cds: List[Tuple[int, int]] = field(default_factory=list)   # 1-based closed
    cover: int = 0
    depth: int = 0


@dataclass
class Options:
    """include/DataClass.h:22-59 In3str1v defaults; InInt=-1 preset at :3444."""
    min_mapq: int = -1      # -q
    xflags: int = 1796      # -x
    min_dep: int = 1        # -d (clamped >=1 at :233)
    win: int = 0            # -w (clamped >=1 at :211-215 when given)
    site_output: bool = False   # -a
    feature: str = "CDS"    # -f
    gff: Optional[str] = None   # -g
    bed: Optional[str] = None   # -b
    indexed: bool = True    # False for -s / SAM input / missing index (:4124)


def _cpp_split(s: str, delims: str) -> List[str]:
    """include/comm.h:98-109 split(): tokens between any of the delimiter characters, empties dropped."""
    out, cur = [], []
    for ch in s:
        if ch in delims:
            if cur:
                out.append("".join(cur)); cur = []
        else:
            cur.append(ch)
    if cur:
        out.append("".join(cur))
    return out


def _open_text(path: str):
    with open(path, "rb") as fh:
        magic = fh.read(2)
    return gzip.open(path, "rt") if magic == b"\x1f\x8b" else open(path, "rt")


def sniff_gff_mode(path: str) -> int:
    """:162-181 -- first 167 lines; the LAST matching line decides (Parent -> 1, transcript_id -> 2)."""
    mode = 0
    with _open_text(path) as fh:
        for k, line in enumerate(fh):
            if k >= 167:
                break
            line = line.rstrip("\n")
            if len(line) < 2 or line[0] == "#":
                continue
            if "Parent" in line:
                mode = 1
            elif "transcript_id" in line:
                mode = 2
    return mode


def sniff_bed_mode(path: str) -> int:
    """:268-288 -- BED4 if either of the first two lines has exactly 4 tokens."""
    with _open_text(path) as fh:
        l1 = fh.readline().rstrip("\n"); l2 = fh.readline().rstrip("\n")
    return 4 if (len(_cpp_split(l1, " \t")) == 4 or len(_cpp_split(l2, " \t")) == 4) else 3


def _add_gene(genes: Dict[int, Dict[str, Gene]], tid: int, gid: str, s: int, e: int):
    g = genes.setdefault(tid, {})
    if gid not in g:                                            # :3601-3604
        g[gid] = Gene(s, e, e - s + 1, [(s, e)])
    else:                                                       # :3638-3641
        x = g[gid]
        x.start = min(x.start, s); x.end = max(x.end, e); x.length += e - s + 1; x.cds.append((s, e))


def _atoi(s: str) -> int:
    m = re.match(r"\s*([+-]?\d+)", s)
    return int(m.group(1)) if m else 0


def parse_regions(opts: Options, names: List[str]) -> Tuple[int, Dict[int, Dict[str, Gene]]]:
    """Returns (mode, genes) for modes 1-4 (:3547-3905)."""
    genes: Dict[int, Dict[str, Gene]] = {}
    idx = {}
    for i, n in enumerate(names):
        idx.setdefault(n, i)                                     # map insert keeps the first (:3502)
    if opts.gff:
        mode = sniff_gff_mode(opts.gff)
        if mode == 0:
            raise ValueError("GFF/GTF format not recognised (:185-189)")
        with _open_text(opts.gff) as fh:
            for line in fh:
                line = line.rstrip("\n")
                if len(line) <= 0 or line[0] == "#":
                    continue
                if mode == 2:
                    line = line.replace('"', "").replace(";", "")   # :3657-3658
                tok = line.split()
                if len(tok) < 3 or tok[2] != opts.feature:
                    continue
                s = _atoi(tok[3]) if len(tok) > 3 else 0
                e = _atoi(tok[4]) if len(tok) > 4 else 0
                if mode == 1:
                    attr = tok[8] if len(tok) > 8 else ""
                    inf = _cpp_split(attr, ",;")                     # :3575
                    gid = _cpp_split(inf[0], "=")[-1] if inf else ""
                    for x in inf[1:]:
                        t2 = _cpp_split(x, "=")
                        if t2 and t2[0] == "Parent":
                            gid = t2[-1]
                else:
                    inf = _cpp_split(line, "\t ")                    # :3668-3669 (transcript_id branch is dead code)
                    gid = inf[9] if len(inf) > 9 else ""
                if tok[0] not in idx:
                    continue                                         # :3590-3593 warning, skipped
                _add_gene(genes, idx[tok[0]], gid, s, e)
        return mode, genes
    if opts.bed:
        mode = sniff_bed_mode(opts.bed)
        with _open_text(opts.bed) as fh:
            for line in fh:
                line = line.rstrip("\n")
                if len(line) <= 0 or line[0] == "#":
                    continue
                tok = line.split()
                if len(tok) < 3:
                    continue
                s, e = _atoi(tok[1]), _atoi(tok[2])
                gid = f"{tok[0]}_{tok[1]}_{tok[2]}" if mode == 3 else (tok[3] if len(tok) > 3 else "")  # :3752 / :3831
                if s > e:
                    continue                                         # :3754-3758
                if tok[0] not in idx:
                    continue
                _add_gene(genes, idx[tok[0]], gid, s, e)
        return mode, genes
    return 0, genes


def merge_regions(genes: Dict[str, Gene]) -> List[Tuple[int, int]]:
    """:3919-3972 -- merge gene spans of one contig into sorted, strictly separated regions."""
    by_start: Dict[int, int] = {}
    for g in genes.values():
        by_start[g.start] = max(by_start.get(g.start, g.end), g.end) if g.start in by_start else g.end
    items = sorted(by_start.items())
    out = [list(items[0])]
    End = items[0][1]
    for s, e in items[1:]:
        if s > End:
            out.append([s, e]); End = e
        elif e > End:
            out[-1][1] = e; End = e
    return [(a, b) for a, b in out]


def indexed_window_skips(regions: List[Tuple[int, int]], genes: Dict[str, Gene], chr_len: int) -> Dict[str, int]:
    """How many times ProDealChrBambai's windowing (:695-784) hands each gene to StatChrDepthWin (:299-308).

    Normally exactly once; a gene whose GeneStart >= the (contig-clipped) MeMEnd of its window is skipped
    unless the window is degenerate (MeMEnd == MeMStart)."""
    W = 10_000_000
    count = {k: 0 for k in genes}
    if not regions:
        return count
    MeMStart = max(regions[0][0], 1)
    MeMEnd = min(MeMStart + W - 1, chr_len)
    for k, (first, second) in enumerate(regions):
        end = min(second + 1, chr_len)
        last = (k + 1 == len(regions))
        if end >= MeMEnd or last:
            MeMEnd = end
            for gid, g in genes.items():
                if MeMEnd - MeMStart != 0:
                    if g.start >= MeMEnd or g.end < MeMStart:
                        continue
                else:
                    if g.start > MeMEnd or g.end < MeMStart:
                        continue
                count[gid] += 1
            if not last:
                MeMStart = regions[k + 1][0]
                if MeMStart - 150 > MeMEnd:
                    MeMStart -= 150
            MeMEnd = min(MeMStart + W, chr_len)
    return count


# --------------------------------------------------------------------------------------------------
# Whole run -> .stat text
# --------------------------------------------------------------------------------------------------
HEADERS = {  # :4069-4090
    0: "#Chr\tLength\tCoveredSite\tTotalDepth\tCoverage(%)\tMeanDepth\n",
    1: "#Chr\tStart\tEnd\tGeneID\tLength\tCoveredSite\tTotalDepth\tCoverage(%)\tMeanDepth\n",
    2: "#Chr\tStart\tEnd\tGeneID\tLength\tCoveredSite\tTotalDepth\tCoverage(%)\tMeanDepth\n",
    3: "#Chr\tStart\tEnd\tRegionID\tLength\tCoveredSite\tTotalDepth\tCoverage(%)\tMeanDepth\n",
    4: "#Chr\tStart\tEnd\tGeneID\tLength\tCoveredSite\tTotalDepth\tCoverage(%)\tMeanDepth\n",
    5: "#Chr\tStart\tEnd\tLength\tCoveredSite\tTotalDepth\tCoverage(%)\tMeanDepth\n",
    6: "#Chr\tStart\tEnd\tLength\tCoveredSite\tTotalDepth\tCoverage(%)\tMeanDepth\n",
}
SUFFIX = {0: ".chr.stat.gz", 1: ".gene.stat.gz", 2: ".gene.stat.gz", 3: ".bed.stat.gz", 4: ".bed.stat.gz",
          5: ".win.stat.gz", 6: ".win.stat.gz"}


def _wrap_i32(v: int) -> int:
    v &= 0xFFFFFFFF
    return v - (1 << 32) if v & 0x80000000 else v


def build_region_model(targets: List[Tuple[str, int]], opts: Options) -> Tuple[int, Dict[int, Dict[str, Gene]]]:
    """(mode, genes) as main() builds them: -g/-b regions (:3547-3905), else pseudo-gene tiles (:3975-4051)."""
    names = [t[0] for t in targets]
    mode, genes = parse_regions(opts, names)
    if not genes:                                                   # :3975-4051
        if opts.win == 0:
            mode, W = 0, 10_000_000
        elif opts.win < 150:
            mode, W = 6, 10_000_000
        else:
            mode, W = 5, opts.win
        for t, (nm, ln) in enumerate(targets):
            a, b = tile_contig(ln, W)
            for s, e in zip(a.tolist(), b.tolist()):
                _add_gene(genes, t, f"{nm}{s}", s, e)
    return mode, genes


def _endpos(rec: Records, i: int) -> int:
    """bam_endpos(): pos + reference length of the CIGAR; pos + 1 for a read flagged unmapped or without a reference-consuming op."""
    p = int(rec.pos[i])
    if int(rec.flag[i]) & 4:
        return p + 1
    n = 0
    for c in rec.cigar[rec.cigar_off[i]:rec.cigar_off[i + 1]]:
        if (0x18d >> (int(c) & 15)) & 1:                 # M D N = X
            n += int(c) >> 4
    return p + (n if n else 1)


def reads_reaching_depth(targets: List[Tuple[str, int]], rec: Records, opts: Options, sorted_hdr: bool = True) -> np.ndarray:
    """Indices of the records the reference hands to its depth loop in region modes (-g / -b), in file order.

    Only the per-site dump (-a) can tell: region statistics never look between regions.
      indexed    : reads returned by the `chr:first-1..second+1` queries of the merged regions (:700-706, :726)
      no index   : every read, unless the header says SO:coordinate -- then the per-contig region cursor of :4610-4646 skips
                   reads ending before the cursor's region, walks the cursor forward, closes a contig after its last region
                   and stops the file once every contig is closed.
    The record filter (:732-733) is NOT applied here."""
    mode, genes = build_region_model(targets, opts)
    if mode not in (1, 2, 3, 4):
        return np.arange(rec.n)
    merged = {t: merge_regions(g) for t, g in genes.items() if g}
    keep = []
    if opts.indexed:
        for i in range(rec.n):
            t = int(rec.tid[i])
            if t not in merged:
                continue
            p, e = int(rec.pos[i]), _endpos(rec, i)
            for first, second in merged[t]:
                wb = max(first - 1, 1) - 1
                if wb >= e:
                    break
                if second + 1 > p:
                    keep.append(i)
                    break
        return np.array(keep, np.int64)
    if not sorted_hdr:
        return np.arange(rec.n)
    at = {t: 0 for t in merged}
    done = {t: False for t in merged}
    open_contigs = len(merged)
    for i in range(rec.n):
        t = int(rec.tid[i])
        if t < 0 or t >= len(targets) or t not in merged or done[t]:
            continue
        if int(rec.mapq[i]) < opts.min_mapq or (int(rec.flag[i]) & opts.xflags):
            continue
        p, e = int(rec.pos[i]), _endpos(rec, i)
        regs = merged[t]
        k = at[t]
        if e < regs[k][0]:
            continue
        stop = False
        if p > regs[k][1]:
            k += 1
            while k < len(regs) and p > regs[k][1]:
                k += 1
            if k == len(regs):
                done[t] = True
                open_contigs -= 1
                if open_contigs == 0:
                    stop = True
            at[t] = min(k, len(regs) - 1)
        if stop:
            break
        keep.append(i)
    return np.array(keep, np.int64)


def run_model(targets: List[Tuple[str, int]], rec: Records, opts: Options,
              depth_override: Optional[List[np.ndarray]] = None, sorted_hdr: bool = True) -> Dict[str, str]:
    """Produce {'stat': text, 'suffix': str, 'site': text-or-None} exactly as the reference would print them.

    depth_override lets tests substitute per-base depth produced by the CUDA engine for the oracle's own."""
    names = [t[0] for t in targets]
    tlen = np.array([t[1] for t in targets], dtype=np.uint32)
    mode, genes = build_region_model(targets, opts)
    whole_contig = (not opts.indexed) or opts.site_output or mode == 6   # :4124-4127 -> SiteInfo 18-bit paths
    mask = 0x3FFFF if whole_contig else 0xFFFFFFFF

    if depth_override is not None:
        depth = depth_override
    elif opts.site_output and mode in (1, 2, 3, 4):
        depth = accumulate(tlen, rec.select(reads_reaching_depth(targets, rec, opts, sorted_hdr)), opts.xflags, opts.min_mapq)
    else:
        depth = accumulate(tlen, rec, opts.xflags, opts.min_mapq)

    # L0: per-gene reduce
    for t in sorted(genes):
        g = genes[t]
        times = None
        if not whole_contig:
            times = indexed_window_skips(merge_regions(g), g, int(tlen[t]))
        for gid, x in g.items():
            s0 = np.array([c[0] - 1 for c in x.cds], np.int32)
            e0 = np.array([c[1] for c in x.cds], np.int32)
            cov, tot = reduce_intervals(depth[t], s0, e0, opts.min_dep, mask)
            k = 1 if times is None else times[gid]
            x.cover = int(cov.sum()) * k
            x.depth = int(tot.sum()) * k

    out: List[str] = [HEADERS[mode]]
    SS_Len = SS_Cov = SS_Tot = 0
    site = None
    if opts.site_output:                                            # :4264-4284
        lines = []
        for t, (nm, ln) in enumerate(targets):
            if t not in genes:
                continue
            d = depth[t] & np.uint32(0x3FFFF)
            lines.append("".join(f"{nm}\t{j}\t{int(v)}\n" for j, v in enumerate(d.tolist())))
        site = "".join(lines)

    if mode == 6:                                                   # :4353-4394
        for t, (nm, ln) in enumerate(targets):
            if t not in genes:
                continue
            a, b = mode6_windows(ln, opts.win)
            cov, tot = reduce_intervals(depth[t], a, b, opts.min_dep, 0x3FFFF)
            for s0, e0, c, d in zip(a.tolist(), b.tolist(), cov.tolist(), tot.tolist()):
                j = s0 + 1
                L = e0 - j + 1
                c32, d32 = _wrap_i32(c), _wrap_i32(d)                # int GeneCover, GeneDepth (:4364)
                cp, md = fmt_ratio_signed(c32, d32, L)
                out.append(f"{nm}\t{j}\t{e0}\t{L}\t{c32}\t{d32}\t{cp}\t{md}\n")
                SS_Cov = (SS_Cov + c32) & 0xFFFFFFFFFFFFFFFF
                SS_Len += L
                SS_Tot = (SS_Tot + d32) & 0xFFFFFFFFFFFFFFFF
    elif mode in (1, 2, 3, 4, 5):                                   # :5011-5050, :5080-5118
        for t in sorted(genes):
            nm = names[t]
            rows: Dict[int, List[str]] = {}
            for gid in sorted(genes[t], key=lambda s: s.encode()):   # std::map<string> order
                x = genes[t][gid]
                cp, md = fmt_ratio_signed(_wrap_i32(x.cover), x.depth, x.length)
                SS_Cov += x.cover; SS_Len += x.length; SS_Tot += x.depth
                if mode == 5:
                    row = f"{nm}\t{x.start}\t{x.end}\t{x.length}\t{x.cover}\t{x.depth}\t{cp}\t{md}"
                else:
                    row = f"{nm}\t{x.start}\t{x.end}\t{gid}\t{x.length}\t{x.cover}\t{x.depth}\t{cp}\t{md}"
                rows.setdefault(x.start, []).append(row)
            for st in sorted(rows):
                out.append("\n".join(rows[st]) + "\n")
    else:                                                           # mode 0 :5051-5079
        for t in sorted(genes):
            L = C = D = 0
            for x in genes[t].values():
                L += x.length; C += x.cover; D += x.depth
            SS_Cov += C; SS_Len += L; SS_Tot += D
            cp, md = fmt_ratio(C, D, L)
            out.append(f"{names[t]}\t{L}\t{C}\t{D}\t{cp}\t{md}\n")
    cp, md = fmt_ratio(SS_Cov, SS_Tot, SS_Len)
    out.append(f"##RegionLength: {SS_Len}\tCoveredSite: {SS_Cov}\tCoverage(%): {cp}\tMeanDepth: {md}\n")
    return {"stat": "".join(out), "suffix": SUFFIX[mode], "site": site, "mode": mode}


def fmt_ratio_signed(cov: int, total: int, length: int) -> Tuple[str, str]:
    """Row ratios where GeneCover is a (possibly negative) int (:5023); not reachable at sane sizes."""
    if cov >= 0 and total >= 0:
        return fmt_ratio(cov, total, length)
    c = cov * 100.0 / length if length else float("nan")
    d = total * 1.0 / length if length else float("nan")
    return f"{c:.2f}", f"{d:.2f}"
