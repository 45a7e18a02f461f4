#!/usr/bin/env python
"""
bench.py -- BASELINE.json's metric on BASELINE.json's config.

Metric : Mreads/s depth-accumulated (and Gbases/s = 150 x reads) on a 3 Gb x 50x synthetic sorted read set,
         per-chr mode: accumulate (filter + CIGAR walk) -> per-base depth -> CoveredSite/TotalDepth per 10 Mb
         pseudo-gene -> per-chr rows + the ##RegionLength summary triple.
Step   : one full pass: pd_reset, pd_push_contig x contigs, pd_finalize, pd_reduce, read back the triples.
N GPUs : cfg5 -- ONE job (the same 3 Gb genome at 200x, 4 G reads) strong-scaled by genomic range (pandepth_b200/shard.py
         assign_ranges + pd_create_ranges: 9 contigs do not divide over 8 GPUs, equal ranges of the concatenated genome do);
         one process per GPU, boundary reads go to both owners and each engine clips; every rank reduces all intervals
         clipped to its range and ONE NCCL all-reduce of the (CoveredSite, TotalDepth) counters, inside the timed region,
         gives every rank the per-chr rows and the ##RegionLength summary triple.

value : reads/s with the decoded (pos, cigar) batches already resident in HBM (device pointers, PD_DEVICE) -- north_star's
        path: filter + CIGAR walk -> per-base depth -> CoveredSite/TotalDepth.
e2e   : the SAME JOB THE REFERENCE DOES, from the same starting point: the BAM file of this workload (page cache) ->
        per-chr triples, through the product feeder (pdh_feed_bam, include/pandepth_host.h: parallel pread into pinned
        pieces -> pd_push_bgzf: BGZF inflate + record walk + accumulate on the device) and pd_reduce.  Host->device copies
        (the compressed file) and the read-back of the triples are inside the timed region.  When the cpu_baseline leg runs,
        the triples are compared with the rows the reference binary printed for the same file.
        e2e_arrays is last round's number: decoded arrays in pinned host memory -> pd_push_contig (8 B/read over PCIe).
--impl reference : the unmodified reference binary (oracle/_ref/pandepth_ref) on the box's host cores on the same BAM file
        (generated once per box by oracle/_ref/bamgen -P and cached under /tmp).
"""
import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

READLEN = 150
N_CONTIGS = 9


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--genome", type=float, default=3.0e9, help="bases per GPU (cfg2: 3e9)")
    ap.add_argument("--coverage", type=float, default=50.0, help="N = 1 (cfg2)")
    ap.add_argument("--coverage-multi", type=float, default=200.0, help="N > 1 (cfg5): the same genome at this coverage, strong-scaled")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-e2e-arrays", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--bam-mb", type=float, default=0.0, help="genome size (Mb) of the BAM both arms read (default: --genome for flavour A, 300 for B)")
    ap.add_argument("--flavour", default="A", choices=["A", "B"], help="BAM flavour of SURVEY 8(d): A = SEQ '*', B = random SEQ+QUAL")
    ap.add_argument("--ref-threads", type=int, default=0)
    ap.add_argument("--feed-threads", type=int, default=0)
    ap.add_argument("--ref-sample-mb", type=float, default=0.0, help=argparse.SUPPRESS)       # older spelling of --bam-mb
    ap.add_argument("--ref-seq", action="store_true", help=argparse.SUPPRESS)                 # older spelling of --flavour B
    a = ap.parse_args()
    if a.ref_seq:
        a.flavour = "B"
    if a.ref_sample_mb and not a.bam_mb:
        a.bam_mb = a.ref_sample_mb
    if not a.bam_mb:
        a.bam_mb = a.genome / 1e6 if a.flavour == "A" else 300.0
    return a


def contig_lengths(genome):
    base = int(genome) // N_CONTIGS
    lens = [base] * N_CONTIGS
    lens[-1] += int(genome) - base * N_CONTIGS
    return lens


# ----------------------------------------------------------------------------------------------------------------
# clocks sampler (recipe line from /opt/skills/guides/B200_PROFILING.md)
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s in sm if s > 0.5 * max(sm)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the unmodified reference binary on a bounded BAM sample
# ----------------------------------------------------------------------------------------------------------------
def cpu_quota_cores():
    """CPUs the container may actually use (cgroup v2 cpu.max), or None when unlimited/unknown."""
    try:
        a, p = open("/sys/fs/cgroup/cpu.max").read().split()
        return None if a == "max" else float(a) / float(p)
    except Exception:
        return None


def ref_paths():
    d = os.path.join(ROOT, "oracle", "_ref")
    return os.path.join(d, "pandepth_ref"), os.path.join(d, "bamgen")


def usable_cores():
    q = cpu_quota_cores()
    n = os.cpu_count() or 1
    return max(1, int(q + 0.5)) if q and q < n else n


def make_sample_bam(sample_mb, coverage, with_seq, threads=0):
    """The BAM of the workload (9 contigs, sorted 150M reads, BGZF level 1, .bai), written once per box by bamgen -P."""
    ref, gen = ref_paths()
    if not (os.path.exists(ref) and os.path.exists(gen)):
        return None, None, "oracle/_ref binaries missing"
    lens = contig_lengths(sample_mb * 1e6)
    reads = int(sum(lens) * coverage / READLEN)
    tag = f"pdb200_bam_{int(sample_mb)}mb_{int(coverage)}x_{'B' if with_seq else 'A'}"
    d = os.path.join(tempfile.gettempdir(), tag)
    bam = os.path.join(d, "s.bam")
    if not os.path.exists(bam + ".ok"):
        os.makedirs(d, exist_ok=True)
        spec = ",".join(f"chr{i + 1}:{l}" for i, l in enumerate(lens))
        cmd = [gen, "-P", "-o", bam, "-c", spec, "-n", str(reads), "-s", "7", "-l", str(READLEN), "-t", str(max(2, min(32, threads or usable_cores())))]
        if with_seq:
            cmd.append("-q")
        subprocess.run(cmd, check=True)
        open(bam + ".ok", "w").write("ok\n")
    return bam, reads, None


def read_chr_stat(path):
    """Rows of a .chr.stat.gz the reference (or our CLI) printed: {chr: (Length, CoveredSite, TotalDepth)}."""
    import gzip
    rows = {}
    with gzip.open(path, "rt") as fh:
        for line in fh:
            if line.startswith("#"):
                continue
            f = line.rstrip("\n").split("\t")
            rows[f[0]] = (int(f[1]), int(f[2]), int(f[3]))
    return rows


def time_reference(bam, reads, threads, steps, warmup, keep_rows=False):
    ref, _ = ref_paths()
    out = tempfile.mkdtemp(prefix="pdb200_ref")
    times = []
    rows = None
    try:
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            subprocess.run([ref, "-i", bam, "-o", os.path.join(out, "o"), "-t", str(threads)], check=True,
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
        if keep_rows:
            rows = read_chr_stat(os.path.join(out, "o.chr.stat.gz"))
    finally:
        shutil.rmtree(out, ignore_errors=True)
    return (times, rows) if keep_rows else times


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = args.ref_threads or min(52, os.cpu_count() or 1)
    bam, reads, why = make_sample_bam(args.bam_mb, args.coverage, args.flavour == "B")
    if bam is None:
        print(json.dumps({"impl": "reference", "unavailable": why}))
        return
    times = time_reference(bam, reads, threads, args.steps, max(1, args.warmup))
    total = sum(times)
    value = reads * len(times) / total / 1e6
    sample = (f"{N_CONTIGS} contigs, {args.bam_mb:.0f} Mb x {args.coverage:.0f}x = {reads} sorted 150M reads, "
              f"BAM level 1 flavour {args.flavour} ({'random SEQ+QUAL' if args.flavour == 'B' else 'SEQ=*'}), .bai present, page cache warm; "
              f"the same file our e2e arm reads")
    line = {
        "impl": "reference", "metric": "Mreads/s depth-accumulated", "value": value, "unit": "Mreads/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": max(1, args.warmup), "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "gbases_per_s": value * READLEN / 1e3,
        "config": {"workload": ("cfg2 (9 contigs, 50x, sorted 150M reads, per-chr mode): " if args.gpus <= 1 else
                                "cfg5's genome (9 contigs, 3 Gb, per-chr mode) on a bounded sample -- 50x instead of 200x, a quarter of the reads; "
                                "the reference's Mreads/s does not depend on the coverage: ") + sample,
                   "cmd": f"pandepth_ref -i s.bam -o out -t {threads}"},
        "cpu_baseline": {"value": value, "unit": "Mreads/s", "cores": threads, "cpu_quota_cores": cpu_quota_cores(), "online_cpus": os.cpu_count(),
                         "kind": "reference", "sample": sample,
                         "note": "reference runs min(t, #contigs)=9 workers + 2 BGZF threads each (PanDepth.cpp:4447-4452,:4511); includes its own BAM decode"},
        "e2e": {"value": value, "unit": "Mreads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------
def tile_plan(lens, coverage, tile=1_000_000):
    """The workload's read set, defined tile by tile so that any rank can materialise exactly the reads of its range:
    tile k of contig t holds round(coverage * tile_len / READLEN) reads whose starts are uniform in the tile (seeded by the
    tile's global index), never running past the contig end."""
    plan = []
    gid = 0
    for t, ln in enumerate(lens):
        last_start = ln - READLEN                       # a read must fit the contig
        for lo in range(0, max(1, last_start + 1), tile):
            hi = min(lo + tile, last_start + 1)
            if hi > lo:
                plan.append((t, lo, hi, int(round(coverage * (hi - lo) / READLEN)), gid))
            gid += 1
    return plan


def rank_reads(plan, t, b, z, device, torch):
    """Sorted start positions of the reads of contig t that can touch [b, z): the tiles intersecting [b - READLEN + 1, z)."""
    gen = torch.Generator(device=device)
    parts = []
    for (tt, lo, hi, n, gid) in plan:
        if tt != t or hi + READLEN - 1 <= b or lo >= z or n == 0:
            continue
        gen.manual_seed(1_000_003 * 7 + gid)
        p = torch.randint(lo, hi, (n,), device=device, dtype=torch.int32, generator=gen)
        parts.append(torch.sort(p).values)
    return torch.cat(parts).contiguous() if parts else torch.zeros(0, dtype=torch.int32, device=device)


def ours(args):
    import ctypes
    import numpy as np
    import torch
    import torch.distributed as dist
    import pandepth_b200 as pd
    from pandepth_b200.shard import assign_ranges, allreduce_sum_u64

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # the job.  N = 1: cfg2 (3 Gb, 50x).  N > 1: cfg5 -- the SAME 3 Gb genome at 200x, strong-scaled by genomic range: rank r
    # owns an equal r-th part of the concatenated genome (pd_create_ranges), is fed the reads overlapping it (boundary reads
    # go to both neighbours, each engine clips), reduces ALL intervals clipped to its range, and one NCCL all-reduce of the
    # (CoveredSite, TotalDepth) arrays -- inside the timed region -- gives every rank the whole table.
    coverage = args.coverage if world == 1 else (args.coverage_multi or 200.0)
    lens = contig_lengths(args.genome)
    tlen = np.array(lens, np.uint32)
    own_beg, own_end = assign_ranges(lens, world)
    plan = tile_plan(lens, coverage)
    R_total = sum(p[3] for p in plan)
    G_total = int(sum(lens))
    mine = [t for t in range(len(lens)) if own_end[rank, t] > own_beg[rank, t]]
    G_local = int(sum(int(own_end[rank, t]) - int(own_beg[rank, t]) for t in mine))

    # a dedicated (non-default) torch stream is handed to the engine, so the CUDA events below are recorded on the very
    # stream every kernel and copy is launched on
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    eng = pd.Engine(tlen, device=local_rank, own_beg=own_beg[rank], own_end=own_end[rank])
    eng.set_stream(stream.cuda_stream)
    eng.set_reduce_hint(1, 0xFFFFFFFF)          # -d 1, uint32 path: known before the reads arrive, as in the CLI

    # synthetic sorted reads on the device: the tiles that can touch this rank's range, per contig
    d_pos, d_cig, n_reads = {}, {}, {}
    for t in mine:
        d_pos[t] = rank_reads(plan, t, int(own_beg[rank, t]), int(own_end[rank, t]), dev, torch)
        n_reads[t] = int(d_pos[t].numel())
        d_cig[t] = torch.full((n_reads[t],), (READLEN << 4) | 0, device=dev, dtype=torch.int32)
    R_local = sum(n_reads.values())             # includes the boundary tiles' reads that neighbours hold too
    torch.cuda.synchronize()

    # per-chr mode intervals: 10 Mb pseudo-genes with the reference's tiling rule (PanDepth.cpp:3997-4046), every contig
    iv_t, iv_s, iv_e = [], [], []
    W = 10_000_000
    for t in range(len(lens)):
        Start, End = 1, 2
        while End <= lens[t]:
            End = min(Start + W - 1, lens[t])
            iv_t.append(t); iv_s.append(Start - 1); iv_e.append(End)
            End += 2; Start += W
    iv_t = np.array(iv_t, np.int32); iv_s = np.array(iv_s, np.int32); iv_e = np.array(iv_e, np.int32)

    def combine(cov, tot):
        if world == 1:
            return cov, tot
        both = allreduce_sum_u64(np.stack([cov, tot]), world, device=dev)      # NCCL all-reduce (sum) of the triples' counters
        return both[0], both[1]

    phase = {"reset": 0.0, "push": 0.0, "reduce": 0.0, "combine": 0.0, "n": 0}      # host-side wall time per phase of the device step

    def step_device():
        t0 = time.perf_counter()
        eng.reset()
        t1 = time.perf_counter()
        for t in mine:
            if n_reads[t]:
                # (PD_READY -- cut kernels of batch k+1 on a side stream under batch k's tile kernel -- was measured on this
                #  step: 5.10 -> 5.06 ms, because the tile kernel itself runs 10 % slower next to them; not used here)
                eng.push_contig(t, d_pos[t].data_ptr(), d_cig[t].data_ptr(), flags=pd.PD_SORTED | pd.PD_DEVICE, n=n_reads[t])
        t2 = time.perf_counter()
        cov, tot = eng.reduce(iv_t, iv_s, iv_e, 1, 0xFFFFFFFF)       # finalizes, reduces, copies the result back (synchronises)
        t3 = time.perf_counter()
        res = combine(cov, tot)
        t4 = time.perf_counter()
        phase["reset"] += t1 - t0; phase["push"] += t2 - t1; phase["reduce"] += t3 - t2; phase["combine"] += t4 - t3; phase["n"] += 1
        return res

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record(stream)
        res = None
        for _ in range(steps):
            res = fn()
        e1.record(stream)
        barrier()
        wall = (time.perf_counter() - w0) * 1e3
        ms = max(e0.elapsed_time(e1), 0.0)
        if world > 1:
            t = torch.tensor([ms, wall], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0].item()), float(t[1].item())
        return ms, wall, res

    sampler = ClockSampler(local_rank) if rank == 0 else None
    # ---- device-resident arm (value) with per-kernel timing --------------------------------------------------
    if sampler:
        sampler.start()
        time.sleep(0.5)              # nvidia-smi needs a moment before its first sample
    for _ in range(args.warmup):
        step_device()
    eng.profile_enable(True)
    for k in phase:
        phase[k] = 0
    ms_dev, _, (cov, tot) = timed(step_device, args.steps, 0)
    host_phase_ms = {k: 1e3 * v / max(1, phase["n"]) for k, v in phase.items() if k != "n"}
    clocks = sampler.stop() if sampler else None
    prof = eng.profile()
    eng.profile_enable(False)
    launches_per_step = eng.stats()["kernel_launches"]

    # property check at full size: every aligned base of the whole job counted exactly once, on every rank's copy of the table
    assert int(tot.sum()) == READLEN * R_total, (int(tot.sum()), READLEN * R_total)
    table = np.zeros((len(lens), 3), np.uint64)
    for i in range(iv_t.size):
        table[iv_t[i], 0] += np.uint64(iv_e[i] - iv_s[i]); table[iv_t[i], 1] += cov[i]; table[iv_t[i], 2] += tot[i]

    # ---- end-to-end arm: the BAM file of the workload -> triples, through the product feeder ---------------------
    e2e = None
    e2e_arrays = None
    ref_rows = None
    cpu_baseline = None
    feed_threads = args.feed_threads or max(1, usable_cores() // world)
    if not args.no_e2e:
        bam_mb = args.bam_mb if world == 1 else args.genome / 1e6
        if rank == 0:
            bam, bam_reads, why = make_sample_bam(bam_mb, coverage, args.flavour == "B")
        barrier()
        if rank != 0:
            bam, bam_reads, why = make_sample_bam(bam_mb, coverage, args.flavour == "B")     # cached by rank 0
        if bam is None:
            e2e = {"value": None, "unit": "Mreads/s", "unavailable": why}
        else:
            H = ctypes.CDLL(os.path.join(ROOT, "pandepth_b200", "libpandepth_host.so"))

            class FS(ctypes.Structure):
                _fields_ = [(n, ctypes.c_uint64) for n in ("file_bytes", "pieces", "bgzf_blocks", "reads")]

            H.pdh_feed_bam_range.restype = ctypes.c_int
            H.pdh_feed_bam_range.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t]
            H.pdh_bai_block_offset.restype = ctypes.c_uint64
            H.pdh_bai_block_offset.argtypes = [ctypes.c_char_p, ctypes.c_int32, ctypes.c_int64]
            # the BAM's own genome (flavour B runs on a smaller genome than the value arm): an engine for ITS contigs
            blens = contig_lengths(bam_mb * 1e6)
            btlen = np.array(blens, np.uint32)
            bbeg, bend = assign_ranges(blens, world)
            same_genome = blens == lens
            beng = eng if same_genome else pd.Engine(btlen, device=local_rank, own_beg=bbeg[rank], own_end=bend[rank])
            if not same_genome:
                beng.set_stream(stream.cuda_stream); beng.set_reduce_hint(1, 0xFFFFFFFF)
            b_t, b_s, b_e = [], [], []
            for t in range(len(blens)):
                Start, End = 1, 2
                while End <= blens[t]:
                    End = min(Start + W - 1, blens[t])
                    b_t.append(t); b_s.append(Start - 1); b_e.append(End)
                    End += 2; Start += W
            b_t = np.array(b_t, np.int32); b_s = np.array(b_s, np.int32); b_e = np.array(b_e, np.int32)
            # this rank's byte range of the file, from the .bai: first block that can hold a read overlapping its first base
            # ... up to the block where the next rank's range starts (that block goes to both; the engines clip)
            f_beg, f_end = 0, 2 ** 64 - 1
            if world > 1:
                mine_b = [t for t in range(len(blens)) if bend[rank, t] > bbeg[rank, t]]
                t0, t1 = mine_b[0], mine_b[-1]
                if rank > 0:
                    f_beg = int(H.pdh_bai_block_offset(bam.encode(), t0, max(0, int(bbeg[rank, t0]) - READLEN)))
                if rank < world - 1:
                    # the first read overlapping the 16 kb window AFTER the one my range ends in starts behind my last base:
                    # everything I need lies before its block (that block is fed too; the engines clip)
                    nxt = int(H.pdh_bai_block_offset(bam.encode(), t1, int(bend[rank, t1]) + 16384))
                    f_end = nxt + 1 if nxt < 2 ** 64 - 1 else f_end
            err = ctypes.create_string_buffer(512)
            fs = FS()

            def step_bam():
                beng.reset()
                rc = H.pdh_feed_bam_range(beng._h, bam.encode(), feed_threads, f_beg, f_end, ctypes.byref(fs), err, 512)
                if rc != 0:
                    raise RuntimeError(f"pdh_feed_bam_range rc={rc}: {err.value.decode()}")
                c, s_ = beng.reduce(b_t, b_s, b_e, 1, 0xFFFFFFFF)
                return combine(c, s_)

            k = max(1, min(args.steps, 3))
            step_bam()                       # warm-up: pinned pieces, device slots, page cache
            beng.reset(); s0 = beng.stats()
            ms_e2e, wall_e2e, (covb, totb) = timed(step_bam, k, 0)
            s1 = beng.stats()
            assert int(totb.sum()) == READLEN * bam_reads, (int(totb.sum()), READLEN * bam_reads)
            h2d = torch.tensor([float(s1["h2d_bytes"]), float(s1["d2h_bytes"])], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(h2d, op=dist.ReduceOp.SUM)
            btable = np.zeros((len(blens), 3), np.uint64)
            for i in range(b_t.size):
                btable[b_t[i], 0] += np.uint64(b_e[i] - b_s[i]); btable[b_t[i], 1] += covb[i]; btable[b_t[i], 2] += totb[i]
            e2e = {"value": bam_reads * k / (wall_e2e * 1e-3) / 1e6, "unit": "Mreads/s",
                   "h2d_bytes_per_step": int(h2d[0].item()), "d2h_bytes_per_step": int(h2d[1].item()), "steps": k,
                   "ms_per_step": wall_e2e / k, "ms_per_step_cuda_events": ms_e2e / k,
                   "timing": "wall clock between barriers (host work is part of the job), max over ranks; CUDA-event time beside it",
                   "input": f"BAM file in the page cache: {os.path.basename(os.path.dirname(bam))}/s.bam, {os.path.getsize(bam)} bytes, {bam_reads} reads, "
                            f"flavour {args.flavour}, genome {bam_mb:.0f} Mb x {coverage:.0f}x",
                   "path": f"pdh_feed_bam_range ({feed_threads} reader threads/rank, pinned pieces) -> pd_push_bgzf (device inflate + record walk + accumulate) -> pd_reduce"
                           + (" -> NCCL all-reduce" if world > 1 else ""),
                   "gbases_per_s": bam_reads * k / (wall_e2e * 1e-3) / 1e6 * READLEN / 1e3,
                   "summary_row": {"RegionLength": int(btable[:, 0].sum()), "CoveredSite": int(btable[:, 1].sum()), "TotalDepth": int(btable[:, 2].sum())}}
            # the reference on the same file (rank 0, N = 1): baseline figure AND a full-size parity check of the rows it printed
            if rank == 0 and world == 1 and not args.no_cpu_baseline:
                threads = args.ref_threads or min(52, os.cpu_count() or 1)
                times, ref_rows = time_reference(bam, bam_reads, threads, 2, 1, keep_rows=True)
                v = bam_reads / min(times) / 1e6
                cpu_baseline = {"value": v, "unit": "Mreads/s", "cores": threads, "cpu_quota_cores": cpu_quota_cores(), "online_cpus": os.cpu_count(),
                                "kind": "reference",
                                "sample": f"the e2e arm's own file: {N_CONTIGS} contigs, {bam_mb:.0f} Mb x {coverage:.0f}x = {bam_reads} reads, flavour {args.flavour}, "
                                          f"best of 2 after one warm-up, pandepth_ref -t {threads} (whole process, decode included)"}
                same = all(ref_rows.get(f"chr{t + 1}") == (int(btable[t, 0]), int(btable[t, 1]), int(btable[t, 2])) for t in range(len(blens)))
                e2e["parity_vs_reference_rows"] = "identical" if same else "DIFFERENT"
                assert same, "e2e triples differ from the rows the reference binary printed for the same BAM"
            if not same_genome:
                beng.close()

        # last round's e2e: decoded arrays in pinned host memory -> pd_push_contig (kept as a second figure)
        if world == 1 and not args.no_e2e_arrays:
            h_pos, h_cig = {}, {}
            for t in mine:
                h_pos[t] = torch.empty(n_reads[t], dtype=torch.int32, pin_memory=True); h_pos[t].copy_(d_pos[t])
                h_cig[t] = torch.empty(n_reads[t], dtype=torch.int32, pin_memory=True); h_cig[t].copy_(d_cig[t])
            torch.cuda.synchronize()

            def step_host():
                eng.reset()
                for t in mine:
                    eng.push_contig(t, h_pos[t].numpy(), h_cig[t].numpy(), flags=pd.PD_SORTED | pd.PD_ASYNC)
                return eng.reduce(iv_t, iv_s, iv_e, 1, 0xFFFFFFFF)

            k = max(1, min(args.steps, 3))
            step_host()
            eng.reset()
            ms_a, _, (cov2, tot2) = timed(step_host, k, 0)
            s1 = eng.stats()
            assert np.array_equal(cov2, cov) and np.array_equal(tot2, tot)
            e2e_arrays = {"value": R_total * k / (ms_a * 1e-3) / 1e6, "unit": "Mreads/s", "h2d_bytes_per_step": int(s1["h2d_bytes"]),
                          "d2h_bytes_per_step": int(s1["d2h_bytes"]), "steps": k, "ms_per_step": ms_a / k,
                          "host_buffers": "decoded (pos, cigar) arrays, pinned (torch pin_memory), pd_push_contig PD_SORTED|PD_ASYNC"}
            del h_pos, h_cig

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (of fallback)"
        kt = prof.get("k_tile_accumulate", {"ms": 0.0, "launches": 0})
        alg_bytes_per_step = 8 * R_local + 4 * G_local                  # K1 = 8 B/read + 4 B/base (this rank's launches)
        n_launch = max(1, kt["launches"])
        avg_ms = kt["ms"] / n_launch
        alg_per_launch = alg_bytes_per_step * args.steps / n_launch
        achieved = alg_per_launch / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
        value = R_total * args.steps / (ms_dev * 1e-3) / 1e6
        # dram traffic of one k_tile_accumulate launch from the committed ncu --set full capture (same workload only)
        traffic = None
        traffic_src = None
        try:
            import glob
            for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json"))):
                tj = json.load(open(f))["kernels"].get("k_tile_accumulate")
                if tj and args.genome == 3.0e9 and coverage == 50.0 and world == 1:
                    traffic = tj["dram_bytes_per_launch"]
                    traffic_src = f"profiles/{os.path.basename(f)}: dram__bytes_read.sum + dram__bytes_write.sum of one launch, committed ncu --set full capture of this workload (not measured by this run)"
        except Exception:
            traffic = None
        cfg = "cfg2" if world == 1 else "cfg5"
        line = {
            "metric": "Mreads/s depth-accumulated", "value": value, "unit": "Mreads/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak" if world == 1 else "strong",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "gbases_per_s": value * READLEN / 1e3,
            "config": {"workload": f"{cfg}: {N_CONTIGS} contigs summing {args.genome / 1e9:.2f} Gb, {coverage:.0f}x = {R_total} sorted 150M reads, "
                                   f"per-chr mode (10 Mb pseudo-genes)" + (f"; ONE job strong-scaled over {world} GPUs by genomic range" if world > 1 else ""),
                       "parallelism": (f"range-shard x{world} (equal cuts of the concatenated genome, boundary reads to both owners, engines clip), "
                                       f"NCCL all-reduce of the per-interval (CoveredSite, TotalDepth) counters inside the timed region") if world > 1 else "1 GPU",
                       "l2": "inputs (8 B/read) and the depth array (4 B/base) are far larger than the 126 MB L2; no flush needed",
                       "timing": "CUDA events on the engine's stream, max over ranks"},
            "roofline": {"bound": "hbm", "kernel": "k_tile_accumulate", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if peak else None, "traffic": traffic, "traffic_source": traffic_src,
                         "algorithmic_bytes_per_launch": alg_per_launch, "avg_launch_ms": avg_ms, "launches": kt["launches"],
                         "peak_source": peak_src,
                         "kernel_share_of_step": kt["ms"] / ms_dev if ms_dev else None},
            "kernels_ms_per_step": {k: v["ms"] / args.steps for k, v in prof.items() if v["launches"]},
            "host_phase_ms_per_step": host_phase_ms,
            "gpu_launches": int(launches_per_step),
            "clocks": clocks,
            "summary_row": {"RegionLength": int(table[:, 0].sum()), "CoveredSite": int(table[:, 1].sum()), "TotalDepth": int(table[:, 2].sum())},
        }
        if e2e:
            line["e2e"] = e2e
        if e2e_arrays:
            line["e2e_arrays"] = e2e_arrays
        if cpu_baseline:
            line["cpu_baseline"] = cpu_baseline
        elif world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = {"value": None, "unit": "Mreads/s", "cores": 0, "kind": "reference", "sample": "e2e arm disabled: no BAM to run the reference on"}
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
