#!/usr/bin/env python
"""
bench.py -- BASELINE.json's metric on BASELINE.json's config.

Metric : Mreads/s depth-accumulated (and Gbases/s = 150 x reads) on a 3 Gb x 50x synthetic sorted read set,
         per-chr mode: accumulate (filter + CIGAR walk) -> per-base depth -> CoveredSite/TotalDepth per 10 Mb
         pseudo-gene -> per-chr rows + the ##RegionLength summary triple.
Step   : one full pass: pd_reset, pd_push_contig x contigs, pd_finalize, pd_reduce, read back the triples.
N GPUs : weak scaling by contig sharding (one process per GPU, no data-path collective): the job is N assemblies of the
         cfg2 genome (9N contigs, 3N Gb, 50x -- a pan-genome of N haplotypes), contigs assigned by LPT bin packing
         (pandepth_b200/shard.py), one all-gather of the per-contig (Length, CoveredSite, TotalDepth) triples at the end.

value : reads/s with the batches already resident in HBM (device pointers, PD_DEVICE).
e2e   : the same pass through the same C-ABI calls with HOST buffers (pinned), host->device copies inside the timed region.
--impl reference : the unmodified reference binary (oracle/_ref/pandepth_ref) on the box's host cores, on a bounded
         BAM sample of the same workload (same 9-contig shape, 50x, 150M reads, scaled-down genome).
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
    ap.add_argument("--coverage", type=float, default=50.0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-sample-mb", type=float, default=300.0, help="genome size (Mb) of the reference arm's BAM sample")
    ap.add_argument("--ref-threads", type=int, default=0)
    ap.add_argument("--ref-seq", action="store_true", help="reference arm: BAM sample with random SEQ+QUAL (flavour B)")
    return ap.parse_args()


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


def make_sample_bam(sample_mb, coverage, with_seq, threads):
    ref, gen = ref_paths()
    if not (os.path.exists(ref) and os.path.exists(gen)):
        return None, None, "oracle/_ref binaries missing"
    lens = contig_lengths(sample_mb * 1e6)
    reads = int(sum(lens) * coverage / READLEN)
    tag = f"pdb200_sample_{int(sample_mb)}mb_{int(coverage)}x_{'B' if with_seq else 'A'}"
    d = os.path.join(tempfile.gettempdir(), tag)
    bam = os.path.join(d, "sample.bam")
    if not os.path.exists(bam + ".bai"):
        os.makedirs(d, exist_ok=True)
        spec = ",".join(f"chr{i + 1}:{l}" for i, l in enumerate(lens))
        cmd = [gen, "-o", bam, "-c", spec, "-n", str(reads), "-s", "7", "-l", str(READLEN), "-t", str(max(2, min(16, threads)))]
        if with_seq:
            cmd.append("-q")
        subprocess.run(cmd, check=True)
    return bam, reads, None


def time_reference(bam, reads, threads, steps, warmup):
    ref, _ = ref_paths()
    out = tempfile.mkdtemp(prefix="pdb200_ref")
    times = []
    try:
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            subprocess.run([ref, "-i", bam, "-o", os.path.join(out, "o"), "-t", str(threads)], check=True,
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    finally:
        shutil.rmtree(out, ignore_errors=True)
    return times


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = args.ref_threads or min(52, os.cpu_count() or 1)
    bam, reads, why = make_sample_bam(args.ref_sample_mb, args.coverage, args.ref_seq, threads)
    if bam is None:
        print(json.dumps({"impl": "reference", "unavailable": why}))
        return
    times = time_reference(bam, reads, threads, args.steps, max(1, args.warmup))
    total = sum(times)
    value = reads * len(times) / total / 1e6
    sample = (f"{N_CONTIGS} contigs, {args.ref_sample_mb:.0f} Mb x {args.coverage:.0f}x = {reads} sorted 150M reads, "
              f"BAM level 1 {'with SEQ+QUAL' if args.ref_seq else 'SEQ=*'}, .bai present, page cache warm")
    line = {
        "impl": "reference", "metric": "Mreads/s depth-accumulated", "value": value, "unit": "Mreads/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": max(1, args.warmup), "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "gbases_per_s": value * READLEN / 1e3,
        "config": {"workload": "cfg2 shape (9 contigs, 50x, sorted 150M reads, per-chr mode), bounded sample: " + sample,
                   "cmd": f"pandepth_ref -i sample.bam -o out -t {threads}"},
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
def ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import pandepth_b200 as pd
    from pandepth_b200.shard import assign_contigs, gather_triples, owned_mask

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # the job: `world` assemblies of the cfg2 genome; LPT gives every rank one assembly's worth of contigs
    per_asm = contig_lengths(args.genome)
    lens = per_asm * world
    owner = assign_contigs(lens, world)
    mine = [t for t in range(len(lens)) if owner[t] == rank]
    tlen = np.array(lens, np.uint32)
    G_local = sum(lens[t] for t in mine)
    reads_per_contig = {t: int(lens[t] * args.coverage / READLEN) for t in mine}
    R_local = sum(reads_per_contig.values())

    # a dedicated (non-default) torch stream is handed to the engine, so the CUDA events below are recorded on the very
    # stream every kernel and copy is launched on
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    eng = pd.Engine(tlen, device=local_rank, owned=owned_mask(owner, rank))
    eng.set_stream(stream.cuda_stream)
    eng.set_reduce_hint(1, 0xFFFFFFFF)          # -d 1, uint32 path: known before the reads arrive, as in the CLI

    # synthetic sorted reads, generated on the device (seeded), one (pos, cigar) pair of arrays per owned contig
    gen = torch.Generator(device=dev); gen.manual_seed(7 + rank)
    d_pos, d_cig = {}, {}
    for t in mine:
        n = reads_per_contig[t]
        p = torch.randint(0, lens[t] - READLEN + 1, (n,), device=dev, dtype=torch.int32, generator=gen)
        d_pos[t] = torch.sort(p).values.contiguous()
        d_cig[t] = torch.full((n,), (READLEN << 4) | 0, device=dev, dtype=torch.int32)
        del p
    torch.cuda.synchronize()

    # per-chr mode intervals: 10 Mb pseudo-genes with the reference's tiling rule (PanDepth.cpp:3997-4046)
    iv_t, iv_s, iv_e = [], [], []
    W = 10_000_000
    for t in mine:
        Start, End = 1, 2
        while End <= lens[t]:
            End = min(Start + W - 1, lens[t])
            iv_t.append(t); iv_s.append(Start - 1); iv_e.append(End)
            End += 2; Start += W
    iv_t = np.array(iv_t, np.int32); iv_s = np.array(iv_s, np.int32); iv_e = np.array(iv_e, np.int32)
    region_len = int((iv_e - iv_s).sum())

    def step_device():
        eng.reset()
        for t in mine:
            eng.push_contig(t, d_pos[t].data_ptr(), d_cig[t].data_ptr(), flags=pd.PD_SORTED | pd.PD_DEVICE, n=reads_per_contig[t])
        cov, tot = eng.reduce(iv_t, iv_s, iv_e, 1, 0xFFFFFFFF)       # finalizes, reduces, copies the result back
        return cov, tot

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        res = None
        for _ in range(steps):
            res = fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, res

    sampler = ClockSampler(local_rank) if rank == 0 else None
    # ---- device-resident arm (value) with per-kernel timing --------------------------------------------------
    if sampler:
        sampler.start()
        time.sleep(0.5)              # nvidia-smi needs a moment before its first sample
    for _ in range(args.warmup):
        step_device()
    eng.profile_enable(True)
    ms_dev, (cov, tot) = timed(step_device, args.steps, 0)
    clocks = sampler.stop() if sampler else None
    prof = eng.profile()
    eng.profile_enable(False)
    launches_per_step = eng.stats()["kernel_launches"]

    # property check at full size: every aligned base counted exactly once, all bases of the tiles seen
    assert int(tot.sum()) == READLEN * R_local, (int(tot.sum()), READLEN * R_local)
    local = np.zeros((len(lens), 3), np.uint64)
    for i in range(iv_t.size):
        local[iv_t[i], 0] += np.uint64(iv_e[i] - iv_s[i]); local[iv_t[i], 1] += cov[i]; local[iv_t[i], 2] += tot[i]
    table = gather_triples(local, owner, rank, world, device=dev)       # NCCL all-gather of the summary triples

    # ---- end-to-end arm: same calls, HOST (pinned) buffers ---------------------------------------------------
    e2e = None
    if not args.no_e2e:
        h_pos, h_cig = {}, {}
        for t in mine:
            h_pos[t] = torch.empty(reads_per_contig[t], dtype=torch.int32, pin_memory=True); h_pos[t].copy_(d_pos[t])
            h_cig[t] = torch.empty(reads_per_contig[t], dtype=torch.int32, pin_memory=True); h_cig[t].copy_(d_cig[t])
        torch.cuda.synchronize()

        def step_host():
            eng.reset()
            for t in mine:
                eng.push_contig(t, h_pos[t].numpy(), h_cig[t].numpy(), flags=pd.PD_SORTED | pd.PD_ASYNC)
            return eng.reduce(iv_t, iv_s, iv_e, 1, 0xFFFFFFFF)

        k = max(1, min(args.steps, 3))
        step_host()
        eng.reset(); s0 = eng.stats()
        ms_e2e, (cov2, tot2) = timed(step_host, k, 0)
        s1 = eng.stats()
        assert np.array_equal(cov2, cov) and np.array_equal(tot2, tot)
        e2e = {"value": R_local * world * k / (ms_e2e * 1e-3) / 1e6, "unit": "Mreads/s",
               "h2d_bytes_per_step": int(s1["h2d_bytes"]), "d2h_bytes_per_step": int(s1["d2h_bytes"]), "steps": k,
               "ms_per_step": ms_e2e / k, "host_buffers": "pinned (torch pin_memory), pd_push_contig PD_SORTED|PD_ASYNC"}
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
        alg_bytes_per_step = 8 * R_local + 4 * G_local                  # K1 = 8 B/read + 4 B/base
        n_launch = max(1, kt["launches"])
        avg_ms = kt["ms"] / n_launch
        alg_per_launch = alg_bytes_per_step * args.steps / n_launch
        achieved = alg_per_launch / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
        value = R_local * world * args.steps / (ms_dev * 1e-3) / 1e6
        # dram traffic of one k_tile_accumulate launch from the committed ncu --set full capture (same workload only)
        traffic = None
        try:
            import glob
            for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json"))):
                tj = json.load(open(f))["kernels"].get("k_tile_accumulate")
                if tj and args.genome == 3.0e9 and args.coverage == 50.0:
                    traffic = tj["dram_bytes_per_launch"]
        except Exception:
            traffic = None
        line = {
            "metric": "Mreads/s depth-accumulated", "value": value, "unit": "Mreads/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "gbases_per_s": value * READLEN / 1e3,
            "config": {"workload": f"cfg2 per GPU: {N_CONTIGS} contigs summing {args.genome / 1e9:.2f} Gb, {args.coverage:.0f}x = "
                                   f"{R_local} sorted 150M reads, per-chr mode (10 Mb pseudo-genes); x{world} assemblies sharded by contig",
                       "parallelism": f"contig-shard x{world}, NCCL all-gather of per-contig triples" if world > 1 else "1 GPU",
                       "l2": "inputs (8 GB/GPU) and the depth array (12 GB/GPU) are far larger than the 126 MB L2; no flush needed",
                       "timing": "CUDA events on the engine's stream, max over ranks"},
            "roofline": {"bound": "hbm", "kernel": "k_tile_accumulate", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if peak else None, "traffic": traffic,
                         "algorithmic_bytes_per_launch": alg_per_launch, "avg_launch_ms": avg_ms, "launches": kt["launches"],
                         "peak_source": peak_src,
                         "kernel_share_of_step": kt["ms"] / ms_dev if ms_dev else None},
            "kernels_ms_per_step": {k: v["ms"] / args.steps for k, v in prof.items() if v["launches"]},
            "gpu_launches": int(launches_per_step),
            "clocks": clocks,
            "summary_row": {"RegionLength": int(table[:, 0].sum()), "CoveredSite": int(table[:, 1].sum()), "TotalDepth": int(table[:, 2].sum())},
        }
        if e2e:
            line["e2e"] = e2e
        if world == 1 and not args.no_cpu_baseline:
            threads = args.ref_threads or min(52, os.cpu_count() or 1)
            bam, reads, why = make_sample_bam(args.ref_sample_mb, args.coverage, args.ref_seq, threads)
            if bam:
                times = time_reference(bam, reads, threads, 3, 1)
                v = reads / min(times) / 1e6
                line["cpu_baseline"] = {"value": v, "unit": "Mreads/s", "cores": threads, "cpu_quota_cores": cpu_quota_cores(),
                                        "online_cpus": os.cpu_count(), "kind": "reference",
                                        "sample": f"{N_CONTIGS} contigs, {args.ref_sample_mb:.0f} Mb x {args.coverage:.0f}x = {reads} reads, "
                                                  f"BAM level 1 {'SEQ+QUAL' if args.ref_seq else 'SEQ=*'}, best of 3, pandepth_ref -t {threads}"}
            else:
                line["cpu_baseline"] = {"value": None, "unit": "Mreads/s", "cores": 0, "kind": "reference", "sample": why}
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
