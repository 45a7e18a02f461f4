#!/usr/bin/env python
"""gpurun_out/{launches.csv, prof_rNN_*.ncu-rep} -> profiles/rNN_* summaries (+ the dram traffic of the dominant kernel as JSON).
usage: python tools/summarize_profiles.py r01"""
import collections, csv, json, os, shutil, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles"); os.makedirs(P, exist_ok=True)

# ---- launch lists --------------------------------------------------------------------------------------------
LISTS = [("launches.csv", "launches", "python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline (cfg2: 3 Gb x 50x), the engine's kernels only (-k regex: torch launches ~2x10^4 kernels to generate the input)"),
         ("launches_bgzf_A.csv", "launches_bgzf_A", "python tools/bench_bgzf.py cA.bam --steps 1 --no-feed (flavour A: 100 M SEQ-less reads, compressed file resident in HBM)"),
         ("launches_bgzf_B.csv", "launches_bgzf_B", "python tools/bench_bgzf.py cB.bam --steps 1 --no-feed (flavour B: 20 M reads with SEQ+QUAL, compressed file resident in HBM)")]
for fname, out, what in LISTS:
    lc = os.path.join(G, fname)
    if not os.path.exists(lc):
        continue
    rows = list(csv.reader(open(lc)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]; kn = hdr.index("Kernel Name"); mv = hdr.index("Metric Value"); mu = hdr.index("Metric Unit")
    agg = collections.OrderedDict(); tot = 0.0
    for r in rows[hi + 1:]:
        if len(r) <= mv: continue
        name = r[kn].split("(")[0].replace("void ", "")
        v = float(r[mv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[mu], 1.0)
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
    ours = {k: v for k, v in agg.items() if k.startswith("pdk::") or k.startswith("k_")}; tot_o = sum(v[1] for v in ours.values()) or 1.0
    L = [f"# ncu launch list, {tag}: {what}",
         "# ncu --metrics gpu__time_duration.sum --clock-control none -c N --csv ...",
         "# per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.",
         "# torch kernels (radix sort, fill, random, copies) generate / stage the synthetic input and run OUTSIDE the timed region.", "",
         f"{'kernel':72s} {'launches':>8s} {'total_us':>12s} {'share':>7s} {'engine share':>13s}"]
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        es = f"{100 * us / tot_o:12.1f}%" if k in ours else ""
        L.append(f"{k[:72]:72s} {n:8d} {us:12.1f} {100 * us / tot:6.1f}% {es}")
    open(os.path.join(P, f"{tag}_{out}_summary.txt"), "w").write("\n".join(L) + "\n")
    shutil.copy(lc, os.path.join(P, f"{tag}_{out}.csv"))
    print("\n".join(L))

# ---- full captures -------------------------------------------------------------------------------------------
KEYS = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of ncu peak"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm throughput %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("launch__registers_per_thread", "registers/thread"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__shared_mem_per_block_dynamic", "dyn smem/block"), ("smsp__inst_executed.sum", "warp instructions"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
        ("smsp__inst_executed_op_shared_atom.sum", "shared atomics (warp instr)"),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "LSU data pipe %"),
        ("smsp__average_warp_latency_per_inst_issued.ratio", "cycles per issued instruction and warp"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard / issue"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction / issue"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (expected 0)")]
UB = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}; US = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1, "second": 1}
T = [f"# ncu --set full --clock-control none --import-source on, {tag}, B200 (sm_100a); one launch per kernel, values per launch.",
     "# tile/reduce/cut: bench.py cfg2 (3 Gb x 50x): one launch = one 333,333,333-base contig, 111,111,111 sorted 150M reads",
     "# tile_mixed:      tools/bench_variants.py --only mixed (same contig, 30 % multi-op CIGARs with cigar_off[])",
     "# bgzf_A / bgzf_B: tools/bench_bgzf.py on 100 M SEQ-less reads / 20 M reads with SEQ+QUAL: one launch = one chunk of BGZF blocks",
     "# scan/scatter:    tools/bench_delta_path.py 3e8 (300 Mb contig, 100 M unsorted reads)", ""]
traffic = {}
for rep in sorted(f for f in os.listdir(G) if f.startswith(f"prof_{tag}_") and f.endswith(".ncu-rep")):
    out = subprocess.run(["ncu", "-i", os.path.join(G, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3: continue
    hdr, units = rows[0], rows[1]; kn = hdr.index("Kernel Name"); seen = set()
    for r in rows[2:]:
        if r[kn] in seen: continue
        seen.add(r[kn]); T.append(f"== {r[kn]}   [{rep}]"); vals = {}
        for k, label in KEYS:
            if k in hdr:
                i = hdr.index(k); T.append(f"   {label:34s} {r[i]:>18s} {units[i]}"); vals[k] = (float(r[i].replace(',', '')), units[i])
        tr = sum(vals[k][0] * UB[vals[k][1]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        t = vals["gpu__time_duration.sum"][0] * US[vals["gpu__time_duration.sum"][1]]
        T.append(f"   {'dram traffic (read+write)':34s} {tr / 1e6:18.1f} MB   -> {tr / t / 1e9:8.1f} GB/s under ncu"); T.append("")
        short = r[kn].split("(")[0].replace("void ", "").split("<")[0]
        key = short if short not in traffic else short + ":" + rep[len(f"prof_{tag}_"):-len(".ncu-rep")]     # e.g. k_tile_accumulate:tile_mixed
        traffic[key] = {"dram_bytes_per_launch": tr, "duration_s_under_ncu": t, "report": rep}
open(os.path.join(P, f"{tag}_ncu_full_summary.txt"), "w").write("\n".join(T) + "\n")
json.dump({"note": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full; cfg2 launch = one 333,333,333-base contig / 111,111,111 reads",
           "kernels": traffic}, open(os.path.join(P, f"{tag}_traffic.json"), "w"), indent=1)
print("\n".join(T[:4])); print({k: round(v["dram_bytes_per_launch"] / 1e6, 1) for k, v in traffic.items()})
