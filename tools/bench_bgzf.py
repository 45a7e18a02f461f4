#!/usr/bin/env python
"""Device-side BGZF decode bench (GPU box): one JSON line per BAM.

  python tools/bench_bgzf.py file.bam [--steps K] [--threads T] [--piece-mb M]

  resident : the whole compressed file resident in HBM, pushed in pieces with PD_DEVICE|PD_ASYNC -> per-kernel device times
             (k_bgzf_decode, k_bgzf_pack, k_cut, k_tile_accumulate, ...) and Mreads/s of decode + accumulate
  feed     : pdh_feed_bam (include/pandepth_host.h): file in the page cache -> pinned pieces -> pd_push_bgzf, wall clock
Both are checked against each other and against TotalDepth = sum of aligned bases where the generator's shape is known.
"""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("bam")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--threads", type=int, default=16)
    ap.add_argument("--piece-mb", type=float, default=0.0)
    ap.add_argument("--no-resident", action="store_true")
    ap.add_argument("--no-feed", action="store_true")
    args = ap.parse_args()
    import torch
    import pandepth_b200 as pd
    from tests import bgzf_util as bu

    data = np.fromfile(args.bam, dtype=np.uint8)
    t0 = time.perf_counter()
    # block table + header (python hop over the headers: fine up to a few hundred thousand blocks)
    tab = bu.block_table(data)
    import struct, zlib
    head = b""
    b0 = 0
    while True:
        o, l, n = int(tab["in_off"][b0]), int(tab["in_len"][b0]), int(tab["out_len"][b0])
        head += zlib.decompress(data[o:o + l].tobytes(), -15)
        b0 += 1
        if len(head) < 12:
            continue
        q = 8 + struct.unpack_from("<I", head, 4)[0]
        if len(head) < q + 4:
            continue
        n_ref = struct.unpack_from("<I", head, q)[0]
        q += 4
        ok = True
        names, lens = [], []
        for _ in range(n_ref):
            if len(head) < q + 4:
                ok = False
                break
            ln = struct.unpack_from("<I", head, q)[0]
            if len(head) < q + 4 + ln + 4:
                ok = False
                break
            names.append(head[q + 4:q + 4 + ln - 1].decode())
            lens.append(struct.unpack_from("<I", head, q + 4 + ln)[0])
            q += 8 + ln
        if ok:
            break
    acc = 0
    fb = 0
    while fb < tab.size and acc + int(tab["out_len"][fb]) <= q:
        acc += int(tab["out_len"][fb])
        fb += 1
    skip = q - acc
    tab = tab[fb:]
    t_parse = time.perf_counter() - t0
    tlen = np.array(lens, np.uint32)
    inflated = int(tab["out_len"].sum())
    comp = int(tab["in_len"].sum())

    line = {"bam": os.path.basename(args.bam), "file_bytes": int(data.size), "blocks": int(tab.size), "inflated_bytes": inflated,
            "ratio": inflated / max(1, comp), "contigs": len(lens), "genome": int(tlen.sum())}
    eng = pd.Engine(tlen, device=0)
    eng.set_reduce_hint(1, 0xFFFFFFFF)
    iv_t = np.arange(len(lens), dtype=np.int32); iv_s = np.zeros(len(lens), np.int32); iv_e = tlen.astype(np.int32)

    # ---- resident: compressed bytes in HBM ----
    if not args.no_resident:
        d = torch.from_numpy(data).cuda()
        pad = torch.zeros(64, dtype=torch.uint8, device="cuda")
        d = torch.cat([d, pad])
        ratio = inflated / max(1, comp)
        piece = int((args.piece_mb or min(96.0, max(1.0, 384.0 / ratio))) * (1 << 20))
        # pieces of ~piece compressed bytes
        bounds = [0]
        start_off = int(tab["in_off"][0])
        for i in range(tab.size):
            if int(tab["in_off"][i]) - start_off >= piece:
                bounds.append(i)
                start_off = int(tab["in_off"][i])
        bounds.append(tab.size)
        pieces = [(tab[a:b].copy(), skip if a == 0 else 0) for a, b in zip(bounds[:-1], bounds[1:]) if b > a]

        def step():
            eng.reset()
            for t, sk in pieces:
                eng.push_bgzf(d.data_ptr(), t, first_skip=sk, flags=pd.PD_DEVICE | pd.PD_ASYNC, n_bytes=int(data.size))
            return eng.reduce(iv_t, iv_s, iv_e, 1, 0xFFFFFFFF)

        cov, tot = step()
        reads = eng.stats()["reads_pushed"]
        for _ in range(8):           # more warm-up: every decode slot (PD_BGZF_DEPTH + 1 of them, used round-robin) has seen the biggest
            step()                   # piece and holds its final buffers (a slot that grows costs a device synchronisation + cudaMalloc)
        eng.profile_enable(True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cov2, tot2 = step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.steps
        prof = eng.profile()
        eng.profile_enable(False)
        assert np.array_equal(tot, tot2)
        st = eng.stats()
        kms = {k: v["ms"] / args.steps for k, v in prof.items() if v["launches"]}
        line["resident"] = {"ms_per_step": dt * 1e3, "Mreads_per_s": reads / dt / 1e6, "inflated_GBps": inflated / dt / 1e9, "compressed_GBps": comp / dt / 1e9,
                            "pieces": len(pieces), "piece_mb": piece / (1 << 20), "reads": int(reads), "kernels_ms_per_step": kms,
                            "decode_kernel_inflated_GBps": inflated / (kms.get("k_bgzf_decode", 0) * 1e-3) / 1e9 if kms.get("k_bgzf_decode") else None,
                            "tile_batches": st["tile_path_batches"], "delta_batches": st["delta_path_batches"], "launches": st["kernel_launches"],
                            "total_depth": int(tot.sum()), "covered": int(cov.sum())}
        del d
        torch.cuda.empty_cache()

    if args.no_feed:
        line["parse_s"] = t_parse
        print(json.dumps(line))
        return

    # ---- feed: the product feeder, file in the page cache ----
    H = ctypes.CDLL(os.path.join(ROOT, "pandepth_b200", "libpandepth_host.so"))

    class FS(ctypes.Structure):
        _fields_ = [(n, ctypes.c_uint64) for n in ("file_bytes", "pieces", "bgzf_blocks", "reads")]

    H.pdh_feed_bam.restype = ctypes.c_int
    H.pdh_feed_bam.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t]
    if args.piece_mb:
        os.environ["PANDEPTH_BGZF_PIECE_MB"] = str(args.piece_mb)
    err = ctypes.create_string_buffer(512)
    fs = FS()

    def feed():
        eng.reset()
        rc = H.pdh_feed_bam(eng._h, args.bam.encode(), args.threads, ctypes.byref(fs), err, 512)
        if rc != 0:
            raise RuntimeError(f"pdh_feed_bam rc={rc}: {err.value.decode()}")
        return eng.reduce(iv_t, iv_s, iv_e, 1, 0xFFFFFFFF)

    covf, totf = feed()          # warm-up: allocates the pinned pieces and the device slots
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        covf, totf = feed()
        times.append(time.perf_counter() - t0)
    if "resident" in line:
        assert np.array_equal(totf, tot) and np.array_equal(covf, cov), "feed and resident paths disagree"
    best = min(times)
    line["feed"] = {"s_best": best, "s_all": times, "Mreads_per_s": fs.reads / best / 1e6, "file_GBps": fs.file_bytes / best / 1e9, "threads": args.threads,
                    "pieces": int(fs.pieces), "reads": int(fs.reads), "total_depth": int(totf.sum())}
    line["parse_s"] = t_parse
    print(json.dumps(line))


if __name__ == "__main__":
    main()
