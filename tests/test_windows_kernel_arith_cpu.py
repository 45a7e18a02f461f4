"""
CPU: the index arithmetic of k_reduce_windows (pandepth_b200/csrc/pd_kernels.cuh) for window widths the GPU suite does not
reach (it runs win = 7, 33, 100, 1000): a line-by-line host translation of the kernel's three warp-level branches -- four
windows per iteration with up to five loads per lane (32 <= win <= 160), the 128-bit aligned fast path, and the 32-bit
lane-strided branch (win <= 512) -- with lanes vectorised in numpy, compared with a direct computation for every width from 32 to
164 and a few up to 512, both depth masks, several -d values, short last windows; any load outside [0, len + padding) fails.
This checks the translation, not the SASS: the widths go into the GPU suite first thing next round.
"""
import numpy as np
lane = np.arange(32, dtype=np.int64)
def emulate(depth, length, start0, win, n_win, min_dep, mask, nwarps=3):
    """line-by-line host translation of k_reduce_windows (pd_kernels.cuh), lanes vectorised; returns (cov, sum)"""
    cov_out = np.full(n_win, 0xDEADBEEF, np.uint64); sum_out = np.full(n_win, 0xDEADBEEF, np.uint64)
    def ld(idx, ok):   # guarded loads: out-of-contig indices must never be touched
        idx = np.where(ok, idx, 0)
        assert np.all((idx >= 0) & (idx < depth.size)), "out-of-bounds load"
        return np.where(ok, depth[idx].astype(np.uint64), 0)
    if 32 <= win <= 160 and min_dep >= 1 and win * mask < (1 << 32):
        WB, NL = 4, 5
        ngroups = (n_win + WB - 1) // WB
        md, mk = min_dep, mask
        vec_ok = md == 1 and (win & 3) == 0 and win <= 128 and (start0 & 3) == 0
        for warp in range(nwarps):
            g = warp
            while g < ngroups:
                k0 = g * WB; lo0 = start0 + k0 * win
                rem64 = length - lo0
                rem = 0 if rem64 <= 0 else min(rem64, WB * 160)
                nw = min(n_win - k0, WB)
                if vec_ok and nw == WB and rem >= WB * win:
                    q = win >> 2
                    for j in range(WB):
                        ok = lane < q
                        c = np.zeros(32, np.uint64); s = np.zeros(32, np.uint64)
                        for comp in range(4):
                            v = ld(lo0 + (j * q + lane) * 4 + comp, ok) & mk
                            c += (v != 0); s += v
                        cov_out[k0 + j] = c.sum() & 0xFFFFFFFF; sum_out[k0 + j] = s.sum() & 0xFFFFFFFF
                    g += nwarps; continue
                x = {}
                for j in range(WB):
                    off = j * win
                    wl = rem - off if rem > off else 0
                    if wl > win: wl = win
                    if j >= nw: wl = 0
                    for i in range(NL):
                        t = lane + 32 * i
                        x[j, i] = ld(lo0 + lane + off + 32 * i, t < wl)
                for j in range(WB):
                    c = np.zeros(32, np.uint64); s = np.zeros(32, np.uint64)
                    for i in range(NL):
                        v = x[j, i] & mk
                        cc = v >= md
                        c += cc; s += np.where(cc, v, 0)
                    if j < nw: cov_out[k0 + j] = c.sum() & 0xFFFFFFFF; sum_out[k0 + j] = s.sum() & 0xFFFFFFFF
                g += nwarps
    elif 32 <= win <= 512 and win * mask < (1 << 32):
        for warp in range(nwarps):
            k = warp
            while k < n_win:
                lo = start0 + k * win; hi = min(lo + win, length)
                c = np.zeros(32, np.uint64); s = np.zeros(32, np.uint64)
                q = lo + lane
                while np.any(q < hi):
                    ok = q < hi
                    v = ld(q, ok) & mask
                    cc = ok & (v >= min_dep)
                    c += cc; s += np.where(cc, v, 0)
                    q = q + 32
                cov_out[k] = c.sum() & 0xFFFFFFFF; sum_out[k] = s.sum() & 0xFFFFFFFF
                k += nwarps
    else:
        return None
    return cov_out, sum_out

def test_window_branches_for_all_widths():
    rng = np.random.default_rng(3)
    checked = 0
    for win in list(range(32, 165)) + [200, 256, 300, 511, 512]:
      for trial in range(3):
            length = int(rng.integers(2, 40 * win))
            depth = rng.integers(0, 6, length + 64).astype(np.uint32)            # + padding elements like the engine's layout
            depth[rng.integers(0, length, 5)] = rng.integers(200000, 4000000000, 5)   # 18-bit wrap material
            depth[length:] = 0xABCDEF                                            # anything read past `len` would show
            for mask in (0x3FFFF, 0xFFFFFFFF):
                for min_dep in (1, 2, 5):
                    n_win = (length - 1 + win - 1) // win if length > 1 else 0      # mode 6: for (j=1; j<CC; j+=W)
                    r = emulate(depth, length, 0, win, n_win, min_dep, mask)
                    if r is None: continue
                    cov, tot = r
                    for k in range(n_win):
                        a, b = k * win, min(k * win + win, length)
                        v = depth[a:b].astype(np.uint64) & mask
                        ok = v >= min_dep
                        assert cov[k] == ok.sum() and tot[k] == (v[ok].sum() & 0xFFFFFFFF), (win, length, mask, min_dep, k, cov[k], ok.sum(), tot[k], v[ok].sum())
                        checked += 1
    assert checked > 20000

