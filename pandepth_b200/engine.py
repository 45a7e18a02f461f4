"""ctypes binding of include/pandepth_b200.h (no logic of its own; every call goes to the CUDA library)."""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np

PD_SORTED, PD_DEVICE, PD_ASYNC, PD_READY, PD_CRC = 1, 2, 4, 8, 16
_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class PDError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"pandepth_b200 error {code}: {msg}")
        self.code = code


class _Stats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint64) for n in ("reads_pushed", "kernel_launches", "h2d_bytes", "d2h_bytes",
                                               "hbm_bytes_allocated", "tile_path_batches", "delta_path_batches", "scan_passes")]


class _KTime(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char * 32), ("launches", ctypes.c_uint64), ("ms", ctypes.c_double)]


def lib_path() -> str:
    # PANDEPTH_B200_LIB selects another build of the same CUDA library (kernel tuning experiments)
    return os.environ.get("PANDEPTH_B200_LIB") or os.path.join(_HERE, "libpandepth_b200.so")


def load_library() -> ctypes.CDLL:
    """Load the CUDA engine.  There is deliberately no fallback: a missing library is an error."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError(f"{path} not built -- run __graft_entry__.build() (nvcc, sm_100a); there is no CPU fallback")
    L = ctypes.CDLL(path)
    vp, i32, i64, u32 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint32
    L.pd_create.restype = vp
    L.pd_create.argtypes = [ctypes.c_int, ctypes.c_int, vp, vp, u32, ctypes.c_int]
    L.pd_create_ranges.restype = vp
    L.pd_create_ranges.argtypes = [ctypes.c_int, ctypes.c_int, vp, vp, vp, u32, ctypes.c_int]
    L.pd_destroy.restype = None; L.pd_destroy.argtypes = [vp]
    L.pd_last_error.restype = ctypes.c_char_p; L.pd_last_error.argtypes = [vp]
    L.pd_set_stream.restype = ctypes.c_int; L.pd_set_stream.argtypes = [vp, vp]
    L.pd_reset.restype = ctypes.c_int; L.pd_reset.argtypes = [vp]
    L.pd_set_reduce_hint.restype = ctypes.c_int; L.pd_set_reduce_hint.argtypes = [vp, u32, u32]
    L.pd_push_contig.restype = ctypes.c_int
    L.pd_push_contig.argtypes = [vp, i32, i64, vp, vp, vp, vp, vp, u32]
    L.pd_push.restype = ctypes.c_int
    L.pd_push.argtypes = [vp, i64, vp, vp, vp, vp, vp, vp, u32]
    L.pd_push_bgzf.restype = ctypes.c_int
    L.pd_push_bgzf.argtypes = [vp, vp, ctypes.c_size_t, vp, i64, u32, u32]
    L.pd_sync.restype = ctypes.c_int; L.pd_sync.argtypes = [vp]
    L.pd_finalize.restype = ctypes.c_int; L.pd_finalize.argtypes = [vp]
    L.pd_reduce.restype = ctypes.c_int
    L.pd_reduce.argtypes = [vp, i64, vp, vp, vp, u32, u32, vp, vp, u32]
    L.pd_reduce_windows.restype = ctypes.c_int
    L.pd_reduce_windows.argtypes = [vp, i32, i32, i32, i64, u32, u32, vp, vp, u32]
    L.pd_depth_read.restype = ctypes.c_int
    L.pd_depth_read.argtypes = [vp, i32, i32, i32, vp, u32]
    L.pd_host_alloc.restype = vp; L.pd_host_alloc.argtypes = [ctypes.c_size_t]
    L.pd_host_free.restype = None; L.pd_host_free.argtypes = [vp]
    L.pd_get_stats.restype = ctypes.c_int; L.pd_get_stats.argtypes = [vp, vp]
    L.pd_profile_enable.restype = ctypes.c_int; L.pd_profile_enable.argtypes = [vp, ctypes.c_int]
    L.pd_profile_get.restype = ctypes.c_int; L.pd_profile_get.argtypes = [vp, vp, ctypes.c_int]
    _LIB = L
    return L


def _p(a):
    """numpy array / int device pointer / None -> c_void_p."""
    if a is None:
        return None
    if isinstance(a, int):
        return ctypes.c_void_p(a)
    return a.ctypes.data_as(ctypes.c_void_p)


def _arr(a, dt):
    """Contiguous array of dtype dt; same-width integer arrays are reinterpreted in place (no copy, stays pinned)."""
    if a is None:
        return None
    a = np.asarray(a)
    dt = np.dtype(dt)
    if a.dtype != dt and a.dtype.kind in "iu" and a.dtype.itemsize == dt.itemsize and a.flags.c_contiguous:
        return a.view(dt)
    return np.ascontiguousarray(a, dtype=dt)


class Engine:
    """One engine per GPU (see include/pandepth_b200.h for the contract of every call)."""

    def __init__(self, target_len: Sequence[int], device: int = 0, owned: Optional[Sequence[int]] = None,
                 xflags: int = 1796, min_mapq: int = -1, own_beg: Optional[Sequence[int]] = None, own_end: Optional[Sequence[int]] = None):
        """owned: per-contig 0/1 mask (contig sharding); own_beg / own_end: per-contig owned range (range sharding)."""
        self._L = load_library()
        self.target_len = np.ascontiguousarray(target_len, dtype=np.uint32)
        if own_beg is not None:
            b = np.ascontiguousarray(own_beg, dtype=np.uint32); z = np.ascontiguousarray(own_end, dtype=np.uint32)
            self._h = self._L.pd_create_ranges(device, int(self.target_len.size), _p(self.target_len), _p(b), _p(z), xflags, min_mapq)
        else:
            own = None if owned is None else np.ascontiguousarray(owned, dtype=np.uint8)
            self._h = self._L.pd_create(device, int(self.target_len.size), _p(self.target_len), _p(own), xflags, min_mapq)
        if not self._h:
            raise PDError(-1, self._L.pd_last_error(None).decode())

    def _ck(self, rc: int):
        if rc != 0:
            raise PDError(rc, self._L.pd_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self._L.pd_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_stream(self, cuda_stream: Optional[int]):
        self._ck(self._L.pd_set_stream(self._h, ctypes.c_void_p(cuda_stream) if cuda_stream else None))

    def reset(self):
        self._ck(self._L.pd_reset(self._h))

    def set_reduce_hint(self, min_dep: int = 1, mask: int = 0xFFFFFFFF):
        self._ck(self._L.pd_set_reduce_hint(self._h, min_dep, mask))

    def push_contig(self, tid: int, pos, cigar, flag=None, mapq=None, cigar_off=None, flags: int = 0, n: Optional[int] = None):
        """Arrays are numpy (host) or, with PD_DEVICE, integer device pointers (then pass n)."""
        if not (flags & PD_DEVICE):
            pos = _arr(pos, np.int32); cigar = _arr(cigar, np.uint32)
            flag = _arr(flag, np.uint16); mapq = _arr(mapq, np.uint8); cigar_off = _arr(cigar_off, np.uint32)
            n = int(pos.size)
            self._keep = (pos, cigar, flag, mapq, cigar_off)
        self._ck(self._L.pd_push_contig(self._h, tid, int(n), _p(pos), _p(flag), _p(mapq), _p(cigar_off), _p(cigar), flags))

    def push(self, tid, pos, cigar, flag=None, mapq=None, cigar_off=None, flags: int = 0):
        tid = _arr(tid, np.int32); pos = _arr(pos, np.int32); cigar = _arr(cigar, np.uint32)
        flag = _arr(flag, np.uint16); mapq = _arr(mapq, np.uint8); cigar_off = _arr(cigar_off, np.uint32)
        self._keep = (tid, pos, cigar, flag, mapq, cigar_off)
        self._ck(self._L.pd_push(self._h, int(tid.size), _p(tid), _p(pos), _p(flag), _p(mapq), _p(cigar_off), _p(cigar), flags))

    def push_bgzf(self, data, blocks, first_skip: int = 0, flags: int = 0, n_bytes: Optional[int] = None):
        """Compressed BGZF chunk -> inflate + record walk + accumulate on the device.  `data` is a uint8 numpy array (or, with
        PD_DEVICE, an integer device pointer plus n_bytes); `blocks` a structured array (in_off u64, in_len u32, out_len u32)."""
        blocks = np.ascontiguousarray(blocks)
        assert blocks.dtype.itemsize == 16
        if not (flags & PD_DEVICE):
            data = _arr(data, np.uint8)
            n_bytes = int(data.size)
        self._keep = (data, blocks)
        self._ck(self._L.pd_push_bgzf(self._h, _p(data), int(n_bytes), _p(blocks), int(blocks.size), first_skip, flags))

    def sync(self):
        self._ck(self._L.pd_sync(self._h))

    def finalize(self):
        self._ck(self._L.pd_finalize(self._h))

    def reduce(self, tid, start0, end0, min_dep: int = 1, mask: int = 0xFFFFFFFF) -> Tuple[np.ndarray, np.ndarray]:
        tid = _arr(tid, np.int32); start0 = _arr(start0, np.int32); end0 = _arr(end0, np.int32)
        cov = np.zeros(tid.size, np.uint64); tot = np.zeros(tid.size, np.uint64)
        self._ck(self._L.pd_reduce(self._h, int(tid.size), _p(tid), _p(start0), _p(end0), min_dep, mask, _p(cov), _p(tot), 0))
        return cov, tot

    def reduce_windows(self, tid: int, start0: int, win: int, n_win: int, min_dep: int = 1, mask: int = 0xFFFFFFFF):
        cov = np.zeros(n_win, np.uint32); tot = np.zeros(n_win, np.uint64)
        self._ck(self._L.pd_reduce_windows(self._h, tid, start0, win, n_win, min_dep, mask, _p(cov), _p(tot), 0))
        return cov, tot

    def depth(self, tid: int, start0: int = 0, n: Optional[int] = None) -> np.ndarray:
        if n is None:
            n = int(self.target_len[tid]) - start0
        out = np.zeros(n, np.uint32)
        self._ck(self._L.pd_depth_read(self._h, tid, start0, n, _p(out), 0))
        return out

    def stats(self) -> dict:
        s = _Stats()
        self._ck(self._L.pd_get_stats(self._h, ctypes.byref(s)))
        return {n: int(getattr(s, n)) for n, _ in _Stats._fields_}

    def profile_enable(self, on: bool = True):
        self._ck(self._L.pd_profile_enable(self._h, 1 if on else 0))

    def profile(self) -> dict:
        buf = (_KTime * 16)()
        k = self._L.pd_profile_get(self._h, ctypes.cast(buf, ctypes.c_void_p), 16)
        if k < 0:
            self._ck(k)
        return {buf[i].name.decode(): {"launches": int(buf[i].launches), "ms": float(buf[i].ms)} for i in range(min(k, 16))}
