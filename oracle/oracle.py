"""ctypes wrapper around liboracle.so (oracle/sst_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (suffix-array-searching_b200/sst_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

MAX = 0x7FFFFFFF
SIMPLE, COMPACT, L1, OVERLAP, MAP = 1, 2, 3, 4, 5
VARIANTS = {"simple": SIMPLE, "compact": COMPACT, "l1": L1, "overlap": OVERLAP, "map": MAP}


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "sst_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B" if force else "-s"])
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        vp, sz, u32p, u64p, u8p = C.c_void_p, C.c_size_t, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint8)
        L.orc_height.restype = sz
        L.orc_height.argtypes = [sz, sz]
        L.orc_layer_size.restype = sz
        L.orc_layer_size.argtypes = [sz, sz, sz, sz]
        L.orc_node_find.restype = sz
        L.orc_node_find.argtypes = [vp, C.c_uint32]
        L.orc_lower_bound.restype = None
        L.orc_lower_bound.argtypes = [vp, sz, vp, sz, vp, vp]
        L.orc_stree_build.restype = vp
        L.orc_stree_build.argtypes = [vp, sz, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_pstree_build.restype = vp
        L.orc_pstree_build.argtypes = [vp, sz, sz, C.c_int]
        L.orc_tree_free.restype = None
        L.orc_tree_free.argtypes = [vp]
        for f in ("orc_tree_nodes", "orc_tree_layers", "orc_tree_levels", "orc_tree_size_bytes"):
            getattr(L, f).restype = sz
            getattr(L, f).argtypes = [vp]
        for f in ("orc_tree_offsets", "orc_tree_layer_sizes", "orc_tree_image", "orc_tree_params", "orc_tree_prefix_map"):
            getattr(L, f).restype = None
            getattr(L, f).argtypes = [vp, vp]
        L.orc_tree_search.restype = None
        L.orc_tree_search.argtypes = [vp, vp, sz, vp, vp]
        L.orc_stree_batch_final.restype = C.c_double
        L.orc_stree_batch_final.argtypes = [vp, vp, sz, vp, C.c_int]
        L.orc_stree_batch_interleave.restype = C.c_double
        L.orc_stree_batch_interleave.argtypes = [vp, vp, sz, vp, C.c_int]
        L.orc_sa_build.restype = None
        L.orc_sa_build.argtypes = [vp, sz, vp]
        L.orc_sa_check.restype = C.c_uint64
        L.orc_sa_check.argtypes = [vp, sz, vp, C.c_int]
        L.orc_sa_search.restype = None
        L.orc_sa_search.argtypes = [vp, sz, vp, vp, vp, sz, vp, vp, vp, vp]
        L.orc_sa_search_cmp.restype = None
        L.orc_sa_search_cmp.argtypes = [vp, sz, vp, vp, vp, sz, vp, vp]
        L.orc_sa_search_mlr.restype = None
        L.orc_sa_search_mlr.argtypes = [vp, sz, vp, vp, vp, sz, vp, vp]
        L.orc_sa_search_batch32.restype = C.c_double
        L.orc_sa_search_batch32.argtypes = [vp, sz, vp, vp, vp, sz, vp, vp, C.c_int]
        L.orc_eytzinger_build.restype = None
        L.orc_eytzinger_build.argtypes = [vp, sz, vp]
        L.orc_eytzinger_search.restype = None
        L.orc_eytzinger_search.argtypes = [vp, sz, vp, sz, vp]
        L.orc_has_avx2.restype = C.c_int
        _lib = L
    return _lib


def _p(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _u32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint32)


def height(n: int, B: int = 16) -> int:
    return lib().orc_height(n, B)


def layer_size(n: int, h: int, H: int, B: int = 16) -> int:
    return lib().orc_layer_size(n, h, H, B)


def node_find(node16, q: int) -> int:
    node = _u32(node16)
    assert node.size == 16
    return lib().orc_node_find(_p(node), q)


def lower_bound(vals, qs):
    """sst/src/binary_search.rs:36-49: (values, indices)."""
    vals, qs = _u32(vals), _u32(qs)
    ov = np.empty(qs.size, np.uint32)
    oi = np.empty(qs.size, np.uint64)
    lib().orc_lower_bound(_p(vals), vals.size, _p(qs), qs.size, _p(ov), _p(oi))
    return ov, oi


class Tree:
    """Restated STree / PartitionedSTree (owns the oracle handle)."""

    def __init__(self, handle, vals):
        if not handle:
            raise ValueError("oracle build returned None")
        self.h = handle
        self._vals = vals  # keep alive

    @classmethod
    def stree(cls, vals, B=16, left_max=False, reverse=False, full=False):
        vals = _u32(vals)
        return cls(lib().orc_stree_build(_p(vals), vals.size, B, int(left_max), int(reverse), int(full)), vals)

    @classmethod
    def pstree(cls, vals, b, variant):
        vals = _u32(vals)
        v = VARIANTS[variant] if isinstance(variant, str) else variant
        h = lib().orc_pstree_build(_p(vals), vals.size, b, v)
        return cls(h, vals) if h else None

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_tree_free(self.h)
            self.h = None

    @property
    def nodes(self):
        return lib().orc_tree_nodes(self.h)

    @property
    def layers(self):
        return lib().orc_tree_layers(self.h)

    @property
    def levels(self):
        return lib().orc_tree_levels(self.h)

    @property
    def size_bytes(self):
        return lib().orc_tree_size_bytes(self.h)

    @property
    def offsets(self):
        out = np.zeros(self.levels, np.uint64)
        lib().orc_tree_offsets(self.h, _p(out))
        return out

    @property
    def layer_sizes(self):
        out = np.zeros(self.levels, np.uint64)
        lib().orc_tree_layer_sizes(self.h, _p(out))
        return out

    @property
    def params(self):
        out = np.zeros(8, np.uint64)
        lib().orc_tree_params(self.h, _p(out))
        keys = ["shift", "parts", "bpp", "l1", "overlap", "has_overlap", "max_bucket", "prefix_map_len"]
        return dict(zip(keys, (int(x) for x in out)))

    @property
    def prefix_map(self):
        out = np.zeros(self.params["prefix_map_len"], np.uint32)
        lib().orc_tree_prefix_map(self.h, _p(out))
        return out

    def image(self):
        out = np.empty(self.nodes * 16, np.uint32)
        lib().orc_tree_image(self.h, _p(out))
        return out

    def search(self, qs, want_slot=False):
        qs = _u32(qs)
        ov = np.empty(qs.size, np.uint32)
        os_ = np.empty(qs.size, np.uint64) if want_slot else None
        lib().orc_tree_search(self.h, _p(qs), qs.size, _p(ov), _p(os_))
        return (ov, os_) if want_slot else ov

    def batch_final(self, qs, threads=1):
        """CPU baseline (s_tree.rs:303-326). Returns (values, seconds)."""
        qs = _u32(qs)
        ov = np.empty(qs.size, np.uint32)
        secs = lib().orc_stree_batch_final(self.h, _p(qs), qs.size, _p(ov), threads)
        if secs < 0:
            raise RuntimeError("batch_final needs a plain B=16 tree")
        return ov, secs

    def batch_interleave(self, qs, threads=1):
        """The reference's fastest CPU scheme, batch_interleave_all_128 (s_tree.rs:684-832). Returns (values, seconds)."""
        qs = _u32(qs)
        ov = np.empty(qs.size, np.uint32)
        secs = lib().orc_stree_batch_interleave(self.h, _p(qs), qs.size, _p(ov), threads)
        if secs < 0:
            raise RuntimeError("batch_interleave needs a plain B=16 tree of height <= 8")
        return ov, secs


def eytzinger_build(vals) -> np.ndarray:
    """sst/src/eytzinger.rs:37-63: one-based BFS layout, element 0 is u32::MAX."""
    vals = _u32(vals)
    out = np.empty(vals.size + 1, np.uint32)
    lib().orc_eytzinger_build(_p(vals), vals.size, _p(out))
    return out


def eytzinger_search(e, qs) -> np.ndarray:
    """sst/src/eytzinger.rs:82-89."""
    e, qs = _u32(e), _u32(qs)
    out = np.empty(qs.size, np.uint32)
    lib().orc_eytzinger_search(_p(e), e.size, _p(qs), qs.size, _p(out))
    return out


# ----------------------------------------------------------------------------------------------
# suffix arrays
# ----------------------------------------------------------------------------------------------
def pack_patterns(pats):
    """list of bytes-like -> (flat uint8 array padded with 64 zero bytes, uint64 offsets[npat+1])."""
    lens = np.fromiter((len(p) for p in pats), dtype=np.uint64, count=len(pats))
    off = np.zeros(len(pats) + 1, np.uint64)
    np.cumsum(lens, out=off[1:])
    flat = np.zeros(int(off[-1]) + 64, np.uint8)
    if len(pats):
        flat[: int(off[-1])] = np.frombuffer(b"".join(bytes(p) for p in pats), np.uint8)
    return flat, off


def sa_build(text: np.ndarray) -> np.ndarray:
    text = np.ascontiguousarray(text, np.uint8)
    sa = np.empty(text.size, np.uint32)
    lib().orc_sa_build(_p(text), text.size, _p(sa))
    return sa


def sa_check(text, sa, threads=1) -> int:
    text = np.ascontiguousarray(text, np.uint8)
    sa = _u32(sa)
    return int(lib().orc_sa_check(_p(text), text.size, _p(sa), threads))


def sa_search(text, sa, flat, off, want_hi=True):
    """sas/src/sa_search.rs:98-112. Returns (lo, hi, pos, probes)."""
    text = np.ascontiguousarray(text, np.uint8)
    sa = _u32(sa)
    npat = off.size - 1
    lo = np.empty(npat, np.uint32)
    hi = np.empty(npat, np.uint32) if want_hi else None
    pos = np.empty(npat, np.uint32)
    cnt = C.c_uint64(0)
    lib().orc_sa_search(_p(text), text.size, _p(sa), _p(flat), _p(off), npat, _p(lo), _p(hi), _p(pos), C.byref(cnt))
    return lo, hi, pos, cnt.value


def sa_search_cmp(text_padded, n, sa, flat, off):
    """sas/src/sa_search.rs:121-136 (16-byte SIMD cmp). text_padded must extend >= 32 B past n."""
    sa = _u32(sa)
    npat = off.size - 1
    lo = np.empty(npat, np.uint32)
    pos = np.empty(npat, np.uint32)
    lib().orc_sa_search_cmp(_p(text_padded), n, _p(sa), _p(flat), _p(off), npat, _p(lo), _p(pos))
    return lo, pos


def sa_search_mlr(text, sa, flat, off):
    text = np.ascontiguousarray(text, np.uint8)
    sa = _u32(sa)
    npat = off.size - 1
    lo = np.empty(npat, np.uint32)
    bc = C.c_uint64(0)
    lib().orc_sa_search_mlr(_p(text), text.size, _p(sa), _p(flat), _p(off), npat, _p(lo), C.byref(bc))
    return lo, bc.value


def sa_search_batch32(text, sa, flat, off, threads=1):
    text = np.ascontiguousarray(text, np.uint8)
    sa = _u32(sa)
    npat = off.size - 1
    lo = np.empty(npat, np.uint32)
    pos = np.empty(npat, np.uint32)
    secs = lib().orc_sa_search_batch32(_p(text), text.size, _p(sa), _p(flat), _p(off), npat, _p(lo), _p(pos), threads)
    return lo, pos, secs


# ----------------------------------------------------------------------------------------------
# input formats (pure Python / numpy restatements)
# ----------------------------------------------------------------------------------------------
def read_fasta(data: bytes) -> np.ndarray:
    """suffix-array-searching/src/util.rs:144-169 read_fasta_file: needletail::parse_fastx_file records.  The format is
    picked from the first byte (needletail 0.5.1, pinned at Cargo.lock:686-687; not vendored): '>' = FASTA (a header line starts with '>', sequence
    lines are concatenated with line ends stripped), '@' = FASTQ (four-line records: header, sequence, '+', qualities;
    only the second line is sequence).  map[] sends A/C/G/T in either case to 0..3 and every other byte to 0."""
    m = np.zeros(256, np.uint8)
    for ch, v in ((b"A", 0), (b"C", 1), (b"G", 2), (b"T", 3), (b"a", 0), (b"c", 1), (b"g", 2), (b"t", 3)):
        m[ch[0]] = v
    out = []
    fastq = data[:1] == b"@"
    for ln, line in enumerate(data.split(b"\n")):
        if (ln % 4 != 1) if fastq else line.startswith(b">"):
            continue
        line = line.replace(b"\r", b"")
        if line:
            out.append(m[np.frombuffer(line, np.uint8)])
    return np.concatenate(out) if out else np.zeros(0, np.uint8)


def kmer_keys(codes, k=16, max_keys=None, sort=True) -> np.ndarray:
    """static-search-tree/src/bin/bench.rs:60-76 (--human): rolling 2-bit pack of k bases masked to 2k bits
    and to i32::MAX; vals[0] = MAX; then bench.rs:89 sorts."""
    codes = np.asarray(codes, np.uint8)
    n = codes.size
    if n < k:
        return np.zeros(0, np.uint32)
    count = n - k + 1 if max_keys is None else min(n - k + 1, max_keys)
    vals = np.empty(count, np.uint32)
    key = 0
    for i in range(k - 1):
        key = (key << 2) | int(codes[i])
    mask = (1 << (2 * k)) - 1
    for i in range(k - 1, k - 1 + count):
        key = ((key << 2) | int(codes[i])) & mask
        vals[i - (k - 1)] = key & MAX
    if count:
        vals[0] = MAX
    if sort:
        vals.sort()
    return vals
