"""sst_b200 -- host-side mirror of the reference's operator API over libsst_b200.so (C ABI).

The reference is Rust; its toolchain is absent here, so this module plays the role of the Rust
shim (see suffix-array-searching_b200/rust/ and INTEGRATION.md): same names, argument meaning
and error behaviour as

  * trait SearchIndex  (static-search-tree/src/lib.rs:30-48):  new / size / layers / query(_one)
  * STree16::new_params (static-search-tree/src/s_tree.rs:72), STree15
  * PartitionedSTree16{,C,L,O,M}::new / try_new (static-search-tree/src/partitioned_s_tree.rs:231-241,354-364)
  * SaNaive::build + binary_search (suffix-array-searching/src/sa_search.rs:30-57,98-112)

Reference errors are panics; here they are `SstError` (status != 0), and `try_new` returns None
where the reference returns None.  There is NO CPU fallback: if the CUDA library or a B200 is
missing every call raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

MAX = 0x7FFFFFFF

LEFT_MAX, REVERSE_STORAGE, FULL_ARRAY = 1, 2, 4
PLAIN, SIMPLE, COMPACT, L1, OVERLAPPING, MAP, EYTZINGER = 0, 1, 2, 3, 4, 5, 6
SCHEME_AUTO, SCHEME_GROUP4, SCHEME_GROUP16, SCHEME_GROUP2, SCHEME_GENERIC, SCHEME_TABLE, SCHEME_BINSEARCH, SCHEME_BUCKETED = 0, 1, 2, 3, 4, 5, 6, 7
SA_BINARY, SA_MLR = 0, 1
ERR_CUDA, ERR_ARG, ERR_CAPACITY, ERR_UNSUPPORTED = 1, 2, 3, 4

_PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("SST_B200_LIB") or os.path.join(_PKG_DIR, "libsst_b200.so")  # (SST_B200_LIB: A/B runs of a differently compiled library)


class SstError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"sst_b200 error {status}: {msg}")
        self.status = status


_lib = None


def lib() -> C.CDLL:
    """Load libsst_b200.so. Fails loudly when it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SstError(ERR_CUDA, f"{LIB_PATH} not built; run `python -c 'import __graft_entry__ as g; g.build()'`")
    L = C.CDLL(LIB_PATH)
    vp, sz, i32, u32 = C.c_void_p, C.c_size_t, C.c_int, C.c_uint32
    sig = {
        "sst_last_error": (C.c_char_p, []),
        "sst_last_status": (i32, []),
        "sst_device_count": (i32, []),
        "sst_version": (C.c_char_p, []),
        "sst_set_option": (i32, [C.c_char_p, C.c_longlong]),
        "sst_get_option": (i32, [C.c_char_p, vp]),
        "sst_reset_options": (None, []),
        "sst_option_count": (i32, []),
        "sst_option_name": (C.c_char_p, [i32]),
        "sst_query_reserve": (i32, [vp, sz, i32]),
        "sst_query_release": (None, []),
        "sst_query_calibrate": (i32, [vp, sz, vp]),
        "sst_bind_thread_to_device": (i32, [i32]),
        "sst_host_alloc": (vp, [sz]),
        "sst_host_free": (None, [vp]),
        "sst_stree_build": (vp, [vp, sz, u32, u32, i32]),
        "sst_stree_build_device": (vp, [vp, sz, u32, u32, i32]),
        "sst_pstree_build": (vp, [vp, sz, u32, i32, i32]),
        "sst_pstree_build_device": (vp, [vp, sz, u32, i32, i32]),
        "sst_eytzinger_build": (vp, [vp, sz, i32]),
        "sst_eytzinger_build_device": (vp, [vp, sz, i32]),
        "sst_index_image_words": (sz, [vp]),
        "sst_index_free": (None, [vp]),
        "sst_index_size_bytes": (sz, [vp]),
        "sst_index_layers": (sz, [vp]),
        "sst_index_len": (sz, [vp]),
        "sst_index_device": (i32, [vp]),
        "sst_index_variant": (i32, [vp]),
        "sst_index_nodes": (sz, [vp]),
        "sst_index_levels": (sz, [vp]),
        "sst_index_offsets": (i32, [vp, vp]),
        "sst_index_image": (i32, [vp, vp]),
        "sst_index_params": (i32, [vp, vp]),
        "sst_index_prefix_map": (i32, [vp, vp]),
        "sst_query": (i32, [vp, vp, sz, vp, vp, i32]),
        "sst_query_device": (i32, [vp, vp, sz, vp, vp, i32, vp]),
        "sst_query_launches": (i32, [vp, i32]),
        "sst_query_plan": (i32, [vp, sz, i32, i32, vp, vp]),
        "sst_last_stage_ms": (i32, [vp, i32]),
        "sst_sa_build": (vp, [vp, sz, i32]),
        "sst_sa_build_device": (vp, [vp, sz, i32]),
        "sst_sa_from_parts": (vp, [vp, sz, vp, i32]),
        "sst_sa_free": (None, [vp]),
        "sst_sa_len": (sz, [vp]),
        "sst_sa_get": (i32, [vp, vp]),
        "sst_sa_check": (i32, [vp, vp]),
        "sst_sa_gather": (i32, [vp, vp, sz, vp]),
        "sst_sa_search": (i32, [vp, vp, vp, sz, i32, vp, vp, vp]),
        "sst_sa_search_device": (i32, [vp, vp, vp, sz, i32, vp, vp, vp, vp]),
        "sst_sa_search_probes": (i32, [vp, vp, vp, sz, vp, vp]),
        "sst_multi_query_device": (i32, [vp, vp, vp, vp, vp, i32]),
        "sst_multi_sa_build": (vp, [vp, sz, vp, i32]),
        "sst_multi_sa_from_parts": (vp, [vp, sz, vp, vp, i32]),
        "sst_multi_sa_search": (i32, [vp, vp, vp, sz, i32, vp, vp, vp]),
        "sst_multi_sa_devices": (i32, [vp]),
        "sst_multi_sa_free": (None, [vp]),
        "sst_fasta_encode": (i32, [vp, sz, vp, vp, i32]),
        "sst_fasta_encode_device": (i32, [vp, sz, vp, vp, i32]),
        "sst_kmer_keys": (i32, [vp, sz, u32, sz, vp, vp, i32, i32]),
        "sst_kmer_keys_device": (i32, [vp, sz, u32, sz, vp, vp, i32, i32]),
        "sst_multi_stree_build": (vp, [vp, sz, u32, u32, vp, i32]),
        "sst_multi_pstree_build": (vp, [vp, sz, u32, i32, vp, i32]),
        "sst_multi_query": (i32, [vp, vp, sz, vp, vp, i32]),
        "sst_multi_devices": (i32, [vp]),
        "sst_multi_free": (None, [vp]),
        "sst_time_query_device": (C.c_double, [vp, vp, sz, vp, vp, i32, i32, i32]),
        "sst_probe_gather64": (C.c_double, [i32, sz, sz, i32, i32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError here == header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


EXPORTS = None  # filled lazily by tests from include/sst_b200.h


def _raise():
    L = lib()
    raise SstError(L.sst_last_status(), (L.sst_last_error() or b"").decode())


def _check(rc: int):
    if rc != 0:
        _raise()


def device_count() -> int:
    return lib().sst_device_count()


def set_option(name: str, value: int) -> None:
    """Tuning / A-B option of the library (csrc/common.cuh: SST_OPTION_LIST); `name` with or without the SST_ prefix."""
    _check(lib().sst_set_option(name.encode(), int(value)))


def get_option(name: str) -> int:
    v = C.c_longlong(0)
    _check(lib().sst_get_option(name.encode(), C.byref(v)))
    return v.value


def reset_options() -> None:
    lib().sst_reset_options()


def option_names():
    L = lib()
    return [L.sst_option_name(i).decode() for i in range(L.sst_option_count())]


class options:
    """Context manager: `with sst.options(BK_MIN_N=0, BK_R=256): ...` sets options and restores the previous values."""

    def __init__(self, **kw):
        self.kw = kw
        self.old = {}

    def __enter__(self):
        for k, v in self.kw.items():
            self.old[k] = get_option(k)
            set_option(k, v)
        return self

    def __exit__(self, *exc):
        for k, v in self.old.items():
            set_option(k, v)
        return False


def bind_thread_to_device(device: int) -> int:
    """CPU affinity of the calling thread := the CPUs local to `device` (NUMA); returns the set size, 0 if unknown."""
    return lib().sst_bind_thread_to_device(device)


class PinnedArray:
    """numpy view over page-locked host memory from sst_host_alloc (full-speed host path)."""

    def __init__(self, n: int, dtype=np.uint32):
        self.nbytes = int(n) * np.dtype(dtype).itemsize
        self._p = lib().sst_host_alloc(self.nbytes)
        if not self._p:
            _raise()
        buf = (C.c_char * max(self.nbytes, 1)).from_address(self._p)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(n))

    def __del__(self):
        if getattr(self, "_p", None) and _lib is not None:
            self.array = None
            _lib.sst_host_free(C.c_void_p(self._p))
            self._p = None


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _host_u32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint32)


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _torch_stream_ptr(t):
    import torch

    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


# ----------------------------------------------------------------------------------------------
# SearchIndex (static-search-tree/src/lib.rs:30-48)
# ----------------------------------------------------------------------------------------------
class SearchIndex:
    """An immutable index on one GPU. `query` mirrors SearchIndex::query / SearchScheme::query."""

    def __init__(self, handle):
        if not handle:
            _raise()
        self._h = C.c_void_p(handle)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _lib is not None:
            _lib.sst_index_free(h)
            self._h = None

    # SearchIndex::size (bytes) / ::layers
    def size(self) -> int:
        return lib().sst_index_size_bytes(self._h)

    def layers(self) -> int:
        return lib().sst_index_layers(self._h)

    def __len__(self) -> int:
        return lib().sst_index_len(self._h)

    @property
    def device(self) -> int:
        return lib().sst_index_device(self._h)

    @property
    def variant(self) -> int:
        return lib().sst_index_variant(self._h)

    # layout introspection (the struct fields of s_tree.rs:14-17 / partitioned_s_tree.rs:19-32)
    @property
    def offsets(self) -> np.ndarray:
        out = np.zeros(lib().sst_index_levels(self._h), np.uint64)
        _check(lib().sst_index_offsets(self._h, _ptr(out)))
        return out

    def image(self) -> np.ndarray:
        out = np.empty(lib().sst_index_image_words(self._h), np.uint32)
        _check(lib().sst_index_image(self._h, _ptr(out)))
        return out

    @property
    def params(self) -> dict:
        out = np.zeros(8, np.uint64)
        _check(lib().sst_index_params(self._h, _ptr(out)))
        keys = ["shift", "parts", "bpp", "l1", "overlap", "has_overlap", "max_bucket", "prefix_map_len"]
        return dict(zip(keys, (int(x) for x in out)))

    @property
    def prefix_map(self) -> np.ndarray:
        out = np.zeros(self.params["prefix_map_len"], np.uint32)
        _check(lib().sst_index_prefix_map(self._h, _ptr(out)))
        return out

    def query(self, qs, scheme: int = SCHEME_AUTO, want_index: bool = False):
        """SearchScheme::query: values of the first key >= q, in query order.

        numpy / list input  -> host path (H2D, kernel, D2H inside the call), returns numpy.
        torch CUDA tensor   -> device path on torch's current stream, returns torch tensors.
        With want_index=True returns (values, indices).
        """
        L = lib()
        if _is_torch(qs):
            import torch

            if not qs.is_cuda:
                return self.query(qs.numpy(), scheme, want_index)
            assert qs.dtype in (torch.int32, torch.uint32) and qs.is_contiguous()
            assert qs.device.index == self.device, "queries must live on the index's device"
            vals = torch.empty_like(qs)
            idx = torch.empty(qs.numel(), dtype=torch.int64, device=qs.device) if want_index else None
            _check(
                L.sst_query_device(
                    self._h, C.c_void_p(qs.data_ptr()), qs.numel(), C.c_void_p(vals.data_ptr()),
                    C.c_void_p(idx.data_ptr()) if want_index else None, scheme, _torch_stream_ptr(qs),
                )
            )
            return (vals, idx) if want_index else vals
        q = _host_u32(qs)
        vals = np.empty(q.size, np.uint32)
        idx = np.empty(q.size, np.uint64) if want_index else None
        _check(L.sst_query(self._h, _ptr(q), q.size, _ptr(vals), _ptr(idx), scheme))
        return (vals, idx) if want_index else vals

    def reserve(self, nq: int, want_index: bool = False) -> None:
        """Pre-size the calling thread's pipeline scratch so that device-side queries of up to nq allocate nothing."""
        _check(lib().sst_query_reserve(self._h, int(nq), int(want_index)))

    def calibrate(self, max_nq: int = 1 << 25) -> int:
        """Measure on this device from which batch size on the reordered-batch pipeline beats the direct kernel and make
        SCHEME_AUTO use it for this index.  Returns the crossover (0: nothing to calibrate; 2**64 - 1: the pipeline never won)."""
        out = C.c_size_t(0)
        _check(lib().sst_query_calibrate(self._h, int(max_nq), C.byref(out)))
        return out.value

    def query_one(self, q: int, scheme: int = SCHEME_AUTO) -> int:
        """SearchScheme::query_one (lib.rs:52-54)."""
        return int(self.query(np.array([q], np.uint32), scheme)[0])

    # `search` is the name of the single-query method on the reference structs (s_tree.rs:196)
    search = query_one


def _build(host_fn, dev_fn, vals, *args, device=0):
    L = lib()
    if _is_torch(vals) and vals.is_cuda:
        assert vals.is_contiguous() and vals.element_size() == 4
        return getattr(L, dev_fn)(C.c_void_p(vals.data_ptr()), vals.numel(), *args, vals.device.index)
    v = _host_u32(vals.numpy() if _is_torch(vals) else vals)
    return getattr(L, host_fn)(_ptr(v), v.size, *args, device)


class STree(SearchIndex):
    """STree<B,16> (static-search-tree/src/s_tree.rs:14-20)."""

    B = 16

    def __init__(self, vals, left_max=False, reverse_storage=False, full_array=False, device=0):
        flags = (LEFT_MAX if left_max else 0) | (REVERSE_STORAGE if reverse_storage else 0) | (FULL_ARRAY if full_array else 0)
        super().__init__(_build("sst_stree_build", "sst_stree_build_device", vals, self.B, flags, device=device))

    @classmethod
    def new(cls, vals, device=0):
        """SearchIndex::new == new_params(vals, false, false, false) (s_tree.rs:47-50)."""
        return cls(vals, device=device)

    @classmethod
    def new_params(cls, vals, left_max, reverse_storage, full_array, device=0):
        """STree::new_params (s_tree.rs:72-77)."""
        return cls(vals, left_max, reverse_storage, full_array, device=device)


class STree16(STree):
    B = 16


class STree15(STree):
    B = 15


class Eytzinger(SearchIndex):
    """Eytzinger (static-search-tree/src/eytzinger.rs:9-78): baseline layout, unsigned compares,
    a query above every key returns 0xffffffff."""

    def __init__(self, vals, device=0):
        super().__init__(_build("sst_eytzinger_build", "sst_eytzinger_build_device", vals, device=device))

    @classmethod
    def new(cls, vals, device=0):
        return cls(vals, device=device)


class PartitionedSTree(SearchIndex):
    """PartitionedSTree<16,16,Tp> (static-search-tree/src/partitioned_s_tree.rs:19-98)."""

    VARIANT = SIMPLE

    def __init__(self, vals, b: int, device=0):
        super().__init__(_build("sst_pstree_build", "sst_pstree_build_device", vals, b, self.VARIANT, device=device))

    @classmethod
    def new(cls, vals, b: int, device=0):
        """`new` = try_new(..).unwrap() (partitioned_s_tree.rs:231-233,354-356)."""
        return cls(vals, b, device=device)

    @classmethod
    def try_new(cls, vals, b: int, device=0):
        """Returns None where the reference returns None (memory caps)."""
        try:
            return cls(vals, b, device=device)
        except SstError as e:
            if e.status == ERR_CAPACITY:
                return None
            raise


class PartitionedSTree16(PartitionedSTree):
    VARIANT = SIMPLE


class PartitionedSTree16C(PartitionedSTree):
    VARIANT = COMPACT


class PartitionedSTree16L(PartitionedSTree):
    VARIANT = L1


class PartitionedSTree16O(PartitionedSTree):
    VARIANT = OVERLAPPING


class PartitionedSTree16M(PartitionedSTree):
    VARIANT = MAP


# ----------------------------------------------------------------------------------------------
# input formats: read_fasta_file (sas/util.rs:144-169) and the --human k-mer keys (sst/bin/bench.rs:60-76)
# ----------------------------------------------------------------------------------------------
def read_fasta(data: bytes, device=0) -> np.ndarray:
    """FASTA text -> uint8 codes 0..3 (headers dropped, line ends stripped, non-ACGT -> 0)."""
    buf = np.frombuffer(data, np.uint8)
    out = np.empty(max(buf.size, 1), np.uint8)
    n = C.c_size_t(0)
    _check(lib().sst_fasta_encode(_ptr(buf) if buf.size else None, buf.size, _ptr(out), C.byref(n), device))
    return out[: n.value].copy()


def read_fasta_file(path, device=0) -> np.ndarray:
    with open(path, "rb") as f:
        return read_fasta(f.read(), device)


def kmer_keys(codes, k=16, max_keys=None, sort=True, device=0) -> np.ndarray:
    """bench.rs:60-76: 31-bit k-mer keys of a 2-bit coded sequence, keys[0] = MAX, optionally sorted."""
    c = np.ascontiguousarray(codes, np.uint8)
    cap = max(c.size - k + 1, 0) if max_keys is None else min(max(c.size - k + 1, 0), max_keys)
    out = np.empty(max(cap, 1), np.uint32)
    n = C.c_size_t(0)
    _check(lib().sst_kmer_keys(_ptr(c), c.size, k, cap if max_keys is None else max_keys, _ptr(out), C.byref(n), int(sort), device))
    return out[: n.value].copy()


# ----------------------------------------------------------------------------------------------
# multi-GPU replicas (bench.rs:558-573 chunking rule, one host thread + stream per device)
# ----------------------------------------------------------------------------------------------
class MultiIndex:
    def __init__(self, handle):
        if not handle:
            _raise()
        self._h = C.c_void_p(handle)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _lib is not None:
            _lib.sst_multi_free(h)
            self._h = None

    @classmethod
    def stree(cls, vals, devices, left_max=False, reverse_storage=False, full_array=False, node_b=16):
        v = _host_u32(vals)
        d = np.ascontiguousarray(devices, np.int32)
        flags = (LEFT_MAX if left_max else 0) | (REVERSE_STORAGE if reverse_storage else 0) | (FULL_ARRAY if full_array else 0)
        return cls(lib().sst_multi_stree_build(_ptr(v), v.size, node_b, flags, _ptr(d), d.size))

    @classmethod
    def pstree(cls, vals, b, variant, devices):
        v = _host_u32(vals)
        d = np.ascontiguousarray(devices, np.int32)
        return cls(lib().sst_multi_pstree_build(_ptr(v), v.size, b, variant, _ptr(d), d.size))

    @property
    def n_devices(self) -> int:
        return lib().sst_multi_devices(self._h)

    def query(self, qs, scheme=SCHEME_AUTO, want_index=False):
        q = _host_u32(qs)
        vals = np.empty(q.size, np.uint32)
        idx = np.empty(q.size, np.uint64) if want_index else None
        _check(lib().sst_multi_query(self._h, _ptr(q), q.size, _ptr(vals), _ptr(idx), scheme))
        return (vals, idx) if want_index else vals


    def query_device(self, shards, scheme=SCHEME_AUTO, want_index=False):
        """Device-resident shards: shards[i] is a contiguous int32/uint32 CUDA tensor on the device of replica i.
        Returns the per-shard value tensors (and index tensors)."""
        import torch

        G = self.n_devices
        assert len(shards) == G
        vals = [torch.empty_like(q) for q in shards]
        idx = [torch.empty(q.numel(), dtype=torch.int64, device=q.device) for q in shards] if want_index else None
        for q in shards:
            torch.cuda.synchronize(q.device)  # the shards may come from other streams
        PP = C.c_void_p * G
        qp = PP(*[q.data_ptr() for q in shards])
        vp_ = PP(*[v.data_ptr() for v in vals])
        ip = PP(*[i.data_ptr() for i in idx]) if want_index else None
        nq = (C.c_size_t * G)(*[q.numel() for q in shards])
        _check(lib().sst_multi_query_device(self._h, qp, nq, vp_, ip, scheme))
        return (vals, idx) if want_index else vals


class MultiSa:
    """Text + suffix array replicated on several GPUs (built once, copied peer to peer), patterns sharded contiguously."""

    def __init__(self, handle):
        if not handle:
            _raise()
        self._h = C.c_void_p(handle)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _lib is not None:
            _lib.sst_multi_sa_free(h)
            self._h = None

    @classmethod
    def build(cls, text, devices):
        t = np.ascontiguousarray(text, np.uint8)
        d = np.ascontiguousarray(devices, np.int32)
        return cls(lib().sst_multi_sa_build(_ptr(t), t.size, _ptr(d), d.size))

    @classmethod
    def from_parts(cls, text, sa, devices):
        t = np.ascontiguousarray(text, np.uint8)
        s = _host_u32(sa)
        d = np.ascontiguousarray(devices, np.int32)
        return cls(lib().sst_multi_sa_from_parts(_ptr(t), t.size, _ptr(s), _ptr(d), d.size))

    @property
    def n_devices(self) -> int:
        return lib().sst_multi_sa_devices(self._h)

    def search(self, flat, off, mode=SA_BINARY, want_hi=True):
        flat = np.ascontiguousarray(flat, np.uint8)
        off = np.ascontiguousarray(off, np.uint64)
        npat = off.size - 1
        lo = np.empty(npat, np.uint32)
        hi = np.empty(npat, np.uint32) if want_hi else None
        pos = np.empty(npat, np.uint32)
        _check(lib().sst_multi_sa_search(self._h, _ptr(flat), _ptr(off), npat, mode, _ptr(lo), _ptr(hi), _ptr(pos)))
        return lo, hi, pos


# ----------------------------------------------------------------------------------------------
# suffix arrays (suffix-array-searching/src/sa_search.rs)
# ----------------------------------------------------------------------------------------------
def pack_patterns(pats):
    """list of bytes-like -> (flat uint8 array, uint64 offsets[npat+1])."""
    lens = np.fromiter((len(p) for p in pats), dtype=np.uint64, count=len(pats))
    off = np.zeros(len(pats) + 1, np.uint64)
    np.cumsum(lens, out=off[1:])
    flat = np.zeros(int(off[-1]), np.uint8)
    if len(pats) and off[-1]:
        flat[:] = np.frombuffer(b"".join(bytes(p) for p in pats), np.uint8)
    return flat, off


class SaNaive:
    """SaNaive (sa_search.rs:11-57) / experiments::SA: text + suffix array resident on one GPU."""

    def __init__(self, handle):
        if not handle:
            _raise()
        self._h = C.c_void_p(handle)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _lib is not None:
            _lib.sst_sa_free(h)
            self._h = None

    @classmethod
    def build(cls, text, device=0):
        """SaNaive::build(t): constructs the suffix array on the GPU (replaces the libsais call, sa_search.rs:33)."""
        if _is_torch(text) and text.is_cuda:
            assert text.is_contiguous() and text.element_size() == 1
            return cls(lib().sst_sa_build_device(C.c_void_p(text.data_ptr()), text.numel(), text.device.index))
        t = np.ascontiguousarray(text, np.uint8)
        return cls(lib().sst_sa_build(_ptr(t), t.size, device))

    @classmethod
    def from_parts(cls, text, sa, device=0):
        t = np.ascontiguousarray(text, np.uint8)
        s = _host_u32(sa)
        assert s.size == t.size
        return cls(lib().sst_sa_from_parts(_ptr(t), t.size, _ptr(s), device))

    def __len__(self):
        return lib().sst_sa_len(self._h)

    @property
    def sa(self) -> np.ndarray:
        out = np.empty(len(self), np.uint32)
        _check(lib().sst_sa_get(self._h, _ptr(out)))
        return out

    def gather(self, positions) -> np.ndarray:
        """sa[positions] (0xffffffff beyond the end) without copying the whole array."""
        p = np.ascontiguousarray(positions, np.uint64)
        out = np.empty(p.size, np.uint32)
        _check(lib().sst_sa_gather(self._h, _ptr(p), p.size, _ptr(out)))
        return out

    def check(self) -> int:
        """Adjacent-suffix strict order (sa_search.rs:36-38): number of violations."""
        v = C.c_uint64(0)
        _check(lib().sst_sa_check(self._h, C.byref(v)))
        return v.value

    def search(self, flat, off, mode=SA_BINARY, want_hi=True):
        """Batched binary_search (sa_search.rs:98-112). Returns (lo, hi, pos); pos = sa[lo] is the reference's return value."""
        flat = np.ascontiguousarray(flat, np.uint8)
        off = np.ascontiguousarray(off, np.uint64)
        npat = off.size - 1
        lo = np.empty(npat, np.uint32)
        hi = np.empty(npat, np.uint32) if want_hi else None
        pos = np.empty(npat, np.uint32)
        _check(lib().sst_sa_search(self._h, _ptr(flat), _ptr(off), npat, mode, _ptr(lo), _ptr(hi), _ptr(pos)))
        return lo, hi, pos

    def search_probes(self, flat, off):
        """The reference's loop itself (plain binary search over [0, n)) with its probe counter `cnt`
        (sa_search.rs:98-112): returns (sa[l], iterations) per pattern."""
        flat = np.ascontiguousarray(flat, np.uint8)
        off = np.ascontiguousarray(off, np.uint64)
        npat = off.size - 1
        pos = np.empty(npat, np.uint32)
        probes = np.empty(npat, np.uint32)
        _check(lib().sst_sa_search_probes(self._h, _ptr(flat), _ptr(off), npat, _ptr(pos), _ptr(probes)))
        return pos, probes

    def binary_search(self, q: bytes) -> int:
        """binary_search(sa, q, cnt) -> sa[l] (sa_search.rs:98-112)."""
        flat, off = pack_patterns([q])
        return int(self.search(flat, off, SA_BINARY, want_hi=False)[2][0])
