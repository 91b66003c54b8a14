"""ctypes binding of libe2i.so (include/e2i.h) -- plumbing for tests, bench.py and the multi-GPU driver.

Host-side mirror of the reference's in-process interface for the hot path: `Index` stands for
`dna_bwt_t` (/root/reference/internal/dna_bwt.hpp:24-420), `navigate` for navigate_one_bwt /
navigate_two_bwts (ebwt2InDel.cpp:555-831), `call` for the cluster scan + find_variants
(:840-1096, 1395-1445, 1510-1560, 1609-1655), `snp_format` for to_file (:1149-1330) and `run` for
run_one_dataset / run_two_datasets / run_two_datasets_da (:1344-1674).

There is no CPU fallback: if libe2i.so is missing or no CUDA device is present every compute
entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("E2I_LIB") or os.path.join(HERE, "libe2i.so")

E2I_OK, E2I_ERR_CUDA, E2I_ERR_SYMBOL, E2I_ERR_ARG, E2I_ERR_MEMORY, E2I_ERR_IO = range(6)


class E2iError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libe2i error {code}: {msg}")
        self.code = code


class Params(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("k_left", "k_right", "K", "max_gap", "max_snvs", "mcov_out",
                                         "complexity", "max_variants_per_position", "term")]


_U64_FIELDS = ("leaves", "nodes", "lcp_values", "lcp_values_leaves", "n_min", "da_values", "n_clusters",
               "clust_size", "events", "clusters_out")
_WORK_FIELDS = ("rank_leaves", "rank_nodes", "rank_call", "bit_updates", "candidates", "levels_leaves",
                "levels_nodes", "max_frontier")
_MS_FIELDS = ("ms_index", "ms_leaves", "ms_nodes", "ms_call", "ms_h2d", "ms_d2h")
_IO_FIELDS = ("kernel_launches", "h2d_bytes", "d2h_bytes")
_WALL_FIELDS = ("ms_format", "ms_wall")


class Stats(C.Structure):
    _fields_ = ([(k, C.c_uint64) for k in _U64_FIELDS] + [("clust_sizes", C.c_uint64 * 201)] +
                [(k, C.c_uint64) for k in _WORK_FIELDS] + [(k, C.c_double) for k in _MS_FIELDS] +
                [(k, C.c_uint64) for k in _IO_FIELDS] + [(k, C.c_double) for k in _WALL_FIELDS] +
                [("da_values_leaves", C.c_uint64)])

    def as_dict(self):
        d = {k: int(getattr(self, k)) for k in _U64_FIELDS + _WORK_FIELDS + _IO_FIELDS}
        d.update({k: float(getattr(self, k)) for k in _MS_FIELDS + _WALL_FIELDS})
        d["clust_sizes"] = list(self.clust_sizes)
        d["da_values_leaves"] = int(self.da_values_leaves)
        return d


class CallRec(C.Structure):
    _fields_ = [("begin", C.c_uint64), ("end", C.c_uint64), ("n0", C.c_uint8), ("n1", C.c_uint8),
                ("right_len", C.c_uint8), ("has_right", C.c_uint8), ("support", C.c_int32 * 8)]


CALL_REC_DTYPE = np.dtype([("begin", "<u8"), ("end", "<u8"), ("n0", "u1"), ("n1", "u1"), ("right_len", "u1"),
                           ("has_right", "u1"), ("support", "<i4", (8,))], align=True)
assert CALL_REC_DTYPE.itemsize == C.sizeof(CallRec) == 56

# every symbol include/e2i.h declares (tests check that the library exports all of them)
SYMBOLS = (
    "e2i_last_error", "e2i_version", "e2i_params_default", "e2i_params_resolve", "e2i_create", "e2i_destroy",
    "e2i_set_frontier_budget", "e2i_trim", "e2i_stream", "e2i_host_alloc", "e2i_host_free", "e2i_index_build", "e2i_index_build_device",
    "e2i_index_slice_align", "e2i_index_super_count", "e2i_index_alloc", "e2i_index_slice_count", "e2i_index_slice_super",
    "e2i_index_slice_pack", "e2i_index_finish", "e2i_index_device", "e2i_index_free", "e2i_index_size", "e2i_index_F", "e2i_index_bytes", "e2i_rank_batch", "e2i_access_batch",
    "e2i_fl_batch", "e2i_rank_batch_device", "e2i_da_load", "e2i_da_load_device", "e2i_bits_fetch",
    "e2i_bits_size", "e2i_bits_free", "e2i_navigate", "e2i_navigate_shard", "e2i_lcpbits_fetch",
    "e2i_lcpbits_device", "e2i_bits_device", "e2i_lcpbits_free", "e2i_call", "e2i_calls_count",
    "e2i_calls_fetch", "e2i_calls_view", "e2i_calls_free", "e2i_snp_format", "e2i_snp_format_gpu", "e2i_call_snp", "e2i_call_device", "e2i_calls_clusters", "e2i_calls_snp", "e2i_calls_snp_device", "e2i_device_free", "e2i_snp_count", "e2i_filter_snp", "e2i_distance", "e2i_buffer_free", "e2i_run",
    "e2i_run_device", "e2i_run_files", "e2i_index_build_file", "e2i_da_load_file", "e2i_index_save", "e2i_index_load", "e2i_ebwt_build", "e2i_run_multi", "e2i_enable_peers", "e2i_or_allreduce",
    "e2i_navigate_ranged", "e2i_comm_local", "e2i_comm_shm", "e2i_comm_barrier", "e2i_comm_free",
)

_lib = None


def lib():
    """Load libe2i.so (built in-tree by `make` / __graft_entry__.build()).  Fails loudly if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise E2iError(E2I_ERR_IO, f"{LIB_PATH} is missing: run `make` (there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, u64, u8p, u64p = C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p
    PP, PS = C.POINTER(Params), C.POINTER(Stats)
    sig = {
        "e2i_last_error": (C.c_char_p, []),
        "e2i_version": (C.c_char_p, []),
        "e2i_params_default": (None, [PP]),
        "e2i_params_resolve": (None, [PP]),
        "e2i_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
        "e2i_destroy": (None, [vp]),
        "e2i_set_frontier_budget": (C.c_int, [vp, u64]),
        "e2i_trim": (C.c_int, [vp]),
        "e2i_stream": (C.c_void_p, [vp]),
        "e2i_host_alloc": (C.c_int, [u64, C.POINTER(vp)]),
        "e2i_host_free": (None, [vp]),
        "e2i_index_build": (C.c_int, [vp, u8p, u64, C.c_uint8, C.POINTER(vp), C.POINTER(u64)]),
        "e2i_index_build_device": (C.c_int, [vp, u8p, u64, C.c_uint8, C.POINTER(vp), C.POINTER(u64)]),
        "e2i_index_slice_align": (u64, []),
        "e2i_index_super_count": (u64, [vp]),
        "e2i_index_alloc": (C.c_int, [vp, u64, C.c_uint8, u64, C.POINTER(vp)]),
        "e2i_index_slice_count": (C.c_int, [vp, vp, u8p, u64, u64, u64, u64p, C.POINTER(u64)]),
        "e2i_index_slice_super": (C.c_int, [vp, vp, u64p, u64p]),
        "e2i_index_slice_pack": (C.c_int, [vp, vp, u8p, u64p, u64p]),
        "e2i_index_finish": (C.c_int, [vp, u64p]),
        "e2i_index_device": (C.c_int, [vp, C.POINTER(vp), C.POINTER(u64)]),
        "e2i_index_free": (None, [vp]),
        "e2i_index_size": (u64, [vp]),
        "e2i_index_F": (C.c_int, [vp, u64p]),
        "e2i_index_bytes": (u64, [vp]),
        "e2i_rank_batch": (C.c_int, [vp, vp, u64p, u64, u64p]),
        "e2i_access_batch": (C.c_int, [vp, vp, u64p, u64, u8p]),
        "e2i_fl_batch": (C.c_int, [vp, vp, u64p, u64, u64p]),
        "e2i_rank_batch_device": (C.c_int, [vp, vp, u64p, u64, u64p, C.POINTER(C.c_float)]),
        "e2i_da_load": (C.c_int, [vp, u8p, u64, C.POINTER(vp)]),
        "e2i_da_load_device": (C.c_int, [vp, u8p, u64, C.POINTER(vp)]),
        "e2i_bits_fetch": (C.c_int, [vp, vp, u64p, u64]),
        "e2i_bits_size": (u64, [vp]),
        "e2i_bits_free": (None, [vp]),
        "e2i_navigate": (C.c_int, [vp, vp, vp, PP, C.POINTER(vp), C.POINTER(vp), PS]),
        "e2i_navigate_shard": (C.c_int, [vp, vp, vp, PP, C.c_int, C.c_int, C.POINTER(vp), C.POINTER(vp), PS]),
        "e2i_lcpbits_fetch": (C.c_int, [vp, vp, u64p, u64p]),
        "e2i_lcpbits_device": (C.c_int, [vp, C.POINTER(vp), C.POINTER(u64), C.POINTER(vp), C.POINTER(u64)]),
        "e2i_bits_device": (C.c_int, [vp, C.POINTER(vp), C.POINTER(u64)]),
        "e2i_lcpbits_free": (None, [vp]),
        "e2i_call": (C.c_int, [vp, vp, vp, vp, vp, PP, u64, u64, C.POINTER(vp), PS]),
        "e2i_calls_count": (u64, [vp]),
        "e2i_calls_fetch": (C.c_int, [vp, vp, vp, vp, u64, C.POINTER(u64)]),
        "e2i_calls_view": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(u64)]),
        "e2i_calls_free": (None, [vp]),
        "e2i_snp_format": (C.c_int, [vp, vp, vp, u64, PP, C.c_int, u64, C.POINTER(vp), C.POINTER(C.c_size_t), PS]),
        "e2i_snp_format_gpu": (C.c_int, [vp, vp, vp, vp, u64, PP, C.c_int, u64, C.POINTER(vp), C.POINTER(C.c_size_t), PS]),
        "e2i_call_snp": (C.c_int, [vp, vp, vp, vp, vp, PP, u64, u64, u64, C.POINTER(vp), C.POINTER(C.c_size_t), PS]),
        "e2i_call_device": (C.c_int, [vp, vp, vp, vp, vp, PP, u64, u64, C.POINTER(vp), PS]),
        "e2i_calls_clusters": (C.c_int, [vp, PP, C.POINTER(u64)]),
        "e2i_calls_snp": (C.c_int, [vp, PP, u64, C.POINTER(vp), C.POINTER(C.c_size_t), PS]),
        "e2i_calls_snp_device": (C.c_int, [vp, PP, u64, C.POINTER(vp), C.POINTER(u64), PS]),
        "e2i_device_free": (None, [vp, vp]),
        "e2i_snp_count": (C.c_int, [vp, vp, vp, u64, PP, C.c_int, C.POINTER(u64)]),
        "e2i_filter_snp": (C.c_int, [C.c_char_p, C.c_size_t, C.c_int32, C.c_int32, C.POINTER(vp), C.POINTER(C.c_size_t)]),
        "e2i_distance": (None, [C.c_char_p, C.c_char_p, C.c_int32, C.c_int32, C.POINTER(C.c_int32)]),
        "e2i_buffer_free": (None, [vp]),
        "e2i_run": (C.c_int, [vp, u8p, u64, u8p, u64, u8p, PP, C.POINTER(vp), C.POINTER(C.c_size_t), PS]),
        "e2i_run_device": (C.c_int, [vp, u8p, u64, u8p, u64, u8p, PP, C.POINTER(vp), C.POINTER(C.c_size_t), PS]),
        "e2i_run_files": (C.c_int, [vp, C.c_char_p, C.c_char_p, C.c_char_p, PP, C.POINTER(vp), C.POINTER(C.c_size_t), PS, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]),
        "e2i_index_build_file": (C.c_int, [vp, C.c_char_p, C.c_uint8, C.POINTER(vp), C.POINTER(u64)]),
        "e2i_da_load_file": (C.c_int, [vp, C.c_char_p, u64, C.POINTER(vp)]),
        "e2i_index_save": (C.c_int, [vp, C.c_char_p]),
        "e2i_index_load": (C.c_int, [vp, C.c_char_p, C.POINTER(vp)]),
        "e2i_ebwt_build": (C.c_int, [vp, u8p, u64, C.c_uint32, u64, C.c_uint8, u8p, u8p]),
        "e2i_run_multi": (C.c_int, [C.POINTER(C.c_int), C.c_int, u8p, u64, u8p, u64, u8p, PP, u64, C.POINTER(vp), C.POINTER(C.c_size_t), PS]),
        "e2i_enable_peers": (C.c_int, [C.POINTER(vp), C.c_int]),
        "e2i_navigate_ranged": (C.c_int, [vp, vp, vp, vp, PP, C.POINTER(vp), C.POINTER(vp), PS]),
        "e2i_comm_local": (C.c_int, [C.c_int, C.POINTER(vp)]),
        "e2i_comm_shm": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.POINTER(vp)]),
        "e2i_comm_barrier": (None, [vp]),
        "e2i_comm_free": (None, [vp]),
        "e2i_or_allreduce": (C.c_int, [vp, C.POINTER(vp), C.c_int, C.c_int, u64]),
    }
    assert set(sig) == set(SYMBOLS)
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _await_producer(t):
    """The library works on the context's own stream.  A torch CUDA tensor handed in may still be being written by
    kernels queued on torch's current stream: wait for them (a no-op when the stream is idle)."""
    if hasattr(t, "is_cuda") and t.is_cuda:
        import torch
        torch.cuda.current_stream(t.device).synchronize()


def _check(rc: int):
    if rc != E2I_OK:
        raise E2iError(rc, lib().e2i_last_error().decode(errors="replace"))


def default_params(**kw) -> Params:
    p = Params()
    lib().e2i_params_default(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def resolve_params(p: Params) -> Params:
    lib().e2i_params_resolve(C.byref(p))
    return p


def distance(a: str, b: str, max_gap: int):
    out = (C.c_int32 * 2)()
    lib().e2i_distance(a.encode(), b.encode(), len(a), max_gap, out)
    return int(out[0]), int(out[1])


def _host_u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data


class SnpText:
    """The .snp text returned by the library (malloc'ed by libe2i, freed with the object).
    `bytes(x)` / `x.tobytes()` copies it out; `len(x)` and `x.view()` do not."""

    def __init__(self, ptr, length: int):
        self._ptr, self._len = ptr, int(length)

    def __len__(self):
        return self._len

    def view(self) -> memoryview:
        return memoryview((C.c_ubyte * self._len).from_address(self._ptr.value)).cast("B") if self._len else memoryview(b"")

    def tobytes(self) -> bytes:
        return C.string_at(self._ptr, self._len) if self._len else b""

    __bytes__ = tobytes

    def __eq__(self, other):
        if isinstance(other, SnpText):
            other = other.view()
        elif isinstance(other, np.ndarray):
            other = memoryview(np.ascontiguousarray(other)).cast("B")
        return self.view() == other

    def __del__(self):
        if getattr(self, "_ptr", None) is not None and _lib is not None:
            _lib.e2i_buffer_free(self._ptr)
            self._ptr = None


class Context:
    """One device context (e2i_ctx): a stream, scratch and the frame allocator.  One per GPU / thread."""

    def __init__(self, device: int = 0, frontier_bytes: int = 0):
        h = C.c_void_p()
        _check(lib().e2i_create(device, C.byref(h)))
        self.h = h
        self.device = device
        if frontier_bytes:
            _check(lib().e2i_set_frontier_budget(self.h, frontier_bytes))

    def close(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.e2i_destroy(self.h)
        self.h = None

    def trim(self):
        """Give the device memory cached by the library back to the driver."""
        _check(lib().e2i_trim(self.h))

    @property
    def stream_ptr(self) -> int:
        """cudaStream_t of the context (wrap with torch.cuda.ExternalStream to record events on it)."""
        return int(lib().e2i_stream(self.h))

    def __del__(self):
        self.close()

    # ---- index ----
    def index(self, bwt, term: int = ord("#")) -> "Index":
        """bwt: numpy uint8 array / bytes (host) or a torch CUDA uint8 tensor (device, 16-byte aligned)."""
        h, bad = C.c_void_p(), C.c_uint64(0)
        if hasattr(bwt, "data_ptr") and bwt.is_cuda:
            _await_producer(bwt)
            rc = lib().e2i_index_build_device(self.h, bwt.data_ptr(), bwt.numel(), term, C.byref(h), C.byref(bad))
        else:
            keep, ptr = _host_u8(np.frombuffer(bwt, dtype=np.uint8) if isinstance(bwt, (bytes, bytearray)) else bwt)
            rc = lib().e2i_index_build(self.h, ptr, len(keep), term, C.byref(h), C.byref(bad))
        if rc == E2I_ERR_SYMBOL:
            raise ValueError(f"forbidden symbol at position {bad.value}")
        _check(rc)
        return Index(self, h)

    def index_file(self, path: str, term: int = ord("#")) -> "Index":
        """Streaming ingest of an ASCII eBWT file (reader thread -> page-locked ring -> copy stream -> counting)."""
        h, bad = C.c_void_p(), C.c_uint64(0)
        rc = lib().e2i_index_build_file(self.h, os.fsencode(path), term, C.byref(h), C.byref(bad))
        if rc == E2I_ERR_SYMBOL:
            raise ValueError(f"forbidden symbol at position {bad.value}")
        _check(rc)
        return Index(self, h)

    def index_load(self, path: str) -> "Index":
        """Packed-index sidecar written by Index.save()."""
        h = C.c_void_p()
        _check(lib().e2i_index_load(self.h, os.fsencode(path), C.byref(h)))
        return Index(self, h)

    def document_array_file(self, path: str, n: int) -> "Bits":
        h = C.c_void_p()
        _check(lib().e2i_da_load_file(self.h, os.fsencode(path), n, C.byref(h)))
        return Bits(self, h)

    def run_files(self, path1: str, path2: str | None = None, path_da: str | None = None, params: Params | None = None, copy: bool = True):
        """Whole path from files (e2i_run_files).  Returns (.snp bytes or SnpText, Stats)."""
        p = params or default_params()
        st = Stats()
        out, ln = C.c_void_p(), C.c_size_t()
        bad = (C.c_uint64 * 2)()
        rc = lib().e2i_run_files(self.h, os.fsencode(path1), os.fsencode(path2) if path2 else None, os.fsencode(path_da) if path_da else None,
                                 C.byref(p), C.byref(out), C.byref(ln), C.byref(st), None, None, bad)
        if rc == E2I_ERR_SYMBOL:
            raise ValueError(lib().e2i_last_error().decode())
        _check(rc)
        text = SnpText(out, ln.value)
        return (text.tobytes() if copy else text), st

    def ebwt_build(self, reads: np.ndarray, second_from: int | None = None, term: int = ord("#")):
        """eBWT of an (m, L) uint8 matrix of reads on the GPU (e2i_ebwt_build).  With second_from the reads from that
        row on belong to the second individual and the ASCII '0'/'1' document array is returned as well."""
        reads = np.ascontiguousarray(reads, dtype=np.uint8)
        m, L = reads.shape
        bwt = np.empty(m * (L + 1), dtype=np.uint8)
        da = np.empty(m * (L + 1), dtype=np.uint8) if second_from is not None else None
        rc = lib().e2i_ebwt_build(self.h, reads.ctypes.data, m, L, m if second_from is None else second_from, term,
                                  bwt.ctypes.data, da.ctypes.data if da is not None else None)
        if rc == E2I_ERR_SYMBOL:
            raise ValueError(lib().e2i_last_error().decode())
        _check(rc)
        return bwt if da is None else (bwt, da)

    def index_alloc(self, n: int, term: int = ord("#"), tile_multiple: int = 1) -> "Index":
        """Empty index for slice-wise construction (Index.slice_count / slice_super / slice_pack / finish)."""
        h = C.c_void_p()
        _check(lib().e2i_index_alloc(self.h, n, term, tile_multiple, C.byref(h)))
        return Index(self, h)

    def document_array(self, da) -> "Bits":
        h = C.c_void_p()
        if hasattr(da, "data_ptr") and da.is_cuda:
            _await_producer(da)
            _check(lib().e2i_da_load_device(self.h, da.data_ptr(), da.numel(), C.byref(h)))
        else:
            keep, ptr = _host_u8(da)
            _check(lib().e2i_da_load(self.h, ptr, len(keep), C.byref(h)))
        return Bits(self, h)

    # ---- phases 2+3 ----
    def navigate(self, b1: "Index", b2: "Index | None" = None, params: Params | None = None,
                 shard: int = 0, n_shards: int = 1, stats: Stats | None = None):
        p = params or default_params()
        st = stats if stats is not None else Stats()
        lh, dh = C.c_void_p(), C.c_void_p()
        _check(lib().e2i_navigate_shard(self.h, b1.h, b2.h if b2 else None, C.byref(p), shard, n_shards,
                                        C.byref(lh), C.byref(dh) if b2 else None, C.byref(st)))
        return LcpBits(self, lh), (Bits(self, dh) if b2 else None), st

    def navigate_ranged(self, comm, b1: "Index", b2: "Index | None" = None, params: Params | None = None, stats: Stats | None = None):
        """Position-range sharded traversal (e2i_navigate_ranged); comm: a handle made by comm_shm()."""
        p = params or default_params()
        st = stats if stats is not None else Stats()
        lh, dh = C.c_void_p(), C.c_void_p()
        _check(lib().e2i_navigate_ranged(self.h, comm, b1.h, b2.h if b2 else None, C.byref(p), C.byref(lh), C.byref(dh) if b2 else None, C.byref(st)))
        return LcpBits(self, lh), (Bits(self, dh) if b2 else None), st

    # ---- phase 4 ----
    def call_snp(self, b1, b2, da, lcp, params: Params | None = None, pos_begin: int = 0, pos_end: int = 2 ** 64 - 1,
                 first_cluster_nr: int = 1, stats: Stats | None = None, copy: bool = True):
        """Phase 4 + the .snp text of [pos_begin, pos_end), formatted on the device (e2i_call_snp)."""
        p = params or default_params()
        st = stats if stats is not None else Stats()
        out, ln = C.c_void_p(), C.c_size_t()
        _check(lib().e2i_call_snp(self.h, b1.h, b2.h if b2 else None, da.h if da else None, lcp.h, C.byref(p),
                                  pos_begin, pos_end, first_cluster_nr, C.byref(out), C.byref(ln), C.byref(st)))
        text = SnpText(out, ln.value)
        return (text.tobytes() if copy else text), st

    def call_device(self, b1, b2, da, lcp, params: Params | None = None, pos_begin: int = 0,
                    pos_end: int = 2 ** 64 - 1, stats: Stats | None = None) -> "DeviceCalls":
        """Phase 4 on [pos_begin, pos_end) with the records left in HBM (e2i_call_device)."""
        p = params or default_params()
        st = stats if stats is not None else Stats()
        ch = C.c_void_p()
        _check(lib().e2i_call_device(self.h, b1.h, b2.h if b2 else None, da.h if da else None, lcp.h, C.byref(p),
                                     pos_begin, pos_end, C.byref(ch), C.byref(st)))
        return DeviceCalls(self, ch, p, st)

    def snp_format(self, recs, left, right, params: Params, two_samples: bool, first_cluster_nr: int = 1,
                   stats: Stats | None = None):
        """snp_format() by the device formatter (e2i_snp_format_gpu): same arguments, same text."""
        st = stats if stats is not None else Stats()
        recs = np.ascontiguousarray(recs, dtype=CALL_REC_DTYPE)
        left = np.ascontiguousarray(left, dtype=np.uint8)
        right = np.ascontiguousarray(right, dtype=np.uint8)
        out, ln = C.c_void_p(), C.c_size_t()
        _check(lib().e2i_snp_format_gpu(self.h, recs.ctypes.data, left.ctypes.data, right.ctypes.data, len(recs), C.byref(params),
                                        1 if two_samples else 0, first_cluster_nr, C.byref(out), C.byref(ln), C.byref(st)))
        return SnpText(out, ln.value).tobytes(), st

    def call(self, b1, b2, da, lcp, params: Params | None = None, pos_begin: int = 0,
             pos_end: int = 2 ** 64 - 1, stats: Stats | None = None, copy: bool = True):
        """Phase 4 on [pos_begin, pos_end).  Returns (recs, left, right, Stats) as numpy arrays; with
        copy=False they are zero-copy views of the context's page-locked result buffer, valid until
        the next call() on this context."""
        p = params or default_params()
        st = stats if stats is not None else Stats()
        ch = C.c_void_p()
        _check(lib().e2i_call(self.h, b1.h, b2.h if b2 else None, da.h if da else None, lcp.h, C.byref(p),
                              pos_begin, pos_end, C.byref(ch), C.byref(st)))
        if not copy:
            try:
                pr, pl, pt, n = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_uint64()
                _check(lib().e2i_calls_view(ch, C.byref(pr), C.byref(pl), C.byref(pt), C.byref(n)))
                n = int(n.value)

                def view(ptr, nbytes, dtype):
                    if not nbytes:
                        return np.zeros(0, dtype=dtype)
                    return np.frombuffer((C.c_ubyte * nbytes).from_address(ptr.value), dtype=dtype)
                return (view(pr, n * CALL_REC_DTYPE.itemsize, CALL_REC_DTYPE), view(pl, n * 8 * p.k_left, np.uint8),
                        view(pt, n * p.k_right, np.uint8), st)
            finally:
                lib().e2i_calls_free(ch)
        try:
            n = int(lib().e2i_calls_count(ch))
            recs = np.zeros(n, dtype=CALL_REC_DTYPE)
            left = np.zeros(n * 8 * p.k_left, dtype=np.uint8)
            right = np.zeros(n * p.k_right, dtype=np.uint8)
            got = C.c_uint64(0)
            if n:
                _check(lib().e2i_calls_fetch(ch, recs.ctypes.data, left.ctypes.data, right.ctypes.data, n, C.byref(got)))
        finally:
            lib().e2i_calls_free(ch)
        return recs, left, right, st

    # ---- whole path ----
    def run(self, bwt1, bwt2=None, da=None, params: Params | None = None, stats: Stats | None = None,
            copy: bool = True):
        """Whole path.  Inputs all on the host (numpy / bytes / pinned torch CPU tensors) -> e2i_run,
        or all torch CUDA uint8 tensors -> e2i_run_device.  Returns (.snp bytes, Stats); with
        copy=False the text comes back as a SnpText that wraps the library's buffer."""
        p = params or default_params()
        st = stats if stats is not None else Stats()
        on_dev = hasattr(bwt1, "is_cuda") and bwt1.is_cuda
        if on_dev:
            _await_producer(bwt1)
        keep = []

        def arg(x):
            if x is None:
                return None, 0
            if hasattr(x, "data_ptr"):
                return x.data_ptr(), x.numel()
            a, ptr = _host_u8(np.frombuffer(x, dtype=np.uint8) if isinstance(x, (bytes, bytearray)) else x)
            keep.append(a)
            return ptr, len(a)

        p1, n1 = arg(bwt1)
        p2, n2 = arg(bwt2)
        pd, _ = arg(da)
        out, ln = C.c_void_p(), C.c_size_t()
        fn = lib().e2i_run_device if on_dev else lib().e2i_run
        rc = fn(self.h, p1, n1, p2, n2, pd, C.byref(p), C.byref(out), C.byref(ln), C.byref(st))
        if rc == E2I_ERR_SYMBOL:
            raise ValueError(lib().e2i_last_error().decode())
        _check(rc)
        text = SnpText(out, ln.value)
        return (text.tobytes() if copy else text), st


def run_multi(devices, bwt1, bwt2=None, da=None, params: Params | None = None, frontier_bytes: int = 0, copy: bool = True):
    """Whole path on several GPUs of this box from ONE process (e2i_run_multi): host inputs (numpy / bytes).
    Returns (.snp bytes or SnpText, Stats)."""
    p = params or default_params()
    st = Stats()
    keep = []

    def arg(x):
        if x is None:
            return None, 0
        a, ptr = _host_u8(np.frombuffer(x, dtype=np.uint8) if isinstance(x, (bytes, bytearray)) else x)
        keep.append(a)
        return ptr, len(a)
    p1, n1 = arg(bwt1)
    p2, n2 = arg(bwt2)
    pd, _ = arg(da)
    devs = (C.c_int * len(devices))(*devices)
    out, ln = C.c_void_p(), C.c_size_t()
    rc = lib().e2i_run_multi(devs, len(devices), p1, n1, p2, n2, pd, C.byref(p), frontier_bytes, C.byref(out), C.byref(ln), C.byref(st))
    if rc == E2I_ERR_SYMBOL:
        raise ValueError(lib().e2i_last_error().decode())
    _check(rc)
    text = SnpText(out, ln.value)
    return (text.tobytes() if copy else text), st


def comm_shm(name: str, rank: int, world: int):
    """Communicator over a POSIX shared-memory segment for the processes of one box (e2i_comm_shm)."""
    h = C.c_void_p()
    _check(lib().e2i_comm_shm(name.encode(), rank, world, C.byref(h)))
    return h


def comm_free(h):
    if h:
        lib().e2i_comm_free(h)


def filter_snp(snp: bytes, m: int, M: int = 0) -> bytes:
    """Coverage filter of the reference's filter_snp tool on an in-memory .snp text."""
    out, ln = C.c_void_p(), C.c_size_t()
    _check(lib().e2i_filter_snp(snp, len(snp), m, M, C.byref(out), C.byref(ln)))
    return SnpText(out, ln.value).tobytes()


def snp_count(recs: np.ndarray, left: np.ndarray, right: np.ndarray, params: Params, two_samples: bool) -> int:
    """Cluster numbers the records consume (clusters_out of snp_format) without building the text."""
    recs = np.ascontiguousarray(recs, dtype=CALL_REC_DTYPE)
    left = np.ascontiguousarray(left, dtype=np.uint8)
    right = np.ascontiguousarray(right, dtype=np.uint8)
    out = C.c_uint64(0)
    _check(lib().e2i_snp_count(recs.ctypes.data, left.ctypes.data, right.ctypes.data, len(recs), C.byref(params),
                               1 if two_samples else 0, C.byref(out)))
    return int(out.value)


def snp_format(recs: np.ndarray, left: np.ndarray, right: np.ndarray, params: Params, two_samples: bool,
               first_cluster_nr: int = 1, stats: Stats | None = None, copy: bool = True):
    st = stats if stats is not None else Stats()
    recs = np.ascontiguousarray(recs, dtype=CALL_REC_DTYPE)
    left = np.ascontiguousarray(left, dtype=np.uint8)
    right = np.ascontiguousarray(right, dtype=np.uint8)
    out, ln = C.c_void_p(), C.c_size_t()
    _check(lib().e2i_snp_format(recs.ctypes.data, left.ctypes.data, right.ctypes.data, len(recs), C.byref(params),
                                1 if two_samples else 0, first_cluster_nr, C.byref(out), C.byref(ln), C.byref(st)))
    text = SnpText(out, ln.value)
    return (text.tobytes() if copy else text), st


class DeviceCalls:
    """Call records of one position range, resident in HBM (Context.call_device)."""

    def __init__(self, ctx: Context, h, params: Params, stats: Stats):
        self.ctx, self.h, self.params, self.stats = ctx, h, params, stats

    def __len__(self):
        return int(lib().e2i_calls_count(self.h))

    def clusters(self) -> int:
        """Cluster numbers the range consumes (device pass, no text)."""
        out = C.c_uint64(0)
        _check(lib().e2i_calls_clusters(self.h, C.byref(self.params), C.byref(out)))
        return int(out.value)

    def snp(self, first_cluster_nr: int = 1, copy: bool = True):
        out, ln = C.c_void_p(), C.c_size_t()
        _check(lib().e2i_calls_snp(self.h, C.byref(self.params), first_cluster_nr, C.byref(out), C.byref(ln), C.byref(self.stats)))
        text = SnpText(out, ln.value)
        return text.tobytes() if copy else text

    def snp_device(self, first_cluster_nr: int = 1):
        """(device pointer, length) of the text, left in device memory; release with free_device()."""
        out, ln = C.c_void_p(), C.c_uint64(0)
        _check(lib().e2i_calls_snp_device(self.h, C.byref(self.params), first_cluster_nr, C.byref(out), C.byref(ln), C.byref(self.stats)))
        return out, int(ln.value)

    def free_device(self, ptr):
        if ptr and self.ctx.h:
            lib().e2i_device_free(self.ctx.h, ptr)

    def close(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.e2i_calls_free(self.h)
        self.h = None

    def __del__(self):
        self.close()


class Index:
    def __init__(self, ctx: Context, h):
        self.ctx, self.h = ctx, h

    def close(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.e2i_index_free(self.h)
        self.h = None

    def __del__(self):
        self.close()

    @property
    def n(self) -> int:
        return int(lib().e2i_index_size(self.h))

    @property
    def nbytes(self) -> int:
        return int(lib().e2i_index_bytes(self.h))

    def F(self) -> np.ndarray:
        out = np.zeros(4, dtype=np.uint64)
        _check(lib().e2i_index_F(self.h, out.ctypes.data))
        return out

    def save(self, path: str):
        """Write the packed index (as it lies in HBM) to `path`; Context.index_load reads it back."""
        _check(lib().e2i_index_save(self.h, os.fsencode(path)))

    # ---- slice-wise construction (multi-GPU) ----
    def slice_count(self, dev_slice, begin: int, n_tiles: int) -> np.ndarray:
        """Counts of one slice; n_tiles = tiles the slice owns (0 = empty slice), see distributed.index_slices."""
        counts = np.zeros(4, dtype=np.uint64)
        bad = C.c_uint64(0)
        rc = lib().e2i_index_slice_count(self.ctx.h, self.h, dev_slice.data_ptr() if dev_slice.numel() else None, begin,
                                         dev_slice.numel(), n_tiles, counts.ctypes.data, C.byref(bad))
        if rc == E2I_ERR_SYMBOL:
            raise ValueError(f"forbidden symbol at position {bad.value}")
        _check(rc)
        return counts

    @property
    def n_super(self) -> int:
        """Entries of the superblock table (one per 2^16 symbols)."""
        return int(lib().e2i_index_super_count(self.h))

    def slice_super(self, before: np.ndarray, n_super: int | None = None) -> np.ndarray:
        n_super = self.n_super if n_super is None else n_super
        before = np.ascontiguousarray(before, dtype=np.uint64)
        out = np.zeros(n_super * 4, dtype=np.uint64)
        _check(lib().e2i_index_slice_super(self.ctx.h, self.h, before.ctypes.data, out.ctypes.data))
        return out

    def slice_pack(self, dev_slice, before: np.ndarray, super_table: np.ndarray):
        before = np.ascontiguousarray(before, dtype=np.uint64)
        super_table = np.ascontiguousarray(super_table, dtype=np.uint64)
        _check(lib().e2i_index_slice_pack(self.ctx.h, self.h, dev_slice.data_ptr() if dev_slice.numel() else None,
                                          before.ctypes.data, super_table.ctypes.data))

    def finish(self, totals: np.ndarray):
        totals = np.ascontiguousarray(totals, dtype=np.uint64)
        _check(lib().e2i_index_finish(self.h, totals.ctypes.data))

    def device_blocks(self):
        """(device pointer, bytes) of the block array, for the all-gather of the slices."""
        p, b = C.c_void_p(), C.c_uint64()
        _check(lib().e2i_index_device(self.h, C.byref(p), C.byref(b)))
        return int(p.value), int(b.value)

    def rank4(self, pos) -> np.ndarray:
        pos = np.ascontiguousarray(pos, dtype=np.uint64)
        out = np.zeros((len(pos), 4), dtype=np.uint64)
        _check(lib().e2i_rank_batch(self.ctx.h, self.h, pos.ctypes.data, len(pos), out.ctypes.data))
        return out

    def access(self, pos) -> np.ndarray:
        pos = np.ascontiguousarray(pos, dtype=np.uint64)
        out = np.zeros(len(pos), dtype=np.uint8)
        _check(lib().e2i_access_batch(self.ctx.h, self.h, pos.ctypes.data, len(pos), out.ctypes.data))
        return out

    def FL(self, pos) -> np.ndarray:
        pos = np.ascontiguousarray(pos, dtype=np.uint64)
        out = np.zeros(len(pos), dtype=np.uint64)
        _check(lib().e2i_fl_batch(self.ctx.h, self.h, pos.ctypes.data, len(pos), out.ctypes.data))
        return out

    def rank4_device(self, dev_pos, dev_out4) -> float:
        """Kernel-only batched rank on device tensors (int64 views of u64); returns milliseconds."""
        ms = C.c_float(0)
        _check(lib().e2i_rank_batch_device(self.ctx.h, self.h, dev_pos.data_ptr(), dev_pos.numel(),
                                           dev_out4.data_ptr(), C.byref(ms)))
        return float(ms.value)


class Bits:
    def __init__(self, ctx: Context, h):
        self.ctx, self.h = ctx, h

    def close(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.e2i_bits_free(self.h)
        self.h = None

    def __del__(self):
        self.close()

    @property
    def n(self) -> int:
        return int(lib().e2i_bits_size(self.h))

    def fetch(self) -> np.ndarray:
        nw = (self.n + 63) // 64
        out = np.zeros(nw, dtype=np.uint64)
        _check(lib().e2i_bits_fetch(self.ctx.h, self.h, out.ctypes.data, nw))
        return out

    def device_words(self):
        """(device pointer, number of uint32 words) of the packed bits, for the cross-GPU OR-reduce."""
        p, w = C.c_void_p(), C.c_uint64()
        _check(lib().e2i_bits_device(self.h, C.byref(p), C.byref(w)))
        return int(p.value), int(w.value)


class LcpBits:
    def __init__(self, ctx: Context, h):
        self.ctx, self.h = ctx, h

    def close(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.e2i_lcpbits_free(self.h)
        self.h = None

    def __del__(self):
        self.close()

    def fetch(self, n: int):
        thr = np.zeros((2 * n + 63) // 64, dtype=np.uint64)
        mn = np.zeros((n + 63) // 64, dtype=np.uint64)
        _check(lib().e2i_lcpbits_fetch(self.ctx.h, self.h, thr.ctypes.data, mn.ctypes.data))
        return thr, mn

    def device_words(self):
        pt, wt, pm, wm = C.c_void_p(), C.c_uint64(), C.c_void_p(), C.c_uint64()
        _check(lib().e2i_lcpbits_device(self.h, C.byref(pt), C.byref(wt), C.byref(pm), C.byref(wm)))
        return (int(pt.value), int(wt.value)), (int(pm.value), int(wm.value))
