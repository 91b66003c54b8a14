"""ctypes binding of the CPU oracle (liboracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product never does.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_BIN = os.path.join(HERE, "_ref", "ebwt2InDel")


class Params(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("k_left", "k_right", "K", "max_gap", "max_snvs", "mcov_out",
                                         "complexity", "max_variants_per_position", "term")]


class Stats(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in (
        "leaves", "nodes", "lcp_values", "lcp_values_leaves", "n_min", "da_values", "max_stack_leaves",
        "max_stack_nodes", "n_clusters", "clust_size", "events", "clusters_out", "rank_leaves",
        "rank_nodes", "rank_call", "lcp_border_updates")] + [("clust_sizes", C.c_uint64 * 201)]

    def as_dict(self):
        d = {k: int(getattr(self, k)) for k, _ in self._fields_ if k != "clust_sizes"}
        d["clust_sizes"] = list(self.clust_sizes)
        return d


_lib = None


def build():
    subprocess.run(["make", "-s", "-C", HERE], check=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        u8p, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint64)
        L.orc_bwt_build.restype = C.c_void_p
        L.orc_bwt_build.argtypes = [u8p, C.c_uint64, C.c_uint8, u64p]
        L.orc_bwt_free.argtypes = [C.c_void_p]
        L.orc_bwt_size.restype = C.c_uint64
        L.orc_bwt_size.argtypes = [C.c_void_p]
        L.orc_bwt_F.argtypes = [C.c_void_p, u64p]
        L.orc_rank4_batch.argtypes = [C.c_void_p, u64p, C.c_uint64, u64p]
        L.orc_access.restype = C.c_uint8
        L.orc_access.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_select.restype = C.c_uint64
        L.orc_select.argtypes = [C.c_void_p, C.c_uint64, C.c_uint8]
        L.orc_FL.restype = C.c_uint64
        L.orc_FL.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_distance.argtypes = [C.c_char_p, C.c_char_p, C.c_int32, C.c_int32, C.POINTER(C.c_int32)]
        L.orc_navigate_one.argtypes = [C.c_void_p, C.POINTER(Params), u64p, u64p, C.POINTER(Stats)]
        L.orc_navigate_two.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(Params), u64p, u64p, u64p, C.POINTER(Stats)]
        L.orc_call.argtypes = [C.c_void_p, C.c_void_p, u64p, u64p, u64p, C.POINTER(Params),
                               C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(Stats)]
        L.orc_run.argtypes = [u8p, C.c_uint64, u8p, C.c_uint64, u8p, C.POINTER(Params),
                              C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(Stats)]
        L.orc_params_default.argtypes = [C.POINTER(Params)]
        L.orc_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def default_params(**kw) -> Params:
    p = Params()
    lib().orc_params_default(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data_as(C.POINTER(C.c_uint8))


def _u64(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint64))


class Bwt:
    def __init__(self, ascii_bytes, term=ord("#")):
        self._keep, ptr = _u8(ascii_bytes)
        bad = C.c_uint64(0)
        self.h = lib().orc_bwt_build(ptr, len(self._keep), term, C.byref(bad))
        if not self.h:
            raise ValueError(f"forbidden symbol at position {bad.value}")
        self.n = len(self._keep)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_bwt_free(self.h)
            self.h = None

    def F(self):
        out = np.zeros(4, dtype=np.uint64)
        lib().orc_bwt_F(self.h, _u64(out))
        return out

    def rank4(self, pos):
        pos = np.ascontiguousarray(pos, dtype=np.uint64)
        out = np.zeros((len(pos), 4), dtype=np.uint64)
        lib().orc_rank4_batch(self.h, _u64(pos), len(pos), _u64(out))
        return out


def distance(a: str, b: str, max_gap: int):
    out = (C.c_int32 * 2)()
    lib().orc_distance(a.encode(), b.encode(), len(a), max_gap, out)
    return int(out[0]), int(out[1])


def navigate_one(bwt: Bwt, p: Params):
    n = bwt.n
    thr = np.zeros((2 * n + 63) // 64 + 1, dtype=np.uint64)
    mn = np.zeros((n + 63) // 64 + 1, dtype=np.uint64)
    st = Stats()
    rc = lib().orc_navigate_one(bwt.h, C.byref(p), _u64(thr), _u64(mn), C.byref(st))
    assert rc == 0
    return thr, mn, st


def navigate_two(b1: Bwt, b2: Bwt, p: Params):
    n = b1.n + b2.n
    thr = np.zeros((2 * n + 63) // 64 + 1, dtype=np.uint64)
    mn = np.zeros((n + 63) // 64 + 1, dtype=np.uint64)
    da = np.zeros((n + 63) // 64 + 1, dtype=np.uint64)
    st = Stats()
    rc = lib().orc_navigate_two(b1.h, b2.h, C.byref(p), _u64(thr), _u64(mn), _u64(da), C.byref(st))
    assert rc == 0
    return thr, mn, da, st


def run(bwt1, bwt2=None, da_ascii=None, params: Params | None = None):
    """Whole path on the oracle.  Returns (.snp bytes, Stats)."""
    p = params or default_params()
    k1, p1 = _u8(bwt1)
    k2, p2 = _u8(bwt2) if bwt2 is not None else (None, None)
    k3, p3 = _u8(da_ascii) if da_ascii is not None else (None, None)
    out, ln, st = C.c_void_p(), C.c_size_t(), Stats()
    rc = lib().orc_run(p1, len(k1), p2, len(k2) if k2 is not None else 0, p3, C.byref(p),
                       C.byref(out), C.byref(ln), C.byref(st))
    if rc:
        raise RuntimeError(f"oracle failed: {rc}")
    snp = C.string_at(out, ln.value)
    lib().orc_free(out)
    return snp, st


# ---- the compiled reference (oracle/_ref/ebwt2InDel) ---------------------------------------

_COUNTERS = {
    "leaves": r"Processed (\d+) suffix-tree leaves",
    "nodes": r"Processed (\d+) suffix-tree nodes",
    "lcp_values": r"Computed (\d+)/\d+ LCP values",
    "lcp_values_leaves": r"Computed (\d+)/\d+ LCP threshold values",
    "n_min": r"Found (\d+) LCP minima",
    "da_values": r"Computed (\d+)/\d+ DA values",
    "n_clusters": r"Analyzed (\d+) clusters",
}


def ref_available() -> bool:
    return os.access(REF_BIN, os.X_OK)


def run_ref(bwt1, bwt2=None, da_ascii=None, flags=(), workdir=None, timed=False):
    """Run the compiled reference on in-memory inputs.  Returns (.snp bytes, counters[, phase seconds])."""
    with tempfile.TemporaryDirectory(dir=workdir) as d:
        f1 = os.path.join(d, "a.ebwt")
        np.ascontiguousarray(bwt1, dtype=np.uint8).tofile(f1)
        cmd = [REF_BIN, "-1", f1]
        if bwt2 is not None:
            f2 = os.path.join(d, "b.ebwt")
            np.ascontiguousarray(bwt2, dtype=np.uint8).tofile(f2)
            cmd += ["-2", f2]
        if da_ascii is not None:
            f3 = os.path.join(d, "da.txt")
            np.ascontiguousarray(da_ascii, dtype=np.uint8).tofile(f3)
            cmd += ["-d", f3]
        out = os.path.join(d, "out.snp")
        cmd += ["-o", out] + [str(x) for x in flags]
        t0 = time.perf_counter()
        marks = {}
        proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, text=True)
        lines = []
        for line in proc.stdout:
            lines.append(line)
            for key, pat in (("phase2", "Phase 2/4"), ("phase3", "Phase 3/4"), ("phase4", "Phase 4/4"), ("done", "Done.")):
                if key not in marks and line.startswith(pat):
                    marks[key] = time.perf_counter() - t0
        proc.wait()
        total = time.perf_counter() - t0
        if proc.returncode != 0:
            raise RuntimeError("reference failed:\n" + "".join(lines[-20:]))
        text = "".join(lines)
        snp = open(out, "rb").read()
    counters = {}
    for k, pat in _COUNTERS.items():
        m = re.findall(pat, text)
        if m:
            counters[k] = int(m[-1])
    if timed:
        ph = {"load": marks.get("phase2", 0.0),
              "leaves": marks.get("phase3", 0.0) - marks.get("phase2", 0.0),
              "nodes": marks.get("phase4", 0.0) - marks.get("phase3", 0.0),
              "call": marks.get("done", total) - marks.get("phase4", 0.0),
              "total": total}
        return snp, counters, ph
    return snp, counters
