"""Generates tests/golden/big/<case>.json: the COMPILED, UNMODIFIED reference (oracle/_ref/ebwt2InDel)
run once on a multi-gigasymbol seeded input (see tests/bigcase.py).  Build container only (needs
oracle/_ref and a few tens of GB of RAM / scratch disk):

    python tests/golden/make_big_golden.py big_c4s30 [--work /tmp/e2i_big] [--threads 8] [--keep]

Input = the workload of ebwt2indel_b200/workloads.py (numpy-seeded genome + read plan), eBWT by the CPU
builder oracle/bcr_build.c.  Recorded: input checksums (so that the GPU test can prove its GPU-built
input is the same string), sha-256 and size of the reference's .snp, the counters and the cluster
histogram it printed, and its wall time per phase (a DRAM-resident single-thread baseline).
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import re
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bigcase  # noqa: E402
from ebwt2indel_b200.workloads import plans_for  # noqa: E402
from oracle import binding as ob  # noqa: E402


def bcr_lib():
    L = C.CDLL(os.path.join(ROOT, "oracle", "libbcr.so"))
    L.orc_bcr_build.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_uint8, C.c_uint64,
                                C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    L.orc_checksum.restype = C.c_uint64
    L.orc_checksum.argtypes = [C.c_void_p, C.c_uint64]
    return L


def flatten(plans):
    """(hap, start per read, revcomp flag per read) over all plans: forward reads then reverse complements, per plan."""
    haps, starts, rcs, off = [], [], [], 0
    for p in plans:
        haps.append(p.hap)
        starts.append(p.starts + off)
        rcs.append(np.zeros(p.n_fwd, np.uint8))
        if p.revcomp:
            starts.append(p.starts + off)
            rcs.append(np.ones(p.n_fwd, np.uint8))
        off += len(p.hap)
    return (np.ascontiguousarray(np.concatenate(haps)), np.ascontiguousarray(np.concatenate(starts).astype(np.int64)),
            np.ascontiguousarray(np.concatenate(rcs)))


def build_ebwt(L, plans, want_owner, threads):
    hap, st, rc = flatten(plans)
    m, rl = len(st), plans[0].read_len
    out = np.empty(m * (rl + 1), np.uint8)
    own = np.empty(m * (rl + 1), np.uint8) if want_owner else None
    second = plans[0].n_reads if len(plans) > 1 else m
    rcode = L.orc_bcr_build(hap.ctypes.data, st.ctypes.data, rc.ctypes.data, m, rl, ord("#"), second, out.ctypes.data,
                            own.ctypes.data if want_owner else None, threads, 1)
    assert rcode == 0, rcode
    return out, own


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("case", choices=sorted(bigcase.CASES))
    ap.add_argument("--work", default="/tmp/e2i_big")
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--keep", action="store_true", help="keep the input files and the .snp in the work directory")
    args = ap.parse_args()
    assert ob.ref_available(), "build the reference first: make -C oracle"
    os.makedirs(args.work, exist_ok=True)
    os.makedirs(bigcase.BIG_DIR, exist_ok=True)
    L = bcr_lib()
    cfg = bigcase.case_config(args.case)
    mode, pl1, pl2 = plans_for(cfg)
    rec = {"case": args.case, "workload": bigcase.CASES[args.case][0], "scale": bigcase.CASES[args.case][1], "mode": mode,
           "flags": [], "reference": "oracle/_ref/ebwt2InDel (unmodified, g++ -Ofast -DNDEBUG), one thread"}
    files = {}
    t0 = time.time()
    b1, own = build_ebwt(L, pl1, mode == 3, args.threads)
    rec["n1"] = int(len(b1))
    rec["bwt1_checksum"] = int(L.orc_checksum(b1.ctypes.data, len(b1)))
    files["1"] = os.path.join(args.work, args.case + ".a.ebwt")
    b1.tofile(files["1"])
    del b1
    if mode == 3:
        da = own + 48
        rec["da_checksum"] = int(L.orc_checksum(da.ctypes.data, len(da)))
        files["d"] = os.path.join(args.work, args.case + ".da.txt")
        da.tofile(files["d"])
        del da, own
    if mode == 2:
        b2, _ = build_ebwt(L, pl2, False, args.threads)
        rec["n2"] = int(len(b2))
        rec["bwt2_checksum"] = int(L.orc_checksum(b2.ctypes.data, len(b2)))
        files["2"] = os.path.join(args.work, args.case + ".b.ebwt")
        b2.tofile(files["2"])
        del b2
    rec["build_seconds"] = time.time() - t0
    print(f"[{args.case}] inputs built in {rec['build_seconds']:.0f} s: {rec}", flush=True)
    del pl1, pl2

    out = os.path.join(args.work, args.case + ".snp")
    cmd = [ob.REF_BIN, "-1", files["1"]]
    if "2" in files:
        cmd += ["-2", files["2"]]
    if "d" in files:
        cmd += ["-d", files["d"]]
    cmd += ["-o", out]
    t0 = time.perf_counter()
    marks, keep_lines = {}, []
    proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, text=True)
    for line in proc.stdout:
        for key, pat in (("phase2", "Phase 2/4"), ("phase3", "Phase 3/4"), ("phase4", "Phase 4/4"), ("done", "Done.")):
            if key not in marks and line.startswith(pat):
                marks[key] = time.perf_counter() - t0
                print(f"[{args.case}] {pat} at {marks[key]:.0f} s", flush=True)
        if not re.match(r"^\s*\d+(\.\d+)?%", line) and "%" not in line[:8]:
            keep_lines.append(line.rstrip("\n"))
    proc.wait()
    total = time.perf_counter() - t0
    assert proc.returncode == 0, "reference failed"
    text = "\n".join(keep_lines)
    counters = {}
    for k, pat in ob._COUNTERS.items():
        m = re.findall(pat, text)
        if m:
            counters[k] = int(m[-1])
    h = hashlib.sha256()
    size = 0
    with open(out, "rb") as f:
        while True:
            chunk = f.read(1 << 24)
            if not chunk:
                break
            h.update(chunk)
            size += len(chunk)
    rec["snp_sha256"] = h.hexdigest()
    rec["snp_bytes"] = size
    rec["counters"] = counters
    rec["stdout_lines"] = keep_lines            # everything but the progress percentages
    rec["ref_seconds"] = {"load": marks.get("phase2", 0.0), "leaves": marks.get("phase3", 0.0) - marks.get("phase2", 0.0),
                          "nodes": marks.get("phase4", 0.0) - marks.get("phase3", 0.0),
                          "call": marks.get("done", total) - marks.get("phase4", 0.0), "total": total}
    rec["ref_host"] = {"cpu": open("/proc/cpuinfo").read().split("model name")[1].split("\n")[0].strip(": \t") if os.path.exists("/proc/cpuinfo") else "?",
                       "threads_used": 1}
    with open(os.path.join(bigcase.BIG_DIR, args.case + ".json"), "w") as f:
        json.dump(rec, f, indent=1)
    print(f"[{args.case}] done: {json.dumps(rec)[:600]}", flush=True)
    if not args.keep:
        for p in list(files.values()) + [out]:
            os.remove(p)


if __name__ == "__main__":
    main()
