"""Regenerates tests/golden/*.npz by running the COMPILED, UNMODIFIED reference (oracle/_ref/ebwt2InDel,
built from /root/reference by oracle/Makefile) on seeded synthetic inputs.

The reference ships no fixtures, so these vectors are the pin: every case stores the
exact input bytes, the flags, the reference's .snp output and the counters it printed.  Run in the
build container only (needs /root/reference):  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ebwt2indel_b200 import synth  # noqa: E402
from oracle import binding as ob  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# flag name -> Params field
FLAG_FIELD = {"-L": "k_left", "-R": "k_right", "-k": "K", "-g": "max_gap", "-v": "max_snvs", "-m": "mcov_out",
              "-c": "complexity", "-q": "max_variants_per_position", "-t": "term"}


def save(name, bwt1, bwt2=None, da=None, flags=()):
    snp, counters = ob.run_ref(bwt1, bwt2, da, flags)
    keys = sorted(counters)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        bwt1=np.asarray(bwt1, dtype=np.uint8),
        bwt2=np.asarray(bwt2 if bwt2 is not None else [], dtype=np.uint8),
        da=np.asarray(da if da is not None else [], dtype=np.uint8),
        flags=np.asarray([str(f) for f in flags], dtype="U16"),
        snp=np.frombuffer(snp, dtype=np.uint8),
        counter_names=np.asarray(keys, dtype="U32"),
        counter_values=np.asarray([counters[k] for k in keys], dtype=np.uint64),
    )
    print(f"{name}: n1={len(bwt1)} snp={len(snp)} bytes {counters}")


def main():
    assert ob.ref_available(), "build the reference first: make -C oracle"
    reads = synth.diploid_reads(6000, 14, 5, 24, 100, seed=11)
    bwt, _ = synth.ebwt_naive(reads)
    save("m1_default", bwt)
    save("m1_flags", bwt, flags=("-m", 2, "-L", 20, "-R", 15, "-k", 10, "-g", 5, "-c", 10, "-v", 3))
    save("m1_q2_g0", bwt, flags=("-q", 2, "-g", 0, "-m", 4))
    dollar = bwt.copy()
    dollar[dollar == ord("#")] = ord("$")
    save("m1_term36", dollar, flags=("-t", 36))
    short = synth.diploid_reads(1500, 6, 2, 30, 40, seed=12)
    sb, _ = synth.ebwt_naive(short)
    save("m1_short_reads", sb, flags=("-L", 12, "-R", 10, "-k", 8, "-c", 6, "-g", 3))
    meta = synth.metagenome_plan(3, 4, 3000, 0.004, 12, 60, seed=9).materialize()
    mb, _ = synth.ebwt_naive(meta)
    save("m1_metagenome", mb, flags=("-m", 2))
    r0, r1 = synth.two_individuals_reads(4000, 10, 4, 24, 100, seed=13)
    merged, da = synth.merged_ebwt_da(r0, r1)
    save("m3_default", merged, da=da)
    save("m3_flags", merged, da=da, flags=("-m", 2, "-L", 24, "-R", 20, "-k", 12, "-g", 6))
    b0, _ = synth.ebwt_naive(r0)
    b1, _ = synth.ebwt_naive(r1)
    save("m2_default", b0, b1)
    save("m2_flags", b0, b1, flags=("-m", 2, "-L", 24, "-R", 20, "-k", 12, "-g", 6, "-q", 2))


def save_stdout():
    """tests/golden/<case>.stdout.txt: what the reference prints (progress percentages dropped, file names
    replaced by placeholders) -- the CLI drop-in test diffs bin/ebwt2InDel's stdout against it."""
    import re
    import subprocess
    import tempfile
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import golden_names, load_golden
    inv = {v: k for k, v in FLAG_FIELD.items()}
    for name in golden_names():
        g = load_golden(name)
        with tempfile.TemporaryDirectory() as d:
            names = {}
            cmd = [ob.REF_BIN]
            for key, flag, fn in (("bwt1", "-1", "a.ebwt"), ("bwt2", "-2", "b.ebwt"), ("da", "-d", "da.txt")):
                if g[key] is not None:
                    path = os.path.join(d, fn)
                    g[key].tofile(path)
                    cmd += [flag, path]
                    names[path] = "<" + key + ">"
            out = os.path.join(d, "out.snp")
            names[out] = "<out>"
            for k, v in g["flags"].items():
                cmd += [inv[k], str(v)]
            text = subprocess.run(cmd + ["-o", out], capture_output=True, text=True, check=True).stdout
        for path, ph in names.items():
            text = text.replace(path, ph)
        lines = [ln for ln in text.split("\n") if not re.match(r"^\s*(LCP: )?\d+(\.\d+)?%", ln) and "%" not in ln[:10]]
        with open(os.path.join(HERE, name + ".stdout.txt"), "w") as f:
            f.write("\n".join(lines))
        print(name, len(lines), "stdout lines")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "stdout":
        save_stdout()
    else:
        main()
        save_stdout()
