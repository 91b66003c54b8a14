"""Multi-gigasymbol parity cases (n > 2^32 and the full-size BASELINE.json configs).

The compiled, unmodified reference needs 0.5 - 3 hours of one host core per case, so it is run ONCE
in the build container by tests/golden/make_big_golden.py, which records, per case, a checksum of
the input eBWT(s) it built with the CPU builder (oracle/bcr_build.c), the sha-256 of the reference's
.snp output and every counter it printed (tests/golden/big/<case>.json).  The GPU tests rebuild the
same seeded input with the GPU builder, check the input checksums, run the product through the C ABI
and compare sha-256 + counters: bit-exact parity at sizes where the 64-bit / multi-superblock code
paths of the kernels actually execute.
"""
from __future__ import annotations

import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
BIG_DIR = os.path.join(HERE, "golden", "big")

# case -> (workload name in ebwt2indel_b200.workloads.CONFIGS, scale)
CASES = {
    "big_c4s30": ("C4", 0.3),       # mode -1, n = 4.53 G  (> 2^32: two superblocks, 64-bit node arithmetic)
    "big_c2": ("C2", 1.0),          # mode -d, n = 1.01 G, full size
    "big_c3": ("C3", 1.0),          # mode -2, 2 x 1.51 G, full size
    "big_c5s25": ("C5", 0.25),      # mode -1 metagenome, n = 5.05 G
    "big_c4": ("C4", 1.0),          # mode -1, n = 15.1 G: the headline configuration itself
    "big_c1": ("C1", 1.0),          # mode -1, n = 40.4 M (quick check of this machinery)
}


def case_config(name: str) -> dict:
    from ebwt2indel_b200.workloads import CONFIGS, scaled
    wl, scale = CASES[name]
    return CONFIGS[wl] if scale == 1.0 else scaled(CONFIGS[wl], scale)


def load_golden(name: str):
    p = os.path.join(BIG_DIR, name + ".json")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        return json.load(f)
