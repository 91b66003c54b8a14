import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
FLAG_FIELD = {"-L": "k_left", "-R": "k_right", "-k": "K", "-g": "max_gap", "-v": "max_snvs", "-m": "mcov_out",
              "-c": "complexity", "-q": "max_variants_per_position", "-t": "term"}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    """Returns dict(bwt1, bwt2|None, da|None, flags{field: value}, snp bytes, counters{name: int})."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    flags = [str(x) for x in z["flags"]]
    fields = {FLAG_FIELD[flags[i]]: int(flags[i + 1]) for i in range(0, len(flags), 2)}
    return {
        "bwt1": z["bwt1"],
        "bwt2": z["bwt2"] if len(z["bwt2"]) else None,
        "da": z["da"] if len(z["da"]) else None,
        "flags": fields,
        "snp": z["snp"].tobytes(),
        "counters": {str(k): int(v) for k, v in zip(z["counter_names"], z["counter_values"])},
    }


def resolved_fields(fields):
    """The reference's "0 means default" rule (ebwt2InDel.cpp:1740-1746) applied to raw flag values."""
    d = {"k_left": 31, "k_right": 30, "K": 16, "max_gap": 10, "max_snvs": 2, "mcov_out": 3, "complexity": 20}
    out = dict(fields)
    for k, v in d.items():
        if out.get(k, 0) == 0:
            out[k] = v
    return out


@pytest.fixture(scope="session")
def oracle():
    from oracle import binding
    if not os.path.exists(binding.LIB_PATH):
        binding.build()
    binding.lib()
    return binding


@pytest.fixture(scope="session")
def e2i():
    """The ctypes binding of libe2i.so.  The library and the CLI binaries are build products (not in
    git): build them when they are missing (nvcc cross-compiles sm_100a without a GPU)."""
    import subprocess
    from ebwt2indel_b200 import api
    needed = [api.LIB_PATH, os.path.join(ROOT, "bin", "ebwt2InDel"), os.path.join(ROOT, "bin", "filter_snp")]
    if not all(os.path.exists(p) for p in needed):
        subprocess.run(["make", "-s", "-j8", "-C", ROOT, "all"], check=True)
    api.lib()
    return api


@pytest.fixture(scope="session")
def gpu_ctx(e2i):
    ctx = e2i.Context(0)
    yield ctx
    ctx.close()
