"""Drop-in check at BASELINE configs[0] size: run the compiled reference and bin/ebwt2InDel on the
same eBWT file with the same argv; compare the .snp bytes and the counter lines of stdout."""
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from ebwt2indel_b200 import api  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C1"
ctx = api.Context(0)
wl = bench.make_workload(bench.CONFIGS[name], torch.device("cuda:0"), ctx)
ctx.close()
pat = re.compile(r"^(Processed|Computed|Found|Analyzed|Stored to file|Average cluster).*$", re.M)
with tempfile.TemporaryDirectory() as d:
    files = bench.write_inputs(d, wl)
    outs = {}
    for tag, exe in (("ref", os.path.join(ROOT, "oracle", "_ref", "ebwt2InDel")), ("b200", os.path.join(ROOT, "bin", "ebwt2InDel"))):
        argv = [exe] + files[:-1] + [os.path.join(d, tag + ".snp")]
        t0 = time.perf_counter()
        r = subprocess.run(argv, capture_output=True, text=True)
        dt = time.perf_counter() - t0
        snp = open(os.path.join(d, tag + ".snp"), "rb").read()
        lines = [ln for ln in pat.findall(r.stdout) if "LCP threshold" not in ln]
        outs[tag] = (snp, lines, dt, r.returncode)
        print(f"{tag}: rc={r.returncode} wall={dt:.2f}s snp={len(snp)} bytes")
    same_snp = outs["ref"][0] == outs["b200"][0]
    ref_lines, my_lines = outs["ref"][1], outs["b200"][1]
    print("snp identical:", same_snp)
    print("counter lines identical:", ref_lines == my_lines)
    if ref_lines != my_lines:
        for a, b in zip(ref_lines, my_lines):
            if a != b:
                print("  ref :", a, "\n  b200:", b)
    print(f"speed-up (wall, files -> .snp): {outs['ref'][2] / outs['b200'][2]:.1f}x")
