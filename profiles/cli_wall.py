"""Wall time of the drop-in binary from FILES to the .snp file (process start to exit) at a bench configuration:
   python profiles/cli_wall.py C4 [dir]
Runs bin/ebwt2InDel twice plainly (the file is in the page cache after being written) and twice with the packed-index
sidecar (E2I_INDEX_CACHE=1: the first run writes <file>.e2ix, the second loads it instead of the ASCII text); checks
the .snp against the golden of the compiled reference when there is one."""
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from ebwt2indel_b200 import api  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C4"
base = sys.argv[2] if len(sys.argv) > 2 else None
ctx = api.Context(0)
t0 = time.perf_counter()
wl = bench.make_workload(bench.CONFIGS[name], torch.device("cuda:0"), ctx)
torch.cuda.synchronize()
print(f"workload {name}: n={wl['n']} built in {time.perf_counter() - t0:.1f} s", flush=True)
golden = os.path.join(ROOT, "tests", "golden", "big", f"big_{name.lower()}.json")
want = json.load(open(golden))["snp_sha256"] if os.path.exists(golden) else None
out = {"workload": name, "n": wl["n"], "runs": []}
with tempfile.TemporaryDirectory(dir=base) as d:
    t0 = time.perf_counter()
    files = bench.write_inputs(d, wl)
    print(f"inputs written in {time.perf_counter() - t0:.1f} s", flush=True)
    del wl
    ctx.close()
    torch.cuda.empty_cache()
    exe = os.path.join(ROOT, "bin", "ebwt2InDel")
    runs = (("plain", {}), ("plain", {}), ("sidecar-write", {"E2I_INDEX_CACHE": "1"}), ("sidecar-load", {"E2I_INDEX_CACHE": "1"}))
    if os.environ.get("E2I_CLI_DEBUG"):                 # one instrumented run: where the wall time goes
        runs = (("debug", {"E2I_DEBUG": "1"}),)
    for tag, env in runs:
        t0 = time.perf_counter()
        r = subprocess.run([exe] + files, capture_output=True, text=True, env={**os.environ, **env}, timeout=600)
        dt = time.perf_counter() - t0
        snp = open(files[-1], "rb").read()
        sha = hashlib.sha256(snp).hexdigest()
        line = [ln for ln in r.stderr.splitlines() if ln.startswith("[e2i] n=")]
        rec = {"mode": tag, "wall_s": round(dt, 3), "rc": r.returncode, "snp_bytes": len(snp), "matches_reference": (sha == want) if want else None,
               "summary": line[-1] if line else r.stderr[-300:]}
        out["runs"].append(rec)
        print(rec, flush=True)
        if tag == "debug":
            print("\n".join(ln for ln in r.stderr.splitlines() if ("cli:" in ln or "run_files" in ln or "navigate:" in ln or "run:" in ln or "call:" in ln)), flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"r02_cli_wall_{name.lower()}.json"), "w"), indent=1)
