"""Per-step phase times of the whole path on one workload (diagnostic; not a bench value)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from ebwt2indel_b200 import api  # noqa: E402

cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "C4s16"]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
dev = torch.device("cuda:0")
ctx = api.Context(0)
t0 = time.perf_counter()
wl = bench.make_workload(cfg, dev, ctx)
torch.cuda.synchronize()
print(f"built n={wl['n']} in {time.perf_counter() - t0:.1f} s", flush=True)
ctx.trim()
torch.cuda.empty_cache()
p = api.default_params()
for i in range(steps):
    t0 = time.perf_counter()
    snp, st = ctx.run(wl["bwt1"], wl["bwt2"], wl["da"], p)
    dt = time.perf_counter() - t0
    print(f"step {i}: wall {dt * 1e3:8.1f} ms | index {st.ms_index:6.1f} leaves {st.ms_leaves:6.1f} nodes {st.ms_nodes:7.1f} "
          f"call {st.ms_call:6.1f} format {st.ms_format:5.1f} | nodes {st.nodes} levels {st.levels_nodes} maxfront {st.max_frontier}",
          flush=True)
ctx.close()
