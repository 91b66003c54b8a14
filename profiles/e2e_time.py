"""e2e step timing with host (pinned) inputs: H2D time and the rest (diagnostic)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from ebwt2indel_b200 import api  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C4"
ctx = api.Context(0)
wl = bench.make_workload(bench.CONFIGS[name], torch.device("cuda:0"), ctx)
h = [None if t is None else t.cpu().pin_memory().numpy() for t in (wl["bwt1"], wl["bwt2"], wl["da"])]
del wl
ctx.trim()
torch.cuda.empty_cache()
p = api.default_params()
for i in range(4):
    t0 = time.perf_counter()
    snp, st = ctx.run(h[0], h[1], h[2], p, copy=False)
    dt = time.perf_counter() - t0
    nb = sum(x.nbytes for x in h if x is not None)
    print(f"step {i}: wall {dt * 1e3:8.1f} ms | h2d {st.ms_h2d:7.1f} ms ({nb / st.ms_h2d / 1e6:5.1f} GB/s) index {st.ms_index:6.1f} "
          f"leaves {st.ms_leaves:6.1f} nodes {st.ms_nodes:7.1f} call {st.ms_call:6.1f} format {st.ms_format:5.1f}", flush=True)
ctx.close()
