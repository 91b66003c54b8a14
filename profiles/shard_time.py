"""Time one traversal shard of an N-way sharded run on a single GPU (tuning aid for the multi-GPU path)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from ebwt2indel_b200 import api  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C4"
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 8
ctx = api.Context(0)
wl = bench.make_workload(bench.CONFIGS[name], torch.device("cuda:0"), ctx)
ctx.trim()
torch.cuda.empty_cache()
p = api.default_params()
b1 = ctx.index(wl["bwt1"])
b2 = ctx.index(wl["bwt2"]) if wl["bwt2"] is not None else None
for rep in range(2):
    for s in (0, ns // 2, ns - 1):
        part, da, st = ctx.navigate(b1, b2, p, shard=s, n_shards=ns)
        print(f"rep {rep} shard {s}/{ns}: leaves {st.ms_leaves:7.1f} ms nodes {st.ms_nodes:7.1f} ms | nodes {st.nodes} sweeps {st.levels_nodes}", flush=True)
        del part, da
ctx.close()
