"""Debug helper: sharded vs unsharded bitvectors on a bench workload."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from ebwt2indel_b200 import api, distributed as dd  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C4s16"
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = api.Context(0)
wl = bench.make_workload(bench.CONFIGS[name], torch.device("cuda:0"), ctx)
p = api.default_params()
b1 = ctx.index(wl["bwt1"])
full, _, fst = ctx.navigate(b1, None, p)
(pt, wt), (pm, wm) = full.device_words()
fthr = dd.wrap_device_words(pt, wt, "cuda:0").clone()
fmin = dd.wrap_device_words(pm, wm, "cuda:0").clone()
torch.cuda.synchronize()
print("full", {k: getattr(fst, k) for k in ("nodes", "leaves", "n_min", "lcp_values", "levels_nodes", "levels_leaves")})
del full
sthr, smin = torch.zeros_like(fthr), torch.zeros_like(fmin)
tot = {}
for s in range(ns):
    part, _, st = ctx.navigate(b1, None, p, shard=s, n_shards=ns)
    (pt, wt), (pm, wm) = part.device_words()
    t, m = dd.wrap_device_words(pt, wt, "cuda:0"), dd.wrap_device_words(pm, wm, "cuda:0")
    print("shard", s, {k: getattr(st, k) for k in ("nodes", "leaves", "n_min", "lcp_values", "levels_nodes")},
          "overlap thr", int((sthr & t).ne(0).sum()), "min", int((smin & m).ne(0).sum()))
    sthr |= t
    smin |= m
    torch.cuda.synchronize()
    for k in ("nodes", "leaves", "n_min", "lcp_values"):
        tot[k] = tot.get(k, 0) + getattr(st, k)
    del part
print("sum", tot)
for nm, a, b in (("thr", sthr, fthr), ("min", smin, fmin)):
    d = (a ^ b)
    idx = d.ne(0).nonzero().flatten()
    print(nm, "differing words", idx.numel(), "missing bits", int(torch.tensor([bin(x & 0xffffffff).count('1') for x in (b & ~a)[idx[:1000]].tolist()]).sum()) if idx.numel() else 0,
          "extra bits", int(torch.tensor([bin(x & 0xffffffff).count('1') for x in (a & ~b)[idx[:1000]].tolist()]).sum()) if idx.numel() else 0,
          "first words", idx[:10].tolist())
