"""The position-range traversal on ONE rank (world = 1: every record addressed to itself, no NVLink), against the
plain traversal: isolates what the ranged scheme costs in kernels (destination counting, compaction, segment cursors).
   E2I_RANGED_NODES=1 python profiles/ranged_w1.py C4            (under ncu: add E2I_PROFILE=1 and --profile-from-start off)"""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ebwt2indel_b200 import api, workloads

cfg = workloads.CONFIGS[sys.argv[1]]
ctx = api.Context(0)
dev = torch.device("cuda:0")
wl = workloads.make_workload_gpu(cfg, dev, ctx)
torch.cuda.synchronize()
b1 = ctx.index(wl["bwt1"])
p = api.default_params()
comm = api.comm_shm("/e2i_w1_%d" % os.getpid(), 0, 1)
prof = bool(os.environ.get("E2I_PROFILE"))
for it in range(1 if prof else 2):
    _, _, st = ctx.navigate(b1, None, p)
    print("plain : leaves %.1f ms nodes %.1f ms" % (st.ms_leaves, st.ms_nodes), flush=True)
if prof:
    torch.cuda.profiler.start()
for it in range(1 if prof else 2):
    _, _, sr = ctx.navigate_ranged(comm, b1, None, p)
    print("ranged: leaves %.1f ms nodes %.1f ms (nodes %d = %d)" % (sr.ms_leaves, sr.ms_nodes, sr.nodes, st.nodes), flush=True)
if prof:
    torch.cuda.profiler.stop()
api.comm_free(comm)
