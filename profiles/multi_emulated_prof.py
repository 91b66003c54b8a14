"""Two ranks of e2i_run_multi on ONE device (threads), position-range sharding for both passes: under
   ncu --kernel-name regex:'expand|compact' this gives the per-kernel device time of a HALF-sized rank without any NVLink
   in the picture.   E2I_RANGED_NODES=1 python profiles/multi_emulated_prof.py C4 2"""
import os
import sys
import time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ebwt2indel_b200 import api, workloads

cfg = workloads.CONFIGS[sys.argv[1]]
world = int(sys.argv[2])
ctx = api.Context(0)
wl = workloads.make_workload_gpu(cfg, torch.device("cuda:0"), ctx)
torch.cuda.synchronize()
h1 = wl["bwt1"].cpu().numpy()
ctx.close()
del wl
torch.cuda.empty_cache()
t = time.time()
snp, st = api.run_multi([0] * world, h1, None, None, api.default_params())
print("ranged x%d on one device: wall %.2f s, nodes %.1f ms, leaves %.1f ms, %d bytes of text" % (world, time.time() - t, st.ms_nodes, st.ms_leaves, len(snp)), flush=True)
