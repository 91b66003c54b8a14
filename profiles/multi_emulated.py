"""e2i_run_multi with the ranks on ONE device (logic check + rough timing): python profiles/multi_emulated.py C4s16 2"""
import sys, time, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ebwt2indel_b200 import api, workloads
cfg = workloads.CONFIGS[sys.argv[1]]; world = int(sys.argv[2])
ctx = api.Context(0)
wl = workloads.make_workload_gpu(cfg, torch.device("cuda:0"), ctx)
h1 = wl["bwt1"].cpu().numpy()
snp1, st1 = ctx.run(h1, None, None, api.default_params())
print("single: nodes ms", st1.ms_nodes, "leaves ms", st1.ms_leaves, "levels", st1.levels_nodes)
ctx.close(); del wl; torch.cuda.empty_cache()
for mode in ("ranged", "subtree"):
    os.environ["E2I_SHARDING"] = mode
    t = time.time()
    snp, st = api.run_multi([0] * world, h1, None, None, api.default_params())
    print(mode, "wall", round(time.time() - t, 2), "nodes ms", st.ms_nodes, "leaves ms", st.ms_leaves, "same text", snp == snp1)
