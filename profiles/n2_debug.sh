# breakdown of the position-range node pass at N ranks: bash profiles/n2_debug.sh [N]
N=${1:-2}
E2I_DEBUG=1 E2I_RANGED_NODES=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config C4 --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline > gpurun_out/r02_c4_n${N}_dbg.json 2> gpurun_out/r02_c4_n${N}_dbg.err
grep "ranged" gpurun_out/r02_c4_n${N}_dbg.err | tail -$((2*N))
