# A/B of the sharding schemes on N GPUs: bash profiles/n2_ab.sh [N] [tag]
N=${1:-2}; TAG=${2:-x}
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config C4 --steps 3 --warmup 2 --e2e-steps 1 --no-cpu-baseline > gpurun_out/r02_c4_n${N}_${name}_${TAG}.json 2> gpurun_out/r02_c4_n${N}_${name}_${TAG}.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/r02_c4_n${N}_${name}_${TAG}.json')); print('${name}', round(d['ms_per_step'],1), {k: round(v,1) for k,v in d['phase_ms'].items()}, d.get('sharded_host_ms'), d.get('matches_single_gpu'))"
}
run hybrid E2I_X=1
run ranged E2I_RANGED_NODES=1
