# one bench line on N GPUs: bash profiles/n2_one.sh N tag [ENV=1 ...]
N=$1; TAG=$2; shift; shift
env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config C4 --steps 3 --warmup 2 --e2e-steps 3 --no-cpu-baseline > gpurun_out/r02_c4_n${N}_${TAG}.json 2> gpurun_out/r02_c4_n${N}_${TAG}.err
python -c "
import json,sys; d=json.load(open('gpurun_out/r02_c4_n${N}_${TAG}.json')); print('${TAG}', round(d['ms_per_step'],1), {k: round(v,1) for k,v in d['phase_ms'].items()}, {k: round(v,1) for k,v in d.get('sharded_host_ms',{}).items()}, d.get('matches_single_gpu'), 'e2e', round(d['e2e']['ms_per_step'],1))"
