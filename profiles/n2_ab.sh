run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --config C4 --steps 3 --warmup 2 --e2e-steps 1 --no-cpu-baseline > gpurun_out/r02_c4_n2_${name}.json 2> gpurun_out/r02_c4_n2_${name}.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/r02_c4_n2_${name}.json')); print('${name}', d['ms_per_step'], d['phase_ms'])"
}
run hybrid2 E2I_X=1
run ranged2 E2I_RANGED_NODES=1
