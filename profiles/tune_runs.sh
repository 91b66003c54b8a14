for rn in 2 3 4 6; do for rl in 1 2; do
echo "RUNS nodes=$rn leaves=$rl"; E2I_RUNS_NODES=$rn E2I_RUNS_LEAVES=$rl timeout -s KILL 200 python bench.py --config C4s16 --steps 3 --warmup 1 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['phase_ms'])"
done; done
