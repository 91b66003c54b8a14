# A/B of library variants on the C4 workload: E2I_LIB selects the .so (profiles/README.md)
for lib in libe2i.so libe2i_w48.so libe2i_w96.so; do
echo "LIB $lib"; E2I_LIB=$PWD/ebwt2indel_b200/$lib timeout -s KILL 400 python bench.py --config C4 --steps 3 --warmup 1 --e2e-steps 1 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['phase_ms'])"
done
