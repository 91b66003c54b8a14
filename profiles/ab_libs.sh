# A/B of library variants on a workload: E2I_LIB selects the .so;  bash profiles/ab_libs.sh <config> lib1.so lib2.so ...
cfg=$1; shift
for lib in "$@"; do
echo "LIB $lib"; E2I_LIB=$PWD/ebwt2indel_b200/$lib timeout -s KILL 400 python bench.py --config $cfg --steps 3 --warmup 1 --e2e-steps 1 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['phase_ms'])"
done
