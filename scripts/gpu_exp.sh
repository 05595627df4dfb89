mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest-exit $?; tail -3 gpurun_out/pytest_gpu.log
run() { name=$1; wl=$2; shift; shift; env "$@" timeout 90 python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/exp_${wl}_$name.json 2>> gpurun_out/exp.err; echo "$name exit $?"; }
run h16 c4 X=1
run h16p100 c4 GK_LAG_PCT=100
run h16p300 c4 GK_LAG_PCT=300
run h16nw c4 GK_DEBUG_NOWAIT=1
run h16 c3 X=1
python scripts/summarize.py gpurun_out/exp_c4_h16*.json gpurun_out/exp_c3_h16.json
