mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
run() { name=$1; shift; env "$@" timeout 300 python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/exp_c4_$name.json 2>> gpurun_out/exp.err; echo "$name exit $?"; }
run base X=1
run l1 GK_P1_L1=1
GK_P1_L1=1 GK_PROFILE=1 timeout 300 python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/exp_c4_prof.json 2> gpurun_out/prof.err; tail -2 gpurun_out/prof.err
python scripts/summarize.py gpurun_out/exp_c4_base.json gpurun_out/exp_c4_l1.json
