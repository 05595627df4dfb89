mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo smoke-exit $?; tail -2 gpurun_out/smoke.log
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest-exit $?; tail -5 gpurun_out/pytest_gpu.log
timeout 200 python bench.py --workload c3 --steps 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo c3-exit $?
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo c4-exit $?
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo ref-exit $?
python scripts/summarize.py gpurun_out/bench_c3.json gpurun_out/bench_c4.json
if [ "$1" = "ncu" ]; then
B="python bench.py --workload ${2:-c4} --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
timeout 200 $B > gpurun_out/plain1.log 2>&1 && timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${2:-c4}.csv $B > gpurun_out/ncu1.log 2>&1; echo ncu1-exit $?
timeout 200 $B > gpurun_out/plain2.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:phi_step -s ${3:-30} -c 2 -o gpurun_out/prof_${2:-c4} $B > gpurun_out/ncu2.log 2>&1; echo ncu2-exit $?
fi
