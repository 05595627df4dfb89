mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo smoke-exit $?; tail -2 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest-exit $?; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --workload c3 --steps 3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo c3-exit $?
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo c4-exit $?
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --compress 0 --e2e-steps 0 > gpurun_out/bench_c4_nocompress.json 2> gpurun_out/bench_c4_nocompress.err; echo c4nc-exit $?
if [ "$1" = "ncu" ]; then
B="python bench.py --workload c3 --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
timeout 300 $B > gpurun_out/plain1.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c3.csv $B > gpurun_out/ncu1.log 2>&1; echo ncu1-exit $?
timeout 300 $B > gpurun_out/plain2.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:phi_step -s 30 -c 2 -o gpurun_out/prof_c3 $B > gpurun_out/ncu2.log 2>&1; echo ncu2-exit $?
fi
