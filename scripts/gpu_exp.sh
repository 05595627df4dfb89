mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest-exit $?; tail -2 gpurun_out/pytest_gpu.log
timeout 120 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo c4-exit $?
python scripts/summarize.py gpurun_out/bench_c4.json; grep -o '"plan_seconds": [0-9.]*' gpurun_out/bench_c4.json
