mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 400 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo pytest-exit $?; tail -12 gpurun_out/pytest_gpu.log
