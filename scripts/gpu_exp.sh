mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest-exit $?; tail -5 gpurun_out/pytest_gpu.log
