mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest-exit $?; tail -3 gpurun_out/pytest_gpu.log
GK_STEP_KERNEL=tma timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fixture or synthetic or bottleneck" > gpurun_out/pytest_gpu_tma.log 2>&1; echo pytest-tma-exit $?; tail -3 gpurun_out/pytest_gpu_tma.log
