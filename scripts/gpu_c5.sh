N=${1:-2}; WL=${2:-c5s}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sharded" > gpurun_out/pytest_sharded.log 2>&1; echo pytest-exit $?; tail -2 gpurun_out/pytest_sharded.log
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/mgpu_check.py > gpurun_out/mgpu_check_$N.log 2>&1; echo check-exit $?; grep -E "rank [0-9]/$N n=40000" gpurun_out/mgpu_check_$N.log
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload $WL --steps 1 --warmup 3 --e2e-steps 0 > gpurun_out/bench_${WL}_n${N}.json 2> gpurun_out/bench_${WL}_n$N.err; echo "bench-exit $?"
grep -o '"ms_per_step": [0-9.]*\|"rank_step_expand_ms": [^}]*\]\]' gpurun_out/bench_${WL}_n${N}.json | head -3
