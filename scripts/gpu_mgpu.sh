N=${1:-2}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
if [ "$2" = "check" ]; then
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/mgpu_check.py > gpurun_out/mgpu_check_$N.log 2>&1; echo check-exit $?; grep -E "rank|Error|error" gpurun_out/mgpu_check_$N.log | tail -6
fi
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_c4_n$N.json 2> gpurun_out/bench_c4_n$N.err; echo bench-exit $?
tail -2 gpurun_out/bench_c4_n$N.err; grep -o '"ms_per_step": [0-9.]*\|"ms_per_launch_avg": [0-9.]*\|"expand_ms": [0-9.]*\|"seconds_per_step": [0-9.]*' gpurun_out/bench_c4_n$N.json
