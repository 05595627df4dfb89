# multi-GPU round on N GPUs: GPU tests that need several devices, NCCL parity check, bench c4
N=${1:-2}
mkdir -p gpurun_out
TAG=${2:-r3}
timeout 300 python -m pytest tests -m gpu -x -q -k "gpus or multi or real or shard" > gpurun_out/${TAG}_pytest_n$N.log 2>&1; echo pytest-exit $?; tail -2 gpurun_out/${TAG}_pytest_n$N.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/mgpu_check.py > gpurun_out/${TAG}_mgpu_check_n$N.log 2>&1; echo check-exit $?; grep -E "rank|Error|error|assert" gpurun_out/${TAG}_mgpu_check_n$N.log | tail -8
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_c4_n$N.json 2> gpurun_out/${TAG}_bench_c4_n$N.err; echo bench-exit $?
tail -3 gpurun_out/${TAG}_bench_c4_n$N.err
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench_c4_n$N.json"))
print("ms/pass", d["ms_per_step"], "resident", d["resident"]["ms_per_step"], "roofline", d["roofline"]["frac"], "step ms", d["roofline"]["ms_per_launch_avg"])
print("rank_step_expand", d["roofline"].get("rank_step_expand_ms"), "nvlink", d["roofline"].get("nvlink_gbs"))
print("e2e", d["e2e"]["seconds_per_step"], d["e2e"]["last_call_seconds"], "parity", d["parity_check"])
PY
