# BASELINE configs[4] on 8 GPUs: gen.phi (ranked, sharded) + gen.gc of the founders
mkdir -p gpurun_out
TAG=${1:-r3}
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 2 --warmup 3 --workload c5 --no-cpu-baseline --no-parity-check --e2e-steps 1 > gpurun_out/${TAG}_bench_c5_n8.json 2> gpurun_out/${TAG}_bench_c5_n8.err; echo phi-exit $?
tail -3 gpurun_out/${TAG}_bench_c5_n8.err
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --steps 1 --warmup 1 --workload c5 --op gc > gpurun_out/${TAG}_bench_c5_gc_n8.json 2> gpurun_out/${TAG}_bench_c5_gc_n8.err; echo gc-exit $?
tail -2 gpurun_out/${TAG}_bench_c5_gc_n8.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench_c5_n8.json"))
    print("phi ms/pass", d["ms_per_step"], "resident", d["resident"]["ms_per_step"], "mean", d["config"]["phi_mean"], "cuts", d["config"]["cuts"][-5:], "ranked", d["config"]["ranked"])
    print("rank_step_expand", d["roofline"].get("rank_step_expand_ms"), "nvlink", d["roofline"].get("nvlink_gbs"))
    print("e2e", d["e2e"])
except Exception as ex:
    print("phi line missing", ex)
try:
    d = json.load(open("gpurun_out/${TAG}_bench_c5_gc_n8.json"))
    print("gc ms", d["ms_per_step"], d["config"], d["roofline"])
except Exception as ex:
    print("gc line missing", ex)
PY
