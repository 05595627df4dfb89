mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
for ST in 1 2; do for PF in 0 1; do
  export GK_STAGES=$ST GK_PREFETCH=$PF
  timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "fixture or synthetic or bottleneck" > gpurun_out/pytest_s${ST}p${PF}.log 2>&1; echo "pytest s$ST p$PF exit $?"
  timeout 300 python bench.py --workload c3 --steps 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/exp_c3_s${ST}p${PF}.json 2>> gpurun_out/exp.err
  timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/exp_c4_s${ST}p${PF}.json 2>> gpurun_out/exp.err
done; done
python scripts/summarize.py gpurun_out/exp_c*.json
