# end-of-round measurements on one B200: bench lines, ncu launch list, ncu full captures of the two dominant kernels
TAG=${1:-r02b}
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_c4.json 2> gpurun_out/${TAG}_bench_c4.err; echo c4-exit $?
python bench.py --workload c3 --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; echo c3-exit $?
python bench.py --workload c1 --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_c1.json 2> gpurun_out/${TAG}_bench_c1.err; echo c1-exit $?
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_reference_arm_c4.json 2> gpurun_out/${TAG}_reference_arm_c4.err; echo ref-exit $?
python scripts/prof_one.py c4 > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_c4.csv python scripts/prof_one.py c4 > gpurun_out/${TAG}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:phi_step_g4 -s 6 -c 2 -f -o gpurun_out/${TAG}_prof_step python scripts/prof_one.py c4 > gpurun_out/${TAG}_ncu_step.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:expand_sym -c 1 -f -o gpurun_out/${TAG}_prof_expand python scripts/prof_one.py c4 > gpurun_out/${TAG}_ncu_expand.log 2>&1
ls -la gpurun_out/${TAG}_* | tail -12
