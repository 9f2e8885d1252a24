# Round evidence: GPU tests, smoke, the bench line (with CPU baseline), the other BASELINE configs, the ncu launch list of one step
# and ncu --set full captures of the GEMM kernels. Every ncu run follows a plain run of the same command that exited 0.
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r01_gpu_tests.log 2>&1; tail -2 gpurun_out/r01_gpu_tests.log
python __graft_entry__.py smoke > gpurun_out/r01_smoke.log 2>&1; tail -2 gpurun_out/r01_smoke.log
python bench.py --steps 40 --warmup 5 > gpurun_out/r01_bench.json 2> gpurun_out/r01_bench.err
python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r01_bench_c3.json 2> gpurun_out/r01_bench_c3.err
python bench.py --workload c4 --steps 20 --warmup 3 > gpurun_out/r01_bench_c4.json 2> gpurun_out/r01_bench_c4.err
python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r01_bench_c5.json 2> gpurun_out/r01_bench_c5.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-graph"
$CMD > gpurun_out/plain_l.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 320 -c 160 --csv --log-file gpurun_out/r01_launches.csv $CMD > gpurun_out/ncu_l.log 2>&1
$CMD > gpurun_out/plain_f.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"layer_gemm_pair|wgrad_pair_kernel" -s 24 -c 6 -o gpurun_out/r01_prof_gemm -f $CMD > gpurun_out/ncu_f.log 2>&1
tail -2 gpurun_out/ncu_f.log
