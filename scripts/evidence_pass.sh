# Round evidence: GPU tests, the bench line (with CPU baseline), the ncu launch list of one step and ncu --set full captures of
# the GEMM kernels. Every ncu run follows a plain run of the same command that exited 0.
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r01_gpu_tests.log 2>&1; tail -2 gpurun_out/r01_gpu_tests.log
python bench.py --steps 40 --warmup 5 > gpurun_out/r01_bench.json 2> gpurun_out/r01_bench.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-graph"
$CMD > gpurun_out/plain_l.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 330 -c 170 --csv --log-file gpurun_out/r01_launches.csv $CMD > gpurun_out/ncu_l.log 2>&1
$CMD > gpurun_out/plain_f.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"layer_gemm_pair|wgrad_pair_kernel" -s 24 -c 6 -o gpurun_out/r01_prof_gemm -f $CMD > gpurun_out/ncu_f.log 2>&1
tail -2 gpurun_out/ncu_f.log
