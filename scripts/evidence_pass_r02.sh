# Round-2 evidence: GPU tests, smoke, the bench line (CPU + eager-CUDA baselines), the other BASELINE configs, the ncu launch list of
# one step and ncu --set full captures of the dominant kernels. Every ncu run follows a plain run of the same command that exited 0.
set -x
python -m pytest tests -m gpu -q -s > gpurun_out/r02_gpu_tests.log 2>&1; tail -2 gpurun_out/r02_gpu_tests.log
python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
python bench.py --steps 40 --warmup 5 --kernel-table > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err
python bench.py --workload c4 --steps 20 --warmup 3 --screen-molecules 1250000 > gpurun_out/r02_bench_c4.json 2> gpurun_out/r02_bench_c4.err
python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c5.json 2> gpurun_out/r02_bench_c5.err
python bench.py --workload c1 --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_c1.json 2> gpurun_out/r02_bench_c1.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eager-cuda-baseline --no-sustained --no-e2e --no-graph"
$CMD > gpurun_out/plain_l.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 150 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu_l.log 2>&1
$CMD > gpurun_out/plain_f.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"layer_gemm_pair|wgrad_pair_kernel|embed_edge_init_kernel|embed_bwd_mma_kernel|layer_bwd_epilogue|seg_reduce_ell" -s 30 -c 12 -o gpurun_out/r02_prof_step -f $CMD > gpurun_out/ncu_f.log 2>&1
tail -2 gpurun_out/ncu_f.log
CMD3="python bench.py --workload c3 --batch 2048 --steps 1 --warmup 3 --no-cpu-baseline --no-eager-cuda-baseline --no-sustained --no-e2e --no-graph"
$CMD3 > gpurun_out/plain_c3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"layer_gemm_pair" -s 10 -c 2 -o gpurun_out/r02_prof_c3 -f $CMD3 > gpurun_out/ncu_c3.log 2>&1
tail -2 gpurun_out/ncu_c3.log
ls -la gpurun_out/*.ncu-rep
