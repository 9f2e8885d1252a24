# Round-2 evidence (final build), part A: GPU tests, smoke, the bench line (CPU + eager-CUDA baselines), configs[0], the ncu launch list
# of one step and ncu --set full captures of the dominant kernels. Every ncu run follows a plain run of the same command that exited 0.
# usage (on the GPU box): bash scripts/evidence_pass_r02.sh [A|B]     part B = the other BASELINE configs on one GPU
set -x
PART=${1:-A}
if [ "$PART" = "A" ]; then
python -m pytest tests -m gpu -q -s > gpurun_out/r02_gpu_tests.log 2>&1; tail -2 gpurun_out/r02_gpu_tests.log
python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
timeout 300 python bench.py --steps 40 --warmup 5 --kernel-table > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
timeout 100 python bench.py --workload c1 --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_c1.json 2> gpurun_out/r02_bench_c1.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eager-cuda-baseline --no-sustained --no-e2e --no-graph"
timeout 100 $CMD > gpurun_out/plain_l.log 2>&1 &&
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 150 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu_l.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"layer_gemm_pair|wgrad_pair_kernel|embed_edge_init_kernel|layer_bwd_epilogue|seg_reduce_ell|pooled_message_sum" -s 30 -c 16 -o gpurun_out/r02_prof_step -f $CMD > gpurun_out/ncu_f.log 2>&1
tail -2 gpurun_out/ncu_f.log
ls -la gpurun_out/*.ncu-rep
else
timeout 240 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline --no-eager-cuda-baseline > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err
timeout 200 python bench.py --workload c4 --steps 20 --warmup 3 --no-cpu-baseline --screen-molecules 1250000 > gpurun_out/r02_bench_c4.json 2> gpurun_out/r02_bench_c4.err
timeout 240 python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu-baseline --no-eager-cuda-baseline > gpurun_out/r02_bench_c5.json 2> gpurun_out/r02_bench_c5.err
fi
