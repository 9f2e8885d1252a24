# ncu --set full captures of the secondary kernels (one launch each), after the same command exited 0 without ncu
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_m.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"wgrad_tma_kernel|seg_reduce_v4|embedding_bag_bwd_partial|embedding_bag_sum_kernel|layer_bwd_epilogue" -s 40 -c 14 -o gpurun_out/prof_misc -f $CMD > gpurun_out/ncu_m.log 2>&1
tail -3 gpurun_out/ncu_m.log
