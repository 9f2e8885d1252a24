# one ncu --set full capture of the K2 forward kernel inside a short bench run (after the same command exited 0 without ncu)
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/plain_f.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:layer_gemm_pair -s 19 -c 3 -o gpurun_out/prof_k2 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_f.log 2>&1
tail -3 gpurun_out/ncu_f.log
