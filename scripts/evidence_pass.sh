set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r01_gpu_tests.log 2>&1; tail -2 gpurun_out/r01_gpu_tests.log
python bench.py --steps 40 --warmup 5 > gpurun_out/bench13.json 2> gpurun_out/bench13.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/plain_l.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 420 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_l.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/plain_f.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:layer_gemm_pair -s 9 -c 1 -o gpurun_out/r01_prof_k2_pair -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_f.log 2>&1
ls -la gpurun_out | tail -8
