set -x
timeout 300 python -m pytest tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2_mgpu1.log
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/r2_bench_cfg2.json 2> gpurun_out/r2_bench_cfg2.err
tail -c 600 gpurun_out/r2_bench_cfg2.err
for c in 3 4 5; do timeout 600 python bench.py --config $c --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_cfg$c.json 2> gpurun_out/r2_bench_cfg$c.err; tail -c 400 gpurun_out/r2_bench_cfg$c.err; done
cat gpurun_out/r2_mgpu1.log
