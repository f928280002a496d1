#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (both arms), ncu launch list + full capture of our kernels.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.txt
echo "== bench (ours)"; timeout 1200 python bench.py > $OUT/bench.json 2> $OUT/bench.err; tail -c 1500 $OUT/bench.json; tail -5 $OUT/bench.err
echo "== bench (reference arm)"; timeout 900 python bench.py --impl reference > $OUT/bench_reference.json 2> $OUT/bench_reference.err; tail -c 600 $OUT/bench_reference.json
echo "== ncu launch list (bench.py, 1 warm-up + 1 timed sampling step of 1 denoising step per arm)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3500 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 1 --warmup 1 --nb-steps 1 --no-cpu-baseline --no-extras > $OUT/ncu_bench.log 2>&1
python tools/summarize_launches.py $OUT/launches.csv > $OUT/launches_summary.txt 2>&1; tail -12 $OUT/launches_summary.txt
echo "== ncu --set full on our kernels (same command)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_tc_kernel|iadb_step_kernel|pack_kernel|combine_kernel|groupnorm_nhwc|add_bias_nhwc' \
  -c 24 -f -o $OUT/prof_bench python bench.py --steps 1 --warmup 1 --nb-steps 2 --no-cpu-baseline --no-extras > $OUT/ncu_full.log 2>&1
echo "== ncu --set full, micro driver (cfg1 B=4 and cfg2 B=64 get_noise, K2 at three shapes)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_tc_kernel|iadb_step_kernel|pack_kernel|combine_kernel|groupnorm_nhwc|add_bias_nhwc' \
  -c 40 -f -o $OUT/prof_micro env NO_GRAPH_TIMING=1 python tools/k_micro.py --k2 --iters 1 --flush write > $OUT/ncu_micro.log 2>&1
ls -la $OUT
