#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (both arms), ncu launch list + full capture of our kernels.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh [tag]
# gpurun only copies back <= 64 MiB: ncu reports are summarised on the box (tools/ncu_summary.py) and
# dropped when large; the text summaries are what gets committed under profiles/.
TAG=${1:-r02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
KERNELS='gemv_kernel|gemm_tc_kernel|snapshot_u8|iadb_step_kernel|ddim_step_kernel|to_u8_kernel|pack_kernel|combine_kernel|groupnorm_nhwc|add_bias_nhwc|attention_small|upsample2x|linear_tc|shortcut_tc|conv_in3x3'
keep_small() { if [ -f "$1" ] && [ $(stat -c %s "$1") -gt 20000000 ]; then rm -f "$1"; fi; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.txt
echo "== bench (ours)"; timeout 1200 python bench.py > $OUT/bench.json 2> $OUT/bench.err; tail -c 1500 $OUT/bench.json; tail -5 $OUT/bench.err
echo "== bench (reference arm)"; timeout 900 python bench.py --impl reference > $OUT/bench_reference.json 2> $OUT/bench_reference.err; tail -c 600 $OUT/bench_reference.json
if [ -x build/stream_probe ]; then echo "== stream probe (make probes)"; (timeout 120 build/stream_probe 7 0; timeout 120 build/stream_probe 7 1) > $OUT/stream_probe.txt 2>&1; tail -60 $OUT/stream_probe.txt; fi
if [ -x build/get_noise_probe ]; then (timeout 60 build/get_noise_probe 4 3 9 0; timeout 60 build/get_noise_probe 4 3 9 1; timeout 60 build/get_noise_probe 4 3 9 2; timeout 60 build/get_noise_probe 64 3 9 1) > $OUT/get_noise_probe.txt 2>&1; grep -E "^#|parity|whole call|span" $OUT/get_noise_probe.txt; fi
if [ -z "$SKIP_NCU" ]; then
echo "== ncu launch list (bench.py, 1 warm-up + 1 timed sampling step of 1 denoising step per arm)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 2500 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 1 --warmup 1 --nb-steps 1 --no-cpu-baseline --no-extras --no-reference-gpu > $OUT/ncu_bench.log 2>&1
python tools/summarize_launches.py $OUT/launches.csv > $OUT/launches_summary.txt 2>&1; tail -12 $OUT/launches_summary.txt
echo "== ncu --set full, get_noise + K2 micro driver (cfg1 B=4, cfg2 B=64, K2 at three shapes)"
timeout 400 ncu --set full --clock-control none -k regex:"$KERNELS" -c 58 -f -o $OUT/prof_micro \
  env NO_GRAPH_TIMING=1 python tools/k_micro.py --k2 --iters 0 --flush write > $OUT/ncu_micro.log 2>&1
python tools/ncu_summary.py $OUT/prof_micro.ncu-rep > $OUT/ncu_micro_summary.txt 2>&1; keep_small $OUT/prof_micro.ncu-rep
echo "== ncu --set full, K5 at two UNet shapes"
K5_SHAPES=128x64,256x32 timeout 300 ncu --set full --clock-control none -k regex:"$KERNELS" -s 2 -c 2 -f -o $OUT/prof_k5 \
  python tools/k5_micro.py > $OUT/ncu_k5.log 2>&1
python tools/ncu_summary.py $OUT/prof_k5.ncu-rep > $OUT/ncu_k5_summary.txt 2>&1; keep_small $OUT/prof_k5.ncu-rep
cat $OUT/ncu_micro_summary.txt $OUT/ncu_k5_summary.txt | cut -c1-150
fi
rm -f $OUT/launches.csv
ls -la $OUT
