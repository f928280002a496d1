# 2-GPU box: the seed-parity test over NCCL + weak-scaling bench lines (N=1 and N=2) with checksums
set -x
nvidia-smi -L
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2_mgpu2_test.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_worker.py > gpurun_out/r2_mgpu2_worker.log 2>&1
timeout 900 python bench.py --gpus 1 --steps 2 --warmup 3 --no-cpu-baseline --no-reference-gpu --no-extras > gpurun_out/r2_scale_n1.json 2> gpurun_out/r2_scale_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline --no-reference-gpu --no-extras > gpurun_out/r2_scale_n2.json 2> gpurun_out/r2_scale_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --config 4 --gpus 2 --steps 1 --warmup 3 --no-cpu-baseline --no-reference-gpu --no-extras > gpurun_out/r2_scale_cfg4_n2.json 2> gpurun_out/r2_scale_cfg4_n2.err
cat gpurun_out/r2_mgpu2_test.log; grep MGPU gpurun_out/r2_mgpu2_worker.log
python - <<'PY'
import json
for f in ("r2_scale_n1", "r2_scale_n2", "r2_scale_cfg4_n2"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "n_gpus", d["n_gpus"], "value", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), d["checksum"]["rank0_shard_sha256_16"], d["checksum"].get("global_batch_sha256_16"))
    except Exception as e:
        print(f, "ERR", e)
PY
