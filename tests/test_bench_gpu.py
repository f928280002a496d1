"""bench.py's own arm on the GPU at a tiny size: the JSON line carries every key of the contract and the per-config
workloads run.  (Sizes are far below the BASELINE configs: this checks the plumbing, not the numbers.)"""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*extra):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "3", "--nb-steps", "3",
                          "--batch", "2", "--no-cpu-baseline", "--ref-gpu-steps", "2", *extra],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, (out.stdout + out.stderr)[-3000:]
    return json.loads(out.stdout.strip().splitlines()[-1])


def test_bench_line_has_the_contract_keys():
    line = _run()
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "roofline_get_noise", "roofline_step",
                "reference_gpu", "checksum", "get_noise", "unet"):
        assert key in line, key
    assert line["unit"] == "images/s" and line["higher_is_better"] is True and line["scaling"] == "weak" and line["vs_baseline"] is None
    assert line["value"] > 0 and line["e2e"]["value"] > 0 and line["gpu_launches"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 2 * 3 * 64 * 64 * 4 == line["e2e"]["d2h_bytes_per_step"]
    r = line["roofline"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in r, key
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    g = line["roofline_get_noise"]
    assert g["cfg1"]["kernel"].startswith("gemv_kernel") and 0 < g["cfg1"]["frac"] < 1
    assert line["reference_gpu"]["value"] > 0 and line["reference_gpu"]["steps_run"] == 2
    assert line["checksum"]["rank0_shard_sha256_16"] == line["checksum"]["global_batch_sha256_16"]      # one GPU: the same batch
    assert "workload" in line["config"] and line["config"]["batch_per_gpu"] == 2


@pytest.mark.parametrize("cfg", [3, 4, 5])
def test_other_baseline_configs_run(cfg):
    line = _run("--config", str(cfg), "--no-extras", "--no-reference-gpu")
    assert line["value"] > 0 and f"configs[{cfg - 1}]" in line["config"]["workload"]
