"""bench.py's reference arm runs anywhere (CPU only) and prints the contract's JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-batch", "1", "--cpu-steps", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "images/s" and line["higher_is_better"] is True
    for key in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config",
                "cpu_baseline", "e2e", "gpu_launches"):
        assert key in line, key
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["gpu_launches"] == 0
    assert "workload" in line["config"]
    # the reference arm's config describes what it RAN (bounded sample), not the GPU arm's batch
    assert line["config"]["batch_per_gpu"] == 1 and line["config"]["nb_steps_run"] == 1
    assert line["config"]["extrapolation_factor"] == 250 and "host CPU" in line["config"]["device"]
    # ms_per_step is the measured wall time of one bounded sample: the whole run must fit the time the driver saw
    assert line["ms_per_step"] * line["steps"] / 1e3 < 120


def test_reference_arm_covers_the_other_configs():
    for cfg, T in ((5, 250), (3, 100)):
        out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", str(cfg), "--steps", "1",
                              "--warmup", "0", "--cpu-batch", "1", "--cpu-steps", "2"], capture_output=True, text=True, timeout=600,
                             cwd=ROOT)
        assert out.returncode == 0, out.stderr[-2000:]
        line = json.loads(out.stdout.strip().splitlines()[-1])
        assert line["config"]["nb_steps"] == T and line["value"] > 0 and f"configs[{cfg - 1}]" in line["config"]["workload"]


def test_non_zero_rank_of_the_reference_arm_exits_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True,
                         text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
