"""bench.py's reference arm (`--impl reference`: the reference's own code (oracle/_ref) on the host cores, the driver's baseline run) on a
tiny configuration -- CPU only: it must print ONE JSON line with the contract's keys and run without a GPU.  Also the
torchrun case: under WORLD_SIZE > 1 rank 0 alone reports, the other ranks exit quietly."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ARGS = ["--impl", "reference", "--steps", "1", "--warmup", "0", "--ring", "2"]


def _run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + ARGS, capture_output=True, text=True, env=env,
                          timeout=300)


def test_reference_arm_prints_the_contract_line():
    res = _run()
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "detect_frames_per_s" and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_under_torchrun_only_rank0_reports():
    other = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert other.returncode == 0 and other.stdout.strip() == ""
    first = _run({"RANK": "0", "LOCAL_RANK": "0", "WORLD_SIZE": "2"})
    assert first.returncode == 0 and json.loads(first.stdout.strip().splitlines()[-1])["impl"] == "reference"
