"""bench.py contract on CPU: the reference arm runs without a GPU, prints exactly one JSON line
with the required keys, and times the unmodified reference (oracle/_ref)."""

import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libpkref.so")):
        pytest.skip("oracle/_ref not built")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--config", "2", "--steps", "1", "--warmup", "0"],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300, check=True)
    lines = [l for l in out.stdout.decode().splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
                "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
                "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["value"] > 0
    assert d["vs_baseline"] is None and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"]


def test_reference_arm_nonzero_rank_is_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--config", "2", "--steps", "1", "--warmup", "0", "--gpus", "2"],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120, env=env, check=True)
    assert out.stdout.decode().strip() == ""


def test_flops_per_frame_matches_survey():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.flops_per_frame(bench.CONFIGS["3"]) == 17530880
    assert bench.flops_per_frame(bench.CONFIGS["4"]) == 84901888
