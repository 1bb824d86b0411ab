"""bench.py's reference arm end to end on the CPU (no GPU needed): one JSON line with the contract's keys, produced by
the UNMODIFIED reference model through its own API.  Also pins that both arms would print the same ``config``."""
import json
import os
import subprocess
import sys

import pytest

import clip_common as C

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not C.have_reference(), reason="reference tree not available")
def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="8")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--height", "192", "--width", "320"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "stdout must carry exactly one line: %r" % r.stdout[:500]
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "impl"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["value"] > 0
    # rank != 0 under torchrun: exits 0 without a line
    r1 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                        capture_output=True, text=True, env=dict(env, RANK="1", WORLD_SIZE="2"), timeout=120)
    assert r1.returncode == 0 and r1.stdout.strip() == ""
    # the GPU arm builds its config from the same function and the same arguments
    sys.path.insert(0, ROOT)
    import bench
    ns = type("A", (), dict(height=192, width=320, detections=40, config="GoMatching_ICDAR15"))()
    assert bench.clip_config(ns) == d["config"]


def test_gpu_arm_refuses_to_run_without_cuda():
    """north_star: no CPU fallback -- the product arm must fail loudly on a box without a GPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("this box has a GPU")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True,
                       text=True, timeout=300)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
