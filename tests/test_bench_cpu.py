"""bench.py's contract pieces that run without a GPU: the reference arm (`--impl reference`, the unmodified lambda3
on the host cores) at a tiny size, and the clock sampler's behaviour on a box without NVML."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "lambda3")


@pytest.mark.skipif(not os.path.exists(REF), reason="reference binary (oracle/_ref/lambda3) not built")
def test_reference_arm_prints_one_json_line(tmp_path):
    env = dict(os.environ, LAMBDA_B200_CACHE=str(tmp_path))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--n-seqs", "2000",
                        "--n-queries", "200", "--steps", "2", "--warmup", "1"], capture_output=True, text=True, env=env,
                       check=True)
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "searchp_query_seqs_per_s" and d["unit"] == "queries/s"
    assert d["value"] > 0 and d["steps"] == 2 and d["warmup"] == 1 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("searchp: 200x300aa")


def test_reference_arm_other_ranks_do_nothing(tmp_path):
    env = dict(os.environ, LAMBDA_B200_CACHE=str(tmp_path), RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_clock_sampler_without_nvml():
    sys.path.insert(0, ROOT)
    import bench
    s = bench.ClockSampler(0)
    s.start()
    s.begin()
    s.end()
    out = s.summary()
    assert set(out) >= {"sm_mhz", "sm_max_mhz", "reasons", "samples", "source"}
    assert out["samples"] == 0 or out["sm_mhz"] > 0
