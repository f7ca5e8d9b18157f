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


FAKE_ENGINE = r"""
import sys, numpy as np, torch
sys.path.insert(0, ROOT)
import lambda_b200, bench
from lambda_b200._abi import HIT_DT, STATS_DT

torch.cuda.set_device = lambda *a, **k: None
torch.cuda.synchronize = lambda *a, **k: None
torch.Tensor.pin_memory = lambda self: self
torch.Tensor.cuda = lambda self, *a, **k: self

class FakeIndex:
    device_bytes = 123456789
    @staticmethod
    def load(path, device=0, keep_ids=True):
        return FakeIndex()

class FakeSearcher:
    def __init__(self, ix, domain, streams=None, **kw):
        self.kw = kw
    def search(self, res, offs, copy=True):
        n = (offs.numel() if hasattr(offs, "numel") else len(offs)) - 1
        hits = np.zeros(n, HIT_DT)
        hits["q_id"] = np.arange(n)
        hits["aln_len"] = 1
        st = np.zeros(1, STATS_DT)[0]
        for k, v in dict(ms_total=10.0, ms_seed=1.0, ms_sort_merge=0.5, ms_extend_score=5.0, ms_extend_trace=3.0, ms_h2d=0.1,
                         ms_host=0.4, cells_score=4e9, cells_trace=3e8, kernel_launches=120, hits_final=n,
                         n_extensions_score=5 * n, n_extensions_trace=n).items():
            st[k] = v
        return hits, st
    def m8(self, hits, ids):
        return ["fake\n" for _ in range(len(hits))]

lambda_b200.Index = FakeIndex
lambda_b200.Searcher = FakeSearcher
sys.argv = ["bench.py"] + ARGS
bench.main()
"""


@pytest.mark.skipif(not os.path.exists(REF), reason="reference binary (oracle/_ref/lambda3) not built")
@pytest.mark.parametrize("extra", [[], ["--band", "32"], ["--workload", "searchn", "--n-seqs", "3"]])
def test_our_arm_assembles_the_contract_line_with_a_fake_engine(tmp_path, extra):
    """everything of bench.py's own arm except the CUDA library: workload generation, timing loop, clocks, roofline
    objects, cpu_baseline, JSON keys -- with Index / Searcher replaced by stand-ins returning fixed stage times"""
    env = dict(os.environ, LAMBDA_B200_CACHE=str(tmp_path))
    args = ["--n-seqs", "2000", "--n-queries", "300", "--steps", "2", "--warmup", "1", "--cpu-sample", "100"] + extra
    code = f"ROOT = {ROOT!r}\nARGS = {args!r}\n" + FAKE_ENGINE
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "roofline", "roofline_trace", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["vs_baseline"] is None
    assert d["ms_per_step"] == pytest.approx(10.0) and d["value"] == pytest.approx(300 / 0.010)
    assert d["gpu_launches"] == 240
    rf = d["roofline"]
    assert rf["achieved"] == pytest.approx(4e9 / 5e-3 / 1e9 * 10) and rf["frac"] == pytest.approx(rf["achieved"] / rf["peak"])
    assert rf["traffic"] is None or isinstance(rf["traffic"], float)
    rt = d["roofline_trace"]
    assert rt["bound"] == "int16-alu" and rt["achieved"] == pytest.approx(3e8 / 3e-3 / 1e9 * 10) and 0 < rt["frac"] < 1
    assert rt["hbm"]["achieved"] == pytest.approx(1 * 3e8 / 3e-3 / 1e9) and 0 < rt["hbm"]["frac"] < 1  # 1 B per cell
    assert d["kernel_sources_sha"] and d["scaling"] == "weak"
    assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} and d["e2e"]["h2d_bytes_per_step"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    if "--band" in extra:
        assert d["config"]["window_band"] == 32 and d["parity_sample"] is None and "cpu_baseline" not in d
    else:
        assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] > 0
        assert d["parity_sample"]["identical"] is False  # the stand-in engine returns nonsense, of course
