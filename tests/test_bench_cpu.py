"""bench.py's reference arm runs without a GPU (it times the oracle port on the host cores): check the
JSON contract of its line on the small C1 workload."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1",
                        "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "VB-MLP train samples/sec" and d["unit"] == "samples/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 2 and d["warmup"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == dict(value=d["value"], unit="samples/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0)


def test_flops_per_sample_matches_survey():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.flops_per_sample(bench.WORKLOADS["c1"]) == 319600
    assert bench.flops_per_sample(bench.WORKLOADS["c2"]) == 124752000
    assert bench.flops_per_sample(bench.WORKLOADS["c3"]) == 762773504               # plain nn.Linear output layer
