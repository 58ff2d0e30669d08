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
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 8 and d["cpu_baseline"]["sample"]   # config.lua:5
    # C1 fits: the full minibatch is timed, nothing is extrapolated, and ms_per_step is time really spent
    assert d["estimated"] is False and d["cpu_baseline"]["rows_timed"] == 100
    assert abs(d["ms_per_step"] - 1e3 * 100 / d["value"]) < 0.5 * d["ms_per_step"]
    assert d["config"]["global_batch"] == 100 and d["config"]["sizes"] == [784, 100, 10]
    assert d["e2e"] == dict(value=d["value"], unit="samples/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0)


def test_flops_per_sample_matches_survey():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.flops_per_sample(bench.WORKLOADS["c1"]) == 319600
    assert bench.flops_per_sample(bench.WORKLOADS["c2"]) == 124752000
    assert bench.flops_per_sample(bench.WORKLOADS["c3"]) == 762773504               # plain nn.Linear output layer


def test_both_arms_describe_the_same_config():
    sys.path.insert(0, ROOT)
    import bench
    for name in ("c1", "c2", "c3"):
        w = bench.WORKLOADS[name]
        a = bench.config_dict(w, 1, w["N"], "weak")
        assert a["workload"] == w["desc"] and a["global_batch"] == w["N"] and a["flops_per_sample"] == bench.flops_per_sample(w)
    c = bench.config_dict(bench.WORKLOADS["c3"], 8, 1024, "strong")
    assert c["global_batch"] == 8192 and c["per_gpu_batch"] == 1024 and c["parallelism"] == "dp8" and c["scaling"] == "strong"
