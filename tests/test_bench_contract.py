"""bench.py's reference arm runs on CPU: check its JSON line against the contract (one line,
the keys the driver reads, rank 0 only under torchrun-style environments)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
        "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


def run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                         env=e, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return [l for l in out.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_c1_line():
    lines = run(["--impl", "reference", "--workload", "c1", "--steps", "1", "--warmup", "0"])
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert KEYS <= set(d), KEYS - set(d)
    assert d["impl"] == "reference" and d["metric"] == "first_hit_Mrays_per_s" and d["unit"] == "Mrays/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    assert "workload" in d["config"] and "MarchingCubesSearch" in d["config"]["workload"]


def test_reference_arm_other_ranks_print_nothing():
    assert run(["--impl", "reference", "--workload", "c1", "--steps", "1"], env={"RANK": "1", "WORLD_SIZE": "2"}) == []


def test_reference_arm_path_tracing_line():
    """The path-tracing reference arm (oracle renderer on a bounded frame) keeps the same contract."""
    lines = run(["--impl", "reference", "--workload", "c4", "--steps", "1", "--warmup", "0"])
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert KEYS <= set(d), KEYS - set(d)
    assert d["metric"] == "path_traced_Msamples_per_s" and d["unit"] == "Msamples/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and "spp" in d["cpu_baseline"]["sample"]
    assert "showcase" in d["config"]["workload"]
