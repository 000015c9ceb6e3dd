"""CPU-side checks of bench.py's contract (no GPU): the reference arm prints ONE JSON line with the agreed keys, the
host-buffer helper is harmless without NVML, and non-zero ranks of the reference arm stay silent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_numa_local_is_a_noop_without_nvml():
    import bench
    before = os.sched_getaffinity(0)
    with bench.numa_local(0) as n:
        assert isinstance(n.state, str) and n.state
    assert os.sched_getaffinity(0) == before


def _run(env_extra, *args):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], env=env, stdout=subprocess.PIPE,
                          stderr=subprocess.PIPE, text=True, timeout=600)


def test_reference_arm_prints_one_json_line():
    r = _run({"RANK": "0", "WORLD_SIZE": "1"}, "--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("MMSE channel estimates/sec") and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("configs[2]")


def test_reference_arm_other_ranks_exit_silently():
    r = _run({"RANK": "1", "WORLD_SIZE": "2"}, "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0 and r.stdout.strip() == ""
