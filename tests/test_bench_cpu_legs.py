"""The CPU legs of bench.py (cpu_baseline, --impl reference, the rollout leg's sampler-equivalent worker) run without a
GPU; this keeps them working.  bench.py is the one place besides tests/ and smoke() that may execute oracle/."""
import importlib.util
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_cpu_oracle_and_sampler_equivalent_legs():
    b = _bench()
    v, dt = b.cpu_oracle_throughput(3, 2000, 2)
    assert v > 1000 and dt > 0
    w = b.cpu_sampler_equivalent(3, seconds=0.5)
    assert 1 < w < v          # a batch-1 forward per agent and step is far slower than the env step itself


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-sample-steps", "16000"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "env-steps/sec" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0 and d["gpu_launches"] == 0
