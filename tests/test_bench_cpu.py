"""CPU: the bench contract's reference arm runs without a GPU (it times the CPU restatement of the reference loop),
and our own arm refuses to run without one (no CPU fallback behind the headline number)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*flags, **kw):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(flags), capture_output=True, text=True,
                          cwd=ROOT, env=env, timeout=kw.get("timeout", 600))


def test_reference_arm_prints_the_contract_line():
    out = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-steps", "20", "--demos", "300")
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "env-steps/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["steps"] == 1 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and "Maze" in d["metric"]


def test_own_arm_fails_loudly_without_a_gpu():
    out = _run("--steps", "1", "--warmup", "1", "--no-cpu-baseline", timeout=300)
    assert out.returncode != 0
    assert not [l for l in out.stdout.splitlines() if l.startswith("{") and '"value"' in l]
    assert "CUDA" in out.stderr or "cuda" in out.stderr
