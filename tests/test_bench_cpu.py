"""CPU tests of bench.py's driver contract: the reference arm prints ONE JSON line with the keys the driver reads
(tiny model here; the full-size arm is the same code path), non-zero ranks of a torchrun launch exit without work, and
our arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    e["CUDA_VISIBLE_DEVICES"] = ""
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                          cwd=ROOT, env=e, timeout=300)


def test_reference_arm_prints_the_contract_line():
    r = _run(["--impl", "reference", "--tiny", "--steps", "2", "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "noisy_vlm_samples_per_sec" and d["unit"] == "samples/s"
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["gpu_launches"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb["sample"]
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["scaling"] in ("weak", "strong") and d["dtype"] == "f32" and d["data"] == "synthetic"


def test_reference_arm_runs_on_rank_0_only():
    r = _run(["--impl", "reference", "--tiny", "--gpus", "2", "--steps", "1", "--warmup", "0"],
             env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_our_arm_needs_a_gpu():
    r = _run(["--tiny", "--steps", "1", "--warmup", "1"])
    assert r.returncode != 0 and "no CPU fallback" in r.stderr
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]
