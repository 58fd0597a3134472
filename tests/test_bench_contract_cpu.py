"""The reference arm of bench.py runs on host cores only: its one-line JSON contract is checked here."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "denoising_system_steps_per_sec"
    assert d["unit"] == "system*steps/s" and d["higher_is_better"] is True and d["scaling"] == "strong"
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1 and d["vs_baseline"] is None
    assert "workload" in d["config"] and not any(k in d["config"] for k in ("model", "seq_len", "global_batch"))
    cb, e2e = d["cpu_baseline"], d["e2e"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert e2e["value"] == d["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
