"""CPU: bench.py's driver contract that does not need a GPU — the reference arm prints one JSON line with the agreed
keys, and the product arm refuses to run (loudly, non-zero) on a box without a CUDA device: no CPU fallback."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-1000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "simulated_games_per_sec" and d["unit"] == "games/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    # the real Python reference when it can be imported (/root/reference here, oracle/_ref byte code on the GPU box),
    # else the C port of its algorithm; the port's rate is reported beside it either way
    import ref_harness

    assert d["cpu_baseline"]["kind"] == ("reference" if ref_harness.reference_available() else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["cpu_port"]["kind"] == "port" and d["cpu_port"]["value"] > d["value"] * (10 if d["cpu_baseline"]["kind"] == "reference" else 0.2)
    assert d["config"]["matchups"] == 1 and d["config"]["sims_per_matchup_per_gpu_per_step"] == 10_000_000
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_product_arm_has_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        return   # (on the GPU box the product arm is exercised by the driver itself)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "3"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stdout + r.stderr)
