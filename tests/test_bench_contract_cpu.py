"""bench.py contract, the part that runs without a GPU: the reference arm (`--impl reference`) times the CPU
port of the decode step on the host cores and prints ONE JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_json_line():
    env = {**os.environ, "OMP_NUM_THREADS": "1"}          # what torchrun exports to its workers: must be overridden
    # a small context keeps the sample to a few seconds; the keys and their meaning do not depend on it
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--batch", "4", "--ctx", "512"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "tokens/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("decode tokens/s (Llama-3-8B AWQ-int4, b=4, ctx=512")
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["n_gpus"] == 1
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and "decode_ref.c" in cb["sample"]
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count()
    assert cb["cores"] == avail                            # every host core, despite OMP_NUM_THREADS=1
    assert d["e2e"] == {"value": d["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = {**os.environ, "RANK": "1", "WORLD_SIZE": "2"}
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
