"""CPU: the parts of bench.py's contract that do not need a GPU -- the reference arm prints exactly ONE JSON line with the
required keys (timing the UNMODIFIED reference staged by oracle/stage_ref.py on the host cores; the oracle port only when
nothing was staged), honours --steps / --warmup and prints its true batch size, other ranks of a torchrun launch stay silent and exit 0, and the
product arm refuses to run without a CUDA device instead of falling back to anything."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e, timeout=600)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = _run(["--impl", "reference", "--steps", "2", "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["steps"] == 2 and d["warmup"] == 1                                   # the driver's K / W are honoured
    assert d["value"] > 0 and abs(d["value"] - 4 / (d["ms_per_step"] / 1e3)) < 2e-2 * d["value"]
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["config"]["batch_per_gpu"] == 4 and d["config"]["global_batch"] == 4 and "sample" in d["config"]   # its TRUE batch
    cb = d["cpu_baseline"]
    from oracle import ref_runner
    assert cb["kind"] == ("reference" if ref_runner.available() else "port")
    assert cb["cores"] == (os.cpu_count() or 1) and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = _run(["--steps", "1", "--warmup", "3"])
    assert r.returncode != 0 and r.stdout.strip() == "" and "CUDA" in (r.stderr + r.stdout)
