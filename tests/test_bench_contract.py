"""bench.py contract checks that need no GPU: the reference arm (CPU restatement) prints one JSON line with the agreed
keys, and the product arm refuses to run without a CUDA device instead of falling back to anything."""
import json
import subprocess
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent


def run_bench(*args, timeout=600):
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_reference_arm_json_line():
    r = run_bench("--impl", "reference", "--workload", "ecoli", "--scale", "0.02", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, "exactly one JSON line"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "greedy_matchtig_unitigs_per_sec" and d["unit"] == "unitigs/s"
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1 and d["higher_is_better"] is True
    cb = d["cpu_baseline"]
    assert d["value"] > 0 and abs(d["value"] - cb["sample_unitigs"] / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    assert d["config"]["workload"] == "ecoli" and "model" not in d["config"]
    # recipe-level keys only: both arms must print the same config object
    assert set(d["config"]) == {"workload", "description", "k", "scale", "reader", "outputs", "candidate_cap", "l2", "parallelism"}
    assert d["config"]["outputs"] == "GFA + duplicate-kmer bitvector"
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["unit"] == d["unit"] and cb["sample"]
    # T_compute (build .. bitvector) and T_wall (parse .. GFA) are reported separately
    assert 0 < d["t_compute_ms"] < d["t_wall_ms"] and d["t_wall_ms"] == d["ms_per_step"]
    assert d["settled_nodes"]["reference_semantics"] > 0
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert d["vs_baseline"] is None and d["data"] == "synthetic"


def test_product_arm_refuses_to_run_without_a_gpu():
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = run_bench("--workload", "ecoli", "--scale", "0.02", "--steps", "1", "--warmup", "3")
    assert r.returncode != 0, "the product path must fail loudly without its CUDA device"
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")], "no bench line may be printed from a CPU-only run"
