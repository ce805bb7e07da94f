"""bench.py's JSON contract, as far as a CPU-only container can exercise it: the reference arm (`--impl reference`) runs
here and must print ONE JSON line with the keys the driver reads; under a fake torchrun environment only rank 0 works;
our own arm must fail loudly without a GPU (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None, timeout=300):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout, env=e)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = _run(["--impl", "reference", "--workload", "c1", "--steps", "2", "--warmup", "1"], env={"OMP_NUM_THREADS": "1"})
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "retrieve_queries_per_sec" and j["unit"] == "queries/s"
    assert j["higher_is_better"] is True and j["steps"] == 2 and j["warmup"] == 1 and j["value"] > 0
    assert j["config"]["rows"] == 10_548 and j["config"]["dims"] == 1536 and j["config"]["k"] == 10
    cb = j["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["value"] == j["value"] and "np.dot + get_top_k" in cb["sample"]
    # torchrun exports OMP_NUM_THREADS=1; the arm must undo that and report the threads BLAS really uses
    assert cb["cores"] >= 1 and (cb["cores"] > 1 or (os.cpu_count() or 1) == 1)
    assert j["e2e"] == {"value": j["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert j["gpu_launches"] == 0 and j["vs_baseline"] is None


def test_reference_arm_other_ranks_exit_without_work():
    r = _run(["--impl", "reference", "--workload", "c1", "--gpus", "2", "--steps", "1", "--warmup", "0"],
             env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, timeout=60)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_our_arm_fails_loudly_without_a_gpu():
    try:
        import torch
        if torch.cuda.is_available():
            import pytest
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    r = _run(["--workload", "c1", "--steps", "1", "--no-cpu-baseline"], timeout=120)
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout) or "svs_b200 has no CPU path" in (r.stderr + r.stdout)
