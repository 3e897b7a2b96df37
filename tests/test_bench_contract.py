"""bench.py prints ONE JSON line carrying every key of the measurement contract (GPU box only)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*extra):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--stars", "4", "--steps", "2", "--warmup", "1",
                          *extra], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


@pytest.mark.gpu
def test_bench_line_has_contract_keys():
    d = _run("--cpu-sample-stars", "1")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1 and d["dtype"] == "f64" and d["value"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    # the serial end-to-end loop cannot beat the kernels-only figure; with two batches in flight (e2e.mode) the tail of one
    # batch's launches overlaps the other's, so only the serial entry is bounded
    assert 0 < d["e2e"]["one_batch_at_a_time"]["value"] <= d["value"] * 1.05
    assert d["e2e"]["value"] >= d["e2e"]["one_batch_at_a_time"]["value"] and "mode" in d["e2e"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in d["roofline"], k
    assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-12
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in d["cpu_baseline"], k
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["value"] > 0
    for k in ("sm_mhz", "sm_max_mhz", "reasons"):
        assert k in d["clocks"], k
    assert d["gpu_launches"] >= d["steps"]
    assert d["all_stars_converged"] is True


def test_bench_reference_arm():
    """The reference arm is CPU-only: committed sample inputs (bench_data/), the oracle port, no product library."""
    d = _run("--impl", "reference", "--cpu-sample-stars", "2")
    assert d["product_library_loaded"] is False
    assert d["impl"] == "reference" and d["value"] > 0 and d["metric"] == "zone_microphysics_evals_per_sec"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1


def test_e2e_entry_reports_both_loops():
    """bench.e2e_entry: `value` is the faster of the serial loop and the loop with two batches in flight, `mode` says
    which, and both figures stay in the line."""
    sys.path.insert(0, ROOT)
    import bench
    e = bench.e2e_entry(1.0e6, 100.0, 2.0e-3, 1.6e-3, 10, 20)
    assert e["value"] == 1.0e6 / 1.6e-3 and "two batches" in e["mode"] and e["h2d_bytes_per_step"] == 10
    assert e["one_batch_at_a_time"]["ms_per_step"] == 2.0 and abs(e["two_batches_in_flight"]["ms_per_step"] - 1.6) < 1e-12
    e = bench.e2e_entry(1.0e6, 100.0, 2.0e-3, None, 10, 20)
    assert e["value"] == 1.0e6 / 2.0e-3 and "two_batches_in_flight" not in e and e["mode"] == "one batch at a time"
    e = bench.e2e_entry(1.0e6, 100.0, 2.0e-3, 2.5e-3, 10, 20)
    assert e["value"] == 1.0e6 / 2.0e-3 and e["mode"] == "one batch at a time" and "two_batches_in_flight" in e
