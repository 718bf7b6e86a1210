"""CPU tests of bench.py's contract: the reference arm (`--impl reference`: the unmodified reference package from oracle/_ref
on the host cores, or the oracle port when it is absent)
prints exactly one JSON line with the keys the driver reads, uses all host cores even when the launcher exported
OMP_NUM_THREADS=1 (torchrun does), and non-zero ranks exit without work; the product arm refuses to run without a GPU."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e, timeout=600)


def test_reference_arm_line():
    r = _run(["--impl", "reference", "--D", "2", "--chi", "6", "--steps", "1", "--warmup", "1"], env={"OMP_NUM_THREADS": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "sweeps/s" and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    from oracle import vendor_ref
    assert d["cpu_baseline"]["kind"] == ("reference" if vendor_ref.import_path() else "port") and d["cpu_baseline"]["value"] == d["value"]
    assert d["value"] == pytest.approx(1e3 / (16.0 * d["ms_per_step"]))      # a step is one site-move, a value-unit sweep is 16
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    assert d["cpu_baseline"]["cores"] == ncpu          # not the launcher's OMP_NUM_THREADS=1


def test_reference_arm_other_ranks_exit_quietly():
    r = _run(["--impl", "reference", "--gpus", "2", "--D", "2", "--chi", "6", "--steps", "1", "--warmup", "1"],
             env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_arm_needs_a_gpu():
    r = _run(["--D", "2", "--chi", "6", "--steps", "1", "--warmup", "1"])
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


def test_metric_and_config_name_the_measured_shape():
    """BASELINE.json's metric string verbatim at the headline shape; other BASELINE configs (bench.py --D --chi) spell their shape out in
    `metric` and in the L2 note, and both arms build the SAME `config` object (the driver compares them)."""
    import argparse
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    head = argparse.Namespace(D=8, chi=256, d=2)
    assert bench.metric_name(head) == bench.METRIC
    assert "D=8" in bench.METRIC and "chi=256" in bench.METRIC and str(base.get("unit", "sweeps/s")).split("/")[0][:5] in bench.METRIC + bench.UNIT
    small = argparse.Namespace(D=4, chi=64, d=2)
    assert "D=4 chi=64" in bench.metric_name(small)
    c_head, c_small = bench.shared_config(head, 4, 4), bench.shared_config(small, 2, 2)
    assert "exceed the 126 MB L2" in c_head["l2"] and "2.00 GiB" in c_head["l2"]
    assert "launch-bound" in c_small["l2"] and c_small["cell"] == "2x2" and "D=4 chi=64" in c_small["workload"]
    assert set(c_head) == set(c_small) == {"workload", "cell", "value_unit", "l2"}
