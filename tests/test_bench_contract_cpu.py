"""bench.py contract, the part that runs without a GPU: the reference arm (`--impl reference`) times
the reference's own CPU path (oracle/_ref when it was built here, else the C restatement) and prints
ONE JSON line with the keys the driver reads; under torchrun only rank 0 prints."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def run_bench(extra, env=None):
    e = dict(os.environ)
    if env:
        e.update(env)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference"] + extra,
                       capture_output=True, text=True, timeout=600, env=e, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    return [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_prints_the_contract_line():
    lines = run_bench(["--gpus", "1", "--steps", "2", "--warmup", "1"])
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "pair_interactions_per_s" and d["unit"] == "pairs/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 2 and d["gpu_launches"] == 0
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["dtype"] == "f64" and d["data"] == "synthetic"
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] == 1 and cb["value"] == d["value"]
    assert "sample" in cb and "rho=1.0" in cb["sample"]
    e2e = d["e2e"]
    assert e2e["value"] == d["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and d["config"]["rebuild_every"] == 20
    # both arms name the SAME workload: `config` is a function of the command line alone (bench_config), the
    # bounded sample the CPU arm timed is stated beside it
    assert d["config"] == _bench().bench_config(_Args(), 1)
    assert "N=1000188 (63 cells/side) rho=1.0" in d["config"]["workload"]
    assert "N=119164" in d["sample"] and d["scaling"] == "weak"


class _Args:
    density, L = 1.0, 100.1


def _bench():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_bench_config_is_the_generator_arithmetic():
    b = _bench()
    assert b.fcc_cells(1.0, 100.1) == (63, 1000188)      # BASELINE config 3
    assert b.fcc_cells(1.0, 50.0) == (31, 119164) and b.fcc_cells(0.5, 50.0) == (25, 62500)   # the reference's own boxes
    c = b.bench_config(_Args(), 8)
    assert "N=131072000 (320 cells/side) rho=0.8" in c["workload"] and c["parallelism"] == "z-slab x8"


def test_reference_arm_at_n_gpus_names_config_5_and_samples_its_density():
    lines = run_bench(["--gpus", "2", "--steps", "2", "--warmup", "1"], env={"RANK": "0", "WORLD_SIZE": "2"})
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["config"] == _bench().bench_config(_Args(), 2) and d["scaling"] == "strong" and d["n_gpus"] == 2
    assert "rho=0.8" in d["cpu_baseline"]["sample"] and "N=97556" in d["sample"]


def test_reference_arm_other_ranks_stay_silent():
    assert run_bench(["--gpus", "2", "--steps", "2", "--warmup", "1"], env={"RANK": "1", "WORLD_SIZE": "2"}) == []
