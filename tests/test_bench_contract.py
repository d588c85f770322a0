"""bench.py's JSON contract, checked on the CPU through the reference arm (the GPU arm needs a B200 and is
exercised by tests/test_gpu_nccl.py): one JSON line with the keys the driver reads, the same `metric` and
`config` strings the GPU arm prints, a cpu_baseline block and an e2e block with zero copy bytes."""
import json
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--N", "600", "--M", "2000", "--steps", "1",
                          "--warmup", "1", "--cpu-rows", "64"], cwd=ROOT, capture_output=True, text=True, check=True)
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "impl"):
        assert k in j, k
    assert j["impl"] == "reference" and j["unit"] == "queries/s" and j["higher_is_better"] is True
    assert j["value"] > 0 and j["vs_baseline"] is None and j["data"] == "synthetic"
    assert j["config"]["N"] == 600 and j["config"]["M"] == 2000 and "workload" in j["config"]
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and "64 query rows" in cb["sample"]
    assert j["e2e"] == {"value": j["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_ranks_other_than_zero_stay_silent_in_the_reference_arm():
    import os
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--gpus", "2", "--N", "600", "--M", "2000"],
                         cwd=ROOT, capture_output=True, text=True, env=env, check=True)
    assert out.stdout.strip() == ""


def test_throttled_runs_are_rejected():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", ROOT / "bench.py")
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    ok = {"sm_mhz": 1575.0, "sm_max_mhz": 1965.0, "reasons": ["sw_power_cap"]}
    assert not b.clocks_rejected(ok) and not b.clocks_rejected(None)
    assert not b.clocks_rejected({"sm_mhz": 1950.0, "sm_max_mhz": 1965.0, "reasons": []})
    assert b.clocks_rejected(dict(ok, reasons=["sw_power_cap", "hw_thermal_slowdown"]))
    assert b.clocks_rejected(dict(ok, reasons=["sw_thermal_slowdown"])) and b.clocks_rejected(dict(ok, reasons=["hw_slowdown"]))
    assert b.clocks_rejected({"sm_mhz": 900.0, "sm_max_mhz": 1965.0, "reasons": []})      # clock lock left behind
    assert not b.clocks_rejected({"sm_mhz": None, "sm_max_mhz": 1965.0, "reasons": ["no samples: x"]})


def test_core_binding_is_optional():
    """bench.py binds a rank to the cores NVML lists as local to its GPU; without NVML or a GPU it leaves the process alone."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("bench_mod2", ROOT / "bench.py")
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    before = os.sched_getaffinity(0)
    n = b.bind_to_gpu_cores(0)
    assert n == 0 or n == len(os.sched_getaffinity(0))
    if n == 0:
        assert os.sched_getaffinity(0) == before
