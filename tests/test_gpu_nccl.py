"""Real NCCL run of bench.py on 2 GPUs at a reduced size (skipped with fewer than 2 GPUs)."""
import json
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900, method="thread")]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_bench_matches_one_gpu():
    args = ["--N", "20000", "--M", "100000", "--steps", "1", "--warmup", "1", "--no-cpu-baseline"]
    one = subprocess.run([sys.executable, "bench.py"] + args, cwd=ROOT, capture_output=True, text=True, check=True)
    two = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29531", "bench.py", "--gpus", "2"] + args,
                         cwd=ROOT, capture_output=True, text=True)
    assert two.returncode == 0, two.stderr[-3000:]
    last = lambda out: json.loads([l for l in out.strip().splitlines() if l.startswith("{")][-1])
    a, b = last(one.stdout), last(two.stdout)
    assert b["n_gpus"] == 2
    assert a["quality"]["num_pairs"] == b["quality"]["num_pairs"]
    for k in ("top1", "top10", "mrr_first_schema", "mrr_last_schema"):
        assert a["quality"][k] == pytest.approx(b["quality"][k], rel=1e-12), k
    # both runs checked sampled rows of their timed result against the oracle, bit for bit
    assert a["verified_rows"]["ok"] and b["verified_rows"]["ok"] and b["verified_rows"]["rows"] > 0
    assert b["e2e"]["same_result_as_device_arm"]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("exchange", ["alltoall", "none"])
def test_two_gpu_exchanges(exchange):
    """Both layouts over real NCCL (the default picks by size): column shards + list all-to-all, and query slabs with
    the prepared-operand all-gather (fp32 master rows on a side stream)."""
    args = ["--N", "40000", "--M", "60000", "--steps", "1", "--warmup", "1", "--no-cpu-baseline", "--exchange", exchange]
    two = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", "bench.py", "--gpus", "2"] + args,
                         cwd=ROOT, capture_output=True, text=True)
    assert two.returncode == 0, two.stderr[-3000:]
    b = json.loads([l for l in two.stdout.strip().splitlines() if l.startswith("{")][-1])
    assert b["verified_rows"]["ok"] and b["e2e"]["same_result_as_device_arm"]
