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
    a = json.loads(one.stdout.strip().splitlines()[-1])
    b = json.loads(two.stdout.strip().splitlines()[-1])
    assert b["n_gpus"] == 2
    assert a["quality"]["num_pairs"] == b["quality"]["num_pairs"]
    for k in ("top1_vanilla", "top10_vanilla", "mrr_vanilla", "mrr_combined"):
        assert a["quality"][k] == pytest.approx(b["quality"][k], rel=1e-12), k
