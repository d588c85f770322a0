"""The BLAS port (CPU baseline of bench.py) against the bit-exact C oracle.  No GPU."""
import numpy as np

from oracle import numpy_port


def test_numpy_port_matches_oracle(oracle, synthetic):
    img, chk, _ = synthetic.make_numpy(96, 3000, 128, T=512, seed=21)
    lam = (0.3, 0.2, 0.5)
    o = oracle.evaluate(img, chk, T=512, schema_mask=15, candidates="all", lam=lam, kmax=20, cutoff=100)
    p = numpy_port.evaluate(img, chk, T=512, schemas=(0, 1, 2, 3), lam=lam, kmax=20, cutoff=100, slab=1024)
    assert np.array_equal(p["pair_chunk"], o["pair_chunk"])
    assert np.abs(p["pair_sim"] - o["pair_sim"]).max() < 1e-6      # BLAS summation order differs
    assert np.abs(p["topk_score"] - o["topk_score"]).max() < 1e-6
    assert (p["topk_idx"] == o["topk_idx"]).mean() > 0.999           # only near-ties may swap
    assert (p["pair_rank"] == o["pair_rank"]).mean() > 0.999


def test_numpy_port_row_sample(oracle, synthetic):
    img, chk, _ = synthetic.make_numpy(64, 1000, 64, T=64, seed=22)
    rows = np.array([3, 10, 11, 40])
    p = numpy_port.evaluate(img, chk, T=64, schemas=(0, 3), lam=(0.1, 0.1, 0.2), kmax=10, cutoff=100, rows=rows)
    o = oracle.evaluate(img, chk, T=64, schema_mask=9, candidates="all", lam=(0.1, 0.1, 0.2), kmax=10, cutoff=100)
    assert (p["topk_idx"] == o["topk_idx"][:, rows]).mean() > 0.99
