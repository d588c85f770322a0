"""Re-runs the unmodified reference (where /root/reference exists) and checks the
committed golden vectors were not edited by hand.  Skipped on the GPU box."""
import json

import pytest

from conftest import GOLDEN, unhex
from oracle import reference_harness as rh

pytestmark = pytest.mark.skipif(not rh.available(), reason="/root/reference is not on this machine")


def test_weak_vectors_still_match_reference():
    _, ins = rh.load_reference()
    g = json.loads((GOLDEN / "weak_vectors.json").read_text())
    for c in g["positional"][:200]:
        assert ins.compute_positional_alignment({"bbox": c["image"]}, {"bbox": c["chunk"]}) == unhex(c["expect"])
    for c in g["lexical_text"]:
        assert ins.compute_lexical_alignment({"text": c["text"]}, c["terms"]) == unhex(c["expect"])


def test_reference_tree_untouched():
    assert not (rh.REF / "evaluation_results").exists()


def test_oracle_fuzz_against_the_live_reference():
    """Beyond the committed goldens: thousands of random inputs through the reference's own functions and the
    oracle (H6, H7, H8).  Positional: bit-exact except where CPython's pow(x, 2) is not the rounded square (the
    distance branch; 1 ulp allowed there, as in test_oracle_golden.py)."""
    import math
    import numpy as np
    from oracle import oracle
    _, ins = rh.load_reference()
    rng = np.random.default_rng(2024)

    def box():
        kind = rng.integers(0, 10)
        if kind == 0:
            return None
        if kind == 1:
            return [0.0, 0.0, 0.0, 0.0]
        if kind == 2:
            return rng.uniform(0, 800, 3).tolist()            # wrong length
        x0, y0 = rng.uniform(0, 600), rng.uniform(0, 780)
        w, h = rng.uniform(-5, 200), rng.uniform(-5, 200)     # some inverted / zero-size boxes
        if kind == 3:
            w = 0.0
        b = [x0, y0, x0 + w, y0 + h]
        return [float(np.float32(v)) for v in b] if kind >= 7 else b
    n_dist = 0
    for _ in range(6000):
        ib, cb = box(), box()
        want = ins.compute_positional_alignment({"bbox": ib} if ib is not None else {}, {"bbox": cb} if cb is not None else {})
        got = oracle.positional(ib, cb)
        if got != want:
            assert math.isclose(got, want, rel_tol=0, abs_tol=2.3e-16), (ib, cb, got, want)
            n_dist += 1
    assert n_dist < 600  # the pow() corner is rare
    alphabet = list("abcAB -") + ["é", "ß", "Ж"]
    for _ in range(400):
        text = "".join(rng.choice(alphabet, size=int(rng.integers(0, 60))))
        terms = ["".join(rng.choice(alphabet, size=int(rng.integers(0, 4)))).lower() for _ in range(int(rng.integers(0, 40)))]
        want = ins.compute_lexical_alignment({"text": text}, terms)
        bits = oracle.term_bitsets([text], terms) if terms else np.zeros((1, 1), np.uint64)
        hits = int(np.unpackbits(bits.view(np.uint8)).sum())
        assert oracle.lexical(hits, len(terms)) == want, (text, terms)
