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
