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


def test_metric_functions_on_random_corpora(tmp_path):
    """The unmodified compute_top_k_accuracy / compute_mrr / compute_average_similarity over the fake pgvector
    against the oracle's same-page evaluation, on random corpora with NULL pages, empty pages, duplicated embeddings
    (ties) and un-normalised rows -- beyond the one committed golden corpus."""
    import numpy as np
    from oracle import oracle
    for seed in range(6):
        rng = np.random.default_rng(100 + seed)
        N, M, D = int(rng.integers(5, 30)), int(rng.integers(20, 120)), 8
        n_pages = max(1, M // 6)
        ce = rng.standard_normal((M, D)).astype(np.float32)
        ce[rng.integers(0, M, 6)] = ce[0]                                   # duplicates -> ties
        ie = rng.standard_normal((N, D)).astype(np.float32) * rng.uniform(0.5, 3.0, (N, 1)).astype(np.float32)
        cpage = [None if rng.random() < 0.05 else int(rng.integers(0, n_pages)) for _ in range(M)]
        ipage = [None if rng.random() < 0.1 else int(rng.integers(0, n_pages + 2)) for _ in range(N)]
        cman = [f"m{(p or 0) % 3}" for p in cpage]
        iman = [f"m{(p or 0) % 3}" for p in ipage]
        t = dict(image_ids=[f"i{i}" for i in range(N)], image_manual=iman, image_page=ipage, image_emb=ie,
                 chunk_ids=[f"c{j}" for j in range(M)], chunk_manual=cman, chunk_page=cpage, chunk_emb=ce, alignments=[])
        db = rh.FakeDB({"vanilla_clip": t})
        ev, _ = rh.load_reference(db, output_dir=tmp_path)
        s = "vanilla_clip"
        want = (ev.compute_top_k_accuracy(s, [1, 5, 10]), ev.compute_mrr(s), ev.compute_average_similarity(s))
        ids, null = {}, np.uint64(0xFFFFFFFFFFFFFFFF)
        key = lambda man, page: null if page is None else np.uint64(ids.setdefault((man, page), len(ids)))
        img = dict(emb=ie, key=np.array([key(m, p) for m, p in zip(iman, ipage)], np.uint64), bbox=None, terms=None)
        chk = dict(emb=ce, key=np.array([key(m, p) for m, p in zip(cman, cpage)], np.uint64), bbox=None, terms=None)
        o = oracle.evaluate(img, chk, schema_mask=1, candidates="same_page", kmax=10, cutoff=100)
        got = oracle.metrics_from_ranks(o["pair_rank"][0], o["pair_sim"])
        assert len(ev.get_image_text_pairs(s)) == got["num_pairs"], seed
        assert want[0] == got["top_k"] and want[1] == got["mrr"] and want[2] == got["avg_similarity"], seed
