"""The CPU oracle against the golden vectors made from the unmodified reference
(tests/golden/make_golden.py).  No GPU."""
import hashlib
import json

import numpy as np
import pytest

from conftest import random_texts_and_terms, GOLDEN, unhex


def test_positional_known_answers(oracle):
    cases = json.loads((GOLDEN / "weak_vectors.json").read_text())["positional"]
    assert len(cases) > 600
    for c in cases:
        got = oracle.positional(c["image"], c["chunk"])
        want = unhex(c["expect"])
        # bit-exact except where CPython's pow(x, 2) (glibc, < 1 ulp) is not the rounded square
        assert got == want or abs(got - want) <= 2.3e-16, c


def test_positional_appendix_a(oracle):
    # SURVEY.md Appendix A, P1..P13
    assert oracle.positional([0, 0, 10, 10], [5, 5, 15, 15]) == 0.14285714285714285
    assert oracle.positional([0, 0, 10, 10], [10, 0, 20, 10]) == 0.99
    assert oracle.positional([0, 0, 10, 10], [100, 100, 110, 110]) == 0.8585786437626906
    assert oracle.positional([0, 0, 10, 10], [2000, 0, 2010, 10]) == 0.0
    assert oracle.positional(None, [0, 0, 10, 10]) == 0.0
    assert oracle.positional([0, 0, 10], [0, 0, 10, 10]) == 0.0
    assert oracle.positional([10, 0, 0, 10], [2, 2, 8, 8]) == 1.0
    assert oracle.positional([72.0, 120.5, 300.25, 410.75], [72.0, 430.0, 523.3, 442.1]) == 0.7963274533718401


def test_lexical_known_answers(oracle):
    cases = json.loads((GOLDEN / "weak_vectors.json").read_text())["lexical"]
    for c in cases:
        assert oracle.lexical(c["hits"], c["T"]) == unhex(c["expect"]), c
    assert oracle.lexical(1, 200) == 0.05  # Appendix A L11: not > 0.05


def test_lexical_substring_semantics(oracle):
    """The oracle's term sets against the reference's own compute_lexical_alignment answers (golden), and against
    Python's `in` on random unicode strings (what src/insert_clip_embeddings.py:150 evaluates)."""
    cases = json.loads((GOLDEN / "weak_vectors.json").read_text())["lexical_text"]
    for c in cases:
        bits = oracle.term_bitsets([c["text"]], c["terms"])
        hits = int(np.unpackbits(bits.view(np.uint8)).sum())
        assert oracle.lexical(hits, len(c["terms"])) == unhex(c["expect"]), c
    texts, terms = random_texts_and_terms(np.random.default_rng(5), 60, 150)
    bits = oracle.term_bitsets(texts, terms)
    for j, text in enumerate(texts):
        low = text.lower()
        for t, term in enumerate(terms):
            assert bool((int(bits[j, t >> 6]) >> (t & 63)) & 1) == (term in low), (j, t)
    assert bits.shape == (60, 3) and not (bits[:, 2] >> np.uint64(150 - 128)).any()  # no bits beyond T


def test_cosine_orders_agree(oracle):
    rng = np.random.default_rng(0)
    for D in (4, 64, 100, 512, 768, 1024):
        a, b = rng.standard_normal(D).astype(np.float32), rng.standard_normal(D).astype(np.float32)
        c, s = oracle.cosine(a, b), oracle.cosine(a, b, sequential=True)
        ref = float(np.dot(a.astype(np.float64), b.astype(np.float64)) /
                    np.sqrt(np.dot(a.astype(np.float64), a.astype(np.float64)) * np.dot(b.astype(np.float64), b.astype(np.float64))))
        assert abs(c - s) < 1e-6 and abs(c - ref) < 1e-6  # pgvector's loop order is compiler-defined
        assert oracle.cosine(a, a) == pytest.approx(1.0, abs=1e-6)


def _oracle_same_page(oracle, corpus, mask=1, kmax=100):
    return oracle.evaluate(corpus.img, corpus.chk, T=corpus.n_terms, schema_mask=mask, candidates="same_page",
                           kmax=kmax, cutoff=max(kmax, 100))


def test_small_corpus_pairs_topk_metrics(oracle, small_corpus):
    d, c = small_corpus
    exp = d["expect"]
    r = _oracle_same_page(oracle, c)
    off, pc = r["pair_offsets"], r["pair_chunk"]
    pairs = [[c.image_ids[i], c.chunk_ids[pc[p]], c.image_manual[i], c.image_page[i]]
             for i in range(len(c.image_ids)) for p in range(off[i], off[i + 1])]
    assert pairs == exp["pairs"]
    for i, iid in enumerate(c.image_ids):
        got = [(c.chunk_ids[j], s) for j, s in zip(r["topk_idx"][0, i, :10], r["topk_score"][0, i, :10]) if j >= 0]
        want = [(cid, unhex(s)) for cid, s in exp["top10"][iid]]
        assert got == want, iid
        got100 = [c.chunk_ids[j] for j in r["topk_idx"][0, i] if j >= 0]
        assert got100 == exp["top100"][iid]
    assert [unhex(x) for x in exp["pair_similarity"]] == r["pair_sim"].tolist()
    m = oracle.metrics_from_ranks(r["pair_rank"][0], r["pair_sim"], k_values=(1, 5, 10, 20))
    assert {str(k): v for k, v in m["top_k"].items()} == {k: unhex(v) for k, v in exp["top_k_1_5_10_20"].items()}
    assert m["mrr"] == unhex(exp["mrr"])
    assert m["avg_similarity"] == unhex(exp["avg_similarity"])
    # metrics.json: all four schemas rank identically in the reference
    ref = json.loads(exp["metrics_json"])
    m3 = oracle.metrics_from_ranks(r["pair_rank"][0], r["pair_sim"])
    for s in ("vanilla_clip", "clip_lexical", "clip_positional", "clip_combined"):
        assert ref[s]["top_k"] == {str(k): v for k, v in m3["top_k"].items()}
        assert ref[s]["mrr"] == float(m3["mrr"]) and ref[s]["avg_similarity"] == float(m3["avg_similarity"])
        assert ref[s]["num_pairs"] == m3["num_pairs"]


def test_small_corpus_alignments(oracle, small_corpus):
    d, c = small_corpus
    names = ("lexical", "positional", "combined")
    for si, schema in enumerate(("vanilla_clip", "clip_lexical", "clip_positional", "clip_combined")):
        want = [(a, b, unhex(s), t) for a, b, s, t in d["expect"]["alignments"][schema]]
        # the reference's Python loop joins page None with page None (the SQL join does not): key_py
        assert si in (0, 2) or any("pNone" in w[0] for w in want)
        off, pc, rec = oracle.alignments(dict(c.img, key=c.img["key_py"]), dict(c.chk, key=c.chk["key_py"]),
                                         T=c.n_terms, schema=si)
        got = [(c.image_ids[i], c.chunk_ids[pc[p]], float(rec[p, t]), names[t])
               for i in range(len(c.image_ids)) for p in range(off[i], off[i + 1]) for t in range(3) if rec[p, t] != 0.0]
        assert len(got) == len(want)
        for g, w in zip(got, want):
            assert g[:2] == w[:2] and g[3] == w[3]
            assert g[2] == w[2] or abs(g[2] - w[2]) <= 2.3e-16


def test_config1_metrics(oracle, synthetic):
    g = json.loads((GOLDEN / "config1.json").read_text())
    img, chk, meta = synthetic.make_numpy(g["N"], g["M"], g["D"], seed=g["seed"])
    h = hashlib.sha256()
    for a in (img["emb"], chk["emb"], img["key"], chk["key"]):
        h.update(np.ascontiguousarray(a).tobytes())
    if h.hexdigest() != g["input_sha256"]:
        pytest.skip("numpy generated different synthetic inputs than when the golden file was made")
    r = oracle.evaluate(img, chk, schema_mask=1, candidates="same_page", kmax=100, cutoff=100)
    m = oracle.metrics_from_ranks(r["pair_rank"][0], r["pair_sim"], k_values=(1, 5, 10, 20))
    assert m["num_pairs"] == g["num_pairs"]
    assert {str(k): v for k, v in m["top_k"].items()} == {k: unhex(v) for k, v in g["top_k_20"].items()}
    assert m["mrr"] == unhex(g["mrr"]) and m["avg_similarity"] == unhex(g["avg_similarity"])


def test_all_mode_is_consistent_with_same_page_mode(oracle, synthetic):
    """Ranking against every chunk restricted to the page's chunks gives the same-page order."""
    img, chk, _ = synthetic.make_numpy(40, 320, 64, T=64, seed=7)
    a = oracle.evaluate(img, chk, T=64, schema_mask=15, candidates="all", lam=(0.3, 0.2, 0.5), kmax=320, cutoff=320)
    s = oracle.evaluate(img, chk, T=64, schema_mask=15, candidates="same_page", lam=(0.3, 0.2, 0.5), kmax=16, cutoff=320)
    assert np.array_equal(a["pair_sim"], s["pair_sim"])
    for si in range(4):
        for i in range(40):
            page = set(s["topk_idx"][si, i][s["topk_idx"][si, i] >= 0].tolist())
            order = [j for j in a["topk_idx"][si, i].tolist() if j in page]
            assert order == [j for j in s["topk_idx"][si, i].tolist() if j >= 0]
