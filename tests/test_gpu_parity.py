"""Parity of the CUDA path (through the C ABI) with the CPU oracle and the golden
vectors.  Needs a B200:  python -m pytest tests -m gpu

Bar: indices, ranks and metric values bit-exact; scores bit-exact where both sides
follow the canonical summation order (tolerance 1e-5 written where it is not)."""
import json

import numpy as np
import pytest

from conftest import GOLDEN, random_texts_and_terms, unhex

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300, method="thread")]

ALL4 = ["vanilla_clip", "clip_lexical", "clip_positional", "clip_combined"]


@pytest.fixture(scope="module")
def eng(pkg):
    e = pkg.AlignmentEngine(0)
    yield e
    e.close()


@pytest.fixture(scope="module", params=[1, 2], ids=["cta_group2", "b_multicast"])
def eng_pairs(pkg, request):
    """An engine whose fused kernel runs on CTA pairs: 1 = tcgen05.mma.cta_group::2 (one M = 256 MMA per pair),
    2 = two cta_group::1 kernels in a cluster that fill one B ring together by TMA multicast."""
    e = pkg.AlignmentEngine(0)
    e.set_option("cta_pairs", request.param)
    yield e
    e.close()


def load(eng, img, chk, T=0):
    eng.set_images(img["emb"], img["key"], img.get("bbox"), img.get("terms"))
    eng.set_chunks(chk["emb"], chk["key"], chk.get("bbox"), chk.get("terms"), n_terms=T)


def as_inputs(d, how):
    """The same table as pageable numpy arrays, page-locked torch tensors or CUDA tensors."""
    if how == "pageable":
        return d
    import torch
    out = {}
    for k, v in d.items():
        if v is None:
            out[k] = None
            continue
        t = torch.from_numpy(v.view(np.int64) if v.dtype == np.uint64 else v)
        out[k] = t.pin_memory() if how == "pinned" else t.cuda()
    return out


def check_against_oracle(oracle, eng, img, chk, T, *, schemas=ALL4, candidates, lam=(0.0, 0.0), ks=(1, 5, 10, 20),
                         cutoff=100, path="auto", kprime=0, eps_scale=0.0, pipeline_rows=0, inputs="pageable",
                         pinned_outputs=False):
    load(eng, as_inputs(img, inputs), as_inputs(chk, inputs), T)
    r = eng.run(schemas, candidates=candidates, k_values=ks, mrr_cutoff=cutoff, weak_weight=lam, path=path,
                kprime=kprime, eps_scale=eps_scale, pipeline_rows=pipeline_rows, pinned_outputs=pinned_outputs)
    mask = sum({"vanilla_clip": 1, "clip_lexical": 2, "clip_positional": 4, "clip_combined": 8}[s] for s in schemas)
    o = oracle.evaluate(img, chk, T=T, schema_mask=mask, candidates=candidates, lam=(lam[0], lam[1], lam[0] + lam[1]),
                        kmax=max(ks), cutoff=max(max(ks), cutoff))
    off, pc = eng.pairs()
    assert np.array_equal(off, o["pair_offsets"]) and np.array_equal(pc, o["pair_chunk"])
    assert np.array_equal(r["topk_idx"], o["topk_idx"])
    assert np.array_equal(r["topk_score"], o["topk_score"])
    assert np.array_equal(r["pair_rank"], o["pair_rank"])
    assert np.array_equal(r["pair_sim"], o["pair_sim"])
    for si in range(len(schemas)):
        for q, k in enumerate(ks):
            assert r["hits"][si, q] == np.count_nonzero((o["pair_rank"][si] >= 1) & (o["pair_rank"][si] <= k))
        rr = sum(1.0 / x for x in o["pair_rank"][si].tolist() if 1 <= x <= cutoff)
        assert r["rr_sum"][si] == pytest.approx(rr, rel=1e-12, abs=1e-12)
    assert r["sim_sum"] == pytest.approx(float(o["pair_sim"].sum()), rel=1e-12, abs=1e-12)
    assert r["num_pairs"] == len(pc)
    return r


# ----------------------------------------------------------------------------- same-page (reference) mode
def test_small_corpus_metrics_json_is_byte_identical(pkg, small_corpus, tmp_path, capsys):
    d, corpus = small_corpus
    ev = pkg.evaluate_alignments if hasattr(pkg, "evaluate_alignments") else None
    import importlib
    ev = importlib.import_module(pkg.__name__ + ".evaluate_alignments")
    ev.clear_schemas()
    for s in ALL4:
        ev.register_schema(s, corpus)
    ev.OUTPUT_DIR = tmp_path
    ev.print_metrics_report()
    assert (tmp_path / "metrics.json").read_text() == d["expect"]["metrics_json"]
    exp = d["expect"]
    s = "vanilla_clip"
    assert [list(p) for p in ev.get_image_text_pairs(s)] == exp["pairs"]
    for iid in corpus.image_ids:
        assert ev.get_top_k_similar_chunks(iid, s, 10) == [(c, unhex(v)) for c, v in exp["top10"][iid]]
        assert [c for c, _ in ev.get_top_k_similar_chunks(iid, s, 100)] == exp["top100"][iid]
    assert {str(k): v for k, v in ev.compute_top_k_accuracy(s, [1, 5, 10, 20]).items()} == \
        {k: unhex(v) for k, v in exp["top_k_1_5_10_20"].items()}
    assert ev.compute_mrr(s) == unhex(exp["mrr"])
    assert ev.compute_average_similarity(s) == unhex(exp["avg_similarity"])
    for (a, b, _, _), v in list(zip(exp["pairs"], exp["pair_similarity"]))[:50]:
        assert ev.compute_similarity(a, b, s) == unhex(v)
    # a pair that is not on the same page goes through the single-pair path
    from oracle import oracle
    i, j = 0, len(corpus.chunk_ids) - 1
    assert ev.compute_similarity(corpus.image_ids[i], corpus.chunk_ids[j], s) == \
        oracle.cosine(corpus.img["emb"][i], corpus.chk["emb"][j])
    for sc in ALL4[1:]:
        got = ev.get_weak_supervision_scores(sc)
        # the whole `alignments` table of the unmodified reference, page-None pairs included; weak_score is REAL
        want = {k: sorted(unhex(x) for x in v) for k, v in exp["weak_scores"][sc].items()}
        assert set(got) == set(want)
        for k in got:
            assert sorted(got[k]) == want[k], (sc, k)
    ev.clear_schemas()


def test_small_corpus_alignment_records(pkg, small_corpus):
    import importlib
    ins = importlib.import_module(pkg.__name__ + ".insert_clip_embeddings")
    d, corpus = small_corpus
    for schema, (ul, up) in {"clip_lexical": (True, False), "clip_positional": (False, True),
                             "clip_combined": (True, True)}.items():
        want = [(a, b, unhex(s), t) for a, b, s, t in d["expect"]["alignments"][schema]]
        # the insert loop pairs page None with page None (:377-380): three such records in the lexical schemas
        assert not ul or any("pNone" in a for a, _, _, _ in want)
        got = ins.compute_alignment_records(corpus, ul, up)
        assert len(got) == len(want)
        assert [(g[0], g[1], g[3]) for g in got] == [(w[0], w[1], w[3]) for w in want]
        assert all(g[2] == w[2] or abs(g[2] - w[2]) <= 2.3e-16 for g, w in zip(got, want))


# ----------------------------------------------------------------------------- ingest: lexical term sets
def test_term_bitsets_golden_and_random(oracle, eng):
    """mmalign_term_bitsets against the reference's compute_lexical_alignment answers (golden), the oracle and
    Python's `in` (src/insert_clip_embeddings.py:149-150), on strings with multi-byte UTF-8, empty terms and texts,
    duplicate terms, terms longer than the text and matches that end at the last byte."""
    cases = json.loads((GOLDEN / "weak_vectors.json").read_text())["lexical_text"]
    for c in cases:
        bits = eng.term_bitsets([c["text"].lower()], c["terms"])
        assert np.array_equal(bits, oracle.term_bitsets([c["text"]], c["terms"]))
        hits = int(np.unpackbits(bits.view(np.uint8)).sum())
        assert oracle.lexical(hits, len(c["terms"])) == unhex(c["expect"]), c
    for seed, m, T, W in [(1, 300, 150, None), (2, 1000, 64, None), (3, 50, 1, 4), (4, 2000, 700, None), (5, 7, 0, None)]:
        texts, terms = random_texts_and_terms(np.random.default_rng(seed), m, max(T, 6))
        terms = terms[:T]
        bits = eng.term_bitsets([t.lower() for t in texts], terms, W)
        assert np.array_equal(bits, oracle.term_bitsets(texts, terms, W)), (seed, m, T)
        for j in range(0, m, 37):
            low = texts[j].lower()
            for t, term in enumerate(terms):
                assert bool((int(bits[j, t >> 6]) >> (t & 63)) & 1) == (term in low), (j, t)
    assert eng.term_bitsets([], ["a"]).shape == (0, 1)


def test_term_bitsets_table_scale(oracle, eng, pkg, small_corpus):
    """A 20k-chunk table of word-like text with 512 terms (device-side matching equals the oracle), and the
    golden corpus ingested through build_corpus on the GPU."""
    rng = np.random.default_rng(8)
    vocab = ["".join(rng.choice(list("abcdefghijklmnopqrstuvwxyz"), size=int(rng.integers(2, 10)))) for _ in range(3000)]
    texts = [" ".join(rng.choice(vocab, size=int(rng.integers(5, 120)))) for _ in range(20000)]
    terms = list(rng.choice(vocab, size=500)) + ["valve seat", "o-ring", " ", "e", "zzzzzz", "a b", "the", "ing", "qu", "x", "tion", "er "]
    bits = eng.term_bitsets(texts, terms)
    assert np.array_equal(bits, oracle.term_bitsets(texts, terms))
    d, c = small_corpus
    z = np.load(GOLDEN / "small_corpus.npz")
    g = pkg.build_corpus(d["images"], d["chunks"], z["img_emb"], z["chk_emb"], d["lexical_components"], engine=eng)
    assert np.array_equal(g.chk["terms"], c.chk["terms"]) and g.n_terms == c.n_terms


def test_weak_functions_known_answers(pkg):
    import importlib
    ins = importlib.import_module(pkg.__name__ + ".insert_clip_embeddings")
    g = json.loads((GOLDEN / "weak_vectors.json").read_text())
    for c in g["positional"][:120]:
        got, want = ins.compute_positional_alignment({"bbox": c["image"]}, {"bbox": c["chunk"]}), unhex(c["expect"])
        assert got == want or abs(got - want) <= 2.3e-16, c
    for c in g["lexical_text"]:
        assert ins.compute_lexical_alignment({"text": c["text"]}, c["terms"]) == unhex(c["expect"])


def test_config1_matches_reference_metrics(oracle, eng, synthetic):
    g = json.loads((GOLDEN / "config1.json").read_text())
    img, chk, _ = synthetic.make_numpy(g["N"], g["M"], g["D"], seed=g["seed"])
    r = check_against_oracle(oracle, eng, img, chk, 512, schemas=["vanilla_clip"], candidates="same_page")
    P = r["num_pairs"]
    assert P == g["num_pairs"]
    for q, k in enumerate((1, 5, 10, 20)):
        assert r["hits"][0, q] / P == unhex(g["top_k_20"][str(k)])
    m = oracle.metrics_from_ranks(r["pair_rank"][0], r["pair_sim"])
    assert m["mrr"] == unhex(g["mrr"]) and m["avg_similarity"] == unhex(g["avg_similarity"])


# ----------------------------------------------------------------------------- full N x M mode
@pytest.mark.parametrize("N,M,D", [(40, 320, 64), (130, 300, 128), (1, 1, 64), (257, 1031, 512)])
def test_exact_scan_matches_oracle(oracle, eng, synthetic, N, M, D):
    img, chk, _ = synthetic.make_numpy(N, M, D, T=64, seed=11)
    check_against_oracle(oracle, eng, img, chk, 64, candidates="all", lam=(0.3, 0.2), path="exact")


def test_fused_raw_scores_match_bf16_matmul(eng, synthetic):
    import torch
    for N, M, D in [(128, 256, 64), (200, 700, 128), (300, 1000, 512)]:
        img, chk, _ = synthetic.make_numpy(N, M, D, seed=5)
        load(eng, img, chk)
        got = eng.debug_scores()
        a = torch.from_numpy(img["emb"]).cuda().bfloat16().float()  # rows are unit norm already
        b = torch.from_numpy(chk["emb"]).cuda().bfloat16().float()
        torch.backends.cuda.matmul.allow_tf32 = False
        want = (a.double() @ b.double().T).float().cpu().numpy()
        # tolerance: the tensor cores accumulate in fp32 with truncation; measured 2.9e-5 at D=512.
        # The row certificate (rescore.cu) budgets D * 2.4e-7 = 1.2e-4 for it.
        assert np.abs(got - want).max() < D * 1.2e-7, (N, M, D, np.abs(got - want).max())


def test_cta_pair_raw_scores(eng_pairs, synthetic):
    """The CTA-pair kernel's tile scores against an fp64 product of the very bf16 operands: row blocks in pairs
    (odd counts padded by a phantom block), B halves of 128 columns per CTA, ragged last tiles, D from 64 to 1024
    (A rows resident in shared memory up to D = 512, streamed with B beyond)."""
    import torch
    for N, M, D in [(128, 256, 64), (256, 512, 64), (200, 700, 128), (300, 1000, 512), (130, 300, 768), (700, 1500, 1024),
                    (1000, 40000, 256)]:
        img, chk, _ = synthetic.make_numpy(N, M, D, seed=5)
        load(eng_pairs, img, chk)
        got = eng_pairs.debug_scores()
        a, b = eng_pairs.debug_operands()
        want = (torch.from_numpy(a).cuda().double() @ torch.from_numpy(b).cuda().double().T).cpu().numpy()
        assert np.abs(got - want).max() < D * 1.2e-7, (N, M, D, np.abs(got - want).max())


@pytest.mark.parametrize("N,M,D,ks,cutoff", [
    (300, 1000, 128, (1, 5, 10), 20), (1000, 5000, 512, (1, 5, 10, 20), 100), (130, 2000, 64, (1, 5, 10, 20), 100),
    (513, 3000, 256, (10,), 10), (256, 4096, 768, (1, 5, 10, 20), 30), (200, 3000, 1024, (1, 5, 10, 20), 100),
    (20000, 30000, 128, (1, 5, 10, 20), 100),
])
def test_cta_pair_path_matches_oracle(oracle, eng_pairs, synthetic, N, M, D, ks, cutoff):
    img, chk, _ = synthetic.make_numpy(N, M, D, T=512, seed=3)
    if N > 5000:      # many units per CTA pair: the oracle checks a sample of rows, pipelined in small slabs
        load(eng_pairs, img, chk, 512)
        r = eng_pairs.run(ALL4, candidates="all", k_values=ks, mrr_cutoff=cutoff, weak_weight=(0.3, 0.2), pipeline_rows=8192)
        rows = np.arange(0, N, 97)
        sub = {k: (v[rows] if v is not None else None) for k, v in img.items()}
        o = oracle.evaluate(sub, chk, T=512, schema_mask=15, candidates="all", lam=(0.3, 0.2, 0.5), kmax=max(ks), cutoff=cutoff)
        assert np.array_equal(r["topk_idx"][:, rows], o["topk_idx"]) and np.array_equal(r["topk_score"][:, rows], o["topk_score"])
        assert r["stats"]["slabs"] == 3 and r["stats"]["eps_violations"] == 0
        return
    r = check_against_oracle(oracle, eng_pairs, img, chk, 512, candidates="all", lam=(0.3, 0.2), ks=ks, cutoff=cutoff)
    assert r["stats"]["fused_launches"] == 1


def test_cta_pair_ties_and_groups(oracle, eng_pairs, synthetic):
    img, chk, _ = synthetic.make_numpy(256, 3000, 128, T=64, seed=9)
    chk["emb"][100:400] = chk["emb"][100]          # 300 identical chunks: a wall of equal scores
    img["emb"][7] = chk["emb"][100]
    r = check_against_oracle(oracle, eng_pairs, img, chk, 64, candidates="all", lam=(0.1, 0.1), ks=(1, 5, 10), cutoff=20,
                             kprime=20)
    assert r["stats"]["rows_rescanned"] > 0
    check_against_oracle(oracle, eng_pairs, img, chk, 64, candidates="all", lam=(0.1, 0.1), ks=(1, 5, 10), cutoff=20,
                         pipeline_rows=128, inputs="pinned", pinned_outputs=True)


@pytest.mark.parametrize("N,M,D,ks,cutoff", [
    (300, 1000, 128, (1, 5, 10), 20),
    (1000, 5000, 512, (1, 5, 10, 20), 100),
    (130, 2000, 64, (1, 5, 10, 20), 100),
    (513, 3000, 256, (10,), 10),
    (256, 4096, 768, (1, 5, 10, 20), 30),
    (200, 3000, 1024, (1, 5, 10, 20), 100),
])
def test_fused_path_matches_oracle(oracle, eng, synthetic, N, M, D, ks, cutoff):
    img, chk, _ = synthetic.make_numpy(N, M, D, T=512, seed=3)
    r = check_against_oracle(oracle, eng, img, chk, 512, candidates="all", lam=(0.3, 0.2), ks=ks, cutoff=cutoff)
    assert r["stats"]["fused_launches"] == 1


def test_fused_path_with_ties_and_forced_rescan(oracle, eng, synthetic):
    img, chk, _ = synthetic.make_numpy(256, 3000, 128, T=64, seed=9)
    chk["emb"][100:400] = chk["emb"][100]          # 300 identical chunks: a wall of equal scores
    img["emb"][7] = chk["emb"][100]
    r = check_against_oracle(oracle, eng, img, chk, 64, candidates="all", lam=(0.1, 0.1), ks=(1, 5, 10), cutoff=20,
                             kprime=20)             # K' = K: most rows cannot be certified
    assert r["stats"]["rows_rescanned"] > 0


@pytest.mark.parametrize("N,M,D,cutoff,eps_scale", [(300, 20000, 128, 100, 16.0), (2500, 3000, 64, 20, 16.0),
                                                      (64, 900, 64, 100, 100.0), (200, 30000, 256, 100, 100.0)])
def test_two_stage_exact_scan(oracle, eng, synthetic, N, M, D, cutoff, eps_scale):
    """An inflated error bound (eps_scale) leaves (almost) no row certified, so every one goes through the exact
    scan: stage 1 (whole-GPU prefilter against the row's threshold) for the first 2048 failed rows, the streaming
    scan for the rest and for rows whose survivors overflow stage 1 (eps_scale=100: the threshold excludes nothing)."""
    img, chk, _ = synthetic.make_numpy(N, M, D, T=64, seed=13)
    r = check_against_oracle(oracle, eng, img, chk, 64, candidates="all", lam=(0.3, 0.2), ks=(1, 5, 10, 20), cutoff=cutoff,
                             eps_scale=eps_scale)
    assert r["stats"]["rows_rescanned"] > (N // 2 if eps_scale >= 100 else 0)


@pytest.mark.parametrize("path,cand", [("auto", "all"), ("exact", "all"), ("auto", "same_page")])
def test_row_slab(oracle, eng, synthetic, path, cand):
    """mmalign_run restricted to image rows [row0, row0 + rows): outputs are the slab's slice of the full result."""
    img, chk, _ = synthetic.make_numpy(500, 3000, 128, T=64, seed=17)
    load(eng, img, chk, 64)
    ks, cutoff, lam = (1, 5, 10), 30, (0.3, 0.2)
    o = oracle.evaluate(img, chk, T=64, schema_mask=15, candidates=cand, lam=(lam[0], lam[1], lam[0] + lam[1]), kmax=10,
                        cutoff=cutoff)
    for r0, rows in [(0, 200), (200, 300), (128, 129), (499, 1), (500, 0)]:
        r = eng.run(ALL4, candidates=cand, k_values=ks, mrr_cutoff=cutoff, weak_weight=lam, path=path, slab=(r0, rows))
        p0, p1 = o["pair_offsets"][r0], o["pair_offsets"][r0 + rows]
        assert r["num_pairs"] == p1 - p0
        assert np.array_equal(r["topk_idx"], o["topk_idx"][:, r0:r0 + rows])
        assert np.array_equal(r["topk_score"], o["topk_score"][:, r0:r0 + rows])
        assert np.array_equal(r["pair_rank"], o["pair_rank"][:, p0:p1])
        assert np.array_equal(r["pair_sim"], o["pair_sim"][p0:p1])


# ----------------------------------------------------------------------------- the slab pipeline of mmalign_run
@pytest.mark.parametrize("inputs,pinned_out", [("pageable", False), ("pinned", True), ("device", False)])
@pytest.mark.parametrize("cand,path,eps_scale", [("all", "auto", 0.0), ("all", "auto", 100.0), ("all", "exact", 0.0),
                                                 ("same_page", "auto", 0.0)])
def test_slab_pipeline_matches_oracle(oracle, eng, synthetic, inputs, pinned_out, cand, path, eps_scale):
    """mmalign_run cut into pipeline slabs (uploads of later slabs and downloads of earlier ones beside the kernels):
    the same bytes as the oracle, with a ragged last slab, from pageable, page-locked and device inputs."""
    N = 700
    img, chk, _ = synthetic.make_numpy(N, 3000, 128, T=64, seed=31)
    for pr, want_slabs in [(256, 3), (128, 6), (1000, 1), (-1, 1)]:
        r = check_against_oracle(oracle, eng, img, chk, 64, candidates=cand, lam=(0.3, 0.2), ks=(1, 5, 10), cutoff=30,
                                 path=path, eps_scale=eps_scale, pipeline_rows=pr, inputs=inputs, pinned_outputs=pinned_out)
        assert r["stats"]["slabs"] == want_slabs
        if cand == "all" and path == "auto":
            assert r["stats"]["fused_launches"] == want_slabs
            if eps_scale > 1:
                assert r["stats"]["rows_rescanned"] > 0


@pytest.mark.parametrize("order", ["images_first", "chunks_first"])
@pytest.mark.parametrize("eps_scale", [0.0, 100.0])
def test_first_slab_by_column_groups(oracle, pkg, synthetic, order, eps_scale):
    """Page-locked tables uploaded in small pieces (mmalign_set_option piece_bytes): the chunk table arrives in four
    column groups and the first slab of the pipelined run is contracted group by group -- a row of that slab has four
    times the lists of the other slabs' rows.  Same bytes as the oracle, whichever table is set first."""
    img, chk, _ = synthetic.make_numpy(700, 6000, 128, T=64, seed=53)
    e = pkg.AlignmentEngine(0)
    try:
        e.set_option("piece_bytes", 128 * 1024)      # 256-row pieces: 3 image pieces, 24 chunk pieces
        pi, pc = as_inputs(img, "pinned"), as_inputs(chk, "pinned")
        for rep in range(2):
            if order == "images_first":
                e.set_images(pi["emb"], pi["key"], pi["bbox"], None)
                e.set_chunks(pc["emb"], pc["key"], pc["bbox"], pc["terms"], n_terms=64)
            else:
                e.set_chunks(pc["emb"], pc["key"], pc["bbox"], pc["terms"], n_terms=64)
                e.set_images(pi["emb"], pi["key"], pi["bbox"], None)
            r = e.run(ALL4, candidates="all", k_values=(1, 5, 10), mrr_cutoff=30, weak_weight=(0.3, 0.2), pipeline_rows=256,
                      pinned_outputs=True, eps_scale=eps_scale)
            o = oracle.evaluate(img, chk, T=64, schema_mask=15, candidates="all", lam=(0.3, 0.2, 0.5), kmax=10, cutoff=30)
            assert r["stats"]["slabs"] == 3 and r["stats"]["fused_launches"] == 4 + 2
            assert np.array_equal(r["topk_idx"], o["topk_idx"]) and np.array_equal(r["topk_score"], o["topk_score"])
            assert np.array_equal(r["pair_rank"], o["pair_rank"]) and np.array_equal(r["pair_sim"], o["pair_sim"])
            # a second run on the same tables has nothing in flight: one launch per slab
            r2 = e.run(ALL4, candidates="all", k_values=(1, 5, 10), mrr_cutoff=30, weak_weight=(0.3, 0.2), pipeline_rows=256)
            assert r2["stats"]["fused_launches"] == 3 and np.array_equal(r2["topk_idx"], o["topk_idx"])
    finally:
        e.close()


@pytest.mark.parametrize("eps_scale", [0.0, 100.0])
@pytest.mark.parametrize("inputs", ["device", "pinned"])
def test_rescoring_beside_contraction(oracle, pkg, synthetic, eps_scale, inputs):
    """mmalign_set_option k2_sms: the exact rescoring of slab s on a few SMs while slab s+1 is contracted (two
    streams, alternating list buffers).  Same bytes as the oracle, also when every row goes through the exact scan."""
    img, chk, _ = synthetic.make_numpy(900, 4000, 128, T=64, seed=59)
    e = pkg.AlignmentEngine(0)
    try:
        e.set_option("k2_sms", 8)
        for pr, slabs in [(128, 8), (256, 4), (512, 2), (1024, 1)]:
            r = check_against_oracle(oracle, e, img, chk, 64, candidates="all", lam=(0.3, 0.2), ks=(1, 5, 10), cutoff=30,
                                     eps_scale=eps_scale, pipeline_rows=pr, inputs=inputs, pinned_outputs=inputs == "pinned")
            assert r["stats"]["slabs"] == slabs and r["stats"]["k2_sms"] == (8 if slabs > 1 else 0)
        # the automatic slab size (four waves of the contraction's SMs) at a size that has several of them
        big_i, big_c, _ = synthetic.make_numpy(80000, 2000, 64, T=64, seed=61)
        load(e, big_i, big_c, 64)
        r = e.run(ALL4, candidates="all", k_values=(1, 5, 10), mrr_cutoff=30, weak_weight=(0.3, 0.2))
        o = oracle.evaluate(big_i, big_c, T=64, schema_mask=15, candidates="all", lam=(0.3, 0.2, 0.5), kmax=10, cutoff=30)
        assert r["stats"]["slabs"] == 3 and r["stats"]["k2_sms"] == 8   # 2 x 2 waves of 140 row blocks + the rest
        assert np.array_equal(r["topk_idx"], o["topk_idx"]) and np.array_equal(r["pair_rank"], o["pair_rank"])
    finally:
        e.close()


def test_slab_pipeline_with_row_range_and_empty_tables(oracle, eng, synthetic):
    img, chk, _ = synthetic.make_numpy(500, 3000, 128, T=64, seed=17)
    load(eng, img, chk, 64)
    ks, cutoff, lam = (1, 5, 10), 30, (0.3, 0.2)
    o = oracle.evaluate(img, chk, T=64, schema_mask=15, candidates="all", lam=(lam[0], lam[1], lam[0] + lam[1]), kmax=10,
                        cutoff=cutoff)
    for r0, rows in [(100, 390), (499, 1), (500, 0)]:
        r = eng.run(ALL4, candidates="all", k_values=ks, mrr_cutoff=cutoff, weak_weight=lam, slab=(r0, rows), pipeline_rows=128)
        p0, p1 = o["pair_offsets"][r0], o["pair_offsets"][r0 + rows]
        assert r["stats"]["slabs"] == -(-rows // 128)
        assert np.array_equal(r["topk_idx"], o["topk_idx"][:, r0:r0 + rows])
        assert np.array_equal(r["topk_score"], o["topk_score"][:, r0:r0 + rows])
        assert np.array_equal(r["pair_rank"], o["pair_rank"][:, p0:p1]) and np.array_equal(r["pair_sim"], o["pair_sim"][p0:p1])
    empty = dict(emb=np.zeros((0, 128), np.float32), key=np.zeros(0, np.uint64), bbox=np.zeros((0, 4)), terms=None)
    load(eng, empty, chk, 64)
    r = eng.run(ALL4, candidates="all", pipeline_rows=128)
    assert r["num_pairs"] == 0 and r["topk_idx"].shape == (4, 0, 10) and r["stats"]["slabs"] == 0


def test_set_calls_are_stream_ordered(oracle, eng, synthetic):
    """set_* returns before its uploads have run; a second set_* of the same table, a run, and mmalign_sync all
    see the right data (page-locked inputs are borrowed until then)."""
    a_img, a_chk, _ = synthetic.make_numpy(300, 2000, 128, T=64, seed=1)
    b_img, b_chk, _ = synthetic.make_numpy(300, 2000, 128, T=64, seed=2)
    pa, pb = (as_inputs(a_img, "pinned"), as_inputs(a_chk, "pinned")), (as_inputs(b_img, "pinned"), as_inputs(b_chk, "pinned"))
    load(eng, *pa, 64)
    load(eng, *pb, 64)   # overwrites the upload buffers of the first call while its preparation may still run
    eng.sync()
    r = eng.run(ALL4, candidates="all", k_values=(1, 5, 10), mrr_cutoff=30, weak_weight=(0.3, 0.2))
    o = oracle.evaluate(b_img, b_chk, T=64, schema_mask=15, candidates="all", lam=(0.3, 0.2, 0.5), kmax=10, cutoff=30)
    assert np.array_equal(r["topk_idx"], o["topk_idx"]) and np.array_equal(r["pair_rank"], o["pair_rank"])
    load(eng, *pa, 64)
    r = eng.run(ALL4, candidates="all", k_values=(1, 5, 10), mrr_cutoff=30, weak_weight=(0.3, 0.2))
    o = oracle.evaluate(a_img, a_chk, T=64, schema_mask=15, candidates="all", lam=(0.3, 0.2, 0.5), kmax=10, cutoff=30)
    assert np.array_equal(r["topk_idx"], o["topk_idx"]) and np.array_equal(r["topk_score"], o["topk_score"])


# ----------------------------------------------------------------------------- image term sets (AND-popcount)
@pytest.mark.parametrize("cand,path", [("all", "auto"), ("all", "exact"), ("same_page", "auto")])
def test_image_term_sets(oracle, eng, synthetic, cand, path):
    """Images with their own term sets: hits = popcount(chunk & image) (BASELINE.json north_star), both weights non-zero."""
    img, chk, _ = synthetic.make_numpy(300, 3000, 128, T=100, p_term=0.05, seed=37, img_p_term=0.5)
    assert img["terms"] is not None
    r = check_against_oracle(oracle, eng, img, chk, 100, candidates=cand, lam=(0.3, 0.2), ks=(1, 5, 10, 20), cutoff=100,
                             path=path)
    plain = dict(img, terms=None)
    load(eng, plain, chk, 100)
    q = eng.run(ALL4, candidates=cand, k_values=(1, 5, 10, 20), mrr_cutoff=100, weak_weight=(0.3, 0.2), path=path)
    assert not np.array_equal(q["topk_score"][1], r["topk_score"][1])      # the image sets do change the lexical schema
    assert np.array_equal(q["topk_score"][0], r["topk_score"][0])          # ... and leave vanilla_clip alone
    # the alignments producer with image term sets
    load(eng, img, chk, 100)
    for si, schema in enumerate(ALL4):
        off, pc, rec = oracle.alignments(img, chk, T=100, schema=si)
        assert np.array_equal(eng.alignments(schema), rec)
        assert si == 0 or np.count_nonzero(rec) > 0


# ----------------------------------------------------------------------------- half-precision encoder rows
@pytest.mark.parametrize("dtype", ["float16", "bfloat16"])
@pytest.mark.parametrize("where", ["device", "host"])
def test_half_precision_rows(oracle, eng, synthetic, dtype, where):
    """fp16 / bf16 rows (an encoder batch) through mmalign_set_*_half: exactly the result of the same values as fp32."""
    import torch
    img, chk, _ = synthetic.make_numpy(300, 3000, 128, T=64, seed=67)
    td = getattr(torch, dtype)
    hi, hc = torch.from_numpy(img["emb"]).to(td), torch.from_numpy(chk["emb"]).to(td)
    img32, chk32 = dict(img, emb=hi.float().numpy()), dict(chk, emb=hc.float().numpy())   # the values the halves hold
    if where == "device":
        hi, hc = hi.cuda(), hc.cuda()
    elif dtype == "float16":
        hi, hc = hi.numpy(), hc.numpy()      # numpy float16 arrays on the host
    eng.set_images(hi, img["key"], img["bbox"], None)
    eng.set_chunks(hc, chk["key"], chk["bbox"], chk["terms"], n_terms=64)
    for cand in ("all", "same_page"):
        r = eng.run(ALL4, candidates=cand, k_values=(1, 5, 10), mrr_cutoff=30, weak_weight=(0.3, 0.2))
        o = oracle.evaluate(img32, chk32, T=64, schema_mask=15, candidates=cand, lam=(0.3, 0.2, 0.5), kmax=10, cutoff=30)
        assert np.array_equal(r["topk_idx"], o["topk_idx"]) and np.array_equal(r["topk_score"], o["topk_score"])
        assert np.array_equal(r["pair_rank"], o["pair_rank"]) and np.array_equal(r["pair_sim"], o["pair_sim"])


# ----------------------------------------------------------------------------- deep lists, error model
def test_deep_lists_are_exact_to_the_full_depth(oracle, eng, synthetic):
    """deep_idx / deep_score on the fused path (include/mmalign.h): final to max(Kmax, mrr_cutoff) entries."""
    img, chk, _ = synthetic.make_numpy(300, 5000, 128, T=64, seed=41)
    load(eng, img, chk, 64)
    r = eng.run(ALL4, candidates="all", k_values=(1, 5, 10), mrr_cutoff=100, weak_weight=(0.3, 0.2), deep=True)
    o = oracle.evaluate(img, chk, T=64, schema_mask=15, candidates="all", lam=(0.3, 0.2, 0.5), kmax=100, cutoff=100)
    assert r["deep_idx"].shape == (4, 300, 100)
    assert np.array_equal(r["deep_idx"], o["topk_idx"]) and np.array_equal(r["deep_score"], o["topk_score"])
    assert np.array_equal(r["topk_idx"], o["topk_idx"][:, :, :10]) and np.array_equal(r["pair_rank"], o["pair_rank"])


def _adversarial(kind, N, M, D, rng):
    if kind == "same_sign":      # every product positive: the partial sums grow monotonically to ~0.64
        a, b = np.abs(rng.standard_normal((N, D))), np.abs(rng.standard_normal((M, D)))
    elif kind == "cancel":       # large alternating components: sum |a_k b_k| ~ 1 while the dot product is ~ 0
        base = rng.standard_normal(D)
        sign = np.where(np.arange(D) % 2 == 0, 1.0, -1.0)
        a = base * sign + 0.05 * rng.standard_normal((N, D))
        b = base + 0.05 * rng.standard_normal((M, D))
    elif kind == "spiky":        # a few components carry the norm (large exponent spread inside one MMA)
        a, b = rng.standard_normal((N, D)) * 1e-3, rng.standard_normal((M, D)) * 1e-3
        for x in (a, b):
            x[np.arange(len(x)), rng.integers(0, D, len(x))] = 1.0
            x[np.arange(len(x)), rng.integers(0, D, len(x))] = -0.7
    else:                        # aligned: all rows close to one direction, scores near 1
        base = rng.standard_normal(D)
        a, b = base + 0.1 * rng.standard_normal((N, D)), base + 0.1 * rng.standard_normal((M, D))
    unit = lambda x: (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)
    return unit(a), unit(b)


@pytest.mark.parametrize("kind", ["same_sign", "cancel", "spiky", "aligned"])
@pytest.mark.parametrize("D", [512, 1024])
def test_error_model_on_adversarial_inputs(oracle, eng, synthetic, kind, D):
    """The certificate budgets D * 2.4e-7 for the tensor cores' fp32 accumulation.  Inputs
    chosen against that model -- same-sign products, heavy cancellation, a wide exponent spread, scores near 1 --
    must (a) stay within half the budget on the raw tile scores, (b) never trip the run-time check of the bound
    (stats eps_violations), (c) give the oracle's result."""
    import torch
    N, M = 256, 3000
    img, chk, _ = synthetic.make_numpy(N, M, D, T=64, seed=47)
    img["emb"], chk["emb"] = _adversarial(kind, N, M, D, np.random.default_rng(D + len(kind)))
    load(eng, img, chk, 64)
    got = eng.debug_scores()
    a, b = eng.debug_operands()   # the bf16 rows the tensor cores multiplied, widened to fp32 (exact)
    want = (torch.from_numpy(a).cuda().double() @ torch.from_numpy(b).cuda().double().T).cpu().numpy()
    assert np.abs(got - want).max() < D * 1.2e-7, (kind, D, np.abs(got - want).max())
    # K0's operands are the nearest bf16 of the fp32-normalised rows (up to the rounding of the normalisation itself)
    ref = img["emb"] / np.linalg.norm(img["emb"].astype(np.float64), axis=1, keepdims=True)
    assert np.abs(a - ref).max() <= 2.0 ** -8 * np.abs(ref).max() * 1.01   # half an ulp of an 8-bit significand
    r = check_against_oracle(oracle, eng, img, chk, 64, candidates="all", lam=(0.3, 0.2), ks=(1, 5, 10, 20), cutoff=100)
    assert r["stats"]["eps_violations"] == 0


def test_edge_shapes(oracle, eng, synthetic, pkg):
    img, chk, _ = synthetic.make_numpy(5, 64, 64, T=64, seed=1)
    empty = dict(emb=np.zeros((0, 64), np.float32), key=np.zeros(0, np.uint64), bbox=np.zeros((0, 4)), terms=None)
    load(eng, empty, chk, 64)
    r = eng.run(ALL4, candidates="all")
    assert r["num_pairs"] == 0 and r["topk_idx"].shape == (4, 0, 10)
    echk = dict(empty, terms=np.zeros((0, 1), np.uint64))
    load(eng, img, echk, 64)
    r = eng.run(ALL4, candidates="all")
    assert r["num_pairs"] == 0 and (r["topk_idx"] == -1).all()
    # NULL pages never join
    img["key"][:] = np.uint64(0xFFFFFFFFFFFFFFFF)
    check_against_oracle(oracle, eng, img, chk, 64, candidates="all")
    check_against_oracle(oracle, eng, img, chk, 64, candidates="same_page")
    with pytest.raises(pkg.MMAlignError):
        eng.run(["clip_lexical"], k_values=[0])


def test_large_run_properties(eng, synthetic):
    """100k x 200k x 512: size-independent checks (no oracle at this size)."""
    import torch
    N, M, D = 100_000, 200_000, 512
    img, chk, meta = synthetic.make_torch(N, M, D, device="cuda")
    load(eng, img, chk, 512)
    r = eng.run(ALL4, candidates="all", k_values=(1, 5, 10, 20), mrr_cutoff=100, weak_weight=(0.3, 0.2))
    assert r["num_pairs"] == 8 * N
    idx, sc = r["topk_idx"], r["topk_score"]
    assert (np.diff(sc, axis=2) <= 0).all()                                   # sorted
    ties = np.diff(sc, axis=2) == 0
    assert (np.diff(idx, axis=2)[ties] > 0).all()                              # lower index first
    assert (idx >= 0).all() and (idx < M).all()
    planted = meta["planted"].cpu().numpy()
    assert (idx[0, :, 0] == planted).mean() > 0.5                               # the planted chunk usually wins
    # vanilla top-1 score equals an fp64 recomputation of that pair's cosine
    rows = np.arange(0, N, 997)
    a = img["emb"][rows].double()
    b = chk["emb"][torch.from_numpy(idx[0, rows, 0]).cuda()].double()
    cos = ((a * b).sum(1) / (a.norm(dim=1) * b.norm(dim=1))).cpu().numpy()
    assert np.abs(cos - sc[0, rows, 0]).max() < 1e-5
    # hits are monotone in k, and consistent with the per-pair ranks
    assert (np.diff(r["hits"], axis=1) >= 0).all()
    for si in range(4):
        assert r["hits"][si, 3] == np.count_nonzero((r["pair_rank"][si] >= 1) & (r["pair_rank"][si] <= 20))
    # the exact scan of a sample of rows agrees with the fused path
    sub = dict(emb=img["emb"][rows], key=img["key"][rows], bbox=img["bbox"][rows], terms=None)
    eng.set_images(sub["emb"], sub["key"], sub["bbox"], None)
    e = eng.run(ALL4, candidates="all", k_values=(1, 5, 10, 20), mrr_cutoff=100, weak_weight=(0.3, 0.2), path="exact")
    assert np.array_equal(e["topk_idx"], idx[:, rows]) and np.array_equal(e["topk_score"], sc[:, rows])


# ----------------------------------------------------------------------------- documented limits
def _page_heavy_corpus(synthetic, n_on_page):
    """3 images, 700 chunks: n_on_page chunks share page (7, 3) with image 0; image 1 has a NULL page."""
    img, chk, _ = synthetic.make_numpy(3, 700, 64, T=64, seed=21)
    key = lambda manual, page: np.uint64((manual << 32) | page)
    chk["key"][:] = [key(1000 + j, 1) for j in range(700)]          # every chunk alone on its page ...
    chk["key"][50:50 + n_on_page] = key(7, 3)                       # ... except one crowded page
    img["key"][:] = [key(7, 3), np.uint64(0xFFFFFFFFFFFFFFFF), key(1000 + 5, 1)]
    return img, chk


@pytest.mark.parametrize("cand", ["same_page", "all"])
def test_page_with_512_chunks_is_the_limit(oracle, eng, synthetic, pkg, cand):
    img, chk = _page_heavy_corpus(synthetic, 512)
    r = check_against_oracle(oracle, eng, img, chk, 64, candidates=cand, lam=(0.3, 0.2), ks=(1, 5, 10, 20), cutoff=100)
    assert r["num_pairs"] == 513
    img, chk = _page_heavy_corpus(synthetic, 513)
    load(eng, img, chk, 64)
    with pytest.raises(pkg.MMAlignError) as e:
        eng.run(ALL4, candidates=cand, k_values=(1, 5, 10), mrr_cutoff=100)
    assert e.value.code == -5 and "512 same-page chunks" in str(e.value)


@pytest.mark.parametrize("pipeline_rows,kprime", [(0, 0), (256, 0), (256, 300)])
def test_rows_of_both_rescoring_kernels_in_one_launch(oracle, eng, synthetic, pipeline_rows, kprime):
    """The warp-per-row kernels (select / gather / rank) rank rows with at most 32 same-page chunks and a union of at
    most 256 candidates; the others are handed to the block-per-row kernel through a row list.  Here both kinds sit
    in the same launch: every third image is on a page of 33-60 chunks, the rest on pages of one to eight, some have
    no page at all; K' = 300 additionally pushes unions past 256 entries; several pipeline slabs offset the row lists."""
    N, M, D = 700, 6000, 128
    img, chk, _ = synthetic.make_numpy(N, M, D, T=64, seed=33)
    key = lambda manual, page: np.uint64((manual << 32) | page)
    rng = np.random.default_rng(5)
    pos, page = 0, 0
    while pos < M:                                                   # chunk pages of 1..8 rows, every fifth one of 33..60
        page += 1
        n = int(rng.integers(33, 61)) if page % 5 == 0 else int(rng.integers(1, 9))
        chk["key"][pos:pos + n] = key(page // 50, page)
        pos += n
    crowded = [p for p in range(1, page + 1) if p % 5 == 0]
    light = [p for p in range(1, page + 1) if p % 5 != 0]
    for i in range(N):
        p = crowded[i % len(crowded)] if i % 3 == 0 else light[(7 * i) % len(light)]
        img["key"][i] = np.uint64(0xFFFFFFFFFFFFFFFF) if i % 41 == 0 else key(p // 50, p)
    r = check_against_oracle(oracle, eng, img, chk, 64, candidates="all", lam=(0.3, 0.2), ks=(1, 5, 10, 20), cutoff=100,
                             kprime=kprime, pipeline_rows=pipeline_rows)
    assert r["stats"]["fused_launches"] >= 1
    off, _ = eng.pairs()
    c = np.diff(off)
    assert (c > 32).sum() > 100 and ((c > 0) & (c <= 32)).sum() > 100 and (c == 0).sum() > 5


@pytest.mark.parametrize("N,M,D", [(1900, 40000, 128), (2500, 60000, 64)])
def test_unions_wider_than_the_warp_kernels_are_cut_back(oracle, eng, synthetic, N, M, D):
    """Few row blocks against many columns: the fused kernel splits the columns ~10 ways to fill the GPU, a row then has
    ~20 lists and their union above the completeness threshold outgrows the 256 slots of the warp-per-row rescoring,
    which cuts it back to its best 208 entries and raises the threshold (select_kernel).  Exact all the same."""
    img, chk, _ = synthetic.make_numpy(N, M, D, T=64, seed=61)
    r = check_against_oracle(oracle, eng, img, chk, 64, candidates="all", lam=(0.3, 0.2), ks=(1, 5, 10, 20), cutoff=100)
    assert r["stats"]["fused_launches"] >= 1


def test_widest_lists_k256_and_cutoff256(oracle, eng, synthetic, pkg):
    """Kmax = mrr_cutoff = 256 (the documented maximum): K' = 491."""
    img, chk, _ = synthetic.make_numpy(300, 6000, 64, T=64, seed=23)
    r = check_against_oracle(oracle, eng, img, chk, 64, candidates="all", lam=(0.3, 0.2), ks=(1, 256), cutoff=256)
    assert r["stats"]["kprime"] == 491 and r["topk_idx"].shape == (4, 300, 256)
    load(eng, img, chk, 64)
    for bad in (dict(k_values=(257,)), dict(mrr_cutoff=257)):
        with pytest.raises(pkg.MMAlignError):
            eng.run(ALL4, candidates="all", **bad)


def test_512_entry_lists(oracle, eng, synthetic):
    """K' = 2000 makes every list keep > 113 entries, which selects the 512-entry list variant of the fused kernel;
    the union of a row's lists then exceeds what the rescoring kernel holds per row, so those rows take the exact
    scan -- the result is still the oracle's."""
    img, chk, _ = synthetic.make_numpy(400, 20000, 64, T=64, seed=25)
    r = check_against_oracle(oracle, eng, img, chk, 64, candidates="all", lam=(0.3, 0.2), ks=(1, 5, 10, 20), cutoff=100,
                             kprime=2000)
    assert r["stats"]["kprime"] == 2000
