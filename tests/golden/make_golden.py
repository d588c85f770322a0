"""Generates the golden vectors under tests/golden/ by running the UNMODIFIED
reference code (/root/reference/src/*.py) in this container.

    python tests/golden/make_golden.py

Third-party modules the reference imports but this image lacks are stubbed and
PostgreSQL/pgvector is replaced by oracle.reference_harness.FakeDB (whose only
own arithmetic is the pgvector cosine restatement).  /root/reference does not
exist on the GPU box, so the outputs are committed:

  weak_vectors.json    compute_lexical_alignment / compute_positional_alignment
                       known answers (SURVEY.md Appendix A + 600 random cases)
  small_corpus.json/.npz  a 48-image x 256-chunk corpus with string ids, NULL
                       pages, tied embeddings, broken bboxes: pairs, per-image
                       top-10/top-100, metrics.json and the `alignments` records
                       of all four schemas
  config1.json         BASELINE config 1 (1k x 5k x 512, seeded): metrics of the
                       reference functions, plus an input checksum
"""
from __future__ import annotations

import hashlib
import io
import json
import sys
import tempfile
from contextlib import redirect_stdout
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import importlib  # noqa: E402

from oracle import reference_harness as rh  # noqa: E402

PKG = "multimodal-alignment-of-noisy-image-text-pairs-using-weak-supervision_b200"
synthetic = importlib.import_module(PKG + ".synthetic")
OUT = Path(__file__).resolve().parent
SCHEMAS = ["vanilla_clip", "clip_lexical", "clip_positional", "clip_combined"]


def fl(x):
    """JSON-safe exact float: hex string."""
    return float(x).hex()


# --------------------------------------------------------------------------- weak vectors
def weak_vectors(ins):
    rng = np.random.default_rng(1234)
    pos_cases = [
        ([0, 0, 10, 10], [5, 5, 15, 15]), ([1, 2, 3, 4], [1, 2, 3, 4]),
        ([0, 0, 100, 100], [25, 25, 75, 75]), ([0, 0, 10, 10], [10, 0, 20, 10]),
        ([0, 0, 10, 10], [100, 100, 110, 110]), ([0, 0, 10, 10], [2000, 0, 2010, 10]),
        ([5, 0, 5, 10], [0, 0, 10, 10]), ([0, 0, 10, 10], [0, 3, 10, 3]),
        (None, [0, 0, 10, 10]), ([0, 0, 10, 10], [0, 0, 0, 0]), ([0, 0, 10], [0, 0, 10, 10]),
        ([72.0, 120.5, 300.25, 410.75], [72.0, 430.0, 523.3, 442.1]), ([10, 0, 0, 10], [2, 2, 8, 8]),
        ([], [0, 0, 1, 1]), ([0, 0, 10, 10], None),
    ]
    for _ in range(600):
        def box():
            x0, y0 = rng.uniform(0, 600), rng.uniform(0, 780)
            w, h = rng.uniform(-5, 250), rng.uniform(-5, 200)
            if rng.random() < 0.1:
                w = 0.0
            if rng.random() < 0.1:
                h = 0.0
            b = [x0, y0, x0 + w, y0 + h]
            if rng.random() < 0.3:  # fp32-representable, like PyMuPDF rectangles
                b = [float(np.float32(v)) for v in b]
            if rng.random() < 0.15:  # integer grid: exact ties / touching edges
                b = [float(round(v / 25) * 25) for v in b]
            return b
        pos_cases.append((box(), box()))
    pos = [dict(image=a, chunk=b, expect=fl(ins.compute_positional_alignment({"bbox": a}, {"bbox": b})))
           for a, b in pos_cases]
    lex = []
    for T in [0, 1, 3, 9, 10, 11, 12, 19, 20, 21, 50, 64, 200, 512, 513, 1000]:
        for hits in sorted({0, 1, 2, 3, 5, 7, T // 20, T // 10, T // 10 + 1, T}):
            if hits > T:
                continue
            terms = [f"term{t:04d}x" for t in range(T)]
            text = " ".join(t.upper() if i % 2 else t for i, t in enumerate(terms[:hits]))
            lex.append(dict(T=T, hits=hits,
                            expect=fl(ins.compute_lexical_alignment({"text": text}, terms))))
    # substring semantics (Appendix A L2, L8)
    lex_text = [
        dict(text="Remove the Filter and pump", terms=["filter", "pump", "valve"]),
        dict(text="the pumps are filtered", terms=["pump", "filter"] + [f"zz{i}" for i in range(10)]),
        dict(text="", terms=["a"]),
        dict(text="ABC", terms=["abc", "bc", "c", "d", "B"]),
    ]
    for c in lex_text:
        c["expect"] = fl(ins.compute_lexical_alignment({"text": c["text"]}, c["terms"]))
    return dict(positional=pos, lexical=lex, lexical_text=lex_text)


# --------------------------------------------------------------------------- small corpus
def small_corpus():
    rng = np.random.default_rng(20261018)
    D = 64
    vocab = ["pump", "filter", "valve", "hose", "motor", "bolt", "gasket", "seal", "rotor",
             "nozzle", "lever", "spring", "washer", "bracket", "sensor"]
    filler = ["the", "remove", "install", "check", "and", "then", "carefully", "unit", "figure"]
    lexical_components = {"components": [{"term": t, "count": 3} for t in vocab[:12]]}
    images, chunks = [], []
    img_emb, chk_emb = [], []
    pages = [("manA", p) for p in range(1, 9)] + [("manB", p) for p in range(1, 6)]

    def bbox(kind=None):
        k = kind if kind is not None else rng.integers(0, 12)
        x0, y0 = float(np.float32(rng.uniform(36, 400))), float(np.float32(rng.uniform(36, 600)))
        w, h = float(np.float32(rng.uniform(20, 176))), float(np.float32(rng.uniform(8, 156)))
        if k == 0:
            return None
        if k == 1:
            return [0, 0, 0, 0]
        if k == 2:
            return [x0, y0, x0 + w]
        if k == 3:
            return [x0, y0, x0, y0 + h]
        return [x0, y0, x0 + w, y0 + h]

    cidx = 0
    for man, p in pages:
        n_c = int(rng.integers(0, 13)) if (man, p) != ("manA", 1) else 12
        base = rng.standard_normal(D).astype(np.float32)
        for c in range(n_c):
            e = rng.standard_normal(D).astype(np.float32)
            if c in (3, 7) and n_c > 7:  # exact duplicates on a page -> tied similarities
                e = base.copy()
            e /= np.linalg.norm(e)
            words = list(rng.choice(filler, 6)) + list(rng.choice(vocab, int(rng.integers(0, 5))))
            rng.shuffle(words)
            text = " ".join(w.capitalize() if rng.random() < 0.3 else w for w in words)
            chunks.append({"chunk_id": f"{man}_p{p}_c{c}", "manual_id": man, "page": p,
                           "bbox": bbox(), "text": text})
            chk_emb.append(e)
            cidx += 1
    # chunks without a page, and on a page with no image
    for c in range(3):
        e = rng.standard_normal(D).astype(np.float32)
        chunks.append({"chunk_id": f"manA_pNone_c{c}", "manual_id": "manA", "page": None,
                       "bbox": bbox(5), "text": "pump valve"})
        chk_emb.append(e / np.linalg.norm(e))
    while len(chunks) < 256:
        e = rng.standard_normal(D).astype(np.float32)
        man, p = pages[int(rng.integers(0, len(pages)))]
        c = len(chunks)
        chunks.append({"chunk_id": f"{man}_p{p}_x{c}", "manual_id": man, "page": p,
                       "bbox": bbox(), "text": " ".join(rng.choice(vocab + filler, 8))})
        chk_emb.append(e / np.linalg.norm(e))
    chk_emb = np.stack(chk_emb).astype(np.float32)
    for i in range(48):
        man, p = pages[int(rng.integers(0, len(pages)))]
        if i == 5:
            man, p = "manA", None
        if i == 6:
            man, p = "manC", 1  # manual without chunks
        same = [j for j, c in enumerate(chunks) if c["manual_id"] == man and c["page"] == p and p is not None]
        e = rng.standard_normal(D).astype(np.float32)
        if same and rng.random() < 0.7:
            e = chk_emb[same[int(rng.integers(0, len(same)))]] + 1.5 * e / np.linalg.norm(e)
        if i == 9:
            e *= 3.7  # un-normalised row (insert_clip_embeddings.py:294-297 fallback)
        else:
            e /= np.linalg.norm(e)
        images.append({"image_id": f"{man}_p{p}_img{i}", "manual_id": man, "page": p,
                       "bbox": bbox(), "bbox_source": "x", "caption": None,
                       "filename": f"img{i}.png", "image_type": "raster_image"})
        img_emb.append(e.astype(np.float32))
    img_emb = np.stack(img_emb).astype(np.float32)

    tables = {}
    for s in SCHEMAS:
        tables[s] = dict(image_ids=[i["image_id"] for i in images],
                         image_manual=[i["manual_id"] for i in images],
                         image_page=[i["page"] for i in images], image_emb=img_emb,
                         chunk_ids=[c["chunk_id"] for c in chunks],
                         chunk_manual=[c["manual_id"] for c in chunks],
                         chunk_page=[c["page"] for c in chunks], chunk_emb=chk_emb,
                         alignments=[])
    db = rh.FakeDB(tables)
    tmp = Path(tempfile.mkdtemp())
    ev, ins = rh.load_reference(db, output_dir=tmp)

    # alignments: the reference's own insert_embeddings(), unmodified, with its inputs on disk
    (tmp / "image_metadata.json").write_text(json.dumps(images))
    (tmp / "text_chunks.json").write_text(json.dumps(chunks))
    (tmp / "filtered_lexical_components.json").write_text(json.dumps(lexical_components))
    ins.IMAGE_METADATA_FILE = tmp / "image_metadata.json"
    ins.TEXT_CHUNKS_FILE = tmp / "text_chunks.json"
    ins.LEXICAL_COMPONENTS_FILE = tmp / "filtered_lexical_components.json"
    ins.IMAGES_DIR = tmp
    captured = {}

    def fake_execute_values(cur, sql, records):
        if "alignments" in sql:
            captured["rows"] = list(records)
    ins.execute_values = fake_execute_values
    ins.psycopg2 = type("pg", (), {"connect": staticmethod(db.connect)})
    flags = {"vanilla_clip": (False, False), "clip_lexical": (True, False),
             "clip_positional": (False, True), "clip_combined": (True, True)}
    align = {}
    with redirect_stdout(io.StringIO()):
        for s, (ul, up) in flags.items():
            captured.clear()
            ins.insert_embeddings(s, use_lexical=ul, use_positional=up)
            rows = captured.get("rows", [])
            align[s] = [[a, b, fl(sc), ty] for a, b, sc, ty in rows]
            tables[s]["alignments"] = [(ty, sc) for _, _, sc, ty in rows]

    expect = {}
    with redirect_stdout(io.StringIO()):
        ev.print_metrics_report()
    expect["metrics_json"] = (tmp / "metrics.json").read_text()
    s = "vanilla_clip"
    expect["pairs"] = [list(p) for p in ev.get_image_text_pairs(s)]
    expect["top10"] = {i["image_id"]: [[c, fl(v)] for c, v in ev.get_top_k_similar_chunks(i["image_id"], s, 10)]
                       for i in images}
    expect["top100"] = {i["image_id"]: [c for c, _ in ev.get_top_k_similar_chunks(i["image_id"], s, 100)]
                        for i in images}
    expect["top_k_1_5_10_20"] = {str(k): fl(v) for k, v in ev.compute_top_k_accuracy(s, [1, 5, 10, 20]).items()}
    expect["mrr"] = fl(ev.compute_mrr(s))
    expect["avg_similarity"] = fl(ev.compute_average_similarity(s))
    expect["pair_similarity"] = [fl(ev.compute_similarity(a, b, s)) for a, b, _, _ in expect["pairs"]]
    expect["weak_scores"] = {sc: {k: [fl(x) for x in v] for k, v in ev.get_weak_supervision_scores(sc).items()}
                             for sc in SCHEMAS}
    expect["alignments"] = align
    np.savez_compressed(OUT / "small_corpus.npz", img_emb=img_emb, chk_emb=chk_emb)
    (OUT / "small_corpus.json").write_text(json.dumps(
        dict(images=images, chunks=chunks, lexical_components=lexical_components, expect=expect), indent=1))


# --------------------------------------------------------------------------- config 1
def config1():
    N, M, D = 1000, 5000, 512
    img, chk, meta = synthetic.make_numpy(N, M, D, seed=0x5EED0001)
    h = hashlib.sha256()
    for a in (img["emb"], chk["emb"], img["key"], chk["key"]):
        h.update(np.ascontiguousarray(a).tobytes())
    man = lambda k: f"man{int(k) >> 32}"
    page = lambda k: int(int(k) & 0xFFFFFFFF)
    t = dict(image_ids=[f"img{i}" for i in range(N)], image_manual=[man(k) for k in img["key"]],
             image_page=[page(k) for k in img["key"]], image_emb=img["emb"],
             chunk_ids=[f"chk{j}" for j in range(M)], chunk_manual=[man(k) for k in chk["key"]],
             chunk_page=[page(k) for k in chk["key"]], chunk_emb=chk["emb"], alignments=[])

    class FastDB(rh.FakeDB):  # same answers, page lookup by dict instead of a scan
        pass
    db = FastDB({"vanilla_clip": t})
    index = {}
    for j in range(M):
        index.setdefault((t["chunk_manual"][j], t["chunk_page"][j]), []).append(j)
    rh._Cursor._same_page = lambda self, tt, i: index.get((tt["image_manual"][i], tt["image_page"][i]), [])
    ev, _ = rh.load_reference(db, output_dir=Path(tempfile.mkdtemp()))
    s = "vanilla_clip"
    out = dict(N=N, M=M, D=D, seed=0x5EED0001, input_sha256=h.hexdigest(),
               num_pairs=len(ev.get_image_text_pairs(s)),
               top_k={str(k): fl(v) for k, v in ev.compute_top_k_accuracy(s, [1, 5, 10]).items()},
               top_k_20={str(k): fl(v) for k, v in ev.compute_top_k_accuracy(s, [1, 5, 10, 20]).items()},
               mrr=fl(ev.compute_mrr(s)), avg_similarity=fl(ev.compute_average_similarity(s)),
               db_connections=db.n_connect)
    (OUT / "config1.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    assert rh.available(), "needs /root/reference"
    _, ins = rh.load_reference()
    (OUT / "weak_vectors.json").write_text(json.dumps(weak_vectors(ins), indent=0))
    small_corpus()
    config1()
    print("golden vectors written to", OUT)
