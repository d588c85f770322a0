"""pgvector / PostgreSQL COPY interop (pgvector_io.py): the byte layouts of the PostgreSQL documentation and of
pgvector's vector_send, written out by hand here, and a round trip of the golden corpus through both tables."""
import importlib
import json
import struct

import numpy as np
import pytest

from conftest import GOLDEN, PKG_NAME, OracleIngest


@pytest.fixture(scope="module")
def pg():
    return importlib.import_module(PKG_NAME + ".pgvector_io")


def test_hand_written_stream(pg):
    """One tuple (text, int4 NULL, float4[] with a NULL element, vector) and the trailer, byte by byte."""
    stream = (b"PGCOPY\n\xff\r\n\0" + b"\x00\x00\x00\x00" + b"\x00\x00\x00\x00"      # signature, flags, extension
              + b"\x00\x04"                                                             # 4 fields
              + b"\x00\x00\x00\x05" + "imég".encode()                                   # text 'imég' (5 bytes in UTF-8)
              + b"\xff\xff\xff\xff"                                                     # int4 NULL
              + b"\x00\x00\x00\x20"                                                     # float4[]: 32 bytes
              + struct.pack(">iiiii", 1, 1, 700, 2, 1) + b"\x00\x00\x00\x04\x3f\xc0\x00\x00" + b"\xff\xff\xff\xff"
              + b"\x00\x00\x00\x0c" + b"\x00\x02\x00\x00" + b"\x3f\x80\x00\x00" + b"\xc0\x20\x00\x00"   # vector [1, -2.5]
              + b"\xff\xff")
    cols = [("id", pg.TEXT), ("page", pg.INT4), ("bbox", pg.FLOAT4_ARRAY), ("emb", pg.VECTOR)]
    r = pg.read_copy_binary(stream, cols)
    assert r["id"] == ["imég"] and r["page"] == [None] and r["bbox"] == [[1.5, None]]
    assert r["emb"][0].dtype == np.float32 and r["emb"][0].tolist() == [1.0, -2.5]
    assert pg.write_copy_binary(cols, [("imég", None, [1.5, None], [1.0, -2.5])]) == stream
    with pytest.raises(ValueError):
        pg.read_copy_binary(stream, cols[:2])                                           # field count mismatch
    with pytest.raises(ValueError):
        pg.read_copy_binary(b"not a copy stream at all....", cols)


def test_text_forms(pg):
    assert pg.parse_vector_text("[0.1,-2,3e-3]").tolist() == [np.float32(0.1), -2.0, np.float32(3e-3)]
    assert pg.parse_vector_text("[]").shape == (0,)
    assert pg.parse_real_array_text("{72,120.5,300.25,410.75}") == [72.0, 120.5, 300.25, 410.75]
    assert pg.parse_real_array_text("{}") == [] and pg.parse_real_array_text(None) is None
    assert pg.parse_real_array_text("{1,NULL}") == [1.0, None]
    assert pg.parse_real_array_text("{0.1}") == [float(np.float32(0.1))]           # REAL is float4
    with pytest.raises(ValueError):
        pg.parse_vector_text("{1,2}")


def test_golden_corpus_round_trip(pg, pkg, small_corpus):
    """images / text_chunks of the golden corpus -> COPY streams -> Corpus: same ids, page keys, embeddings and term
    sets as the corpus built from the JSON records; bboxes are the float4 roundings (the table type is REAL[])."""
    d, want = small_corpus
    z = np.load(GOLDEN / "small_corpus.npz")

    def bbox(r):
        b = r.get("bbox")
        return None if b is None else list(b)
    img_rows = [(r["image_id"], r.get("manual_id"), r.get("page"), bbox(r), z["img_emb"][i]) for i, r in enumerate(d["images"])]
    chk_rows = [(r["chunk_id"], r.get("manual_id"), r.get("page"), r["text"], bbox(r), z["chk_emb"][j])
                for j, r in enumerate(d["chunks"])]
    got = pg.corpus_from_copy(pg.write_copy_binary(pg.IMAGE_COLUMNS, img_rows), pg.write_copy_binary(pg.CHUNK_COLUMNS, chk_rows),
                              d["lexical_components"], engine=OracleIngest())
    assert got.image_ids == want.image_ids and got.chunk_ids == want.chunk_ids
    assert np.array_equal(got.img["key"], want.img["key"]) and np.array_equal(got.chk["key"], want.chk["key"])
    assert np.array_equal(got.img["emb"], want.img["emb"]) and np.array_equal(got.chk["emb"], want.chk["emb"])
    assert np.array_equal(got.chk["terms"], want.chk["terms"])
    assert np.array_equal(got.img["bbox"], want.img["bbox"].astype(np.float32).astype(np.float64))
    assert np.array_equal(got.chk["bbox"], want.chk["bbox"].astype(np.float32).astype(np.float64))
    # rankings back out: (image_id, chunk_id, rank, similarity) rows, padding dropped
    idx = np.array([[3, 1, -1], [0, -1, -1]])
    sc = np.array([[0.5, 0.25, -np.inf], [0.125, -np.inf, -np.inf]])
    back = pg.read_copy_binary(pg.rankings_to_copy(got, idx, sc, row0=2), pg.RANKING_COLUMNS)
    assert back["image_id"] == [got.image_ids[2]] * 2 + [got.image_ids[3]]
    assert back["chunk_id"] == [got.chunk_ids[3], got.chunk_ids[1], got.chunk_ids[0]]
    assert back["rank"] == [1, 2, 1] and back["similarity"] == [0.5, 0.25, 0.125]


def _random_stream(pg, rng, n, D):
    """A text_chunks-shaped stream with ragged ids and texts (fields at every byte alignment), NULL pages, NULL / empty /
    three-element / NULL-element boxes."""
    rows, embs = [], rng.standard_normal((n, D)).astype(np.float32)
    for i in range(n):
        kind = int(rng.integers(0, 8))
        box = [float(np.float32(x)) for x in rng.uniform(0, 700, 4)]
        bbox = {0: None, 1: [], 2: box[:3], 3: [box[0], None, box[2], box[3]]}.get(kind, box)
        rows.append((f"chunk{'x' * int(rng.integers(0, 7))}{i}", None if kind == 5 else "man" + "é" * int(rng.integers(0, 3)),
                     None if kind == 6 else int(rng.integers(-5, 4000)), "t" * int(rng.integers(0, 40)), bbox, embs[i]))
    return pg.write_copy_binary(pg.CHUNK_COLUMNS, rows), rows, embs


def test_copy_scan_matches_the_python_reader(pg, pkg):
    """mmalign_copy_scan (the C walk of the tuples; host-only, no GPU needed) finds the fields the Python reader reads."""
    import ctypes as C
    L = pkg._native.load()
    rng = np.random.default_rng(5)
    data, rows, _ = _random_stream(pg, rng, 300, 16)
    buf = np.frombuffer(data, np.uint8)
    nc = len(pg.CHUNK_COLUMNS)
    n = L.mmalign_copy_scan(buf.ctypes.data, len(buf), nc, None, None, 0)
    assert n == 300
    off, ln = np.zeros((n, nc), np.int64), np.zeros((n, nc), np.int32)
    assert L.mmalign_copy_scan(buf.ctypes.data, len(buf), nc, off.ctypes.data, ln.ctypes.data, n) == n
    want = pg.read_copy_binary(data, pg.CHUNK_COLUMNS)
    for i in range(n):
        assert data[off[i, 0]:off[i, 0] + ln[i, 0]].decode() == want["chunk_id"][i]
        assert (ln[i, 1] == -1) == (want["manual_id"][i] is None) and (ln[i, 2] == -1) == (want["page"][i] is None)
        assert ln[i, 5] == 4 + 4 * 16
    # error codes: bad signature, truncation, field count, capacity
    assert L.mmalign_copy_scan(buf.ctypes.data + 1, len(buf) - 1, nc, None, None, 0) == -1
    assert L.mmalign_copy_scan(buf.ctypes.data, len(buf) - 5, nc, None, None, 0) == -2
    assert L.mmalign_copy_scan(buf.ctypes.data, len(buf), nc - 1, None, None, 0) == -3
    assert L.mmalign_copy_scan(buf.ctypes.data, len(buf), nc, off.ctypes.data, ln.ctypes.data, n - 1) == -4
    empty = np.frombuffer(pg.write_copy_binary(pg.CHUNK_COLUMNS, []), np.uint8)
    assert L.mmalign_copy_scan(empty.ctypes.data, len(empty), nc, None, None, 0) == 0


@pytest.mark.gpu
def test_copy_decode_on_the_gpu_matches_the_python_reader(pg, pkg, small_corpus):
    """mmalign_copy_decode: vectors, boxes and pages of a COPY stream decoded on the GPU equal the field-by-field Python
    decode; the golden corpus ingested that way gives the same Corpus."""
    eng = pkg.AlignmentEngine(0)
    try:
        rng = np.random.default_rng(6)
        for n, D in [(1, 4), (257, 64), (1000, 512), (33, 1024)]:
            data, rows, embs = _random_stream(pg, rng, n, D)
            recs, emb = pg.records_from_copy(data, pg.CHUNK_COLUMNS, engine=eng)
            want_recs, want_emb = pg.records_from_copy(data, pg.CHUNK_COLUMNS)
            assert np.array_equal(emb, want_emb) and np.array_equal(emb, embs)
            for g, w in zip(recs, want_recs):
                assert g["chunk_id"] == w["chunk_id"] and g["manual_id"] == w["manual_id"] and g["page"] == w["page"]
                assert g["text"] == w["text"]
                ok = w["bbox"] is not None and len(w["bbox"]) == 4 and None not in w["bbox"]
                assert g["bbox"] == (w["bbox"] if ok else (None if w["bbox"] is None else [0.0] * 4))
        assert pg.records_from_copy(pg.write_copy_binary(pg.CHUNK_COLUMNS, []), pg.CHUNK_COLUMNS, engine=eng)[0] == []
        with pytest.raises(pkg.MMAlignError):   # a vector of another dimension
            bad = pg.write_copy_binary(pg.CHUNK_COLUMNS, [("a", "m", 1, "t", None, np.ones(8, np.float32)),
                                                         ("b", "m", 1, "t", None, np.ones(4, np.float32))])
            pg.records_from_copy(bad, pg.CHUNK_COLUMNS, engine=eng)
        # the golden corpus through both decoders
        d, want = small_corpus
        z = np.load(GOLDEN / "small_corpus.npz")
        box = lambda r: None if r.get("bbox") is None else list(r["bbox"])
        img = pg.write_copy_binary(pg.IMAGE_COLUMNS, [(r["image_id"], r.get("manual_id"), r.get("page"), box(r), z["img_emb"][i])
                                                      for i, r in enumerate(d["images"])])
        chk = pg.write_copy_binary(pg.CHUNK_COLUMNS, [(r["chunk_id"], r.get("manual_id"), r.get("page"), r["text"], box(r),
                                                      z["chk_emb"][j]) for j, r in enumerate(d["chunks"])])
        a = pg.corpus_from_copy(img, chk, d["lexical_components"], engine=eng)
        b = pg.corpus_from_copy(img, chk, d["lexical_components"], engine=OracleIngest())
        for side in ("img", "chk"):
            for f in ("emb", "key", "bbox", "key_py"):
                assert np.array_equal(getattr(a, side)[f], getattr(b, side)[f]), (side, f)
        assert np.array_equal(a.chk["terms"], b.chk["terms"]) and a.image_ids == b.image_ids and a.chunk_ids == b.chunk_ids
    finally:
        eng.close()
