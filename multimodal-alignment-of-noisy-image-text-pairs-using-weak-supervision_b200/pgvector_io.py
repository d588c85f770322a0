"""pgvector / PostgreSQL interop: bulk COPY streams <-> the dense arrays of the C ABI.

The reference keeps its corpus in the `images` / `text_chunks` tables of src/setup_vector_db.py:102-131
(`clip_embedding vector(D)`, `bbox REAL[]`, `page INTEGER`, `manual_id VARCHAR`, ids `VARCHAR UNIQUE`,
`id SERIAL`).  A deployment that already has those tables can feed this path without re-encoding:

    COPY (SELECT image_id, manual_id, page, bbox, clip_embedding FROM vanilla_clip.images ORDER BY id)
        TO STDOUT WITH (FORMAT binary)

and get rankings back as a COPY stream for an `alignments`-style table.  `ORDER BY id` keeps the insertion
order, which is the dense index order ("lower index wins a tie").

Formats follow the PostgreSQL documentation of COPY BINARY (signature, int32 flags, int32 extension length,
per tuple an int16 field count and per field an int32 byte length, -1 = NULL; trailer int16 -1), the array
wire format (ndim, has-null flag, element OID, per dimension length + lower bound, per element length + data)
and pgvector's `vector_send` (int16 dim, int16 unused, dim big-endian float4).  No PostgreSQL server exists in
the build environment: the tests round-trip streams built to those specifications ("parity unpinned" against a
live server, like the cosine itself -- DESIGN.md section 2).

Two decoders: read_copy_binary (pure Python / numpy, field by field -- the specification written down, and the
checker of the other one) and records_from_copy(..., engine=...), which walks the tuples in C (mmalign_copy_scan)
and decodes the bulk columns -- vectors, boxes, pages -- on the GPU (mmalign_copy_decode, csrc/ingest.cu).
"""
from __future__ import annotations

import io
import struct
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

SIGNATURE = b"PGCOPY\n\xff\r\n\0"
FLOAT4_OID = 700

# column kinds understood here
TEXT, INT4, FLOAT4, FLOAT8, FLOAT4_ARRAY, VECTOR = "text", "int4", "float4", "float8", "float4[]", "vector"


# --------------------------------------------------------------------------------------------- decoding
def _decode_field(kind: str, b: memoryview):
    if kind == TEXT:
        return bytes(b).decode("utf-8")
    if kind == INT4:
        return struct.unpack(">i", b)[0]
    if kind == FLOAT4:
        return struct.unpack(">f", b)[0]
    if kind == FLOAT8:
        return struct.unpack(">d", b)[0]
    if kind == VECTOR:  # pgvector vector_send: int16 dim, int16 unused, float4[dim]
        dim, unused = struct.unpack_from(">hh", b, 0)
        if unused != 0 or len(b) != 4 + 4 * dim:
            raise ValueError("malformed vector field")
        return np.frombuffer(b, dtype=">f4", count=dim, offset=4).astype(np.float32)
    if kind == FLOAT4_ARRAY:
        ndim, _hasnull, oid = struct.unpack_from(">iii", b, 0)
        if ndim == 0:
            return []
        if ndim != 1 or oid != FLOAT4_OID:
            raise ValueError(f"expected a one-dimensional float4 array (ndim={ndim}, element oid={oid})")
        n, _lbound = struct.unpack_from(">ii", b, 12)
        out, o = [], 20
        for _ in range(n):
            ln = struct.unpack_from(">i", b, o)[0]
            o += 4
            if ln == -1:
                out.append(None)
            else:
                out.append(struct.unpack_from(">f", b, o)[0])
                o += ln
        return out
    raise ValueError(f"unknown column kind {kind!r}")


def read_copy_binary(data: bytes, columns: Sequence[Tuple[str, str]]) -> Dict[str, list]:
    """Parses a `COPY ... TO STDOUT WITH (FORMAT binary)` stream.  columns = [(name, kind), ...] in SELECT order.
    Returns {name: list of values (None for SQL NULL)}; vector columns hold float32 arrays."""
    mv = memoryview(data)
    if bytes(mv[:11]) != SIGNATURE:
        raise ValueError("not a PostgreSQL binary COPY stream")
    flags, ext = struct.unpack_from(">ii", mv, 11)
    if flags & (1 << 16):
        raise ValueError("COPY streams WITH OIDS are not supported")
    o = 19 + ext
    out: Dict[str, list] = {name: [] for name, _ in columns}
    while True:
        nf = struct.unpack_from(">h", mv, o)[0]
        o += 2
        if nf == -1:
            break
        if nf != len(columns):
            raise ValueError(f"tuple with {nf} fields, {len(columns)} columns declared")
        for name, kind in columns:
            ln = struct.unpack_from(">i", mv, o)[0]
            o += 4
            if ln == -1:
                out[name].append(None)
            else:
                out[name].append(_decode_field(kind, mv[o:o + ln]))
                o += ln
    return out


# --------------------------------------------------------------------------------------------- encoding
def _encode_field(kind: str, v) -> bytes:
    if kind == TEXT:
        return str(v).encode("utf-8")
    if kind == INT4:
        return struct.pack(">i", int(v))
    if kind == FLOAT4:
        return struct.pack(">f", float(v))
    if kind == FLOAT8:
        return struct.pack(">d", float(v))
    if kind == VECTOR:
        a = np.asarray(v, dtype=">f4")
        return struct.pack(">hh", a.shape[0], 0) + a.tobytes()
    if kind == FLOAT4_ARRAY:
        if len(v) == 0:
            return struct.pack(">iii", 0, 0, FLOAT4_OID)
        body = b"".join(struct.pack(">i", -1) if x is None else struct.pack(">if", 4, float(x)) for x in v)
        return struct.pack(">iiiii", 1, int(any(x is None for x in v)), FLOAT4_OID, len(v), 1) + body
    raise ValueError(f"unknown column kind {kind!r}")


def write_copy_binary(columns: Sequence[Tuple[str, str]], rows: Iterable[Sequence]) -> bytes:
    """The inverse of read_copy_binary: a stream `COPY table (cols...) FROM STDIN WITH (FORMAT binary)` accepts."""
    buf = io.BytesIO()
    buf.write(SIGNATURE + struct.pack(">ii", 0, 0))
    for row in rows:
        buf.write(struct.pack(">h", len(columns)))
        for (_, kind), v in zip(columns, row):
            if v is None:
                buf.write(struct.pack(">i", -1))
            else:
                b = _encode_field(kind, v)
                buf.write(struct.pack(">i", len(b)) + b)
    buf.write(struct.pack(">h", -1))
    return buf.getvalue()


# --------------------------------------------------------------------------------------------- text forms
def parse_vector_text(s: str) -> np.ndarray:
    """pgvector's text form '[0.1,0.2,...]' -- what `SELECT clip_embedding` returns to psycopg2
    (src/evaluate_alignments.py:86-93) -- as float32, each component rounded like the server's float4 input."""
    s = s.strip()
    if not (s.startswith("[") and s.endswith("]")):
        raise ValueError("not a pgvector literal")
    body = s[1:-1].strip()
    return np.array([float(x) for x in body.split(",")] if body else [], dtype=np.float32)


def parse_real_array_text(s: Optional[str]) -> Optional[List[Optional[float]]]:
    """PostgreSQL's text form of REAL[]: '{72,120.5,300.25,410.75}', '{}' or None (SQL NULL)."""
    if s is None:
        return None
    s = s.strip()
    if not (s.startswith("{") and s.endswith("}")):
        raise ValueError("not an array literal")
    body = s[1:-1].strip()
    if not body:
        return []
    return [None if x.strip().upper() == "NULL" else float(np.float32(float(x))) for x in body.split(",")]


# --------------------------------------------------------------------------------------------- tables <-> corpus
IMAGE_COLUMNS = [("image_id", TEXT), ("manual_id", TEXT), ("page", INT4), ("bbox", FLOAT4_ARRAY), ("clip_embedding", VECTOR)]
CHUNK_COLUMNS = [("chunk_id", TEXT), ("manual_id", TEXT), ("page", INT4), ("text", TEXT), ("bbox", FLOAT4_ARRAY),
                 ("clip_embedding", VECTOR)]


def _records_from_copy_gpu(data: bytes, columns: Sequence[Tuple[str, str]], engine):
    names = [n for n, _ in columns]
    kinds = dict(columns)
    col = lambda name: names.index(name) if name in names else -1
    off, ln, emb, bbox, page, page_null = engine.copy_decode(data, len(columns), col("clip_embedding"), col("bbox"), col("page"))
    mv = memoryview(data)
    recs = []
    for i in range(off.shape[0]):
        r = {}
        for f, name in enumerate(names):
            if name == "clip_embedding":
                continue
            if name == "page":
                r[name] = None if page_null[i] else int(page[i])
            elif name == "bbox":  # NULL stays None; anything else that is not four numbers decoded to zeros (scores 0.0)
                r[name] = None if ln[i, f] < 0 else [float(x) for x in bbox[i]]
            elif ln[i, f] < 0:
                r[name] = None
            elif kinds[name] == TEXT:
                r[name] = bytes(mv[off[i, f]:off[i, f] + ln[i, f]]).decode("utf-8")
            else:
                r[name] = _decode_field(kinds[name], mv[off[i, f]:off[i, f] + ln[i, f]])
        recs.append(r)
    return recs, emb


def records_from_copy(data: bytes, columns: Sequence[Tuple[str, str]], engine=None):
    """COPY stream of the `images` / `text_chunks` table -> (records like the reference's JSON files, embeddings [n, D]).
    The records feed corpus.build_corpus unchanged.  With an engine the tuple walk runs in C and the bulk columns are
    decoded on the GPU; a bbox that is not four non-NULL numbers then reads as [0, 0, 0, 0], which scores 0.0 exactly like
    the missing box it stands for (src/insert_clip_embeddings.py:161-169)."""
    if engine is not None and hasattr(engine, "copy_decode"):
        return _records_from_copy_gpu(data, columns, engine)
    cols = read_copy_binary(data, columns)
    names = [n for n, _ in columns if n != "clip_embedding"]
    n = len(cols[columns[0][0]])
    recs = [{k: cols[k][i] for k in names} for i in range(n)]
    embs = cols["clip_embedding"]
    if any(e is None for e in embs):
        raise ValueError("NULL clip_embedding: the tables declare it NOT NULL (src/setup_vector_db.py:110, :126)")
    D = len(embs[0]) if n else 0
    if any(len(e) != D for e in embs):
        raise ValueError("clip_embedding rows of different dimension")
    emb = np.stack(embs).astype(np.float32) if n else np.zeros((0, 0), np.float32)
    return recs, emb


def corpus_from_copy(images_copy: bytes, chunks_copy: bytes, lexical_components=None, engine=None):
    """Two COPY streams (IMAGE_COLUMNS / CHUNK_COLUMNS, ORDER BY id) -> corpus.Corpus, ready for
    evaluate_alignments.register_schema.  bbox values are float4 on this route (the table type is REAL[]); the
    reference computes its alignment records from the JSON doubles before they are stored."""
    from .corpus import build_corpus
    images, ie = records_from_copy(images_copy, IMAGE_COLUMNS, engine)
    chunks, ce = records_from_copy(chunks_copy, CHUNK_COLUMNS, engine)
    return build_corpus(images, chunks, ie, ce, lexical_components, engine=engine)


RANKING_COLUMNS = [("image_id", TEXT), ("chunk_id", TEXT), ("rank", INT4), ("similarity", FLOAT8)]


def rankings_to_copy(corpus, topk_idx: np.ndarray, topk_score: np.ndarray, row0: int = 0) -> bytes:
    """Top-K lists of one schema ([rows, K] global chunk indices / scores, -1 padded; rows are images row0..) as a
    binary COPY stream of (image_id, chunk_id, rank, similarity) rows."""
    def rows():
        for r in range(topk_idx.shape[0]):
            for k in range(topk_idx.shape[1]):
                j = int(topk_idx[r, k])
                if j < 0:
                    break
                yield (corpus.image_ids[row0 + r], corpus.chunk_ids[j], k + 1, float(topk_score[r, k]))
    return write_copy_binary(RANKING_COLUMNS, rows())
