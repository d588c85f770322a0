"""Host-side ingest: the reference's record formats -> the dense arrays of the C ABI.

Record formats: data/processed/image_metadata.json (src/pdf_processor.py:406-415),
text_chunks.json (:686-692) and filtered_lexical_components.json (:1013-1019, read at
src/insert_clip_embeddings.py:232-248).  String ids never cross the ABI: items keep
their list order as dense indices ("lower index wins a tie" = table insertion order).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

NULL_KEY = np.uint64(0xFFFFFFFFFFFFFFFF)


def page_keys(records: Sequence[dict], page_ids: Dict[tuple, int], python_join: bool = False) -> np.ndarray:
    """(manual_id, page) -> dense 64-bit key.

    SQL join (default; src/evaluate_alignments.py:61 `i.manual_id = t.manual_id AND i.page = t.page`): a NULL
    manual or page never joins, so it becomes MMALIGN_NULL_KEY.
    python_join=True: the alignment loop of insert_embeddings() compares with `!=`
    (src/insert_clip_embeddings.py:377-380), for which None equals None -- every (manual_id, page) tuple,
    None included, gets a real key."""
    out = np.empty(len(records), np.uint64)
    for i, r in enumerate(records):
        page = r.get("page")
        if not python_join and (page is None or r.get("manual_id") is None):
            out[i] = NULL_KEY
        else:
            out[i] = page_ids.setdefault((r.get("manual_id"), page), len(page_ids))
    return out


def bbox_array(records: Sequence[dict]) -> np.ndarray:
    """Missing / wrong-length boxes become all-zero rows, which score 0.0 exactly like
    src/insert_clip_embeddings.py:161-169 (the zero-width test :172 catches them)."""
    out = np.zeros((len(records), 4), np.float64)
    for i, r in enumerate(records):
        b = r.get("bbox")
        if b and len(b) == 4:
            out[i] = b
    return out


def _engine():
    from .engine import default_engine
    return default_engine(0, "scratch")


def term_bitsets(chunks: Sequence[dict], terms: Sequence[str], engine=None) -> np.ndarray:
    """bit t of row j = terms[t] occurs as a substring of chunk j's lower-cased text
    (src/insert_clip_embeddings.py:149-150).  str.lower() (Unicode-aware) runs here; the matching of every
    (chunk, term) pair runs on the GPU (csrc/ingest.cu) -- there is no host fallback."""
    eng = engine if engine is not None else _engine()
    return eng.term_bitsets([c["text"].lower() for c in chunks], list(terms))


@dataclass
class Corpus:
    image_ids: List[str]
    chunk_ids: List[str]
    image_manual: List[str]
    image_page: list
    img: dict                 # emb, key, bbox, terms(None), key_py (keys of the Python-side join, see page_keys)
    chk: dict                 # emb, key, bbox, terms, key_py
    terms: List[str] = field(default_factory=list)
    image_index: Dict[str, int] = field(default_factory=dict)
    chunk_index: Dict[str, int] = field(default_factory=dict)

    @property
    def n_terms(self) -> int:
        return len(self.terms)


def build_corpus(images: Sequence[dict], chunks: Sequence[dict], image_emb, chunk_emb,
                 lexical_components: Optional[dict | Sequence[str]] = None, engine=None) -> Corpus:
    if isinstance(lexical_components, dict):  # insert_clip_embeddings.py:237-239
        terms = [c["term"] for c in lexical_components.get("components", [])]
    else:
        terms = list(lexical_components or [])
    page_ids: Dict[tuple, int] = {}
    img = dict(emb=np.ascontiguousarray(image_emb, np.float32), key=page_keys(images, page_ids),
               bbox=bbox_array(images), terms=None)
    chk = dict(emb=np.ascontiguousarray(chunk_emb, np.float32), key=page_keys(chunks, page_ids),
               bbox=bbox_array(chunks), terms=term_bitsets(chunks, terms, engine))
    py_ids: Dict[tuple, int] = {}
    img["key_py"] = page_keys(images, py_ids, python_join=True)
    chk["key_py"] = page_keys(chunks, py_ids, python_join=True)
    c = Corpus(image_ids=[r["image_id"] for r in images], chunk_ids=[r["chunk_id"] for r in chunks],
               image_manual=[r.get("manual_id") for r in images], image_page=[r.get("page") for r in images],
               img=img, chk=chk, terms=terms)
    c.image_index = {s: i for i, s in enumerate(c.image_ids)}
    c.chunk_index = {s: j for j, s in enumerate(c.chunk_ids)}
    return c
