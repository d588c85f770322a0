"""Drop-in for the weak-supervision half of the reference's src/insert_clip_embeddings.py.

Same function names and argument meaning; the arithmetic runs in the CUDA library
(rescore.cu: alignments_kernel, fp64, no fused multiply-add) -- no CPU fallback.
CLIP encoding and the database inserts of that file are out of scope.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np

from .corpus import bbox_array, term_bitsets
from .engine import AlignmentEngine, default_engine


def _engine() -> AlignmentEngine:
    return default_engine(0, "scratch")


def _records(eng, images, chunks, img_key, chk_key, terms, schema, raw=False):
    e = np.zeros((len(images), 4), np.float32)
    e[:, 0] = 1.0
    f = np.zeros((len(chunks), 4), np.float32)
    f[:, 0] = 1.0
    eng.set_images(e, img_key, bbox_array(images), None)
    eng.set_chunks(f, chk_key, bbox_array(chunks), term_bitsets(chunks, terms, eng), n_terms=len(terms))
    return eng.alignments(schema, raw=raw)


def compute_lexical_alignment(text_chunk: Dict, lexical_components: List[str]) -> float:
    """src/insert_clip_embeddings.py:144-156."""
    rec = _records(_engine(), [{}], [text_chunk], np.zeros(1, np.uint64), np.zeros(1, np.uint64),
                   list(lexical_components), "clip_lexical", raw=True)
    return float(rec[0, 0])


def compute_positional_alignment(image: Dict, chunk: Dict) -> float:
    """src/insert_clip_embeddings.py:159-210."""
    rec = _records(_engine(), [image], [{"text": "", **chunk}], np.zeros(1, np.uint64), np.zeros(1, np.uint64),
                   [], "clip_positional", raw=True)
    return float(rec[0, 1])


SCHEMA_FLAGS = {(False, False): "vanilla_clip", (True, False): "clip_lexical",
                (False, True): "clip_positional", (True, True): "clip_combined"}


def compute_alignment_records(corpus, use_lexical: bool, use_positional: bool,
                              engine: AlignmentEngine | None = None) -> List[Tuple[str, str, float, str]]:
    """The `alignment_records` list of src/insert_clip_embeddings.py:369-414 for a Corpus, in
    the reference's loop order (image-major, chunk order; lexical before positional).

    That loop compares manuals and pages with `!=` (:377-380), so page None matches page None -- unlike
    the SQL join of the evaluation, which never joins NULLs.  The records follow the loop: the tables go to
    the library with the keys of the Python-side join (corpus.page_keys(python_join=True))."""
    if not (use_lexical or use_positional):
        return []
    eng = engine or _engine()
    eng.set_images(corpus.img["emb"], corpus.img.get("key_py", corpus.img["key"]), corpus.img["bbox"], None)
    eng.set_chunks(corpus.chk["emb"], corpus.chk.get("key_py", corpus.chk["key"]), corpus.chk["bbox"], corpus.chk["terms"],
                   n_terms=corpus.n_terms)
    off, pc = eng.pairs()
    rec = eng.alignments(SCHEMA_FLAGS[(bool(use_lexical), bool(use_positional))])
    out = []
    names = ("lexical", "positional", "combined")
    for i in range(len(corpus.image_ids)):
        for p in range(off[i], off[i + 1]):
            for t in range(3):
                if rec[p, t] != 0.0:
                    out.append((corpus.image_ids[i], corpus.chunk_ids[pc[p]], float(rec[p, t]), names[t]))
    return out
