"""Drop-in for the reference's src/evaluate_alignments.py -- same public functions,
argument meaning, return types, error behaviour and metrics.json layout -- with the
PostgreSQL/pgvector round trips replaced by one pass of the CUDA library per schema.

Where the reference opens a database connection, this module looks the schema up in
an in-process registry filled by `register_schema()` (arrays + record lists, the
content of the `images` / `text_chunks` tables).  A schema that was never registered
behaves like a schema missing from the database.

Limits the reference does not have (the library answers MMALIGN_ELIMIT / MMALIGN_EINVAL = MMAlignError): k and
the MRR cut-off at most 256, at most 512 chunks on one (manual, page).

Two knobs the reference does not have (both default to the reference's behaviour):
  CANDIDATES   "same_page": rank an image against the chunks of its own manual+page
               only (the SQL join, :128-131); "all": rank against every chunk.
  WEAK_WEIGHT  (lam_lex, lam_pos): weight of the schema's weak-supervision records in
               the ranking score; (0, 0) = the reference, which never ranks by them.
"""
from __future__ import annotations

import json
from collections import defaultdict
from pathlib import Path
from typing import Dict, List, Tuple

import numpy as np

from .corpus import Corpus
from .engine import AlignmentEngine, SCHEMAS as _ALL_SCHEMAS, default_engine
from .insert_clip_embeddings import compute_alignment_records

# Schemas to evaluate (src/evaluate_alignments.py:29)
SCHEMAS = list(_ALL_SCHEMAS)

# Output directory (src/evaluate_alignments.py:32-34); created on first write, not at import
OUTPUT_DIR = Path.cwd() / "evaluation_results"

CANDIDATES = "same_page"
WEAK_WEIGHT = (0.0, 0.0)
DEVICE = 0

_FLAGS = {"vanilla_clip": (False, False), "clip_lexical": (True, False),
          "clip_positional": (False, True), "clip_combined": (True, True)}
_REGISTRY: Dict[str, "_Schema"] = {}


def _engine() -> AlignmentEngine:
    return default_engine(DEVICE)


class _Schema:
    def __init__(self, name: str, corpus: Corpus):
        self.name, self.corpus = name, corpus
        self.cache = {}

    def results(self, depth: int):
        """One library pass: per-image top-`depth` lists, per-pair ranks and similarities."""
        key = (CANDIDATES, tuple(WEAK_WEIGHT))
        hit = self.cache.get(key)
        if hit is not None and hit["depth"] >= depth:
            return hit
        c, eng = self.corpus, _engine()
        eng.set_images(c.img["emb"], c.img["key"], c.img["bbox"], c.img["terms"])
        eng.set_chunks(c.chk["emb"], c.chk["key"], c.chk["bbox"], c.chk["terms"], n_terms=c.n_terms)
        depth = max(depth, 100)
        r = eng.run(self.name, candidates=CANDIDATES, k_values=[depth], mrr_cutoff=depth,
                    weak_weight=WEAK_WEIGHT, want=("topk", "pairs"))
        off, pc = eng.pairs()
        hit = dict(depth=depth, topk_idx=r["topk_idx"][0], topk_score=r["topk_score"][0],
                   pair_rank=r["pair_rank"][0], pair_sim=r["pair_sim"], offsets=off, pair_chunk=pc)
        self.cache[key] = hit
        return hit


def register_schema(schema: str, corpus: Corpus) -> None:
    """Loads one schema's tables (the job of insert_embeddings() in the reference)."""
    _REGISTRY[schema] = _Schema(schema, corpus)


def clear_schemas() -> None:
    _REGISTRY.clear()


def connect_db():
    """The reference opens a PostgreSQL connection here (:37-45); this build has no
    database, the registry stands in for it."""
    return _REGISTRY


def _get(schema: str) -> _Schema:
    if schema not in _REGISTRY:
        raise KeyError(f'relation "{schema}.images" does not exist')
    return _REGISTRY[schema]


def get_image_text_pairs(schema: str) -> List[Tuple[str, str, str, str]]:
    """All (image_id, chunk_id, manual_id, page) with equal manual and page (:48-69)."""
    s = _get(schema)
    r, c = s.results(100), s.corpus
    off, pc = r["offsets"], r["pair_chunk"]
    return [(c.image_ids[i], c.chunk_ids[pc[p]], c.image_manual[i], c.image_page[i])
            for i in range(len(c.image_ids)) for p in range(off[i], off[i + 1])]


def compute_similarity(image_id: str, chunk_id: str, schema: str) -> float:
    """Cosine similarity of one image/chunk pair (:72-106)."""
    s = _get(schema)
    c = s.corpus
    i, j = c.image_index[image_id], c.chunk_index[chunk_id]
    r = s.results(100)
    off, pc = r["offsets"], r["pair_chunk"]
    p = off[i] + np.searchsorted(pc[off[i]:off[i + 1]], j)
    if p < off[i + 1] and pc[p] == j:
        return float(r["pair_sim"][p])
    eng = default_engine(DEVICE, "scratch")  # not a true pair: score it alone, beside the schema's tables
    eng.set_images(c.img["emb"][i:i + 1], np.zeros(1, np.uint64))
    eng.set_chunks(c.chk["emb"][j:j + 1], np.zeros(1, np.uint64))
    return float(eng.run("vanilla_clip", k_values=[1], want=("pairs",))["pair_sim"][0])


def get_top_k_similar_chunks(image_id: str, schema: str, k: int = 10) -> List[Tuple[str, float]]:
    """Top K most similar text chunks for an image (:109-143)."""
    s = _get(schema)
    r = s.results(k)
    i = s.corpus.image_index[image_id]
    out = []
    for j, sc in zip(r["topk_idx"][i, :k], r["topk_score"][i, :k]):
        if j < 0:
            break
        out.append((s.corpus.chunk_ids[j], float(sc)))
    return out


def get_weak_supervision_scores(schema: str) -> Dict[str, List[float]]:
    """Weak supervision alignment scores by type (:146-166); the column is REAL (fp32)."""
    s = _get(schema)
    ul, up = _FLAGS.get(schema, (False, False))
    by_type = defaultdict(list)
    for _, _, score, ty in sorted(compute_alignment_records(s.corpus, ul, up, default_engine(DEVICE, "scratch")),
                                  key=lambda r: r[3]):
        by_type[ty].append(float(np.float32(score)))
    return dict(by_type)


def compute_top_k_accuracy(schema: str, k_values: List[int] = [1, 5, 10]) -> Dict[int, float]:
    """Top-K accuracy over all true pairs (:169-193)."""
    s = _get(schema)
    r = s.results(max(k_values))
    ranks = r["pair_rank"]
    if len(ranks) == 0:
        return {k: 0.0 for k in k_values}
    return {k: int(np.count_nonzero((ranks >= 1) & (ranks <= k))) / len(ranks) for k in k_values}


def compute_mrr(schema: str) -> float:
    """Mean reciprocal rank, truncated at rank 100 (:196-216)."""
    ranks = _get(schema).results(100)["pair_rank"]
    if len(ranks) == 0:
        return 0.0
    return np.mean([1.0 / r if 1 <= r <= 100 else 0.0 for r in ranks.tolist()])


def compute_average_similarity(schema: str) -> float:
    """Average similarity of the true pairs (:219-231)."""
    sims = _get(schema).results(100)["pair_sim"]
    if len(sims) == 0:
        return 0.0
    return np.mean(sims.tolist())


def print_metrics_report():
    """Comprehensive metrics report and evaluation_results/metrics.json (:356-435)."""
    print("\n" + "=" * 80)
    print("MULTIMODAL ALIGNMENT EVALUATION REPORT")
    print("=" * 80 + "\n")
    all_metrics = {}
    for schema in SCHEMAS:
        print(f"\n📊 Schema: {schema.upper().replace('_', ' ')}")
        print("-" * 80)
        try:
            if schema not in _REGISTRY:
                print("  ⚠️  Schema not found in database")
                continue
            top_k_acc = compute_top_k_accuracy(schema, [1, 5, 10])
            mrr = compute_mrr(schema)
            avg_sim = compute_average_similarity(schema)
            pairs = get_image_text_pairs(schema)
            print(f"  Total Image-Text Pairs: {len(pairs)}")
            print(f"  Average Similarity: {avg_sim:.4f}")
            print(f"  Mean Reciprocal Rank (MRR): {mrr:.4f}")
            print(f"  Top-1 Accuracy: {top_k_acc[1]:.4f} ({top_k_acc[1] * 100:.2f}%)")
            print(f"  Top-5 Accuracy: {top_k_acc[5]:.4f} ({top_k_acc[5] * 100:.2f}%)")
            print(f"  Top-10 Accuracy: {top_k_acc[10]:.4f} ({top_k_acc[10] * 100:.2f}%)")
            if schema in ["clip_lexical", "clip_positional", "clip_combined"]:
                try:
                    scores_by_type = get_weak_supervision_scores(schema)
                    if scores_by_type:
                        print("  Weak Supervision Alignments:")
                        for align_type, scores in scores_by_type.items():
                            print(f"    - {align_type}: {len(scores)} pairs, avg score: {np.mean(scores):.4f}")
                except Exception:
                    pass
            all_metrics[schema] = {"top_k": top_k_acc, "mrr": mrr, "avg_similarity": avg_sim,
                                   "num_pairs": len(pairs)}
        except Exception as e:
            print(f"  ❌ Error evaluating schema: {e}")
            continue
    OUTPUT_DIR.mkdir(exist_ok=True)
    metrics_file = OUTPUT_DIR / "metrics.json"
    with open(metrics_file, "w") as f:
        json.dump(all_metrics, f, indent=2)
    print(f"\n✅ Metrics saved to {metrics_file}")
    print("\n" + "=" * 80)


def main():
    """Complete evaluation (:438-456).  The three PNG charts of the reference are
    presentation only and out of scope here."""
    print("🔍 Starting evaluation...")
    try:
        print_metrics_report()
        print(f"\n✅ Evaluation complete! Results saved to {OUTPUT_DIR}/")
    except Exception as e:
        print(f"❌ Evaluation failed: {e}")
        raise


if __name__ == "__main__":
    main()
