"""Runs the UNMODIFIED reference modules against an in-process fake pgvector.

TEST INFRASTRUCTURE ONLY (see oracle/mmalign_oracle.c).  Works only where
/root/reference exists (this container, not the GPU box); it is how the golden
vectors under tests/golden/ were made and how the oracle is pinned.

* The reference's two modules are loaded from where they lie, with the absent
  third-party imports (psycopg2, matplotlib, seaborn, dotenv) stubbed in
  sys.modules and pathlib.Path.mkdir neutralised during the import
  (src/evaluate_alignments.py:34 would otherwise create a directory inside the
  read-only tree).
* FakeDB answers the seven SQL shapes of SURVEY.md Appendix B from numpy arrays.
  The one piece of arithmetic it supplies itself is pgvector's `<=>`
  (oracle.cosine: the restated algorithm, "parity unpinned"); everything else
  -- pair enumeration, top-K membership, MRR, means -- is the reference's own code.
"""
from __future__ import annotations

import importlib.util
import re
import sys
import types
from pathlib import Path

import numpy as np

from . import oracle

REF = Path("/root/reference")


def available() -> bool:
    return (REF / "src" / "evaluate_alignments.py").exists()


class FakeDB:
    """tables[schema] = dict(image_ids, image_manual, image_page, image_emb,
    chunk_ids, chunk_manual, chunk_page, chunk_emb, alignments=[(type, score)])."""

    def __init__(self, tables):
        self.tables = tables
        self.n_connect = 0
        for t in tables.values():
            t["_img_index"] = {s: i for i, s in enumerate(t["image_ids"])}
            t["_chk_index"] = {s: i for i, s in enumerate(t["chunk_ids"])}

    def connect(self, **kw):
        self.n_connect += 1
        return _Conn(self)


class _Conn:
    def __init__(self, db):
        self.db = db

    def cursor(self):
        return _Cursor(self.db)

    def close(self):
        pass

    def commit(self):
        pass


class _Cursor:
    def __init__(self, db):
        self.db = db
        self.rows = []

    def close(self):
        pass

    def fetchone(self):
        return self.rows[0] if self.rows else None

    def fetchall(self):
        return list(self.rows)

    def _same_page(self, t, i):
        man, page = t["image_manual"][i], t["image_page"][i]
        if page is None:
            return []  # SQL: NULL never equals NULL in a join condition
        return [j for j in range(len(t["chunk_ids"]))
                if t["chunk_manual"][j] == man and t["chunk_page"][j] == page]

    def execute(self, sql, params=None):
        q = " ".join(sql.split())
        m = re.search(r"FROM (\w+)\.(images|text_chunks|alignments)", q)
        schema = m.group(1) if m else None
        t = self.db.tables.get(schema) if schema else None
        if "information_schema.tables" in q:                       # S7
            self.rows = [(1 if params[0] in self.db.tables else 0,)]
        elif q.startswith("SELECT DISTINCT i.image_id, t.chunk_id"):  # S1
            rows = []
            for i, iid in enumerate(t["image_ids"]):
                for j in self._same_page(t, i):
                    rows.append((iid, t["chunk_ids"][j], t["image_manual"][i], t["image_page"][i]))
            self.rows = rows
        elif q.startswith("SELECT clip_embedding FROM") and ".images WHERE image_id" in q:  # S2
            self.rows = [(("img", schema, t["_img_index"][params[0]]),)]
        elif q.startswith("SELECT clip_embedding FROM") and ".text_chunks WHERE chunk_id" in q:  # S3
            self.rows = [(("chk", schema, t["_chk_index"][params[0]]),)]
        elif q.startswith("SELECT 1 - (%s::vector <=> %s::vector)"):  # S4
            a, b = (self._vec(p) for p in params)
            self.rows = [(oracle.cosine(a, b),)]
        elif q.startswith("SELECT chunk_id, 1 - (clip_embedding <=> %s::vector)"):  # S5
            emb, image_id, k = params
            t = self.db.tables[emb[1]]
            a = self._vec(emb)
            i = t["_img_index"][image_id]
            cand = self._same_page(t, i)
            sims = [oracle.cosine(t["chunk_emb"][j], a) for j in cand]  # t.clip_embedding <=> q
            order = sorted(range(len(cand)), key=lambda r: (-sims[r], cand[r]))[:k]
            self.rows = [(t["chunk_ids"][cand[r]], sims[r]) for r in order]
        elif q.startswith("SELECT alignment_type, weak_score FROM"):  # S6
            # weak_score column is REAL (src/setup_vector_db.py:141): fp32 on the way back
            self.rows = sorted(((ty, float(np.float32(sc))) for ty, sc in t.get("alignments", [])),
                               key=lambda r: r[0])
        else:
            raise AssertionError("unexpected SQL: " + q)

    def _vec(self, token):
        kind, schema, idx = token
        t = self.db.tables[schema]
        return t["image_emb"][idx] if kind == "img" else t["chunk_emb"][idx]


def _stub_modules(fake_connect):
    pg = types.ModuleType("psycopg2")
    pg.connect = fake_connect
    extras = types.ModuleType("psycopg2.extras")
    extras.execute_values = lambda *a, **k: None
    pg.extras = extras
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sns = types.ModuleType("seaborn")
    dotenv = types.ModuleType("dotenv")
    dotenv.load_dotenv = lambda *a, **k: None
    return {"psycopg2": pg, "psycopg2.extras": extras, "matplotlib": mpl,
            "matplotlib.pyplot": plt, "seaborn": sns, "dotenv": dotenv}


def load_reference(db: FakeDB | None = None, output_dir: Path | None = None):
    """Returns (evaluate_alignments, insert_clip_embeddings) reference modules."""
    if not available():
        raise RuntimeError("/root/reference is not present on this machine")
    stubs = _stub_modules(db.connect if db else (lambda **k: None))
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    orig_mkdir = Path.mkdir
    Path.mkdir = lambda self, *a, **k: None
    try:
        mods = []
        for name in ("evaluate_alignments", "insert_clip_embeddings"):
            spec = importlib.util.spec_from_file_location("_ref_" + name, REF / "src" / f"{name}.py")
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            mods.append(mod)
    finally:
        Path.mkdir = orig_mkdir
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    ev, ins = mods
    ev.psycopg2 = stubs["psycopg2"]
    if output_dir is not None:
        ev.OUTPUT_DIR = Path(output_dir)
    return ev, ins
