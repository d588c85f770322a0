"""ctypes front end of the CPU oracle (oracle/mmalign_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package never
imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None

NULL_KEY = np.uint64(0xFFFFFFFFFFFFFFFF)
SCHEMAS = ["vanilla_clip", "clip_lexical", "clip_positional", "clip_combined"]


def build(force: bool = False) -> Path:
    so = _HERE / "liboracle.so"
    src = _HERE / "mmalign_oracle.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B", "liboracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(str(build()))
        fp, dp, u64p, i64p, i32p = (C.POINTER(C.c_float), C.POINTER(C.c_double),
                                    C.POINTER(C.c_uint64), C.POINTER(C.c_int64),
                                    C.POINTER(C.c_int32))
        L.orc_dot.restype = C.c_float
        L.orc_dot.argtypes = [fp, fp, C.c_int]
        L.orc_dot_sequential.restype = C.c_float
        L.orc_dot_sequential.argtypes = [fp, fp, C.c_int]
        L.orc_cosine.restype = C.c_double
        L.orc_cosine.argtypes = [fp, fp, C.c_int]
        L.orc_cosine_sequential.restype = C.c_double
        L.orc_cosine_sequential.argtypes = [fp, fp, C.c_int]
        L.orc_lexical.restype = C.c_double
        L.orc_lexical.argtypes = [C.c_int64, C.c_int64]
        L.orc_positional.restype = C.c_double
        L.orc_positional.argtypes = [dp, dp]
        L.orc_weak_records.restype = C.c_int
        L.orc_weak_records.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, dp]
        L.orc_pairs.restype = C.c_int64
        L.orc_pairs.argtypes = [u64p, C.c_int64, u64p, C.c_int64, i64p, i64p]
        L.orc_eval.restype = C.c_int
        L.orc_eval.argtypes = [fp, u64p, dp, u64p, C.c_int64, fp, u64p, dp, u64p, C.c_int64,
                               C.c_int, C.c_int, C.c_int64, C.c_uint32, C.c_int,
                               C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
                               i64p, i64p, i64p, dp, i32p, dp, C.c_int]
        L.orc_alignments.restype = C.c_int
        L.orc_alignments.argtypes = [dp, u64p, C.c_int64, dp, u64p, C.c_int, C.c_int64, C.c_int,
                                     i64p, i64p, dp]
        L.orc_term_bitsets.restype = C.c_int
        L.orc_term_bitsets.argtypes = [C.c_void_p, i64p, C.c_int64, C.c_void_p, i64p, C.c_int, C.c_int, u64p]
        _LIB = L
    return _LIB


def _p(a, ct):
    return None if a is None else a.ctypes.data_as(C.POINTER(ct))


def _c(a, dt):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)


def dot(a, b):
    a, b = _c(a, np.float32), _c(b, np.float32)
    return float(lib().orc_dot(_p(a, C.c_float), _p(b, C.c_float), a.shape[-1]))


def cosine(a, b, sequential=False):
    a, b = _c(a, np.float32), _c(b, np.float32)
    f = lib().orc_cosine_sequential if sequential else lib().orc_cosine
    return float(f(_p(a, C.c_float), _p(b, C.c_float), a.shape[-1]))


def lexical(hits: int, T: int) -> float:
    return float(lib().orc_lexical(int(hits), int(T)))


def positional(img_bbox, chk_bbox) -> float:
    """bbox = 4 floats, or None / wrong length (mapped to zeros, see the C header)."""
    def norm(b):
        if b is None or len(b) != 4:
            return np.zeros(4, np.float64)
        return np.asarray(b, np.float64)
    a, b = norm(img_bbox), norm(chk_bbox)
    return float(lib().orc_positional(_p(a, C.c_double), _p(b, C.c_double)))


def weak_records(use_lex, use_pos, lex, pos):
    out = np.zeros(3, np.float64)
    m = lib().orc_weak_records(int(use_lex), int(use_pos), float(lex), float(pos), _p(out, C.c_double))
    return m, out


def pairs(img_key, chk_key):
    img_key, chk_key = _c(img_key, np.uint64), _c(chk_key, np.uint64)
    N, M = len(img_key), len(chk_key)
    off = np.zeros(N + 1, np.int64)
    P = lib().orc_pairs(_p(img_key, C.c_uint64), N, _p(chk_key, C.c_uint64), M, _p(off, C.c_int64), None)
    pc = np.zeros(max(P, 1), np.int64)
    lib().orc_pairs(_p(img_key, C.c_uint64), N, _p(chk_key, C.c_uint64), M, _p(off, C.c_int64),
                    _p(pc, C.c_int64))
    return off, pc[:P]


def evaluate(img, chk, *, T=0, schema_mask=1, candidates="same_page", lam=(0.0, 0.0, 0.0),
             kmax=10, cutoff=100, nthreads=0):
    """img / chk: dicts with emb [n,D] f32, key [n] u64, bbox [n,4] f64, terms [n,W] u64 or None.

    Returns dict(topk_idx [S,N,kmax], topk_score, pair_offsets [N+1], pair_chunk [P],
    pair_rank [S,P], pair_sim [P]).
    """
    ie, ce = _c(img["emb"], np.float32), _c(chk["emb"], np.float32)
    N, D = ie.shape
    M = ce.shape[0]
    ik, ck = _c(img["key"], np.uint64), _c(chk["key"], np.uint64)
    ib = _c(img.get("bbox") if img.get("bbox") is not None else np.zeros((N, 4)), np.float64)
    cb = _c(chk.get("bbox") if chk.get("bbox") is not None else np.zeros((M, 4)), np.float64)
    ct = _c(chk.get("terms"), np.uint64)
    it = _c(img.get("terms"), np.uint64)
    W = 0 if ct is None else ct.shape[1]
    if ct is None:
        ct = np.zeros((M, 1), np.uint64)
        W = 0
    off, pc = pairs(ik, ck)
    P = len(pc)
    S = bin(schema_mask & 15).count("1")
    ti = np.full((S, N, kmax), -1, np.int64)
    ts = np.full((S, N, kmax), -np.inf, np.float64)
    pr = np.zeros((S, max(P, 1)), np.int32)
    ps = np.zeros(max(P, 1), np.float64)
    pcb = _c(pc if P else np.zeros(1, np.int64), np.int64)
    cand = {"same_page": 0, "all": 1}[candidates] if isinstance(candidates, str) else int(candidates)
    lib().orc_eval(_p(ie, C.c_float), _p(ik, C.c_uint64), _p(ib, C.c_double), _p(it, C.c_uint64), N,
                   _p(ce, C.c_float), _p(ck, C.c_uint64), _p(cb, C.c_double), _p(ct, C.c_uint64), M,
                   D, W, int(T), schema_mask, cand, lam[0], lam[1], lam[2], kmax, cutoff,
                   _p(off, C.c_int64), _p(pcb, C.c_int64), _p(ti, C.c_int64), _p(ts, C.c_double),
                   _p(pr, C.c_int32), _p(ps, C.c_double), nthreads)
    return dict(topk_idx=ti, topk_score=ts, pair_offsets=off, pair_chunk=pc,
                pair_rank=pr[:, :P], pair_sim=ps[:P])


def alignments(img, chk, *, T, schema: int):
    """Records of the `alignments` table for one schema: rec [P,3] (lexical, positional, combined)."""
    N, M = len(img["key"]), len(chk["key"])
    ib = _c(img.get("bbox") if img.get("bbox") is not None else np.zeros((N, 4)), np.float64)
    cb = _c(chk.get("bbox") if chk.get("bbox") is not None else np.zeros((M, 4)), np.float64)
    ct = _c(chk.get("terms"), np.uint64)
    it = _c(img.get("terms"), np.uint64)
    W = 0 if ct is None else ct.shape[1]
    if ct is None:
        ct = np.zeros((M, 1), np.uint64)
    off, pc = pairs(img["key"], chk["key"])
    P = len(pc)
    rec = np.zeros((max(P, 1), 3), np.float64)
    pcb = _c(pc if P else np.zeros(1, np.int64), np.int64)
    lib().orc_alignments(_p(ib, C.c_double), _p(it, C.c_uint64), N, _p(cb, C.c_double),
                         _p(ct, C.c_uint64), W, int(T), schema, _p(off, C.c_int64),
                         _p(pcb, C.c_int64), _p(rec, C.c_double))
    return off, pc, rec[:P]


def pack_strings(strings):
    """UTF-8 bytes of `strings`, concatenated, with the [n+1] int64 offsets."""
    enc = [s.encode("utf-8") if isinstance(s, str) else bytes(s) for s in strings]
    off = np.zeros(len(enc) + 1, np.int64)
    np.cumsum([len(b) for b in enc], out=off[1:])
    return np.frombuffer(b"".join(enc) or b"\0", np.uint8).copy(), off


def term_bitsets(texts, terms, term_words=None):
    """bits [m, W] u64: bit t of row j = terms[t] in texts[j].lower()  (insert_clip_embeddings.py:149-150)."""
    tb, to = pack_strings([t.lower() for t in texts])
    pb, po = pack_strings(terms)
    W = term_words or max(1, (len(terms) + 63) // 64)
    bits = np.zeros((len(texts), W), np.uint64)
    rc = lib().orc_term_bitsets(tb.ctypes.data, _p(to, C.c_int64), len(texts), pb.ctypes.data, _p(po, C.c_int64),
                                len(terms), W, _p(bits, C.c_uint64))
    assert rc == 0
    return bits


def metrics_from_ranks(pair_rank, pair_sim, k_values=(1, 5, 10), mrr_cutoff=100):
    """evaluate_alignments.py:182-192, :203-216, :226-231 on the per-pair arrays."""
    P = len(pair_sim)
    if P == 0:
        return {"top_k": {k: 0.0 for k in k_values}, "mrr": 0.0, "avg_similarity": 0.0, "num_pairs": 0}
    top_k = {k: int(np.count_nonzero((pair_rank >= 1) & (pair_rank <= k))) / P for k in k_values}
    rr = [1.0 / r if 1 <= r <= mrr_cutoff else 0.0 for r in pair_rank.tolist()]
    return {"top_k": top_k, "mrr": np.mean(rr), "avg_similarity": np.mean(pair_sim.tolist()),
            "num_pairs": P}
