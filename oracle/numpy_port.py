"""numpy/BLAS port of the scoring + ranking path: the CPU baseline of bench.py.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (see oracle/mmalign_oracle.c).  Follows the same
reference sites as the C oracle but uses what a CPU implementation would really use for
speed: float32 `A @ B.T` through the BLAS numpy links against (all host cores), in column
slabs; selection of each slab's best candidates in float32 with `np.argpartition`, row blocks
spread over a thread pool (numpy releases the GIL there); float64 cosine, clamp and ordering
(`np.lexsort`, lower index wins ties) only for the kept candidates; the chunk table's norms and
sorted page keys computed once per table; vectorised weak-supervision terms
(src/insert_clip_embeddings.py:144-210, :369-414).
BLAS sums in its own order, so scores agree with the C oracle to ~1e-7, not bit for bit;
tests/test_numpy_port.py pins it against the oracle.
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_CHUNK_CACHE = {}
_POOL = None


def _pool():
    global _POOL
    if _POOL is None:
        _POOL = ThreadPoolExecutor(max_workers=os.cpu_count() or 1)
    return _POOL


def _chunk_side(chk):
    """Per chunk table, once: squared norms and the page keys in sorted order (the join of evaluate_alignments.py:57-63)."""
    B = chk["emb"]
    tag = (id(B), B.shape, id(chk["key"]))
    hit = _CHUNK_CACHE.get(tag)
    if hit is None:
        Bc = np.ascontiguousarray(B, np.float32)
        nb = np.einsum("ij,ij->i", Bc, Bc).astype(np.float64)
        ck = np.asarray(chk["key"], np.uint64)
        order = np.argsort(ck, kind="stable")
        _CHUNK_CACHE.clear()
        hit = _CHUNK_CACHE[tag] = (Bc, nb, order, ck[order])
    return hit

_LEX = (False, True, False, True)
_POS = (False, False, True, True)


def _positional(ib, cb):
    """Vectorised src/insert_clip_embeddings.py:159-210 for paired rows ib, cb [P,4] (f64)."""
    bad = ((ib[:, 2] - ib[:, 0] == 0) | (ib[:, 3] - ib[:, 1] == 0) |
           (cb[:, 2] - cb[:, 0] == 0) | (cb[:, 3] - cb[:, 1] == 0))
    x1, y1 = np.maximum(ib[:, 0], cb[:, 0]), np.maximum(ib[:, 1], cb[:, 1])
    x2, y2 = np.minimum(ib[:, 2], cb[:, 2]), np.minimum(ib[:, 3], cb[:, 3])
    apart = (x2 <= x1) | (y2 <= y1)
    dx = (ib[:, 0] + ib[:, 2]) / 2 - (cb[:, 0] + cb[:, 2]) / 2
    dy = (ib[:, 1] + ib[:, 3]) / 2 - (cb[:, 1] + cb[:, 3]) / 2
    dist_score = np.maximum(0.0, 1.0 - np.sqrt(dx * dx + dy * dy) / 1000.0)
    inter = np.maximum(0, x2 - x1) * np.maximum(0, y2 - y1)
    uni = (ib[:, 2] - ib[:, 0]) * (ib[:, 3] - ib[:, 1]) + (cb[:, 2] - cb[:, 0]) * (cb[:, 3] - cb[:, 1]) - inter
    with np.errstate(divide="ignore", invalid="ignore"):
        iou = np.where(uni == 0, 0.0, inter / uni)
    return np.where(bad, 0.0, np.where(apart, dist_score, iou))


def _weak(schema, lex, pos, lam):
    ul, up = _LEX[schema], _POS[schema]
    have_l = ul & (lex > 0.05)
    have_p = up & (pos > 0.05)
    if ul and up:
        both = have_l & have_p
        comb = (lex + pos) / 2
        return np.where(both, np.where(comb > 0.1, lam[2] * comb, 0.0),
                        lam[0] * np.where(have_l, lex, 0.0) + lam[1] * np.where(have_p, pos, 0.0))
    return lam[0] * np.where(have_l, lex, 0.0) + lam[1] * np.where(have_p, pos, 0.0)


def evaluate(img, chk, *, T=0, schemas=(0,), lam=(0.0, 0.0, 0.0), kmax=10, cutoff=100, rows=None, slab=1 << 16):
    """Full N x M ranking ("all" candidates) of the image rows `rows` (default: all).

    Returns dict(topk_idx [S,R,kmax], topk_score, pair_rows [P] (index into rows), pair_chunk [P],
    pair_rank [S,P], pair_sim [P])."""
    A = np.ascontiguousarray(img["emb"], np.float32)
    B, nb, order, sk = _chunk_side(chk)
    rows = np.arange(A.shape[0]) if rows is None else np.asarray(rows)
    A = A[rows]
    R, M = A.shape[0], B.shape[0]
    kneed = max(kmax, cutoff)
    na = np.einsum("ij,ij->i", A, A).astype(np.float64)
    # ---- same-page pairs (evaluate_alignments.py:57-63) via the sorted chunk keys
    ik = np.asarray(img["key"], np.uint64)[rows]
    lo, hi = np.searchsorted(sk, ik, "left"), np.searchsorted(sk, ik, "right")
    hi = np.where(ik == np.uint64(0xFFFFFFFFFFFFFFFF), lo, hi)
    cnt = hi - lo
    off = np.concatenate([[0], np.cumsum(cnt)])
    P = int(off[-1])
    pair_rows = np.repeat(np.arange(R), cnt)
    pair_chunk = order[np.concatenate([np.arange(l, h) for l, h in zip(lo, hi)])] if P else np.zeros(0, np.int64)
    # ---- cosine of every (row, chunk) in slabs; running top-(kneed + max page size)
    # selection in float32 (16 entries of margin for its rounding); the kept candidates get the float64 cosine below
    keep = min(M, kneed + int(cnt.max(initial=0)) + 16)
    best_a = np.full((R, 0), -np.inf, np.float32)   # approximate cosine (selection only)
    best_d = np.zeros((R, 0), np.float32)           # the raw float32 dot product
    best_j = np.zeros((R, 0), np.int64)
    ra = (1.0 / np.sqrt(na)).astype(np.float32)
    rb = (1.0 / np.sqrt(nb)).astype(np.float32)
    blocks = [(r0, min(R, r0 + max(1, -(-R // (os.cpu_count() or 1))))) for r0 in
              range(0, R, max(1, -(-R // (os.cpu_count() or 1))))]

    def top_of(x, k):  # indices of the k largest per row, row blocks on the thread pool
        w = x.shape[1]
        if k >= w:
            return np.tile(np.arange(w), (x.shape[0], 1))
        out = np.empty((x.shape[0], k), np.int64)

        def one(b):
            out[b[0]:b[1]] = np.argpartition(x[b[0]:b[1]], w - k, axis=1)[:, w - k:]
        list(_pool().map(one, blocks))
        return out
    for c0 in range(0, M, slab):
        c1 = min(M, c0 + slab)
        dots = A @ B[c0:c1].T                              # float32 sgemm, all host cores
        approx = dots * ra[:, None] * rb[None, c0:c1]
        part = top_of(approx, min(keep, c1 - c0))
        best_a = np.concatenate([best_a, np.take_along_axis(approx, part, 1)], 1)
        best_d = np.concatenate([best_d, np.take_along_axis(dots, part, 1)], 1)
        best_j = np.concatenate([best_j, part + c0], 1)
        if best_a.shape[1] > keep:
            sel = top_of(best_a, keep)
            best_a, best_d, best_j = (np.take_along_axis(v, sel, 1) for v in (best_a, best_d, best_j))
    best_s = best_d.astype(np.float64) / np.sqrt(na[:, None] * nb[best_j])
    np.clip(best_s, -1.0, 1.0, out=best_s)
    best_s = 1.0 - (1.0 - best_s)
    # ---- exact pair similarities and weak terms
    if P:
        d = np.einsum("ij,ij->i", A[pair_rows], B[pair_chunk]).astype(np.float64)
        pair_sim = 1.0 - (1.0 - np.clip(d / np.sqrt(na[pair_rows] * nb[pair_chunk]), -1.0, 1.0))
        terms = chk.get("terms")
        if terms is not None and T > 0:
            t = np.asarray(terms, np.uint64)[pair_chunk]
            if img.get("terms") is not None:
                t = t & np.asarray(img["terms"], np.uint64)[rows][pair_rows]
            hits = np.bitwise_count(t).sum(1).astype(np.float64)
            lex = np.minimum(1.0, hits / max(T * 0.1, 1))
        else:
            lex = np.zeros(P)
        pos = _positional(np.asarray(img["bbox"], np.float64)[rows][pair_rows],
                          np.asarray(chk["bbox"], np.float64)[pair_chunk])
    else:
        pair_sim = np.zeros(0)
    S_ = len(schemas)
    topk_idx = np.full((S_, R, kmax), -1, np.int64)
    topk_score = np.full((S_, R, kmax), -np.inf)
    pair_rank = np.zeros((S_, P), np.int32)
    for si, s in enumerate(schemas):
        w = _weak(s, lex, pos, lam) if P else np.zeros(0)
        for r in range(R):
            j, sc = best_j[r], best_s[r].copy()
            p0, p1 = off[r], off[r + 1]
            if p1 > p0:
                pj = pair_chunk[p0:p1]
                own = np.isin(j, pj)
                j = np.concatenate([j[~own], pj])
                sc = np.concatenate([sc[~own], pair_sim[p0:p1] + w[p0:p1]])
            o = np.lexsort((j, -sc))[:kneed]
            jj, ss = j[o], sc[o]
            n = min(kmax, len(o))
            topk_idx[si, r, :n], topk_score[si, r, :n] = jj[:n], ss[:n]
            if p1 > p0:
                pos_in = {int(c): q + 1 for q, c in enumerate(jj[:cutoff])}
                pair_rank[si, p0:p1] = [pos_in.get(int(c), 0) for c in pair_chunk[p0:p1]]
    return dict(topk_idx=topk_idx, topk_score=topk_score, pair_rows=pair_rows, pair_chunk=pair_chunk,
                pair_rank=pair_rank, pair_sim=pair_sim)


def blas_info():
    try:
        from threadpoolctl import threadpool_info
        info = [i for i in threadpool_info() if i.get("user_api") == "blas"]
        if info:
            return f"{info[0].get('internal_api')} {info[0].get('version')} threads={info[0].get('num_threads')}"
    except Exception:
        pass
    return "unknown BLAS"
