"""Synthetic corpora of the shapes BASELINE.json names.

Embeddings are iid N(0,1) rows, L2-normalised; chunk j lies on page j // 8 of
manual page // 64 (8 chunks per page); image i lies on a uniformly drawn page and
is pulled towards one of that page's chunks (`signal`) so that Top-K / MRR are
not trivial; bboxes are boxes on a 612 x 792 pt page with 2 % all-zero
(invalid) entries; term sets are T Bernoulli(p) bits per chunk and, as in the
reference (src/insert_clip_embeddings.py:144-156 takes no image argument),
all-ones for images (terms=None) -- unless `img_p_term` > 0, which gives every
image its own Bernoulli(img_p_term) term set, the AND-popcount form of the lexical
term that BASELINE.json's north_star words (hits = popcount(chunk & image)).
"""
from __future__ import annotations

import numpy as np

CHUNKS_PER_PAGE = 8
PAGES_PER_MANUAL = 64


def page_keys(pages: np.ndarray) -> np.ndarray:
    pages = pages.astype(np.uint64)
    return ((pages // np.uint64(PAGES_PER_MANUAL)) << np.uint64(32)) | pages


def _bboxes(rng, n):
    x0 = rng.uniform(36, 400, n).astype(np.float32)
    y0 = rng.uniform(36, 600, n).astype(np.float32)
    w = rng.uniform(20, 176, n).astype(np.float32)
    h = rng.uniform(8, 156, n).astype(np.float32)
    b = np.stack([x0, y0, x0 + w, y0 + h], 1).astype(np.float64)
    b[rng.random(n) < 0.02] = 0.0
    return b


def _pack_bits(bits, T):
    M, W64 = bits.shape
    W = W64 // 64
    bits[:, T:] = False
    terms = np.packbits(bits.reshape(M, W, 8, 8)[:, :, ::-1, :], axis=-1, bitorder="little")
    return np.ascontiguousarray(terms[:, :, ::-1, 0]).view(np.uint64).reshape(M, W)


def make_numpy(N, M, D, *, T=512, p_term=0.01, signal=4.5, seed=0x5EED0000, img_p_term=0.0):
    """Small/medium corpora on the host (tests, golden vectors, CPU baseline)."""
    rng = np.random.default_rng(seed)
    ce = rng.standard_normal((M, D), dtype=np.float32)
    ce /= np.linalg.norm(ce, axis=1, keepdims=True)
    cpage = np.arange(M, dtype=np.int64) // CHUNKS_PER_PAGE
    n_pages = max(1, M // CHUNKS_PER_PAGE)
    ipage = rng.integers(0, n_pages, N)
    pick = np.minimum(ipage * CHUNKS_PER_PAGE + rng.integers(0, CHUNKS_PER_PAGE, N), M - 1)
    u = rng.standard_normal((N, D), dtype=np.float32)
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    ie = ce[pick] + np.float32(signal) * u
    ie /= np.linalg.norm(ie, axis=1, keepdims=True)
    W = (T + 63) // 64
    terms = _pack_bits(rng.random((M, W * 64)) < p_term, T)
    img = dict(emb=ie.astype(np.float32), key=page_keys(ipage), bbox=_bboxes(rng, N), terms=None)
    chk = dict(emb=ce, key=page_keys(cpage), bbox=_bboxes(rng, M), terms=terms)
    if img_p_term > 0:  # drawn last, so the rest of the corpus is the same with and without image term sets
        img["terms"] = _pack_bits(rng.random((N, W * 64)) < img_p_term, T)
    return img, chk, dict(T=T, planted=pick)


def make_torch(N, M, D, *, T=512, p_term=0.01, signal=4.5, seed=0x5EED0000, device="cuda",
               row0=0, rows=None, img_row0=0, img_rows=None, img_p_term=0.0):
    """Full-size corpora generated on the device, by GLOBAL row index in blocks of 65536 rows,
    so that any chunk shard [row0, row0+rows) and any image slab [img_row0, img_row0+img_rows) is
    identical whatever the world size."""
    import torch

    BLK = 65536
    rows = M if rows is None else rows

    def normal_rows(tag, lo, hi, dim):
        out = torch.empty((hi - lo, dim), dtype=torch.float32, device=device)
        b = lo // BLK
        while b * BLK < hi:
            g = torch.Generator(device=device)
            g.manual_seed((seed * 1000003 + tag * 7919 + b) & 0x7FFFFFFFFFFF)
            blk = torch.randn((BLK, dim), generator=g, dtype=torch.float32, device=device)
            s, e = max(lo, b * BLK), min(hi, (b + 1) * BLK)
            out[s - lo:e - lo] = blk[s - b * BLK:e - b * BLK]
            b += 1
        return out

    def unit(x):
        return x / x.norm(dim=1, keepdim=True)

    def uniform_rows(tag, lo, hi, dim):
        out = torch.empty((hi - lo, dim), dtype=torch.float64, device=device)
        b = lo // BLK
        while b * BLK < hi:
            g = torch.Generator(device=device)
            g.manual_seed((seed * 1000003 + tag * 7919 + b) & 0x7FFFFFFFFFFF)
            blk = torch.rand((BLK, dim), generator=g, dtype=torch.float64, device=device)
            s, e = max(lo, b * BLK), min(hi, (b + 1) * BLK)
            out[s - lo:e - lo] = blk[s - b * BLK:e - b * BLK]
            b += 1
        return out

    def bboxes(tag, lo, hi):
        r = uniform_rows(tag, lo, hi, 5)
        x0 = (36 + r[:, 0] * 364).float().double()
        y0 = (36 + r[:, 1] * 564).float().double()
        w = (20 + r[:, 2] * 156).float().double()
        h = (8 + r[:, 3] * 148).float().double()
        b = torch.stack([x0, y0, (x0 + w).float().double(), (y0 + h).float().double()], 1)
        b[r[:, 4] < 0.02] = 0.0
        return b.contiguous()

    def keys(pages):
        return ((pages // PAGES_PER_MANUAL) << 32) | pages  # int64 view of the u64 key

    # chunk shard
    ce = unit(normal_rows(1, row0, row0 + rows, D))
    cpage = torch.arange(row0, row0 + rows, device=device, dtype=torch.int64) // CHUNKS_PER_PAGE
    W = (T + 63) // 64
    weights = (1 << torch.arange(63, device=device, dtype=torch.int64))

    def term_rows(tag, lo, hi, p):
        tb = uniform_rows(tag, lo, hi, W * 64) < p
        tb[:, T:] = False
        tb = tb.view(hi - lo, W, 64)
        t = (tb[:, :, :63].long() * weights).sum(-1)
        return torch.where(tb[:, :, 63], t | torch.iinfo(torch.int64).min, t).contiguous()

    terms = term_rows(3, row0, row0 + rows, p_term)
    chk = dict(emb=ce, key=keys(cpage), bbox=bboxes(4, row0, row0 + rows), terms=terms)
    # images
    i0 = img_row0
    i1 = N if img_rows is None else img_row0 + img_rows
    N = i1 - i0
    n_pages = max(1, M // CHUNKS_PER_PAGE)
    r = uniform_rows(5, i0, i1, 2)
    ipage = (r[:, 0] * n_pages).long().clamp_(0, n_pages - 1)
    pick = (ipage * CHUNKS_PER_PAGE + (r[:, 1] * CHUNKS_PER_PAGE).long()).clamp_(0, M - 1)
    # the planted chunk row may live on another shard: regenerate it from its global index
    u = unit(normal_rows(2, i0, i1, D))
    ie = torch.empty((N, D), dtype=torch.float32, device=device)
    SL = 1 << 18
    for s in range(0, N, SL):
        pk = pick[s:s + SL]
        blocks = torch.unique(pk // BLK)
        src = torch.empty((len(pk), D), dtype=torch.float32, device=device)
        for b in blocks.tolist():
            rowsb = unit(normal_rows(1, b * BLK, min(M, (b + 1) * BLK), D))
            sel = (pk // BLK) == b
            src[sel] = rowsb[pk[sel] - b * BLK]
        ie[s:s + SL] = unit(src + signal * u[s:s + SL])
    img = dict(emb=ie, key=keys(ipage), bbox=bboxes(6, i0, i1),
               terms=term_rows(7, i0, i1, img_p_term) if img_p_term > 0 else None)
    return img, chk, dict(T=T, planted=pick)
