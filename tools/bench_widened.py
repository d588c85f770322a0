"""Measurement of the two kernels the path was widened by, on one B200:
   term_bitsets_kernel  ingest of the lexical term sets (src/insert_clip_embeddings.py:149-150)
   alignments_kernel    the `alignments` table producer (src/insert_clip_embeddings.py:369-414)
Each: device-resident inputs, CUDA events, best of 5 after a warm-up, algorithmic bytes / time against the measured
HBM copy bandwidth (MEASURED_PEAKS.json), and the CPU oracle timed on a bounded sample beside it.
   python tools/bench_widened.py [--chunks 400000] [--pairs-images 1000000]
Prints one JSON line per kernel."""
import argparse
import importlib
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
PKG = "multimodal-alignment-of-noisy-image-text-pairs-using-weak-supervision_b200"


def best_ms(fn, n=5):
    import torch
    fn()
    torch.cuda.synchronize()
    out = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1))
    return min(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunks", type=int, default=400_000)
    ap.add_argument("--terms", type=int, default=512)
    ap.add_argument("--pairs-images", type=int, default=1_000_000)
    ap.add_argument("--skip-alignments", action="store_true")
    a = ap.parse_args()
    import torch
    pkg = importlib.import_module(PKG)
    synthetic = importlib.import_module(PKG + ".synthetic")
    from oracle import oracle
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {"hbm_gbs": 6650.0}
    eng = pkg.AlignmentEngine(0)
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(0)

    # ---- ingest: word-like text, ~60 words of 2..9 letters per chunk, T terms drawn from the vocabulary
    letters = np.frombuffer(b"abcdefghijklmnopqrstuvwxyz", np.uint8)
    vocab = [bytes(rng.choice(letters, size=int(rng.integers(2, 10)))) + b" " for _ in range(4000)]
    vlen = np.array([len(v) for v in vocab])
    vcat = np.frombuffer(b"".join(vocab), np.uint8)
    vstart = np.concatenate([[0], np.cumsum(vlen)[:-1]])
    m = a.chunks
    words = rng.integers(20, 100, m)                       # words per chunk
    widx = rng.integers(0, len(vocab), int(words.sum()))   # word ids, all chunks
    wl = vlen[widx]
    total = int(wl.sum())
    pos = np.concatenate([[0], np.cumsum(wl)[:-1]])
    text = np.empty(total, np.uint8)
    src = np.repeat(vstart[widx] - pos, wl) + np.arange(total)   # gather index into vcat
    text[:] = vcat[src]
    off = np.zeros(m + 1, np.int64)
    np.cumsum(np.add.reduceat(wl, np.concatenate([[0], np.cumsum(words)[:-1]])), out=off[1:])
    terms = [vocab[i][:-1].decode() for i in rng.choice(len(vocab), a.terms, replace=False)]
    W = (a.terms + 63) // 64
    d_text, d_off = torch.from_numpy(text).to(dev), torch.from_numpy(off).to(dev)
    d_bits = torch.empty((m, W), dtype=torch.int64, device=dev)
    ms = best_ms(lambda: eng.term_bitsets_device(d_text, d_off, terms, d_bits))
    algo = total + 8 * (m + 1) + 8 * W * m
    # parity of the timed run against the oracle on a sample of chunks, and the oracle's own speed
    sample = np.sort(rng.choice(m, 2000, replace=False))
    texts_s = [bytes(text[off[j]:off[j + 1]]).decode() for j in sample]
    t0 = time.perf_counter()
    want = oracle.term_bitsets(texts_s, terms)
    dt = time.perf_counter() - t0
    assert np.array_equal(d_bits[torch.from_numpy(sample).to(dev)].cpu().numpy().view(np.uint64), want)
    print(json.dumps({"kernel": "term_bitsets_kernel", "chunks": m, "terms": a.terms, "text_bytes": total, "ms": ms,
                      "chunks_per_s": m / ms * 1e3,
                      "roofline": {"bound": "hbm", "achieved": algo / ms / 1e6, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                   "frac": algo / ms / 1e6 / peaks["hbm_gbs"],
                                   "note": "algorithmic bytes = text + offsets + term sets; the kernel is bounded by its byte comparisons"},
                      "cpu_baseline": {"value": len(sample) / dt, "unit": "chunks/s", "cores": 1, "kind": "port",
                                       "sample": f"{len(sample)} chunks x {a.terms} terms, oracle/mmalign_oracle.c: orc_term_bitsets"}}),
          flush=True)

    if a.skip_alignments:
        eng.close()
        return
    # ---- alignments: the records of all same-page pairs of the synthetic corpus (8 pairs per image)
    N = a.pairs_images
    img, chk, _ = synthetic.make_torch(N, N, 64, T=512, device=dev)
    eng.set_images(img["emb"], img["key"], img["bbox"], None)
    eng.set_chunks(chk["emb"], chk["key"], chk["bbox"], chk["terms"], n_terms=512)
    P = eng.num_pairs()
    rec = torch.empty((P, 3), dtype=torch.float64, device=dev)
    ms = best_ms(lambda: eng.alignments_device("clip_combined", rec))
    algo = P * (8 + 4 + 32 + 64 + 32 + 24)  # offsets probe + chunk id, image bbox, chunk term set, chunk bbox, record
    to_np = lambda d, n: {k: (v[:n].cpu().numpy() if v is not None else None) for k, v in d.items()}
    n_s = 2000  # the oracle restates the SQL join as a nested loop: O(sample x chunks)
    ih, ch = to_np(img, n_s), to_np(chk, N)
    for d in (ih, ch):
        d["key"] = d["key"].view(np.uint64)
        if d["terms"] is not None:
            d["terms"] = d["terms"].view(np.uint64)
    t0 = time.perf_counter()
    o_off, o_pc, o_rec = oracle.alignments(ih, ch, T=512, schema=3)
    dt = time.perf_counter() - t0
    assert np.array_equal(rec[:len(o_rec)].cpu().numpy(), o_rec)
    print(json.dumps({"kernel": "alignments_kernel", "images": N, "chunks": N, "pairs": P, "ms": ms, "pairs_per_s": P / ms * 1e3,
                      "roofline": {"bound": "hbm", "achieved": algo / ms / 1e6, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                   "frac": algo / ms / 1e6 / peaks["hbm_gbs"],
                                   "note": "algorithmic bytes per pair: pair index 12, image bbox 32, chunk terms 64, chunk bbox 32, record 24"},
                      "cpu_baseline": {"value": len(o_rec) / dt, "unit": "pairs/s", "cores": 1, "kind": "port",
                                       "sample": f"the pairs of the first {n_s} images (oracle/mmalign_oracle.c: orc_alignments, incl. the pair join)"}}),
          flush=True)
    eng.close()


if __name__ == "__main__":
    main()
