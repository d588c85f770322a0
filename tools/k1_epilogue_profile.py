"""Where the fused kernel's epilogue warps spend their cycles (profiling build: make -C .../csrc prof).
   MMALIGN_LIB=<...>/csrc/libmmalign_prof.so python tools/k1_epilogue_profile.py [--N ..] [--M ..] [--D ..]
Prints, per epilogue warp and tile: cycles waiting for the tensor pipe, waiting for TMEM loads, filtering, compacting."""
import argparse
import ctypes as C
import importlib
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
PKG = "multimodal-alignment-of-noisy-image-text-pairs-using-weak-supervision_b200"
ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=8 * 148 * 128)
ap.add_argument("--M", type=int, default=1_000_000)
ap.add_argument("--D", type=int, default=512)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--compact", type=int, default=-1)
ap.add_argument("--pairs", type=int, default=0, help="cta_pairs option of the engine (0, 1 = cta_group::2, 2 = B multicast)")
a = ap.parse_args()
os.environ.setdefault("MMALIGN_LIB", str(ROOT / PKG / "csrc" / "libmmalign_prof.so"))
pkg = importlib.import_module(PKG)
synthetic = importlib.import_module(PKG + ".synthetic")
L = pkg._native.load()
img, chk, _ = synthetic.make_torch(a.N, a.M, a.D, T=512, device="cuda")
eng = pkg.AlignmentEngine(0)
eng.set_option("cta_pairs", a.pairs)
if a.compact >= 0:
    eng.set_option("compact_one", a.compact)
eng.set_images(img["emb"], img["key"], img["bbox"], None)
eng.set_chunks(chk["emb"], chk["key"], chk["bbox"], chk["terms"], n_terms=512)
buf = (C.c_ulonglong * 16)()
k2 = (C.c_ulonglong * 16)()
for rep in range(a.reps):
    L.mmalign_profile_counters(buf, 1)
    L.mmalign_profile_k2(k2, 1)
    r = eng.run(["vanilla_clip", "clip_lexical", "clip_positional", "clip_combined"], candidates="all",
                k_values=(1, 5, 10, 20), mrr_cutoff=100, weak_weight=(0.3, 0.2), device_outputs=True)
    L.mmalign_profile_counters(buf, 0)
    L.mmalign_profile_k2(k2, 0)
    w = [int(x) for x in k2]
    rows_ = max(w[8], 1)
    names = ["stage row", "same-page entries", "list sweep", "sort approx", "theta", "candidate gathers + cosine", "sort exact",
             "merge + outputs"]
    print(f"rep {rep}: rescore {r['stats']['rescore_us'] / 1e3:.2f} ms; cycles per row (one CTA): " +
          ", ".join(f"{n} {w[q] / (max(w[12], 1) if q >= 5 and w[12] else rows_):.0f}" for q, n in enumerate(names)) + f"; total {sum(w[:8]) / rows_:.0f}; sampled rows {w[8]}, handed to the block kernel {w[9]}, union entries per row {w[10] / rows_:.1f}, re-scored {w[11] / rows_:.1f}", flush=True)
    v = [int(x) for x in buf]
    tiles, warps = max(v[4], 1), max(v[7], 1)
    us = r["stats"]["fused_us"]
    print(f"rep {rep}: fused {us / 1e3:.2f} ms = {2.0 * a.N * a.M * a.D / us / 1e6:.0f} TFLOP/s; per epilogue warp and tile (cycles): "
          f"wait-for-MMA {v[0] / tiles:.0f}, wait-for-TMEM-load {v[1] / tiles:.0f}, filter {v[2] / tiles:.0f}, "
          f"routine compaction {v[3] / tiles:.0f} ({v[5] / tiles:.3f} per tile), total {v[6] / tiles:.0f}; "
          f"warps {warps}, tiles/warp {tiles / warps:.0f}; per 32x32 chunk: maxima+vote {v[8] / tiles / 4:.0f} cycles, "
          f"P(chunk has a hit) {v[9] / tiles / 4:.3f}, 8-column groups with a hit per chunk {v[10] / tiles / 4:.3f}, "
          f"hit path {v[11] / max(v[9], 1):.0f} cycles per chunk with a hit = {v[11] / max(v[10], 1):.0f} per group; "
          f"per compacted list (all call sites): load {v[12] / max(v[15], 1):.0f}, bisection {v[13] / max(v[15], 1):.0f}, "
          f"write-back {v[14] / max(v[15], 1):.0f} cycles; lists {v[15]}", flush=True)
eng.close()
