"""Runs the fused path at a size one ncu capture can replay (default 4 waves of row blocks x 262144 chunks), once
per kernel variant:  python tools/k1_probe.py [--N ..] [--M ..] [--D ..] [--variants 0,1] [--reps 2]
Prints the library's CUDA-event time of the fused kernel per variant (cta_pairs = 0 / 1)."""
import argparse
import importlib
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
PKG = "multimodal-alignment-of-noisy-image-text-pairs-using-weak-supervision_b200"
ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=4 * 148 * 128)
ap.add_argument("--M", type=int, default=262144)
ap.add_argument("--D", type=int, default=512)
ap.add_argument("--variants", default="0,1")
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
pkg = importlib.import_module(PKG)
synthetic = importlib.import_module(PKG + ".synthetic")
img, chk, _ = synthetic.make_torch(a.N, a.M, a.D, T=512, device="cuda")
for v in [int(x) for x in a.variants.split(",")]:
    eng = pkg.AlignmentEngine(0)
    eng.set_option("cta_pairs", v)
    eng.set_images(img["emb"], img["key"], img["bbox"], None)
    eng.set_chunks(chk["emb"], chk["key"], chk["bbox"], chk["terms"], n_terms=512)
    for rep in range(a.reps):
        r = eng.run(["vanilla_clip", "clip_lexical", "clip_positional", "clip_combined"], candidates="all",
                    k_values=(1, 5, 10, 20), mrr_cutoff=100, weak_weight=(0.3, 0.2), device_outputs=True)
        us = r["stats"]["fused_us"]
        print(f"cta_pairs={v} rep={rep}: fused {us / 1e3:.3f} ms = {2.0 * a.N * a.M * a.D / us / 1e6:.1f} TFLOP/s, "
              f"rescore {r['stats']['rescore_us'] / 1e3:.3f} ms, hits {r['hits'][0].tolist()}", flush=True)
    eng.close()
