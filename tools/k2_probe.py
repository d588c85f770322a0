"""The exact rescoring (K2) on its own clock: one mmalign_run over N x M synthetic rows, the library's CUDA-event
times of the contraction and the rescoring.  Under ncu this is the command the K2 captures of profiles/ come from:
   ncu --set full --import-source on -k regex:"select_kernel|gather_kernel|rank_kernel" -o k2 python tools/k2_probe.py
"""
import argparse, importlib, pathlib, sys

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
PKG = "multimodal-alignment-of-noisy-image-text-pairs-using-weak-supervision_b200"
ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=8 * 148 * 128)
ap.add_argument("--M", type=int, default=1_000_000)
ap.add_argument("--D", type=int, default=512)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--schemas", default="vanilla_clip,clip_lexical,clip_positional,clip_combined")
a = ap.parse_args()
pkg = importlib.import_module(PKG)
synthetic = importlib.import_module(PKG + ".synthetic")
img, chk, _ = synthetic.make_torch(a.N, a.M, a.D, T=512, device="cuda")
eng = pkg.AlignmentEngine(0)
eng.set_images(img["emb"], img["key"], img["bbox"], None)
eng.set_chunks(chk["emb"], chk["key"], chk["bbox"], chk["terms"], n_terms=512)
for rep in range(a.reps):
    r = eng.run(a.schemas.split(","), candidates="all", k_values=(1, 5, 10, 20), mrr_cutoff=100, weak_weight=(0.3, 0.2),
                device_outputs=True)
    st = r["stats"]
    rows = a.N
    gathered = st["candidates_rescored"] * a.D * 4 if "candidates_rescored" in st else 0
    print(f"rep {rep}: fused {st['fused_us'] / 1e3:.2f} ms, rescore {st['rescore_us'] / 1e3:.3f} ms "
          f"({st['rescore_us'] / rows * 1e3:.1f} ns per row), rows rescanned {st['rows_rescanned']}, "
          f"hits {r['hits'][:, 2].tolist()}", flush=True)
eng.close()
