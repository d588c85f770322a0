import importlib, sys, time, torch
sys.path.insert(0, ".")
PKG = "multimodal-alignment-of-noisy-image-text-pairs-using-weak-supervision_b200"
pkg = importlib.import_module(PKG); syn = importlib.import_module(PKG + ".synthetic")
N, M, D = int(sys.argv[1]), int(sys.argv[2]), 512
img, chk, _ = syn.make_torch(N, M, D, device="cuda")
eng = pkg.AlignmentEngine(0)
def t(label, f):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize()
    print(f"{label:28s} {1e3 * (time.perf_counter() - t0):9.2f} ms"); return r
for it in range(3):
    print("--- iteration", it)
    t("set_images", lambda: eng.set_images(img["emb"], img["key"], img["bbox"], None))
    t("set_chunks", lambda: eng.set_chunks(chk["emb"], chk["key"], chk["bbox"], chk["terms"], n_terms=512))
    t("num_pairs", lambda: eng.num_pairs())
    t("torch.empty outputs", lambda: [torch.empty((4, N, 20), dtype=torch.int64, device="cuda"), torch.empty((4, N, 20), dtype=torch.float64, device="cuda")])
    r = t("run(device_outputs)", lambda: eng.run(pkg.SCHEMAS, candidates="all", k_values=(1, 5, 10, 20), weak_weight=(0.3, 0.2), device_outputs=True))
    print("   stats", r["stats"])
