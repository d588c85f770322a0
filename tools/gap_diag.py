"""Diagnostic: distribution of (exact 100th best - 160th best cosine) per row versus the
certificate's epsilon, on the bench's synthetic data.  Run on the GPU box."""
import importlib, sys, torch
sys.path.insert(0, ".")
PKG = "multimodal-alignment-of-noisy-image-text-pairs-using-weak-supervision_b200"
syn = importlib.import_module(PKG + ".synthetic")
N, M, D = 1_000_000, 1_000_000, 512
img, chk, meta = syn.make_torch(N, M, D, device="cuda")
rows = torch.arange(0, N, 500, device="cuda")
a, b = img["emb"][rows], chk["emb"]
top = torch.empty((len(rows), 200), device="cuda")
for s in range(0, len(rows), 250):
    sc = a[s:s + 250] @ b.T
    same = (img["key"][rows[s:s + 250]][:, None] == chk["key"][None, :])
    sc[same] = -1.0
    top[s:s + 250] = sc.topk(200, dim=1).values
gap = top[:, 99] - top[:, 159]
ea = (a - a.bfloat16().float()).norm(dim=1)
eb = (b - b.bfloat16().float()).norm(dim=1)
eps = ea * 1.001 + eb.max() * 1.001 + D * 2.4e-7 + 2e-6
print("rows", len(rows), "gap mean %.5f sd %.5f min %.5f" % (gap.mean(), gap.std(), gap.min()))
print("s100 mean %.4f s160 mean %.4f" % (top[:, 99].mean(), top[:, 159].mean()))
print("eps mean %.5f max %.5f ; ea mean %.5f eb max %.5f eb mean %.5f" % (eps.mean(), eps.max(), ea.mean(), eb.max(), eb.mean()))
print("fraction gap < eps: %.5f" % (gap < eps).float().mean())
for k in (128, 160, 192, 224, 256):
    g = top[:, 99] - top[:, k - 1] if k <= 200 else None
    if g is not None:
        print(k, "frac fail %.5f" % (g < eps).float().mean())
bf = (a.bfloat16().float() @ b[:200000].bfloat16().float().T) - (a @ b[:200000].T)
print("bf16 score error: rms %.2e max %.2e" % (bf.pow(2).mean().sqrt(), bf.abs().max()))
