"""Wall time of the UNMODIFIED reference metric functions (compute_top_k_accuracy, compute_mrr,
compute_average_similarity of /root/reference/src/evaluate_alignments.py) on BASELINE config 1
(1k images x 5k chunks, D=512, vanilla_clip, same-page candidates) over the in-process fake pgvector of
oracle/reference_harness.py.  Runs only where /root/reference exists (the build container); SURVEY section 8d (i).
   python tools/time_reference_config1.py"""
import os
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "golden"))
import importlib  # noqa: E402

from oracle import reference_harness as rh  # noqa: E402

synthetic = importlib.import_module("multimodal-alignment-of-noisy-image-text-pairs-using-weak-supervision_b200.synthetic")

N, M, D = 1000, 5000, 512
img, chk, _ = synthetic.make_numpy(N, M, D, seed=0x5EED0001)
man = lambda k: f"man{int(k) >> 32}"
page = lambda k: int(int(k) & 0xFFFFFFFF)
t = dict(image_ids=[f"img{i}" for i in range(N)], image_manual=[man(k) for k in img["key"]],
         image_page=[page(k) for k in img["key"]], image_emb=img["emb"],
         chunk_ids=[f"chk{j}" for j in range(M)], chunk_manual=[man(k) for k in chk["key"]],
         chunk_page=[page(k) for k in chk["key"]], chunk_emb=chk["emb"], alignments=[])
db = rh.FakeDB({"vanilla_clip": t})
index = {}
for j in range(M):
    index.setdefault((t["chunk_manual"][j], t["chunk_page"][j]), []).append(j)
rh._Cursor._same_page = lambda self, tt, i: index.get((tt["image_manual"][i], tt["image_page"][i]), [])
ev, _ = rh.load_reference(db, output_dir=Path(tempfile.mkdtemp()))
s = "vanilla_clip"
t0 = time.perf_counter()
acc = ev.compute_top_k_accuracy(s, [1, 5, 10])
mrr = ev.compute_mrr(s)
avg = ev.compute_average_similarity(s)
dt = time.perf_counter() - t0
P = len(ev.get_image_text_pairs(s))
import json  # noqa: E402
rec = {"value": N / dt, "unit": "queries/s", "seconds": dt, "images": N, "pairs": P, "cores": 1, "host_vcpus": os.cpu_count(),
       "fake_db_connections": db.n_connect, "kind": "reference",
       "what": "UNMODIFIED compute_top_k_accuracy + compute_mrr + compute_average_similarity of src/evaluate_alignments.py:169-231 "
               "at BASELINE config 1 (1k x 5k x 512, vanilla_clip) over the in-process fake pgvector of oracle/reference_harness.py "
               "(single Python thread; against a real PostgreSQL every one of the connections is a TCP round trip)",
       "where": "build container (the reference tree does not travel to the GPU box)",
       "metrics": {"top_k": {str(k): v for k, v in acc.items()}, "mrr": float(mrr), "avg_similarity": float(avg)}}
(ROOT / "profiles" / "bench" / "r2_reference_as_written_config1.json").write_text(json.dumps(rec, indent=1) + "\n")
print(f"reference metric functions, config 1: {dt:.2f} s for {N} images / {P} pairs on {os.cpu_count()} vCPU "
      f"(single Python thread) = {N / dt:.1f} images/s, {db.n_connect} fake-DB connections; top_k={acc} mrr={mrr:.6f} avg={avg:.6f}")
