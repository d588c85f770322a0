"""What paces the fused kernel (K1)?  The tuning build of the library (csrc: `make tune`) can take the kernel apart:
MMALIGN_K1_DIAG=1 hands every accumulator back unread (no epilogue work), =2 stops the operand loads after the ring's
first fill (the MMAs re-read shared memory), =3 both (the tensor pipe and its barriers alone).  Each variant runs
K1 alone (mmalign_fused_pass) back to back for about a second, so that the power cap is in the number.
   MMALIGN_LIB=.../csrc/libmmalign_tune.so python tools/k1_diag.py [--N rows] [--M cols] [--D dim]"""
import argparse, importlib, os, pathlib, sys, time

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
PKG = "multimodal-alignment-of-noisy-image-text-pairs-using-weak-supervision_b200"
os.environ.setdefault("MMALIGN_LIB", str(ROOT / PKG / "csrc" / "libmmalign_tune.so"))
ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=16 * 148 * 128)
ap.add_argument("--M", type=int, default=1_000_000)
ap.add_argument("--D", type=int, default=512)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--variants", default="0,1,2,3")
ap.add_argument("--pairs", type=int, default=0)
ap.add_argument("--compact", type=int, default=-1, help="compact_one option (-1 = library default)")
ap.add_argument("--epi-sleep", type=int, default=-1, help="epi_sleep_ns option (-1 = library default)")
a = ap.parse_args()
import torch
pkg = importlib.import_module(PKG)
synthetic = importlib.import_module(PKG + ".synthetic")
img, chk, _ = synthetic.make_torch(a.N, a.M, a.D, T=512, device="cuda")
eng = pkg.AlignmentEngine(0)
eng.set_option("cta_pairs", a.pairs)
if a.compact >= 0:
    eng.set_option("compact_one", a.compact)
if a.epi_sleep >= 0:
    eng.set_option("epi_sleep_ns", a.epi_sleep)
eng.set_images(img["emb"], img["key"], img["bbox"], None)
eng.set_chunks(chk["emb"], chk["key"], chk["bbox"], chk["terms"], n_terms=512)
import threading
import pynvml
pynvml.nvmlInit()
_h = pynvml.nvmlDeviceGetHandleByIndex(0)


class Sampler(threading.Thread):  # SM clock and board power while a variant runs (NVML, every 20 ms)
    def __init__(self):
        super().__init__(daemon=True)
        self.stop, self.mhz, self.watts = False, [], []

    def run(self):
        while not self.stop:
            self.mhz.append(pynvml.nvmlDeviceGetClockInfo(_h, pynvml.NVML_CLOCK_SM))
            self.watts.append(pynvml.nvmlDeviceGetPowerUsage(_h) / 1000.0)
            time.sleep(0.02)


names = {0: "whole kernel", 1: "no epilogue work", 2: "no operand loads", 3: "neither (tensor pipe + barriers)",
         4: "accumulator handed back right after its loads (pair kernel; lists not valid)",
         32: "TMEM loads only (no filter, no lists)", 8: "no epilogue work in quadrant 1 (the warps that share the MMA warp's scheduler)", 16: "no epilogue work in quadrant 2"}
for v in [int(x) for x in a.variants.split(",")]:
    os.environ["MMALIGN_K1_DIAG"] = str(v)
    kw = dict(shard=(0, a.M), k_values=(1, 5, 10, 20), mrr_cutoff=100, weak_weight=(0.3, 0.2))
    eng.fused_pass(["vanilla_clip"], **kw)
    torch.cuda.synchronize()
    ts = []
    smp = Sampler()
    smp.start()
    for rep in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.fused_pass(["vanilla_clip"], **kw)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    smp.stop = True
    smp.join()
    half = len(smp.mhz) // 2  # the second half of the run: the cap has settled
    mhz, watts = sorted(smp.mhz[half:])[len(smp.mhz[half:]) // 2], sorted(smp.watts[half:])[len(smp.watts[half:]) // 2]
    ms = sum(ts[-3:]) / len(ts[-3:])
    print(f"cta_pairs={a.pairs} compact={a.compact} epi_sleep={a.epi_sleep} diag={v} ({names[v]}): {ms:.2f} ms = {2.0 * a.N * a.M * a.D / ms / 1e9:.0f} TFLOP/s  "
          f"(all reps: {', '.join(f'{t:.1f}' for t in ts)}); median SM clock {mhz} MHz, board power {watts:.0f} W", flush=True)
eng.close()
