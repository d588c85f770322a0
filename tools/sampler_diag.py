"""Which NVML query perturbs a running step?  Runs the 1M x 1M step with different samplers."""
import importlib, sys, threading, time, torch, pynvml
sys.path.insert(0, ".")
PKG = "multimodal-alignment-of-noisy-image-text-pairs-using-weak-supervision_b200"
pkg = importlib.import_module(PKG); syn = importlib.import_module(PKG + ".synthetic")
N = M = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
img, chk, _ = syn.make_torch(N, M, 512, device="cuda")
eng = pkg.AlignmentEngine(0)
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
def step():
    t0 = time.perf_counter()
    eng.set_images(img["emb"], img["key"], img["bbox"], None)
    eng.set_chunks(chk["emb"], chk["key"], chk["bbox"], chk["terms"], n_terms=512)
    eng.run(pkg.SCHEMAS, candidates="all", k_values=(1, 5, 10, 20), weak_weight=(0.3, 0.2), device_outputs=True)
    return round(1e3 * (time.perf_counter() - t0))
def with_sampler(name, fn, period):
    stop = threading.Event(); calls = []
    def loop():
        while not stop.is_set():
            t0 = time.perf_counter(); fn(); calls.append(1e3 * (time.perf_counter() - t0)); stop.wait(period)
    th = threading.Thread(target=loop, daemon=True)
    if fn: th.start()
    times = [step() for _ in range(5)]
    stop.set()
    print(f"{name:34s} steps {times}  nvml call ms: max {max(calls) if calls else 0:.1f} n {len(calls)}")
step(); step()
with_sampler("no sampler", None, 0)
with_sampler("clock only, 200 ms", lambda: pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), 0.2)
with_sampler("reasons only, 200 ms", lambda: pynvml.nvmlDeviceGetCurrentClocksEventReasons(h), 0.2)
with_sampler("power only, 200 ms", lambda: pynvml.nvmlDeviceGetPowerUsage(h), 0.2)
with_sampler("clock+reasons, 200 ms", lambda: (pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)), 0.2)
with_sampler("no sampler again", None, 0)
