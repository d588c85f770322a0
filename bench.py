#!/usr/bin/env python
"""bench.py -- top-K retrieval throughput of the alignment-scoring hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One step = one pass of the hot path over one batch of synthetic input: K0 operand
preparation + pair index, K1 fused tcgen05 score/top-K', K2 exact rescoring and ranking,
exact rescan of uncertified rows, K4 metric sums (and, for N > 1 GPUs, the NCCL merge).
Workload = BASELINE.json config 5, the one `metric` is quoted on: 1M images x 1M chunks,
D=512, all four schemas in one pass, K in {1,5,10,20} + MRR@100, full N x M candidates.
Chunks are sharded over the GPUs (total work fixed: "strong" scaling).

`value`   queries/s with the inputs already resident in HBM (device pointers through the C ABI).
`e2e`     the same step through the same C-ABI calls with HOST (pinned) input buffers and HOST
          output buffers: host->device and device->host copies inside the timed region.
`--impl reference`  the CPU port of the reference path (oracle/numpy_port.py, BLAS on all
          host cores) on a bounded row sample of the same workload.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
PKG = "multimodal-alignment-of-noisy-image-text-pairs-using-weak-supervision_b200"
SCHEMAS = ["vanilla_clip", "clip_lexical", "clip_positional", "clip_combined"]
K_VALUES = (1, 5, 10, 20)
MRR_CUTOFF = 100
WEAK = (0.3, 0.2)
T_TERMS = 512
METRIC = "top-10 retrieval queries/s at 1M x 1M, D=512 (all four schemas, K<=20 + MRR@100)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--N", type=int, default=1_000_000)
    ap.add_argument("--M", type=int, default=1_000_000)
    ap.add_argument("--D", type=int, default=512)
    ap.add_argument("--cpu-rows", type=int, default=0, help="rows of the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--kprime", type=int, default=0)
    ap.add_argument("--exchange", default="auto", choices=["auto", "alltoall", "allgather", "none"],
                    help="multi-GPU: none = contraction and rescoring both sharded by query rows (no list exchange); "
                         "alltoall = contraction sharded by chunk columns, rescoring by query rows; auto (default) = none "
                         "when every rank's query slab fills the GPU, else alltoall; allgather = fully sharded variant "
                         "(distributed.AllGatherScorer)")
    return ap.parse_args()


def peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        p = json.loads(f.read_text())
        return dict(bf16=p["bf16_tflops"], bf16_sustained=p["bf16_tflops_sustained"], hbm=p["hbm_gbs"], src="measured")
    return dict(bf16=1590.0, bf16_sustained=1400.0, hbm=6650.0, src="fallback")


class ClockSampler:
    """SM clock and throttle reasons during the timed region, sampled every 200 ms through NVML in this
    process (an `nvidia-smi -lms` poller was measured to stall the driver for tens of ms per query)."""

    def __init__(self, index):
        self.index, self.samples, self._stop, self.thread, self.err = index, [], threading.Event(), None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)
            return

        def loop():
            R = pynvml
            names = {"hw_slowdown": R.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": R.nvmlClocksEventReasonHwThermalSlowdown,
                     "sw_thermal_slowdown": R.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": R.nvmlClocksEventReasonSwPowerCap}
            it, reasons, pw = 0, [], 0.0
            while not self._stop.is_set():
                try:
                    sm = R.nvmlDeviceGetClockInfo(h, R.NVML_CLOCK_SM)
                    # the throttle-reason and power queries were measured (tools/sampler_diag.py) to stall a running
                    # step by 150-450 ms now and then; the SM clock query does not.  So: clock every 200 ms,
                    # reasons and power every 2 s.
                    if it % 10 == 0:
                        mask = R.nvmlDeviceGetCurrentClocksEventReasons(h)
                        reasons = [n for n, b in names.items() if mask & b]
                        pw = R.nvmlDeviceGetPowerUsage(h) / 1000.0
                    self.samples.append((time.time(), sm, pw, reasons))
                except Exception as e:  # noqa: BLE001
                    self.err = repr(e)
                it += 1
                self._stop.wait(0.2)
        self.thread = threading.Thread(target=loop, daemon=True)
        self.thread.start()

    def stop(self, t_begin=0.0, t_end=1e30):
        """Summary of the samples taken in [t_begin, t_end] (the timed region)."""
        self._stop.set()
        if self.thread:
            self.thread.join(1.0)
        win = [x for x in self.samples if t_begin <= x[0] <= t_end + 0.2]
        try:  # one more look at the throttle reasons right at the end of the region
            import pynvml as R
            h = R.nvmlDeviceGetHandleByIndex(self.index)
            mask = R.nvmlDeviceGetCurrentClocksEventReasons(h)
            extra = [n for n, b in (("hw_slowdown", R.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", R.nvmlClocksEventReasonHwThermalSlowdown),
                                    ("sw_thermal_slowdown", R.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", R.nvmlClocksEventReasonSwPowerCap)) if mask & b]
            if win:
                win.append((win[-1][0], win[-1][1], win[-1][2], extra))
        except Exception:  # noqa: BLE001
            pass
        if not win:
            return {"sm_mhz": None, "sm_max_mhz": getattr(self, "max_sm", None), "reasons": ["no samples: " + str(self.err)]}
        sm = [x[1] for x in win]
        reasons = sorted({r for x in win for r in x[3]})
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(self.max_sm), "reasons": reasons,
                "power_w_max": max(x[2] for x in win), "samples": len(win),
                "source": "NVML: SM clock every 200 ms, throttle reasons and power every 2 s"}


REJECT = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown")


def clocks_rejected(clocks) -> bool:
    """True when the timed region saw a hardware / thermal slowdown, or SM clocks far below the maximum with no
    throttle reason at all (a leftover clock lock).  sw_power_cap alone is normal for a dense GEMM on a 1 kW part."""
    if not clocks or clocks.get("sm_mhz") is None:
        return False
    reasons = clocks.get("reasons") or []
    if any(r in REJECT for r in reasons):
        return True
    return not reasons and clocks["sm_mhz"] < 0.6 * (clocks.get("sm_max_mhz") or 0)


# ------------------------------------------------------------------------------------------- reference arm
def cpu_sample(img_h, chk_h, rows, budget_note):
    """The CPU port on `rows` query rows against the full chunk table; returns (queries/s, seconds)."""
    from oracle import numpy_port
    t0 = time.perf_counter()
    numpy_port.evaluate(img_h, chk_h, T=T_TERMS, schemas=(0, 1, 2, 3), lam=(WEAK[0], WEAK[1], WEAK[0] + WEAK[1]),
                        kmax=max(K_VALUES), cutoff=MRR_CUTOFF, rows=rows)
    dt = time.perf_counter() - t0
    return len(rows) / dt, dt


def host_corpus(args, synthetic, device):
    """Synthetic corpus as host numpy arrays (generated on the GPU when there is one: same generator)."""
    import torch
    if device is not None:
        img, chk, _ = synthetic.make_torch(args.N, args.M, args.D, T=T_TERMS, device=device)
        to = lambda d: {k: (v.cpu().numpy() if v is not None else None) for k, v in d.items()}
        img, chk = to(img), to(chk)
        for d in (img, chk):
            d["key"] = d["key"].view(np.uint64)
            if d["terms"] is not None:
                d["terms"] = d["terms"].view(np.uint64)
        torch.cuda.empty_cache()
        return img, chk
    img, chk, _ = synthetic.make_numpy(args.N, args.M, args.D, T=T_TERMS)
    return img, chk


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.exchange == "auto":  # the same config line as the GPU arm prints for this --gpus
        distributed = importlib.import_module(PKG + ".distributed")
        args.exchange = "none" if distributed.slab_size(args.N, args.gpus) >= 148 * 128 else "alltoall"
    import torch
    synthetic = importlib.import_module(PKG + ".synthetic")
    from oracle import numpy_port
    dev = "cuda" if torch.cuda.is_available() else None
    img, chk = host_corpus(args, synthetic, dev)
    cores = os.cpu_count()
    rows_n = args.cpu_rows or 256
    rng = np.random.default_rng(0)
    times = []
    for it in range(args.warmup + args.steps):
        rows = np.sort(rng.choice(args.N, size=min(rows_n, args.N), replace=False))
        qps, dt = cpu_sample(img, chk, rows, "")
        if it == 0 and not args.cpu_rows:  # size the sample so that one step is ~5 s
            rows_n = int(max(64, min(args.N, rows_n * 5.0 / max(dt, 1e-3))))
        if it >= args.warmup:
            times.append((len(rows), dt))
    q = sum(n for n, _ in times)
    t = sum(d for _, d in times)
    value = q / t
    line = {
        "impl": "reference", "metric": METRIC,
        "value": value, "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * t / len(times), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": cores, "kind": "port",
                         "sample": f"{times[-1][0]} query rows per step x all {args.M} chunks, D={args.D}, 4 schemas, "
                                   f"K<=20 + MRR@100; numpy sgemm ({numpy_port.blas_info()}) + argpartition/lexsort"},
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload(args, G):
    ex = getattr(args, "exchange", "alltoall")
    if G == 1:
        how = "one GPU"
    elif ex == "allgather":
        how = f"chunks sharded over {G} GPUs, images replicated (fully sharded variant)"
    else:
        how = (f"each of {G} GPUs ingests 1/{G} of the images and of the chunks (an NVLink all-gather inside the step replicates "
               "them); " + ("contraction sharded by chunk columns, exact rescoring by query rows" if ex == "alltoall"
                            else "contraction and exact rescoring sharded by query rows"))
    return {"workload": f"BASELINE config 5: {args.N} images x {args.M} chunks, D={args.D}, all four schemas in one pass, "
                        f"K in {list(K_VALUES)} + MRR@{MRR_CUTOFF}, candidates=all, weak_weight={WEAK}",
            "N": args.N, "M": args.M, "D": args.D, "schemas": 4, "k_values": list(K_VALUES), "mrr_cutoff": MRR_CUTOFF,
            "sharding": how,
            "l2": "inputs (2 x %.1f GB bf16 operands) are far larger than the 126 MB L2" % (args.N * args.D * 2 / 1e9)}


# ------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    pkg = importlib.import_module(PKG)
    synthetic = importlib.import_module(PKG + ".synthetic")
    distributed = importlib.import_module(PKG + ".distributed")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a B200: the scoring path is CUDA-only (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    P = peaks()
    N, M, D = args.N, args.M, args.D
    r0, r1 = distributed.shard_range(M, world, rank)
    eng = pkg.AlignmentEngine(local)
    run_kw = dict(schemas=SCHEMAS, k_values=K_VALUES, mrr_cutoff=MRR_CUTOFF, weak_weight=WEAK, kprime=args.kprime)
    phase_ms, step_ms = {}, {}
    if args.exchange == "auto":
        args.exchange = "none" if distributed.slab_size(N, world) >= 148 * 128 else "alltoall"
    if args.exchange in ("alltoall", "none") or world == 1:
        # every rank ingests its slab of the images and its shard of the chunks
        i0, i1 = distributed.slab_range(N, world, rank)
        img, chk, meta = synthetic.make_torch(N, M, D, T=T_TERMS, device=dev, row0=r0, rows=r1 - r0, img_row0=i0,
                                              img_rows=i1 - i0)
        sharded = distributed.ShardedScorer(eng, world, rank, dev, contraction="rows" if args.exchange == "none" else "columns")

        def step(im, ck, host_out):
            t0 = time.perf_counter()
            sharded.load(im, ck, N=N, M=M, n_terms=T_TERMS)
            t2 = time.perf_counter()
            out = sharded.run(host_outputs=host_out, **run_kw)
            t3 = time.perf_counter()
            step_ms.setdefault("host" if host_out else "device", []).append(round(1e3 * (t3 - t0), 1))
            phase_ms["host" if host_out else "device"] = dict(load=1e3 * (t2 - t0), run=1e3 * (t3 - t2),
                                                              **{k: v for k, v in out["stats"].items() if k.endswith("_us")},
                                                              **({"exchange": out["phases_ms"]} if "phases_ms" in out else {}))
            return out
    else:
        img, chk, meta = synthetic.make_torch(N, M, D, T=T_TERMS, device=dev, row0=r0, rows=r1 - r0)
        sharded = distributed.AllGatherScorer(eng, world, rank, dev)

        def step(im, ck, host_out):
            t0 = time.perf_counter()
            eng.set_images(im["emb"], im["key"], im["bbox"], im["terms"])
            t1 = time.perf_counter()
            eng.set_chunks(ck["emb"], ck["key"], ck["bbox"], ck["terms"], n_terms=T_TERMS, col_offset=r0)
            t2 = time.perf_counter()
            out = sharded.run(host_outputs=host_out, **run_kw)
            t3 = time.perf_counter()
            step_ms.setdefault("host" if host_out else "device", []).append(round(1e3 * (t3 - t0), 1))
            phase_ms["host" if host_out else "device"] = dict(set_images=1e3 * (t1 - t0), set_chunks=1e3 * (t2 - t1),
                                                              run=1e3 * (t3 - t2), **{k: v for k, v in out["stats"].items() if k.endswith("_us")},
                                                              **({"exchange": out["phases_ms"]} if "phases_ms" in out else {}))
            return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = None
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), out

    # ---- device-resident arm
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # started before the warm-up so that its start-up cost is not in the timed region
    for _ in range(args.warmup):
        res = step(img, chk, False)
    fused_us, resc_us, scan_us, launches = [], [], [], []

    def dev_step():
        r = step(img, chk, False)
        fused_us.append(r["stats"]["fused_us"]); resc_us.append(r["stats"]["rescore_us"])
        scan_us.append(r["stats"]["exact_scan_us"])
        # kernels of the library per step (profiles/r1c_launches_summary.txt): K0 + error max of both tables (4), pair index
        # (iota, 10 CUB radix-sort/scan launches, page ranges, offsets scan: 14), then what mmalign_run counts itself
        # (fused, rescore, prefilter, scan, 2 metric kernels), + the list export when sharded
        launches.append(18 + r["stats"]["kernel_launches"] - 2 + (1 if world > 1 else 0))
        return r
    t_begin = time.time()
    ms, res = timed(dev_step, args.steps)
    clocks = sampler.stop(t_begin, time.time()) if rank == 0 else None
    # a throttled (or clock-locked) run is measured again, once; rank 0 saw the clocks, every rank has to follow
    redo = torch.tensor([1 if (rank == 0 and clocks_rejected(clocks)) else 0], device=dev)
    if world > 1:
        dist.broadcast(redo, 0)
    if int(redo.item()):
        first = clocks
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        step(img, chk, False)
        for l_ in (fused_us, resc_us, scan_us, launches):
            l_.clear()
        t_begin = time.time()
        ms, res = timed(dev_step, args.steps)
        if rank == 0:
            clocks = sampler.stop(t_begin, time.time())
            clocks["remeasured_after"] = first
    ms_per_step = ms / args.steps
    value = N / (ms_per_step / 1000.0)

    # ---- end-to-end arm: host (pinned) inputs, host outputs
    e2e = None
    if not args.no_e2e:
        pin = lambda d: {k: (v.cpu().pin_memory() if v is not None else None) for k, v in d.items()}
        img_h, chk_h = pin(img), pin(chk)
        h2d = sum(v.numel() * v.element_size() for d in (img_h, chk_h) for v in d.values() if v is not None)
        for _ in range(min(args.warmup, 2)):
            res_h = step(img_h, chk_h, True)
        ms_h, res_h = timed(lambda: step(img_h, chk_h, True), args.steps)
        io = torch.tensor([float(h2d), float(res_h["d2h_bytes"])], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(io)  # bytes of the whole job: every rank copies its own shard in and its own results out
        e2e = {"value": N / (ms_h / args.steps / 1000.0), "unit": "queries/s", "h2d_bytes_per_step": int(io[0].item()),
               "d2h_bytes_per_step": int(io[1].item()), "ms_per_step": ms_h / args.steps}
        del img_h, chk_h

    if os.environ.get("MMALIGN_BENCH_RANKS"):  # every rank's view of its last step
        print(f"[rank {rank}] phases of the last step:", json.dumps(phase_ms), file=sys.stderr, flush=True)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel (K1), timed with CUDA events inside the library
    t_fused = float(np.mean(fused_us)) * 1e-6
    flops_launch = 2.0 * N * (r1 - r0) * D  # algorithmic: 2*M_local*D per query x N queries per launch
    if world > 1 and args.exchange == "none":
        flops_launch = 2.0 * (i1 - i0) * M * D  # this rank's query slab against every chunk
    achieved = flops_launch / t_fused / 1e12 if t_fused > 0 else 0.0
    traffic = None  # DRAM bytes of one launch, from the committed ncu --set full capture of this very configuration
    tf = ROOT / "profiles" / "r1c_traffic.json"
    if tf.exists():
        t = json.loads(tf.read_text())["fused_score_topk_kernel"]
        if (t["N"], t["M"], t["D"]) == (N, r1 - r0, D):
            traffic = t["dram_bytes_read"] + t["dram_bytes_write"]
    roof = {"bound": "tensor", "kernel": "fused_score_topk_kernel", "achieved": achieved, "peak": P["bf16_sustained"],
            "unit": "TFLOP/s", "frac": achieved / P["bf16_sustained"], "traffic": traffic,
            "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of one launch (profiles/r1c_fused_score_topk_kernel.txt); "
                            "the kernel is tensor-bound: 2*N*M*D flop against (N+M)*D*2 algorithmic operand bytes",
            "peak_source": f"{P['src']} bf16_tflops_sustained (kernel runs ~{t_fused * 1e3:.0f} ms inside the step)",
            "frac_of_burst_peak": achieved / P["bf16"], "ms_per_launch": t_fused * 1e3,
            "share_of_step": t_fused / (ms_per_step / 1000.0),
            "other_phases_ms": {"rescore_kernel": float(np.mean(resc_us)) / 1e3,
                                "exact_scan_kernel": float(np.mean(scan_us)) / 1e3}}
    m = res["metrics"]
    line = {
        "metric": METRIC,
        "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic", "config": workload(args, world), "clocks": clocks,
        "e2e": e2e, "gpu_launches": int(np.sum(launches)), "roofline": roof,
        "quality": {"top1_vanilla": m["top_k"][0][0], "top10_vanilla": m["top_k"][0][2], "mrr_vanilla": m["mrr"][0],
                    "mrr_combined": m["mrr"][-1], "num_pairs": m["num_pairs"],
                    "rows_rescanned": res["stats"]["rows_rescanned"], "kprime": res["stats"]["kprime"],
                    "candidates_rescored_per_row": res["stats"]["candidates_rescored"] / max(N, 1)},
    }
    if world == 1 and not args.no_cpu_baseline:
        to_np = lambda d: {k: (v.cpu().numpy() if v is not None else None) for k, v in d.items()}
        ih, ch = to_np(img), to_np(chk)
        for d_ in (ih, ch):
            d_["key"] = d_["key"].view(np.uint64)
            if d_["terms"] is not None:
                d_["terms"] = d_["terms"].view(np.uint64)
        from oracle import numpy_port
        rng = np.random.default_rng(0)
        rows = np.sort(rng.choice(N, size=min(128, N), replace=False))
        qps, dt = cpu_sample(ih, ch, rows, "")
        n2 = int(max(64, min(N, 128 * 12.0 / max(dt, 1e-3))))
        rows = np.sort(rng.choice(N, size=n2, replace=False))
        qps, dt = cpu_sample(ih, ch, rows, "")
        line["cpu_baseline"] = {"value": qps, "unit": "queries/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": f"{n2} query rows x all {M} chunks in {dt:.1f} s (oracle/numpy_port.py: numpy sgemm "
                                          f"[{numpy_port.blas_info()}] + argpartition/lexsort, same workload)"}
    print("wall ms of every step (warm-up included):", json.dumps(step_ms), file=sys.stderr)
    print("phase wall times of the last step (ms; *_us from CUDA events):", json.dumps(phase_ms), file=sys.stderr)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
