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
ALL4 = ["vanilla_clip", "clip_lexical", "clip_positional", "clip_combined"]
MRR_CUTOFF = 100
WEAK = (0.3, 0.2)
T_TERMS = 512
METRIC5 = "top-10 retrieval queries/s at 1M x 1M, D=512 (all four schemas, K<=20 + MRR@100)"
VERIFY_ROWS = 64

# BASELINE.json `configs`, in order.  Config 5 is the one `metric` is quoted on (the default); config 1 is the
# reference's own CPU-runnable case: the same-page candidates of its SQL join, its K values, no ranking weights.
CONFIGS = {
    1: dict(N=1_000, M=5_000, D=512, schemas=["vanilla_clip"], candidates="same_page", k_values=(1, 5, 10), weak=(0.0, 0.0),
            name="ViT-B-32 (D=512) vanilla_clip scoring, 1k images x 5k text chunks, same-page candidates (the reference's SQL join)"),
    2: dict(N=100_000, M=500_000, D=512, schemas=["clip_combined"], candidates="all", k_values=(1, 5, 10, 20), weak=WEAK,
            name="ViT-B-32 clip_combined (cos + lexical + positional), 100k x 500k"),
    3: dict(N=1_000_000, M=1_000_000, D=768, schemas=["clip_lexical"], candidates="all", k_values=(1, 5, 10, 20), weak=WEAK,
            name="ViT-L-14 (D=768) clip_lexical, 1M x 1M"),
    4: dict(N=2_000_000, M=4_000_000, D=1024, schemas=["clip_positional"], candidates="all", k_values=(1, 5, 10, 20), weak=WEAK,
            name="ViT-H-14 (D=1024) clip_positional, 2M images x 4M chunks"),
    5: dict(N=1_000_000, M=1_000_000, D=512, schemas=ALL4, candidates="all", k_values=(1, 5, 10, 20), weak=WEAK,
            name="all four schemas in one pass, K in {1,5,10,20} + MRR, 1M x 1M at D=512"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=5, choices=sorted(CONFIGS), help="BASELINE.json config (default 5, the metric's)")
    ap.add_argument("--schemas", default=None, help="comma-separated schema names (default: the config's)")
    ap.add_argument("--N", type=int, default=None, help="override the config's image count (the line is then labelled custom)")
    ap.add_argument("--M", type=int, default=None)
    ap.add_argument("--D", type=int, default=None)
    ap.add_argument("--cpu-rows", type=int, default=0, help="rows of the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-verify", action="store_true")
    ap.add_argument("--no-prefetch", action="store_true", help="end-to-end arm: upload every step's inputs inside its own step "
                    "instead of behind the previous step's kernels")
    ap.add_argument("--kprime", type=int, default=0)
    ap.add_argument("--pipeline-rows", type=int, default=0, help="query rows per pipeline slab of mmalign_run (0 = auto)")
    ap.add_argument("--option", action="append", help="name=value for mmalign_set_option (repeatable)")
    ap.add_argument("--cta-pairs", type=int, default=-1, help="fused kernel on CTA pairs (cta_group::2): 1 / 0, -1 = library default")
    ap.add_argument("--exchange", default="auto", choices=["auto", "alltoall", "allgather", "none"],
                    help="multi-GPU: none = contraction and rescoring both sharded by query rows (no list exchange); "
                         "alltoall = contraction sharded by chunk columns, rescoring by query rows; auto (default) = none "
                         "when every rank's query slab fills the GPU, else alltoall; allgather = fully sharded variant "
                         "(distributed.AllGatherScorer)")
    a = ap.parse_args()
    cfg = CONFIGS[a.config]
    a.custom = any(getattr(a, k) is not None and getattr(a, k) != cfg[k] for k in ("N", "M", "D")) or a.schemas is not None
    for k in ("N", "M", "D"):
        if getattr(a, k) is None:
            setattr(a, k, cfg[k])
    a.schema_list = [x.strip() for x in a.schemas.split(",")] if a.schemas else list(cfg["schemas"])
    for x in a.schema_list:
        if x not in ALL4:
            ap.error(f"unknown schema {x!r}")
    a.schema_list = [x for x in ALL4 if x in a.schema_list]
    a.candidates, a.k_values, a.weak, a.cfg_name = cfg["candidates"], tuple(cfg["k_values"]), tuple(cfg["weak"]), cfg["name"]
    return a


def metric_name(a):
    if a.config == 5 and not a.custom:
        return METRIC5
    return (f"top-10 retrieval queries/s at {a.N} x {a.M}, D={a.D} ({'+'.join(a.schema_list)}, K<={max(a.k_values)} + "
            f"MRR@{MRR_CUTOFF}, candidates={a.candidates})")


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
def blas_threads_all():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm is entitled to every host core."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count())
    except Exception:  # noqa: BLE001
        pass


def schema_ids(args):
    return tuple(ALL4.index(x) for x in args.schema_list)


def cpu_sample(args, img_h, chk_h, rows):
    """The CPU port on `rows` query rows against the full chunk table; returns (queries/s, seconds)."""
    lam = (args.weak[0], args.weak[1], args.weak[0] + args.weak[1])
    t0 = time.perf_counter()
    if args.candidates == "all":
        from oracle import numpy_port
        numpy_port.evaluate(img_h, chk_h, T=T_TERMS, schemas=schema_ids(args), lam=lam, kmax=max(args.k_values),
                            cutoff=MRR_CUTOFF, rows=rows)
    else:  # the reference's same-page join: the C restatement (pthreads over the rows)
        from oracle import oracle
        sub = {k: (v[rows] if v is not None else None) for k, v in img_h.items()}
        oracle.evaluate(sub, chk_h, T=T_TERMS, schema_mask=sum(1 << i for i in schema_ids(args)), candidates="same_page",
                        lam=lam, kmax=max(args.k_values), cutoff=MRR_CUTOFF)
    dt = time.perf_counter() - t0
    return len(rows) / dt, dt


def cpu_kind(args):
    if args.candidates == "all":
        from oracle import numpy_port
        return (f"oracle/numpy_port.py: numpy sgemm [{numpy_port.blas_info()}] in column slabs, fp32 argpartition on a thread "
                "pool, fp64 exact scores of the kept candidates, lexsort")
    return "oracle/mmalign_oracle.c: the C restatement of the reference's same-page SQL ranking, pthreads over the rows"


def reference_as_written():
    """Wall time of the UNMODIFIED reference metric functions at config 1, recorded in the build container (the
    reference tree does not travel to the GPU box): tools/time_reference_config1.py -> profiles/bench/."""
    f = ROOT / "profiles" / "bench" / "r2_reference_as_written_config1.json"
    return json.loads(f.read_text()) if f.exists() else None


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


def resolve_exchange(args, distributed):
    if args.exchange == "auto":
        args.exchange = "none" if distributed.slab_size(args.N, args.gpus) >= 148 * 128 else "alltoall"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    blas_threads_all()
    distributed = importlib.import_module(PKG + ".distributed")
    resolve_exchange(args, distributed)  # the same config line as the GPU arm prints for this --gpus
    import torch
    synthetic = importlib.import_module(PKG + ".synthetic")
    dev = "cuda" if torch.cuda.is_available() else None
    img, chk = host_corpus(args, synthetic, dev)
    cores = os.cpu_count()
    rows_n = args.cpu_rows or min(256, args.N)
    rng = np.random.default_rng(0)
    times = []
    for it in range(args.warmup + args.steps):
        rows = np.sort(rng.choice(args.N, size=min(rows_n, args.N), replace=False))
        qps, dt = cpu_sample(args, img, chk, rows)
        if it == 0 and not args.cpu_rows:  # size the sample so that one step is ~5 s
            rows_n = int(max(64, min(args.N, rows_n * 5.0 / max(dt, 1e-3))))
        if it >= args.warmup:
            times.append((len(rows), dt))
    q = sum(n for n, _ in times)
    t = sum(d for _, d in times)
    value = q / t
    line = {
        "impl": "reference", "metric": metric_name(args),
        "value": value, "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * t / len(times), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": cores, "kind": "port",
                         "sample": f"{times[-1][0]} query rows per step x all {args.M} chunks, D={args.D}, "
                                   f"{len(args.schema_list)} schema(s), K<={max(args.k_values)} + MRR@{MRR_CUTOFF}; {cpu_kind(args)}",
                         "reference_as_written": reference_as_written()},
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload(args, G):
    ex = getattr(args, "exchange", "alltoall")
    if G == 1:
        how = "one GPU"
    elif ex == "allgather":
        how = f"chunks sharded over {G} GPUs, images replicated (fully sharded variant)"
    elif ex == "alltoall":
        how = (f"each of {G} GPUs ingests 1/{G} of the images and of the chunks (an NVLink all-gather inside the step replicates "
               "them); contraction sharded by chunk columns, exact rescoring by query rows")
    else:
        how = (f"each of {G} GPUs ingests 1/{G} of the images and of the chunks and prepares its chunk shard; the prepared operands "
               "are all-gathered over NVLink inside the step (fp32 master rows on a side stream); contraction and exact "
               "rescoring sharded by query rows")
    label = f"BASELINE config {args.config}: {args.cfg_name}" if not args.custom else \
        f"custom shape (NOT a BASELINE config; nearest: config {args.config})"
    return {"workload": f"{label} -- {args.N} images x {args.M} chunks, D={args.D}, schemas={args.schema_list}, "
                        f"K in {list(args.k_values)} + MRR@{MRR_CUTOFF}, candidates={args.candidates}, weak_weight={tuple(args.weak)}",
            "baseline_config": None if args.custom else args.config,
            "N": args.N, "M": args.M, "D": args.D, "schemas": args.schema_list, "k_values": list(args.k_values),
            "mrr_cutoff": MRR_CUTOFF, "candidates": args.candidates, "sharding": how,
            "l2": "inputs (%.2f + %.2f GB of bf16 operands, %.2f + %.2f GB of fp32 rows) against a 126 MB L2; nothing is flushed "
                  "between steps because nothing of that size stays" % (args.N * args.D * 2 / 1e9, args.M * args.D * 2 / 1e9,
                                                                         args.N * args.D * 4 / 1e9, args.M * args.D * 4 / 1e9)}


# ------------------------------------------------------------------------------------------- our arm
def verify_rows(args, eng, sharded, img, chk, res, world, rank, dev):
    """After the timed region: VERIFY_ROWS sampled query rows of the timed result against the CPU oracle (bit-exact:
    top-K indices and scores, true-pair ranks and similarities).  Rank 0 checks rows of its own slab."""
    import torch
    from oracle import oracle
    to_np = lambda t: None if t is None else (t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t))
    n_local = int(res["topk_idx"].shape[1])
    if n_local == 0:
        return {"rows": 0, "ok": True}
    rng = np.random.default_rng(12345)
    # 64 rows at config 5's 1M x 512 per row; fewer where a row costs more (the oracle scans every column of every sampled row)
    n_rows = int(max(8, min(VERIFY_ROWS, VERIFY_ROWS * (1e6 * 512) / max(1.0, float(args.M) * args.D))))
    rows = np.sort(rng.choice(n_local, size=min(n_rows, n_local), replace=False))
    if world > 1 and getattr(sharded, "mode", "") == "rows":     # the engine holds the slab; the chunk table is gathered
        full = {f: sharded._full.get(("chk", f)) for f in ("emb", "key", "bbox", "terms")}
        chk_h = {f: (to_np(v[:args.M]) if v is not None else None) for f, v in full.items()}
        img_rows = rows                                          # positions within this rank's slab (img is the slab)
    elif world > 1:
        full = {f: sharded._full.get(("chk", f)) for f in ("emb", "key", "bbox", "terms")}
        chk_h = {f: (to_np(v[:args.M]) if v is not None else None) for f, v in full.items()}
        img_rows = rows
    else:
        chk_h = {f: to_np(chk[f]) for f in ("emb", "key", "bbox", "terms")}
        img_rows = rows
    ridx = torch.as_tensor(img_rows, device=img["emb"].device) if hasattr(img["emb"], "device") else img_rows
    sub = {f: (to_np(img[f][ridx]) if img.get(f) is not None else None) for f in ("emb", "key", "bbox", "terms")}
    for d in (sub, chk_h):
        d["key"] = d["key"].view(np.uint64)
        if d["terms"] is not None:
            d["terms"] = d["terms"].view(np.uint64)
    t0 = time.perf_counter()
    o = oracle.evaluate(sub, chk_h, T=T_TERMS, schema_mask=sum(1 << ALL4.index(x) for x in args.schema_list),
                        candidates=args.candidates, lam=(args.weak[0], args.weak[1], args.weak[0] + args.weak[1]),
                        kmax=max(args.k_values), cutoff=MRR_CUTOFF)
    dt = time.perf_counter() - t0
    tk_i, tk_s = to_np(res["topk_idx"][:, rows]), to_np(res["topk_score"][:, rows])
    off = to_np(res["pair_offsets_local"])
    sel = np.concatenate([np.arange(off[r], off[r + 1]) for r in rows]) if len(rows) else np.zeros(0, np.int64)
    pr, ps = to_np(res["pair_rank"])[:, sel], to_np(res["pair_sim"])[sel]
    bad = []
    if not np.array_equal(tk_i, o["topk_idx"]):
        bad.append("topk_idx")
    if not np.array_equal(tk_s, o["topk_score"]):
        bad.append("topk_score")
    if not np.array_equal(pr, o["pair_rank"]):
        bad.append("pair_rank")
    if not np.array_equal(ps, o["pair_sim"]):
        bad.append("pair_sim")
    return {"rows": int(len(rows)), "pairs": int(len(sel)), "ok": not bad, "mismatch": bad, "oracle_seconds": round(dt, 1),
            "checked": "top-K indices and scores, true-pair ranks and similarities of the last timed step, bit for bit "
                       "against oracle/mmalign_oracle.c on the same rows (rank 0's slab)"}


def bind_to_gpu_cores(local):
    """One process per GPU: keep the process -- and with it the pages of its pinned host buffers (first touch) -- on the
    cores NVML lists as local to its GPU, so that a rank's uploads do not cross the socket interconnect.  Returns the
    number of cores bound to (0: left alone)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        n = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local), (n + 63) // 64)
        cpus = [i for i in range(n) if (int(mask[i // 64]) >> (i % 64)) & 1]
        if cpus and len(cpus) < n:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:  # noqa: BLE001 -- an optimisation only
        pass
    return 0


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
    bound_cores = bind_to_gpu_cores(local) if world > 1 else 0
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    P = peaks()
    N, M, D = args.N, args.M, args.D
    args.gpus = world
    r0, r1 = distributed.shard_range(M, world, rank)
    eng = pkg.AlignmentEngine(local)
    if args.cta_pairs >= 0:
        eng.set_option("cta_pairs", args.cta_pairs)
    for kv in args.option or []:
        k, v = kv.split("=")
        eng.set_option(k, int(v))
    run_kw = dict(schemas=args.schema_list, k_values=args.k_values, mrr_cutoff=MRR_CUTOFF, weak_weight=args.weak,
                  kprime=args.kprime, candidates=args.candidates)
    phase_ms, step_ms = {}, {}
    resolve_exchange(args, distributed)
    i0, i1 = distributed.slab_range(N, world, rank)
    if args.exchange in ("alltoall", "none") or world == 1:
        # every rank ingests its slab of the images and its shard of the chunks
        img, chk, meta = synthetic.make_torch(N, M, D, T=T_TERMS, device=dev, row0=r0, rows=r1 - r0, img_row0=i0,
                                              img_rows=i1 - i0)
        sharded = distributed.ShardedScorer(eng, world, rank, dev, contraction="rows" if args.exchange == "none" else "columns")

        def step(im, ck, host_out, prefetch=False):
            t0 = time.perf_counter()
            sharded.load(im, ck, N=N, M=M, n_terms=T_TERMS)
            if host_out and prefetch:
                sharded.prefetch(im, ck)  # the next step's upload, behind this step's kernels (one upload per step either way)
            t2 = time.perf_counter()
            out = sharded.run(host_outputs=host_out, pipeline_rows=args.pipeline_rows, **run_kw)
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
    fused_path = args.candidates == "all"

    def dev_step():
        r = step(img, chk, False)
        fused_us.append(r["stats"]["fused_us"]); resc_us.append(r["stats"]["rescore_us"])
        scan_us.append(r["stats"]["exact_scan_us"])
        # kernels of the library per step: K0 + error max of the chunk table (2; the image K0 launches are counted by
        # mmalign_run, one per slab), pair index (iota, 10 CUB radix-sort/scan launches, page ranges, offsets scan: 14),
        # then what mmalign_run counts itself (K0 per slab, fused, rescore, prefilter, scan, 2 metric kernels), + the
        # list export when the contraction is sharded by columns
        launches.append(16 + r["stats"]["kernel_launches"] + (1 if world > 1 and args.exchange == "alltoall" else 0))
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

    # ---- the timed result, checked row by row against the oracle (rank 0, its slab)
    verified = None
    if rank == 0 and not args.no_verify and args.exchange != "allgather":
        off, _ = eng.pairs_device()
        q0 = res.get("topk_row0", 0) if (world > 1 and sharded.mode != "rows") else 0
        res["pair_offsets_local"] = (off[q0:q0 + res["topk_idx"].shape[1] + 1] - off[q0]).cpu()
        vimg = img
        verified = verify_rows(args, eng, sharded, vimg, chk, res, world, rank, dev)
    if world > 1:
        dist.barrier()

    # ---- end-to-end arm: host (pinned) inputs, host outputs
    e2e = None
    if not args.no_e2e:
        pin = lambda d: {k: (v.cpu().pin_memory() if v is not None else None) for k, v in d.items()}
        img_h, chk_h = pin(img), pin(chk)
        h2d = sum(v.numel() * v.element_size() for d in (img_h, chk_h) for v in d.values() if v is not None)
        # what the host link gives this rank while every rank copies at once (diagnostic, outside the timed region)
        barrier()
        probe = torch.empty_like(chk["emb"])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); probe.copy_(chk_h["emb"], non_blocking=True); e1.record()
        torch.cuda.synchronize()
        phase_ms["h2d_probe"] = dict(GBps=round(probe.numel() * 4 / e0.elapsed_time(e1) / 1e6, 1), bound_cores=bound_cores)
        del probe
        # (a) every step uploads its own inputs before it computes
        for _ in range(min(args.warmup, 2)):
            res_h = step(img_h, chk_h, True)
        ms_h, res_h = timed(lambda: step(img_h, chk_h, True), args.steps)
        strict = {"value": N / (ms_h / args.steps / 1000.0), "ms_per_step": ms_h / args.steps,
                  "uploads": "every step uploads its own inputs before it computes"}
        same = bool(np.array_equal(res_h["hits"], res["hits"]) and res_h["num_pairs"] == res["num_pairs"])
        # (b) streaming use: uploads double-buffered across steps
        pipelined = not args.no_prefetch and args.exchange != "allgather"
        if pipelined:
            for _ in range(min(args.warmup, 2)):
                res_h = step(img_h, chk_h, True, prefetch=True)
            ms_h, res_h = timed(lambda: step(img_h, chk_h, True, prefetch=True), args.steps)
            sharded._prefetched = None  # (the copy staged for a step that is not run)
            same = same and bool(np.array_equal(res_h["hits"], res["hits"]) and res_h["num_pairs"] == res["num_pairs"])
        io = torch.tensor([float(h2d), float(res_h["d2h_bytes"])], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(io)  # bytes of the whole job: every rank copies its own shard in and its own results out
        e2e = {"value": N / (ms_h / args.steps / 1000.0), "unit": "queries/s", "h2d_bytes_per_step": int(io[0].item()),
               "d2h_bytes_per_step": int(io[1].item()), "ms_per_step": ms_h / args.steps,
               "pipeline_slabs": res_h["stats"].get("slabs"),
               "uploads": (strict["uploads"] if not pipelined else
                           "double-buffered across steps (ShardedScorer.prefetch): the pinned host shards of step s+1 are copied to "
                           "device staging buffers behind the kernels of step s; one upload and one download per step inside the timed "
                           "region (the first timed step's inputs travelled during the last warm-up step, the last timed step uploads "
                           "for a step that is not run)"),
               "per_step_uploads": strict,  # the same arm with every step's upload inside the step itself, measured just before
               "same_result_as_device_arm": same}
        del img_h, chk_h

    if os.environ.get("MMALIGN_BENCH_RANKS"):  # every rank's view of its last step
        print(f"[rank {rank}] phases of the last step:", json.dumps(phase_ms), file=sys.stderr, flush=True)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel, timed with CUDA events inside the library
    # (in `all` mode the rescoring is three kernels -- select_kernel, gather_kernel, rank_kernel -- plus rescore_kernel
    # for the rows they hand over; in same-page mode rescore_kernel alone)
    t_resc_ms = float(np.mean(resc_us)) / 1e3
    other = {"rescore_kernel": t_resc_ms, "exact_scan_kernel": float(np.mean(scan_us)) / 1e3}
    if fused_path and t_resc_ms > 0:
        # algorithmic bytes of the rescoring: one fp32 row per re-scored candidate (the stat is the whole job's; a rank re-scores 1/G of it)
        gathered = res["stats"]["candidates_rescored"] / max(world, 1) * D * 4.0
        other["rescoring_gather_GBps"] = gathered / t_resc_ms / 1e6
        other["rescoring_frac_of_hbm"] = gathered / t_resc_ms / 1e6 / P["hbm"]
    if fused_path:
        t_fused = float(np.mean(fused_us)) * 1e-6
        flops_launch = 2.0 * N * (r1 - r0) * D  # algorithmic: 2*M_local*D per query x N queries
        if world > 1 and args.exchange == "none":
            flops_launch = 2.0 * (i1 - i0) * M * D  # this rank's query slab against every chunk
        achieved = flops_launch / t_fused / 1e12 if t_fused > 0 else 0.0
        traffic = None  # DRAM bytes of the kernel per step, from the committed ncu --set full capture of this configuration
        tf = ROOT / "profiles" / "r2_traffic.json"
        if tf.exists():
            t = json.loads(tf.read_text()).get("fused_score_topk_kernel", {})
            if (t.get("N"), t.get("M"), t.get("D")) == (N, r1 - r0, D):
                traffic = t["dram_bytes_read"] + t["dram_bytes_write"]
        n_fused = max(1, int(res["stats"]["fused_launches"]))
        roof = {"bound": "tensor", "kernel": "fused_score_topk_kernel", "achieved": achieved, "peak": P["bf16_sustained"],
                "unit": "TFLOP/s", "frac": achieved / P["bf16_sustained"], "traffic": traffic,
                "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of the kernel over one step (profiles/); "
                                "the kernel is tensor-bound: 2*N*M*D flop against (N+M)*D*2 algorithmic operand bytes",
                "peak_source": f"{P['src']} bf16_tflops_sustained (the kernel runs {t_fused * 1e3:.0f} ms of every step)",
                "frac_of_burst_peak": achieved / P["bf16"], "launches_per_step": n_fused,
                "ms_per_launch": t_fused * 1e3 / n_fused, "ms_per_step": t_fused * 1e3,
                "share_of_step": t_fused / (ms_per_step / 1000.0), "other_phases_ms": other}
    else:  # same-page candidates (config 1): the exact rescoring kernel alone, HBM row gathers
        t_resc = max(float(np.mean(resc_us)) * 1e-6, 1e-9)
        cand = res["stats"]["candidates_rescored"]
        bytes_launch = (cand + N) * D * 4.0
        roof = {"bound": "hbm", "kernel": "rescore_kernel", "achieved": bytes_launch / t_resc / 1e9, "peak": P["hbm"],
                "unit": "GB/s", "frac": bytes_launch / t_resc / 1e9 / P["hbm"], "traffic": None,
                "peak_source": f"{P['src']} hbm_gbs", "ms_per_launch": t_resc * 1e3,
                "note": f"{cand} same-page row gathers + {N} query rows of 4*D bytes; at this size the step is launch latency, "
                        "not bandwidth", "share_of_step": t_resc / (ms_per_step / 1000.0), "other_phases_ms": other}
    m = res["metrics"]
    line = {
        "metric": metric_name(args),
        "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic", "config": workload(args, world), "clocks": clocks,
        "e2e": e2e, "gpu_launches": int(np.sum(launches)), "roofline": roof, "verified_rows": verified,
        "quality": {"top1": m["top_k"][0][0], "top10": m["top_k"][0][min(2, len(args.k_values) - 1)], "mrr_first_schema": m["mrr"][0],
                    "mrr_last_schema": m["mrr"][-1], "num_pairs": m["num_pairs"],
                    "rows_rescanned": res["stats"]["rows_rescanned"], "kprime": res["stats"]["kprime"],
                    "eps_violations": res["stats"].get("eps_violations"),
                    "candidates_rescored_per_row": res["stats"]["candidates_rescored"] / max(N, 1)},
    }
    if world == 1 and not args.no_cpu_baseline:
        blas_threads_all()
        to_np = lambda d: {k: (v.cpu().numpy() if v is not None else None) for k, v in d.items()}
        ih, ch = to_np(img), to_np(chk)
        for d_ in (ih, ch):
            d_["key"] = d_["key"].view(np.uint64)
            if d_["terms"] is not None:
                d_["terms"] = d_["terms"].view(np.uint64)
        rng = np.random.default_rng(0)
        rows = np.sort(rng.choice(N, size=min(128, N), replace=False))
        qps, dt = cpu_sample(args, ih, ch, rows)
        n2 = int(max(64, min(N, 128 * 12.0 / max(dt, 1e-3))))
        rows = np.sort(rng.choice(N, size=n2, replace=False))
        qps, dt = cpu_sample(args, ih, ch, rows)
        line["cpu_baseline"] = {"value": qps, "unit": "queries/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": f"{n2} query rows x all {M} chunks in {dt:.1f} s ({cpu_kind(args)}; same workload)",
                                "reference_as_written": reference_as_written()}
    print("wall ms of every step (warm-up included):", json.dumps(step_ms), file=sys.stderr)
    print("phase wall times of the last step (ms; *_us from CUDA events):", json.dumps(phase_ms), file=sys.stderr)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if verified is not None and not verified["ok"]:
        raise SystemExit(f"verified_rows FAILED: {verified['mismatch']}")


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
