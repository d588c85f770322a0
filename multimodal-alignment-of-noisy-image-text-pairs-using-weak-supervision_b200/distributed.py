"""Multi-GPU plumbing: one process per GPU.  torch.distributed (NCCL over
NVLink) carries the exchanges of a step; every merge/count/reduction on the data is a kernel of
the CUDA library.

ShardedScorer (default).  The similarity contraction -- all of the FLOPs -- is sharded by CHUNK columns;
the exact rescoring and ranking is sharded by QUERY rows:

  load   every rank ingests 1/G of the images and 1/G of the chunks (host->device or device-resident) and
         an all-gather over NVLink replicates the tables (4 GB at config 5: ~20x faster than PCIe would be)
  1.     fused tcgen05 pass: every image row against this rank's chunk columns -> per-row candidate lists
  2.     all-to-all of the candidate lists: rank g receives, from every rank, the lists of query slab g
  3.     exact rescoring, certificate, rescan and metric sums of slab g against the whole chunk table --
         exactly the single-GPU second half, so top-K lists, true-pair ranks and similarities need no
         further merge; rank g owns the results of its slab
  4.     all-reduce(sum) of the metric sums.

AllGatherScorer (fully sharded variant: no rank ever holds another rank's chunk rows; for chunk tables
too large to replicate).  Images replicated, three exchanges:

  0. all-reduce(max) of the per-row list thresholds (and of the chunk rounding-error bound), then
     all-reduce(sum) of the per-row certificate counts            (mmalign_fused_pass / rescore_pass)
  1. all-to-all of each rank's exact top-K lists  -> mmalign_merge_topk: rank g merges and owns the lists
     of query slab g (`topk_row0`, ceil(N/G) rows)
  2. all-gather of each rank's true pairs (image, chunk, score per schema)
        -> mmalign_count_beating against the rank's own exact lists
        -> all-reduce(sum) of the counts            (global rank of every true pair)
  3. all-reduce(sum) of the metric sums from mmalign_reduce_metrics.

ShardedScorer with contraction="rows" (what "auto" picks when every rank's query slab fills its GPU): rank g
contracts and re-scores its own query slab against the whole chunk table, so no candidate list travels.  Each
rank prepares (K0) the 1/G of the chunk table it ingests; the prepared bf16 operands, norms and page keys are
all-gathered first and the contraction starts, while the fp32 master rows follow on a side stream for the
exact rescoring (mmalign_prep_rows / mmalign_set_chunks_prepared / mmalign_rescore_after).

With world size 1 no collective runs.
"""
from __future__ import annotations

import numpy as np


def shard_range(M: int, world: int, rank: int):
    """Contiguous column range of `rank`: ceil(M / world) columns each, the last one shorter."""
    per = -(-M // world) if world > 0 else M
    lo = min(M, rank * per)
    return lo, min(M, lo + per)


def metrics_from_sums(hits, rr_sum, sim_sum, num_pairs):
    """evaluate_alignments.py:192 (hits / len(pairs)), :216, :231 (means) from the reduced sums."""
    P = int(num_pairs)
    if P == 0:
        z = np.zeros_like(np.asarray(hits, np.float64))
        return dict(top_k=z.tolist(), mrr=[0.0] * len(rr_sum), avg_similarity=0.0, num_pairs=0)
    return dict(top_k=(np.asarray(hits, np.float64) / P).tolist(), mrr=(np.asarray(rr_sum) / P).tolist(),
                avg_similarity=float(sim_sum) / P, num_pairs=P)


def slab_size(N: int, world: int) -> int:
    """Query rows per rank: whole 128-row blocks (the fused kernel's row tile), ceil(blocks / world) each."""
    blocks = -(-N // 128)
    return max(1, -(-blocks // max(world, 1))) * 128


def slab_range(N: int, world: int, rank: int):
    per = slab_size(N, world)
    lo = min(N, rank * per)
    return lo, min(N, lo + per)


class ShardedScorer:
    """Contraction sharded by chunk columns or by query rows, rescoring by query rows (module docstring)."""
    FIELDS = ("emb", "key", "bbox", "terms")

    def __init__(self, engine, world: int = 1, rank: int = 0, device=None, dist=None, contraction: str = "auto"):
        """contraction="columns": the fused pass of rank g covers every image row against chunk shard g and the
        candidate lists are exchanged (all-to-all).  contraction="rows": rank g contracts its own query slab against
        the whole chunk table -- no list exchange, and the lists of a row warm up 2*splits times instead of
        2*splits*G times (the fused kernel's per-list warm-up is the one cost that grows with G in the column
        layout).  "auto" (default): rows when every slab fills the GPU on its own (>= 148 row blocks of 128
        queries), columns for query-poor shapes."""
        self.eng, self.world, self.rank, self.device = engine, world, rank, device
        if contraction not in ("auto", "columns", "rows"):
            raise ValueError("contraction must be 'auto', 'columns' or 'rows'")
        self.contraction = contraction
        if dist is None and world > 1:
            import torch.distributed as dist
        self.dist = dist
        self._full = {}
        self._side = None
        self._prefetched = None  # (image ids, staged device copies, event) of the next step: prefetch()
        self._pf_stream = None
        self.N = self.M = 0
        self.mode = "rows"
        self.MAX = getattr(getattr(dist, "ReduceOp", None), "MAX", "max") if dist is not None else "max"

    def _mode(self, N):
        if self.contraction != "auto":
            return self.contraction
        return "rows" if slab_size(N, self.world) >= 148 * 128 else "columns"

    # -- ingest ---------------------------------------------------------------------------------------
    @staticmethod
    def _as_tensor(x):
        import torch
        if x is None:
            return None
        if not type(x).__module__.startswith("torch"):
            x = torch.from_numpy(x.view("int64") if x.dtype.kind == "u" else x)
        return x.view(torch.int64) if x.dtype == torch.uint64 else x

    def _buffer(self, name, shape, dtype, like):
        import torch
        buf = self._full.get(name)
        if buf is None or tuple(buf.shape) != tuple(shape) or buf.dtype != dtype:
            buf = torch.empty(shape, dtype=dtype, device=self.device if self.device is not None else like.device)
            self._full[name] = buf
        return buf

    def _gather(self, buf, slot):
        import torch
        if buf.dtype == torch.int16:  # bf16 bit patterns travel as bytes (gloo has no 16-bit integer type)
            buf, slot = buf.view(torch.uint8), slot.view(torch.uint8)
        if hasattr(self.dist, "all_gather_into_tensor"):
            self.dist.all_gather_into_tensor(buf, slot)
        else:
            parts = [torch.empty_like(slot) for _ in range(self.world)]
            self.dist.all_gather(parts, slot.contiguous())
            per = slot.shape[0]
            for g, p_ in enumerate(parts):
                buf[g * per:(g + 1) * per].copy_(p_)

    def _replicate(self, side, shard, total, per):
        """All-gather of one table: every rank contributes `per` rows (the last ranks fewer; their slots are
        padded) into persistent [G * per, ...] buffers.  Returns the dict of full tables cut to `total` rows."""
        G, out = self.world, {}
        for f in self.FIELDS:
            x = self._as_tensor(shard.get(f))
            if x is None:
                out[f] = None
                continue
            buf = self._buffer((side, f), (G * per,) + tuple(x.shape[1:]), x.dtype, x)
            slot = buf[self.rank * per:(self.rank + 1) * per]
            slot[:x.shape[0]].copy_(x, non_blocking=True)  # host->device for pinned host shards
            self._gather(buf, slot)
            out[f] = buf[:total]
        return out

    def _load_rows(self, img, chk, N, M, n_terms):
        """Query-row layout: the rank keeps its own image slab; the chunk table is prepared where it is ingested
        (K0 of 1/G of the rows on every rank) and the PREPARED operands travel first -- bf16 rows, norms, rounding
        errors and page keys are all the contraction and the pair index need -- while the fp32 master rows, boxes
        and term sets, which only the exact rescoring reads, follow on a side stream beside the contraction."""
        import torch
        G, eng = self.world, self.eng
        per = -(-M // G) if M else 1
        x = {f: self._as_tensor(chk.get(f)) for f in self.FIELDS}
        m, D = int(x["emb"].shape[0]), int(x["emb"].shape[1])
        cuda = self.device is not None and torch.device(self.device).type == "cuda"
        full, slot = {}, {}
        for f in self.FIELDS:
            if x[f] is None:
                full[f] = None
                continue
            full[f] = self._buffer(("chk", f), (G * per,) + tuple(x[f].shape[1:]), x[f].dtype, x[f])
            slot[f] = full[f][self.rank * per:(self.rank + 1) * per]
            slot[f][:m].copy_(x[f], non_blocking=True)
        like = x["emb"]
        full["bf16"] = self._buffer(("chk", "bf16"), (G * per, D), torch.int16, like)
        full["norm2"] = self._buffer(("chk", "norm2"), (G * per,), torch.float32, like)
        full["err"] = self._buffer(("chk", "err"), (G * per,), torch.float32, like)
        for f in ("bf16", "norm2", "err"):
            slot[f] = full[f][self.rank * per:(self.rank + 1) * per]
        st = torch.cuda.current_stream().cuda_stream if cuda else None
        if m:
            eng.prep_rows(slot["emb"][:m], slot["bf16"][:m], slot["norm2"][:m], slot["err"][:m], stream=st)
        for f in ("bf16", "norm2", "err", "key"):     # what the contraction and the pair index read
            self._gather(full[f], slot[f])
        event = None
        if cuda:                                      # what only the exact rescoring reads: beside the contraction
            if self._side is None:
                self._side = torch.cuda.Stream(device=self.device)
            self._side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._side):
                for f in ("emb", "bbox", "terms"):
                    if full[f] is not None:
                        self._gather(full[f], slot[f])
                event = torch.cuda.Event()
                event.record(self._side)
        else:
            for f in ("emb", "bbox", "terms"):
                if full[f] is not None:
                    self._gather(full[f], slot[f])
        cutm = lambda t: None if t is None else t[:M]
        eng.set_chunks_prepared(cutm(full["emb"]), cutm(full["key"]), cutm(full["bbox"]), cutm(full["terms"]),
                                cutm(full["bf16"]), cutm(full["norm2"]), cutm(full["err"]), n_terms=n_terms, stream=st)
        if event is not None:
            eng.rescore_after(event)
        eng.set_images(img["emb"], img["key"], img.get("bbox"), img.get("terms"))

    def prefetch(self, img, chk):
        """Streaming use (one load + run per batch of tables): start the upload of the NEXT step's host shards -- pinned
        torch tensors -- into device staging buffers on a side stream, behind whatever the GPU is computing; the next
        load() given the same objects picks the staged copies up instead of uploading.  Two sets of staging buffers
        alternate, because the step in flight still reads the set its own load() was given."""
        import torch
        if self.device is None or torch.device(self.device).type != "cuda":
            return
        if self._pf_stream is None:
            self._pf_stream = torch.cuda.Stream(device=self.device)
            self._pf_parity = 0
        self._pf_parity ^= 1
        staged = []
        self._pf_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self._pf_stream):
            for side, d in (("img", img), ("chk", chk)):
                o = {}
                for f in self.FIELDS:
                    x = self._as_tensor(d.get(f))
                    if x is None:
                        o[f] = None
                        continue
                    buf = self._buffer(("prefetch", self._pf_parity, side, f), tuple(x.shape), x.dtype, x)
                    buf.copy_(x, non_blocking=True)
                    o[f] = buf
                staged.append(o)
            ev = torch.cuda.Event()
            ev.record(self._pf_stream)
        self._prefetched = (id(img), id(chk), staged[0], staged[1], ev)

    def _take_prefetched(self, img, chk):
        pf = self._prefetched
        self._prefetched = None
        if pf is None or pf[0] != id(img) or pf[1] != id(chk):
            return img, chk
        pf[4].synchronize()  # (the copy had the previous step's kernels to finish behind)
        return pf[2], pf[3]

    def load(self, img, chk, *, N: int, M: int, n_terms: int = 0):
        """img: image rows slab_range(N, world, rank); chk: chunk rows shard_range(M, world, rank) -- dicts of
        emb / key / bbox / terms (CUDA or pinned host torch tensors, or numpy arrays)."""
        img, chk = self._take_prefetched(img, chk)
        self.N, self.M = int(N), int(M)
        self.mode = self._mode(N) if self.world > 1 else "rows"
        if self.world > 1 and self.mode == "rows":
            self._load_rows(img, chk, N, M, n_terms)
            return
        if self.world > 1:
            img = self._replicate("img", img, N, slab_size(N, self.world))
            chk = self._replicate("chk", chk, M, -(-M // self.world) if M else 1)
        self.eng.set_images(img["emb"], img["key"], img.get("bbox"), img.get("terms"))
        self.eng.set_chunks(chk["emb"], chk["key"], chk.get("bbox"), chk.get("terms"), n_terms=n_terms, col_offset=0)

    # -- one step --------------------------------------------------------------------------------------
    def run(self, *, schemas, k_values, mrr_cutoff=100, weak_weight=(0.0, 0.0), kprime=0, host_outputs=False,
            candidates="all", path="auto", eps_scale=0.0, pipeline_rows=0):
        eng = self.eng
        kw = dict(candidates=candidates, k_values=k_values, mrr_cutoff=mrr_cutoff, weak_weight=weak_weight,
                  kprime=kprime, path=path, eps_scale=eps_scale)
        out_kw = dict(want=("topk", "pairs", "sums"), device_outputs=not host_outputs, pinned_outputs=host_outputs)
        if self.world == 1:
            r = eng.run(schemas, pipeline_rows=pipeline_rows, **out_kw, **kw)
            r["topk_row0"] = 0
        else:
            import time
            import torch
            dist, G, N, M = self.dist, self.world, self.N, self.M
            marks = [("start", time.perf_counter())]

            def mark(name):
                marks.append((name, time.perf_counter()))
            row0, row1 = slab_range(N, G, self.rank)
            mode = self.mode
            if mode == "rows":  # the engine holds this rank's slab only
                r = eng.run(schemas, pipeline_rows=pipeline_rows, **out_kw, **kw)
                mark("fused + rescore slab")
            else:
                lo, hi = shard_range(M, G, self.rank)
                eng.fused_pass(schemas, shard=(lo, hi - lo), k_values=k_values, mrr_cutoff=mrr_cutoff,
                               weak_weight=weak_weight, kprime=kprime, n_ranks=G)
                mark("fused_pass")
                per = slab_size(N, G)
                stride = torch.tensor([eng.list_stride()], dtype=torch.int32, device=self.device)
                dist.all_reduce(stride, op=self.MAX)
                stride = int(stride.item())
                keys, count, tau = eng.export_lists(G, per, stride)
                rk, rc, rt = torch.empty_like(keys), torch.empty_like(count), torch.empty_like(tau)
                dist.all_to_all_single(rk, keys)
                dist.all_to_all_single(rc, count)
                dist.all_to_all_single(rt, tau)
                mark("list exchange")
                r = eng.run(schemas, slab=(row0, row1 - row0), imported=(rk, rc, rt), **out_kw, **kw)
                mark("rescore slab")
            r["topk_row0"], r["contraction"] = row0, mode
            S, nk = r["hits"].shape
            st = r["stats"]
            packed = torch.tensor(np.concatenate([r["hits"].reshape(-1).astype(np.float64), r["rr_sum"],
                                                  [r["sim_sum"], float(r["num_pairs"]), float(st["rows_rescanned"]),
                                                   float(st["candidates_rescored"])]]),
                                  dtype=torch.float64, device=self.device)
            dist.all_reduce(packed)
            packed = packed.cpu().numpy()
            mark("metric sums")
            r["hits"] = np.rint(packed[:S * nk]).astype(np.int64).reshape(S, nk)
            r["rr_sum"], r["sim_sum"], r["num_pairs"] = packed[S * nk:S * nk + S], float(packed[-4]), int(round(packed[-3]))
            st["rows_rescanned"], st["candidates_rescored"] = int(round(packed[-2])), int(round(packed[-1]))  # whole job
            r["phases_ms"] = {b[0]: round(1e3 * (b[1] - a[1]), 2) for a, b in zip(marks, marks[1:])}
        r["metrics"] = metrics_from_sums(r["hits"], r["rr_sum"], r["sim_sum"], r["num_pairs"])
        r["d2h_bytes"] = sum(r[k].nbytes for k in ("topk_idx", "topk_score", "pair_rank", "pair_sim")) + 200 \
            if host_outputs else 0
        return r


class AllGatherScorer:
    """Fully sharded variant (module docstring): chunks never leave their rank, images replicated."""

    def __init__(self, engine, world: int = 1, rank: int = 0, device=None, dist=None):
        self.eng, self.world, self.rank, self.device = engine, world, rank, device
        if dist is None and world > 1:
            import torch.distributed as dist
        self.dist = dist
        self._pinned = {}
        self.MAX = getattr(getattr(dist, "ReduceOp", None), "MAX", "max") if dist is not None else "max"

    def run(self, *, schemas, k_values, mrr_cutoff=100, weak_weight=(0.0, 0.0), kprime=0, host_outputs=False,
            candidates="all", path="auto"):
        eng = self.eng
        kw = dict(candidates=candidates, k_values=k_values, mrr_cutoff=mrr_cutoff, weak_weight=weak_weight,
                  kprime=kprime, path=path)
        if self.world == 1:
            r = eng.run(schemas, want=("topk", "pairs", "sums"), device_outputs=not host_outputs,
                        pinned_outputs=host_outputs, **kw)
            r["metrics"] = metrics_from_sums(r["hits"], r["rr_sum"], r["sim_sum"], r["num_pairs"])
            r["d2h_bytes"] = sum(r[k].nbytes for k in ("topk_idx", "topk_score", "pair_rank", "pair_sim")) + 200 \
                if host_outputs else 0
            return r
        import time
        import torch
        dist = self.dist
        G = self.world
        marks = []

        def mark(name):
            torch.cuda.synchronize() if torch.cuda.is_available() else None
            marks.append((name, time.perf_counter()))
        mark("start")
        # 0. the depth K' of the candidate lists is shared by the ranks: every rank keeps its share, the ranks
        #    agree on a per-row threshold, and each re-scores only what lies above it (~K'/G entries per row)
        ses = eng.sharded_session(schemas, k_values=k_values, mrr_cutoff=mrr_cutoff, weak_weight=weak_weight,
                                  kprime=kprime, n_ranks=G)
        tau = ses.fused_pass()
        mark("fused_pass")
        dist.all_reduce(tau, op=self.MAX)
        eps = torch.tensor([ses.chunk_err_max()], dtype=torch.float32, device=tau.device)
        dist.all_reduce(eps, op=self.MAX)
        mark("tau exchange")
        r, cert = ses.rescore_pass(tau, float(eps.item()))
        mark("rescore_pass")
        dist.all_reduce(cert)
        # a row is certified when, over all ranks, at least kneed entries lie provably above everything left out
        bad = (cert < ses.kneed).any(dim=0) & torch.isfinite(tau)
        rows = torch.nonzero(bad).to(torch.int32).flatten().contiguous()
        ses.rescan_rows(rows)
        mark("certificate + rescan")
        S, N, K = r["topk_idx"].shape
        # 1. global top-K lists, query-sharded: rank g merges the lists of query slab g (one all-to-all; every
        #    rank receives 1/G of what an all-gather would deliver) and owns the result for those queries
        slab = -(-N // G)
        row0 = min(N, self.rank * slab)
        rows_here = max(0, min(N, row0 + slab) - row0)
        if hasattr(dist, "all_to_all_single"):
            pad = slab * G - N
            ti, ts = r["topk_idx"], r["topk_score"]
            if pad:
                ti = torch.cat([ti, ti.new_full((S, pad, K), -1)], dim=1)
                ts = torch.cat([ts, ts.new_full((S, pad, K), float("-inf"))], dim=1)
            send_i = ti.view(S, G, slab, K).permute(1, 0, 2, 3).contiguous()
            send_s = ts.view(S, G, slab, K).permute(1, 0, 2, 3).contiguous()
            recv_i, recv_s = torch.empty_like(send_i), torch.empty_like(send_s)
            dist.all_to_all_single(recv_i, send_i)
            dist.all_to_all_single(recv_s, send_s)
            m_idx, m_score = eng.merge_topk(recv_i.view(G, S * slab, K), recv_s.view(G, S * slab, K))
            m_idx, m_score = m_idx.view(S, slab, K)[:, :rows_here], m_score.view(S, slab, K)[:, :rows_here]
        else:  # backends without all-to-all: all-gather, merge everything, keep the slab
            gi = [torch.empty_like(r["topk_idx"]) for _ in range(G)]
            gs = [torch.empty_like(r["topk_score"]) for _ in range(G)]
            dist.all_gather(gi, r["topk_idx"].contiguous())
            dist.all_gather(gs, r["topk_score"].contiguous())
            m_idx, m_score = eng.merge_topk(torch.stack(gi).view(G, S * N, K), torch.stack(gs).view(G, S * N, K))
            m_idx = m_idx.view(S, N, K)[:, row0:row0 + rows_here]
            m_score = m_score.view(S, N, K)[:, row0:row0 + rows_here]
        mark("top-K gather + merge")
        # 2. global rank of every true pair
        off, pc = eng.pairs_device()
        P = int(pc.shape[0])
        sizes = [torch.zeros(1, dtype=torch.int64, device=pc.device) for _ in range(G)]
        dist.all_gather(sizes, torch.tensor([P], dtype=torch.int64, device=pc.device))
        sizes = [int(s.item()) for s in sizes]
        Pmax = max(max(sizes), 1)
        q_img = torch.zeros(Pmax, dtype=torch.int64, device=pc.device)
        q_chk = torch.zeros(Pmax, dtype=torch.int64, device=pc.device)
        q_sc = torch.full((S, Pmax), float("inf"), dtype=torch.float64, device=pc.device)
        if P:
            q_img[:P] = torch.repeat_interleave(torch.arange(N, device=pc.device), off[1:] - off[:-1])
            q_chk[:P] = pc
            q_sc[:, :P] = r["pair_score"]
        a_img = [torch.empty_like(q_img) for _ in range(G)]
        a_chk = [torch.empty_like(q_chk) for _ in range(G)]
        a_sc = [torch.empty_like(q_sc) for _ in range(G)]
        dist.all_gather(a_img, q_img)
        dist.all_gather(a_chk, q_chk)
        dist.all_gather(a_sc, q_sc)
        counts = eng.count_beating(r["deep_idx"], r["deep_score"], torch.cat(a_img), torch.cat(a_chk),
                                   torch.cat(a_sc, dim=1).contiguous())
        dist.all_reduce(counts)
        kneed = r["deep_idx"].shape[2]
        mine = counts[:, self.rank * Pmax:self.rank * Pmax + P]
        pair_rank = torch.where(mine < kneed, mine + 1, torch.zeros_like(mine)).to(torch.int32).contiguous()
        mark("pair ranks")
        # 3. metric sums
        hits, rr, sim = eng.reduce_metrics(pair_rank, r["pair_sim"], k_values, mrr_cutoff)
        packed = torch.tensor(np.concatenate([hits.reshape(-1).astype(np.float64), rr, [sim, float(P)]]),
                              dtype=torch.float64, device=pc.device)
        dist.all_reduce(packed)
        packed = packed.cpu().numpy()
        mark("metric sums")
        nk = len(k_values)
        hits_g = np.rint(packed[:S * nk]).astype(np.int64).reshape(S, nk)
        rr_g, sim_g, P_g = packed[S * nk:S * nk + S], packed[-2], int(round(packed[-1]))
        out = dict(topk_idx=m_idx, topk_score=m_score, pair_rank=pair_rank, pair_sim=r["pair_sim"], hits=hits_g,
                   rr_sum=rr_g, sim_sum=float(sim_g), num_pairs=P_g, stats=r["stats"],
                   metrics=metrics_from_sums(hits_g, rr_g, sim_g, P_g), d2h_bytes=0, topk_row0=row0,
                   phases_ms={b[0]: round(1e3 * (b[1] - a[1]), 2) for a, b in zip(marks, marks[1:])})
        if host_outputs:  # page-locked host buffers, reused across steps
            for k in ("topk_idx", "topk_score", "pair_rank", "pair_sim"):
                t = out[k]
                if not t.is_cuda:
                    out[k] = t.numpy()
                    continue
                buf = self._pinned.get(k)
                if buf is None or buf.shape != t.shape or buf.dtype != t.dtype:
                    buf = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                    self._pinned[k] = buf
                buf.copy_(t, non_blocking=True)
                out[k] = buf.numpy()
            if self._pinned:
                torch.cuda.synchronize()
            out["d2h_bytes"] = sum(out[k].nbytes for k in ("topk_idx", "topk_score", "pair_rank", "pair_sim")) + 200
        return out
