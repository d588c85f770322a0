"""Multi-GPU plumbing: one process per GPU, chunks sharded by contiguous column range,
images replicated.  torch.distributed (NCCL over NVLink) carries the
three exchanges of a step; every merge/count/reduction on the data is a kernel of the
CUDA library:

  0. all-reduce(max) of the per-row list thresholds (and of the chunk rounding-error bound), then
     all-reduce(sum) of the per-row certificate counts            (mmalign_fused_pass / rescore_pass)
  1. all-to-all of each rank's exact top-K lists  -> mmalign_merge_topk: rank g merges and owns the lists
     of query slab g (`topk_row0`, ceil(N/G) rows)
  2. all-gather of each rank's true pairs (image, chunk, score per schema)
        -> mmalign_count_beating against the rank's own exact lists
        -> all-reduce(sum) of the counts            (global rank of every true pair)
  3. all-reduce(sum) of the metric sums from mmalign_reduce_metrics.

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


class ShardedScorer:
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
