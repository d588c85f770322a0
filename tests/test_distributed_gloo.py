"""world_size-2 gloo run of the multi-GPU plumbing (distributed.ShardedScorer) on the CPU.

The CUDA kernels cannot run here, so a stand-in engine answers the four library calls
(run / merge_topk / count_beating / reduce_metrics) from the CPU oracle; what is under test is
the host logic: shard ranges, the three exchanges, padding of ragged pair lists, global ranks,
metric reduction.  The merged result must equal the oracle's single-process result."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG_NAME

MASK = 15
LAM = (0.3, 0.2, 0.5)
KS = (1, 5, 10, 20)
CUTOFF = 30


class OracleEngine:
    def __init__(self, oracle, img, chk_shard, T, col_offset):
        self.o, self.img, self.chk, self.T, self.off0 = oracle, img, chk_shard, T, col_offset
        self.N = len(img["key"])

    def sharded_session(self, schemas, *, k_values, mrr_cutoff, weak_weight, kprime, n_ranks):
        eng = self

        class Session:
            kneed = max(max(k_values), mrr_cutoff)

            def fused_pass(self):  # the stand-in keeps every local column: complete above -inf
                return torch.full((eng.N,), float("-inf"))

            def chunk_err_max(self):
                return 0.0

            def rescore_pass(self, tau, eps):
                r = eng.run(schemas, k_values=k_values, mrr_cutoff=mrr_cutoff, weak_weight=weak_weight)
                return r, torch.full((4, eng.N), self.kneed, dtype=torch.int32)

            def rescan_rows(self, rows):
                assert rows.numel() == 0
        return Session()

    def run(self, schemas, *, k_values, mrr_cutoff, weak_weight, **_):
        kmax, kneed = max(k_values), max(max(k_values), mrr_cutoff)
        lam = (weak_weight[0], weak_weight[1], weak_weight[0] + weak_weight[1])
        a = self.o.evaluate(self.img, self.chk, T=self.T, schema_mask=MASK, candidates="all", lam=lam, kmax=kneed, cutoff=kneed)
        s = self.o.evaluate(self.img, self.chk, T=self.T, schema_mask=MASK, candidates="same_page", lam=lam, kmax=64, cutoff=64)
        off, pc = a["pair_offsets"], a["pair_chunk"]
        P = len(pc)
        score = np.zeros((4, P))
        for si in range(4):
            for i in range(self.N):
                lut = dict(zip(s["topk_idx"][si, i].tolist(), s["topk_score"][si, i].tolist()))
                for p in range(off[i], off[i + 1]):
                    score[si, p] = lut[pc[p]]
        g = lambda x: np.where(x >= 0, x + self.off0, -1)
        t = torch.from_numpy
        self._pairs = (t(off), t(pc + self.off0))
        return dict(topk_idx=t(g(a["topk_idx"][:, :, :kmax]).copy()), topk_score=t(a["topk_score"][:, :, :kmax].copy()),
                    pair_rank=t(a["pair_rank"]), pair_sim=t(a["pair_sim"]), pair_score=t(score),
                    deep_idx=t(g(a["topk_idx"])), deep_score=t(a["topk_score"]), stats={})

    def pairs_device(self):
        return self._pairs

    def merge_topk(self, gi, gs):
        G, L, K = gi.shape
        i, s = gi.numpy().transpose(1, 0, 2).reshape(L, G * K), gs.numpy().transpose(1, 0, 2).reshape(L, G * K)
        oi, os_ = np.full((L, K), -1, np.int64), np.full((L, K), -np.inf)
        for l in range(L):
            ok = i[l] >= 0
            o = np.lexsort((i[l][ok], -s[l][ok]))[:K]
            oi[l, :len(o)], os_[l, :len(o)] = i[l][ok][o], s[l][ok][o]
        return torch.from_numpy(oi), torch.from_numpy(os_)

    def count_beating(self, deep_idx, deep_score, q_img, q_chk, q_sc):
        di, ds = deep_idx.numpy(), deep_score.numpy()
        S, nq = q_sc.shape
        out = np.zeros((S, nq), np.int32)
        for s in range(S):
            for q in range(nq):
                i, j, sc = int(q_img[q]), int(q_chk[q]), float(q_sc[s, q])
                beats = (di[s, i] >= 0) & ((ds[s, i] > sc) | ((ds[s, i] == sc) & (di[s, i] < j)))
                out[s, q] = int(beats.sum())
        return torch.from_numpy(out)

    def reduce_metrics(self, pair_rank, pair_sim, k_values, mrr_cutoff):
        r = pair_rank.numpy()
        hits = np.array([[np.count_nonzero((r[s] >= 1) & (r[s] <= k)) for k in k_values] for s in range(r.shape[0])], np.int64)
        rr = np.array([sum(1.0 / x for x in r[s].tolist() if 1 <= x <= mrr_cutoff) for s in range(r.shape[0])])
        return hits, rr, float(pair_sim.numpy().sum())


def worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle
    synthetic = importlib.import_module(PKG_NAME + ".synthetic")
    distributed = importlib.import_module(PKG_NAME + ".distributed")
    img, chk, _ = synthetic.make_numpy(24, 203, 64, T=64, seed=31)
    lo, hi = distributed.shard_range(203, world, rank)
    shard = {k: (v[lo:hi] if v is not None else None) for k, v in chk.items()}
    eng = OracleEngine(oracle, img, shard, 64, lo)
    sc = distributed.AllGatherScorer(eng, world, rank, None, dist=dist)
    r = sc.run(schemas=None, k_values=KS, mrr_cutoff=CUTOFF, weak_weight=LAM[:2], host_outputs=True)
    # gather each rank's pair ranks for the global check
    q.put((rank, lo, r["topk_idx"], r["topk_score"], r["pair_rank"], r["hits"], r["rr_sum"], r["sim_sum"],
           r["num_pairs"], r["metrics"]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_merge_equals_single_process(oracle, synthetic):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=180) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    img, chk, _ = synthetic.make_numpy(24, 203, 64, T=64, seed=31)
    o = oracle.evaluate(img, chk, T=64, schema_mask=MASK, candidates="all", lam=LAM, kmax=max(KS), cutoff=CUTOFF)
    for rank, lo, ti, ts, pr, hits, rr, sim, P, metrics in got:
        q0 = rank * 12  # query slab of the rank: ceil(24 / 2) rows
        assert np.array_equal(ti, o["topk_idx"][:, q0:q0 + 12]) and np.array_equal(ts, o["topk_score"][:, q0:q0 + 12])
        assert P == len(o["pair_chunk"])
        for si in range(4):
            for qi, k in enumerate(KS):
                assert hits[si, qi] == np.count_nonzero((o["pair_rank"][si] >= 1) & (o["pair_rank"][si] <= k))
            want_rr = sum(1.0 / x for x in o["pair_rank"][si].tolist() if 1 <= x <= CUTOFF)
            assert rr[si] == pytest.approx(want_rr, rel=1e-12)
        assert sim == pytest.approx(float(o["pair_sim"].sum()), rel=1e-12)
        assert metrics["num_pairs"] == P
    # each rank's slice of the true pairs carries the GLOBAL rank
    hi0 = got[1][1]
    mine0 = o["pair_chunk"] < hi0
    assert np.array_equal(got[0][4], o["pair_rank"][:, mine0]) and np.array_equal(got[1][4], o["pair_rank"][:, ~mine0])


def test_shard_range(pkg):
    d = importlib.import_module(PKG_NAME + ".distributed")
    for M in (0, 1, 7, 8, 1000003):
        for G in (1, 2, 4, 8):
            r = [d.shard_range(M, G, k) for k in range(G)]
            assert r[0][0] == 0 and r[-1][1] == M and all(a[1] == b[0] for a, b in zip(r, r[1:]))


# ------------------------------------------------------------------------------------------------------------------
# ShardedScorer (default exchange): ingest all-gather with ragged shards, candidate-list all-to-all, slab ownership
# ------------------------------------------------------------------------------------------------------------------
class OracleSlabEngine:
    """Stand-in for the engine calls ShardedScorer makes.  The 'fused pass' keeps EVERY column of the rank's shard
    (complete above -inf), so the exchange must deliver, for every row of a slab, every global column exactly once;
    the stand-in for mmalign_rescore_slab checks that and answers from the oracle."""

    def __init__(self, oracle, T):
        self.o, self.T, self.device = oracle, T, 0

    @staticmethod
    def _np(x, dt):
        return None if x is None else np.ascontiguousarray(x.numpy() if hasattr(x, "numpy") else x).view(dt) \
            if dt == np.uint64 else (None if x is None else np.ascontiguousarray(x.numpy() if hasattr(x, "numpy") else x, dtype=dt))

    def set_images(self, emb, key, bbox=None, terms=None):
        self.img = dict(emb=self._np(emb, np.float32), key=self._np(key, np.uint64), bbox=self._np(bbox, np.float64), terms=None)
        self.N = len(self.img["key"])

    def set_chunks(self, emb, key, bbox=None, terms=None, n_terms=0, col_offset=0):
        assert col_offset == 0  # every rank holds the whole table
        self.chk = dict(emb=self._np(emb, np.float32), key=self._np(key, np.uint64), bbox=self._np(bbox, np.float64),
                        terms=self._np(terms, np.uint64))
        self.M = len(self.chk["key"])

    # query-row layout: the prepared chunk table arrives whole, the engine holds this rank's image slab only
    def prep_rows(self, emb, bf16, norm2, err, stream=None):
        bf16.copy_(emb.to(torch.bfloat16).view(torch.int16))   # (stand-in: only the plumbing is under test)
        norm2.copy_((emb * emb).sum(1))
        err.zero_()

    def set_chunks_prepared(self, emb, key, bbox, terms, bf16, norm2, err, n_terms=0, col_offset=0, stream=None):
        assert torch.equal(bf16, emb.to(torch.bfloat16).view(torch.int16))   # every rank's K0 output arrived in place
        assert torch.allclose(norm2, (emb * emb).sum(1))
        self.set_chunks(emb, key, bbox, terms, n_terms, col_offset)

    def rescore_after(self, event):
        raise AssertionError("no side stream on the CPU")

    def fused_pass(self, schemas, *, shard, **_):
        self.shard = shard

    def list_stride(self):
        return max(1, self.shard[1])

    def export_lists(self, n_dest, slab_rows, stride):
        lo, n = self.shard
        keys = torch.zeros((n_dest, slab_rows, stride), dtype=torch.int64)
        keys[:, :, :n] = torch.arange(lo, lo + n)  # score bits 0 in the high half
        count = torch.full((n_dest, slab_rows), n, dtype=torch.int32)
        count.view(-1)[self.N:] = 0                                     # padding rows beyond N
        tau = torch.full((n_dest, slab_rows), float("-inf"))
        return keys, count, tau

    def run(self, schemas, *, k_values, mrr_cutoff, weak_weight, slab=None, imported=None, **_):
        row0, rows = slab if slab is not None else (0, self.N)
        if imported is not None:  # column layout: every global column exactly once, from the rank that owns it
            keys, count, tau = imported
            for r in range(rows):
                cols = torch.cat([keys[g, r, :int(count[g, r])] for g in range(keys.shape[0])]).numpy() & 0xFFFFFFFF
                assert np.array_equal(np.sort(cols), np.arange(self.M)), (row0, r)
            assert torch.isinf(tau[:, :rows]).all()
        lam = (weak_weight[0], weak_weight[1], weak_weight[0] + weak_weight[1])
        kmax, kneed = max(k_values), max(max(k_values), mrr_cutoff)
        o = self.o.evaluate(self.img, self.chk, T=self.T, schema_mask=MASK, candidates="all", lam=lam, kmax=kmax, cutoff=kneed)
        p0, p1 = o["pair_offsets"][row0], o["pair_offsets"][row0 + rows]
        pr, ps = o["pair_rank"][:, p0:p1], o["pair_sim"][p0:p1]
        hits = np.array([[np.count_nonzero((pr[s] >= 1) & (pr[s] <= k)) for k in k_values] for s in range(4)], np.int64)
        rr = np.array([sum(1.0 / x for x in pr[s].tolist() if 1 <= x <= mrr_cutoff) for s in range(4)])
        return dict(topk_idx=o["topk_idx"][:, row0:row0 + rows], topk_score=o["topk_score"][:, row0:row0 + rows],
                    pair_rank=pr, pair_sim=ps, hits=hits, rr_sum=rr, sim_sum=float(ps.sum()), num_pairs=int(p1 - p0),
                    stats=dict(rows_rescanned=0, candidates_rescored=0))


def slab_worker(rank, world, port, q, N, M, contraction):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle
    synthetic = importlib.import_module(PKG_NAME + ".synthetic")
    distributed = importlib.import_module(PKG_NAME + ".distributed")
    img, chk, _ = synthetic.make_numpy(N, M, 64, T=64, seed=33)
    cut = lambda d, lo, hi: {k: (v[lo:hi] if v is not None else None) for k, v in d.items()}
    sc = distributed.ShardedScorer(OracleSlabEngine(oracle, 64), world, rank, None, dist=dist, contraction=contraction)
    sc.load(cut(img, *distributed.slab_range(N, world, rank)), cut(chk, *distributed.shard_range(M, world, rank)),
            N=N, M=M, n_terms=64)
    r = sc.run(schemas=None, k_values=KS, mrr_cutoff=CUTOFF, weak_weight=LAM[:2], host_outputs=True)
    q.put((rank, r["topk_row0"], r["topk_idx"], r["topk_score"], r["pair_rank"], r["hits"], r["rr_sum"], r["sim_sum"],
           r["num_pairs"], r["metrics"]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,N,M,contraction", [(2, 300, 203, "columns"), (3, 130, 50, "auto"), (2, 300, 203, "rows")])
def test_slab_exchange_equals_single_process(oracle, synthetic, world, N, M, contraction):
    """N=300 over 2 ranks: slabs of 256 and 44 rows (whole 128-row blocks); N=130 over 3 ranks: the third slab is
    empty; M=203 / 50: ragged chunk shards, padded in the all-gather."""
    distributed = importlib.import_module(PKG_NAME + ".distributed")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=slab_worker, args=(r, world, port, q, N, M, contraction)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=240) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    img, chk, _ = synthetic.make_numpy(N, M, 64, T=64, seed=33)
    o = oracle.evaluate(img, chk, T=64, schema_mask=MASK, candidates="all", lam=LAM, kmax=max(KS), cutoff=CUTOFF)
    covered = 0
    for rank, row0, ti, ts, pr, hits, rr, sim, P, metrics in got:
        lo, hi = distributed.slab_range(N, world, rank)
        assert row0 == lo and ti.shape[1] == hi - lo
        assert np.array_equal(ti, o["topk_idx"][:, lo:hi]) and np.array_equal(ts, o["topk_score"][:, lo:hi])
        p0, p1 = o["pair_offsets"][lo], o["pair_offsets"][hi]
        assert np.array_equal(pr, o["pair_rank"][:, p0:p1])
        covered += hi - lo
        # the metric sums are the whole job's on every rank
        assert P == len(o["pair_chunk"]) and metrics["num_pairs"] == P
        for si in range(4):
            for qi, k in enumerate(KS):
                assert hits[si, qi] == np.count_nonzero((o["pair_rank"][si] >= 1) & (o["pair_rank"][si] <= k))
            want_rr = sum(1.0 / x for x in o["pair_rank"][si].tolist() if 1 <= x <= CUTOFF)
            assert rr[si] == pytest.approx(want_rr, rel=1e-12)
        assert sim == pytest.approx(float(o["pair_sim"].sum()), rel=1e-12)
    assert covered == N


def test_slab_range(pkg):
    d = importlib.import_module(PKG_NAME + ".distributed")
    for N in (0, 1, 127, 128, 129, 1000, 1000000):
        for G in (1, 2, 3, 8):
            r = [d.slab_range(N, G, k) for k in range(G)]
            assert r[0][0] == 0 and r[-1][1] == N and all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert all(a[0] % 128 == 0 or a[0] == N for a in r) and d.slab_size(N, G) * G >= N


def test_prefetch_is_a_device_feature(pkg):
    """ShardedScorer.prefetch stages the next step's pinned host shards on the GPU; without a CUDA device it does nothing
    and load() takes its inputs as they are (the GPU side: tests/test_gpu_multirank.py::test_prefetched_uploads_*)."""
    d = importlib.import_module(PKG_NAME + ".distributed")
    sc = d.ShardedScorer(engine=None, world=1, rank=0, device=None)
    img, chk = {"emb": np.zeros((2, 64), np.float32)}, {"emb": np.zeros((3, 64), np.float32)}
    sc.prefetch(img, chk)
    assert sc._prefetched is None
    assert sc._take_prefetched(img, chk) == (img, chk)
    other = {"emb": np.zeros((2, 64), np.float32)}
    sc._prefetched = (id(other), id(chk), "staged images", "staged chunks", None)   # staged for other objects: ignored, dropped
    assert sc._take_prefetched(img, chk) == (img, chk) and sc._prefetched is None
