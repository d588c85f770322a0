"""The multi-rank paths with the REAL kernels on one GPU: the ranks run as threads, each with its own
engine and shard; a thread-barrier stand-in for torch.distributed carries the exchanges (the kernels of
the ranks never wait on each other, only the host threads do).  Results must equal the single-process
oracle.  The real NCCL run is tests/test_gpu_nccl.py (needs 2 GPUs)."""
import importlib
import threading

import numpy as np
import pytest
import torch

from conftest import PKG_NAME

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300, method="thread")]

ALL4 = ["vanilla_clip", "clip_lexical", "clip_positional", "clip_combined"]


class ThreadDist:
    def __init__(self, world):
        self.world, self.slots, self.bar = world, {}, threading.Barrier(world)

    def bind(self, rank):
        d = self
        class _D:
            def all_gather(self, out, t):
                torch.cuda.synchronize()
                d.slots[rank] = t
                d.bar.wait()
                for r in range(d.world):
                    out[r].copy_(d.slots[r])
                torch.cuda.synchronize()
                d.bar.wait()

            def all_gather_into_tensor(self, out, t):
                n = t.shape[0]
                self.all_gather([out[r * n:(r + 1) * n] for r in range(d.world)], t.clone())

            def all_to_all_single(self, out, inp):
                torch.cuda.synchronize()
                d.slots[rank] = inp
                d.bar.wait()
                n = inp.shape[0] // d.world
                for r in range(d.world):
                    out[r * n:(r + 1) * n].copy_(d.slots[r][rank * n:(rank + 1) * n])
                torch.cuda.synchronize()
                d.bar.wait()

            def all_reduce(self, t, op="sum"):
                torch.cuda.synchronize()
                d.slots[rank] = t.clone()
                d.bar.wait()
                if op == "max":
                    s = torch.stack([d.slots[r] for r in range(d.world)]).amax(dim=0).to(t.dtype)
                else:
                    s = sum(d.slots[r] for r in range(d.world))
                torch.cuda.synchronize()
                d.bar.wait()
                t.copy_(s)
        return _D()


def _run_ranks(world, rank_main):
    td = ThreadDist(world)
    results, errors = {}, []

    def guarded(rank):
        try:
            torch.cuda.set_device(0)
            results[rank] = rank_main(rank, td.bind(rank))
        except BaseException as e:  # noqa: BLE001
            errors.append(e)
            td.bar.abort()
    threads = [threading.Thread(target=guarded, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    return results


@pytest.mark.parametrize("contraction", ["columns", "rows"])
@pytest.mark.parametrize("N,M,D,world,kprime", [(300, 2003, 128, 2, 0), (201, 5000, 512, 4, 0), (130, 4000, 64, 8, 0),
                                                (700, 3000, 64, 3, 0), (5, 3, 64, 4, 0), (400, 6000, 64, 2, -1)])
def test_column_sharded_contraction_row_sharded_rescoring(pkg, oracle, synthetic, N, M, D, world, kprime, contraction):
    """ShardedScorer: each rank ingests 1/G of both tables, contracts against its chunk columns, exchanges the
    candidate lists and owns the exact results of its query slab.  D=64 makes near-ties and rescans likely;
    kprime=-1 stands for eps_scale=100: an inflated error bound sends uncertified rows through the exact scan."""
    distributed = importlib.import_module(PKG_NAME + ".distributed")
    img, chk, _ = synthetic.make_numpy(N, M, D, T=512, seed=43)
    ks, cutoff, lam = (1, 5, 10, 20), 100, (0.3, 0.2)
    eps_scale, kprime = (100.0, 0) if kprime < 0 else (0.0, kprime)
    cut = lambda d, lo, hi: {k: (v[lo:hi] if v is not None else None) for k, v in d.items()}

    def rank_main(rank, dist):
        eng = pkg.AlignmentEngine(0)
        sc = distributed.ShardedScorer(eng, world, rank, torch.device("cuda", 0), dist=dist, contraction=contraction)
        sc.load(cut(img, *distributed.slab_range(N, world, rank)), cut(chk, *distributed.shard_range(M, world, rank)),
                N=N, M=M, n_terms=512)
        r = sc.run(schemas=ALL4, k_values=ks, mrr_cutoff=cutoff, weak_weight=lam, host_outputs=True, kprime=kprime,
                   eps_scale=eps_scale)
        r = {k: (np.array(v) if isinstance(v, np.ndarray) else v) for k, v in r.items()}
        eng.close()
        return r
    results = _run_ranks(world, rank_main)
    o = oracle.evaluate(img, chk, T=512, schema_mask=15, candidates="all", lam=(lam[0], lam[1], lam[0] + lam[1]),
                        kmax=max(ks), cutoff=cutoff)
    rescanned = 0
    for rank in range(world):
        r = results[rank]
        q0, q1 = distributed.slab_range(N, world, rank)
        assert r["topk_row0"] == q0 and r["topk_idx"].shape[1] == q1 - q0
        assert np.array_equal(r["topk_idx"], o["topk_idx"][:, q0:q1]) and np.array_equal(r["topk_score"], o["topk_score"][:, q0:q1])
        p0, p1 = o["pair_offsets"][q0], o["pair_offsets"][q1]
        assert np.array_equal(r["pair_rank"], o["pair_rank"][:, p0:p1])
        assert np.array_equal(r["pair_sim"], o["pair_sim"][p0:p1])
        assert r["num_pairs"] == len(o["pair_chunk"])
        for si in range(4):
            for q, k in enumerate(ks):
                assert r["hits"][si, q] == np.count_nonzero((o["pair_rank"][si] >= 1) & (o["pair_rank"][si] <= k))
        assert r["sim_sum"] == pytest.approx(float(o["pair_sim"].sum()), rel=1e-12)
        rescanned += r["stats"]["rows_rescanned"]
    if eps_scale:
        assert rescanned > 0


@pytest.mark.parametrize("N,M,D,world", [(300, 2003, 128, 2), (201, 5000, 512, 4), (130, 4000, 64, 8)])
def test_fully_sharded_equals_oracle(pkg, oracle, synthetic, N, M, D, world):
    distributed = importlib.import_module(PKG_NAME + ".distributed")
    img, chk, _ = synthetic.make_numpy(N, M, D, T=512, seed=41)
    ks, cutoff, lam = (1, 5, 10, 20), 100, (0.3, 0.2)
    td = ThreadDist(world)
    results, errors = {}, []

    def rank_main(rank):
        try:
            torch.cuda.set_device(0)
            lo, hi = distributed.shard_range(M, world, rank)
            eng = pkg.AlignmentEngine(0)
            eng.set_images(img["emb"], img["key"], img["bbox"], None)
            eng.set_chunks(chk["emb"][lo:hi], chk["key"][lo:hi], chk["bbox"][lo:hi], chk["terms"][lo:hi], n_terms=512,
                           col_offset=lo)
            sc = distributed.AllGatherScorer(eng, world, rank, torch.device("cuda", 0), dist=td.bind(rank))
            results[rank] = (lo, hi, sc.run(schemas=ALL4, k_values=ks, mrr_cutoff=cutoff, weak_weight=lam, host_outputs=True))
            eng.close()
        except BaseException as e:  # noqa: BLE001
            errors.append(e)
            td.bar.abort()
    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    o = oracle.evaluate(img, chk, T=512, schema_mask=15, candidates="all", lam=(lam[0], lam[1], lam[0] + lam[1]),
                        kmax=max(ks), cutoff=cutoff)
    for rank in range(world):
        lo, hi, r = results[rank]
        q0 = r["topk_row0"]
        q1 = q0 + r["topk_idx"].shape[1]
        assert q1 - q0 == max(0, min(N, q0 + -(-N // world)) - q0)
        assert np.array_equal(r["topk_idx"], o["topk_idx"][:, q0:q1]) and np.array_equal(r["topk_score"], o["topk_score"][:, q0:q1])
        mine = (o["pair_chunk"] >= lo) & (o["pair_chunk"] < hi)
        assert np.array_equal(r["pair_rank"], o["pair_rank"][:, mine])
        assert np.array_equal(r["pair_sim"], o["pair_sim"][mine])
        assert r["num_pairs"] == len(o["pair_chunk"])
        for si in range(4):
            for q, k in enumerate(ks):
                assert r["hits"][si, q] == np.count_nonzero((o["pair_rank"][si] >= 1) & (o["pair_rank"][si] <= k))


@pytest.mark.parametrize("world,contraction", [(1, "rows"), (2, "rows"), (3, "columns")])
def test_prefetched_uploads_give_the_same_results(pkg, oracle, synthetic, world, contraction):
    """Streaming use: while step s computes, ShardedScorer.prefetch copies the pinned host shards of step s+1 into device
    staging buffers (two sets alternate); the next load() picks them up.  Three steps over two different corpora
    (A, B, A): every step's results are the oracle's for ITS corpus."""
    distributed = importlib.import_module(PKG_NAME + ".distributed")
    N, M, D = 600, 5000, 128
    ks, cutoff, lam = (1, 5, 10, 20), 100, (0.3, 0.2)
    corpora = [synthetic.make_numpy(N, M, D, T=512, seed=s)[:2] for s in (81, 82)]
    want = [oracle.evaluate(i, c, T=512, schema_mask=15, candidates="all", lam=(lam[0], lam[1], lam[0] + lam[1]),
                            kmax=max(ks), cutoff=cutoff) for i, c in corpora]

    def pinned(d, lo, hi):
        out = {}
        for k, v in d.items():
            if v is None:
                out[k] = None
                continue
            v = np.ascontiguousarray(v[lo:hi])
            out[k] = torch.from_numpy(v.view(np.int64) if v.dtype == np.uint64 else v).pin_memory()
        return out

    def rank_main(rank, dist):
        eng = pkg.AlignmentEngine(0)
        sc = distributed.ShardedScorer(eng, world, rank, torch.device("cuda", 0), dist=dist if world > 1 else None,
                                       contraction=contraction)
        host = [(pinned(i, *distributed.slab_range(N, world, rank)), pinned(c, *distributed.shard_range(M, world, rank)))
                for i, c in corpora]
        got = []
        for step, which in enumerate((0, 1, 0)):
            assert (sc._prefetched is not None) == (step > 0)
            sc.load(*host[which], N=N, M=M, n_terms=512)
            assert sc._prefetched is None  # (load consumed what the previous step staged)
            nxt = (1, 0, 1)[step]
            sc.prefetch(*host[nxt])
            r = sc.run(schemas=ALL4, k_values=ks, mrr_cutoff=cutoff, weak_weight=lam, host_outputs=True)
            got.append({k: (np.array(v) if isinstance(v, np.ndarray) else v) for k, v in r.items()})
        eng.close()
        return got
    results = _run_ranks(world, rank_main)
    for rank in range(world):
        q0, q1 = distributed.slab_range(N, world, rank)
        for step, which in enumerate((0, 1, 0)):
            r, o = results[rank][step], want[which]
            p0, p1 = o["pair_offsets"][q0], o["pair_offsets"][q1]
            assert np.array_equal(r["topk_idx"], o["topk_idx"][:, q0:q1]) and np.array_equal(r["topk_score"], o["topk_score"][:, q0:q1])
            assert np.array_equal(r["pair_rank"], o["pair_rank"][:, p0:p1]) and np.array_equal(r["pair_sim"], o["pair_sim"][p0:p1])
