"""AlignmentEngine: thin Python owner of one mmalign context (one per process and GPU).

Arrays may be numpy arrays (host) or torch tensors (host, pinned host or CUDA);
only their addresses cross the ABI.  Results come back as numpy arrays unless
`device_outputs=True`, in which case they are CUDA torch tensors.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Optional, Sequence

import numpy as np

from . import _native
from ._native import CAND, PATHS, SCHEMA_BITS

SCHEMAS = ["vanilla_clip", "clip_lexical", "clip_positional", "clip_combined"]


class MMAlignError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"mmalign error {code}: {msg}")
        self.code = code


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _prep(x, np_dtype, torch_dtype_name):
    """Returns (keepalive, address) of a contiguous array of the wanted dtype."""
    if x is None:
        return None, None
    if _is_torch(x):
        import torch
        td = getattr(torch, torch_dtype_name)
        if x.dtype != td:
            if torch_dtype_name == "int64" and x.dtype == torch.uint64:
                x = x.view(torch.int64)
            else:
                x = x.to(td)
        x = x.contiguous()
        return x, x.data_ptr()
    a = np.ascontiguousarray(x, dtype=np_dtype)
    return a, a.ctypes.data


def schema_mask(schemas: Iterable[str] | str | int) -> int:
    if isinstance(schemas, int):
        return schemas
    if isinstance(schemas, str):
        schemas = [schemas]
    m = 0
    for s in schemas:
        if s not in SCHEMA_BITS:
            raise ValueError(f"unknown schema {s!r}; expected one of {SCHEMAS}")
        m |= SCHEMA_BITS[s]
    return m


class AlignmentEngine:
    def __init__(self, device: int = 0):
        self._L = _native.load()
        self._ctx = C.c_void_p()
        rc = self._L.mmalign_create(C.byref(self._ctx), int(device))
        if rc != 0:
            raise MMAlignError(rc, self._L.mmalign_last_error(None).decode())
        self.device = int(device)
        self._keep = {}
        self._pinned = {}
        self.N = self.M = self.D = 0
        self.col_offset = 0

    # -- lifetime ---------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self._L.mmalign_destroy(self._ctx)
            self._ctx = C.c_void_p()
            self._keep = {}

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise MMAlignError(rc, self._L.mmalign_last_error(self._ctx).decode())

    # -- corpus -----------------------------------------------------------------
    @staticmethod
    def _half_dtype(emb):
        """1 / 2 (MMALIGN_F16 / MMALIGN_BF16) for half-precision rows (an encoder batch), else 0."""
        name = str(getattr(emb, "dtype", ""))
        return 1 if name.endswith("float16") and "bfloat16" not in name else (2 if "bfloat16" in name else 0)

    def _set(self, which, emb, key, bbox, terms, n_terms=0, col_offset=0):
        half = self._half_dtype(emb)
        if half:  # fp16 / bf16 rows go down as they are (mmalign_set_*_half widens them on the device)
            e = emb.contiguous() if _is_torch(emb) else np.ascontiguousarray(emb)
            pe = e.data_ptr() if _is_torch(e) else e.ctypes.data
        else:
            e, pe = _prep(emb, np.float32, "float32")
        k, pk = _prep(key, np.uint64, "int64")
        b, pb = _prep(bbox, np.float64, "float64")
        t, pt = _prep(terms, np.uint64, "int64")
        n, D = int(e.shape[0]), int(e.shape[1])
        if k.shape[0] != n:
            raise ValueError("page_key length differs from the number of embedding rows")
        W = 0 if t is None else int(t.shape[1])
        if which == "images":
            rc = self._L.mmalign_set_images_half(self._ctx, pe, half, pk, pb, pt, n, D, W) if half else \
                self._L.mmalign_set_images(self._ctx, pe, pk, pb, pt, n, D, W)
            self.N, self.D = n, D
        else:
            rc = self._L.mmalign_set_chunks_half(self._ctx, pe, half, pk, pb, pt, n, D, W, int(n_terms), int(col_offset)) \
                if half else self._L.mmalign_set_chunks(self._ctx, pe, pk, pb, pt, n, D, W, int(n_terms), int(col_offset))
            self.M, self.col_offset = n, int(col_offset)
        self._keep[which] = (e, k, b, t)  # device inputs are borrowed by the library
        self._check(rc)

    def set_images(self, emb, page_key, bbox=None, terms=None):
        self._set("images", emb, page_key, bbox, terms)

    def set_chunks(self, emb, page_key, bbox=None, terms=None, n_terms: int = 0, col_offset: int = 0):
        self._set("chunks", emb, page_key, bbox, terms, n_terms, col_offset)

    def set_option(self, name: str, value: int):
        """mmalign_set_option (e.g. "piece_bytes")."""
        self._check(self._L.mmalign_set_option(self._ctx, name.encode(), int(value)))

    def sync(self):
        """Waits for the uploads and preparation queued by set_images / set_chunks (mmalign_sync)."""
        self._check(self._L.mmalign_sync(self._ctx))

    def prep_rows(self, emb, bf16_out, norm2_out, err_out, stream=None):
        """K0 of device rows into caller-owned device buffers (mmalign_prep_rows); torch CUDA tensors."""
        n, D = int(emb.shape[0]), int(emb.shape[1])
        st = None if stream is None else C.c_void_p(int(stream))
        self._check(self._L.mmalign_prep_rows(self._ctx, emb.data_ptr(), n, D, bf16_out.data_ptr(), norm2_out.data_ptr(),
                                              err_out.data_ptr(), st))

    def set_chunks_prepared(self, emb, page_key, bbox, terms, bf16, norm2, err, n_terms: int = 0, col_offset: int = 0,
                            stream=None):
        """The gathered chunk table with its prepared operands, all CUDA tensors (mmalign_set_chunks_prepared)."""
        m, D = int(emb.shape[0]), int(emb.shape[1])
        W = 0 if terms is None else int(terms.shape[1])
        st = None if stream is None else C.c_void_p(int(stream))
        self._keep["chunks"] = (emb, page_key, bbox, terms, bf16, norm2, err)
        self._check(self._L.mmalign_set_chunks_prepared(
            self._ctx, emb.data_ptr(), page_key.data_ptr(), bbox.data_ptr(), None if terms is None else terms.data_ptr(),
            bf16.data_ptr(), norm2.data_ptr(), err.data_ptr(), m, D, W, int(n_terms), int(col_offset), st))
        self.M, self.col_offset = m, int(col_offset)

    def rescore_after(self, event):
        """The next run's exact rescoring waits for `event` (a torch.cuda.Event): mmalign_rescore_after."""
        self._keep["resc_event"] = event
        self._check(self._L.mmalign_rescore_after(self._ctx, C.c_void_p(int(event.cuda_event))))

    def num_pairs(self) -> int:
        p = C.c_int64()
        self._check(self._L.mmalign_num_pairs(self._ctx, C.byref(p)))
        return p.value

    def num_pairs_range(self, row0: int, rows: int) -> int:
        p = C.c_int64()
        self._check(self._L.mmalign_num_pairs_range(self._ctx, int(row0), int(rows), C.byref(p)))
        return p.value

    def pairs(self):
        """(pair_offsets [N+1], pair_chunk [P]) -- evaluate_alignments.py:48-69 in (image, chunk) order."""
        P = self.num_pairs()
        off = np.zeros(self.N + 1, np.int64)
        pc = np.zeros(max(P, 1), np.int64)
        self._check(self._L.mmalign_get_pairs(self._ctx, off.ctypes.data, pc.ctypes.data if P else None))
        return off, pc[:P]

    def pairs_device(self):
        """Same as pairs() but as CUDA tensors (no host copy)."""
        import torch
        P = self.num_pairs()
        dev = torch.device("cuda", self.device)
        off = torch.empty(self.N + 1, dtype=torch.int64, device=dev)
        pc = torch.empty(max(P, 1), dtype=torch.int64, device=dev)
        self._check(self._L.mmalign_get_pairs(self._ctx, off.data_ptr(), pc.data_ptr() if P else None))
        return off, pc[:P]

    # -- scoring ----------------------------------------------------------------
    @staticmethod
    def _params(schemas, candidates, k_values, mrr_cutoff, weak_weight, lam_comb, path, kprime, n_ranks=0,
                shard=None, slab=None, eps_scale=0.0, pipeline_rows=0):
        mask = schema_mask(schemas)
        ks = [int(k) for k in k_values]
        prm = _native.Params()
        prm.schema_mask = mask
        prm.candidates = CAND[candidates] if isinstance(candidates, str) else int(candidates)
        prm.n_k = len(ks)
        for i, k in enumerate(ks[:8]):
            prm.k_list[i] = k
        prm.mrr_cutoff = int(mrr_cutoff)
        prm.lam_lex, prm.lam_pos = float(weak_weight[0]), float(weak_weight[1])
        prm.lam_comb = float(weak_weight[0] + weak_weight[1]) if lam_comb is None else float(lam_comb)
        prm.path = PATHS[path] if isinstance(path, str) else int(path)
        prm.kprime = int(kprime)
        prm.n_ranks = int(n_ranks)
        prm.eps_scale = float(eps_scale)
        prm.pipeline_rows = int(pipeline_rows)
        if shard is not None:  # (first chunk row, rows) the fused pass contracts against
            prm.shard_col0, prm.shard_cols = int(shard[0]), int(shard[1])
        if slab is not None:   # (first image row, rows) that are ranked
            prm.slab_row0, prm.slab_rows = int(slab[0]), int(slab[1])
        return prm, mask, bin(mask).count("1"), ks

    def run(self, schemas="vanilla_clip", *, candidates="same_page", k_values: Sequence[int] = (1, 5, 10),
            mrr_cutoff: int = 100, weak_weight=(0.0, 0.0), lam_comb: Optional[float] = None,
            path="auto", kprime: int = 0, want=("topk", "pairs", "sums"), device_outputs=False,
            pinned_outputs=False, deep=False, stream=None, slab=None, imported=None, eps_scale=0.0,
            pipeline_rows=0):
        """slab=(row0, rows): rank only those image rows (outputs are sized by the slab).
        imported=(keys, count, tau): candidate lists received from the ranks' fused passes
        (mmalign_rescore_slab) instead of running the fused kernel here.
        pipeline_rows: query rows per pipeline slab of mmalign_run (0 = auto, -1 = one slab)."""
        prm, mask, S, ks = self._params(schemas, candidates, k_values, mrr_cutoff, weak_weight, lam_comb, path, kprime,
                                        slab=slab, eps_scale=eps_scale, pipeline_rows=pipeline_rows)
        kmax = max(ks) if ks else 0
        kneed = max(kmax, int(mrr_cutoff))
        if slab is None or tuple(slab) == (0, 0):
            P, N = self.num_pairs(), self.N
        else:
            P, N = self.num_pairs_range(*slab), int(slab[1])
        res, out = {}, _native.Out()
        if device_outputs:
            import torch
            dev = torch.device("cuda", self.device)

            def alloc(name, shape, dt):
                t = torch.empty(shape, dtype={"i64": torch.int64, "f64": torch.float64, "i32": torch.int32}[dt], device=dev)
                res[name] = t
                return t.data_ptr()
        elif pinned_outputs:
            import torch

            def alloc(name, shape, dt):
                # page-locked result buffers, allocated once and REUSED by later calls of the same shape
                key = (name, tuple(shape), dt)
                t = self._pinned.get(key)
                if t is None:
                    t = torch.empty(shape, dtype={"i64": torch.int64, "f64": torch.float64, "i32": torch.int32}[dt],
                                    pin_memory=True)
                    self._pinned = {k: v for k, v in self._pinned.items() if k[0] != name}
                    self._pinned[key] = t
                res[name] = t.numpy()
                return t.data_ptr()
        else:
            def alloc(name, shape, dt):
                a = np.empty(shape, dtype={"i64": np.int64, "f64": np.float64, "i32": np.int32}[dt])
                res[name] = a
                return a.ctypes.data
        if "topk" in want:
            out.topk_idx = alloc("topk_idx", (S, N, kmax), "i64")
            out.topk_score = alloc("topk_score", (S, N, kmax), "f64")
        if "pairs" in want:
            out.pair_rank = alloc("pair_rank", (S, P), "i32")
            out.pair_sim = alloc("pair_sim", (P,), "f64")
        if deep:
            out.pair_score = alloc("pair_score", (S, P), "f64")
            out.deep_idx = alloc("deep_idx", (S, N, kneed), "i64")
            out.deep_score = alloc("deep_score", (S, N, kneed), "f64")
        hits = np.zeros((S, len(ks)), np.int64)
        rr = np.zeros(S, np.float64)
        sim = np.zeros(1, np.float64)
        npairs = np.zeros(1, np.int64)
        stats = np.zeros(16, np.int64)
        if "sums" in want:
            out.hits, out.rr_sum, out.sim_sum = hits.ctypes.data, rr.ctypes.data, sim.ctypes.data
        out.num_pairs, out.stats = npairs.ctypes.data, stats.ctypes.data
        st = None if stream is None else C.c_void_p(int(stream))
        if imported is None:
            self._check(self._L.mmalign_run(self._ctx, C.byref(prm), C.byref(out), st))
        else:
            keys, count, tau = imported  # [n_src, list_rows, stride], [n_src, list_rows] x 2
            n_src, list_rows, stride = keys.shape
            self._check(self._L.mmalign_rescore_slab(self._ctx, C.byref(prm), keys.data_ptr(), count.data_ptr(),
                                                     tau.data_ptr(), int(n_src), int(list_rows), int(stride),
                                                     C.byref(out), st))
        res.update(hits=hits, rr_sum=rr, sim_sum=float(sim[0]), num_pairs=int(npairs[0]),
                   stats=dict(rows_rescanned=int(stats[0]), candidates_rescored=int(stats[1]),
                              fused_launches=int(stats[2]), kernel_launches=int(stats[3]),
                              kprime=int(stats[4]), fused_us=int(stats[5]), rescore_us=int(stats[6]),
                              exact_scan_us=int(stats[7]), eps_violations=int(stats[8]), slabs=int(stats[9]),
                              k2_sms=int(stats[10])),
                   schemas=[s for s in SCHEMAS if SCHEMA_BITS[s] & mask], k_values=ks)
        return res

    def alignments(self, schema: str, raw: bool = False):
        """Records of the `alignments` table (insert_clip_embeddings.py:369-414): rec [P,3];
        raw=True gives the un-thresholded (lexical, positional, 0) scores."""
        P = self.num_pairs()
        rec = np.zeros((max(P, 1), 3), np.float64)
        self._check(self._L.mmalign_alignments(self._ctx, SCHEMA_BITS[schema] | (0x100 if raw else 0),
                                               rec.ctypes.data, None))
        return rec[:P]

    def term_bitsets(self, texts_lower: Sequence[str], terms: Sequence[str], term_words: Optional[int] = None):
        """The chunks' lexical term sets (mmalign_term_bitsets): bits [m, W] uint64, bit t of row j =
        terms[t] is a substring of texts_lower[j] (src/insert_clip_embeddings.py:149-150; the caller lower-cases)."""
        def pack(strings):
            enc = [s.encode("utf-8") for s in strings]
            off = np.zeros(len(enc) + 1, np.int64)
            if enc:
                np.cumsum([len(b) for b in enc], out=off[1:])
            return np.frombuffer(b"".join(enc) or b"\0", np.uint8), off
        tb, to = pack(texts_lower)
        pb, po = pack(terms)
        W = int(term_words or max(1, (len(terms) + 63) // 64))
        bits = np.zeros((len(texts_lower), W), np.uint64)
        self._check(self._L.mmalign_term_bitsets(self._ctx, tb.ctypes.data, to.ctypes.data, len(texts_lower),
                                                 pb.ctypes.data, po.ctypes.data, len(terms), W, bits.ctypes.data, None))
        return bits

    def term_bitsets_device(self, text, text_off, terms: Sequence[str], bits):
        """mmalign_term_bitsets on device-resident tensors: text uint8 [bytes], text_off int64 [m + 1],
        bits int64 [m, W] (out).  The term table (tiny) goes up with the call."""
        enc = [s.encode("utf-8") for s in terms]
        po = np.zeros(len(enc) + 1, np.int64)
        if enc:
            np.cumsum([len(b) for b in enc], out=po[1:])
        pb = np.frombuffer(b"".join(enc) or b"\0", np.uint8)
        m, W = int(bits.shape[0]), int(bits.shape[1])
        self._check(self._L.mmalign_term_bitsets(self._ctx, text.data_ptr(), text_off.data_ptr(), m, pb.ctypes.data,
                                                 po.ctypes.data, len(terms), W, bits.data_ptr(), None))
        return bits

    def alignments_device(self, schema: str, rec, raw: bool = False):
        """mmalign_alignments into a device tensor rec float64 [P, 3]."""
        self._check(self._L.mmalign_alignments(self._ctx, SCHEMA_BITS[schema] | (0x100 if raw else 0), rec.data_ptr(), None))
        return rec

    def debug_scores(self):
        out = np.zeros((self.N, self.M), np.float32)
        self._check(self._L.mmalign_debug_scores(self._ctx, out.ctypes.data, None))
        return out

    def copy_decode(self, data: bytes, n_cols: int, vec_col: int, bbox_col: int = -1, page_col: int = -1, D: int = 0):
        """A PostgreSQL binary COPY stream -> (field_off [n, n_cols], field_len [n, n_cols], emb [n, D] f32, bbox [n, 4] f64,
        page [n] i32, page_null [n] bool): tuple walk on the host (mmalign_copy_scan), bulk columns decoded on the GPU
        (mmalign_copy_decode).  D = 0: taken from the first row's vector field."""
        buf = np.frombuffer(data, np.uint8)
        n = int(self._L.mmalign_copy_scan(buf.ctypes.data, len(buf), n_cols, None, None, 0))
        if n < 0:
            raise ValueError(f"not a usable binary COPY stream (mmalign_copy_scan code {n})")
        off = np.zeros((max(n, 1), n_cols), np.int64)
        ln = np.zeros((max(n, 1), n_cols), np.int32)
        got = int(self._L.mmalign_copy_scan(buf.ctypes.data, len(buf), n_cols, off.ctypes.data, ln.ctypes.data, n))
        assert got == n
        off, ln = off[:n], ln[:n]
        if n and D == 0 and vec_col >= 0:
            D = (int(ln[0, vec_col]) - 4) // 4
        emb = np.zeros((n, max(D, 0)), np.float32)
        bbox = np.zeros((n, 4), np.float64)
        page = np.zeros(n, np.int32)
        null = np.zeros(n, np.uint8)
        if n:
            self._check(self._L.mmalign_copy_decode(self._ctx, buf.ctypes.data, len(buf), off.ctypes.data, ln.ctypes.data, n,
                                                    n_cols, vec_col, bbox_col, page_col, D,
                                                    emb.ctypes.data if vec_col >= 0 else None, bbox.ctypes.data,
                                                    page.ctypes.data, null.ctypes.data, None))
        return off, ln, emb, bbox, page, null.astype(bool)

    def debug_operands(self):
        """The bf16 operands of the fused kernel as float32 arrays (validation hook)."""
        a = np.zeros((self.N, self.D), np.uint16)
        b = np.zeros((self.M, self.D), np.uint16)
        self._check(self._L.mmalign_debug_operands(self._ctx, a.ctypes.data, b.ctypes.data, None))
        widen = lambda x: (x.astype(np.uint32) << 16).view(np.float32)
        return widen(a), widen(b)

    # -- sharded run, default exchange: contraction sharded by chunk columns, rescoring by query rows -----
    def fused_pass(self, schemas, *, shard, k_values, mrr_cutoff=100, weak_weight=(0.0, 0.0), lam_comb=None,
                   kprime=0, n_ranks=1):
        """K1 of every image row against chunk rows [shard[0], shard[0] + shard[1]); the lists stay in the context."""
        prm, *_ = self._params(schemas, "all", k_values, mrr_cutoff, weak_weight, lam_comb, "auto", kprime, n_ranks,
                               shard=shard)
        if shard[1] == 0 and shard[0] == 0:
            raise ValueError("fused_pass: an empty shard must be given as (M, 0), not (0, 0) = the whole table")
        self._check(self._L.mmalign_fused_pass(self._ctx, C.byref(prm), None, None))

    def list_stride(self) -> int:
        v = C.c_int32()
        self._check(self._L.mmalign_list_stride(self._ctx, C.byref(v)))
        return int(v.value)

    def export_lists(self, n_dest: int, slab_rows: int, stride: int):
        """(keys [n_dest, slab_rows, stride] int64 view of u64, count [n_dest, slab_rows], tau [n_dest, slab_rows])."""
        import torch
        dev = torch.device("cuda", self.device)
        keys = torch.empty((n_dest, slab_rows, stride), dtype=torch.int64, device=dev)
        count = torch.empty((n_dest, slab_rows), dtype=torch.int32, device=dev)
        tau = torch.empty((n_dest, slab_rows), dtype=torch.float32, device=dev)
        self._check(self._L.mmalign_export_lists(self._ctx, int(n_dest), int(slab_rows), int(stride), keys.data_ptr(),
                                                 count.data_ptr(), tau.data_ptr(), None))
        return keys, count, tau

    # -- sharded passes, fully sharded variant (device tensors; see include/mmalign.h and distributed.py) ------------
    def sharded_session(self, schemas, *, k_values, mrr_cutoff=100, weak_weight=(0.0, 0.0), lam_comb=None, kprime=0,
                        n_ranks=1):
        return _ShardedSession(self, schemas, k_values, mrr_cutoff, weak_weight, lam_comb, kprime, n_ranks)

    # -- multi-GPU helpers (device tensors) ---------------------------------------
    def merge_topk(self, gathered_idx, gathered_score):
        """[G, L, K] gathered lists -> merged [L, K] (torch CUDA tensors)."""
        import torch
        G, L, K = gathered_idx.shape
        oi = torch.empty((L, K), dtype=torch.int64, device=gathered_idx.device)
        os_ = torch.empty((L, K), dtype=torch.float64, device=gathered_idx.device)
        self._check(self._L.mmalign_merge_topk(self._ctx, gathered_idx.data_ptr(), gathered_score.data_ptr(),
                                               G, L, K, oi.data_ptr(), os_.data_ptr(), None))
        return oi, os_

    def count_beating(self, deep_idx, deep_score, q_image, q_chunk, q_score):
        import torch
        S, N, K = deep_idx.shape
        n_q = int(q_image.shape[0])
        counts = torch.zeros((S, n_q), dtype=torch.int32, device=deep_idx.device)
        if n_q:
            self._check(self._L.mmalign_count_beating(self._ctx, deep_idx.data_ptr(), deep_score.data_ptr(), N, S, K,
                                                      n_q, q_image.data_ptr(), q_chunk.data_ptr(),
                                                      q_score.data_ptr(), counts.data_ptr(), None))
        return counts

    def reduce_metrics(self, pair_rank, pair_sim, k_values, mrr_cutoff=100):
        S, P = pair_rank.shape
        ks = np.asarray(list(k_values), np.int32)
        hits = np.zeros((S, len(ks)), np.int64)
        rr = np.zeros(S, np.float64)
        sim = np.zeros(1, np.float64)
        self._check(self._L.mmalign_reduce_metrics(self._ctx, pair_rank.data_ptr(),
                                                   pair_sim.data_ptr() if pair_sim is not None else None, S, P,
                                                   ks.ctypes.data, len(ks), int(mrr_cutoff), hits.ctypes.data,
                                                   rr.ctypes.data, sim.ctypes.data, None))
        return hits, rr, float(sim[0])


_DEFAULT = {}


def default_engine(device: int = 0, role: str = "main") -> AlignmentEngine:
    """The process-wide engine of a device, shared by the modules that mirror the reference (one context per role:
    "main" holds the registered schema's tables, "scratch" serves one-off calls that must not disturb them)."""
    key = (int(device), role)
    eng = _DEFAULT.get(key)
    if eng is None or not eng._ctx.value:
        eng = _DEFAULT[key] = AlignmentEngine(device)
    return eng


class _ShardedSession:
    """The three passes of a sharded run on one rank (mmalign_fused_pass / rescore_pass / rescan_rows)."""

    def __init__(self, eng, schemas, k_values, mrr_cutoff, weak_weight, lam_comb, kprime, n_ranks):
        import torch
        self.eng, self.torch = eng, torch
        self.prm, self.mask, self.S, self.ks = eng._params(schemas, "all", k_values, mrr_cutoff, weak_weight, lam_comb,
                                                           "auto", kprime, n_ranks)
        self.kmax, self.kneed = max(self.ks), max(max(self.ks), int(mrr_cutoff))
        self.dev = torch.device("cuda", eng.device)
        self.out = None

    def fused_pass(self):
        """K1 on this rank's shard; returns tau_row [N] (device)."""
        tau = self.torch.empty(self.eng.N, dtype=self.torch.float32, device=self.dev)
        self.eng._check(self.eng._L.mmalign_fused_pass(self.eng._ctx, C.byref(self.prm), tau.data_ptr(), None))
        return tau

    def chunk_err_max(self) -> float:
        v = C.c_float()
        self.eng._check(self.eng._L.mmalign_chunk_err_max(self.eng._ctx, C.byref(v)))
        return float(v.value)

    def _outputs(self):
        t, S, N, P = self.torch, self.S, self.eng.N, self.eng.num_pairs()
        res = dict(topk_idx=t.empty((S, N, self.kmax), dtype=t.int64, device=self.dev),
                   topk_score=t.empty((S, N, self.kmax), dtype=t.float64, device=self.dev),
                   pair_rank=t.empty((S, P), dtype=t.int32, device=self.dev),
                   pair_sim=t.empty((P,), dtype=t.float64, device=self.dev),
                   pair_score=t.empty((S, P), dtype=t.float64, device=self.dev),
                   deep_idx=t.empty((S, N, self.kneed), dtype=t.int64, device=self.dev),
                   deep_score=t.empty((S, N, self.kneed), dtype=t.float64, device=self.dev))
        out = _native.Out()
        for k, v in res.items():
            setattr(out, k, v.data_ptr())
        return res, out

    def rescore_pass(self, tau_global, eps_chunk_global: float):
        """Exact scores of this rank's entries above the global tau.  Returns (results, cert_count [S,N])."""
        res, out = self._outputs()
        stats = np.zeros(16, np.int64)
        out.stats = stats.ctypes.data
        cert = self.torch.empty((self.S, self.eng.N), dtype=self.torch.int32, device=self.dev)
        self.eng._check(self.eng._L.mmalign_rescore_pass(self.eng._ctx, C.byref(self.prm), tau_global.data_ptr(),
                                                         C.c_float(eps_chunk_global), C.byref(out), cert.data_ptr(), None))
        res["stats"] = dict(rows_rescanned=0, candidates_rescored=int(stats[1]), fused_launches=int(stats[2]),
                            kernel_launches=int(stats[3]), kprime=int(stats[4]), fused_us=int(stats[5]),
                            rescore_us=int(stats[6]), exact_scan_us=0)
        self.res, self.out = res, out
        return res, cert

    def rescan_rows(self, rows):
        """Exact scan of `rows` (int32 device tensor) into the outputs of the last rescore_pass."""
        if rows.numel():
            self.out.stats = None
            self.eng._check(self.eng._L.mmalign_rescan_rows(self.eng._ctx, C.byref(self.prm), rows.data_ptr(),
                                                            int(rows.numel()), C.byref(self.out), None))
            self.res["stats"]["rows_rescanned"] = int(rows.numel())
