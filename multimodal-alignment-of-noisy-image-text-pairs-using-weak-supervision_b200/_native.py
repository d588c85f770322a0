"""ctypes binding of csrc/libmmalign.so (the C ABI declared in include/mmalign.h).

There is no fallback: if the shared library is missing or fails to load, import
of the engine raises.  The library is built in-tree by `build()` (nvcc, sm_100a).
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

CSRC = Path(__file__).resolve().parent / "csrc"
LIB_PATH = CSRC / "libmmalign.so"
SOURCES = ["api.cu", "prep.cu", "rescore.cu", "fused_tc.cu", "ingest.cu", "common.cuh"]

NULL_KEY = 0xFFFFFFFFFFFFFFFF
SCHEMA_BITS = {"vanilla_clip": 1, "clip_lexical": 2, "clip_positional": 4, "clip_combined": 8}
CAND = {"same_page": 0, "all": 1}
PATHS = {"auto": 0, "exact": 1, "fused": 2}


class Params(C.Structure):
    _fields_ = [("schema_mask", C.c_uint32), ("candidates", C.c_int32), ("n_k", C.c_int32),
                ("k_list", C.c_int32 * 8), ("mrr_cutoff", C.c_int32), ("lam_lex", C.c_double),
                ("lam_pos", C.c_double), ("lam_comb", C.c_double), ("path", C.c_int32),
                ("kprime", C.c_int32), ("n_ranks", C.c_int32), ("eps_scale", C.c_float),
                ("shard_col0", C.c_int64), ("shard_cols", C.c_int64), ("slab_row0", C.c_int64),
                ("slab_rows", C.c_int64), ("pipeline_rows", C.c_int32), ("reserved", C.c_int32)]


class Out(C.Structure):
    _fields_ = [("topk_idx", C.c_void_p), ("topk_score", C.c_void_p), ("pair_rank", C.c_void_p),
                ("pair_sim", C.c_void_p), ("hits", C.c_void_p), ("rr_sum", C.c_void_p),
                ("sim_sum", C.c_void_p), ("num_pairs", C.c_void_p), ("pair_score", C.c_void_p),
                ("deep_idx", C.c_void_p), ("deep_score", C.c_void_p), ("stats", C.c_void_p)]


EXPORTS = ["mmalign_abi_version", "mmalign_create", "mmalign_destroy", "mmalign_last_error",
           "mmalign_set_images", "mmalign_set_chunks", "mmalign_num_pairs", "mmalign_get_pairs",
           "mmalign_run", "mmalign_alignments", "mmalign_merge_topk", "mmalign_count_beating",
           "mmalign_reduce_metrics", "mmalign_debug_scores", "mmalign_fused_pass", "mmalign_chunk_err_max",
           "mmalign_rescore_pass", "mmalign_rescan_rows", "mmalign_list_stride", "mmalign_export_lists",
           "mmalign_rescore_slab", "mmalign_num_pairs_range", "mmalign_term_bitsets", "mmalign_sync",
           "mmalign_prep_rows", "mmalign_set_chunks_prepared", "mmalign_rescore_after", "mmalign_debug_operands", "mmalign_set_option",
           "mmalign_set_images_half", "mmalign_set_chunks_half", "mmalign_copy_scan", "mmalign_copy_decode",
           "mmalign_check_report"]
ABI_VERSION = 3

_lib = None


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile the CUDA sources for sm_100a into csrc/libmmalign.so (nvcc cross-compiles without a GPU)."""
    newest = max((CSRC / s).stat().st_mtime for s in SOURCES)
    hdr = CSRC.parent.parent / "include" / "mmalign.h"
    newest = max(newest, hdr.stat().st_mtime)
    if force or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < newest:
        cmd = ["make", "-C", str(CSRC), "-j4"] + (["-B"] if force else [])
        r = subprocess.run(cmd, capture_output=not verbose, text=True)
        if r.returncode != 0:
            raise RuntimeError("building libmmalign.so failed:\n" + (r.stdout or "") + (r.stderr or ""))
    return LIB_PATH


def load():
    """Loads libmmalign.so; raises (loudly) when it is absent -- there is no CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    import os
    path = Path(os.environ.get("MMALIGN_LIB", LIB_PATH))  # development: try another build of the same ABI
    if not path.exists():
        raise RuntimeError(f"{path} is missing: build it with __graft_entry__.build() "
                           "(the scoring path is CUDA-only and has no fallback)")
    L = C.CDLL(str(path))
    vp, i64, i32, u32, dbl = C.c_void_p, C.c_int64, C.c_int32, C.c_uint32, C.c_double
    L.mmalign_abi_version.restype = C.c_int
    L.mmalign_create.argtypes = [C.POINTER(vp), C.c_int]
    L.mmalign_destroy.argtypes = [vp]
    L.mmalign_destroy.restype = None
    L.mmalign_last_error.argtypes = [vp]
    L.mmalign_last_error.restype = C.c_char_p
    L.mmalign_set_images.argtypes = [vp, vp, vp, vp, vp, i64, i32, i32]
    L.mmalign_set_chunks.argtypes = [vp, vp, vp, vp, vp, i64, i32, i32, i64, i64]
    L.mmalign_num_pairs.argtypes = [vp, C.POINTER(i64)]
    L.mmalign_get_pairs.argtypes = [vp, vp, vp]
    L.mmalign_run.argtypes = [vp, C.POINTER(Params), C.POINTER(Out), vp]
    L.mmalign_alignments.argtypes = [vp, u32, vp, vp]
    L.mmalign_merge_topk.argtypes = [vp, vp, vp, i32, i64, i32, vp, vp, vp]
    L.mmalign_count_beating.argtypes = [vp, vp, vp, i64, i32, i32, i64, vp, vp, vp, vp, vp]
    L.mmalign_reduce_metrics.argtypes = [vp, vp, vp, i32, i64, vp, i32, i32, vp, vp, vp, vp]
    L.mmalign_debug_scores.argtypes = [vp, vp, vp]
    L.mmalign_fused_pass.argtypes = [vp, C.POINTER(Params), vp, vp]
    L.mmalign_chunk_err_max.argtypes = [vp, C.POINTER(C.c_float)]
    L.mmalign_rescore_pass.argtypes = [vp, C.POINTER(Params), vp, C.c_float, C.POINTER(Out), vp, vp]
    L.mmalign_rescan_rows.argtypes = [vp, C.POINTER(Params), vp, i64, C.POINTER(Out), vp]
    L.mmalign_list_stride.argtypes = [vp, C.POINTER(i32)]
    L.mmalign_export_lists.argtypes = [vp, i32, i64, i32, vp, vp, vp, vp]
    L.mmalign_rescore_slab.argtypes = [vp, C.POINTER(Params), vp, vp, vp, i32, i64, i32, C.POINTER(Out), vp]
    L.mmalign_num_pairs_range.argtypes = [vp, i64, i64, C.POINTER(i64)]
    L.mmalign_term_bitsets.argtypes = [vp, vp, vp, i64, vp, vp, i32, i32, vp, vp]
    L.mmalign_sync.argtypes = [vp]
    L.mmalign_prep_rows.argtypes = [vp, vp, i64, i32, vp, vp, vp, vp]
    L.mmalign_set_chunks_prepared.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i64, i32, i32, i64, i64, vp]
    L.mmalign_rescore_after.argtypes = [vp, vp]
    L.mmalign_debug_operands.argtypes = [vp, vp, vp, vp]
    L.mmalign_set_option.argtypes = [vp, C.c_char_p, i64]
    L.mmalign_set_images_half.argtypes = [vp, vp, i32, vp, vp, vp, i64, i32, i32]
    L.mmalign_set_chunks_half.argtypes = [vp, vp, i32, vp, vp, vp, i64, i32, i32, i64, i64]
    L.mmalign_copy_scan.argtypes = [vp, i64, i32, vp, vp, i64]
    L.mmalign_copy_scan.restype = C.c_int64
    L.mmalign_copy_decode.argtypes = [vp, vp, i64, vp, vp, i64, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp]
    L.mmalign_check_report.argtypes = [C.POINTER(C.c_uint32)]
    for name in EXPORTS:
        getattr(L, name)
        if name not in ("mmalign_destroy", "mmalign_last_error", "mmalign_abi_version", "mmalign_copy_scan"):
            getattr(L, name).restype = C.c_int
    if L.mmalign_abi_version() != ABI_VERSION:
        raise RuntimeError(f"{path} has ABI version {L.mmalign_abi_version()}, this package needs {ABI_VERSION}: rebuild it")
    _lib = L
    return L


def check_report():
    """mmalign_check_report: dict(checked=bool, violations={file: (count, first line)}).  `checked` is True only for
    the checked build of the library (csrc/Makefile `make check`, selected with MMALIGN_LIB)."""
    out = (C.c_uint32 * 9)()
    rc = load().mmalign_check_report(out)
    if rc != 0:
        raise RuntimeError(f"mmalign_check_report failed with {rc}")
    files = ("fused_tc.cu", "rescore.cu", "prep.cu", "ingest.cu")
    return dict(checked=bool(out[0]), violations={f: (int(out[1 + 2 * q]), int(out[2 + 2 * q])) for q, f in enumerate(files)})
