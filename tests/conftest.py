import importlib
import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

PKG_NAME = "multimodal-alignment-of-noisy-image-text-pairs-using-weak-supervision_b200"
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def pkg():
    return importlib.import_module(PKG_NAME)


@pytest.fixture(scope="session")
def synthetic():
    return importlib.import_module(PKG_NAME + ".synthetic")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.lib()
    return o


class OracleIngest:
    """Stand-in for the engine's ingest call in the CPU-only tests: the term sets come from the oracle.
    (tests/test_gpu_parity.py::test_term_bitsets_* hold the GPU ingest to the same answers.)"""

    def term_bitsets(self, texts_lower, terms, term_words=None):
        from oracle import oracle as o
        return o.term_bitsets(texts_lower, terms, term_words)


@pytest.fixture(scope="session")
def small_corpus(pkg):
    d = json.loads((GOLDEN / "small_corpus.json").read_text())
    z = np.load(GOLDEN / "small_corpus.npz")
    corpus = pkg.build_corpus(d["images"], d["chunks"], z["img_emb"], z["chk_emb"], d["lexical_components"],
                              engine=OracleIngest())
    return d, corpus


def unhex(x):
    return float.fromhex(x)


def random_texts_and_terms(rng, m, T):
    """Texts and terms over a small alphabet (so that matches are common) with multi-byte UTF-8, upper case,
    empty strings, duplicates, terms longer than texts and terms that end exactly at the end of a text."""
    alphabet = list("abcAB \n") + ["é", "ß", "Ж", "中", "😀"]
    def word(lo, hi):
        return "".join(rng.choice(alphabet, size=int(rng.integers(lo, hi + 1))))
    texts = [word(0, 80) for _ in range(m)]
    terms = [word(0, 4).lower() for _ in range(T)]
    terms[0] = ""                                  # occurs everywhere
    terms[1] = terms[2] = "ab"                     # duplicates count twice
    terms[3] = "Ab"                                # never matches lower-cased text
    terms[4] = texts[1][-3:].lower() if len(texts[1]) >= 3 else "zz"   # suffix of a text
    terms[5] = "a" * 200                           # longer than every text
    texts[0] = ""
    return texts, terms


def pytest_sessionfinish(session, exitstatus):
    """A run against the checked build of the library (MMALIGN_LIB=.../libmmalign_check.so: every kernel tests its own
    indices and invariants) ends with its report: one line on stdout and in gpurun_out/, and a failing exit status if
    any check fired."""
    import os
    if "check" not in os.path.basename(os.environ.get("MMALIGN_LIB", "")):
        return
    try:
        import torch
        if not torch.cuda.is_available():
            return
        rep = importlib.import_module(PKG_NAME + "._native").check_report()
    except Exception as e:  # noqa: BLE001 -- the report must not mask the suite's own result
        print(f"\nchecked build: no report ({e})")
        return
    bad = {f: v for f, v in rep["violations"].items() if v[0]}
    line = (f"checked build: checked={rep['checked']}, tests exit status {int(exitstatus)}, "
            f"violations {rep['violations']}")
    print("\n" + line)
    out = ROOT / "gpurun_out"
    if out.is_dir():
        (out / "checked_build_report.txt").write_text(line + "\n")
    if rep["checked"] and bad and session.exitstatus == 0:
        session.exitstatus = 1
