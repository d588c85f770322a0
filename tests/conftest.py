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


@pytest.fixture(scope="session")
def small_corpus(pkg):
    d = json.loads((GOLDEN / "small_corpus.json").read_text())
    z = np.load(GOLDEN / "small_corpus.npz")
    corpus = pkg.build_corpus(d["images"], d["chunks"], z["img_emb"], z["chk_emb"], d["lexical_components"])
    return d, corpus


def unhex(x):
    return float.fromhex(x)
