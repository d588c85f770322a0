"""B200-native alignment scoring and retrieval.

The compute path is the CUDA library csrc/libmmalign.so behind the C ABI of
include/mmalign.h; importing `engine` fails loudly when it has not been built.
"""
from . import _native  # noqa: F401
from .corpus import Corpus, build_corpus  # noqa: F401
from .engine import AlignmentEngine, MMAlignError, SCHEMAS  # noqa: F401

__all__ = ["AlignmentEngine", "MMAlignError", "SCHEMAS", "Corpus", "build_corpus"]
