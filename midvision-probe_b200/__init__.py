"""mvmatch -- B200-native (sm_100a) dense-correspondence matching for midvision-probe.

``correspondence`` mirrors the reference's ``evals/utils/correspondence.py`` one function to one;
``spair`` holds the SPair matching the reference inlines in its eval script; ``evaluation`` shards
pairs over GPUs and reduces integer hit counts; ``affinity`` serves the two adjacent similarity consumers (MaskCut's
affinity matrix, the 2AFC cosine evaluation).  All arithmetic is in ``lib/libmvmatch.so``
(sources in ``csrc/``, C ABI in ``include/mvmatch.h``).
"""
from . import _lib, affinity, correspondence, evaluation, spair, transformations  # noqa: F401
from ._lib import MvMatchError, load  # noqa: F401
from .build import build_lib  # noqa: F401

__version__ = "0.1.0"
