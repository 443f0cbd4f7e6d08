"""Loader of the CUDA library ``libfdal.so`` (built in-tree by ``build.py``).

There is deliberately no fallback: if the shared object is missing or cannot
be loaded, importing the product path raises.
"""
from __future__ import annotations

import os

from . import _binding as b

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libfdal.so")
_api = None


def load() -> b.Api:
    global _api
    if _api is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -m fictitious_domain_al_preconditioners_b200.build` "
                "(nvcc, sm_100a). There is no CPU fallback."
            )
        _api = b.Api(LIB_PATH, "fdal_", extra=b.DEVICE_SIGNATURES)
    return _api
