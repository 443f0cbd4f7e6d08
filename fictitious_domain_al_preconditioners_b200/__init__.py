"""B200-native augmented-Lagrangian solve path.

Host-side mirror of the reference operator API
(``augmented_lagrangian_preconditioner.h``) on top of the C ABI in
``include/fdal.h``; all arithmetic runs in the CUDA library ``csrc/libfdal.so``.
"""
from . import _binding as abi  # noqa: F401
from .context import (  # noqa: F401
    ALConfig,
    ALContext,
    FdalError,
    IterationNumberControl,
    NoConvergence,
    ReductionControl,
    SolverControl,
)

__version__ = "0.1.0"
