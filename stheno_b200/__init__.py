"""stheno_b200 -- a B200-native (sm_100a) implementation of the Stheno GP-inference hot path.

Drop-in for the part of ``stheno`` / ``stheno.torch`` that sits behind ``f(x, noise).logpdf(y)`` and
``f | (f(x, noise), y)``: ``GP``, ``Measure``, ``FDD`` / ``Normal``, ``Obs`` / ``PseudoObs*``, ``cross``, the kernels
``EQ, Exp/Matern12, Matern32, Matern52, Linear, Delta`` with ``c * k``, ``k + k``, ``k * k``, ``k.stretch(l)``, and
the global jitter ``B.epsilon`` (surface listed in SURVEY.md section 8b; reference ``stheno/__init__.py:1-28``).
All arithmetic runs in the hand-written CUDA kernels of ``csrc/`` through the C-ABI ``include/gpk.h``.
"""
from . import B  # noqa: F401
from . import matrix  # noqa: F401
from .kernels import *  # noqa: F401,F403
from .lazy import *  # noqa: F401,F403
from .matrix import Dense, Diagonal, Zero  # noqa: F401
from .mo import *  # noqa: F401,F403
from .model import *  # noqa: F401,F403
from .random import *  # noqa: F401,F403

__version__ = "0.1.0"


class BreakingChangeWarning(UserWarning):
    """Kept for import compatibility with ``stheno/__init__.py:21-28``."""
