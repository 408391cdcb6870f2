"""``from stheno_b200.torch import GP, EQ, ...`` -- mirror of ``stheno/torch.py:1-5`` (the reference's torch shim
re-exports everything; here torch CUDA tensors are the native array type already)."""
from . import *  # noqa: F401,F403
from . import B  # noqa: F401
