"""Size / output-dimensionality inference for multi-output kernels (``stheno/mo/infer.py:16-102``)."""
from ..kernels import (Kernel, PosteriorKernel, ProductKernel, ReversedKernel, ScaledKernel, StretchedKernel,
                       SubspaceKernel, SumKernel, num_elements)

__all__ = ["infer_size", "dimensionality"]


def infer_size(k, x):
    """Size of ``k`` evaluated at ``x`` (numeric input, FDD, or tuple of those)."""
    from ..model.fdd import FDD

    if isinstance(x, tuple):
        return sum(infer_size(k, xi) for xi in x)
    if isinstance(x, FDD):
        return num_elements(x)
    d = dimensionality(k)
    if d is None:
        raise RuntimeError(f"Could not infer dimensionality of {k}.")
    return num_elements(x) * d


def _check_and_merge(k, *ds):
    ds = [d for d in ds if d]
    if not ds:
        return None
    if not all(d == ds[0] for d in ds[1:]):
        raise RuntimeError(f"Inferred dimensionalities for kernel {k} do not match. ")
    return ds[0]


def dimensionality(k):
    """Output dimensionality of ``k`` (None if it cannot be inferred)."""
    from .kernel import CrossKernel, MultiOutputKernel, MultiOutputMean

    if isinstance(k, (MultiOutputKernel, MultiOutputMean)):
        return len(k.ps)
    if isinstance(k, CrossKernel):
        return None  # the reference wraps this case in AmbiguousDimensionalityKernel (stheno/mo/adk.py)
    if isinstance(k, (SumKernel, ProductKernel)):
        return _check_and_merge(k, dimensionality(k.a), dimensionality(k.b))
    if isinstance(k, (ScaledKernel, StretchedKernel, ReversedKernel)):
        return dimensionality(k.k)
    if isinstance(k, PosteriorKernel):
        return _check_and_merge(k, dimensionality(k.k_ij), dimensionality(k.k_zi), dimensionality(k.k_zj))
    if isinstance(k, SubspaceKernel):
        return _check_and_merge(k, dimensionality(k.k_zi), dimensionality(k.k_zj))
    return 1
