from .infer import dimensionality, infer_size
from .kernel import CrossKernel, MultiOutputKernel, MultiOutputMean

__all__ = ["MultiOutputKernel", "MultiOutputMean", "CrossKernel", "infer_size", "dimensionality"]
