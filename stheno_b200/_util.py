"""Host-side plumbing: moving user inputs to the device and results back to where the user's data lives."""
import numpy as np
import torch

NUMPY = "numpy"


def device():
    if not torch.cuda.is_available():
        raise RuntimeError(
            "stheno_b200 needs a CUDA device: the GP hot path runs only on the sm_100a kernels (no CPU fallback)"
        )
    return torch.device("cuda", torch.cuda.current_device())


# Test seam: the CPU test-suite replaces this with ``lambda: torch.device('cpu')`` together with a fake ``ops``.
_device_fn = device


def is_numeric(x):
    return isinstance(x, (int, float, np.ndarray, np.number, torch.Tensor, list))


def origin_of(x):
    """Where results derived from ``x`` should be returned: 'numpy' or a torch device."""
    if isinstance(x, torch.Tensor):
        return x.device
    return NUMPY


def to_dev(x, dtype=None):
    """numpy / list / scalar / torch (any device) -> tensor on the compute device.  Float64 unless the data is
    already float32 (the reference's dtype-follows-input rule, ``stheno/model/fdd.py:63,115-117``)."""
    dev = _device_fn()
    if isinstance(x, torch.Tensor):
        t = x
        if not t.is_floating_point():
            t = t.to(torch.float64)
    else:
        a = np.asarray(x)
        if a.dtype == np.float32:
            t = torch.from_numpy(np.ascontiguousarray(a))
        else:
            t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64))
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    if t.device != dev:
        t = t.to(dev, non_blocking=True)
    return t


def from_dev(t, origin):
    if t is None:
        return None
    if origin == NUMPY:
        return t.detach().cpu().numpy()
    if isinstance(origin, torch.device) and t.device != origin:
        return t.to(origin)
    return t


def uprank(t, rank=2):
    """``B.uprank``: scalars / vectors become column matrices."""
    while t.dim() < rank:
        t = t.unsqueeze(-1) if t.dim() >= 1 else t.reshape(1, 1)
    return t


def batch_flatten(t, keep):
    """``[..., a, b] -> ([B, a, b], batch_shape)`` keeping the last ``keep`` dims."""
    bs = t.shape[: t.dim() - keep]
    return t.reshape((-1,) + tuple(t.shape[t.dim() - keep :])), tuple(bs)
