"""Multi-output glue: kernels / means over tuples of FDDs (``stheno/mo/kernel.py:14-99``, ``mo/mean.py:9-46``,
``mo/input.py:7-36``).  The joint covariance of several processes is the block matrix ``[[k_{p_i p_j}(x_i, x_j)]]``;
each block is one fused kernel-matrix launch written straight into its place of the joint matrix -- or, when the joint is
factorised, of the padded lower-triangular workspace (``matrix.BlockDense``)."""
import torch

from .. import matrix as M
from ..kernels import Kernel, Mean, as_input, pairwise
from .._util import from_dev

__all__ = ["MultiOutputKernel", "MultiOutputMean", "CrossKernel"]


class MultiOutputKernel(Kernel):
    """``MultiOutputKernel(measure, *ps)``: the kernel of ``cross(*ps)``."""

    symmetric = False

    def __init__(self, measure, *ps):
        self.measure, self.ps = measure, ps

    def _pairwise_multi(self, x, y, same):
        from ..model.fdd import FDD

        if not isinstance(x, FDD):
            x = tuple(p(x) for p in self.ps)
        if not isinstance(y, FDD):
            y = x if same and not isinstance(x, FDD) else tuple(p(y) for p in self.ps)
        if isinstance(x, FDD) and isinstance(y, FDD):
            same_pts = x.x is y.x
            return pairwise(self.measure.kernels[x.p, y.p], x.x, None if same_pts else y.x)
        return pairwise(self, x if isinstance(x, tuple) else (x,), y if isinstance(y, tuple) else (y,))

    def _elwise_multi(self, x, y, same):
        from ..kernels import _elwise_any
        from ..model.fdd import FDD

        if isinstance(x, FDD) != isinstance(y, FDD):
            raise ValueError('Unclear combination of arguments given to "elwise".')
        if isinstance(x, FDD):
            return _elwise_any(self.measure.kernels[x.p, y.p], x.x, y.x, x.x is y.x)
        xs = tuple(p(x) for p in self.ps)
        ys = xs if same else tuple(p(y) for p in self.ps)
        return torch.cat([self._elwise_multi(xi, yi, same) for xi, yi in zip(xs, ys)], dim=-2)

    def _matrix(self, x, y, same):
        return self._pairwise_multi(x, y, same)

    def render(self):
        ks = [str(self.measure.kernels[p]) for p in self.ps]
        return "MultiOutputKernel({})".format(", ".join(ks))


class CrossKernel(Kernel):
    """Covariance between ``cross(*ps)`` and another process ``j`` of the measure: ``mok(x, FDD(j, y))``
    (left rule, ``right=True``) or ``mok(FDD(i, x), y)`` (``stheno/model/measure.py:416-422``)."""

    symmetric = False

    def __init__(self, mok, other, right):
        self.mok, self.other, self.right = mok, other, right

    def _wrap(self, x, y, same):
        from ..model.fdd import FDD

        if self.right:
            return x, FDD(self.other, x if same else y)
        return FDD(self.other, x), (x if same else y)

    def _matrix(self, x, y, same):
        xx, yy = self._wrap(x, y, same)
        return self.mok._pairwise_multi(xx, yy, False)

    def _pairwise_multi(self, x, y, same):
        return self._matrix(x, y, same)

    def _elwise_multi(self, x, y, same):
        xx, yy = self._wrap(x, y, same)
        return self.mok._elwise_multi(xx, yy, False)

    def render(self):
        return f"CrossKernel({self.mok.render()})"


def _block(rows):
    """``B.block``: assemble a dense block matrix from a grid of structured blocks."""
    org = next((b.origin for r in rows for b in r if getattr(b, "origin", None) is not None), None)
    if M.BlockDense.eligible(rows):
        # kept as a grid: K1 writes every block straight into its place when numbers are needed (matrix.BlockDense)
        return M.BlockDense(rows, org)
    dense_rows = [torch.cat([M.dense(b) for b in r], dim=-1) for r in rows]
    return M.Dense(torch.cat(dense_rows, dim=-2), org)


def mo_pairwise(k, x, y, same):
    """``pairwise`` for tuple / FDD inputs or multi-output kernels (``stheno/mo/input.py:7-19``)."""
    if isinstance(x, tuple) or isinstance(y, tuple):
        xs = x if isinstance(x, tuple) else (x,)
        ys = y if isinstance(y, tuple) else (y,)
        rows = []
        for i, xi in enumerate(xs):
            row = []
            for j, yj in enumerate(ys):
                row.append(pairwise(k, xi, None if (same and i == j) else yj))
            rows.append(row)
        return _block(rows)
    if hasattr(k, "_pairwise_multi"):
        return k._pairwise_multi(x, y, same)
    return k._matrix(x, y, same)  # kernels that forward arbitrary inputs to their parts (Posterior*, Sum, ...)


def mo_elwise_dev(k, x, y, same):
    from ..kernels import _elwise_any

    if isinstance(x, tuple) or isinstance(y, tuple):
        xs = x if isinstance(x, tuple) else (x,)
        ys = y if isinstance(y, tuple) else (y,)
        if len(xs) != len(ys):
            raise ValueError('"elwise" must be called with similarly sized tuples.')
        return torch.cat([_elwise_any(k, xi, yi, same and xi is yi) for xi, yi in zip(xs, ys)], dim=-2)
    if hasattr(k, "_elwise_multi"):
        return k._elwise_multi(x, y, same)
    return k._elwise_dev(x, y, same)


def mo_elwise(k, x, y, same):
    from ..kernels import _origin_of_input

    return from_dev(mo_elwise_dev(k, x, y, same), _origin_of_input(x))


class MultiOutputMean(Mean):
    """``MultiOutputMean(measure, *ps)`` (``stheno/mo/mean.py:9-46``)."""

    def __init__(self, measure, *ps):
        self.measure, self.ps = measure, ps

    def dev(self, x):
        from ..model.fdd import FDD

        if isinstance(x, FDD):
            return self.measure.means[x.p].dev(x.x)
        if isinstance(x, tuple):
            return torch.cat([self.dev(xi) for xi in x], dim=-2)
        return self.dev(tuple(p(x) for p in self.ps))

    def render(self):
        ms = [str(self.measure.means[p]) for p in self.ps]
        return "MultiOutputMean({})".format(", ".join(ms))
