"""``Measure``: the joint model -- means and kernels of all processes with covariance-propagation rules
(``stheno/model/measure.py:25-492``; hot-path subset: independent GPs, sum, scalar multiple, stretch, conditioning,
cross, logpdf, sample)."""
from types import FunctionType

import numpy as np
import torch

from ..kernels import FunctionScaledKernel, ZeroKernel, num_elements
from ..lazy import LazyMatrix, LazyVector
from .._util import from_dev
from .fdd import FDD, _input_meta
from .gp import GP, assert_same_measure
from .observations import AbstractObservations, AbstractPseudoObservations, Observations, combine

__all__ = ["Measure"]


def _is_numeric(v):
    return isinstance(v, (int, float, np.number, np.ndarray, torch.Tensor))


def _left_scaled(k, f):
    """``TensorProductKernel(f, ones) * k``: ``f(x) k(x, y)`` (``stheno/model/measure.py:250``)."""
    return k if isinstance(k, ZeroKernel) else FunctionScaledKernel(k, f, None)


class Measure:
    default = None

    def __init__(self):
        self.ps = []
        self._pids = set()
        self.means = LazyVector()
        self.kernels = LazyMatrix()
        self._gps_by_name = {}
        self._names_by_gp = {}
        self._prev_default = None

    def __enter__(self):
        self._prev_default = Measure.default
        Measure.default = self
        return self

    def __exit__(self, exc_type, exc_val, exc_tb):
        Measure.default = self._prev_default

    def __hash__(self):
        return id(self)

    def __eq__(self, other):
        return self is other

    def __getitem__(self, key):
        if isinstance(key, str):
            return self._gps_by_name[key]
        return self._names_by_gp[id(key)]

    def name(self, p, name):
        if id(p) in self._names_by_gp:
            del self._gps_by_name[self._names_by_gp[id(p)]]
            del self._names_by_gp[id(p)]
        if name in self._gps_by_name:
            raise RuntimeError(f'Name "{name}" for "{p}" already taken by "{self[name]}".')
        self._gps_by_name[name] = p
        self._names_by_gp[id(p)] = name

    def _add_p(self, p):
        self.ps.append(p)
        self._pids.add(id(p))
        p._measures.append(self)

    def _update(self, p, mean, kernel, left_rule, right_rule=None):
        self.means[p] = mean
        self.kernels[p] = kernel
        self.kernels.add_left_rule(id(p), self._pids, left_rule)
        if right_rule:
            self.kernels.add_right_rule(id(p), self._pids, right_rule)
        else:
            self.kernels.add_right_rule(id(p), self._pids, lambda i: self.kernels[p, i].reversed())
        self._add_p(p)
        return p

    def add_gp(self, mean, kernel, left_rule, right_rule=None):
        p = GP()
        self._update(p, mean, kernel, left_rule, right_rule)
        return p

    def __call__(self, arg):
        """``measure(p)``: a copy of ``p`` under this measure; ``measure(fdd)``: the FDD under this measure
        (``measure.py:139-154``)."""
        if isinstance(arg, FDD):
            return self(arg.p)(arg.x, arg.noise)
        p = arg
        p_copy = GP()
        return self._update(
            p_copy,
            self.means[p],
            self.kernels[p],
            lambda j: self.kernels[p, j],
            lambda i: self.kernels[i, p],
        )

    def add_independent_gp(self, p, mean, kernel):
        self.means[p] = mean
        self.kernels[p] = kernel
        self.kernels.add_left_rule(id(p), self._pids, lambda j: ZeroKernel())
        self.kernels.add_right_rule(id(p), self._pids, lambda i: ZeroKernel())
        self._add_p(p)
        return p

    def sum(self, p_sum, a, b):
        """``measure.py:180-216``."""
        if not isinstance(a, GP):
            a, b = b, a
        if isinstance(b, GP):
            p1, p2 = a, b
            assert_same_measure(p1, p2)
            return self._update(
                p_sum,
                self.means[p1] + self.means[p2],
                (self.kernels[p1] + self.kernels[p2] + self.kernels[p1, p2] + self.kernels[p2, p1]),
                lambda j: self.kernels[p1, j] + self.kernels[p2, j],
            )
        p, other = a, b
        return self._update(p_sum, self.means[p] + other, self.kernels[p], lambda j: self.kernels[p, j])

    def mul(self, p_mul, a, b):
        """``measure.py:218-270``: scalar multiples, ``GP * function`` and the moment-matched ``GP * GP``."""
        if not isinstance(a, GP):
            a, b = b, a
        p, other = a, b
        if isinstance(other, GP):
            p1, p2 = p, other
            assert_same_measure(p1, p2)
            m1, m2 = self.means[p1], self.means[p2]
            mean_fn = lambda m: (lambda x: m.dev(x))  # the mean as a plain function of the points
            term1 = self.sum(GP(), self.mul(GP(), mean_fn(m1), p2), self.mul(GP(), p1, mean_fn(m2)))
            term2 = self.add_independent_gp(
                GP(),
                -(m1 * m2),
                self.kernels[p1] * self.kernels[p2] + self.kernels[p1, p2] * self.kernels[p2, p1],
            )
            return self.sum(p_mul, term1, term2)
        if isinstance(other, FunctionType):
            f = other
            return self._update(
                p_mul,
                f * self.means[p],
                f * self.kernels[p],
                lambda j: _left_scaled(self.kernels[p, j], f),
            )
        return self._update(
            p_mul,
            self.means[p] * other,
            self.kernels[p] * other**2,
            lambda j: self.kernels[p, j] * other,
        )

    def stretch(self, p_stretched, p, stretch):
        """``measure.py:289-305``."""
        return self._update(
            p_stretched,
            self.means[p].stretch(stretch),
            self.kernels[p].stretch(stretch),
            lambda j: self.kernels[p, j].stretch(stretch, 1),
        )

    def shift(self, p_shifted, p, shift):
        """``measure.py:272-287``."""
        return self._update(
            p_shifted,
            self.means[p].shift(shift),
            self.kernels[p].shift(shift),
            lambda j: self.kernels[p, j].shift(shift, 0),
        )

    def select(self, p_selected, p, *dims):
        """``measure.py:307-325``."""
        return self._update(
            p_selected,
            self.means[p].select(dims),
            self.kernels[p].select(dims),
            lambda j: self.kernels[p, j].select(dims, None),
        )

    def transform(self, p_transformed, p, f):
        """``measure.py:327-345``."""
        return self._update(
            p_transformed,
            self.means[p].transform(f),
            self.kernels[p].transform(f),
            lambda j: self.kernels[p, j].transform(f, None),
        )

    def diff(self, p_diff, p, dim=0):
        """``measure.py:343-360``: the derivative process, its kernel ``d^2 k / dx dy`` and cross-kernels ``dk / dx``."""
        return self._update(
            p_diff,
            self.means[p].diff(dim),
            self.kernels[p].diff(dim),
            lambda j: self.kernels[p, j].diff(dim, None),
        )

    def condition(self, *args):
        """``measure | obs`` -> posterior measure whose means / kernels are built on demand (``measure.py:362-401``)."""
        if len(args) == 1 and isinstance(args[0], AbstractObservations):
            obs = args[0]
        elif len(args) == 2 and isinstance(args[0], FDD):
            obs = Observations(args[0], args[1])
        elif len(args) == 1 and isinstance(args[0], tuple):
            obs = Observations(*args[0])
        else:
            obs = Observations(*args)
        posterior = Measure()
        posterior.ps = list(self.ps)
        posterior._pids = set(self._pids)
        posterior.means.add_rule(posterior._pids, lambda i: obs.posterior_mean(self, i))
        posterior.kernels.add_rule(posterior._pids, lambda i, j: obs.posterior_kernel(self, i, j))
        for p in posterior.ps:
            p._measures.append(posterior)
        return posterior

    def __or__(self, args):
        if isinstance(args, tuple):
            return self.condition(*args)
        return self.condition(args)

    def cross(self, p_cross, *ps):
        """``measure.py:403-423``."""
        from ..mo.kernel import CrossKernel, MultiOutputKernel, MultiOutputMean

        mok = MultiOutputKernel(self, *ps)
        return self._update(
            p_cross,
            MultiOutputMean(self, *ps),
            mok,
            lambda j: CrossKernel(mok, j, right=True),
            lambda i: CrossKernel(mok, i, right=False),
        )

    def sample(self, *args):
        """``measure.sample([state,] [n,] *fdds)`` -- joint samples of several FDDs (``measure.py:425-461``)."""
        state, n = None, 1
        args = list(args)
        if args and isinstance(args[0], torch.Generator):
            state = args.pop(0)
        if args and isinstance(args[0], int):
            n = args.pop(0)
        fdds = args
        joint = self(combine(*fdds))
        res = joint.sample(state, n) if state is not None else joint.sample(n)
        if state is not None:
            state, sample = res
        else:
            sample = res
        lengths = [num_elements(fdd) for fdd in fdds]
        i, samples = 0, []
        for length in lengths:
            samples.append(sample[..., i : i + length, :])
            i += length
        if state is not None:
            return (state,) + tuple(samples)
        return samples[0] if len(samples) == 1 else tuple(samples)

    def logpdf(self, *args):
        """``measure.logpdf(fdd, y)``, ``measure.logpdf((fdd1, y1), (fdd2, y2))``, ``measure.logpdf(obs)``
        (sparse observations give the ELBO) (``measure.py:463-489``)."""
        if len(args) == 1 and isinstance(args[0], AbstractPseudoObservations):
            return args[0].elbo(self)
        if len(args) == 1 and isinstance(args[0], Observations):
            obs = args[0]
            return self(obs.fdd).logpdf(obs.y)
        if len(args) == 2 and isinstance(args[0], FDD):
            return self(args[0]).logpdf(args[1])
        fdd, y = combine(*args)
        return self(fdd).logpdf(y)
