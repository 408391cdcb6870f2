"""Conditioning: exact ``Observations`` and the sparse ``PseudoObservations`` family
(``stheno/model/observations.py:28-414``)."""
import numpy as np
import torch

from .. import B
from .. import matrix as M
from .. import ops
from ..kernels import (PosteriorKernel, PosteriorMean, SubspaceKernel, _cross_rows, _elwise_any, num_elements,
                       pairwise)
from .._util import batch_flatten, from_dev, to_dev, uprank
from .fdd import FDD, _input_meta
from .gp import cross

__all__ = [
    "combine", "AbstractObservations", "AbstractPseudoObservations", "Observations", "Obs", "PseudoObservations",
    "SparseObservations", "PseudoObs", "SparseObs", "PseudoObservationsFITC", "PseudoObsFITC",
    "PseudoObservationsDTC", "PseudoObsDTC",
]


def combine(*args):
    """Combine FDDs -- or ``(fdd, y)`` pairs -- into one joint FDD (and stacked ``y``) (``observations.py:28-47``)."""
    if all(isinstance(a, FDD) for a in args):
        fdds = args
        combined_noise = M.block_diag(*[fdd.noise for fdd in fdds])
        return cross(*[fdd.p for fdd in fdds])(tuple(fdds), combined_noise)
    fdds, ys = zip(*args)
    combined_fdd = combine(*fdds)
    dtype = _input_meta(combined_fdd.x)[0]
    combined_y = torch.cat([uprank(to_dev(y, dtype)) for y in ys], dim=-2)
    return combined_fdd, combined_y


class AbstractObservations:
    def __init__(self, *args):
        if len(args) == 2 and isinstance(args[0], FDD):
            fdd, y = args
        else:
            fdd, y = combine(*args)
        y_shape = tuple(np.shape(y)) if not isinstance(y, torch.Tensor) else tuple(y.shape)
        y = uprank(to_dev(y, _input_meta(fdd.x)[0]))
        if y.shape[-1] != 1:
            raise ValueError(f"Invalid shape of observed values {y_shape}.")
        # Missing data: one device reduction + flag; the gather path only runs when NaNs exist.
        nan = torch.isnan(y[..., :, 0])
        if bool(nan.any()):
            avail = ~nan
            fdd = fdd.take(avail)
            y = y[avail]
        self.fdd = fdd
        self.y = y

    def posterior_kernel(self, measure, p_i, p_j):  # pragma: no cover
        raise NotImplementedError("Posterior kernel construction not implemented.")

    def posterior_mean(self, measure, p):  # pragma: no cover
        raise NotImplementedError("Posterior mean construction not implemented.")


class Observations(AbstractObservations):
    """Exact observations ``(f(x, noise), y)``.  The factor of ``K_x`` is computed once per measure and reused by
    every prediction (``observations.py:127-141``); ``L^-1 (y - m(x))`` rides along in that factorisation."""

    def __init__(self, *args):
        AbstractObservations.__init__(self, *args)
        self._K_x = {}

    def K_x(self, measure):
        try:
            return self._K_x[id(measure)]
        except KeyError:
            K_x = M.add(pairwise(measure.kernels[self.fdd.p], self.fdd.x), self.fdd.noise)
            K_x = M._densify(K_x, full=True)  # low-rank structure is exploited by logpdf; conditioning uses the dense factor
            if isinstance(K_x, M.KernelDense):
                K_x.full_precision = True  # posterior means / variances are read element-wise off this factor
            if isinstance(K_x, M.Dense):
                diff = self.y - measure.means[self.fdd.p].dev(self.fdd.x)
                d3, _ = batch_flatten(diff, 2)
                K_x.attach_rhs(("obs", id(self)), d3.transpose(1, 2).contiguous())
            self._K_x[id(measure)] = K_x
            return K_x

    def posterior_kernel(self, measure, p_i, p_j):
        if num_elements(self.fdd.x) == 0:
            return measure.kernels[p_i, p_j]
        return PosteriorKernel(
            measure.kernels[p_i, p_j],
            measure.kernels[self.fdd.p, p_i],
            measure.kernels[self.fdd.p, p_j],
            self.fdd.x,
            self.K_x(measure),
        )

    def posterior_mean(self, measure, p):
        if num_elements(self.fdd.x) == 0:
            return measure.means[p]
        return PosteriorMean(
            measure.means[p],
            measure.means[self.fdd.p],
            measure.kernels[self.fdd.p, p],
            self.fdd.x,
            self.K_x(measure),
            self.y,
            rhs_key=("obs", id(self)),
        )


class AbstractPseudoObservations(AbstractObservations):
    """Observations through inducing points ``u`` (VFE / FITC / DTC), ``observations.py:171-336``."""

    method = None

    def __init__(self, u, *args):
        AbstractObservations.__init__(self, *args)
        if isinstance(u, tuple):
            u = combine(*u)
        self.u = u
        self._K_z, self._elbo, self._mu, self._A = {}, {}, {}, {}

    def _get(self, store, measure):
        try:
            return store[id(measure)]
        except KeyError:
            self._compute(measure)
            return store[id(measure)]

    def K_z(self, measure):
        return self._get(self._K_z, measure)

    def elbo(self, measure):
        e = self._get(self._elbo, measure)
        return from_dev(e, _input_meta(self.fdd.x)[2])

    def mu(self, measure):
        return self._get(self._mu, measure)

    def A(self, measure):
        return self._get(self._A, measure)

    def posterior_kernel(self, measure, p_i, p_j):
        return PosteriorKernel(
            measure.kernels[p_i, p_j],
            measure.kernels[self.u.p, p_i],
            measure.kernels[self.u.p, p_j],
            self.u.x,
            self.K_z(measure),
        ) + SubspaceKernel(
            measure.kernels[self.u.p, p_i],
            measure.kernels[self.u.p, p_j],
            self.u.x,
            self.A(measure),
        )

    def posterior_mean(self, measure, p):
        return PosteriorMean(
            measure.means[p],
            measure.means[self.u.p],
            measure.kernels[self.u.p, p],
            self.u.x,
            self.K_z(measure),
            self.mu(measure),
        )

    def _compute(self, measure):
        """``observations.py:279-336`` with ``W^T = K_xz L_z^-T`` kept in row form ``[n, m]`` (rows = data points):
        the m^2 n flops of the solve and of ``A = I + W K_n^-1 W^T`` both run on the tensor-core GEMM."""
        p_x, x, noise_x = self.fdd.p, self.fdd.x, self.fdd.noise
        p_z, z, noise_z = self.u.p, self.u.x, self.u.noise
        if self._wants_grad(measure):
            return self._compute_grad(measure)
        K_z = M.add(pairwise(measure.kernels[p_z], z), noise_z)  # :286
        self._K_z[id(measure)] = K_z
        K_n = noise_x  # :290
        if not isinstance(K_n, M.Diagonal):
            raise RuntimeError(
                f'Kernel matrix of observation noise must be diagonal, not "{type(K_n).__name__}".'
            )
        K_z = M._densify(K_z)
        ch_z = K_z.chol()  # :300
        m, m_pad = ch_z.n, ch_z.n_pad
        mean_z = measure.means[p_z].dev(z)
        y_bar = uprank(self.y) - measure.means[p_x].dev(x)
        yb3, _ = batch_flatten(y_bar, 2)  # [B, n, 1]
        kn = K_n.diag
        kn3 = kn.reshape(-1, kn.shape[-1])
        if kn3.shape[0] != ch_z.batch:
            kn3 = kn3.expand(ch_z.batch, -1)
        streamed = self._stream_plan(measure, ch_z)
        if streamed is not None:
            A, prod, det_kn, yky, trace_part = self._accumulate_streamed(measure, ch_z, kn3, yb3, *streamed)
        else:
            A, prod, det_kn, yky, trace_part = self._accumulate_materialised(measure, ch_z, kn3, yb3)
        ops.symmetrize_(A, m_pad)
        A_mat = M.Dense(A[:, :m, :m])
        ch_A = A_mat.chol()
        half = ch_A.half_solve(prod.unsqueeze(1))  # [B, 1, m]  L_A^-1 prod
        sol = ch_A.full_solve(prod.unsqueeze(1))  # A^-1 prod
        Lz_pad = ch_z.L_lower_()  # strict upper triangle zeroed in place (no copy); identity on the padding
        solp = torch.zeros(ch_z.batch, ops.TILE, m_pad, dtype=A.dtype, device=A.device)
        solp[:, :1, :m] = sol
        mu_rows = ops.gemm_nt(solp, Lz_pad)  # row 0 = (L_z A^-1 prod)^T
        mu = mean_z + mu_rows[:, 0, :m].reshape(mean_z.shape[:-1]).unsqueeze(-1)  # :329
        self._mu[id(measure)] = mu
        # stored "A" = L_z A L_z^T (:323) as two NT products (A is symmetric)
        U = ops.gemm_nt(Lz_pad, A)
        LAL = ops.gemm_nt(U, Lz_pad)
        self._A[id(measure)] = M.Dense(LAL[:, :m, :m].reshape(K_z.shape), K_z.origin)
        # ELBO (:333-336)
        det_part = det_kn + ch_A.logdet
        iqf_part = yky - (half * half).sum((-1, -2))
        elbo = -0.5 * (det_part + iqf_part + trace_part)
        bs = K_z.shape[:-2]
        self._elbo[id(measure)] = elbo.reshape(bs) if bs else elbo[0]


    # -- differentiable route (generic_grad.py): used only when something that feeds the ELBO requires grad -----------------
    def _wants_grad(self, measure):
        from ..generic_grad import kernel_needs_grad
        from ..kernels import Input

        if not torch.is_grad_enabled():
            return False
        p_x, p_z = self.fdd.p, self.u.p
        ks = [measure.kernels[p_z], measure.kernels[p_z, p_x], measure.kernels[p_x]]
        if any(kernel_needs_grad(k) for k in ks):
            return True
        ts = [self.y]
        for v in (self.fdd.x, self.u.x):
            if isinstance(v, Input):
                ts.append(v.t)
        for nz in (self.fdd.noise, self.u.noise):
            if isinstance(nz, M.Diagonal):
                ts.append(nz.diag)
        return any(isinstance(t, torch.Tensor) and t.requires_grad for t in ts)

    def _compute_grad(self, measure):
        from ..generic_grad import sparse_compute_torch
        from ..kernels import Input

        p_x, x, K_n = self.fdd.p, self.fdd.x, self.fdd.noise
        p_z, z, noise_z = self.u.p, self.u.x, self.u.noise
        if not isinstance(K_n, M.Diagonal):
            raise RuntimeError(
                f'Kernel matrix of observation noise must be diagonal, not "{type(K_n).__name__}".'
            )
        if not isinstance(x, Input) or not isinstance(z, Input):
            raise NotImplementedError("gradients of a sparse approximation over multi-output inputs are not implemented")
        if isinstance(noise_z, M.Zero):
            nz = None
        elif isinstance(noise_z, M.Diagonal):
            nz = noise_z.diag
        else:
            raise NotImplementedError("gradients with a dense inducing-point noise are not implemented")
        y_bar = uprank(self.y) - measure.means[p_x].dev(x)
        K_z, LAL, mu, elbo = sparse_compute_torch(
            self.method, measure.kernels[p_z], measure.kernels[p_z, p_x], measure.kernels[p_x], z.t, x.t, K_n.diag, nz,
            y_bar, measure.means[p_z].dev(z), B.epsilon)
        self._K_z[id(measure)] = M.Dense(K_z, z.origin)
        self._mu[id(measure)] = mu
        self._A[id(measure)] = M.Dense(LAL, z.origin)
        self._elbo[id(measure)] = elbo


    # -- the two ways to form A = I + W K_n^-1 W^T, prod = W K_n^-1 ybar and the scalars ------------------------------------
    def _stream_plan(self, measure, ch_z):
        """``(flat, scales, x_input, z_input)`` when the problem can be streamed (one problem, numeric inputs, a symmetric
        cross-kernel that fits one K1 descriptor), else None."""
        from ..kernels import Input, _is_multi

        p_x, x, p_z, z = self.fdd.p, self.fdd.x, self.u.p, self.u.x
        if ch_z.batch != 1 or _is_multi(x) or _is_multi(z) or not isinstance(x, Input) or not isinstance(z, Input):
            return None
        if x.batch_shape or z.batch_shape:
            return None
        k_zx = measure.kernels[p_z, p_x]
        if not k_zx.symmetric:
            return None
        flat, scales = k_zx._flat()
        if flat is None or not flat.terms:
            return None
        return flat, scales, x, z

    def _accumulate_streamed(self, measure, ch_z, kn3, yb3, flat, scales, x, z):
        """``gpk_sparse_accumulate`` over chunks of data points: O(chunk m + m^2) device memory."""
        acc = ops.SparseAccumulator(flat, z.scaled(scales), ch_z, self.method, chunk=B.sparse_chunk)
        xg = x.scaled(scales)  # [G, 1, n, d]
        n = x.n
        kd = None
        if self.method in ("vfe", "fitc"):
            kd = _elwise_any(measure.kernels[self.fdd.p], x, None, True)[..., 0].reshape(-1)  # :304
        kn1, yb1 = kn3[0], yb3[0, :, 0]
        for a in range(0, n, acc.chunk):
            b_ = min(n, a + acc.chunk)
            acc.add(xg[:, :, a:b_], None if kd is None else kd[a:b_], kn1[a:b_], yb1[a:b_])
        sc = acc.scalars
        return acc.A, acc.prod[: ch_z.n].unsqueeze(0), sc[0].reshape(1), sc[1].reshape(1), sc[2].reshape(1)

    def _accumulate_materialised(self, measure, ch_z, kn3, yb3):
        """Batched / multi-output / non-flattenable problems: ``W^T = K_xz L_z^-T`` held as one ``[B, n_pad, m_pad]`` buffer."""
        p_x, x = self.fdd.p, self.fdd.x
        p_z, z = self.u.p, self.u.x
        m, m_pad = ch_z.n, ch_z.n_pad
        Wt, n = _cross_rows(measure.kernels[p_z, p_x], z, x, ch_z)  # :285
        ch_z.solve_rows_(Wt)  # :301
        trace_part = torch.zeros(ch_z.batch, dtype=Wt.dtype, device=Wt.device)
        if self.method in ("vfe", "fitc"):
            K_x_diag = _elwise_any(measure.kernels[p_x], x, None, True)[..., 0]  # :304
            _, Q_x_diag = ops.row_dot_sq(Wt, n, m_pad, None)  # :305
            corr = K_x_diag.reshape(ch_z.batch, n) - Q_x_diag  # :306
            if self.method == "vfe":
                trace_part = (corr / kn3).sum(-1)  # :308-310
            else:
                kn3 = kn3 + corr  # :311-313
        n_pad = Wt.shape[1]
        rs = torch.rsqrt(kn3)
        Wt[:, :n] *= rs.unsqueeze(-1)
        WsT = torch.empty(ch_z.batch, m_pad, n_pad, dtype=Wt.dtype, device=Wt.device)
        ops.transpose(Wt, n_pad, m_pad, out=WsT)
        del Wt
        A = torch.zeros(ch_z.batch, m_pad, m_pad, dtype=WsT.dtype, device=WsT.device)
        A.diagonal(dim1=1, dim2=2).fill_(1.0)
        ops.gemm_nt(WsT, WsT, A, alpha=1.0, beta=1.0, lower=True)  # :322
        ybs = torch.zeros(ch_z.batch, n_pad, dtype=WsT.dtype, device=WsT.device)
        ybs[:, :n] = yb3[..., 0] * rs
        prod, _ = ops.row_dot_sq(WsT, m, n_pad, ybs, want_sq=False)  # :327
        det_kn = torch.log(2 * B.pi * kn3).sum(-1)
        yky = (yb3[..., 0] ** 2 / kn3).sum(-1)
        return A, prod, det_kn, yky, trace_part

class PseudoObservations(AbstractPseudoObservations):
    """VFE (Titsias, 2009)."""

    method = "vfe"


class PseudoObservationsFITC(AbstractPseudoObservations):
    """FITC (Snelson & Ghahramani, 2006)."""

    method = "fitc"


class PseudoObservationsDTC(AbstractPseudoObservations):
    """DTC (Csato & Opper, 2002; Seeger et al., 2003)."""

    method = "dtc"


Obs = Observations
PseudoObs = PseudoObservations
PseudoObsFITC = PseudoObservationsFITC
PseudoObsDTC = PseudoObservationsDTC
SparseObs = PseudoObservations
SparseObservations = PseudoObservations
