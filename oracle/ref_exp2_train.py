"""ORACLE (test infrastructure only) -- exp2 ``train()``: the taped two-level step and its gradient with respect to the
interpolation weights, /root/reference/exp2/p_hat/quasi_simulation/object.py:1202-1287 (kernels: train_hexagon.py,
temp_kernel.py, cublas.py:93-102,156-166,259-292 there).

The reference differentiates the step with ``wp.Tape``; Warp is not installed here and its tape cannot be emulated kernel by
kernel, so this file restates the FORWARD map statement for statement in torch (the same physics the numpy oracle
``ref_kernels`` restates, which is pinned by the reference-made ``ref_exp2.npz``) and lets torch.autograd produce the gradient.
PARITY UNPINNED for the gradient itself: no reference execution produces it here; it is held instead to central differences
of this forward map in float64 (tests/test_oracle.py) -- the derivative of a pinned function.  Conventions taken from the
reference's kernels: the loss is ``atomic_max`` of |residual component| (one-hot adjoint at the arg-max) minus
``alpha (1 - row sum)^2`` over the rows of w with |1 - row sum| >= 1e-6; the update is ``w += 1e-3 * grad`` clamped to
[0, 1] (``Valueaxpy``); ``solve3x3`` is the 3-step CG with its early exits, differentiated as executed."""
from __future__ import annotations

import numpy as np
import torch


def _t(a, dtype):
    return torch.as_tensor(np.asarray(a), dtype=dtype)


def det3(m):
    return (m[..., 0, 0] * (m[..., 1, 1] * m[..., 2, 2] - m[..., 1, 2] * m[..., 2, 1])
            + m[..., 0, 1] * (m[..., 1, 2] * m[..., 2, 0] - m[..., 1, 0] * m[..., 2, 2])
            + m[..., 0, 2] * (m[..., 1, 0] * m[..., 2, 1] - m[..., 1, 1] * m[..., 2, 0]))


def inv3(m):
    d = det3(m)
    c = torch.stack([
        torch.stack([m[..., 1, 1] * m[..., 2, 2] - m[..., 1, 2] * m[..., 2, 1], m[..., 0, 2] * m[..., 2, 1] - m[..., 0, 1] * m[..., 2, 2],
                     m[..., 0, 1] * m[..., 1, 2] - m[..., 0, 2] * m[..., 1, 1]], -1),
        torch.stack([m[..., 1, 2] * m[..., 2, 0] - m[..., 1, 0] * m[..., 2, 2], m[..., 0, 0] * m[..., 2, 2] - m[..., 0, 2] * m[..., 2, 0],
                     m[..., 0, 2] * m[..., 1, 0] - m[..., 0, 0] * m[..., 1, 2]], -1),
        torch.stack([m[..., 1, 0] * m[..., 2, 1] - m[..., 1, 1] * m[..., 2, 0], m[..., 0, 1] * m[..., 2, 0] - m[..., 0, 0] * m[..., 2, 1],
                     m[..., 0, 0] * m[..., 1, 1] - m[..., 0, 1] * m[..., 1, 0]], -1)], -2)
    return c / d[..., None, None]


def F0(x, hexs, sfg):
    """sum_a x_a (x) dN_a(q): [Nh, q, 3, 3]   (train_hexagon.py:10-11, temp_kernel.py)."""
    return torch.einsum("ear,aqc->eqrc", x[hexs], sfg)


def solve3x3(A, b):
    """cublas.py:59-91 from x = 0, vectorised over the rows; masks reproduce the early exits."""
    x = torch.zeros_like(b)
    r = b
    rn = (r * r).sum(-1)
    active = rn >= 1e-10
    p = r
    for _ in range(3):
        Ap = torch.einsum("nij,nj->ni", A, p)
        dot = (p * Ap).sum(-1)
        act = active & (dot >= 1e-10)
        alpha = torch.where(act, rn / torch.where(act, dot, torch.ones_like(dot)), torch.zeros_like(dot))
        x = x + alpha[:, None] * p
        r = r - alpha[:, None] * Ap
        old = rn
        rn = torch.where(act, (r * r).sum(-1), rn)
        active = act & (rn >= 1e-10)
        beta = torch.where(active, rn / torch.where(active, old, torch.ones_like(old)), torch.zeros_like(rn))
        p = torch.where(active[:, None], r + beta[:, None] * p, p)
    return x


class Exp2Train:
    """The taped step as a function of w.  ``ref`` is an oracle.ref_exp2.RefExp2Object (its integer maps and fixed operators)."""

    def __init__(self, ref, dtype=torch.float64, alpha=0.1, lr=1e-3):
        self.ref, self.dtype, self.alpha, self.lr = ref, dtype, alpha, lr
        self.mu, self.la = float(ref.mu), float(ref.la)
        self.sfg = _t(ref.sfg, dtype)                                   # [a, q, 3]
        self.hex0 = torch.as_tensor(np.asarray(ref.hexs[0]), dtype=torch.long)
        self.hex1 = torch.as_tensor(np.asarray(ref.hexs[1]), dtype=torch.long)
        self.det0, self.inv0 = _t(ref.det[0], dtype), _t(ref.inv[0], dtype)
        ut = ref.Ut_noOrder[0]
        off = np.asarray(ut["offsets"])
        self.ut_rows = torch.as_tensor(np.repeat(np.arange(len(off) - 1), np.diff(off)), dtype=torch.long)
        self.ut_cols = torch.as_tensor(np.asarray(ut["columns"]), dtype=torch.long)
        self.ut_offsets = off
        self.n1 = len(off) - 1
        self.x_init = _t(ref.dev_x_init, dtype)
        self.X = _t(ref.X, dtype)
        self.v2i0 = torch.as_tensor(np.asarray(ref.v2i[0]), dtype=torch.long)
        self.i2v0 = torch.as_tensor(np.asarray(ref.i2v[0]), dtype=torch.long)
        self.v2i1 = torch.as_tensor(np.asarray(ref.v2i[1]), dtype=torch.long)

        def csr(m):
            o = np.asarray(m["offsets"])
            w = np.asarray(m["values"])
            w = w[:, 0, 0] if w.ndim == 3 else w
            return (torch.as_tensor(np.repeat(np.arange(len(o) - 1), np.diff(o)), dtype=torch.long),
                    torch.as_tensor(np.asarray(m["columns"]), dtype=torch.long), _t(w, dtype), len(o) - 1)
        self.ut_hat = csr(ref.Ut_hat[0])
        self.us_hat = csr(ref.Us_hat[0])

    @staticmethod
    def _mv(m, x):
        rows, cols, w, n = m
        y = torch.zeros((n, 3), dtype=x.dtype)
        return y.index_add(0, rows, w[:, None] * x[cols])

    def force(self, x, hexs, det, inv):
        """Elastic nodal forces in vertex order (train_hexagon.py compute_partial_elastic_energy_X: -(P g_b) det)."""
        F = F0(x, hexs, self.sfg) @ inv
        E = 0.5 * (F.transpose(-1, -2) @ F - torch.eye(3, dtype=x.dtype))
        S = 2.0 * self.mu * E + self.la * torch.einsum("eqii->eq", E)[..., None, None] * torch.eye(3, dtype=x.dtype)
        P = F @ S
        g = torch.einsum("aqc,eqcj->eqaj", self.sfg, inv)               # g_a = inv^T dN_a
        fe = -torch.einsum("eqrj,eqaj->eqar", P, g) * det[:, :, None, None]
        out = torch.zeros_like(x)
        return out.index_add(0, hexs.reshape(-1), fe.sum(1).reshape(-1, 3))

    def coarse_diag(self, x1, det, inv):
        """Diagonal blocks (vertex order) of the coarse Hessian: the only blocks jacobi_iteration reads."""
        F = F0(x1, self.hex1, self.sfg) @ inv
        I = torch.eye(3, dtype=x1.dtype)
        E = 0.5 * (F.transpose(-1, -2) @ F - I)
        S = 2.0 * self.mu * E + self.la * torch.einsum("eqii->eq", E)[..., None, None] * I
        FFT = F @ F.transpose(-1, -2)
        g = torch.einsum("aqc,eqcj->eqaj", self.sfg, inv)
        u = torch.einsum("eqrj,eqaj->eqar", F, g)
        t = torch.einsum("eqij,eqaj->eqai", S, g) * det[:, :, None, None]
        s1 = (t * g).sum(-1)                                            # t_a . g_a
        s2 = (g * g).sum(-1)
        cmu, cla = det * self.mu, det * self.la
        K = (s1[..., None, None] * I + (cmu[:, :, None] * s2)[..., None, None] * FFT[:, :, None]
             + (cmu + cla)[:, :, None, None, None] * u[..., :, None] * u[..., None, :])
        D = torch.zeros((self.n1, 3, 3), dtype=x1.dtype)
        return D.index_add(0, self.hex1.reshape(-1), K.sum(1).reshape(-1, 3, 3))

    def forward(self, w, x0):
        """loss(w) of one training step started from the fine state x0 (vertex order); also returns the pieces a test wants."""
        ref = self.ref
        r0 = self.force(x0, self.hex0, self.det0, self.inv0)            # outside the tape in the reference (no w in it)
        b1 = self._mv(self.ut_hat, r0[self.i2v0])                       # index order
        xs1 = torch.zeros((self.n1, 3), dtype=w.dtype).index_add(0, self.ut_rows, w[:, None] * self.x_init[self.ut_cols])
        J = F0(xs1, self.hex1, self.sfg)
        det1, inv1 = det3(J), inv3(J)
        x1 = torch.zeros((self.n1, 3), dtype=w.dtype).index_add(0, self.ut_rows, w[:, None] * self.X[self.ut_cols])
        D = self.coarse_diag(x1, det1, inv1)
        D_index = torch.zeros_like(D).index_copy(0, self.v2i1, D)
        d1 = solve3x3(D_index, b1)
        d0 = self._mv(self.us_hat, d1)                                  # index order
        xn = x0.index_add(0, self.i2v0, d0)
        res = self.force(xn, self.hex0, self.det0, self.inv0)[self.i2v0]
        rowsum = torch.zeros(self.n1, dtype=w.dtype).index_add(0, self.ut_rows, w)
        dev = 1.0 - rowsum
        reg = torch.where(dev.abs() >= 1e-6, dev * dev, torch.zeros_like(dev)).sum()
        loss = res.abs().max() - self.alpha * reg
        return loss, dict(res=res, d1=d1, D=D_index, x1=x1, xs1=xs1, xn=xn, b1=b1)

    def step(self, w_np, x0_np):
        """(loss, grad [nnz], updated w) of one iteration of train()'s loop body."""
        w = torch.tensor(np.asarray(w_np), dtype=self.dtype, requires_grad=True)
        loss, aux = self.forward(w, _t(x0_np, self.dtype))
        g, = torch.autograd.grad(loss, w)
        w_new = torch.clamp(w.detach() + self.lr * g, 0.0, 1.0)        # Valueaxpy (cublas.py:156-166)
        return float(loss.detach()), g.numpy(), w_new.numpy(), {k: v.detach().numpy() for k, v in aux.items()}
