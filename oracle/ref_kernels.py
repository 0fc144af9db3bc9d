"""ORACLE (test infrastructure only) -- numpy restatement of the reference's device kernels.

Follows, thread body by thread body, ``exp1/quasi_simulation/quai_hexagon.py`` and
``exp1/quasi_simulation/cublas.py`` (paths relative to /root/reference), vectorised over
the launch dimension.  ``dtype`` selects the float32 arithmetic of the reference or a
float64 twin (used to show both the oracle and the CUDA path sit inside fp32 round-off of
the exact result).  Summation orders of atomics are not defined by the reference.

Warp builtins restated here (warp-lang is an un-vendored dependency, version unpinned):
``wp.determinant``/``wp.inverse`` (cofactor formulas) and ``wp.eig3`` (McAdams et al.
2011 "Computing the SVD of 3x3 matrices with minimal branching", the cyclic approximate-
Givens Jacobi eigen-analysis Warp's ``svd.h`` is derived from: 4 sweeps over the lower
triangle, V from the accumulated quaternion, d = diag(V^T A V)).  ``eig3`` parity is
therefore *unpinned*; everything that depends on it says so.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
from __future__ import annotations

import numpy as np

I3 = np.eye(3)


# ----------------------------------------------------------------------------------------------
# small dense helpers (Warp builtins)
# ----------------------------------------------------------------------------------------------
def det3(m):
    return (m[..., 0, 0] * (m[..., 1, 1] * m[..., 2, 2] - m[..., 1, 2] * m[..., 2, 1])
            - m[..., 0, 1] * (m[..., 1, 0] * m[..., 2, 2] - m[..., 1, 2] * m[..., 2, 0])
            + m[..., 0, 2] * (m[..., 1, 0] * m[..., 2, 1] - m[..., 1, 1] * m[..., 2, 0]))


def inv3(m):
    d = det3(m)
    out = np.empty_like(m)
    out[..., 0, 0] = m[..., 1, 1] * m[..., 2, 2] - m[..., 1, 2] * m[..., 2, 1]
    out[..., 0, 1] = m[..., 0, 2] * m[..., 2, 1] - m[..., 0, 1] * m[..., 2, 2]
    out[..., 0, 2] = m[..., 0, 1] * m[..., 1, 2] - m[..., 0, 2] * m[..., 1, 1]
    out[..., 1, 0] = m[..., 1, 2] * m[..., 2, 0] - m[..., 1, 0] * m[..., 2, 2]
    out[..., 1, 1] = m[..., 0, 0] * m[..., 2, 2] - m[..., 0, 2] * m[..., 2, 0]
    out[..., 1, 2] = m[..., 0, 2] * m[..., 1, 0] - m[..., 0, 0] * m[..., 1, 2]
    out[..., 2, 0] = m[..., 1, 0] * m[..., 2, 1] - m[..., 1, 1] * m[..., 2, 0]
    out[..., 2, 1] = m[..., 0, 1] * m[..., 2, 0] - m[..., 0, 0] * m[..., 2, 1]
    out[..., 2, 2] = m[..., 0, 0] * m[..., 1, 1] - m[..., 0, 1] * m[..., 1, 0]
    return out / d[..., None, None]


# ----------------------------------------------------------------------------------------------
# quai_hexagon.py
# ----------------------------------------------------------------------------------------------
def _F0(x, hexagons, sfg):
    """F[row,col] = sum_i x[hex[i]][row] * shapeFuncGrad[i][q][col]   (quai_hexagon.py:10-16)."""
    xe = x[hexagons]                                   # [Nh,8,3]
    return np.einsum("ear,aqc->eqrc", xe, sfg)


def prepare_kernal(x, hexagons, sfg, dtype=np.float32):
    """quai_hexagon.py:4-18 -> (det_pX_peps [Nh,8], inverse_pX_peps [Nh,8,3,3])."""
    x = x.astype(dtype); sfg = sfg.astype(dtype)
    J = _F0(x, hexagons, sfg)
    return det3(J).astype(dtype), inv3(J).astype(dtype)


def sum_8(det):
    """cublas.py:215-218."""
    v = det[:, 0].copy()
    for k in range(1, 8):
        v = v + det[:, k]
    return v


def prepare_mass(vol, hexagons, n_vert):
    """quai_hexagon.py:20-26: the FULL element volume goes to each of its 8 nodes."""
    m = np.zeros(n_vert, dtype=vol.dtype)
    np.add.at(m, hexagons.reshape(-1), np.repeat(vol, 8))
    return m


def _FE(x, hexagons, sfg, inv):
    F = _F0(x, hexagons, sfg) @ inv                    # F = F @ inverse_pX_peps
    E = 0.5 * (np.swapaxes(F, -1, -2) @ F - np.eye(3, dtype=F.dtype))
    return F, E


def compute_elastic_energy(x, hexagons, sfg, det, inv, mu, la, dtype=np.float32):
    """quai_hexagon.py:28-55."""
    x, sfg, det, inv = (a.astype(dtype) for a in (x, sfg, det, inv))
    F, E = _FE(x, hexagons, sfg, inv)
    tr = np.trace(E, axis1=-2, axis2=-1)
    psi = (E * E).sum(axis=(-1, -2)) * dtype(mu) + dtype(0.5) * dtype(la) * tr * tr
    return (psi * det).sum(dtype=dtype)


def compute_gravity_energy(x, m, g, pin, pin_pos, control_mag, dtype=np.float32):
    """quai_hexagon.py:57-64 (pin_pos is the per-vertex all_pin_pos array)."""
    x = x.astype(dtype)
    e = -m.astype(dtype) * dtype(g) * x[:, 1]
    d = pin_pos.astype(dtype) - x
    e = e + np.where(pin == 1, dtype(0.5) * dtype(control_mag) * (d * d).sum(-1), dtype(0))
    return e.sum(dtype=dtype)


def elastic_force(x, hexagons, v2i, sfg, det, inv, mu, la, grad, dtype=np.float32):
    """quai_hexagon.py:109-133: grad[v2i[hex[e][i]]] -= P inv^T dN_i det  (v2i=None -> _noOrder, :135-158)."""
    x, sfg, det, inv = (a.astype(dtype) for a in (x, sfg, det, inv))
    F, E = _FE(x, hexagons, sfg, inv)
    tr = np.trace(E, axis1=-2, axis2=-1)
    S = dtype(2.0) * dtype(mu) * E + dtype(la) * tr[..., None, None] * np.eye(3, dtype=dtype)
    P = F @ S
    PinvT = P @ np.swapaxes(inv, -1, -2)               # [Nh,8q,3,3]
    tem = np.einsum("eqrc,aqc->eqar", PinvT, sfg) * det[:, :, None, None]   # [Nh,q,a,3]
    tem = tem.sum(axis=1)                              # atomics over q: order undefined in the reference
    idx = hexagons if v2i is None else v2i[hexagons]
    np.subtract.at(grad, idx.reshape(-1), tem.reshape(-1, 3).astype(grad.dtype))
    return grad


def gravity_force(m, g, grad, i2v):
    """quai_hexagon.py:176-186 (i2v=None -> _noOrder)."""
    mm = m if i2v is None else m[i2v]
    grad[:, 1] += (mm * grad.dtype.type(g)).astype(grad.dtype)
    return grad


def fixed_force(x, v2i, pin_list, grad, fixed_x, control_mag):
    """quai_hexagon.py:160-172: grad[v2i[p]] += k (fixed_x[id] - x[p])."""
    t = grad.dtype.type
    rows = pin_list if v2i is None else v2i[pin_list]
    grad[rows] += t(control_mag) * (fixed_x.astype(grad.dtype) - x[pin_list].astype(grad.dtype))
    return grad


def elastic_hessian_blocks(x, hexagons, sfg, det, inv, mu, la, dtype=np.float32, chunk=2048):
    """quai_hexagon.py:69-106, literal dF/dE/dP chain.  Yields (e0, blocks[ne,8a,8b,3,3]) with the
    sum over quadrature points done in q order (atomic order is undefined in the reference)."""
    x, sfg, det, inv = (a.astype(dtype) for a in (x, sfg, det, inv))
    mu = dtype(mu); la = dtype(la)
    eye = np.eye(3, dtype=dtype)
    for e0 in range(0, hexagons.shape[0], chunk):
        hx = hexagons[e0:e0 + chunk]
        iv = inv[e0:e0 + chunk]; dt = det[e0:e0 + chunk]
        F, E = _FE(x, hx, sfg, iv)                                   # [n,q,3,3]
        trE = np.trace(E, axis1=-2, axis2=-1)
        S = dtype(2.0) * mu * E + la * trE[..., None, None] * eye
        g = np.einsum("bqk,eqkj->eqbj", sfg, iv)                     # dF[i,:] = g_b  (:95)
        # dF(b,i) = e_i g_b^T ;  dE = 0.5 (dF^T F + F^T dF)
        dF = np.einsum("ic,eqbj->eqbicj", eye, g)                    # [n,q,b,i,3,3]
        Fx = F[:, :, None, None]
        dE = dtype(0.5) * (np.swapaxes(dF, -1, -2) @ Fx + np.swapaxes(Fx, -1, -2) @ dF)
        trdE = np.trace(dE, axis1=-2, axis2=-1)
        dP = dF @ S[:, :, None, None] + Fx @ (dtype(2.0) * mu * dE + la * trdE[..., None, None] * eye)
        dPinvT = dP @ np.swapaxes(iv, -1, -2)[:, :, None, None]      # [n,q,b,i,3,3]
        tem = np.einsum("eqbijc,aqc->eqabij", dPinvT, sfg) * dt[:, :, None, None, None, None]
        # temAnsForHessian[j,i] = temAns[j]  -> block[a,b][j,i]
        blocks = np.swapaxes(tem, -1, -2).sum(axis=1)                # [n,a,b,j,i]
        yield e0, blocks.astype(dtype)


def elastic_hessian(x, hexagons, sfg, det, inv, mu, la, values, offset, dtype=np.float32):
    """Adds the element blocks into ``values`` through the element->nnz map (quai_hexagon.py:106)."""
    for e0, blocks in elastic_hessian_blocks(x, hexagons, sfg, det, inv, mu, la, dtype):
        n = blocks.shape[0]
        np.add.at(values, offset[e0 * 64:(e0 + n) * 64], blocks.reshape(-1, 3, 3).astype(values.dtype))
    return values


def compute_fix_hessian(v2i_coarse, fix, control_mag, offset, hessian, fix_idx, fix_value):
    """cublas.py:234-242."""
    t = hessian.dtype.type
    for idx in np.nonzero(fix == 1)[0]:
        for j in range(8):
            if fix_idx[idx][j] != -1:
                jdx = fix_idx[idx][j]
                hessian[offset + v2i_coarse[jdx]] += t(fix_value[idx][j]) * t(fix_value[idx][j]) * t(control_mag) * np.eye(3, dtype=hessian.dtype)
    return hessian


# ----------------------------------------------------------------------------------------------
# cublas.py
# ----------------------------------------------------------------------------------------------
def solve3x3(A, b, x):
    """cublas.py:59-91, vectorised with per-lane 'returned' masks.  Falling off the end after
    3 iterations is taken as 'return x' (SURVEY 8c-iv).  Returns (x, n_not_spd)."""
    t = A.dtype.type
    x = x.copy()
    r = b.copy()
    r_norm = (r * r).sum(-1)
    done = r_norm < t(1e-10)
    p = r.copy()
    not_spd = 0
    for _ in range(3):
        Ap = np.einsum("nij,nj->ni", A, p)
        dot = (p * Ap).sum(-1)
        not_spd += int(((dot < 0) & ~done).sum())
        done = done | (dot < t(1e-10))
        with np.errstate(divide="ignore", invalid="ignore"):
            alpha = r_norm / dot
        act = ~done
        x = np.where(act[:, None], x + alpha[:, None] * p, x)
        r = np.where(act[:, None], r - alpha[:, None] * Ap, r)
        old = r_norm
        r_norm = np.where(act, (r * r).sum(-1), r_norm)
        done = done | (act & (r_norm < t(1e-10)))
        with np.errstate(divide="ignore", invalid="ignore"):
            beta = r_norm / old
        p = np.where((~done)[:, None], r + beta[:, None] * p, p)
    return x, not_spd


_GAMMA = 5.828427124
_CSTAR = 0.923879532
_SSTAR = 0.3826834323


def _approx_givens(a11, a12, a22):
    ch = 2 * (a11 - a22)
    sh = a12
    t = a11.dtype.type
    b = t(_GAMMA) * sh * sh < ch * ch
    with np.errstate(divide="ignore", invalid="ignore"):     # ch = sh = 0: w = inf, the select takes the constants
        w = t(1.0) / np.sqrt(ch * ch + sh * sh)
        ch = np.where(b, w * ch, t(_CSTAR))
        sh = np.where(b, w * sh, t(_SSTAR))
    return ch, sh


def _jacobi_conj(xi, yi, zi, s, qV):
    s11, s21, s22, s31, s32, s33 = s
    ch, sh = _approx_givens(s11, s21, s22)
    scale = ch * ch + sh * sh
    a = (ch * ch - sh * sh) / scale
    b = (2 * sh * ch) / scale
    _s11, _s21, _s22, _s31, _s32, _s33 = s11, s21, s22, s31, s32, s33
    s11 = a * (a * _s11 + b * _s21) + b * (a * _s21 + b * _s22)
    s21 = a * (-b * _s11 + a * _s21) + b * (-b * _s21 + a * _s22)
    s22 = -b * (-b * _s11 + a * _s21) + a * (-b * _s21 + a * _s22)
    s31 = a * _s31 + b * _s32
    s32 = -b * _s31 + a * _s32
    s33 = _s33
    tmp = [qV[0] * sh, qV[1] * sh, qV[2] * sh]
    sh = sh * qV[3]
    qV = [qV[0] * ch, qV[1] * ch, qV[2] * ch, qV[3] * ch]
    qV[zi] = qV[zi] + sh
    qV[3] = qV[3] - tmp[zi]
    qV[xi] = qV[xi] + tmp[yi]
    qV[yi] = qV[yi] - tmp[xi]
    # re-arrange for the next rotation
    return (s22, s32, s33, s21, s31, s11), qV


def eig3(A):
    """Restatement of wp.eig3 (see module docstring; parity unpinned).  Returns (V, d)."""
    t = A.dtype.type
    n = A.shape[0]
    s = (A[:, 0, 0].copy(), A[:, 1, 0].copy(), A[:, 1, 1].copy(), A[:, 2, 0].copy(), A[:, 2, 1].copy(), A[:, 2, 2].copy())
    qV = [np.zeros(n, A.dtype), np.zeros(n, A.dtype), np.zeros(n, A.dtype), np.ones(n, A.dtype)]
    for _ in range(4):
        s, qV = _jacobi_conj(0, 1, 2, s, qV)
        s, qV = _jacobi_conj(1, 2, 0, s, qV)
        s, qV = _jacobi_conj(2, 0, 1, s, qV)
    x, y, z, w = qV
    V = np.empty((n, 3, 3), A.dtype)
    V[:, 0, 0] = 1 - 2 * (y * y + z * z); V[:, 0, 1] = 2 * (x * y - w * z); V[:, 0, 2] = 2 * (x * z + w * y)
    V[:, 1, 0] = 2 * (x * y + w * z); V[:, 1, 1] = 1 - 2 * (x * x + z * z); V[:, 1, 2] = 2 * (y * z - w * x)
    V[:, 2, 0] = 2 * (x * z - w * y); V[:, 2, 1] = 2 * (y * z + w * x); V[:, 2, 2] = 1 - 2 * (x * x + y * y)
    u = np.swapaxes(V, -1, -2) @ A @ V
    d = np.stack([u[:, 0, 0], u[:, 1, 1], u[:, 2, 2]], axis=-1)
    return V, d


def spd_matrix33f(blocks, value):
    """cublas.py:115-125 applied to every block (diagonal AND off-diagonal)."""
    V, d = eig3(blocks)
    d = np.where(d < 0, blocks.dtype.type(value), d)
    return (V * d[:, None, :]) @ np.swapaxes(V, -1, -2)


def inf_norm(x):
    """cublas.py:38-48."""
    return np.abs(x).max() if x.size else x.dtype.type(0)
