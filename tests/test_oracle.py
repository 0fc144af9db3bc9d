"""CPU: the oracle against mathematics and against itself.

The reference ships no tests or golden vectors; the numpy oracle is therefore checked with the
identities the reference's own derivation implies (force = -dE/dx, Hessian = -d force/dx, H_ab = H_ba^T,
Galerkin product vs dense triple product, Gauss-Seidel sweep vs a dense triangular solve), its float32 path
against its float64 twin, and the independent C restatement (oracle/femb_oracle.c) against the numpy one.
"""
import numpy as np
import pytest

from fem_simulation_b200 import hierarchy, mesh
from oracle import ref_kernels as K
from oracle import ref_solvers


@pytest.fixture(scope="module")
def small():
    pts, hx = mesh.jagged_mesh(10, 9, 11, 0.05, seed=1)
    ref = ref_solvers.RefObject(pts, hx, 0.05, [], dtype=np.float64)
    rng = np.random.default_rng(0)
    x = ref.x_gpu_layer[0] + rng.normal(size=ref.x_gpu_layer[0].shape) * 0.003
    return ref, x


def test_force_is_minus_energy_gradient(small):
    ref, x = small
    g = np.zeros_like(x); ref.force(x, g)                       # index order
    rng = np.random.default_rng(1)
    for _ in range(3):
        d = rng.normal(size=x.shape)
        eps = 1e-6
        fd = (ref.total_energy(x + eps * d) - ref.total_energy(x - eps * d)) / (2 * eps)
        an = -(g[ref.v2i[0]] * d).sum()
        assert abs(fd - an) <= 1e-6 * max(1.0, abs(an))


def test_hessian_is_minus_force_jacobian_and_symmetric(small):
    ref, x = small
    p = ref.pat[0]
    vals = np.zeros_like(ref.MF_value); ref.hessian(0, x, vals)
    rng = np.random.default_rng(2)
    d = rng.normal(size=x.shape)                                # vertex order
    d_idx = np.zeros_like(d); d_idx[ref.v2i[0]] = d
    Hd = np.zeros_like(d)
    np.add.at(Hd, p["row"], np.einsum("nij,nj->ni", vals, d_idx[p["col"]]))
    eps = 1e-6
    f1 = np.zeros_like(x); K.elastic_force(x + eps * d, ref.hexs[0], ref.v2i[0], ref.sfg, ref.det[0], ref.inv[0], ref.mu, ref.la, f1, np.float64)
    f0 = np.zeros_like(x); K.elastic_force(x - eps * d, ref.hexs[0], ref.v2i[0], ref.sfg, ref.det[0], ref.inv[0], ref.mu, ref.la, f0, np.float64)
    fd = -(f1 - f0) / (2 * eps)
    assert np.abs(fd - Hd).max() <= 1e-6 * np.abs(Hd).max()
    # symmetry through the pattern: block (r,c) == block (c,r)^T
    n = ref.dims[0]
    key = p["row"].astype(np.int64) * n + p["col"]
    tkey = p["col"].astype(np.int64) * n + p["row"]
    order = np.argsort(key); pos = order[np.searchsorted(key[order], tkey)]
    assert np.abs(vals - np.swapaxes(vals[pos], 1, 2)).max() <= 1e-12 * np.abs(vals).max()


def test_closed_form_block_matches_literal_chain(small):
    """The closed form the CUDA kernel uses (femb_element.cu) vs the literal dF/dE/dP chain of quai_hexagon.py:86-106."""
    ref, x = small
    hx = ref.hexs[0][:64]
    mu, la = 250.0, 30.0
    for e0, blocks in K.elastic_hessian_blocks(x, hx, ref.sfg, ref.det[0][:64], ref.inv[0][:64], mu, la, np.float64):
        F, E = K._FE(x, hx, ref.sfg, ref.inv[0][:64])
        trE = np.trace(E, axis1=-2, axis2=-1)
        S = 2 * mu * E + la * trE[..., None, None] * np.eye(3)
        g = np.einsum("bqk,eqkj->eqbj", ref.sfg, ref.inv[0][:64])
        u = np.einsum("eqij,eqbj->eqbi", F, g)
        t = np.einsum("eqij,eqbj->eqbi", S, g)
        FFT = F @ np.swapaxes(F, -1, -2)
        det = ref.det[0][:64]
        s1 = np.einsum("eqaj,eqbj->eqab", g, t)
        s2 = np.einsum("eqaj,eqbj->eqab", g, g)
        blk = (s1[..., None, None] * np.eye(3) + mu * s2[..., None, None] * FFT[:, :, None, None]
               + mu * np.einsum("eqbj,eqai->eqabji", u, u) + la * np.einsum("eqaj,eqbi->eqabji", u, u))
        blk = (blk * det[:, :, None, None, None, None]).sum(axis=1)
        assert np.abs(blk - blocks).max() <= 1e-12 * np.abs(blocks).max()


def test_gauss_seidel_equals_triangular_solves(small):
    """One symmetric sweep == (D+U) dx1 = b ; (D+L) dx2 = b - U dx1 with the block diagonal inverted exactly."""
    ref, x = small
    ref.MF_value[...] = ref.MF_value_fixed; ref.hessian(0, x, ref.MF_value)
    ref.set_LUD(0)
    p = ref.pat[0]; n = ref.dims[0]
    rng = np.random.default_rng(3)
    b = rng.normal(size=(n, 3)) * 1e3          # large rhs: the 1e-10 cut-offs of solve3x3 (cublas.py:67-86) are negligible
    ref.dev_B_fixed[0][...] = b
    ref.PerformGaussSeidel(0, 1)
    A = np.zeros((3 * n, 3 * n))
    for s in range(p["nnz"]):
        r, c = p["row"][s], p["col"][s]
        A[3 * r:3 * r + 3, 3 * c:3 * c + 3] = ref.MF_value[s]
    blk = np.kron(np.arange(n)[:, None] >= np.arange(n)[None, :], np.ones((3, 3), bool))      # block lower incl. diagonal
    DL, DU = np.where(blk, A, 0.0), np.where(blk.T, A, 0.0)
    U, L = A - DL, A - DU
    dx1 = np.linalg.solve(DU, b.reshape(-1))
    dx2 = np.linalg.solve(DL, b.reshape(-1) - U @ dx1)
    assert np.abs(ref.dev_delta_x[0].reshape(-1) - dx2).max() <= 1e-6 * np.abs(dx2).max()
    assert L.shape == A.shape


def test_galerkin_equals_dense_triple_product(small):
    ref, x = small
    ref.MF_value[...] = ref.MF_value_fixed; ref.hessian(0, x, ref.MF_value)
    spd = K.spd_matrix33f
    K.spd_matrix33f = lambda blocks, value: blocks            # isolate the triple product from the eigen-clamp
    try:
        ref.galerkin()
    finally:
        K.spd_matrix33f = spd
    p, pc = ref.pat[0], ref.pat[1]
    n, nc = ref.dims[0], ref.dims[1]
    A = np.zeros((3 * n, 3 * n))
    for s in range(p["nnz"]):
        A[3 * p["row"][s]:3 * p["row"][s] + 3, 3 * p["col"][s]:3 * p["col"][s] + 3] = ref.MF_value[s]
    P = np.zeros((3 * n, 3 * nc))
    us = ref.Us_hat[0]
    rows = np.repeat(np.arange(us["nrow"]), np.diff(us["offsets"]))
    for r, c, w in zip(rows, us["columns"], us["values"]):
        P[3 * r:3 * r + 3, 3 * c:3 * c + 3] = w * np.eye(3)
    Ac = P.T @ A @ P
    for s in range(pc["nnz"]):
        r, c = pc["row"][s], pc["col"][s]
        assert np.abs(Ac[3 * r:3 * r + 3, 3 * c:3 * c + 3] - ref.UtAUs_values[0][s]).max() <= 1e-10 * np.abs(Ac).max()


def test_spd_matrix33f_on_symmetric_blocks():
    """On symmetric blocks the restated eig3 must reproduce the block when it is SPD and clamp negative eigenvalues."""
    rng = np.random.default_rng(4)
    M = rng.normal(size=(200, 3, 3))
    S = M @ np.swapaxes(M, 1, 2) + 0.1 * np.eye(3)
    out = K.spd_matrix33f(S, 1e-3)
    assert np.abs(out - S).max() <= 5e-3 * np.abs(S).max()     # 4 approximate-Givens sweeps: not exact (McAdams 2011)
    N = -S
    out = K.spd_matrix33f(N, 1e-3)
    w = np.linalg.eigvalsh(out)
    assert w.min() >= -1e-9 and np.abs(w - 1e-3).max() <= 1e-6


def test_fp32_oracle_tracks_fp64_twin():
    pts, hx = mesh.cube_mesh(9, 0.05)
    r32 = ref_solvers.RefObject(pts, hx, 0.05, [])
    r64 = ref_solvers.RefObject(pts, hx, 0.05, [], dtype=np.float64, setup=r32.S)
    r32.NewtonMultigrid(2); r64.NewtonMultigrid(2)
    a, b = r32.x_gpu_layer[0], r64.x_gpu_layer[0]
    assert np.abs(a - b).max() <= 1e-6 * np.abs(b).max()
    assert r32.not_spd == 0


def test_c_oracle_matches_numpy_oracle():
    from oracle import c_oracle
    pts, hx = mesh.cube_mesh(9, 0.05)
    H = hierarchy.build_hierarchy(pts, hx, 0.05)
    plans = [hierarchy.build_galerkin_plan(H, l) for l in range(H.layer - 1)]
    for method, step in (("NewtonMultigrid", "newton_mg_step"), ("FASv3", "fasv3_step")):
        ref = ref_solvers.RefObject(pts, hx, 0.05, [])
        co = c_oracle.from_ref(ref, H, plans, native=False)
        for _ in range(2):
            getattr(co, step)()
        getattr(ref, method)(2)
        assert np.abs(co.x - ref.x_gpu_layer[0]).max() <= 1e-6 * np.abs(ref.x_gpu_layer[0]).max()
        disp = ref.x_gpu_layer[0] - ref.S["x0"]
        assert np.abs((co.x - ref.S["x0"]) - disp).max() <= 1e-3 * np.abs(disp).max()


def test_c_oracle_assembly_gather_matches_literal_chain():
    """oracle/femb_oracle.c assembles without atomics (element scratch + per-row gather through hex_update_offset, closed-form
    element blocks): same values as the literal dF/dE/dP chain with the kernel's scatter (orc_hessian_literal) and as the
    numpy oracle, on a ragged two-level mesh, level 0 (with the pin diagonal) and level 1; forces likewise."""
    from oracle import c_oracle
    pts, hx = mesh.jagged_mesh(9, 10, 9, 0.05, seed=5)
    H = hierarchy.build_hierarchy(pts, hx, 0.05)
    plans = [hierarchy.build_galerkin_plan(H, l) for l in range(H.layer - 1)]
    ref = ref_solvers.RefObject(pts, hx, 0.05, [])
    co = c_oracle.from_ref(ref, H, plans, native=False)
    rng = np.random.default_rng(3)
    x = (ref.x_gpu_layer[0] + rng.normal(size=pts.shape).astype(np.float32) * 0.003).astype(np.float32)
    diag = np.zeros(H.dims[0], np.float32); diag[H.vertex2index[0][ref.pin_list]] = ref.control_mag
    a = np.zeros_like(ref.MF_value); b = np.zeros_like(a); a.fill(5.0)
    co.hessian(0, x, a, diag); co.hessian_literal(0, x, b, diag)
    v = ref.MF_value_fixed.copy(); ref.hessian(0, x, v)
    scale = np.abs(v).max()
    assert np.abs(a - b).max() <= 1e-5 * scale and np.abs(a - v).max() <= 1e-5 * scale
    x1 = (ref.x_gpu_layer[1] * np.float32(1.01)).astype(np.float32)
    a1 = np.zeros_like(ref.UtAUs_values[0]); b1 = np.zeros_like(a1)
    co.hessian(1, x1, a1); co.hessian_literal(1, x1, b1)
    assert np.abs(a1 - b1).max() <= 1e-5 * np.abs(b1).max()
    f = np.zeros_like(x); co.force(x, f)
    g = np.zeros_like(x); ref.force(x, g)
    assert np.abs(f - g).max() <= 2e-5 * np.abs(g).max()


def test_galerkin_from_element_matrices_identity():
    """Ground work for the next optimisation: on level 0 (no eigen-clamp between assembly and product)
    the Galerkin operator Ut_hat A Us_hat equals the sum over fine hexahedra e of P_e^T K_e P_e (+ P^T D_pin P), where P_e
    only involves the 8 corners of the coarse hexahedron that contains e (`Hierarchy.fine_hex_parent`) with plain trilinear
    weights 1, 1/2, 1/4, 1/8.  Checked against the oracle's bsr_mm product on a ragged mesh."""
    import scipy.sparse as sp
    from fem_simulation_b200 import hierarchy, mesh
    from oracle import ref_kernels as K
    pts, hx = mesh.jagged_mesh(8, 9, 8, 0.05, seed=11, keep=0.7)
    H = hierarchy.build_hierarchy(pts, hx, 0.05)
    assert H.layer == 2 and len(H.fine_hex_parent) == 1
    par = H.fine_hex_parent[0]
    chex = H.hexs[1]
    ut = H.Ut_noOrder_hat[0]                                      # coarse x fine, vertex order, un-normalised weights
    P = sp.csr_matrix((ut.weights.astype(np.float64), ut.columns, ut.offsets), shape=(ut.nrow, ut.ncol)).T.tocsr()   # fine x coarse
    # (1) every parent of every node of a fine hexahedron is a corner of its coarse hexahedron, with trilinear weights
    fc = H.fine_coarse[0]
    for e in range(0, hx.shape[0], 7):
        corners = set(int(c) for c in chex[par[e]])
        for a in range(8):
            v = int(hx[e, a])
            row = P.getrow(v)
            assert set(int(c) for c in row.indices) <= corners
            slot = int(np.nonzero(fc[par[e]] == v)[0][0])
            r = np.array([slot // 9, (slot // 3) % 3, slot % 3])
            for c, w in zip(row.indices, row.data):
                b = int(np.nonzero(chex[par[e]] == c)[0][0])
                bits = np.array([(b >> 2) & 1, (b >> 1) & 1, b & 1])
                assert w == np.prod(1.0 - np.abs(r / 2.0 - bits))
    # (2) the identity itself, in float64
    rng = np.random.default_rng(0)
    x = (pts + rng.normal(size=pts.shape) * 0.002).astype(np.float64)
    det, inv = K.prepare_kernal(pts.astype(np.float64), hx, H.shapeFuncGrad, dtype=np.float64)
    p0, p1 = H.pattern[0], H.pattern[1]
    vals = np.zeros((p0.nnz, 3, 3))
    K.elastic_hessian(x, hx, H.shapeFuncGrad, det, inv, 250.0, 0.0, vals, p0.hex_update_offset, dtype=np.float64)
    d_pin = np.zeros(H.dims[0]); d_pin[hierarchy.default_pins(pts, 0.05)] = 100.0
    vals[p0.off_d + H.vertex2index[0]] += d_pin[:, None, None] * np.eye(3)
    n0, n1 = H.dims[0], H.dims[1]
    i2v0, i2v1 = H.index2vertex[0], H.index2vertex[1]
    rows_v, cols_v = i2v0[p0.row], i2v0[p0.col]                   # the blocks in VERTEX order
    want = np.zeros((n1, n1, 3, 3))
    for a in range(3):
        for b in range(3):
            Aab = sp.csr_matrix((vals[:, a, b], (rows_v, cols_v)), shape=(n0, n0))
            want[:, :, a, b] = (P.T @ Aab @ P).toarray()           # Ut_hat A Us_hat
    Pd = P.toarray()
    got = np.zeros((n1, n1, 3, 3))
    for e0, blocks in K.elastic_hessian_blocks(x, hx, H.shapeFuncGrad, det, inv, 250.0, 0.0, dtype=np.float64):
        for k in range(blocks.shape[0]):
            e = e0 + k
            cs = chex[par[e]]
            Pe = Pd[np.ix_(hx[e], cs)]                             # 8 fine nodes x 8 coarse corners
            got[np.ix_(cs, cs)] += np.einsum("aA,abij,bB->ABij", Pe, blocks[k], Pe)
    got += np.einsum("iI,i,iJ->IJ", Pd, d_pin, Pd)[:, :, None, None] * np.eye(3)
    assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()
    # and it is the matrix the oracle's Galerkin produces on the coarse pattern
    assert np.count_nonzero(np.abs(want).sum(axis=(2, 3))) <= p1.nnz


def test_fas_multilevel_oracle_reduces_to_fasv0_and_converges():
    """RefObject.FAS_multilevel (the generalisation of FASv0 to more than two levels): levels=2 is FASv0 bit for
    bit; on a corrected 3-level hierarchy the elastic force keeps falling from cycle to cycle."""
    from fem_simulation_b200 import mesh
    from oracle import ref_solvers
    from oracle import ref_kernels as K
    pts, hx = mesh.cube_mesh(9, 0.05)
    a = ref_solvers.RefObject(pts, hx, 0.05, [])
    b = ref_solvers.RefObject(pts, hx, 0.05, [])
    a.FASv0(3); b.FAS_multilevel(3, levels=2)
    assert np.array_equal(a.dev_x_solved[0], b.dev_x_solved[0])
    pts, hx = mesh.jagged_mesh(17, 16, 18, 0.05, seed=3, keep=0.3)
    r = ref_solvers.RefObject(pts, hx, 0.05, [], corrected_spacing=True)
    assert r.layer >= 3
    rng = np.random.default_rng(0)
    r.x_gpu_layer[0][...] += (rng.normal(size=r.x_gpu_layer[0].shape) * 0.002).astype(np.float32)   # a strained start state
    norms = []
    for _ in range(4):
        if norms:
            r.x_gpu_layer[0][...] = r.dev_x_solved[0]
        r.FAS_multilevel(1, levels=3, smoothing=1)
        f = np.zeros_like(r.dev_x_solved[0])
        r.force(r.dev_x_solved[0], f, gravity=False, pins=False)
        norms.append(K.inf_norm(f))
    assert np.all(np.isfinite(norms)) and norms[-1] < norms[0], norms


def test_exp2_train_oracle_forward_and_gradient():
    """oracle/ref_exp2_train.py (torch restatement of the taped step of exp2 train(), object.py:1202-1287 of exp2/p_hat): its
    forward map reproduces one step of the numpy oracle's compare loop (pinned by ref_exp2.npz), and its autograd gradient with
    respect to the interpolation weights equals central differences of that map in float64."""
    torch = pytest.importorskip("torch")
    from fem_simulation_b200 import mesh
    from oracle.ref_exp2 import RefExp2Object
    from oracle.ref_exp2_train import Exp2Train
    pts, hx = mesh.cube_mesh(9, 0.05)
    r = RefExp2Object(pts, hx, 0.05, [], dtype=np.float64)
    x0 = r.x_gpu_layer[0].copy()
    x0[500] += np.array([7e-4, -4e-4, 9e-4])
    rng = np.random.default_rng(3)
    w = r.w.astype(np.float64) * (1.0 + 0.05 * rng.standard_normal(r.w.shape))      # off the default: every term of the chain is live
    tr = Exp2Train(r, dtype=torch.float64)
    loss, g, w_new, aux = tr.step(w, x0)
    # forward: one step of the pinned compare loop with the same weights
    r.x_gpu_layer[0][...] = x0
    r.compare_loop(1, w)
    assert np.abs(aux["xn"] - r.x_gpu_layer[0]).max() <= 1e-12 * np.abs(x0).max()
    p = r.pat[1]
    D_ref = r.UtAUs_values[0][p["off_d"]:p["off_d"] + r.dims[1]]
    assert np.abs(aux["D"] - D_ref).max() <= 1e-11 * np.abs(D_ref).max()
    assert np.isfinite(loss) and np.abs(g).max() > 0
    # gradient: central differences along random directions and on the largest entries
    f = lambda wv: float(tr.forward(torch.tensor(wv, dtype=torch.float64), torch.tensor(x0, dtype=torch.float64))[0])
    for k in range(3):
        d = rng.standard_normal(w.shape)
        d /= np.linalg.norm(d)
        h = 1e-6
        fd = (f(w + h * d) - f(w - h * d)) / (2 * h)
        assert abs(fd - g @ d) <= 1e-5 * max(abs(fd), np.abs(g).max()), (k, fd, g @ d)
    for i in np.argsort(-np.abs(g))[:3]:
        e = np.zeros_like(w); e[i] = 1.0
        h = 1e-6
        fd = (f(w + h * e) - f(w - h * e)) / (2 * h)
        assert abs(fd - g[i]) <= 1e-5 * np.abs(g).max(), (i, fd, g[i])
    assert np.all(w_new >= 0) and np.all(w_new <= 1)
