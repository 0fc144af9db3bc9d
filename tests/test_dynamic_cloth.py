"""Dynamic (implicit-Euler) hexahedral step and mass-spring cloth step: CUDA path vs the CPU oracle
(oracle/ref_dynamic.py <- exp1/simulation, oracle/ref_cloth.py <- exp1/cloth_simulation)."""
import numpy as np
import pytest

from fem_simulation_b200 import cloth as cloth_mod
from fem_simulation_b200 import mesh
from oracle.ref_cloth import RefCloth
from oracle.ref_dynamic import RefDynamicObject


def rel(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / max(np.abs(b).max(), 1e-30))


def test_cloth_maps_bit_exact():
    """Vectorised grid / pattern / offset maps == the literal loops of exp1/cloth_simulation/object.py:31-176."""
    r = RefCloth()
    x, face, edge = cloth_mod.cloth_grid()
    p = cloth_mod.cloth_pattern(len(x), edge)
    assert np.array_equal(x, r.x_cpu) and np.array_equal(edge, r.edge)
    assert p["nnz"] == r.MF_nnz == 29057 and len(edge) == 12416 and len(x) == 4225
    for k, v in (("MF_row", r.MF_row), ("MF_col", r.MF_col), ("offset", r.offset), ("d_offset", r.d_offset)):
        assert np.array_equal(p[k], v), k
    # the [L;D;U] slots are a permutation and keep rows contiguous
    assert np.array_equal(np.sort(p["ldu_slot"]), np.arange(p["nnz"]))


@pytest.mark.gpu
@pytest.mark.parametrize("graphs", [False, True])
def test_cloth_frames(graphs):
    from fem_simulation_b200 import Cloth
    c = Cloth(use_graphs=graphs)
    r = RefCloth()
    assert rel(c.rest_length.cpu().numpy(), r.rest_length) <= 1e-6
    for _ in range(6):
        c.step(); r.render()
    assert rel(c.MF_value_gpu.cpu().numpy(), r.MF_value) <= 1e-5
    x, xr = c.dev_x.cpu().numpy(), r.dev_x
    assert rel(x, xr) <= 1e-5
    assert np.abs((x - r.x_cpu) - (xr - r.x_cpu)).max() <= 1e-3 * np.abs(xr - r.x_cpu).max()
    assert rel(c.dev_v.cpu().numpy(), r.dev_v) <= 1e-3
    assert np.array_equal(x[[0, 65]], r.x_cpu[[0, 65]])              # frozen vertices (cublas.py:22-26 there)


@pytest.mark.gpu
@pytest.mark.parametrize("pins", [[], [100]])
def test_dynamic_frames(pins):
    import torch
    from fem_simulation_b200 import DynamicObject
    pts, hx = mesh.jagged_mesh(10, 9, 11, 0.05, seed=1)
    d = DynamicObject((pts, hx), 0.05, list(pins))
    r = RefDynamicObject(pts, hx, 0.05, list(pins))
    assert d.control_mag == 10.0 and d.g == -9.8 and d.N_pin == r.N_pin
    for _ in range(4):
        d.step(); r.render()
    x, xr = d.dev_x.cpu().numpy(), r.dev_x
    assert rel(d.MF_value_gpu.cpu().numpy(), r.MF_value) <= 1e-5
    assert rel(x, xr) <= 1e-5
    assert np.abs((x - r.S["x0"]) - (xr - r.S["x0"])).max() <= 2e-3 * np.abs(xr - r.S["x0"]).max()
    assert rel(d.dev_v.cpu().numpy(), r.dev_v) <= 2e-3
    # drag constraint: Control_Kernel + Fixed_Update_Kernel + spring force
    v = int(np.argmin(pts[:, 1]))
    d.drag(v, pts[v] + np.array([0.05, -0.02, 0.0], np.float32))
    r.more_fixed[...] = (((r.dev_x - r.dev_x[v]) ** 2).sum(-1) < 0.002).astype(np.int32)
    assert np.array_equal(d.more_fixed_gpu.cpu().numpy(), r.more_fixed)
    tgt = pts[v] + np.array([0.05, -0.02, 0.0], np.float32)
    for _ in range(2):
        dirv = tgt - r.dev_x[v]
        dirv = dirv * min(float(np.linalg.norm(dirv)), 0.5)
        d.step(); r.render(dirv)
    assert rel(d.dev_x.cpu().numpy(), r.dev_x) <= 1e-5
    # the quasi-static solvers of the dynamic class apply spd_matrix33f on level 0 and carry m/dt^2 in MF_value_fixed
    d2 = DynamicObject((pts, hx), 0.05, list(pins))
    r2 = RefDynamicObject(pts, hx, 0.05, list(pins))
    d2.Newton(2)
    import oracle.ref_kernels as K
    for _ in range(2):
        r2.grad[...] = 0; r2.MF_value[...] = r2.MF_value_fixed
        r2.force(r2.x_gpu_layer[0], r2.grad); r2.hessian(0, r2.x_gpu_layer[0], r2.MF_value)
        r2.MF_value[...] = K.spd_matrix33f(r2.MF_value, r2.spd_value)
        r2.set_LUD(0); r2.dev_B_fixed[0][...] = r2.grad
        r2.PerformConjugateGradient(0, 5); r2.update_deltaX(r2.x_gpu_layer[0], r2.dev_delta_x[0])
    assert rel(d2.x_gpu_layer[0].cpu().numpy(), r2.x_gpu_layer[0]) <= 1e-5
    torch.cuda.synchronize()


@pytest.mark.gpu
def test_exp2_compare_forward():
    """exp2 compare(): classic weights and a perturbed ("trained") weight vector, CUDA vs oracle."""
    import torch
    from fem_simulation_b200.object_exp2 import Exp2Object
    from oracle.ref_exp2 import RefExp2Object
    pts, hx = mesh.jagged_mesh(10, 9, 11, 0.05, seed=1)
    o = Exp2Object((pts, hx), 0.05, [])
    r = RefExp2Object(pts, hx, 0.05, [])
    pv, pert = 300, (0.0007, -0.0004, 0.0009)
    rng = np.random.default_rng(11)
    w_tr = (r.w * (1.0 + 0.05 * rng.standard_normal(len(r.w)))).astype(np.float32)
    o.compare(4, perturb_vertex=pv, perturbation=pert, trained=torch.from_numpy(w_tr))
    r.x_gpu_layer[0][pv] += np.array(pert, np.float32)
    x_start = r.x_gpu_layer[0].copy()
    r.compare_loop(4, r.w)
    classic = list(r.plot_InfNorm)
    x_classic = r.x_gpu_layer[0].copy()
    r.x_gpu_layer[0][...] = x_start
    r.plot_InfNorm = []
    r.compare_loop(4, w_tr)
    assert rel(np.array(o.plot_InfNorm), np.array(classic)) <= 1e-4
    assert rel(np.array(o.plot_InfNorm_newton), np.array(r.plot_InfNorm)) <= 1e-4
    assert rel(o.x_gpu_layer[0].cpu().numpy(), r.x_gpu_layer[0]) <= 1e-5
    assert np.abs(x_classic - r.x_gpu_layer[0]).max() > 0          # the trained weights do change the iteration
    assert o.Ut_noOrder_value.numel() == o.Ut_noOrder[0].nnz


@pytest.mark.gpu
def test_exp2_train_step_matches_oracle():
    """exp2 train(): loss and gradient with respect to the interpolation weights from the hand-written adjoint
    kernels (csrc/femb_exp2.cu) vs torch.autograd through the oracle's restatement of the taped step (oracle/ref_exp2_train.py,
    float64; parity of the gradient is unpinned -- no Warp tape here -- but that restatement's forward map is the pinned compare
    loop and its gradient equals central differences, tests/test_oracle.py).  Tolerance = the float32-vs-float64 gap of the oracle
    itself on the same inputs."""
    import torch
    from fem_simulation_b200.object_exp2 import Exp2Object
    from oracle.ref_exp2 import RefExp2Object
    from oracle.ref_exp2_train import Exp2Train
    pts, hx = mesh.cube_mesh(9, 0.05)
    obj = Exp2Object((pts, hx), 0.05, [], use_graphs=False)
    r = RefExp2Object(pts, hx, 0.05, [], dtype=np.float64)
    rng = np.random.default_rng(3)
    w = (r.w.astype(np.float64) * (1.0 + 0.05 * rng.standard_normal(r.w.shape))).astype(np.float32)
    x0 = r.x_gpu_layer[0].astype(np.float32).copy()
    x0[500] += np.array([7e-4, -4e-4, 9e-4], np.float32)
    l64, g64, w64, _ = Exp2Train(r, dtype=torch.float64).step(w.astype(np.float64), x0.astype(np.float64))
    l32, g32, _, _ = Exp2Train(r, dtype=torch.float32).step(w, x0)
    obj.load_weights(torch.from_numpy(w))
    loss, grad = obj.train_step(torch.from_numpy(x0).to(obj.device), update=True)
    g = grad.cpu().numpy().astype(np.float64)
    scale = np.abs(g64).max()
    noise = np.abs(g32.astype(np.float64) - g64).max() / scale
    err = np.abs(g - g64).max() / scale
    assert abs(loss - l64) <= max(1e-5, 4 * abs(l32 - l64) / abs(l64)) * abs(l64), (loss, l64, l32)
    assert err <= max(1e-4, 4 * noise), (err, noise)
    assert noise <= 1e-2
    # Valueaxpy: w + 1e-3 grad clamped to [0, 1]
    w_new = obj.Ut_noOrder_value.cpu().numpy()
    assert np.abs(w_new - np.clip(w + np.float32(1e-3) * grad.cpu().numpy(), 0.0, 1.0)).max() <= 1e-7
    # and the driver: a few iterations with given perturbations keep the weights in range and return finite losses
    losses = obj.train(3, perturb_vertex=500, perturbations=[[7e-4, -4e-4, 9e-4]] * 3)
    assert len(losses) == 3 and np.all(np.isfinite(losses))
    wv = obj.Ut_noOrder_value
    assert float(wv.min()) >= 0.0 and float(wv.max()) <= 1.0
