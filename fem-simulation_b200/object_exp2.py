"""``Exp2Object`` -- forward path of exp2's ``compare`` (classic vs trained position-restriction P~).

Mirrors /root/reference/exp2/p_hat/quasi_simulation/object.py:1289-1357 (the canonical copy): per step
  R0 = elastic force(x) ; log |R0|inf ; b1 = Ut_hat R0 ;
  Ut_noOrder <- Kronecker(w) ; coarse REST shape re-derived from Ut_noOrder x_init (prepare_kernal) ;
  x1 = Ut_noOrder X  (X = the state copied at construction, :825-826 -- not the current x) ;
  coarse Hessian at x1 on zeroed values (no pins) ; ONE Jacobi straight on the diagonal ; dx0 = Us_hat dx1 ; x += dx0.
The weights w (``Ut_noOrder_value``, float32[nnz(Ut_noOrder[0])]) start as the default row-normalised weights
(p_hat_init, cublas.py:254-256 there) and can be replaced by a trained ``.pth`` (``load_weights``, the format of
``assets/p/Ut_noOrder_0.pth``).

``train()`` (object.py:1202-1287 there: Warp autodiff, ``wp.Tape``, through this step) is here the same forward pass plus
hand-written vector-Jacobian kernels (csrc/femb_exp2.cu): one-hot cotangent of the ``atomic_max`` loss, minus the fine Hessian
for the force, ``Ut_hat`` for ``Us_hat^T``, forward-mode dual numbers for ``solve3x3`` and for the coarse diagonal blocks as a
function of the coarse positions AND the coarse rest shape, then ``w += 1e-3 grad`` clamped to [0, 1] (``Valueaxpy``).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .object import Object


class Exp2Object(Object):
    def __init__(self, mesh_path, dx, pinList, **kw):
        kw.setdefault("pin_displacement", 0.0)               # exp2 object.py:795 displaces by 0.0
        super().__init__(mesh_path, dx, pinList, **kw)
        if self.layer < 2:
            raise ValueError("exp2 needs at least two levels")
        self.dev_x_init = self.x_gpu_layer[0].clone()        # :742-743 (rest state)
        self.X = self.x_gpu_layer[0].clone()                 # :825-826
        self._w_default = self.Ut_noOrder[0].weights.clone()
        self.Ut_noOrder_value = self._w_default.clone()      # :445-446 p_hat_init
        self.plot_InfNorm_newton = []

    def load_weights(self, path_or_tensor):
        w = torch.load(path_or_tensor) if isinstance(path_or_tensor, str) else path_or_tensor
        w = torch.as_tensor(w, dtype=torch.float32).to(self.device).reshape(-1)
        if w.numel() != self.Ut_noOrder[0].nnz:
            raise ValueError(f"weights have {w.numel()} entries, Ut_noOrder[0] has {self.Ut_noOrder[0].nnz}")
        self.Ut_noOrder_value.copy_(w)                      # in place: the captured CUDA graph reads this tensor

    def save_p_hat(self, path):
        torch.save(self.Ut_noOrder_value.detach().cpu(), path)

    def _exp2_body(self):
        x = self.x_gpu_layer[0]
        self._force(x, self.dev_R[0], gravity=False, pins=False)                                # :1309-1310
        self._transfer(self.Ut_hat[0], 1.0, self.dev_R[0], 0.0, self.dev_B_fixed[1])           # :1314
        self.Ut_noOrder[0].set_weights(self.Ut_noOrder_value)                                   # Kronecker_product, :1316
        self._transfer(self.Ut_noOrder[0], 1.0, self.dev_x_init, 0.0, self.dev_x_solved[1])    # :1317
        lv = self.levels[1]
        _lib.check(self.lib.femb_prepare_rest(lv.n_hex, lv.hex, self.dev_x_solved[1].data_ptr(), lv.detJ, lv.invJ, self._stream()),
                   "femb_prepare_rest")                                                         # :1318
        self._transfer(self.Ut_noOrder[0], 1.0, self.X, 0.0, self.x_gpu_layer[1])              # :1319
        self._hessian(1, self.x_gpu_layer[1], self.UtAUs_values[0])                             # :1320-1321
        self.PerformJacobi(layer=1, iterations=1)                                               # :1322
        self._transfer(self.Us_hat[0], 1.0, self.dev_delta_x[1], 0.0, self.dev_delta_x[0])     # :1323
        self.update_deltaX(x, self.dev_delta_x[0])                                              # :1324

    def compare(self, iterations=100, perturb_vertex=732, perturbation=(0.0007, -0.0004, 0.0009), trained=None):
        """Both loops of compare(); the reference's unseeded random perturbation of vertex 732 (:1291-1296) is an argument.
        Fills plot_x / plot_InfNorm (classic) and plot_InfNorm_newton (trained weights, if given)."""
        if perturb_vertex is not None and perturb_vertex < self.N_verts:
            self.x_cpu[perturb_vertex] += torch.tensor(perturbation, dtype=torch.float32)
        runs = [(self._w_default, self.plot_InfNorm, True)]
        if trained is not None:
            self.load_weights(trained)
            runs.append((self.Ut_noOrder_value.clone(), self.plot_InfNorm_newton, False))
        for w, log, first in runs:
            self.x_gpu_layer[0].copy_(self.x_cpu)
            self.Ut_noOrder_value.copy_(w)
            for step in range(1, iterations + 1):
                self._run_step("exp2", self._exp2_body)
                if first:
                    self.plot_x.append(step)
                log.append(self.inf_norm(self.dev_R[0]))         # force evaluated at the start of the step (:1311-1313)

    # ---- train(): object.py:1202-1287 of exp2/p_hat --------------------------------------------------------------------------
    def _train_buffers(self):
        if getattr(self, "_tb", None) is None:
            dev, n0, n1 = self.device, self.dims[0], self.dims[1]
            z = lambda *shape: torch.zeros(shape, device=dev)
            self._tb = dict(res=z(n0, 3), onehot=z(n0, 3), lam0=z(n0, 3), lam1=z(n1, 3), dbar_i=z(n1, 9), dbar_v=z(n1, 9),
                            G=z(max(self.hexs[1].shape[0], 1) * 48), lam_x1=z(n1, 3), lam_r1=z(n1, 3),
                            grad=torch.zeros_like(self.Ut_noOrder_value), reg=z(n1), maxv=z(1),
                            A0=self.new_values(0))
        return self._tb

    def train_step(self, x_start, alpha=0.1, lr=1e-3, update=True):
        """One iteration of train()'s loop from the fine state `x_start` ([N,3], vertex order): returns (loss, grad) and, if
        `update`, applies Valueaxpy to ``Ut_noOrder_value``.  Forward pass = the taped kernels of the reference in their order."""
        tb = self._train_buffers()
        st = self._stream
        x = self.x_gpu_layer[0]
        x.copy_(x_start)
        w = self.Ut_noOrder_value
        lv1 = self.levels[1]
        p1 = self.hierarchy.pattern[1]
        # outside the tape (:1247-1249): R0, b1
        self._force(x, self.dev_R[0], gravity=False, pins=False)
        self._transfer(self.Ut_hat[0], 1.0, self.dev_R[0], 0.0, self.dev_B_fixed[1])
        # taped forward (:1252-1264)
        self.Ut_noOrder[0].set_weights(w)
        self._transfer(self.Ut_noOrder[0], 1.0, self.dev_x_init, 0.0, self.dev_x_solved[1])
        _lib.check(self.lib.femb_prepare_rest(lv1.n_hex, lv1.hex, self.dev_x_solved[1].data_ptr(), lv1.detJ, lv1.invJ, st()), "femb_prepare_rest")
        self._transfer(self.Ut_noOrder[0], 1.0, self.X, 0.0, self.x_gpu_layer[1])
        self._hessian(1, self.x_gpu_layer[1], self.UtAUs_values[0])
        self.PerformJacobi(layer=1, iterations=1)
        self._transfer(self.Us_hat[0], 1.0, self.dev_delta_x[1], 0.0, self.dev_delta_x[0])
        self.update_deltaX(x, self.dev_delta_x[0])
        self._force(x, tb["res"], gravity=False, pins=False)
        # loss = max |res| - alpha sum (1 - rowsum)^2 ; backward
        n0, n1 = self.dims[0], self.dims[1]
        _lib.check(self.lib.femb_inf_norm(3 * n0, tb["res"].data_ptr(), tb["maxv"].data_ptr(), st()), "femb_inf_norm")
        _lib.check(self.lib.femb_absmax_onehot(3 * n0, tb["res"].data_ptr(), tb["maxv"].data_ptr(), tb["onehot"].data_ptr(), st()), "femb_absmax_onehot")
        # d res / d x = -K(x) (elastic Hessian at the updated state, index order like res): lam0 = -K onehot
        self._hessian(0, x, tb["A0"])
        _lib.check(self.lib.femb_spmv(C.byref(self.levels[0]), tb["A0"].data_ptr(), _lib.PART_ALL, -1.0, tb["onehot"].data_ptr(), 0.0,
                                      tb["lam0"].data_ptr(), st()), "femb_spmv")
        self._transfer(self.Ut_hat[0], 1.0, tb["lam0"], 0.0, tb["lam1"])                       # Us_hat^T = Ut_hat
        D = self.UtAUs_values[0][p1.off_d:p1.off_d + n1]
        _lib.check(self.lib.femb_solve3x3_vjp(n1, D.data_ptr(), self.dev_B_fixed[1].data_ptr(), tb["lam1"].data_ptr(), tb["dbar_i"].data_ptr(), st()),
                   "femb_solve3x3_vjp")
        tb["dbar_v"].copy_(tb["dbar_i"][self.dev_vertex2index[1].long()])                       # index order -> vertex order
        _lib.check(self.lib.femb_coarse_diag_vjp(C.byref(lv1), self.x_gpu_layer[1].data_ptr(), self.dev_x_solved[1].data_ptr(), self.mu, self.la,
                                                 tb["dbar_v"].data_ptr(), tb["G"].data_ptr(), tb["lam_x1"].data_ptr(), tb["lam_r1"].data_ptr(), st()),
                   "femb_coarse_diag_vjp")
        _lib.check(self.lib.femb_exp2_weight_grad(C.byref(self.Ut_noOrder[0].c), w.data_ptr(), tb["lam_x1"].data_ptr(), self.X.data_ptr(),
                                                  tb["lam_r1"].data_ptr(), self.dev_x_init.data_ptr(), alpha, tb["grad"].data_ptr(),
                                                  tb["reg"].data_ptr(), st()), "femb_exp2_weight_grad")
        loss = float(tb["maxv"].item()) - alpha * float(tb["reg"].sum().item())
        grad = tb["grad"].clone()
        if update:
            _lib.check(self.lib.femb_exp2_weight_update(w.numel(), w.data_ptr(), tb["grad"].data_ptr(), lr, st()), "femb_exp2_weight_update")
        return loss, grad

    def train(self, iterations=100, perturb_vertex=732, perturbations=None, alpha=0.1, lr=1e-3, save_path=None, seed=None):
        """train(): every iteration restarts from the rest state with vertex `perturb_vertex` displaced by (rand(3) - 0.5) * 0.002
        (:1239-1244; `perturbations` [iterations,3] replaces the unseeded draw), runs the taped step and updates the weights.
        Returns the list of losses; `save_path` = save_p_hat."""
        g = torch.Generator().manual_seed(seed) if seed is not None else None
        losses = []
        for it in range(iterations):
            x0 = self.dev_x_init.clone()
            d = (torch.as_tensor(perturbations[it], dtype=torch.float32) if perturbations is not None
                 else (torch.rand(3, generator=g) - 0.5) * 0.002)
            if perturb_vertex is not None and perturb_vertex < self.N_verts:
                x0[perturb_vertex] += d.to(self.device)
            loss, _ = self.train_step(x0, alpha=alpha, lr=lr)
            losses.append(loss)
        if save_path is not None:
            self.save_p_hat(save_path)
        return losses
