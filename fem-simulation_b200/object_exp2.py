"""``Exp2Object`` -- forward path of exp2's ``compare`` (classic vs trained position-restriction P~).

Mirrors /root/reference/exp2/p_hat/quasi_simulation/object.py:1289-1357 (the canonical copy): per step
  R0 = elastic force(x) ; log |R0|inf ; b1 = Ut_hat R0 ;
  Ut_noOrder <- Kronecker(w) ; coarse REST shape re-derived from Ut_noOrder x_init (prepare_kernal) ;
  x1 = Ut_noOrder X  (X = the state copied at construction, :825-826 -- not the current x) ;
  coarse Hessian at x1 on zeroed values (no pins) ; ONE Jacobi straight on the diagonal ; dx0 = Us_hat dx1 ; x += dx0.
The weights w (``Ut_noOrder_value``, float32[nnz(Ut_noOrder[0])]) start as the default row-normalised weights
(p_hat_init, cublas.py:254-256 there) and can be replaced by a trained ``.pth`` (``load_weights``, the format of
``assets/p/Ut_noOrder_0.pth``).  ``train()`` (Warp autodiff through this step) is out of scope.
"""
from __future__ import annotations

import torch

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
        from . import _lib
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
