"""One assembly of the cube (lattice fold, then the fold-plan path) for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fem_simulation_b200 import Object
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
obj = Object(f"cube:{n}", 1.0 / n, [], use_graphs=False, layers=2)
x = obj.x_gpu_layer[0] + torch.randn_like(obj.x_gpu_layer[0]) * 1e-4
for _ in range(2):
    obj._hessian(0, x, obj.MF_value_gpu, obj.diag_fixed)
    obj._force_hessian(x, obj.grad_gpu, obj.MF_value_gpu, obj.diag_fixed)
    obj._hessian(0, x, obj.MF_value_gpu, obj.diag_fixed, level=obj.level_without_lattice(0))
torch.cuda.synchronize()
