"""One Gauss-Seidel iteration on the cube (for `ncu -k regex:gs_color`): launches 0-7 = backward sweep (U only), 8-15 = forward sweep."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fem_simulation_b200 import Object
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
obj = Object(f"cube:{n}", 1.0 / n, [], use_graphs=False, layers=2)
x = obj.x_gpu_layer[0]
x.add_(torch.randn_like(x) * 1e-4)
obj._force_hessian(x, obj.dev_B_fixed[0], obj.MF_value_gpu, obj.diag_fixed)
obj.PerformGaussSeidel(0, 2)
torch.cuda.synchronize()
