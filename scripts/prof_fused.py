"""One launch of each assembly path on the cube for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fem_simulation_b200 import Object
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
obj = Object(f"cube:{n}", 1.0 / n, [], use_graphs=False, layers=2, fused_assembly=True)
x = obj.x_gpu_layer[0]
x.add_(torch.randn_like(x) * 1e-4)
for _ in range(2):
    obj._force_hessian(x, obj.grad_gpu, obj.MF_value_gpu, obj.diag_fixed)
torch.cuda.synchronize()
print("ok")
