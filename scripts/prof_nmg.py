"""Two Newton-multigrid steps of the cube, eagerly (no graph), for ncu kernel filters."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fem_simulation_b200 import Object
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
obj = Object(f"cube:{n}", 1.0 / n, [], use_graphs=False)
obj.NewtonMultigrid(2)
torch.cuda.synchronize()
