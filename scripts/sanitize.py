"""Tiny end-to-end run for compute-sanitizer (memcheck / racecheck): every kernel of the Newton-multigrid and FASv3 steps on a
ragged 3-level mesh, eagerly (no CUDA graphs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fem_simulation_b200 import Object, mesh

pts, hx = mesh.jagged_mesh(18, 17, 19, 0.05, seed=2, keep=0.7, origin=(-0.3, 0.1, 0.2))
obj = Object((pts, hx), 0.05, [], use_graphs=False, pipe_chunk_hexes=160)
obj.NewtonMultigrid(1)
obj.FASv3(1)
obj.lib.femb_set_pipeline(1)
obj.FASv3(1)
obj.lib.femb_set_pipeline(0)
obj.Newton(1)
print("energy", obj.total_energy(), "launches", obj.lib.femb_launch_count())
torch.cuda.synchronize()
