"""Voxelise the reference's bundled `assets/objs/horse.obj` at dx = 0.05 (BASELINE config 0: "the repo's bundled mesh") with
fem_simulation_b200.voxelize and store the mesh as a fixture: /root/reference does not exist on the GPU box.
Run in the build container:  python tests/golden/make_horse_fixture.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from fem_simulation_b200 import voxelize  # noqa: E402

pts, hx = voxelize.voxelize_obj("/root/reference/assets/objs/horse.obj", 0.05)
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "horse_0.05.npz")
np.savez_compressed(out, points=pts.astype(np.float32), hexagons=hx.astype(np.int32), dx=np.float32(0.05))
print(out, pts.shape, hx.shape)
