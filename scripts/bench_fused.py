"""Assembly stage on the cube: one-kernel lattice path (femb_fused.cu) vs element kernel + row fold."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fem_simulation_b200 import Object

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
chunk_list = [int(c) for c in (sys.argv[3].split(",") if len(sys.argv) > 3 else ["0"])]
obj = Object(f"cube:{n}", 1.0 / n, [], use_graphs=False, layers=2, fused_assembly=True)
x = obj.x_gpu_layer[0]
x.add_(torch.randn_like(x) * 1e-4)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
H = obj.hierarchy
asm_bytes = 352 * H.hexs[0].shape[0] + 36 * H.pattern[0].nnz + 256 * H.hexs[0].shape[0]


def timed(fn, label, nbytes=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(reps):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    us = ev[0].elapsed_time(ev[1]) / reps * 1e3
    extra = f"  {nbytes / us / 1e3:8.1f} GB/s algorithmic" if nbytes else ""
    print(f"{label:58s} {us:9.1f} us{extra}", flush=True)


lv0 = obj.level_without_lattice(0)
va, vb = obj.new_values(0), obj.new_values(0)
fa, fb = torch.zeros_like(x), torch.zeros_like(x)
obj._force_hessian(x, fa, va, obj.diag_fixed)
obj._force_hessian(x, fb, vb, obj.diag_fixed, level=lv0)
torch.cuda.synchronize()
print("max |fused - two-kernel| values", float((va - vb).abs().max()), "scale", float(vb.abs().max()),
      "force", float((fa - fb).abs().max()), "scale", float(fb.abs().max()), flush=True)
for c in chunk_list:
    obj.levels[0].lat_chunks = c
    timed(lambda: obj._hessian(0, x, obj.MF_value_gpu, obj.diag_fixed), f"hessian level0 fused chunks={c}", asm_bytes)
    timed(lambda: obj._force_hessian(x, obj.grad_gpu, obj.MF_value_gpu, obj.diag_fixed), f"force+hessian level0 fused chunks={c}", asm_bytes)
obj.levels[0].lat_chunks = 0
timed(lambda: obj._hessian(0, x, obj.MF_value_gpu, obj.diag_fixed, level=lv0), "hessian level0 two-kernel", asm_bytes)
timed(lambda: obj._force_hessian(x, obj.grad_gpu, obj.MF_value_gpu, obj.diag_fixed, level=lv0), "force+hessian level0 two-kernel", asm_bytes)
timed(lambda: obj._hessian(1, obj.x_gpu_layer[1], obj.UtAUs_values[0]), "hessian level1 fused")
timed(lambda: obj._hessian(1, obj.x_gpu_layer[1], obj.UtAUs_values[0], level=obj.level_without_lattice(1)), "hessian level1 two-kernel")
timed(lambda: obj._force(x, obj.grad_gpu), "force level0")
