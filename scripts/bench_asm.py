"""Assembly stage on the cube: lattice fold (plane-major element matrices) vs the general fold-plan path, per kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fem_simulation_b200 import Object

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
obj = Object(f"cube:{n}", 1.0 / n, [], use_graphs=False, layers=2)
x = obj.x_gpu_layer[0] + torch.randn_like(obj.x_gpu_layer[0]) * 1e-4
nh, nv = obj.N_hexagons, obj.N_verts
Z = obj.hierarchy.pattern[0].nnz
asm_bytes = 352 * nh + 36 * Z + 256 * nh


def timed(fn, what, nbytes=None, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    extra = f"  {nbytes / us / 1e3:8.1f} GB/s algorithmic" if nbytes else ""
    print(f"{what:60s} {us:9.1f} us{extra}", flush=True)
    return us


va, vb = obj.new_values(0), obj.new_values(0)
fa, fb = torch.zeros_like(x), torch.zeros_like(x)
obj._force_hessian(x, fa, va, obj.diag_fixed)
obj._force_hessian(x, fb, vb, obj.diag_fixed, level=obj.level_without_lattice(0))
torch.cuda.synchronize()
print("max |lattice - general| values", float((va - vb).abs().max()), "scale", float(vb.abs().max()),
      "force equal", bool(torch.equal(fa, fb)))
timed(lambda: obj._hessian(0, x, obj.MF_value_gpu, obj.diag_fixed), "hessian level0 lattice fold", asm_bytes)
timed(lambda: obj._force_hessian(x, obj.grad_gpu, obj.MF_value_gpu, obj.diag_fixed), "force+hessian level0 lattice fold", asm_bytes)
lv = obj.level_without_lattice(0)
timed(lambda: obj._hessian(0, x, obj.MF_value_gpu, obj.diag_fixed, level=lv), "hessian level0 fold plan", asm_bytes)
timed(lambda: obj._force_hessian(x, obj.grad_gpu, obj.MF_value_gpu, obj.diag_fixed, level=lv), "force+hessian level0 fold plan", asm_bytes)
timed(lambda: obj._hessian(1, obj.x_gpu_layer[1], obj.UtAUs_values[0]), "hessian level1 lattice fold")
timed(lambda: obj._force(x, obj.grad_gpu), "force level0")
# femb_set_assembly_async modes (TMA tile store, bulk-copied rest data, L2 prefetch): bit-identical results, per-variant time
ref = None
for store_tma in (0, 1, 2, 3):
    obj.lib.femb_set_assembly_async(store_tma)
    v = obj.new_values(0)
    f = torch.zeros_like(x)
    obj._force_hessian(x, f, v, obj.diag_fixed)
    torch.cuda.synchronize()
    if ref is None:
        ref = (v, f)
    else:
        print(f"assembly async mode={store_tma}: values equal {bool(torch.equal(v, ref[0]))} forces equal {bool(torch.equal(f, ref[1]))}", flush=True)
    timed(lambda: obj._hessian(0, x, obj.MF_value_gpu, obj.diag_fixed), f"hessian level0, assembly async mode={store_tma}", asm_bytes)
    timed(lambda: obj._force_hessian(x, obj.grad_gpu, obj.MF_value_gpu, obj.diag_fixed), f"force+hessian level0, assembly async mode={store_tma}", asm_bytes)
    timed(lambda: obj._hessian(1, obj.x_gpu_layer[1], obj.UtAUs_values[0]), f"hessian level1, assembly async mode={store_tma}")
    timed(lambda: obj._force(x, obj.grad_gpu), f"force level0, assembly async mode={store_tma}")
for mb in (3, 5, 4):
    obj.lib.femb_set_fold_config(4, mb)
    timed(lambda: obj._hessian(0, x, obj.MF_value_gpu, obj.diag_fixed), f"hessian level0 lattice fold, fold min blocks {mb}", asm_bytes)
