"""Micro-benchmark of the assembly stage (element Hessians + fold) on the cube: fused vs two-kernel path."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fem_simulation_b200 import Object

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
modes = [int(c) for c in (sys.argv[3] if len(sys.argv) > 3 else "101")]
chunk = int(os.environ.get("PIPE_CHUNK", 2 * 740 * 16))
obj = Object(f"cube:{n}", 1.0 / n, [], use_graphs=False, layers=2, pipe_chunk_hexes=chunk)
print("chunks", obj.levels[0].pipe_chunks, "chunk hexes", chunk)
x = obj.x_gpu_layer[0]
x.add_(torch.randn_like(x) * 1e-4)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]


def timed(fn, label):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(reps):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    print(f"{label:40s} {ev[0].elapsed_time(ev[1]) / reps * 1e3:9.1f} us", flush=True)


for fused in modes:
    obj.lib.femb_set_pipeline(fused)
    timed(lambda: obj._hessian(0, x, obj.MF_value_gpu, obj.diag_fixed), f"hessian level0 pipeline={fused}")
    timed(lambda: obj._force_hessian(x, obj.grad_gpu, obj.MF_value_gpu, obj.diag_fixed), f"force+hessian level0 pipeline={fused}")
    timed(lambda: obj._hessian(1, obj.x_gpu_layer[1], obj.UtAUs_values[0]), f"hessian level1 pipeline={fused}")
obj.lib.femb_set_pipeline(0)
for rpw, mb in ((4, 4), (2, 4), (8, 4), (1, 4), (4, 5), (2, 5), (8, 5), (4, 4)):
    obj.lib.femb_set_fold_config(rpw, mb)
    timed(lambda: obj._hessian(0, x, obj.MF_value_gpu, obj.diag_fixed), f"hessian level0 fold rows/warp={rpw} min_blocks={mb}")
obj.lib.femb_set_fold_config(4, 4)
timed(lambda: obj._force(x, obj.grad_gpu), "force level0")
if os.environ.get("ASM_ONLY"):
    sys.exit(0)
if os.environ.get("GALERKIN_ONLY"):
    obj.galerkin()
    timed(obj.galerkin, "galerkin 0->1")
    sys.exit(0)
obj._force_hessian(x, obj.dev_B_fixed[0], obj.MF_value_gpu, obj.diag_fixed)
for rows, subs, pdl in (((32, 4, 1), (16, 8, 1), (8, 8, 1), (8, 4, 1), (16, 4, 1), (16, 8, 0)) if len(modes) > 1 else ()):
    obj.lib.femb_set_pdl(pdl)
    obj.lib.femb_set_gs_threads_per_row(subs)
    obj.lib.femb_set_gs_rows_per_cta(rows)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        obj.PerformGaussSeidel(0, 3)
    timed(g.replay, f"GS level0 x3 rows/CTA={rows} threads/row={subs} pdl={pdl} (graph)")
    timed(lambda: obj._residual(0, obj.dev_B_fixed[0], obj.dev_delta_x[0], obj.dev_R[0]), f"   residual level0 (full pass) rows/CTA={rows} threads/row={subs}")
obj.lib.femb_set_gs_threads_per_row(8)
obj.lib.femb_set_gs_rows_per_cta(16)
timed(lambda: obj._residual(0, obj.dev_B_fixed[0], obj.dev_delta_x[0], obj.dev_R[0]), "residual level0 (full pass)")
obj.galerkin()
timed(obj.galerkin, "galerkin 0->1")

# coarse levels: cooperative single launch vs PDL colour launches
del obj
torch.cuda.empty_cache()
obj = Object(f"cube:{n}", 1.0 / n, [], use_graphs=False, layers=3)
obj._force_hessian(obj.x_gpu_layer[0], obj.dev_B_fixed[0], obj.MF_value_gpu, obj.diag_fixed)
obj.galerkin()
for lvl in (1, 2):
    obj.dev_B_fixed[lvl].normal_()
    for lim in (-1, 0, -1):
        obj.lib.femb_set_gs_coop_max_blocks(lim)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            obj.PerformGaussSeidel(lvl, 3)
        timed(g.replay, f"GS level{lvl} x3 coop_max_blocks={lim} (graph)")
obj.lib.femb_set_gs_coop_max_blocks(-1)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    obj.galerkin()
timed(g.replay, "galerkin 0->1->2 (graph)")
