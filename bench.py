#!/usr/bin/env python
"""bench.py -- FAS-multigrid (default) / Newton-multigrid quasi-static step throughput on the synthetic voxel cube
(BASELINE.json metric: "Newton/FAS-MG steps/sec and HBM GB/s", configs[4]).

    python bench.py --gpus N --steps K --warmup W [--impl ours|reference] [--n 100] [--method fasv3|nmg]

One "step" = one loop body of ``Object.FASv3`` (object.py:1467-1511 of the reference: fine force + Hessian assembly,
ONE symmetric 8-colour GS iteration on the fine level, state and residual restriction, re-discretised coarse Hessian with
the coarse pin penalty, ONE Jacobi iteration on the coarse level, prolongation; no post-smoothing) or, with
``--method nmg``, of ``Object.NewtonMultigrid`` (object.py:1231-1296: force + element Hessians + assembly, Galerkin
coarse operators, one V-cycle with 3+3 symmetric GS per level, full Newton update; also reported in the default line as
``newton_mg``).  Workload at N=1: cube n=100 (1 000 000 hexahedra, 1 030 301 vertices, 3.09 M dof), the reference's
level-count formula (5 levels).  ``value`` = million hexahedron-steps per second (= steps/s on the 1 M-hexahedron cube)
with all inputs resident in HBM (CUDA-graph replay, CUDA events); ``e2e`` = the same step through the public API with
the state vector coming from pinned host memory and going back every step.  N>1: one process per GPU, each rank owns
one z-slab of an N-times-longer box (weak scaling, NCCL halo exchange; see DESIGN.md "multi-GPU").

``--impl reference`` times the C/OpenMP restatement of the reference algorithm (the reference has no CPU path and
cannot run here: SURVEY.md 8c) on ALL host cores of the box (the thread count is set explicitly: torchrun exports
OMP_NUM_THREADS=1) on the stated workload at N=1 (a few steps of the same n^3 cube) and, at N>1, on one rank's n^3 block
of the N-times-longer box (labelled ``extrapolated``).

Every ``ours`` line carries ``parity``: the state after 2 steps from the rest state on the GPU path vs the same 2 steps of
the C oracle on the same mesh (positions, max-norm relative), so that a throughput number never travels without a
correctness number measured in the same process.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", "--edge", dest="n", type=int, default=100,
                    help="voxels per cube edge (100 -> 1 M hexahedra); use the spelling --edge under torchrun (its parser claims --n)")
    ap.add_argument("--method", default="fasv3", choices=["nmg", "fasv3"],
                    help="fasv3 (BASELINE configs[4], the mesh-partitioned series) or nmg (Newton-multigrid, the north-star target)")
    ap.add_argument("--strong", action="store_true",
                    help="N > 1: partition ONE n^3 cube over the ranks (strong scaling) instead of an n x n x (n N) box (weak)")
    ap.add_argument("--comm", default="auto", choices=["auto", "p2p", "nccl", "torch"],
                    help="auto: C-ABI exchanges inside one CUDA graph per rank, through peer memory (NVLink stores + flags) when every "
                         "rank can map its neighbours, else NCCL send/recv; p2p / nccl: force one; torch: torch.distributed exchanges "
                         "between segment graphs")
    ap.add_argument("--layers", type=int, default=0, help="0 = the reference's level-count formula")
    ap.add_argument("--cpu-n", type=int, default=0, help="cube edge of the CPU-baseline sample (0 = the benchmark's own n)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the 2-step comparison with the C oracle after the timed region")
    ap.add_argument("--no-nmg", action="store_true", help="skip the extra Newton-multigrid timing at N=1")
    ap.add_argument("--corrected-hierarchy", action="store_true",
                    help="coarsen level l with spacing dx*2^l (SURVEY 8f-3) instead of the reference's 2*dx for every l>=1; labelled in config")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = set()
        for r in self.rows:
            for k, nm in enumerate(names):
                if len(r) > 5 + k and r[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# DRAM bytes (read + write) per launch measured with `ncu --set full` on the cube n=100 (profiles/README.md), per million hexahedra
ASM_TRAFFIC_PER_MHEX = 4.22e9      # element_kernel<1,1,0,1,1> 0.364 + 1.381 GB, fold_lattice_kernel 1.527 + 0.947 GB (profiles/r02c_assembly_blocked_tma_ncu_raw.txt)
GS_TRAFFIC_PER_MHEX = 142e6        # average colour launch of a 3-iteration sweep: 154.8 MB per L|U launch (profiles/r01b_gs_color_ncu_raw.txt), ~80 MB per U-only launch


def pass_bytes(V, Z):
    """Algorithmic bytes of one full-matrix pass: values+col of every block, row pointer, b, x, out."""
    return Z * (36 + 4) + V * (4 + 12 + 12 + 12)


def step_bytes(obj, method):
    """Algorithmic bytes of one step from the actual V, Z, N_h of every level (SURVEY.md 8d formulas)."""
    H = obj.hierarchy
    V = H.dims
    Z = [p.nnz for p in H.pattern]
    Nh = [h.shape[0] for h in H.hexs]
    L = H.layer
    force = 352 * Nh[0] + 24 * V[0]
    assembly = 352 * Nh[0] + Z[0] * 36 + 256 * Nh[0]
    if method == "nmg":
        total = force + assembly
        for l in range(L):
            gs_it = 3 if l == L - 1 else 6
            total += (2 * gs_it - 0.5 * (1 if l == L - 1 else 2)) * pass_bytes(V[l], Z[l])     # first half-sweep skips L
            if l < L - 1:
                nnzP = H.Ut_hat[l].nnz
                total += pass_bytes(V[l], Z[l])                                                  # residual
                total += 2 * (nnzP * 8 + (V[l] + V[l + 1]) * 12 + V[l + 1] * 4)                   # restrict + prolong
                total += pass_bytes(V[l], Z[l])                                                  # rhs update
                total += Z[l] * 40 + Z[l + 1] * 36                                               # Galerkin RAP
        return total
    nnzP = H.Ut_hat[0].nnz
    total = 2 * force + assembly + 1.5 * pass_bytes(V[0], Z[0])
    total += 3 * (nnzP * 8 + (V[0] + V[1]) * 12 + V[1] * 4)
    total += 352 * Nh[1] + Z[1] * 36 + 256 * Nh[1] + V[1] * (36 + 24)
    return total


def gs_launch_bytes(obj, iterations=3):
    """Average algorithmic bytes of one colour launch of the row-pass kernel in femb_gauss_seidel(level 0)."""
    H = obj.hierarchy
    total = (2 * iterations - 0.5) * pass_bytes(H.dims[0], H.pattern[0].nnz)
    return total / (16 * iterations)


def bind_to_gpu_numa_node(local_rank):
    """Pin this process to the CPUs next to its GPU (NVML's ideal affinity) before anything is allocated, so that the pinned host
    buffers of the end-to-end loop are first-touched on the GPU's own NUMA node (8 ranks sharing one node was what limited the
    round-1 e2e scaling).  Returns a short description for the bench line; never fatal."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        cpus = sorted(os.sched_getaffinity(0))
        return f"process bound to the GPU's NUMA-local CPUs ({len(cpus)} cpus: {cpus[0]}-{cpus[-1]})"
    except Exception as e:                                   # noqa: BLE001
        return f"not bound ({type(e).__name__})"


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    numa = bind_to_gpu_numa_node(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    from fem_simulation_b200 import Object, _lib
    lib = _lib.load()
    n = args.n
    method = args.method
    t0 = time.time()
    nz_total = n if (args.strong or world == 1) else n * world
    if world > 1:
        # weak scaling: every rank owns an n x n x n block of an n x n x (n*world) box, cut into z-slabs;
        # strong scaling: the n^3 cube itself is cut
        from fem_simulation_b200 import hierarchy as _h
        from fem_simulation_b200.distributed import SlabObject
        import numpy as _np
        glayers = 2 if method == "fasv3" else (args.layers or _h.level_count(_np.zeros(3, _np.float32), _np.array([1.0, 1.0, nz_total / n], _np.float32), 1.0 / n))
        so = SlabObject(n, n, nz_total, 1.0 / n, device=str(dev), layers=glayers, comm=args.comm,
                        corrected_hierarchy=args.corrected_hierarchy)
        obj = so.obj
        step = so.fasv3_step if method == "fasv3" else so.nmg_step
    else:
        obj = Object(f"cube:{n}", 1.0 / n, [], device=str(dev), layers=(args.layers or None), use_graphs=True,
                     corrected_hierarchy=args.corrected_hierarchy)
        body = obj._newton_mg_body if method == "nmg" else (lambda: obj._fas_v123_body(obj.x_gpu_layer[0], False, True, True, True))
        key = "nmg" if method == "nmg" else "fasv3"
        step = lambda: obj._run_step(key, body)
    setup_s = time.time() - t0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def maxed(ms):
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- warm-up (first call also captures the CUDA graph on the 1-GPU path) --------------------
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = lib.femb_launch_count()
    if world == 1:
        obj.use_graphs = False
        step()                                     # one eager step only to count this step's launches
        obj.use_graphs = True
    else:
        so.graph_segments = False                  # one eager step to count the launches behind the graphs
        so.n_exchanges = 0
        step()
        so.graph_segments = True
        exchanges_per_step = so.n_exchanges
    launches_per_step = lib.femb_launch_count() - l0
    barrier()
    # ---- device-resident timing -------------------------------------------------------------------
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    ms = maxed(ev0.elapsed_time(ev1))
    # ---- end-to-end: state from pinned host memory in, new state back out, every step ------------
    x_host = obj.x_gpu_layer[0].cpu().pin_memory()
    out_host = torch.empty_like(x_host).pin_memory()
    barrier()
    ev0.record()
    for _ in range(args.steps):
        obj.x_gpu_layer[0].copy_(x_host, non_blocking=True)
        step()
        out_host.copy_(obj.x_gpu_layer[0], non_blocking=True)
        x_host, out_host = out_host, x_host
    ev1.record()
    barrier()
    ms_e2e = maxed(ev0.elapsed_time(ev1))
    # ---- dominant stage of the FAS step: femb_hessian = element_kernel + fold_lattice_kernel (timed through the C ABI) ----
    reps = 5
    vals_tmp = obj.MF_value_gpu
    for _ in range(2):
        obj._hessian(0, obj.x_gpu_layer[0], vals_tmp, obj.diag_fixed)
    torch.cuda.synchronize(dev)
    ev0.record()
    for _ in range(reps):
        obj._hessian(0, obj.x_gpu_layer[0], vals_tmp, obj.diag_fixed)
    ev1.record()
    torch.cuda.synchronize(dev)
    asm_ms = ev0.elapsed_time(ev1) / reps
    # ---- dominant kernel of the smoother: one colour launch of femb_gauss_seidel on the fine level ----------------
    obj.PerformGaussSeidel(0, 3)
    torch.cuda.synchronize(dev)
    ev0.record()
    for _ in range(reps):
        obj.PerformGaussSeidel(0, 3)
    ev1.record()
    torch.cuda.synchronize(dev)
    gs_ms_per_launch = ev0.elapsed_time(ev1) / (reps * 48)
    # ---- the Newton-multigrid step as well (1 GPU): the other half of BASELINE.json's metric --------
    nmg = None
    if world == 1 and method == "fasv3" and not args.no_nmg:
        for _ in range(3):
            obj._run_step("nmg", obj._newton_mg_body)
        torch.cuda.synchronize(dev)
        ev0.record()
        for _ in range(args.steps):
            obj._run_step("nmg", obj._newton_mg_body)
        ev1.record()
        torch.cuda.synchronize(dev)
        t_nmg = ev0.elapsed_time(ev1) / args.steps
        sbn = step_bytes(obj, "nmg")
        nmg = {"ms_per_step": t_nmg, "steps_per_sec": 1e3 / t_nmg, "algorithmic_bytes_per_step": sbn,
               "achieved_GBps": sbn / (t_nmg * 1e-3) / 1e9}
    parity = None
    if world > 1 and not args.no_parity:
        parity = parity_check_distributed(args, so, step, dev, rank, world)
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    H = obj.hierarchy
    ach = gs_launch_bytes(obj) / (gs_ms_per_launch * 1e-3) / 1e9
    asm_bytes = 352 * H.hexs[0].shape[0] + 36 * H.pattern[0].nnz + 256 * H.hexs[0].shape[0]
    sb = step_bytes(obj, method)
    total_hexes = n * n * nz_total
    hexes_per_rank = total_hexes // world
    mhex = total_hexes / 1e6
    steps_per_s = args.steps / (ms * 1e-3)
    if nmg is not None:
        nmg["frac_of_peak"] = nmg["achieved_GBps"] / peak
    line = {
        "metric": "fas_mg_mhex_steps_per_sec" if method == "fasv3" else "newton_mg_mhex_steps_per_sec",
        "value": steps_per_s * mhex, "unit": "Mhex-steps/s (million hexahedron-steps per second; = steps/s on a 1 M-hexahedron mesh)",
        "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong" if (args.strong and world > 1) else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": (f"voxel box {n}x{n}x{nz_total}: {total_hexes} hexahedra, "
                                f"{'FASv3 two-level quasi-static step (object.py:1467-1511)' if method == 'fasv3' else 'Newton-multigrid step (V-cycle 3+3 sym. 8-colour GS, Galerkin; object.py:1231-1296)'}, "
                                "mu=250 la=0 g=-4.9 k=100, default top-layer pins"),
                   "steps_per_sec": steps_per_s, "hexahedra": total_hexes, "local_levels": H.layer, "local_dims": H.dims,
                   "hierarchy": ("corrected spacing dx*2^l (not the reference's behaviour)" if args.corrected_hierarchy else
                                 "reference coarsening (object.py:175-179: levels >= 3 repeat level 2)"),
                   "parallelism": (f"{world} ranks, z-slab domain decomposition (cuts {so.cuts}, {so.overlap} overlap cell layers below each slab), "
                                   f"{exchanges_per_step} neighbour halo exchanges per step (vector ghosts only, no matrix entries), no all-reduce "
                                   "on the data path; " + ("C-ABI exchanges on the step's stream, transport: " + so.comm.transport + "; ONE CUDA graph per "
                                   "rank per step" if so.comm is not None else "torch.distributed exchanges between CUDA-graph segments")) if world > 1 else
                                  "single GPU, CUDA-graph replay of the step body",
                   "l2": "working set (matrix values %.2f GB per rank) >> 126 MB L2, no flush needed" % (H.pattern[0].nnz * 36 / 1e9),
                   "setup_s": round(setup_s, 1), "host_affinity": numa, "timing": "CUDA events, barrier + synchronize on both sides, max over ranks"},
        "e2e": {"value": args.steps / (ms_e2e * 1e-3) * mhex, "unit": "Mhex-steps/s", "h2d_bytes_per_step": H.dims[0] * 12 * world,
                "d2h_bytes_per_step": H.dims[0] * 12 * world},
        "gpu_launches": int(launches_per_step * args.steps * world),
        "step_hbm": {"algorithmic_bytes_per_step_per_rank": sb, "achieved_GBps_per_rank": sb / (ms / args.steps * 1e-3) / 1e9,
                     "frac_of_peak": sb / (ms / args.steps * 1e-3) / 1e9 / peak},
        "roofline": {"bound": "hbm", "kernel": "femb_hessian: element_kernel<1,1> + fold_lattice_kernel (level 0, the two launches of "
                     "the assembly stage; the element kernel is bound by the L1 / shared-memory pipe, see DESIGN.md)",
                     "achieved": asm_bytes / (asm_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": asm_bytes / (asm_ms * 1e-3) / 1e9 / peak, "traffic": ASM_TRAFFIC_PER_MHEX * H.hexs[0].shape[0] / 1e6,
                     "traffic_note": "dram__bytes_read+write of the two kernels, ncu --set full at cube n=100 (profiles/), scaled by hexahedra",
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                     "us_per_launch": asm_ms * 1e3, "algorithmic_bytes_per_launch": asm_bytes,
                     "algorithmic_bytes_formula": "352*N_h + 36*Z + 256*N_h (SURVEY.md 8d, Hessian assembly)"},
        "roofline_smoother": {"bound": "hbm", "kernel": "gs_color_kernel<8,16> (femb_gauss_seidel, fine level, one colour launch)",
                              "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": GS_TRAFFIC_PER_MHEX * H.hexs[0].shape[0] / 1e6,
                              "us_per_launch": gs_ms_per_launch * 1e3, "algorithmic_bytes_per_launch": gs_launch_bytes(obj)},
        "clocks": clocks,
    }
    if nmg is not None:
        line["newton_mg"] = nmg
    if world > 1 and not args.no_parity:
        line["parity"] = parity
    if world == 1 and not args.no_parity:
        line["parity"], co = parity_check(args, obj, step, H)
        if not args.no_cpu_baseline:
            co.x[...] = obj.x_cpu.numpy()
            line["cpu_baseline"] = cpu_baseline(args, H, co=co if not args.cpu_n else None)
    elif not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(args, H)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


try:
    _ALL_CPUS = set(os.sched_getaffinity(0))          # before bind_to_gpu_numa_node narrows it
except Exception:                                     # noqa: BLE001
    _ALL_CPUS = set()


def host_threads():
    """Threads of the CPU legs: every CPU this process was given at start (the NUMA binding of the GPU arm is lifted for them)."""
    if _ALL_CPUS:
        try:
            os.sched_setaffinity(0, _ALL_CPUS)
        except Exception:                             # noqa: BLE001
            pass
        return len(_ALL_CPUS)
    return max(1, os.cpu_count() or 1)


def cpu_oracle(args, n, H=None):
    """The C/OpenMP restatement of the reference algorithm (oracle/femb_oracle.c) on the n^3 cube, on all host cores."""
    from fem_simulation_b200 import hierarchy, mesh
    from oracle import c_oracle
    pts, hx = mesh.cube_mesh(n)
    if H is None:
        H = hierarchy.build_hierarchy(pts, hx, 1.0 / n, args.layers or None, corrected_spacing=args.corrected_hierarchy)
    plans = [hierarchy.build_galerkin_plan(H, l) for l in range(H.layer - 1)] if args.method == "nmg" else []
    co = c_oracle.from_hierarchy(H, plans, pts, 1.0 / n)
    co._plans_src = plans
    co.set_threads(host_threads())
    return co, H, hx.shape[0]


def cpu_baseline(args, H=None, steps=3, warmup=1, budget_s=30.0, co=None):
    """CPU baseline on a bounded sample: `steps` steps (fewer if the time budget runs out) of the benchmark's own n^3 cube
    -- the N=1 workload itself; at N>1 one rank's block of the longer box."""
    n = args.cpu_n or args.n
    if co is None:
        co, H, nh = cpu_oracle(args, n, H if n == args.n else None)
    else:
        nh = n ** 3
    fn = co.newton_mg_step if args.method == "nmg" else co.fasv3_step
    for _ in range(warmup):
        fn()
    t0 = time.time()
    done = 0
    while done < steps and (done == 0 or time.time() - t0 < budget_s):
        fn()
        done += 1
    dt = (time.time() - t0) / done
    return {"value": nh / 1e6 / dt, "unit": "Mhex-steps/s", "cores": co.threads, "kind": "port",
            "sample": f"oracle/femb_oracle.c (C + OpenMP, {co.threads} threads, atomics-free fixed-pattern assembly), {done} timed "
                      f"step(s) after {warmup} warm-up on cube n={n} ({nh} hexahedra, {co.L} levels), {dt:.3f} s/step",
            "s_per_sample_step": dt, "sample_steps": done, "sample_n": n}


def parity_check_distributed(args, so, step, dev, rank, world, steps=3):
    """N > 1: (a) the same code path on a SMALL box cut over the N ranks vs rank 0 solving that box alone on its GPU (max-norm
    relative difference of the positions after `steps` steps); (b) fingerprints of the full-size run (owned-vertex checksum,
    |f|inf) so that the throughput line carries the state it was measured on."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from fem_simulation_b200 import Object, mesh
    from fem_simulation_b200.distributed import SlabObject, hierarchy_align
    method = args.method
    layers = 2 if method == "fasv3" else 3
    m = 12
    nz = max(hierarchy_align(layers) * 2, 8) * world
    small = SlabObject(m, m, nz, 1.0 / m, device=str(dev), layers=layers, comm=args.comm)
    for _ in range(steps):
        small.fasv3_step() if method == "fasv3" else small.nmg_step()
    ids, xs = small.owned_positions()
    n_glob = (m + 1) * (m + 1) * (nz + 1)
    xg = torch.zeros((n_glob, 3), device=dev)
    xg[ids] = xs
    dist.all_reduce(xg)
    out = {"small_box": f"{m}x{m}x{nz}, {layers} levels, {steps} steps, {world} ranks vs 1 GPU"}
    if rank == 0:
        pts, hx = mesh.box_mesh(m, m, nz, 1.0 / m)
        ref = Object((pts, hx), 1.0 / m, [], device=str(dev), layers=layers, use_graphs=False)
        ref.FASv3(steps) if method == "fasv3" else ref.NewtonMultigrid(steps)
        xr = ref.x_gpu_layer[0]
        out["x_rel_err_vs_1gpu"] = float((xg - xr).abs().max() / xr.abs().max())
        out["displacement_rel_err_vs_1gpu"] = float((xg - xr).abs().max() / (xr - torch.from_numpy(pts).to(dev)).abs().max())
    dist.barrier()
    out["full_size_x_checksum"] = so.checksum()
    out["full_size_force_inf_norm"] = so.force_inf_norm()
    return out


def parity_check(args, obj, step, H, steps=2):
    """State after `steps` steps from the rest state: GPU path vs the C oracle on the same mesh (1 GPU)."""
    import numpy as np
    import torch
    obj.x_gpu_layer[0].copy_(obj.x_cpu)
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
    xg = obj.x_gpu_layer[0].cpu().numpy().astype(np.float64)
    fg = obj.inf_norm(obj.grad_gpu if args.method == "nmg" else obj.dev_B_fixed[0])
    co, _, _ = cpu_oracle(args, args.n, H)
    fn = co.newton_mg_step if args.method == "nmg" else co.fasv3_step
    for _ in range(steps):
        fn()
    xc = co.x.astype(np.float64)
    fc = float(np.abs(co.grad if args.method == "nmg" else co.vec[0]["B_fixed"]).max())
    x0 = obj.x_cpu.numpy().astype(np.float64)
    # what fp32 itself is worth here: the same oracle in double precision (the per-vertex forces of this mesh sit at the
    # 1e-10 cut-offs of solve3x3, cublas.py:67-86, so round-off switches rows on and off)
    from fem_simulation_b200 import mesh
    from oracle import c_oracle
    c64 = c_oracle.from_hierarchy(H, co._plans_src, mesh.cube_mesh(args.n)[0], 1.0 / args.n, dtype=np.float64)
    c64.set_threads(host_threads())
    f64 = c64.newton_mg_step if args.method == "nmg" else c64.fasv3_step
    for _ in range(steps):
        f64()
    x64 = c64.x.astype(np.float64)
    disp = max(np.abs(x64 - x0).max(), 1e-30)
    return {"steps": steps, "x_rel_err": float(np.abs(xg - xc).max() / np.abs(xc).max()),
            "displacement_rel_err": float(np.abs(xg - xc).max() / max(np.abs(xc - x0).max(), 1e-30)),
            "displacement_rel_err_vs_fp64_oracle": float(np.abs(xg - x64).max() / disp),
            "fp32_oracle_vs_fp64_oracle": float(np.abs(xc - x64).max() / disp),
            "force_inf_norm": fg, "force_inf_norm_ref": fc, "x_checksum": float(np.abs(xg).sum()),
            "x_checksum_ref": float(np.abs(xc).sum()),
            "ref": "oracle/femb_oracle.c, same mesh, same steps from the rest state (force norm: at the start of the last step); "
                   "fp32_oracle_vs_fp64_oracle is the round-off level of the algorithm itself on this mesh, in displacement units"}, co


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t_all = time.time()
    cb = cpu_baseline(args, steps=max(1, min(args.steps, 5)), warmup=1 if args.warmup > 0 else 0, budget_s=60.0)
    n = args.n
    hexes = n ** 3 * args.gpus
    extrapolated = args.gpus > 1 or cb["sample_n"] != n
    what = ('FASv3 two-level quasi-static step (object.py:1467-1511)' if args.method == 'fasv3' else
            'Newton-multigrid step (V-cycle 3+3 sym. 8-colour GS, Galerkin; object.py:1231-1296)')
    line = {"impl": "reference", "metric": "fas_mg_mhex_steps_per_sec" if args.method == "fasv3" else "newton_mg_mhex_steps_per_sec",
            "value": cb["value"], "unit": "Mhex-steps/s (million hexahedron-steps per second; = steps/s on a 1 M-hexahedron mesh)",
            "n_gpus": args.gpus, "steps": cb["sample_steps"], "warmup": 1 if args.warmup > 0 else 0,
            "ms_per_step": 1e3 * cb["s_per_sample_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "extrapolated": extrapolated,
            "config": {"workload": (f"voxel box {n}x{n}x{n * args.gpus}: {hexes} hexahedra, {what}, "
                                    "mu=250 la=0 g=-4.9 k=100, default top-layer pins"),
                       "timed": (f"cube n={cb['sample_n']} ({cb['sample_n'] ** 3} hexahedra), {cb['sample_steps']} steps; ms_per_step is that "
                                 "sample's; value = sample hexahedra / sample time" +
                                 ("" if not extrapolated else " (one rank's block of the longer box: throughput is size-independent to first order)")),
                       "arm": "C/OpenMP restatement of the reference algorithm on all host cores (thread count set explicitly); "
                              "the reference itself is GPU-only Warp code and cannot run here (SURVEY.md 8c)",
                       "wall_s": round(time.time() - t_all, 1)},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "Mhex-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
