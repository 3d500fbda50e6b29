#!/usr/bin/env python
"""Benchmark of the MFD hot path on B200 (BASELINE.json metric):

    MFD assembly Mcells/s + PCG s/iter & solve time, 256^3 hex, 1/2/4/8 B200

    python bench.py --gpus N --steps K --warmup W          # our arm (CUDA); N>1 under torchrun
    python bench.py --impl reference --steps K --warmup W  # the reference algorithm on host cores

Workload (BASELINE.json configs[2], SURVEY.md 8d "C3"): n^3 unit-cube hex mesh (n = 256),
K = diag(exp g1, exp g2, exp g3) iid per cell (seed 256), M_E method 1, Dirichlet p=1 on x-,
p=0 on x+, no-flow elsewhere, f = 0.  One step = numeric assembly of the saddle-point CSR
(fused M_E construction + scatter) followed by the full solve (hybrid set-up, PCG to rtol,
recovery of [v;p]).  `value` is the assembly throughput (cells of all ranks / slowest rank's
time); PCG s/iter and time-to-solution ride in the same JSON line.  Inputs are resident in HBM
for `value`; `e2e` goes through pinned host buffers.  With N > 1 the cell set is split into
z-slabs (strong scaling: the 256^3 problem is fixed), see mimpy_b200/partition.py.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "mfd_assembly_mcells_per_s"
UNIT = "Mcells/s"
ASM_BYTES_PER_CELL = 700.0     # 332 B read + 368 B written per hex cell
# saturation step as built (twophase.cu), structured hex cells, r = 3 faces per cell: k_cell_frac_flow (s, p read, f_w of the
# cell written: 24 B) and k_saturation_cells<HEX> (per face: face_to_flux 4 + upwind 4 + u_t 8 + area 8; f_w of the cell's
# upwinded faces written 24; f_w of the neighbours 8; p, p_prev, phi, vol, s_w 40 + s_w written 8); every face counted once
SAT_BYTES_PER_CELL = 24.0 + 3 * 24.0 + 24 + 8 + 48


def measured_traffic(n, world):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the assembly kernel, from the committed
    `ncu --set full` capture (profiles/asm_traffic.json, written by scripts/ncu_traffic.py from the .ncu-rep).
    The record carries the MD5 of the kernel sources it was captured with: if they changed since, the number
    is stale and None is reported instead."""
    import hashlib
    p = os.path.join(ROOT, "profiles", "asm_traffic.json")
    if world != 1 or not os.path.exists(p):
        return None, "no capture for this configuration"
    try:
        rec = json.load(open(p))
        h = hashlib.md5()
        for f in rec["sources"]:
            h.update(open(os.path.join(ROOT, f), "rb").read())
        if h.hexdigest() != rec["sources_md5"]:
            return None, "kernel sources changed since the capture %s" % rec.get("capture")
        if int(rec["n"]) != int(n):
            return None, "capture is for n = %s" % rec["n"]
        return float(rec["dram_bytes_read"]) + float(rec["dram_bytes_write"]), rec.get("capture")
    except Exception as e:
        return None, "unreadable capture record: %s" % e


def auxmg_iteration_bytes(rows, nnz, nc, nf, patterns=True, blocks=True, fp32_cycle=True):
    """Bytes ONE iteration of the benched PCG (auxiliary-space multigrid preconditioner) has to move, per GPU -- the
    counts of DESIGN 3 / 4, each array touched once:
      SpMV of the face system      8 nnz + 21 rows with column patterns (12 nnz + 20 rows indexed)
      two vector kernels           88 rows (x, r, p updated; A p, z read; dots fused)
      face -> cell restriction     116 nc (r of the six faces, weights, right-hand side of level 0 written)
        + cell-block smoother      84 nc (21 floats per cell; the six residuals are in registers already)
      cell -> face prolongation    55 nf (correction of the two cells, weights, D^-1 r or the block sums, z written, r.z)
      V(1,1) cycle                 64 nc on FP32 mirrors / 136 nc on FP64 levels at level 0, x 8/7 for the coarser levels."""
    spmv = (8. * nnz + 21. * rows) if patterns else (12. * nnz + 20. * rows)
    cycle = (64. if fp32_cycle else 136.) * nc * 8. / 7.
    return spmv + 88. * rows + (116. + (84. if blocks else 0.)) * nc + 55. * nf + cycle


def aux_levels(nx, ny, nz, coarsest=64):
    """Levels of the auxiliary-space multigrid (csrc/auxmg.cu: halve until <= coarsest cells)."""
    lv = 1
    while nx * ny * nz > coarsest:
        nx, ny, nz = (nx + 1) // 2, (ny + 1) // 2, (nz + 1) // 2
        lv += 1
    return lv, (nx, ny, nz)


def launches_per_step(precond, n, nz_local, iters, world=1, rounds=1):
    """Kernels of OUR library launched by one bench step (assembly, set-up, solve): a count from the launch
    sites of csrc/ (pcg.cu:launch_iteration, auxmg.cu:mimpy_auxmg_apply, comm.cu), not a measurement."""
    comm = world > 1
    halo = 2 if comm else 0            # k_peer_halo_push + k_peer_halo_add
    red = 1 if comm else 0             # k_peer_allreduce
    setup = 1 + 4 + 1                  # assembly; fused hybrid set-up (zero, set-up, diag, invert); face rhs
    if precond != "auxmg":
        return setup + 3 + (3 + halo + 2 * red) * iters + rounds * 2
    if not comm:
        lv, _ = aux_levels(n, n, nz_local)
        cycle = 2 + 3 * (lv - 1) + 1   # face restrict / prolong, (pre-smooth, restrict, prolong) per level, coarse solve
    else:
        lv, (cx, cy, cz) = aux_levels(n, n, nz_local, 32768)   # distributed levels, then the agglomerated ones
        glv, _ = aux_levels(cx, cy, cz * world)
        cycle = 2 + 3 * (lv - 1) + 2 * (2 * (lv - 1) - 1) + 2 + 3 * (glv - 1) + 1   # + ghost-layer copies, gather, extract
    per_iter = 3 + cycle + halo * 2 + red * 2    # SpMV, two vector updates; halo on Ap and z; two scalar all-reduces
    return setup + (2 + 2 * lv) + (3 + cycle + 1) + per_iter * iters + rounds * 2   # + recover and saddle check per round


def pcg_bytes_per_row(nnz_per_row):
    return 12.0 * nnz_per_row + 108.0  # SURVEY.md 8d (indexed SpMV: the roofline numerator stays the survey's count)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        threading.Thread.__init__(self, daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.stop_flag = False
        self.max_mhz = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                self.samples.append(float(parts[0]))
                self.max_mhz = float(parts[1])
                for nme, v in zip(names, parts[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nme)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def lognormal_k(nc, seed, sigma=1.0):
    rng = np.random.default_rng(seed)
    g = rng.standard_normal((nc, 3)) * sigma
    k = np.zeros((nc, 9))
    k[:, 0], k[:, 4], k[:, 8] = np.exp(g[:, 0]), np.exp(g[:, 1]), np.exp(g[:, 2])
    return k


# ------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference algorithm on host cores
# ------------------------------------------------------------------------------------------
WORKLOAD = "C3: %d^3 hex, log-normal diag K (seed 256), method 1, Dirichlet x-/x+, no-flow else"


def cpu_assembly_sample(n, seed=256):
    """build_lhs + build_rhs of the reference algorithm (oracle port, per-cell Python/NumPy loop
    exactly like mfd.py:545-599) on an n^3 sample of the workload. Returns (seconds, cells, mfd)."""
    from oracle import mimpy_oracle as orc

    k = lognormal_k(n ** 3, seed)
    o = orc.hex_mesh(n, n, n, 1., 1., 1., k_array=k)
    zero = lambda p: np.array([0., 0., 0.])
    t0 = time.perf_counter()
    f = orc.OracleMFD(o, 1)
    f.apply_dirichlet_from_function(0, lambda p: 1.)
    f.apply_dirichlet_from_function(1, lambda p: 0.)
    for b in (2, 3, 4, 5):
        f.apply_neumann_from_function(b, zero)
    f.build_lhs()
    f.build_rhs()
    return time.perf_counter() - t0, n ** 3, f


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n = args.ref_n
    times = []
    for s in range(args.warmup + args.steps):
        dt, cells, f = cpu_assembly_sample(n)
        if s >= args.warmup:
            times.append(dt)
    per = float(np.mean(times))
    value = cells / per / 1e6
    # the reference's solve (SuperLU spsolve, mfd.py:1349-1358) on a smaller sample: 356 s at 30^3
    # (SURVEY 6), super-linear, so n = ref_solve_n keeps the whole arm within a few minutes
    _, _, f = cpu_assembly_sample(args.ref_solve_n)
    t0 = time.perf_counter()
    f.solve()
    t_solve = time.perf_counter() - t0
    sample = "%d^3-cell sample (%d cells) of the 256^3 workload; assembly is O(cells)" % (n, cells)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        # the product arm's workload (same string, same cell count); each step is a bounded sample of it
        "config": {"workload": WORKLOAD % args.n, "cells": args.n ** 3,
                   "sample": "each step assembles a %d^3 block of it (%d cells), throughput is per cell" % (n, cells)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
                         "host_cores_available": os.cpu_count(),
                         "note": "the reference is single-threaded Python (per-cell NumPy/LAPACK calls); 1 core is all it can use",
                         "spsolve_s": t_solve, "spsolve_dof": int(f.flux_dof + f.o.nc),
                         "spsolve_sample": "%d^3 cells" % args.ref_solve_n},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------
# the other BASELINE configs (C2, C4, C5) at their own shapes, one short run each (rank 0, N = 1)
# ------------------------------------------------------------------------------------------
def _timed(ctx, fn, reps=3):
    import torch
    ts, out = [], None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ctx.on_stream():
            e0.record(ctx.stream)
            out = fn()
            e1.record(ctx.stream)
        ctx.sync()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), out


def run_c2(ctx, peak, rtol):
    """BASELINE configs[1]: SPE10-shaped 60x220x85 hexes (6.1 x 3.0 x 0.61 m), k_v = 0.1 k_h, sigma = 2
    log-normal, Dirichlet x-/x+, no-flow elsewhere: the bench path (brick assembly, fused set-up, auxmg PCG)."""
    import torch
    from mimpy_b200 import device

    nx, ny, nz = 60, 220, 85
    dims = (365.76, 670.56, 51.816)
    nc = nx * ny * nz
    g = np.random.default_rng(10).standard_normal((nc, 3))
    k = np.zeros((nc, 9))
    k[:, 0] = k[:, 4] = np.exp(2.0 * g[:, 0])
    k[:, 8] = 0.1 * np.exp(2.0 * g[:, 1])
    dm = device.DeviceMesh.hex(ctx, nx, ny, nz, *dims, cell_k=k)
    fc, areas = device.to_host(ctx, dm.face_centroids), device.to_host(ctx, dm.face_areas)
    eps = 1e-6
    mask = ((fc[:, 1] < eps) | (fc[:, 1] > dims[1] - eps) | (fc[:, 2] < eps) | (fc[:, 2] > dims[2] - eps)).astype(np.uint8)
    plan = device.symbolic(dm, mask)
    xm, xp = np.nonzero(fc[:, 0] < eps)[0], np.nonzero(fc[:, 0] > dims[0] - eps)[0]
    rhs = device.build_rhs(dm, plan, np.concatenate([xm, xp]).astype(np.int32),
                           np.concatenate([-areas[xm], np.zeros(len(xp))]), True, 0., None)
    with ctx.on_stream():
        vals = torch.empty(plan.nnz, dtype=torch.float64, device=ctx.device)
    device.numeric(dm, plan, 1, vals=vals)
    t_asm, _ = _timed(ctx, lambda: device.numeric(dm, plan, 1, vals=vals))
    t_set, hyb = _timed(ctx, lambda: device.HybridSystem(dm, plan, None, precond="auxmg", method=1), reps=2)
    t0 = time.perf_counter()
    sol, info = hyb.solve(rhs, rtol=rtol, maxit=40000, check_every=5, raise_on_noconv=False, norm="saddle", saddle_vals=vals)
    ctx.sync()
    wall = time.perf_counter() - t0
    with ctx.on_stream():
        ax = device.spmv(ctx, plan.indptr, plan.indices, vals, sol, n=plan.ndof)
        res = float(torch.linalg.norm(ax - rhs) / torch.linalg.norm(rhs))
    return {"workload": "C2: 60x220x85 hex (1.12 M cells), k_v = 0.1 k_h, sigma 2 log-normal, method 1", "cells": nc,
            "asm_ms": t_asm, "asm_mcells_per_s": nc / t_asm / 1e3,
            "asm_roofline_frac": ASM_BYTES_PER_CELL * nc / (t_asm / 1e3) / 1e9 / peak,
            "hybrid_setup_ms": t_set, "pcg_iters": int(info.iterations), "pcg_ms_per_iter": float(info.ms_total) / max(int(info.iterations), 1),
            "solve_s": wall, "saddle_rel_residual": res, "rtol": rtol}


def build_c4(ctx, n):
    """The TwoPhase model of BASELINE configs[3] on n^3 hexes (method 6, n^2 rate wells on x-, Dirichlet outflow on x+),
    before initialize_system; returns (model, seconds spent in HexMesh.build_mesh, time step size)."""
    import mimpy_b200.mesh.hexmesh as hexmesh
    import mimpy_b200.models.twophase as twophase

    nc = n ** 3
    rng = np.random.default_rng(128)
    k = np.zeros((nc, 9))
    k[:, 0] = k[:, 4] = k[:, 8] = 1e-13 * np.exp(rng.standard_normal(nc))
    t0 = time.perf_counter()
    mesh = hexmesh.HexMesh()
    mesh.build_mesh(n, n, n, float(n), float(n), float(n), k)
    t_mesh = time.perf_counter() - t0
    zero_flux = lambda p: np.array([0., 0., 0.])
    tp = twophase.TwoPhase()
    tp.set_mesh(mesh, ctx)
    for b in (0, 2, 3, 4, 5):
        tp.apply_flux_boundary_from_function(b, zero_flux)
    tp.apply_pressure_boundary_from_function(1, lambda p: 0.)
    tp.set_initial_p_o(np.zeros(nc))
    tp.set_initial_s_w(np.zeros(nc))
    tp.set_porosities(np.full(nc, .2))
    tp.set_viscosity_water(8.9e-4)
    tp.set_viscosity_oil(1.78e-3)
    tp.set_compressibility_water(0.)
    tp.set_compressibility_oil(0.)
    tp.set_ref_density_water(1000.)
    tp.set_ref_density_oil(1000.)
    tp.set_ref_pressure_oil(0.)
    tp.set_ref_pressure_water(0.)
    tp.set_residual_saturation_water(0.)
    tp.set_residual_saturation_oil(.2)
    tp.set_corey_relperm(2., 2.)
    tp.newton_threshold = 1.e-11      # the reference's ABSOLUTE Newton test is already met at this size
    dt = 1.e4 * (n / 128.) ** 2
    tp.set_time_step_size(dt)
    for c in range(0, nc, n):
        tp.add_rate_well(10. / (n * n), 0., c, "w%d" % c)
    return tp, t_mesh, dt


def run_c4(ctx, peak, n=128, steps=6):
    """BASELINE configs[3]: two-phase IMPES on n^3 hexes through the drop-in TwoPhase API (method 6, n^2 rate
    wells on x-, Dirichlet outflow on x+): ms per IMPES step and the device time of the saturation update."""
    import torch

    nc = n ** 3
    tp, t_mesh, dt = build_c4(ctx, n)
    t0 = time.perf_counter()
    tp.initialize_system()
    ctx.sync()
    t_init = time.perf_counter() - t0
    step_s, sat_ms, iters = [], [], []
    for step in range(1, steps + 1):
        t0 = time.perf_counter()
        tp.update_pressure(step)
        iters.append(int(tp.last_solver_info.iterations))
        if step == 1 or step % 10 == 0:
            tp.find_upwinding_direction()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ctx.on_stream():
            e0.record(ctx.stream)
            tp.update_saturation(step)
            e1.record(ctx.stream)
        ctx.sync()
        torch.cuda.synchronize()
        step_s.append(time.perf_counter() - t0)
        sat_ms.append(e0.elapsed_time(e1))
    tp._sync_host()
    sw = tp.current_s_w
    injected = steps * dt * 10. / 1000. / (0.2 * nc)
    sat_us = float(np.min(sat_ms[1:])) * 1e3
    sat_bytes = SAT_BYTES_PER_CELL * nc
    return {"workload": "C4: two-phase IMPES on %d^3 hex (%.1f M cells), method 6, %d rate wells, drop-in TwoPhase API" % (n, nc / 1e6, n * n),
            "cells": nc, "host_mesh_s": t_mesh, "initialize_system_s": t_init,
            "ms_per_impes_step": float(np.mean(step_s[1:])) * 1e3, "first_step_ms": step_s[0] * 1e3,
            "pcg_iters_per_step": iters, "sat_step_us": sat_us,
            "sat_roofline": {"bound": "hbm", "achieved": sat_bytes / (sat_us / 1e6) / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": sat_bytes / (sat_us / 1e6) / 1e9 / peak, "algorithmic_bytes_per_cell": SAT_BYTES_PER_CELL,
                             "kernels": "k_cell_frac_flow + k_saturation_cells<HEX> + k_rate_wells (twophase.cu): fractional flow once per cell, "
                                        "closed-form face ids, the six faces of a cell through the gather chain together; the "
                                        "reference's per-face formulation (MIMPY_B200_SAT_PERFACE, 378 B/cell) takes 219 us, a "
                                        "single pass with the mobilities per (cell, face) (MIMPY_B200_SAT_FUSED) 309 us",
                             "survey_fused_bytes_per_cell": 80.0},
            "water_mass_balance_rel_err": abs(float(sw.mean()) - injected) / injected,
            "s_w_max": float(sw.max())}


def run_c5(ctx, peak, n=32):
    """BASELINE configs[4] at its timing size (32^3): n^3 hexes with every other cell cut by a plane through its
    centroid (cells of 7..12 faces), full-tensor K, Dirichlet everywhere, through the drop-in MFD API (generic
    warp-group kernels, Jacobi-PCG).  The cuts are host code (Mesh.divide_cells_by_plane with one plane per cell: the
    same halves as the reference's divide_cell_by_plane, tests/test_cutting.py)."""
    import torch
    import mimpy_b200.mesh.hexmesh as hexmesh
    import mimpy_b200.mfd.mfd as mfd
    from mimpy_b200 import device

    nc0 = n ** 3
    rng = np.random.default_rng(5)
    k = np.zeros((nc0, 3, 3))
    for c in range(nc0):
        q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
        k[c] = q @ np.diag(np.exp(rng.standard_normal(3))) @ q.T
    t0 = time.perf_counter()
    mesh = hexmesh.HexMesh()
    mesh.build_mesh(n, n, n, 1., 1., 1., k)
    nrm = np.array([1.2, .2, 1.3])
    nrm /= np.linalg.norm(nrm)
    cut = [c for c in range(nc0) if (c % n + (c // n) % n + c // (n * n)) % 2 == 0]
    mesh.divide_cells_by_plane(cut, np.array([mesh.get_cell_real_centroid(c) for c in cut]), nrm)
    t_mesh = time.perf_counter() - t0
    nc, nf = mesh.get_number_of_cells(), mesh.get_number_of_faces()
    lin = lambda p: 1. + 2. * p[0] - p[1] + .5 * p[2]
    f = mfd.MFD()
    f.set_m_e_construction_method(1)
    f.set_mesh(mesh, ctx)
    for b in range(6):
        f.apply_dirichlet_from_function(b, lin)
    f.build_lhs()
    f.build_rhs()
    dm, plan = f._ensure_plan()
    with ctx.on_stream():
        vals = torch.empty(plan.nnz, dtype=torch.float64, device=ctx.device)
    t_asm, _ = _timed(ctx, lambda: device.numeric(dm, plan, 1, vals=vals))
    sol = f.solve()
    res = float(np.linalg.norm(f.lhs.tocsr() @ sol - f.rhs) / np.linalg.norm(f.rhs))
    ne = np.diff(mesh.cells.compact()[0]).astype(np.float64)
    # generic-kernel count of DESIGN 3: 8 n_E + 104 + 60 r + 8 nnz_M/nc + 16 n_E + 8 per cell
    bytes_cell = 24. * ne.mean() + 112. + 60. * nf / nc + 8. * plan.nnz_m / nc
    return {"workload": "C5: %d^3 hexes, every other cell cut by a plane (%d cells of 6..12 faces, %d faces), full-tensor K, drop-in MFD API" % (n, nc, nf),
            "cells": nc, "host_mesh_and_cuts_s": t_mesh, "asm_ms": t_asm, "asm_mcells_per_s": nc / t_asm / 1e3,
            "asm_roofline": {"bound": "hbm", "achieved": bytes_cell * nc / (t_asm / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": bytes_cell * nc / (t_asm / 1e3) / 1e9 / peak, "algorithmic_bytes_per_cell": bytes_cell,
                             "kernel": "k_local_matrices<G,SCATTER> (generic warp-group kernel); launch-bound at this size"},
            "pcg_iters": int(f.solver_info["iterations"]), "pcg_ms": float(f.solver_info["ms"]), "saddle_rel_residual": res}


def run_generic_path(ctx, peak, n=128):
    """The kernel every polyhedral mesh goes through (k_local_matrices<G,SCATTER>: lane group per cell, any n_E <= 32,
    all eight methods), timed on a mesh large enough to say something about it: the C3 workload at n^3 with the
    fast paths switched off.  Algorithmic bytes: the generic count of DESIGN 3 with n_E = 6."""
    import torch
    from mimpy_b200 import device

    nc = n ** 3
    dm = device.DeviceMesh.hex(ctx, n, n, n, 1., 1., 1., cell_k=lognormal_k(nc, 256))
    plan = device.symbolic(dm, np.zeros(dm.nf, dtype=np.uint8))
    with ctx.on_stream():
        vals = torch.empty(plan.nnz, dtype=torch.float64, device=ctx.device)
    device.numeric(dm, plan, 1, vals=vals, generic_kernel=True)
    t_asm, _ = _timed(ctx, lambda: device.numeric(dm, plan, 1, vals=vals, generic_kernel=True))
    bytes_cell = 24. * 6 + 112. + 60. * dm.nf / nc + 8. * plan.nnz_m / nc
    return {"workload": "generic warp-group kernel on %d^3 hexes (the path of cut / polyhedral meshes), method 1" % n, "cells": nc,
            "asm_ms": t_asm, "asm_mcells_per_s": nc / t_asm / 1e3,
            "asm_roofline": {"bound": "hbm", "achieved": bytes_cell * nc / (t_asm / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": bytes_cell * nc / (t_asm / 1e3) / 1e9 / peak, "algorithmic_bytes_per_cell": bytes_cell,
                             "kernel": "k_local_matrices<8,SCATTER>"}}


def run_extra_configs(ctx, peak, rtol):
    out = {}
    for name, fn in (("c2", lambda: run_c2(ctx, peak, rtol)), ("c4", lambda: run_c4(ctx, peak)), ("c5", lambda: run_c5(ctx, peak)),
                     ("generic_path", lambda: run_generic_path(ctx, peak))):
        try:
            out[name] = fn()
        except Exception as e:  # a failing side config must not take the headline line down
            out[name] = {"error": "%s: %s" % (type(e).__name__, e)}
    return out


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def build_problem(ctx, n, rank, world):
    """This rank's slab of the C3 problem, resident in HBM."""
    import torch
    from mimpy_b200 import device, partition

    slab = partition.SlabPlan(n, n, n, rank, world)
    k_full = lognormal_k(n ** 3, 256)
    k_host = np.ascontiguousarray(k_full[slab.k0 * n * n:slab.k1 * n * n])
    del k_full
    dm = device.DeviceMesh.hex(ctx, n, n, slab.nzl, 1., 1., float(slab.nzl) / n, cell_k=k_host)
    plan = device.symbolic(dm, slab.neumann_mask(), face_flag=slab.face_flag() if world > 1 else None)
    xm, xp = slab.dirichlet_faces()
    areas = device.to_host(ctx, dm.face_areas)
    faces = np.concatenate([xm, xp])
    # value = p(x_f) a_f sigma_bdry with p = 1 on x- (sigma -1), p = 0 on x+   (mfd.py:1157-1162)
    vals = np.concatenate([-areas[xm], np.zeros(len(xp))])
    rhs = device.build_rhs(dm, plan, faces, vals, True, 0., None)
    n_owned = plan.flux_dof
    if world > 1:
        n_owned = partition.attach_halo(ctx, slab, plan)
    return {"slab": slab, "dm": dm, "plan": plan, "rhs": rhs, "k_host": k_host, "n_owned": n_owned}


def run_cuda(args):
    import torch
    import torch.distributed as dist
    from mimpy_b200 import _lib, device, partition

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = _lib.Context(local_rank)
    dev = ctx.device
    n = args.n
    comm = "single"
    if world > 1:
        partition.init_nccl(ctx, rank, world)
        comm = "nccl"
        if os.environ.get("MIMPY_B200_COMM", "peer") != "nccl":
            partition.init_peer(ctx, rank, world, n * n)  # NVLink peer-memory halo + all-reduce
            comm = "nvlink-peer"
    peak, peak_src = measured_peak()

    t_setup = time.perf_counter()
    pb = build_problem(ctx, n, rank, world)
    ctx.sync()
    t_setup = time.perf_counter() - t_setup
    dm, plan, rhs = pb["dm"], pb["plan"], pb["rhs"]
    F = plan.flux_dof
    nc_total = n ** 3
    slab_nz = pb["slab"].nzl
    with ctx.on_stream():
        vals = torch.empty(plan.nnz, dtype=torch.float64, device=dev)
    method = 1

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    check_every = args.check_every if args.check_every > 0 else (5 if args.precond == "auxmg" else 50)

    def owned_norm(v):
        """||v||_2 of a saddle-sized vector over all ranks (flux rows counted by their owner)."""
        with ctx.on_stream():
            s2 = ((v[:pb["n_owned"]] ** 2).sum() + (v[F:] ** 2).sum()).reshape(1).clone()
        ctx.sync()
        if world > 1:
            dist.all_reduce(s2)
        return float(s2) ** 0.5

    # The solve stops when the residual of the ORIGINAL saddle system is <= rtol * ||b||_2: the PCG runs on
    # the face residual against rtol * ||b|| (on this workload it tracks the saddle residual within 10 %,
    # scripts/study_rtol.py), then the recovered [v;p] is checked in the assembled saddle CSR and the PCG
    # continues if the two residuals scale differently (C2).  rtol = 1e-8 leaves |x - x_spsolve|/|x| <= 1e-8
    # at n = 16, 24 (spsolve) and 256 (tests/test_gpu_bench_path.py).
    b_norm = owned_norm(rhs)

    def step(collect):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        barrier()
        with ctx.on_stream():
            evs[0].record(ctx.stream)
            device.numeric(dm, plan, method, vals=vals)
            evs[1].record(ctx.stream)
            hyb = device.HybridSystem(dm, plan, None, precond=args.precond, method=method)  # fused M_E + set-up
            evs[2].record(ctx.stream)
            sol, info = hyb.solve(rhs, rtol=args.rtol, maxit=args.maxit, check_every=check_every,
                                  raise_on_noconv=False, norm="saddle", ref_norm=b_norm, saddle_vals=vals,
                                  owned_norm=owned_norm)
            evs[3].record(ctx.stream)
        ctx.sync()
        barrier()
        t = torch.tensor([evs[0].elapsed_time(evs[1]), evs[1].elapsed_time(evs[2]), evs[2].elapsed_time(evs[3]),
                          float(info.ms_total), evs[0].elapsed_time(evs[3])], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)  # device times, max over ranks
        t = t.tolist()
        out = {"asm_ms": t[0], "setup_ms": t[1], "solve_ms": t[2], "pcg_ms": t[3], "step_ms": t[4],
               "iters": int(info.iterations), "residual": float(info.residual), "rhs_norm": float(info.rhs_norm)}
        out["smoother"] = getattr(hyb, "smoother", None)
        if collect:
            out["sol"] = sol
        del hyb
        return out

    for _ in range(args.warmup):
        step(False)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    res = []
    t0 = time.perf_counter()
    for s in range(args.steps):
        res.append(step(s == args.steps - 1))
    barrier()
    wall = time.perf_counter() - t0
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # ---- the SpMV kernel alone and the plain Jacobi-PCG iteration (bounded: 100 iterations) ----
    with ctx.on_stream():
        hyb = device.HybridSystem(dm, plan, None, precond="jacobi", method=method)
        _, jinfo = hyb.solve(rhs, rtol=1e-30, maxit=100, check_every=50, raise_on_noconv=False)
        _, jinfo = hyb.solve(rhs, rtol=1e-30, maxit=100, check_every=50, raise_on_noconv=False)
        xs = torch.ones(F, dtype=torch.float64, device=dev)
        ys = torch.empty(F, dtype=torch.float64, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        for i in range(reps + 3):
            if i == 3:
                e0.record(ctx.stream)
            if getattr(plan, "sell_npat", 0) > 0:   # the SpMV the PCG runs: column patterns, no index array
                ctx.call("mimpy_b200_spmv_sell_pat", F, _lib.ptr(hyb.sell_ptr), _lib.ptr(hyb.sell_cols), _lib.ptr(plan.sell_pat_id),
                         _lib.ptr(plan.sell_pat_table), plan.sell_npat, _lib.ptr(hyb.a_vals), _lib.ptr(xs), _lib.ptr(ys))
            else:
                ctx.call("mimpy_b200_spmv_sell", F, _lib.ptr(hyb.sell_ptr), _lib.ptr(hyb.sell_cols), _lib.ptr(hyb.a_vals),
                         _lib.ptr(xs), _lib.ptr(ys))
        e1.record(ctx.stream)
    ctx.sync()
    tj = torch.tensor([float(jinfo.ms_total) / max(int(jinfo.iterations), 1), e0.elapsed_time(e1) / reps],
                      dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tj, op=dist.ReduceOp.MAX)
    jacobi_ms_iter, spmv_ms = tj.tolist()
    del hyb, xs, ys

    asm_ms = float(np.mean([r["asm_ms"] for r in res]))
    step_ms = float(np.mean([r["step_ms"] for r in res]))
    iters = res[-1]["iters"]
    pcg_ms = float(np.mean([r["pcg_ms"] for r in res]))
    s_per_iter = pcg_ms / 1e3 / max(iters, 1)
    value = nc_total / (asm_ms / 1e3) / 1e6
    rows = torch.tensor([float(pb["n_owned"]), float(plan.nnz_m)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(rows)
    rows_total, nnz_total = rows.tolist()
    nnz_row = nnz_total / rows_total
    # per-GPU achieved bandwidth of the dominant kernels (algorithmic bytes of the local slab)
    asm_gbs = ASM_BYTES_PER_CELL * (nc_total / world) / (asm_ms / 1e3) / 1e9
    pcg_gbs = pcg_bytes_per_row(nnz_row) * (rows_total / world) / (jacobi_ms_iter / 1e3) / 1e9
    # bytes the SpMV kernel that runs really has to move: with column patterns 8 B per entry + 1 B per row
    # for the pattern id instead of 12 B per entry (the survey's indexed count)
    pat = getattr(plan, "sell_npat", 0) > 0
    spmv_row_bytes = (8.0 * nnz_row + 21.0) if pat else (12.0 * nnz_row + 20.0)
    pcg_row_bytes = spmv_row_bytes + 88.0
    pcg_gbs = pcg_row_bytes * (rows_total / world) / (jacobi_ms_iter / 1e3) / 1e9
    spmv_gbs = spmv_row_bytes * (rows_total / world) / (spmv_ms / 1e3) / 1e9

    aux_roofline = None
    if args.precond == "auxmg":
        try:   # the benched iteration itself (the Jacobi figure above is the survey's 3-kernel iteration)
            blocks = res[-1].get("smoother") == "blocks"
            aux_bytes = auxmg_iteration_bytes(rows_total / world, nnz_total / world, nc_total / world, float(dm.nf),
                                              patterns=pat, blocks=blocks, fp32_cycle=(world == 1))
            aux_gbs = aux_bytes / s_per_iter / 1e9
            aux_roofline = {"bound": "hbm", "achieved": aux_gbs, "peak": peak, "unit": "GB/s", "frac": aux_gbs / peak,
                            "what": "one iteration of the benched PCG (SpMV + 2 vector kernels + restriction%s + V(1,1) cycle on %s "
                                    "levels + prolongation), mean over the solve incl. its convergence checks"
                                    % (" with cell-block solves" if blocks else "", "FP32" if world == 1 else "FP64"),
                            "s_per_iter": s_per_iter, "bytes_per_iteration": aux_bytes,
                            "formula": "8 nnz + 109 rows + (116 + 84 blocks) nc + 55 nf + (64 | 136) nc 8/7 (bench.py:auxmg_iteration_bytes)"}
        except Exception as e:   # a side figure must not take the headline line down
            aux_roofline = {"error": "%s: %s" % (type(e).__name__, e)}

    sol = res[-1]["sol"]
    with ctx.on_stream():
        psum = sol[F:].sum().reshape(1).clone()
        if world > 1:
            dist.all_reduce(psum)
        p_mean = float(psum) / nc_total
        # residual of the ORIGINAL saddle system [M -DIV^T; DIV C][v;p] = b with the assembled CSR; on a
        # slab partition the interface flux rows are completed by the exchange-add of the two halves
        res_vec = device.spmv(ctx, plan.indptr, plan.indices, vals, sol, n=plan.ndof) - rhs
        if world > 1:
            ctx.call("mimpy_b200_halo_exchange_add", _lib.ptr(res_vec))
    saddle_res = owned_norm(res_vec) / b_norm
    sol_p_l2, sol_v_l2 = None, None
    with ctx.on_stream():
        zc = torch.zeros_like(sol)
        zc[F:] = sol[F:]
    sol_p_l2 = owned_norm(zc)
    with ctx.on_stream():
        zc.zero_()
        zc[:F] = sol[:F]
    sol_v_l2 = owned_norm(zc)
    del zc, res_vec

    # ---- e2e: pinned host buffers in and out through the library's host entry points (mimpy_b200/hostpipe.py),
    # every rank its slab.  (a) the headline metric end to end: cell_k (host) -> assembly -> CSR values (host);
    # the assembly runs slab by slab so that the download of one slab's finished rows overlaps the upload
    # of K for the next (PCIe is full duplex; the transfers, not the 2.5 ms kernel, are what this number
    # is made of).  (b) the solve end to end: cell_k (host) -> set-up -> PCG -> [v; p] (host).
    # Each rank times itself (CUDA-synchronised wall clock); the slowest rank counts.  No collective and no
    # barrier inside the timed regions.
    e2e = None
    e2e_solve = None
    if not args.no_e2e:
        from mimpy_b200 import hostpipe
        nc_loc = dm.nc
        k_pin = torch.from_numpy(pb["k_host"]).pin_memory().reshape(nc_loc, 9)
        vals_host = torch.empty(plan.nnz, dtype=torch.float64).pin_memory()
        pipe = hostpipe.AssemblyPipeline(dm, plan, method, n_slabs=8)
        times = []
        for s in range(4):
            barrier()
            t0 = time.perf_counter()
            pipe.run(k_pin, vals_host)
            times.append(time.perf_counter() - t0)
        barrier()
        # what arrived on the host is the matrix of a plain full assembly
        device.numeric(dm, plan, method, vals=vals)
        ctx.sync()
        e2e_ok = bool(torch.equal(vals_host, vals.cpu()))
        tt = torch.tensor([min(times[1:])], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": nc_total / float(tt) / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": int(k_pin.numel() * 8 * world), "d2h_bytes_per_step": int(vals_host.numel() * 8 * world),
               "what": "mimpy_b200.hostpipe.AssemblyPipeline.run, per rank: cell_k from pinned host -> numeric assembly -> CSR "
                       "values to pinned host (wall clock of the slowest rank)" +
                       ("; %d layer slabs, upload / kernel / download pipelined on three streams" % (len(pipe.bounds) - 1)
                        if pipe.piped else ""),
               "host_result_equals_full_assembly": e2e_ok}
        del vals_host, pipe
        if world == 1:
            sol_host = torch.empty(plan.ndof, dtype=torch.float64).pin_memory()
            ts = []
            for s in range(2):
                barrier()
                t0 = time.perf_counter()
                sinfo = hostpipe.solve_to_host(dm, plan, method, k_pin, rhs, sol_host, rtol=args.rtol, precond=args.precond,
                                               saddle_vals=vals, maxit=args.maxit, check_every=check_every,
                                               raise_on_noconv=False)
                ts.append(time.perf_counter() - t0)
            e2e_solve = {"value": nc_total / min(ts) / 1e6, "unit": "Mcells/s (assembled into the face system and solved)",
                         "seconds": min(ts), "pcg_iters": int(sinfo.iterations),
                         "h2d_bytes_per_step": int(k_pin.numel() * 8), "d2h_bytes_per_step": int(sol_host.numel() * 8),
                         "what": "mimpy_b200.hostpipe.solve_to_host: cell_k from pinned host -> fused hybrid set-up -> PCG "
                                 "(residual of the saddle system <= rtol ||b||, verified) -> [v; p] to pinned host",
                         "host_result_equals_device_solve": bool(torch.allclose(sol_host, sol.cpu(), rtol=0, atol=1e-9 * float(sol.abs().max())))}
            del sol_host
        del k_pin

    extra = None
    if not args.no_extras and rank == 0 and world == 1:
        extra = run_extra_configs(ctx, peak, args.rtol)

    cpu = None
    if not args.no_cpu and rank == 0 and world == 1:
        dt, cells, f = cpu_assembly_sample(args.cpu_n)
        cpu = {"value": cells / dt / 1e6, "unit": UNIT, "cores": 1, "kind": "port",
               "host_cores_available": os.cpu_count(),
               "sample": "%d^3 cells of the same workload (build_lhs+build_rhs, %.1f s); O(cells)" % (args.cpu_n, dt)}

    traffic, traffic_src = measured_traffic(n, world)
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD % n,
                       "cells": nc_total, "face_rows": int(rows_total), "face_system_nnz": int(nnz_total),
                       "partition": "z-slabs x%d" % world, "pcg_collectives": comm,
                       "l2": "inputs larger than L2 (>= %.1f GB touched per step per GPU)" % (ASM_BYTES_PER_CELL * nc_total / world / 1e9),
                       "pcg": {"preconditioner": args.precond, "rtol": args.rtol, "check_every": check_every,
                               "stop": "residual of the original saddle system <= rtol * ||b||_2, verified in the assembled CSR "
                                       "(meets |x - x_spsolve|/|x| <= 1e-8: tests/test_gpu_bench_path.py)"}},
            "assembly_ms": asm_ms, "hybrid_setup_ms": float(np.mean([r["setup_ms"] for r in res])),
            "pcg_iters": iters, "pcg_s_per_iter": s_per_iter,
            "solve_time_s": float(np.mean([r["solve_ms"] for r in res])) / 1e3,
            "pcg_rel_residual": res[-1]["residual"] / max(res[-1]["rhs_norm"], 1e-300),
            "saddle_rel_residual": saddle_res, "mean_pressure": p_mean,
            "solution_norms": {"p_l2": sol_p_l2, "v_l2": sol_v_l2,
                               "note": "compare across N: the partitioned solves agree with the 1-GPU solve to the PCG tolerance"},
            "roofline": {"bound": "hbm", "achieved": asm_gbs, "peak": peak, "unit": "GB/s", "frac": asm_gbs / peak,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "kernel": "k_numeric_brick3<1> (pipelined brick: fused M_E + CSR assembly, bulk-copy staging), per GPU",
                         "peak_source": peak_src, "frac_of_8TBs_spec": asm_gbs / 8000.,
                         "algorithmic_bytes_per_cell": ASM_BYTES_PER_CELL},
            "roofline_spmv": {"bound": "hbm", "achieved": spmv_gbs, "peak": peak, "unit": "GB/s", "frac": spmv_gbs / peak,
                              "kernel": "k_spmv_sell_pat (face system, SELL-32, %d column patterns: 1 B/row instead of 4 B/entry; the "
                                        "numerator counts what this kernel moves, 8 nnz + 21 n), per GPU, timed alone" % getattr(plan, "sell_npat", 0),
                              "ms": spmv_ms, "algorithmic_bytes_per_row": spmv_row_bytes,
                              "survey_indexed_bytes_per_row": 12.0 * nnz_row + 20.0},
            "roofline_pcg": {"bound": "hbm", "achieved": pcg_gbs, "peak": peak, "unit": "GB/s", "frac": pcg_gbs / peak,
                             "what": "one Jacobi-PCG iteration (3 kernels), 100-iteration run",
                             "s_per_iter": jacobi_ms_iter / 1e3,
                             "algorithmic_bytes_per_row": pcg_row_bytes, "survey_indexed_bytes_per_row": pcg_bytes_per_row(nnz_row),
                             "rows": int(rows_total)},
            "roofline_pcg_auxmg": aux_roofline,
            "configs": extra, "cpu_baseline": cpu, "e2e": e2e, "e2e_solve": e2e_solve,
            "gpu_launches": int(args.steps * launches_per_step(args.precond, n, slab_nz, iters, world)),
            "clocks": sampler.summary(), "setup_s": t_setup, "wall_s_timed": wall,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--size", dest="n", type=int, default=256, help="cells per axis (256 = BASELINE config)")
    ap.add_argument("--rtol", type=float, default=1e-8)
    ap.add_argument("--maxit", type=int, default=20000)
    ap.add_argument("--check-every", type=int, default=0, help="PCG convergence check interval (0: 5 with auxmg, 50 with jacobi)")
    ap.add_argument("--precond", default="auxmg", choices=["auxmg", "jacobi"])
    ap.add_argument("--cpu-n", type=int, default=32, help="n^3 sample of the cpu_baseline leg (the same sample as --ref-n)")
    ap.add_argument("--ref-n", type=int, default=32)
    ap.add_argument("--ref-solve-n", type=int, default=20)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the short runs of BASELINE configs C2, C4, C5")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_cuda(args)


if __name__ == "__main__":
    sys.exit(main())
