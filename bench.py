#!/usr/bin/env python
"""Benchmark of the electrostatic hot path (BASELINE.json metric: "Poisson solve ms/step & GDoF/s per
CG iter vs HBM roofline; atoms interp/s").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One JSON line on stdout (rank 0).  Layout of the run (DESIGN.md section "Measurement"):

* headline workload "X" = BASELINE.json config 4: the vacuum mesh of apex.ckx + extension_90nm.xyz
  (tests/golden/bench_x90.npz, produced by the reference's own mesher) refined twice -> 2.24e7
  hexahedra / 2.3e7 DoF, matrix 7.8 GB >> L2.  One step = the first PIC step after a re-mesh
  (ProjectRunaway.cpp:449-533): setup(-E0, V0) -> assemble(true) incl. the space-charge RHS of 1e6
  synthetic electrons -> cold-start CG to the reference's absolute tolerance -> check_limits.
  value = DoF x CG iterations / second (GDoF/s per CG iteration), inputs resident in HBM;
  e2e   = the same step through the host-buffer C ABI (particles H2D from pinned memory, potential
          D2H) -- what a Femocs host code would call.
* "native" sub-object = BASELINE.json config 2 on the nanotip_big mesh (tests/golden/mesh_mdbig.npz):
  ms per field step (assemble + solve + extract + interpolate on the 8 937 surface atoms), us per CG
  iteration, atoms/s; its e2e is the full re-meshed MD step from host mesh arrays.
* roofline: algorithmic bytes of the SpMV+dot kernel / its average duration measured live with CUDA
  events on the library's stream (fb_last_solve_profile), against MEASURED_PEAKS.json.
* cpu_baseline: the CPU oracle (restated deal.II SSOR-CG; oracle/) on a bounded sample, rank 0, N=1.

--impl reference times that CPU path alone (rank 0 only).  Nothing here reads /root/reference.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

E0 = -0.5                      # V/A, write_defaults (Main.cpp:25-50)
CG_TOL = 1e-9                  # Config.cpp:64-66 field.cg_tolerance
N_CG = 10000                   # field.n_cg
Q_OVER_EPS0 = -180.9512268     # Pic.h:91
WSP = 0.01                     # Config.cpp:110 electron weight
FALLBACK_HBM_GBS = 6650.0      # /opt/skills/guides/B200_PROFILING.md


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------
# workloads (synthetic, deterministic; no RNG state shared with anything else)
# ------------------------------------------------------------------------------------------------
def load_x_mesh(levels):
    from femocs_b200 import synth
    with np.load(os.path.join(ROOT, "tests", "golden", "bench_x90.npz")) as z:
        nodes, hexs = z["nodes"], z["hexs"]
    for _ in range(levels):
        nodes, hexs = synth.refine_hexes(nodes, hexs)
    return np.ascontiguousarray(nodes), np.ascontiguousarray(hexs), np.ones(len(hexs), np.int32)


def synth_particles(nodes, hexs, n, seed=2024):
    """n electrons at trilinear images of uniform natural coordinates inside randomly chosen
    hexahedra (cell index = hexahedron index: every hexahedron of X is vacuum)."""
    rng = np.random.default_rng(seed)
    cells = rng.integers(0, len(hexs), size=n).astype(np.int32)
    u, v, w = rng.uniform(-0.9, 0.9, size=(3, n))
    su = np.array([-1, 1, 1, -1, -1, 1, 1, -1.0]); sv = np.array([-1, -1, 1, 1, -1, -1, 1, 1.0])
    sw = np.array([-1, -1, -1, -1, 1, 1, 1, 1.0])            # femocs/UCD vertex signs (InterpolatorCells.cpp:1334-1353)
    N = (1 + su[None] * u[:, None]) * (1 + sv[None] * v[:, None]) * (1 + sw[None] * w[:, None]) / 8.0
    xyz = np.einsum("nk,nkd->nd", N, nodes[hexs[cells]])
    return np.ascontiguousarray(xyz), cells


def load_native():
    with np.load(os.path.join(ROOT, "tests", "golden", "mesh_mdbig.npz")) as z:
        return {k: z[k] for k in z.files}


# ------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------
def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def __enter__(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.path or not os.path.exists(self.path):
            return out
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.unlink(self.path)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def oracle_x_sample():
    """CPU baseline: the oracle's restated deal.II path on the X base mesh (level 0, 407 018 DoF):
    setup + assemble(true) + SSOR(1.2)-CG to 1e-9.  Returns (GDoF/s per iteration, dict)."""
    from oracle.oracle import Oracle
    nodes, hexs, mk = load_x_mesh(0)
    o = Oracle()
    o.import_mesh(nodes, hexs, mk)
    return o


def oracle_x_step(o):
    t = time.perf_counter()
    o.setup(-E0, 0.0, False)
    o.assemble(True)
    it = o.solve(N_CG, CG_TOL, 1.2, 0)
    dt = time.perf_counter() - t
    return it, dt


def oracle_x_jacobi(o):
    """the same system with the GPU path's own algorithm (Jacobi-PCG) on the host: the oracle's CG with the OpenMP
    vmult on all cores -- the like-for-like CPU number SURVEY.md section 8d asks for next to the SSOR port"""
    o.setup(-E0, 0.0, False)
    o.assemble(True)
    t = time.perf_counter()
    it = o.solve(N_CG, CG_TOL, 1.2, 1)
    return it, time.perf_counter() - t


def cpu_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------------
# reference arm: the CPU path alone
# ------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    o = oracle_x_sample()
    its, tot = 0, 0.0
    for s in range(args.warmup + args.steps):
        it, dt = oracle_x_step(o)
        if it <= 0:
            raise SystemExit("oracle CG did not converge: %d" % it)
        if s >= args.warmup:
            its += it; tot += dt
    val = o.n_dofs * its / tot / 1e9
    sample = ("X base mesh (refinement level 0: %d DoF, nnz %d) -- setup + assemble + SSOR(1.2)-CG to abs 1e-9 (%d iterations/step); "
              "GDoF/s per iteration is size independent on the CPU once the matrix leaves the caches" % (o.n_dofs, o.nnz, its // args.steps))
    line = {
        "impl": "reference", "metric": "poisson_cg_gdof_per_s", "value": val, "unit": "GDoF/s per CG iteration",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, None),
        "cpu_baseline": {"value": val, "unit": "GDoF/s per CG iteration", "cores": cpu_threads(), "kind": "port", "sample": sample,
                         "note": "SSOR sweeps are serial (as in deal.II); only vmult uses the %d threads" % cpu_threads()},
        "e2e": {"value": val, "unit": "GDoF/s per CG iteration", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, sizes):
    cfg = {"workload": "config 4 'X': apex.ckx+extension_90nm.xyz vacuum mesh refined x%d, Poisson with space-charge RHS of %d "
                       "synthetic electrons, cold-start Jacobi-PCG to abs tol %g (first PIC step after re-mesh)" % (args.levels, args.particles, CG_TOL),
           "refine_levels": args.levels, "n_particles": args.particles, "cg_tolerance": CG_TOL, "n_cg": N_CG, "E0": E0,
           "l2_policy": "inputs larger than L2 (CSR matrix 7.8 GB vs 126 MB L2); native sub-benchmark flushes L2 between steps"}
    if sizes:
        cfg.update(sizes)
    return cfg


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import femocs_b200 as fb

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE line, the JSON: whatever native libraries print there while we run (NCCL announces
    # its version on stdout at communicator creation) is sent to stderr at the file-descriptor level
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- femocs_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    t_setup = time.perf_counter()
    nodes, hexs, mk = load_x_mesh(args.levels)
    pxyz, pcell = synth_particles(nodes, hexs, args.particles)
    log("[rank %d] X mesh: %d hexahedra, %d nodes (%.1f s)" % (rank, len(hexs), len(nodes), time.perf_counter() - t_setup))

    ctx = fb.Context(local)
    if world > 1:
        ctx.init_comm_torch(dist)                 # element-partitioned CG: one rank per GPU, NCCL inside the library
    ctx.set_option("cg_profile", 32)
    if args.dof_order is not None:
        ctx.set_option("dof_order", args.dof_order)
    solver = fb.PoissonSolver(ctx, fb.FieldConfig(E0=E0, cg_tolerance=CG_TOL, n_cg=N_CG, mode="transient"))
    t0 = time.perf_counter()
    assert solver.import_mesh(nodes, hexs, mk), "import_mesh failed"
    part = ctx.partition()
    log("[rank %d] import_mesh: %d rows (+%d ghosts) of %d DoF, local nnz %d (%.1f s host setup + upload)"
        % (rank, part["n_rows"], part["n_ghost"], solver.n_dofs_global, solver.nnz, time.perf_counter() - t0))
    del nodes, hexs, mk
    n, nnz = solver.n_dofs_global, solver.nnz          # global DoF, rank-local non-zeros
    n_loc = part["n_rows"]
    cf = Q_OVER_EPS0 * WSP

    # device-resident inputs (value leg) and pinned host buffers (e2e leg)
    d_pxyz = torch.from_numpy(pxyz).cuda(); d_pcell = torch.from_numpy(pcell).cuda()
    h_pxyz = torch.from_numpy(pxyz).pin_memory(); h_pcell = torch.from_numpy(pcell).pin_memory()
    h_phi = torch.empty(solver.n_vertices, dtype=torch.float64).pin_memory()
    h_phi_np = h_phi.numpy()
    stream = torch.cuda.ExternalStream(ctx.stream)

    def step_dev():
        solver.setup(-E0, 0.0)
        solver.assemble_dev(True, d_pxyz.data_ptr(), d_pcell.data_ptr(), args.particles, cf)
        it = solver.solve()
        solver.check_limits(-1e30, 1e30)
        return it

    def step_e2e():
        solver.setup(-E0, 0.0)
        ctx.check(ctx.L.fb_poisson_assemble(ctx.h, 1, h_pxyz.data_ptr(), h_pcell.data_ptr(), args.particles, cf))
        it = solver.solve()
        solver.check_limits(-1e30, 1e30)
        solver.export_solution(h_phi_np)
        return it

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        launches0 = ctx.kernel_launches
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        iters = 0; its = []
        for _ in range(steps):
            it = fn()
            its.append(it); iters += abs(it)
        e1.record(stream)
        e1.synchronize()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1))
        return ms, iters, its, ctx.kernel_launches - launches0

    with ClockSampler(local) as clk:
        ms, iters, its, launches = timed(step_dev, args.steps, args.warmup)
    clocks = clk.summary()
    spmv_ms, vec_ms, n_samp = solver.solve_profile()
    solve_ms, last_it, _ = solver.solve_stats()
    ms_e2e, iters_e2e, its_e2e, _ = timed(step_e2e, args.steps, max(1, args.warmup - 2))

    # N > 1: ONE system, element-partitioned over the ranks (strong scaling): every rank runs the same iterations
    value = n * iters / (ms * 1e-3) / 1e9
    e2e_value = n * iters_e2e / (ms_e2e * 1e-3) / 1e9
    peak, peak_src = hbm_peak()
    nnz_all = sum_over_ranks(nnz)
    bytes_spmv = 12.0 * nnz + 4.0 * (n_loc + 1) + 16.0 * n_loc          # this rank's share (rank 0 reports)
    bytes_iter = 12.0 * nnz + 4.0 * (n_loc + 1) + 104.0 * n_loc
    achieved = bytes_spmv / (spmv_ms * 1e-3) / 1e9 if spmv_ms > 0 else None
    iter_ms = solve_ms / max(1, last_it)
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("spmv_dram_bytes_per_launch")
        except Exception:
            traffic = None

    line = {
        "metric": "poisson_cg_gdof_per_s", "value": value, "unit": "GDoF/s per CG iteration", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, {"n_dofs": n, "nnz": int(nnz_all), "n_cells": part["n_cells_global"],
                                          "parallelism": ("element-partitioned over %d GPUs (RCB), NCCL p2p halo + all-reduce; rank 0: %d rows, %d ghosts, %d halo values sent"
                                                          % (world, part["n_rows"], part["n_ghost"], part["n_send"])) if world > 1 else "1 GPU",
                                          "cg_iterations_per_step": its, "converged": bool(all(i > 0 for i in its))}),
        "e2e": {"value": e2e_value, "unit": "GDoF/s per CG iteration", "h2d_bytes_per_step": int(pxyz.nbytes + pcell.nbytes),
                "d2h_bytes_per_step": int(h_phi_np.nbytes + 16), "ms_per_step": ms_e2e / args.steps,
                "call": "fb_poisson_setup + fb_poisson_assemble(host particles) + fb_poisson_solve + fb_check_limits + fb_export_solution"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "k_spmv_jds (block-JDS SpMV fused with the d.h dot product)" + (
                         "; rank 0 share, timed INCLUDING the halo exchange and the dot-product all-reduce" if world > 1 else ""),
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                     "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": bytes_spmv, "avg_launch_ms": spmv_ms, "samples": n_samp,
                     "cg_iteration": {"algorithmic_bytes": bytes_iter, "ms": iter_ms,
                                      "achieved": bytes_iter / (iter_ms * 1e-3) / 1e9, "frac": bytes_iter / (iter_ms * 1e-3) / 1e9 / peak,
                                      "vector_kernels_ms": vec_ms, "gdof_per_s": n / (iter_ms * 1e-3) / 1e9}},
    }

    if rank == 0 and world == 1:
        if not args.skip_native:
            del d_pxyz, d_pcell
            line["native"] = native_step(fb, torch, args)
            line["pic"] = pic_step(fb, torch, args)
        if not args.skip_cpu:
            o = oracle_x_sample()
            it, dt = oracle_x_step(o)
            line["cpu_baseline"] = {
                "value": o.n_dofs * abs(it) / dt / 1e9, "unit": "GDoF/s per CG iteration", "cores": cpu_threads(), "kind": "port",
                "sample": "X base mesh (refinement level 0: %d DoF) -- one step: setup + assemble + SSOR(1.2)-CG to abs 1e-9, %d iterations, %.1f s; "
                          "SSOR sweeps serial as in deal.II, vmult on %d threads" % (o.n_dofs, it, dt, cpu_threads())}
            itj, dtj = oracle_x_jacobi(o)
            line["cpu_baseline"]["jacobi_pcg_all_cores"] = {
                "value": o.n_dofs * abs(itj) / dtj / 1e9, "unit": "GDoF/s per CG iteration", "iterations": itj, "solve_s": dtj,
                "cores": cpu_threads(), "note": "same algorithm as the GPU path (Jacobi-PCG, abs 1e-9) on the same base mesh, solve only"}
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    os.close(real_stdout)
    if rank == 0:
        print(json.dumps(line), flush=True)


def native_step(fb, torch, args):
    """BASELINE.json config 2 on the nanotip_big mesh: ms per field step, us per CG iteration, atoms/s;
    e2e = the full re-meshed MD step from host arrays; CPU oracle timed beside it."""
    m = load_native()
    ctx = fb.Context(torch.cuda.current_device())
    conf = fb.FieldConfig(E0=E0, cg_tolerance=CG_TOL, n_cg=N_CG)
    solver = fb.PoissonSolver(ctx, conf)
    assert solver.import_mesh(m["nodes"], m["hexs"], m["hex_markers"])
    interp = fb.Interpolator(ctx); interp.initialize(m)
    atoms = np.ascontiguousarray(m["surf_atoms"]); na = len(atoms)
    d_atoms = torch.from_numpy(atoms).cuda()
    d_cells = torch.empty(na, dtype=torch.int32, device="cuda"); d_sol = torch.empty(na, 5, dtype=torch.float64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.ExternalStream(ctx.stream)
    K, W = max(args.steps, 10), max(args.warmup, 3)

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def step_dev():
        solver.setup(-E0, 0.0); solver.assemble(True)
        it = solver.solve()
        interp.extract_solution(solver, True)
        ctx.check(ctx.L.fb_locate_interpolate_dev(ctx.h, 2, 1, na, d_atoms.data_ptr(), d_cells.data_ptr(), d_sol.data_ptr()))
        ctx.synchronize()
        return it

    def step_e2e():
        s = fb.PoissonSolver(ctx, conf)
        s.import_mesh(m["nodes"], m["hexs"], m["hex_markers"])
        s.setup(-E0, 0.0); s.assemble(True)
        it = s.solve()
        interp.initialize(m); interp.extract_solution(s, True)
        f = fb.FieldReader(interp); f.set_preferences(False, 2, 1); f.interpolate(atoms)
        return it

    def run(fn):
        tot = 0.0; its = 0
        for s in range(W + K):
            flush.fill_(s & 0xff); torch.cuda.synchronize()           # L2 flush between steps
            a, b = ev(), ev()
            a.record(stream); it = fn(); b.record(stream); b.synchronize()
            if s >= W:
                tot += a.elapsed_time(b); its += abs(it)
        return tot / K, its / K

    launches0 = ctx.kernel_launches
    ms_dev, it_dev = run(step_dev)
    launches = (ctx.kernel_launches - launches0) / (W + K)
    solve_ms, it_last, _ = solver.solve_stats()
    # interpolation alone
    a, b = ev(), ev(); reps = 20
    a.record(stream)
    for _ in range(reps):
        ctx.check(ctx.L.fb_locate_interpolate_dev(ctx.h, 2, 1, na, d_atoms.data_ptr(), d_cells.data_ptr(), d_sol.data_ptr()))
    ctx.synchronize(); b.record(stream); b.synchronize()
    interp_ms = a.elapsed_time(b) / reps
    ms_e2e, _ = run(step_e2e)
    out = {"workload": "config 2: nanotip_big mesh (%d DoF, %d hexahedra, nnz %d), Laplace field step + field on %d surface atoms (dim 2, rank 1)"
                       % (solver.n_dofs, solver.n_cells, solver.nnz, na),
           "field_step_ms": ms_dev, "cg_iterations": it_dev, "solve_ms": solve_ms, "us_per_cg_iteration": 1e3 * solve_ms / max(1, it_last),
           "gdof_per_s_per_iteration": solver.n_dofs / (solve_ms * 1e-3 / max(1, it_last)) / 1e9,
           "atoms_interp_per_s": na / (interp_ms * 1e-3), "interp_ms": interp_ms, "gpu_launches_per_step": launches,
           "e2e_remesh_step_ms": ms_e2e,
           "e2e_call": "fb_import_mesh(host mesh) + setup + assemble + solve + fb_interp_initialize + extract + fb_locate_interpolate(host atoms)",
           "regime": "L2-resident (CSR 8.5 MB): latency bound, HBM roofline not applicable"}
    if not args.skip_cpu:
        from oracle.oracle import Oracle
        best = None
        for _ in range(2):
            t = time.perf_counter()
            o = Oracle(); o.import_mesh(m["nodes"], m["hexs"], m["hex_markers"]); o.interp_initialize(m)
            t1 = time.perf_counter()
            o.setup(-E0, 0.0, False); o.assemble(True)
            t2 = time.perf_counter()
            it = o.solve(N_CG, CG_TOL, 1.2, 0)
            t3 = time.perf_counter()
            o.extract_solution(True); o.locate_interpolate(2, 1, atoms)
            t4 = time.perf_counter()
            r = {"remesh_step_ms": 1e3 * (t4 - t), "field_step_ms": 1e3 * (t4 - t1), "solve_ms": 1e3 * (t3 - t2), "cg_iterations": it,
                 "interp_ms": None, "cores": cpu_threads(), "kind": "port (SSOR-CG, deal.II semantics)"}
            t5 = time.perf_counter(); o.locate_interpolate(2, 1, atoms); r["interp_ms"] = 1e3 * (time.perf_counter() - t5)
            r["atoms_interp_per_s"] = na / (r["interp_ms"] * 1e-3)
            if best is None or r["field_step_ms"] < best["field_step_ms"]:
                best = r
        out["cpu_baseline"] = best
    ctx.close()
    return out


def pic_particles(m, n, seed=2024):
    """config 3: n electrons uniform (in natural coordinates) inside the vacuum hexahedra whose centroid lies within
    50 A of the apex, velocities N(0, 0.1 A/fs); returns (pos, vel, solver cell)"""
    rng = np.random.default_rng(seed)
    vac = np.flatnonzero(m["hex_markers"] > 0)
    cent = m["nodes"][m["hexs"][vac]].mean(1)
    apex = m["surf_atoms"][np.argmax(m["surf_atoms"][:, 2])]
    near = np.flatnonzero(np.linalg.norm(cent - apex, axis=1) < 50.0)
    pick = near[rng.integers(0, len(near), size=n)]            # index into the vacuum list = solver cell id
    u, v, w = rng.uniform(-0.9, 0.9, size=(3, n))
    su = np.array([-1, 1, 1, -1, -1, 1, 1, -1.0]); sv = np.array([-1, -1, 1, 1, -1, -1, 1, 1.0]); sw = np.array([-1, -1, -1, -1, 1, 1, 1, 1.0])
    N = (1 + su[None] * u[:, None]) * (1 + sv[None] * v[:, None]) * (1 + sw[None] * w[:, None]) / 8.0
    pos = np.einsum("nk,nkd->nd", N, m["nodes"][m["hexs"][vac[pick]]])
    vel = rng.normal(0.0, 0.1, size=(n, 3))
    return np.ascontiguousarray(pos), np.ascontiguousarray(vel), pick.astype(np.int32)


def pic_step(fb, torch, args):
    """BASELINE.json config 3: PIC on the nanotip_small mesh with 1e6 synthetic electrons resident in HBM.  One step =
    ProjectRunaway::make_pic_step (ProjectRunaway.cpp:492-533) without emission/injection/collisions: update_positions
    (push, periodic images, cell search, clear_lost) -> assemble(space-charge RHS) -> warm-started CG -> check_limits ->
    extract_solution -> update_velocities."""
    with np.load(os.path.join(ROOT, "tests", "golden", "mesh_mdsmall.npz")) as z:
        m = {k: z[k] for k in z.files}
    n_p = args.particles
    pos, vel, cells = pic_particles(m, n_p)
    ctx = fb.Context(torch.cuda.current_device())
    solver = fb.PoissonSolver(ctx, fb.FieldConfig(E0=E0, cg_tolerance=CG_TOL, n_cg=N_CG, mode="transient"))
    assert solver.import_mesh(m["nodes"], m["hexs"], m["hex_markers"])
    interp = fb.Interpolator(ctx); interp.initialize(m)
    lo = m["nodes"].min(0); hi = m["nodes"].max(0)
    box = np.array([lo[0], hi[0], lo[1], hi[1], lo[2], hi[2]])
    dt, q_over_m, cf = 0.5, -17.5882, Q_OVER_EPS0 * WSP * 1e-3      # Pic.h:90; weight scaled so that 1e6 SPs stay within V limits
    stream = torch.cuda.ExternalStream(ctx.stream)
    K, W = max(args.steps, 5), max(args.warmup, 3)
    import ctypes as C

    def fresh():
        return torch.from_numpy(pos).cuda(), torch.from_numpy(vel).cuda(), torch.from_numpy(cells).cuda()

    d_pos, d_vel, d_cell = fresh()
    state = {"n": n_p}
    solver.setup(-E0, 0.0); solver.assemble(True); solver.solve()           # Laplace start, as solve_laplace before the PIC loop

    def step():
        lost = C.c_long(0)
        n = state["n"]
        ctx.check(ctx.L.fb_pic_update_positions_dev(ctx.h, n, d_pos.data_ptr(), d_vel.data_ptr(), d_cell.data_ptr(), dt,
                                                    box.ctypes.data, 1, C.byref(lost)))
        n -= lost.value; state["n"] = n
        solver.assemble_dev(False, d_pos.data_ptr(), d_cell.data_ptr(), n, cf)
        it = solver.solve()
        solver.check_limits(-1e30, 1e30)
        interp.extract_solution(solver, True)
        ctx.check(ctx.L.fb_pic_update_velocities_dev(ctx.h, n, d_pos.data_ptr(), d_cell.data_ptr(), d_vel.data_ptr(), dt, q_over_m))
        ctx.synchronize()
        return it

    for _ in range(W):
        step()
    launches0 = ctx.kernel_launches
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    its = [step() for _ in range(K)]
    b.record(stream); b.synchronize()
    ms = a.elapsed_time(b) / K
    launches = (ctx.kernel_launches - launches0) / K
    n_alive = state["n"]
    # the two particle passes alone
    lost = C.c_long(0)
    a.record(stream)
    ctx.check(ctx.L.fb_pic_update_positions_dev(ctx.h, n_alive, d_pos.data_ptr(), d_vel.data_ptr(), d_cell.data_ptr(), dt, box.ctypes.data, 1, C.byref(lost)))
    b.record(stream); b.synchronize(); push_ms = a.elapsed_time(b); n_alive -= lost.value
    a.record(stream)
    ctx.check(ctx.L.fb_pic_update_velocities_dev(ctx.h, n_alive, d_pos.data_ptr(), d_cell.data_ptr(), d_vel.data_ptr(), dt, q_over_m))
    ctx.synchronize(); b.record(stream); b.synchronize(); vel_ms = a.elapsed_time(b)
    out = {"workload": "config 3: nanotip_small mesh (%d DoF), %d synthetic electrons resident in HBM, dt %.2f fs, periodic box; step = update_positions "
                       "(push + cell search + clear_lost) + space-charge assemble + warm-started CG + check_limits + extract_solution + update_velocities"
                       % (solver.n_dofs, n_p, dt),
           "ms_per_step": ms, "particles_alive": int(n_alive), "cg_iterations_per_step": its, "gpu_launches_per_step": launches,
           "update_positions_ms": push_ms, "update_positions_particles_per_s": n_alive / (push_ms * 1e-3),
           "update_velocities_ms": vel_ms, "update_velocities_particles_per_s": n_alive / (vel_ms * 1e-3),
           "particle_steps_per_s": n_alive / (ms * 1e-3)}
    if not args.skip_cpu:
        from oracle import pic as opic
        from oracle.oracle import Oracle
        o = Oracle(); o.import_mesh(m["nodes"], m["hexs"], m["hex_markers"]); o.interp_initialize(m)
        o.setup(-E0, 0.0, False); o.assemble(True); o.solve(N_CG, CG_TOL, 1.2, 0); o.extract_solution(True)
        ns = min(n_p, 50000)
        t = time.perf_counter()
        p1, v1, c1, _ = opic.update_positions(o, pos[:ns], vel[:ns], cells[:ns], dt, box, True)
        t1 = time.perf_counter()
        opic.update_velocities(o, p1, v1, c1, dt, q_over_m)
        t2 = time.perf_counter()
        out["cpu_baseline"] = {"kind": "port", "cores": 1, "sample": "%d particles of the same set" % ns,
                               "update_positions_particles_per_s": ns / (t1 - t), "update_velocities_particles_per_s": len(c1) / (t2 - t1)}
    ctx.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--levels", type=int, default=2, help="uniform refinements of the X base mesh (2 = 2.3e7 DoF)")
    ap.add_argument("--particles", type=int, default=1000000)
    ap.add_argument("--dof-order", type=int, default=None)
    ap.add_argument("--skip-native", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
