#!/usr/bin/env python
"""Benchmark of the electrostatic hot path (BASELINE.json metric: "Poisson solve ms/step & GDoF/s per
CG iter vs HBM roofline; atoms interp/s").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One JSON line on stdout (rank 0).  Layout of the run (DESIGN.md section "Measurement"):

* headline workload "X" = BASELINE.json config 4: the vacuum mesh of apex.ckx + extension_90nm.xyz
  (tests/golden/bench_x90.npz, produced by the reference's own mesher) refined twice -> 2.24e7
  hexahedra / 2.3e7 DoF, matrix 7.8 GB >> L2.  One step = a cold field solve on a freshly imported mesh
  (ProjectRunaway.cpp:449-533): setup(-E0, V0) -> assemble(true) incl. the space-charge RHS of 1e6
  synthetic electrons -> cold-start CG to the reference's absolute tolerance -> check_limits.
  value = DoF x CG iterations / second (GDoF/s per CG iteration), inputs resident in HBM;
  e2e   = the same step through the host-buffer C ABI (particles H2D from pinned memory, potential
          D2H) -- what a Femocs host code would call.  fb_import_mesh (once per re-mesh) is timed
          separately ("import_mesh_s").
* "time_to_solution" = the SAME system on the X mesh refined ONCE (2.97e6 DoF, matrix 0.93 GB), which the CPU
  can finish: seconds and iterations per cold solve for this arm; both arms of bench.py run it, the B200 arm
  also checks its potential against the CPU oracle's (the pair the time-to-solution ratio comes from).
* "native" sub-object = BASELINE.json config 2 on the nanotip_big mesh (tests/golden/mesh_mdbig.npz):
  ms per field step (assemble + solve + extract + interpolate on the 8 937 surface atoms), us per CG
  iteration, atoms/s (surface atoms, dim 2; ALL atoms, dim 3 = femocs_interpolate_elfield); its e2e is the full
  re-meshed MD step from host mesh arrays.
* "pic" = config 3: PIC on nanotip_small with 1e6 synthetic electrons, population kept steady by re-seeding.
* every leg checks the result of the code path it times against the CPU oracle ONCE before timing ("verified").
* roofline: algorithmic bytes of the SpMV+dot kernel / its average duration measured live with CUDA
  events on the library's stream (fb_last_solve_profile), against MEASURED_PEAKS.json.
* cpu_baseline: the CPU oracle (restated deal.II SSOR-CG; oracle/) on a bounded sample, rank 0, N=1.

--impl reference times that CPU path alone (rank 0 only).  Nothing here reads /root/reference.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

E0 = -0.5                      # V/A, write_defaults (Main.cpp:25-50)
CG_TOL = 1e-9                  # Config.cpp:64-66 field.cg_tolerance
N_CG = 10000                   # field.n_cg
Q_OVER_EPS0 = -180.9512268     # Pic.h:91
Q_OVER_M = -17.5882            # Pic.h:90
WSP = 0.01                     # Config.cpp:110 electron weight
FALLBACK_HBM_GBS = 6650.0      # /opt/skills/guides/B200_PROFILING.md
TTS_LEVEL = 1                  # refinement level of the time-to-solution pair (2.97e6 DoF)
REF_SAMPLE_ITERS = 8           # CG iterations per step of the reference arm's bounded sample


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


def cpu_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def use_all_host_threads():
    """The CPU legs use every core the process may run on, whatever the launcher exported (torchrun sets
    OMP_NUM_THREADS=1): the count is set explicitly in the OpenMP runtime the oracle is linked against."""
    n = cpu_threads()
    os.environ["OMP_NUM_THREADS"] = str(n)
    from oracle.oracle import lib
    lib().fo_set_num_threads(n)
    return n


def rel_diff(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300))


def workload_config(args):
    """identical in both arms (the driver compares it)"""
    return {"workload": "config 4 'X': apex.ckx+extension_90nm.xyz vacuum mesh refined x%d (2.3e7 DoF at x2), Poisson with the space-charge RHS of %d "
                        "synthetic electrons, cold field solve: setup + assemble(true) + CG to abs tol %g + check_limits; fb_import_mesh "
                        "(once per re-mesh) timed separately" % (args.levels, args.particles, CG_TOL),
            "refine_levels": args.levels, "n_particles": args.particles, "cg_tolerance": CG_TOL, "n_cg": N_CG, "E0": E0,
            "time_to_solution_workload": "the same system on the X mesh refined x%d (2.97e6 DoF): complete cold solves by both arms" % TTS_LEVEL,
            "reference_arm_step": "bounded sample: %d SSOR(1.2)-CG iterations (deal.II semantics, all host threads) on the x%d mesh per step"
                                  % (REF_SAMPLE_ITERS, TTS_LEVEL),
            "l2_policy": "inputs larger than L2 (CSR matrix 7.8 GB vs 126 MB L2); native / pic sub-benchmarks flush L2 between steps"}


# ------------------------------------------------------------------------------------------------
# CPU legs (oracle port of the reference's deal.II path)
# ------------------------------------------------------------------------------------------------
class CpuX:
    """the oracle on the X mesh at a refinement level; one object per level, mesh imported once"""

    def __init__(self, level, particles):
        from oracle.oracle import Oracle
        self.threads = use_all_host_threads()
        t = time.perf_counter()
        nodes, hexs, mk = load_x_mesh(level)
        self.pxyz, self.pcell = synth_particles(nodes, hexs, particles)
        self.o = Oracle()
        self.o.import_mesh(nodes, hexs, mk)
        self.level = level
        self.setup_s = time.perf_counter() - t
        self.cf = Q_OVER_EPS0 * WSP

    def assemble(self):
        self.o.setup(-E0, 0.0, False)
        self.o.assemble(True, self.pxyz, self.pcell, self.cf)

    def cold_solve(self, precond=0, max_iter=N_CG):
        """setup + assemble(true, particles) + CG: (signed iterations, seconds, seconds of the CG alone)"""
        t = time.perf_counter()
        self.assemble()
        t1 = time.perf_counter()
        it = self.o.solve(max_iter, CG_TOL, 1.2, precond)
        t2 = time.perf_counter()
        return it, t2 - t, t2 - t1


def cpu_time_to_solution(cx):
    it, tot, cg = cx.cold_solve(0)
    if it <= 0:
        raise SystemExit("oracle CG did not converge: %d" % it)
    return {"mesh": "X refined x%d" % cx.level, "n_dofs": cx.o.n_dofs, "nnz": cx.o.nnz, "seconds": tot, "cg_seconds": cg, "iterations": it,
            "preconditioner": "SSOR(1.2) (deal.II PreconditionSSOR, the reference's solver)", "threads": cx.threads,
            "note": "setup + assemble(true) incl. space charge + CG to abs %g; SSOR sweeps serial as in deal.II, vmult on all threads" % CG_TOL}


# ------------------------------------------------------------------------------------------------
# reference arm: the CPU path alone
# ------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cx = CpuX(TTS_LEVEL, args.particles)
    log("[reference] X x%d: %d DoF, nnz %d, %d threads (%.1f s mesh + import)" % (TTS_LEVEL, cx.o.n_dofs, cx.o.nnz, cx.threads, cx.setup_s))
    tts = cpu_time_to_solution(cx)
    log("[reference] time to solution: %.1f s, %d iterations" % (tts["seconds"], tts["iterations"]))
    its, tot = 0, 0.0
    for s in range(args.warmup + args.steps):
        t = time.perf_counter()
        cx.o.set_solution(np.zeros(cx.o.n_dofs))              # cold start (every Dirichlet value of this workload is 0 V)
        it = cx.o.solve(REF_SAMPLE_ITERS, CG_TOL, 1.2, 0)     # bounded: stops after REF_SAMPLE_ITERS iterations (-#it)
        dt = time.perf_counter() - t
        if s >= args.warmup:
            its += abs(it); tot += dt
    val = cx.o.n_dofs * its / tot / 1e9
    sample = ("X refined x%d (%d DoF, nnz %d; the x%d system of the B200 arm needs ~20 min per CPU solve): per step a cold start + "
              "%d SSOR(1.2)-CG iterations on the assembled system; a complete solve (setup + assemble + CG) is under time_to_solution"
              % (TTS_LEVEL, cx.o.n_dofs, cx.o.nnz, args.levels, REF_SAMPLE_ITERS))
    line = {
        "impl": "reference", "metric": "poisson_cg_gdof_per_s", "value": val, "unit": "GDoF/s per CG iteration",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": val, "unit": "GDoF/s per CG iteration", "cores": cx.threads, "kind": "port", "sample": sample,
                         "note": "SSOR sweeps are serial (as in deal.II); only vmult uses the %d threads" % cx.threads},
        "e2e": {"value": val, "unit": "GDoF/s per CG iteration", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "time_to_solution": tts,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def host_residual(solver):
    """|K phi - b|_2 of the system the library holds, recomputed on the host with scipy (independent of every kernel
    that produced phi), and its normwise backward error eta = |r| / (|A|_inf |x|_2 + |b|_2).  CG monitors the RECURRENCE
    residual (as deal.II's SolverCG does); the true one drifts from it by O(eps |A| |x|), so eta of a correct FP64
    solve sits at or below the unit round-off while a wrong potential gives eta ~ 1e-3 or worse."""
    import scipy.sparse as sp
    g = solver.get_system()
    A = sp.csr_matrix((g["val"], g["col"], g["rowptr"]))
    r = float(np.linalg.norm(A @ g["sol"] - g["rhs"]))
    eta = r / (float(abs(A).sum(1).max()) * float(np.linalg.norm(g["sol"])) + float(np.linalg.norm(g["rhs"])))
    return r, eta


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
    if args.cg_p2p is not None:
        ctx.set_option("cg_p2p", args.cg_p2p)
    solver = fb.PoissonSolver(ctx, fb.FieldConfig(E0=E0, cg_tolerance=CG_TOL, n_cg=N_CG, mode="transient"))
    t0 = time.perf_counter()
    assert solver.import_mesh(nodes, hexs, mk), "import_mesh failed"
    import_s = max_over_ranks(time.perf_counter() - t0)
    part = ctx.partition()
    log("[rank %d] import_mesh: %d rows (+%d ghosts) of %d DoF, local nnz %d (%.1f s host setup + upload)"
        % (rank, part["n_rows"], part["n_ghost"], solver.n_dofs_global, solver.nnz, import_s))
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

    # ---- check what is about to be timed: one step of each leg, both must give the same potential and the host must
    #      find its true residual at the tolerance (world 1; partitioned runs compare the two legs and the iterations)
    verified = {}
    it_dev = step_dev(); phi_dev = solver.export_solution().copy()
    it_e2e = step_e2e()
    verified["dev_vs_e2e_phi_rel"] = rel_diff(phi_dev, h_phi_np)
    verified["iterations"] = [it_dev, it_e2e]
    assert it_dev > 0 and it_e2e > 0, "CG did not converge: %d / %d" % (it_dev, it_e2e)
    assert verified["dev_vs_e2e_phi_rel"] < 1e-8, verified
    if world == 1 and not args.skip_verify:
        r, eta = host_residual(solver)
        verified["host_residual_l2"] = r; verified["backward_error"] = eta
        assert eta < 1e-14, "backward error %g of the potential (true residual %g)" % (eta, r)
        log("[verify] X: |K phi - b| = %.3g recomputed on the host, normwise backward error %.2g" % (r, eta))
    del phi_dev

    with ClockSampler(local) as clk:
        ms, iters, its, launches = timed(step_dev, args.steps, max(0, args.warmup - 1))     # the verification step was warm-up step 1
    clocks = clk.summary()
    spmv_ms, vec_ms, n_samp = solver.solve_profile()
    kernel_id = solver.solve_kernel()
    solve_ms, last_it, _ = solver.solve_stats()
    ms_e2e, iters_e2e, its_e2e, _ = timed(step_e2e, args.steps, max(1, args.warmup - 2))

    # the same cold step with the two-level preconditioner (one GPU, or partitioned in the peer-mapped mode): time to
    # solution is what a FEMOCS user sees, the per-iteration throughput above is what the roofline rates
    two_level = None
    if world == 1 or ctx.comm_mode == 2:
        phi_j = solver.export_solution().copy()
        solver.conf.precond = fb.PRECOND_TWOLEVEL
        ctx.set_option("cg_profile", 0)
        t0 = time.perf_counter(); it_first = step_dev(); first_s = time.perf_counter() - t0      # includes the one-off set-up
        tl_rel = rel_diff(solver.export_solution(), phi_j)
        assert it_first > 0 and tl_rel < 1e-6, (it_first, tl_rel)      # both stop at |r| <= 1e-9; the 1e-8 bar is tested over-converged
        ms_tl, _, its_tl, launches_tl = timed(step_dev, args.steps, 1)
        two_level = {"preconditioner": "Jacobi + aggregation coarse-grid correction (FB_PRECOND_TWOLEVEL)", "seconds_per_step": ms_tl / args.steps * 1e-3,
                     "cg_iterations_per_step": its_tl, "first_step_s_with_setup": first_s, "phi_rel_diff_vs_jacobi_pcg": tl_rel,
                     "jacobi_seconds_per_step": ms / args.steps * 1e-3, "jacobi_iterations": its[-1],
                     "speedup_vs_jacobi_pcg": ms / ms_tl, "gpu_launches": int(launches_tl)}
        solver.conf.precond = fb.PRECOND_JACOBI
        del phi_j
        log("[two-level] X: %.3f s per cold step, %s iterations (Jacobi-PCG %.3f s, %d)" % (two_level["seconds_per_step"], its_tl, ms / args.steps * 1e-3, its[-1]))

    # N > 1: ONE system, element-partitioned over the ranks (strong scaling): every rank runs the same iterations
    value = n * iters / (ms * 1e-3) / 1e9
    e2e_value = n * iters_e2e / (ms_e2e * 1e-3) / 1e9
    peak, peak_src = hbm_peak()
    nnz_all = sum_over_ranks(nnz)
    bytes_spmv = 12.0 * nnz + 4.0 * (n_loc + 1) + 16.0 * n_loc          # this rank's share (rank 0 reports)
    bytes_iter = 12.0 * nnz + 4.0 * (n_loc + 1) + 104.0 * n_loc
    achieved = bytes_spmv / (spmv_ms * 1e-3) / 1e9 if spmv_ms > 0 else None
    iter_ms = solve_ms / max(1, last_it)
    traffic = None; traffic_src = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp) and world == 1 and args.levels == 2:
        try:
            tj = json.load(open(tp))
            if tj.get("spmv_kernel", 304) == kernel_id:
                traffic = tj.get("spmv_dram_bytes_per_launch")
                traffic_src = "static, not measured in this run: " + tj.get("source", tp)
        except Exception:
            traffic = None

    line = {
        "metric": "poisson_cg_gdof_per_s", "value": value, "unit": "GDoF/s per CG iteration", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "sizes": {"n_dofs": n, "nnz": int(nnz_all), "n_cells": part["n_cells_global"],
                  "parallelism": ("element-partitioned over %d GPUs (RCB), %s; rank 0: %d rows, %d ghosts, %d halo values sent"
                                  % (world, {2: "peer-mapped iteration: halo values and dot-product sums stored into the peers' memory over NVLink by the kernels, "
                                                "6 kernels per iteration (pack, SpMV, all-reduce, update, all-reduce, direction) in a CUDA graph, no NCCL inside the loop",
                                             1: "NCCL inside the iteration (grouped send/recv halo + 2 all-reduces, host-issued)"}.get(ctx.comm_mode, "?"),
                                     part["n_rows"], part["n_ghost"], part["n_send"])) if world > 1 else "1 GPU",
                  "comm_mode": ctx.comm_mode,
                  "cg_iterations_per_step": its, "converged": bool(all(i > 0 for i in its)), "seconds_per_step": ms / args.steps * 1e-3,
                  "preconditioner": "Jacobi", "import_mesh_s": import_s},
        "verified": verified,
        "time_to_solution_two_level": two_level,
        "e2e": {"value": e2e_value, "unit": "GDoF/s per CG iteration", "h2d_bytes_per_step": int(pxyz.nbytes + pcell.nbytes),
                "d2h_bytes_per_step": int(h_phi_np.nbytes + 16), "ms_per_step": ms_e2e / args.steps,
                "call": "fb_poisson_setup + fb_poisson_assemble(host particles) + fb_poisson_solve + fb_check_limits + fb_export_solution"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": {306: "k_spmv_jdss (segmented block-JDS SpMV: rows longer than 32 entries split into chained segments; fused with the d.h dot product)",
                                                 304: "k_spmv_jds (block-JDS SpMV fused with the d.h dot product)"}.get(kernel_id, "SpMV variant %d" % kernel_id) + (
                         "; rank 0 share, timed INCLUDING the halo exchange and the dot-product all-reduce" if world > 1 else ""),
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": bytes_spmv, "avg_launch_ms": spmv_ms, "samples": n_samp,
                     "cg_iteration": {"algorithmic_bytes": bytes_iter, "ms": iter_ms,
                                      "achieved": bytes_iter / (iter_ms * 1e-3) / 1e9, "frac": bytes_iter / (iter_ms * 1e-3) / 1e9 / peak,
                                      "vector_kernels_ms": vec_ms, "gdof_per_s": n / (iter_ms * 1e-3) / 1e9}},
    }
    ctx.close()
    del d_pxyz, d_pcell

    if rank == 0 and world == 1:
        cx = None
        if not args.skip_cpu:
            cx = CpuX(TTS_LEVEL, args.particles)
            tts_cpu = cpu_time_to_solution(cx)
            log("[cpu] X x%d: %.1f s, %d SSOR-CG iterations on %d threads" % (TTS_LEVEL, tts_cpu["seconds"], tts_cpu["iterations"], cx.threads))
            line["cpu_baseline"] = {
                "value": cx.o.n_dofs * tts_cpu["iterations"] / tts_cpu["seconds"] / 1e9, "unit": "GDoF/s per CG iteration", "cores": cx.threads,
                "kind": "port",
                "sample": "X refined x%d (%d DoF, nnz %d): ONE complete cold solve -- setup + assemble(true) + SSOR(1.2)-CG to abs %g, %d iterations, "
                          "%.1f s; SSOR sweeps serial as in deal.II, vmult on %d threads"
                          % (TTS_LEVEL, cx.o.n_dofs, cx.o.nnz, CG_TOL, tts_cpu["iterations"], tts_cpu["seconds"], cx.threads),
                "time_to_solution": tts_cpu}
        line["time_to_solution"] = tts_leg(fb, torch, args, cx)
        if cx is not None:
            tts = line["time_to_solution"]
            tts["cpu_seconds"] = tts_cpu["seconds"]; tts["cpu_iterations"] = tts_cpu["iterations"]; tts["cpu_threads"] = cx.threads
            tts["speedup_vs_cpu"] = tts_cpu["seconds"] / tts["seconds"]
            tts["two_level_speedup_vs_cpu"] = tts_cpu["seconds"] / tts["two_level_seconds"]
            phi_cpu = cx.o.export_solution()
            cx.o.set_solution(np.zeros(cx.o.n_dofs))
            t = time.perf_counter(); itj = cx.o.solve(100, CG_TOL, 1.2, 1); cgj = time.perf_counter() - t
            line["cpu_baseline"]["jacobi_pcg_all_cores"] = {
                "value": cx.o.n_dofs * abs(itj) / cgj / 1e9, "unit": "GDoF/s per CG iteration", "iterations_sampled": abs(itj), "seconds": cgj,
                "cores": cx.threads, "note": "the GPU path's own algorithm (Jacobi-PCG) on the same x%d system: 100 iterations, CG only" % TTS_LEVEL}
            del cx
        if not args.skip_native:
            line["native"] = native_step(fb, torch, args)
            line["pic"] = pic_step(fb, torch, args)
            line["heat"] = heat_step(fb, torch, args)
    if dist is not None:
        dist.destroy_process_group()
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    os.close(real_stdout)
    if rank == 0:
        print(json.dumps(line), flush=True)


def tts_leg(fb, torch, args, cx):
    """time to solution on the X mesh refined TTS_LEVEL times: the complete cold solve the CPU arm also runs; when the
    CPU result is at hand (cx) the potentials are compared"""
    nodes, hexs, mk = load_x_mesh(TTS_LEVEL)
    pxyz, pcell = synth_particles(nodes, hexs, args.particles)
    ctx = fb.Context(torch.cuda.current_device())
    solver = fb.PoissonSolver(ctx, fb.FieldConfig(E0=E0, cg_tolerance=CG_TOL, n_cg=N_CG, mode="transient"))
    t0 = time.perf_counter()
    assert solver.import_mesh(nodes, hexs, mk)
    import_s = time.perf_counter() - t0
    h_pxyz = torch.from_numpy(pxyz).pin_memory(); h_pcell = torch.from_numpy(pcell).pin_memory()
    h_phi = torch.empty(solver.n_vertices, dtype=torch.float64).pin_memory().numpy()
    cf = Q_OVER_EPS0 * WSP
    stream = torch.cuda.ExternalStream(ctx.stream)

    def step():
        solver.setup(-E0, 0.0)
        ctx.check(ctx.L.fb_poisson_assemble(ctx.h, 1, h_pxyz.data_ptr(), h_pcell.data_ptr(), args.particles, cf))
        it = solver.solve()
        solver.check_limits(-1e30, 1e30)
        solver.export_solution(h_phi)
        return it

    out = {"mesh": "X refined x%d" % TTS_LEVEL, "n_dofs": solver.n_dofs, "nnz": solver.nnz, "import_mesh_s": import_s}
    it = step()
    assert it > 0
    if cx is not None:
        out["phi_rel_diff_vs_cpu_oracle"] = rel_diff(h_phi, cx.o.export_solution())
        # both sides stop at |r|_2 <= 1e-9 with different preconditioners: the potentials (~1e3 V) agree to the
        # conditioning of K times that residual; the 1e-8 parity bar itself is tested with both sides over-converged
        assert out["phi_rel_diff_vs_cpu_oracle"] < 1e-6, out
        log("[verify] X x%d: potential vs CPU oracle (SSOR-CG): rel %.2e" % (TTS_LEVEL, out["phi_rel_diff_vs_cpu_oracle"]))
    K = max(3, min(args.steps, 5))
    for _ in range(2):
        step()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    its = [step() for _ in range(K)]
    b.record(stream); b.synchronize()
    sec = a.elapsed_time(b) * 1e-3 / K
    out.update(seconds=sec, iterations=its[-1], preconditioner="Jacobi",
               call="fb_poisson_setup + fb_poisson_assemble(host particles) + fb_poisson_solve + fb_check_limits + fb_export_solution",
               gdof_per_s_per_iteration=solver.n_dofs * abs(its[-1]) / sec / 1e9)
    # the same complete cold solves with the two-level preconditioner
    phi_j = h_phi.copy()
    solver.conf.precond = fb.PRECOND_TWOLEVEL
    assert step() > 0
    out["two_level_phi_rel_diff_vs_jacobi"] = rel_diff(h_phi, phi_j)
    assert out["two_level_phi_rel_diff_vs_jacobi"] < 1e-6, out
    step()
    a.record(stream)
    its2 = [step() for _ in range(K)]
    b.record(stream); b.synchronize()
    out.update(two_level_seconds=a.elapsed_time(b) * 1e-3 / K, two_level_iterations=its2[-1])
    ctx.close()
    return out


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
    all_atoms = np.ascontiguousarray(m["atoms"]); nall = len(all_atoms)
    d_atoms = torch.from_numpy(atoms).cuda()
    d_cells = torch.empty(na, dtype=torch.int32, device="cuda"); d_sol = torch.empty(na, 5, dtype=torch.float64, device="cuda")
    d_all = torch.from_numpy(all_atoms).cuda()
    d_cells_all = torch.empty(nall, dtype=torch.int32, device="cuda"); d_sol_all = torch.empty(nall, 5, dtype=torch.float64, device="cuda")
    h_all = torch.from_numpy(all_atoms).pin_memory().numpy()
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

    moved = m["nodes"].copy()
    lo_, hi_ = moved.min(0), moved.max(0)
    inner = np.all((moved > lo_ + 1e-9) & (moved < hi_ - 1e-9), axis=1)
    moved[inner] += 0.02 * np.random.default_rng(11).standard_normal((int(inner.sum()), 3))
    m_moved = dict(m, nodes=moved)
    flip = [0]

    def step_e2e_moved():
        """a host code that moves nodes WITHOUT re-meshing: same connectivity, new coordinates; the import
        keeps numbering / sparsity / SpMV tables (fb_last_import_reused)"""
        flip[0] ^= 1
        mm = m_moved if flip[0] else m
        solver.import_mesh(mm["nodes"], mm["hexs"], mm["hex_markers"])
        assert ctx.last_import_reused
        solver.setup(-E0, 0.0); solver.assemble(True)
        it = solver.solve()
        interp.initialize(mm); interp.extract_solution(solver, True)
        f = fb.FieldReader(interp); f.set_preferences(False, 2, 1); f.interpolate(atoms)
        return it

    def run(fn, st=None):
        st = st or stream
        tot = 0.0; its = 0
        for s in range(W + K):
            flush.fill_(s & 0xff); torch.cuda.synchronize()           # L2 flush between steps
            a, b = ev(), ev()
            a.record(st); it = fn(); b.record(st); b.synchronize()
            if s >= W:
                tot += a.elapsed_time(b); its += abs(it)
        return tot / K, its / K

    # ---- check the timed path once against the CPU oracle: cells bit-exact, potential / field on atoms 1e-8 with both
    #      sides over-converged (1e-11), all atoms (dim 3) included
    verified = None
    o = None
    if not args.skip_cpu:
        from oracle.oracle import Oracle
        use_all_host_threads()
        o = Oracle(); o.import_mesh(m["nodes"], m["hexs"], m["hex_markers"]); o.interp_initialize(m)
        o.setup(-E0, 0.0, False); o.assemble(True); assert o.solve(N_CG, 1e-11, 1.2, 0) > 0
        nod = o.extract_solution(True)
        oc, osol = o.locate_interpolate(2, 1, atoms)
        oc3, osol3 = o.locate_interpolate(3, 1, all_atoms)
        conf.cg_tolerance = 1e-11
        assert step_dev() > 0
        conf.cg_tolerance = CG_TOL
        ctx.check(ctx.L.fb_locate_interpolate_dev(ctx.h, 3, 1, nall, d_all.data_ptr(), d_cells_all.data_ptr(), d_sol_all.data_ptr()))
        ctx.synchronize()
        verified = {"phi_rel": rel_diff(solver.export_solution(), o.export_solution()), "nodal_rel": rel_diff(interp.get_solutions(), nod),
                    "surface_cells_equal": bool(np.array_equal(d_cells.cpu().numpy(), oc)), "surface_field_rel": rel_diff(d_sol.cpu().numpy(), osol),
                    "all_atom_cells_equal": bool(np.array_equal(d_cells_all.cpu().numpy(), oc3)),
                    "all_atom_field_rel": rel_diff(d_sol_all.cpu().numpy(), osol3)}
        assert verified["surface_cells_equal"] and verified["all_atom_cells_equal"], verified
        assert max(verified["phi_rel"], verified["nodal_rel"], verified["surface_field_rel"], verified["all_atom_field_rel"]) < 1e-8, verified
        log("[verify] native: %s" % verified)

    launches0 = ctx.kernel_launches
    ms_dev, it_dev = run(step_dev)
    launches = (ctx.kernel_launches - launches0) / (W + K)
    solve_ms, it_last, _ = solver.solve_stats()
    # interpolation alone: surface atoms (dim 2, rank 1)
    a, b = ev(), ev(); reps = 20
    a.record(stream)
    for _ in range(reps):
        ctx.check(ctx.L.fb_locate_interpolate_dev(ctx.h, 2, 1, na, d_atoms.data_ptr(), d_cells.data_ptr(), d_sol.data_ptr()))
    ctx.synchronize(); b.record(stream); b.synchronize()
    interp_ms = a.elapsed_time(b) / reps
    # all atoms (dim 3, rank 1) = femocs_interpolate_elfield (Femocs.cpp:216-232): device resident, and through the host ABI
    reps3 = 5
    a.record(stream)
    for _ in range(reps3):
        ctx.check(ctx.L.fb_locate_interpolate_dev(ctx.h, 3, 1, nall, d_all.data_ptr(), d_cells_all.data_ptr(), d_sol_all.data_ptr()))
    ctx.synchronize(); b.record(stream); b.synchronize()
    all_ms = a.elapsed_time(b) / reps3
    h_cells = np.zeros(nall, np.int32); h_sol = np.zeros((nall, 5))
    a.record(stream)
    for _ in range(reps3):
        ctx.check(ctx.L.fb_locate_interpolate(ctx.h, 3, 1, nall, h_all.ctypes.data, h_all.ctypes.data + 8, h_all.ctypes.data + 16, 3,
                                              h_cells.ctypes.data, h_sol.ctypes.data))
    b.record(stream); b.synchronize()
    all_e2e_ms = a.elapsed_time(b) / reps3
    ms_e2e, _ = run(step_e2e)
    solver.import_mesh(m["nodes"], m["hexs"], m["hex_markers"])
    ms_e2e_moved, _ = run(step_e2e_moved)
    out = {"workload": "config 2: nanotip_big mesh (%d DoF, %d hexahedra, nnz %d), Laplace field step + field on %d surface atoms (dim 2, rank 1); "
                       "femocs_interpolate_elfield on all %d atoms (dim 3, rank 1)" % (solver.n_dofs, solver.n_cells, solver.nnz, na, nall),
           "verified": verified,
           "field_step_ms": ms_dev, "cg_iterations": it_dev, "solve_ms": solve_ms, "us_per_cg_iteration": 1e3 * solve_ms / max(1, it_last),
           "gdof_per_s_per_iteration": solver.n_dofs / (solve_ms * 1e-3 / max(1, it_last)) / 1e9,
           "atoms_interp_per_s": na / (interp_ms * 1e-3), "interp_ms": interp_ms, "gpu_launches_per_step": launches,
           "all_atoms_interp_per_s": nall / (all_ms * 1e-3), "all_atoms_interp_ms": all_ms,
           "all_atoms_interp_e2e_per_s": nall / (all_e2e_ms * 1e-3), "all_atoms_interp_e2e_ms": all_e2e_ms,
           "e2e_remesh_step_ms": ms_e2e, "e2e_moved_nodes_step_ms": ms_e2e_moved,
           "e2e_moved_nodes_call": "fb_import_mesh(same connectivity, moved nodes: topology tables re-used) + setup + assemble + solve + "
                                   "fb_interp_initialize + extract + fb_locate_interpolate(host atoms)",
           "e2e_call": "fb_import_mesh(host mesh) + setup + assemble + solve + fb_interp_initialize + extract + fb_locate_interpolate(host atoms)",
           "regime": "L2-resident (CSR 8.5 MB): latency bound, HBM roofline not applicable"}
    if o is not None:
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
            t6 = time.perf_counter(); o.locate_interpolate(3, 1, all_atoms); r["all_atoms_interp_ms"] = 1e3 * (time.perf_counter() - t6)
            r["all_atoms_interp_per_s"] = nall / (r["all_atoms_interp_ms"] * 1e-3)
            if best is None or r["field_step_ms"] < best["field_step_ms"]:
                best = r
        out["cpu_baseline"] = best
        ref = ref_interp_baseline(m, atoms, all_atoms)
        if ref:
            out["cpu_baseline_reference_code"] = ref
    ctx.close()
    out["q2"] = native_q2_step(fb, torch, args, m, atoms, run)
    return out


def native_q2_step(fb, torch, args, m, atoms, run):
    """BASELINE.json config 2 as it is worded ("nanotip_big Q2 Laplace solve + surface-atom field interpolation"): the same
    field step with the quadratic element (option fe_degree = 2, csrc/q2.cu), checked against the oracle's FE_Q(2)
    restatement and timed beside it"""
    ctx = fb.Context(torch.cuda.current_device())
    ctx.set_option("fe_degree", 2)
    conf = fb.FieldConfig(E0=E0, cg_tolerance=CG_TOL, n_cg=N_CG)
    solver = fb.PoissonSolver(ctx, conf)
    t0 = time.perf_counter()
    assert solver.import_mesh(m["nodes"], m["hexs"], m["hex_markers"])
    import_ms = 1e3 * (time.perf_counter() - t0)
    interp = fb.Interpolator(ctx); interp.initialize(m)
    na = len(atoms)
    d_atoms = torch.from_numpy(atoms).cuda()
    d_cells = torch.empty(na, dtype=torch.int32, device="cuda"); d_sol = torch.empty(na, 5, dtype=torch.float64, device="cuda")

    def step_dev():
        solver.setup(-E0, 0.0); solver.assemble(True)
        it = solver.solve()
        interp.extract_solution(solver, True)
        ctx.check(ctx.L.fb_locate_interpolate_dev(ctx.h, 2, 1, na, d_atoms.data_ptr(), d_cells.data_ptr(), d_sol.data_ptr()))
        ctx.synchronize()
        return it

    out = {"workload": "config 2 with FE_Q(2): nanotip_big mesh (%d DoF, %d hexahedra, nnz %d), Laplace field step (27-node stiffness assembly + "
                       "CG + extract_solution) + field on %d surface atoms" % (solver.n_dofs, solver.n_cells, solver.nnz, na),
           "import_mesh_ms": import_ms}
    o = None
    if not args.skip_cpu:
        from oracle.oracle import Oracle
        use_all_host_threads()
        t = time.perf_counter()
        o = Oracle(); o.set_fe_degree(2); o.import_mesh(m["nodes"], m["hexs"], m["hex_markers"]); o.interp_initialize(m)
        t1 = time.perf_counter()
        o.setup(-E0, 0.0, False); o.assemble(True)
        t2 = time.perf_counter()
        it_cpu = o.solve(N_CG, CG_TOL, 1.2, 0)
        t3 = time.perf_counter()
        o.extract_solution(True); o.locate_interpolate(2, 1, atoms)
        t4 = time.perf_counter()
        out["cpu_baseline"] = {"field_step_ms": 1e3 * (t4 - t1), "assemble_ms": 1e3 * (t2 - t1), "solve_ms": 1e3 * (t3 - t2), "cg_iterations": it_cpu,
                               "cores": cpu_threads(), "kind": "port (FE_Q(2) restatement, SSOR-CG; the reference fixes shape_degree = 1 at compile time)"}
        assert o.solve(N_CG, 1e-11, 1.2, 0) >= 0
        nod = o.extract_solution(True)
        oc, osol = o.locate_interpolate(2, 1, atoms)
        conf.cg_tolerance = 1e-11
        assert step_dev() > 0
        conf.cg_tolerance = CG_TOL
        v = {"phi_rel": rel_diff(solver.export_solution(), o.export_solution()), "nodal_rel": rel_diff(interp.get_solutions(), nod),
             "surface_cells_equal": bool(np.array_equal(d_cells.cpu().numpy(), oc)), "surface_field_rel": rel_diff(d_sol.cpu().numpy(), osol)}
        assert v["surface_cells_equal"] and max(v["phi_rel"], v["nodal_rel"], v["surface_field_rel"]) < 1e-8, v
        out["verified"] = v
        log("[verify] native Q2: %s" % v)
    st = torch.cuda.ExternalStream(ctx.stream)
    ms_dev, it_dev = run(step_dev, st)
    solve_ms, it_last, _ = solver.solve_stats()
    out.update(field_step_ms=ms_dev, cg_iterations=it_dev, solve_ms=solve_ms, us_per_cg_iteration=1e3 * solve_ms / max(1, it_last),
               gdof_per_s_per_iteration=solver.n_dofs / (solve_ms * 1e-3 / max(1, it_last)) / 1e9, spmv_kernel=ctx.L.fb_last_solve_kernel(ctx.h),
               preconditioner="Jacobi")
    solver.conf.precond = fb.PRECOND_TWOLEVEL
    step_dev()
    ms_tl, it_tl = run(step_dev, st)
    out.update(two_level_field_step_ms=ms_tl, two_level_cg_iterations=it_tl)
    ctx.close()
    return out


def ref_interp_baseline(m, atoms, all_atoms):
    """the interpolation loops timed with the reference's OWN compiled code (oracle/_ref/libfemocs_ref.so, built from
    /root/reference by oracle/Makefile.ref) when that library travelled with the repo; None otherwise"""
    try:
        from oracle import reflib
        if not os.path.exists(reflib.LIB_PATH) or not hasattr(reflib, "time_interpolation"):
            return None
        return reflib.time_interpolation(m, atoms, all_atoms)
    except Exception as e:          # noqa: BLE001  (the reference build is optional test infrastructure)
        log("[cpu] compiled-reference interpolation baseline unavailable: %s" % e)
        return None


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


class PicLeg:
    """BASELINE.json config 3: PIC on the nanotip_small mesh with n synthetic electrons resident in HBM.  One step =
    ProjectRunaway::make_pic_step (ProjectRunaway.cpp:492-533) without emission physics / collisions: update_positions
    (push, periodic images, cell search, clear_lost) -> assemble(space-charge RHS) -> warm-started CG -> check_limits ->
    extract_solution -> update_velocities -> injection.  Injection (Pic::inject_electrons in the reference) is replaced by
    re-seeding: as many electrons as were lost re-enter from the initial set (positions near the apex, thermal
    velocities), so the population stays at n -- without it every electron has left the box after ~20 steps.
    Only _dev entry points are called; tests/test_gpu_dev_entry_points.py runs this very class against the oracle."""

    def __init__(self, fb, torch, m, n_particles, device, dt=0.5, weight_scale=1e-3):
        self.fb, self.torch, self.m, self.n0 = fb, torch, m, int(n_particles)
        self.pos0, self.vel0, self.cells0 = pic_particles(m, self.n0)
        self.ctx = fb.Context(device)
        self.solver = fb.PoissonSolver(self.ctx, fb.FieldConfig(E0=E0, cg_tolerance=CG_TOL, n_cg=N_CG, mode="transient"))
        assert self.solver.import_mesh(m["nodes"], m["hexs"], m["hex_markers"])
        self.interp = fb.Interpolator(self.ctx); self.interp.initialize(m)
        lo = m["nodes"].min(0); hi = m["nodes"].max(0)
        self.box = np.array([lo[0], hi[0], lo[1], hi[1], lo[2], hi[2]])
        # weight scaled so that 1e6 super-particles stay within the potential limits
        self.dt, self.q_over_m, self.cf = float(dt), Q_OVER_M, Q_OVER_EPS0 * WSP * weight_scale
        self.stream = torch.cuda.ExternalStream(self.ctx.stream)
        self.n = 0; self.n_lost_last = 0; self.n_lost_total = 0; self.cursor = 0

    def start(self):
        t = self.torch
        self.d_pos0 = t.from_numpy(self.pos0).cuda(); self.d_vel0 = t.from_numpy(self.vel0).cuda(); self.d_cell0 = t.from_numpy(self.cells0).cuda()
        self.d_pos = self.d_pos0.clone(); self.d_vel = self.d_vel0.clone(); self.d_cell = self.d_cell0.clone()
        self.n = self.n0; self.cursor = 0
        s = self.solver
        s.setup(-E0, 0.0); s.assemble(True)                   # Laplace start, as solve_laplace before the PIC loop
        assert s.solve() > 0
        self.interp.extract_solution(s, True)

    def step(self, reseed=True):
        c, s = self.ctx, self.solver
        lost = C.c_long(0)
        c.check(c.L.fb_pic_update_positions_dev(c.h, self.n, self.d_pos.data_ptr(), self.d_vel.data_ptr(), self.d_cell.data_ptr(), self.dt,
                                                self.box.ctypes.data, 1, C.byref(lost)))
        self.n -= lost.value; self.n_lost_last = lost.value; self.n_lost_total += lost.value
        s.assemble_dev(False, self.d_pos.data_ptr(), self.d_cell.data_ptr(), self.n, self.cf)
        it = s.solve()
        s.check_limits(-1e30, 1e30)
        self.interp.extract_solution(s, True)
        c.check(c.L.fb_pic_update_velocities_dev(c.h, self.n, self.d_pos.data_ptr(), self.d_cell.data_ptr(), self.d_vel.data_ptr(), self.dt, self.q_over_m))
        c.synchronize()
        if reseed and self.n < self.n0:
            k = self.n0 - self.n
            with self.torch.cuda.stream(self.stream):
                idx = (self.torch.arange(k, device="cuda") + self.cursor) % self.n0
                self.d_pos[self.n:self.n0] = self.d_pos0[idx]; self.d_vel[self.n:self.n0] = self.d_vel0[idx]; self.d_cell[self.n:self.n0] = self.d_cell0[idx]
            self.cursor = (self.cursor + k) % self.n0
            self.n = self.n0
            c.synchronize()
        return it

    def state(self):
        n = self.n
        return self.d_pos[:n].cpu().numpy(), self.d_vel[:n].cpu().numpy(), self.d_cell[:n].cpu().numpy()

    def set_velocities(self, vel):
        self.d_vel[:self.n] = self.torch.from_numpy(np.ascontiguousarray(vel)).cuda()
        self.torch.cuda.synchronize()

    def close(self):
        self.ctx.close()


def pic_step(fb, torch, args):
    with np.load(os.path.join(ROOT, "tests", "golden", "mesh_mdsmall.npz")) as z:
        m = {k: z[k] for k in z.files}
    n_p = args.particles
    leg = PicLeg(fb, torch, m, n_p, torch.cuda.current_device())
    ctx, stream = leg.ctx, leg.stream
    K, W = max(args.steps, 5), max(args.warmup, 3)
    leg.start()

    # ---- check the first step of the timed sequence against the CPU oracle on ALL particles: survivors, cells and
    #      positions bit-exact, potential and new velocities to 1e-8 (both sides over-converged for this step)
    verified = None
    o = None
    if not args.skip_cpu:
        from oracle import pic as opic
        from oracle.oracle import Oracle
        use_all_host_threads()
        o = Oracle(); o.import_mesh(m["nodes"], m["hexs"], m["hex_markers"]); o.interp_initialize(m)
        o.setup(-E0, 0.0, False); o.assemble(True); o.solve(N_CG, 1e-11, 1.2, 0); o.extract_solution(True)
        leg.solver.conf.cg_tolerance = 1e-11
        leg.start()
        p1, v1, c1, lost = opic.update_positions(o, leg.pos0, leg.vel0, leg.cells0, leg.dt, leg.box, True)
        o.assemble(False, p1, c1, leg.cf); o.solve(N_CG, 1e-11, 1.2, 0); o.extract_solution(True)
        v2 = opic.update_velocities(o, p1, v1, c1, leg.dt, leg.q_over_m)
        leg.step(reseed=False)
        gp, gv, gc = leg.state()
        verified = {"n_lost_equal": bool(leg.n_lost_last == lost), "cells_equal": bool(np.array_equal(gc, c1)), "positions_equal": bool(np.array_equal(gp, p1)),
                    "phi_rel": rel_diff(leg.solver.export_solution(), o.export_solution()), "velocity_rel": rel_diff(gv, v2), "particles_checked": int(len(c1))}
        assert verified["n_lost_equal"] and verified["cells_equal"] and verified["positions_equal"], verified
        assert verified["phi_rel"] < 1e-8 and verified["velocity_rel"] < 1e-8, verified
        log("[verify] pic: %s" % verified)
        leg.solver.conf.cg_tolerance = CG_TOL
        leg.start()

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(W):
        leg.step()
    launches0 = ctx.kernel_launches
    lost0 = leg.n_lost_total
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0; its = []
    for s in range(K):
        flush.fill_(s & 0xff); torch.cuda.synchronize()
        a.record(stream)
        its.append(leg.step())
        b.record(stream); b.synchronize()
        tot += a.elapsed_time(b)
    ms = tot / K
    launches = (ctx.kernel_launches - launches0) / K
    n_alive = leg.n
    lost_per_step = (leg.n_lost_total - lost0) / K
    assert n_alive >= 0.9 * n_p, "PIC population collapsed: %d of %d alive" % (n_alive, n_p)
    # the two particle passes alone (on the steady population)
    lost = C.c_long(0)
    a.record(stream)
    ctx.check(ctx.L.fb_pic_update_positions_dev(ctx.h, leg.n, leg.d_pos.data_ptr(), leg.d_vel.data_ptr(), leg.d_cell.data_ptr(), leg.dt,
                                                leg.box.ctypes.data, 1, C.byref(lost)))
    b.record(stream); b.synchronize(); push_ms = a.elapsed_time(b); n_pushed = leg.n; leg.n -= lost.value
    a.record(stream)
    ctx.check(ctx.L.fb_pic_update_velocities_dev(ctx.h, leg.n, leg.d_pos.data_ptr(), leg.d_cell.data_ptr(), leg.d_vel.data_ptr(), leg.dt, leg.q_over_m))
    ctx.synchronize(); b.record(stream); b.synchronize(); vel_ms = a.elapsed_time(b)
    out = {"workload": "config 3: nanotip_small mesh (%d DoF), %d synthetic electrons resident in HBM, dt %.2f fs, periodic box; step = update_positions "
                       "(push + cell search + clear_lost) + space-charge assemble + warm-started CG + check_limits + extract_solution + update_velocities "
                       "+ re-seeding of the lost electrons (stands in for Pic::inject_electrons)" % (leg.solver.n_dofs, n_p, leg.dt),
           "verified": verified,
           "ms_per_step": ms, "particles_alive": int(n_alive), "particles_lost_and_reseeded_per_step": lost_per_step,
           "cg_iterations_per_step": its, "gpu_launches_per_step": launches,
           "update_positions_ms": push_ms, "update_positions_particles_per_s": n_pushed / (push_ms * 1e-3),
           "update_velocities_ms": vel_ms, "update_velocities_particles_per_s": leg.n / (vel_ms * 1e-3),
           "particle_steps_per_s": n_alive / (ms * 1e-3)}
    if o is not None:
        from oracle import pic as opic
        ns = min(n_p, 50000)
        t = time.perf_counter()
        p1, v1, c1, _ = opic.update_positions(o, leg.pos0[:ns], leg.vel0[:ns], leg.cells0[:ns], leg.dt, leg.box, True)
        t1 = time.perf_counter()
        opic.update_velocities(o, p1, v1, c1, leg.dt, leg.q_over_m)
        t2 = time.perf_counter()
        out["cpu_baseline"] = {"kind": "port", "cores": 1, "sample": "%d particles of the same set" % ns,
                               "update_positions_particles_per_s": ns / (t1 - t), "update_velocities_particles_per_s": len(c1) / (t2 - t1)}
    leg.close()
    return out


# ------------------------------------------------------------------------------------------------
# config 5: coupled field + current + heat loop on the tip110 mesh
# ------------------------------------------------------------------------------------------------
# rows of PhysicalQuantities::hc_resistivity_data stand-in (T [K], rho); the host code owns the real table
HEAT_TAB_T = np.array([200., 250., 273.15, 300., 350., 400., 450., 500., 600., 800., 1000., 1200., 1357.])
HEAT_TAB_RHO = np.array([10.49, 13.87, 15.43, 17.23, 20.58, 23.95, 27.34, 30.76, 37.72, 52.6, 69.1, 88.2, 104.3])
HEAT_T_AMBIENT = 300.0
HEAT_DT = 4e-11          # [s] heat time step (ProjectRunaway.cpp:549: delta_time * 1e-15)
HEAT_F_REF = 4.0         # [V/Ang] field at which the synthetic emission law gives its nominal current density
HEAT_RAMP = (1.0, 1.04, 0.97, 1.07, 0.94)      # applied-field factor of consecutive steps (a voltage ripple: every step has new emission data)


def synthetic_emission(F):
    """stand-in for EmissionReader::calc_emission (GETELEC, host, out of scope): a Fowler-Nordheim shaped current
    density [A/Ang^2] and Nottingham heat [W/Ang^2] per surface face from the local field norm F [V/Ang]"""
    Fm = HEAT_F_REF
    J = 1.2e-3 * (F / Fm) ** 2 * np.exp(-6.0 * (Fm / np.maximum(F, 1e-3 * Fm) - 1.0))
    return np.ascontiguousarray(J), np.ascontiguousarray(-4e-6 * J)


def heat_step(fb, torch, args):
    """BASELINE.json config 5: ProjectRunaway::run with field_mode = laplace and heat_mode = transient on the tip110 mesh
    (src/ProjectRunaway.cpp:422-447, 535-571): field solve + nodal fields, field on the centroids of the copper_surface
    faces, (synthetic) emission on the host, current solve, heat solve, limits, export of T and current density."""
    with np.load(os.path.join(ROOT, "tests", "golden", "mesh_tip110.npz")) as z:
        m = {k: z[k] for k in z.files}
    dev = torch.cuda.current_device()
    vctx, bctx = fb.Context(dev), fb.Context(dev)
    conf = fb.FieldConfig(E0=E0, cg_tolerance=CG_TOL, n_cg=N_CG)
    hconf = fb.HeatingConfig(cg_tolerance=CG_TOL, n_cg=N_CG)
    solver = fb.PoissonSolver(vctx, conf)
    assert solver.import_mesh(m["nodes"], m["hexs"], m["hex_markers"])
    interp = fb.Interpolator(vctx); interp.initialize(m)
    ch = fb.CurrentHeatSolver(bctx, hconf)
    assert ch.import_mesh(m["nodes"], m["hexs"], m["hex_markers"])
    ch.set_dependencies(HEAT_TAB_T, HEAT_TAB_RHO)
    cen = ch.export_surface_centroids(); nf = len(cen)
    fields = fb.FieldReader(interp); fields.set_preferences(False, 2, 3)
    fields.interpolate(cen)                                          # cells of the face centroids: once per mesh (ProjectRunaway.cpp:273)
    K, W = max(args.steps, 10), max(args.warmup, 3)
    state = {}

    def restart():
        ch.setup(HEAT_T_AMBIENT); state["k"] = 0

    def step():
        amp = HEAT_RAMP[state["k"] % len(HEAT_RAMP)]; state["k"] += 1
        solver.setup(-E0 * amp, 0.0); solver.assemble(True)
        fit = solver.solve(); t_f = solver.solve_stats()[0]
        interp.extract_solution(solver, True)
        fields.calc_interpolation()                                  # surface_fields.calc_interpolation with the cached cells (:576)
        F = np.sqrt((fields.interpolation[:, :3] ** 2).sum(1))
        J, nott = synthetic_emission(F)
        ch.current.set_bcs(J); ch.current.assemble(); cit = ch.current.solve(); t_c = fb.PoissonSolver.solve_stats(ch)[0]
        ch.heat.set_bcs(nott); ch.heat.assemble(HEAT_DT); hit = ch.heat.solve(); t_h = fb.PoissonSolver.solve_stats(ch)[0]
        state["solve_ms"] = state.get("solve_ms", np.zeros(3)) + np.array([t_f, t_c, t_h])
        bad = ch.heat.check_limits(hconf.T_min, hconf.T_max)
        state.update(F=F, J=J, nott=nott, T=ch.heat.export_solution(), rho=ch.current.export_solution_grad(), bad=bad)
        return fit, cit, hit

    verified = None
    o = ob = None
    if not args.skip_cpu:
        from oracle.oracle import Oracle
        use_all_host_threads()
        o = Oracle(); o.import_mesh(m["nodes"], m["hexs"], m["hex_markers"]); o.interp_initialize(m)
        ob = Oracle(); ob.import_bulk_mesh(m["nodes"], m["hexs"], m["hex_markers"]); ob.ch_set_physics(HEAT_TAB_T, HEAT_TAB_RHO)
        assert np.array_equal(ob.surface_centroids(), cen)
        # two coupled steps with both sides over-converged: potential, temperatures and current potential to 1e-8
        conf.cg_tolerance = hconf.cg_tolerance = 1e-14
        restart(); ob.ch_setup(HEAT_T_AMBIENT)
        v2d = ob.vectors()[2]
        for k in range(2):
            fit, cit, hit = step()
            assert fit > 0 and cit > 0 and hit >= 0 and not state["bad"], (fit, cit, hit)
            o.setup(-E0 * HEAT_RAMP[k], 0.0, False); o.assemble(True); assert o.solve(N_CG, 1e-14, 1.2, 0) > 0
            o.extract_solution(True)
            ocells, osol = o.locate_interpolate(2, 3, cen)
            Fo = np.sqrt((osol[:, :3] ** 2).sum(1))
            Jo, no = synthetic_emission(Fo)
            ob.current_assemble(Jo); assert ob.ch_solve(0, N_CG, 1e-14, 1.2, 0) > 0
            ob.heat_assemble(HEAT_DT, no); assert ob.ch_solve(1, N_CG, 1e-14, 1.2, 0) >= 0
        ob.ch_select(0)
        verified = {"surface_field_rel": rel_diff(state["F"], Fo), "temperature_rel": rel_diff(state["T"], ob.ch_solution(1)[v2d]),
                    "current_density_rel": rel_diff(state["rho"], ob.export_solution_grad()), "T_max_K": float(state["T"].max()),
                    "steps_checked": 2}
        assert max(verified["surface_field_rel"], verified["temperature_rel"], verified["current_density_rel"]) < 1e-8, verified
        assert verified["T_max_K"] > HEAT_T_AMBIENT + 1.0, verified     # the workload does heat the tip
        log("[verify] heat: %s" % verified)
        conf.cg_tolerance = hconf.cg_tolerance = CG_TOL

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    stream_v, stream_b = torch.cuda.ExternalStream(vctx.stream), torch.cuda.ExternalStream(bctx.stream)
    restart()
    for _ in range(W):
        step()
    launches0 = vctx.kernel_launches + bctx.kernel_launches
    tot = 0.0; its = []
    state["solve_ms"] = np.zeros(3)
    for s_ in range(K):
        flush.fill_(s_ & 0xff); torch.cuda.synchronize()
        t0 = time.perf_counter()
        its.append(step())
        torch.cuda.synchronize()
        tot += time.perf_counter() - t0                              # every call of the step is synchronous at return (host ABI)
    ms = 1e3 * tot / K
    launches = (vctx.kernel_launches + bctx.kernel_launches - launches0) / K
    # the three solves alone, from the library's own CUDA events (average over the timed steps)
    sm = state["solve_ms"] / K
    parts = {"field_solve_ms": float(sm[0]), "current_solve_ms": float(sm[1]), "heat_solve_ms": float(sm[2])}
    out = {"workload": "config 5: tip110 mesh (vacuum %d DoF / %d hexahedra, bulk %d DoF / %d hexahedra, %d emitting faces); coupled step = Laplace "
                       "field solve + extract_solution + field on the face centroids (dim 2, rank 3) + synthetic emission on the host + "
                       "current assemble/solve + heat assemble(dt = %g s)/solve + check_limits + export of T and current density"
                       % (solver.n_dofs, solver.n_cells, ch.n_dofs, ch.n_cells, nf, HEAT_DT),
           "verified": verified, "ms_per_step": ms, "timing": "host wall clock around synchronous host-ABI calls (inputs and outputs on the host)",
           "cg_iterations_field_current_heat": [list(map(int, i)) for i in its[-3:]], "gpu_launches_per_step": launches,
           "T_max_K": float(state["T"].max()), **parts,
           "regime": "L2-resident systems: latency bound, HBM roofline not applicable"}
    if o is not None:
        ob.ch_setup(HEAT_T_AMBIENT)
        best = None
        for k in range(3):
            t = time.perf_counter()
            o.setup(-E0 * HEAT_RAMP[k], 0.0, False); o.assemble(True); fit = o.solve(N_CG, CG_TOL, 1.2, 0)
            o.extract_solution(True)
            t1 = time.perf_counter()
            osol = o.interpolate(2, 3, cen, ocells)
            Jo, no = synthetic_emission(np.sqrt((osol[:, :3] ** 2).sum(1)))
            t2 = time.perf_counter()
            ob.current_assemble(Jo); cit = ob.ch_solve(0, N_CG, CG_TOL, 1.2, 0)
            t3 = time.perf_counter()
            ob.heat_assemble(HEAT_DT, no); hit = ob.ch_solve(1, N_CG, CG_TOL, 1.2, 0)
            t4 = time.perf_counter()
            r = {"ms_per_step": 1e3 * (t4 - t), "field_ms": 1e3 * (t1 - t), "surface_field_ms": 1e3 * (t2 - t1), "current_ms": 1e3 * (t3 - t2),
                 "heat_ms": 1e3 * (t4 - t3), "cg_iterations_field_current_heat": [fit, cit, hit], "cores": cpu_threads(),
                 "kind": "port (SSOR-CG, deal.II semantics; serial assembly as the reference's copier)"}
            if k > 0 and (best is None or r["ms_per_step"] < best["ms_per_step"]):       # step 0 starts from the ambient state
                best = r
        out["cpu_baseline"] = best
    vctx.close(); bctx.close()
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
    ap.add_argument("--cg-p2p", type=int, default=None, help="0: NCCL inside the partitioned CG iteration instead of the peer-mapped mode")
    ap.add_argument("--skip-native", action="store_true")
    ap.add_argument("--only", default=None, choices=["native", "pic", "heat"],
                    help="development aid: run ONE sub-benchmark alone and print its object (not the contract line)")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-verify", action="store_true", help="skip the host-side residual check of the X leg (7 GB D2H)")
    args = ap.parse_args()
    if args.only:
        import torch
        import femocs_b200 as fb
        print(json.dumps({args.only: {"native": native_step, "pic": pic_step, "heat": heat_step}[args.only](fb, torch, args)}), flush=True)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
