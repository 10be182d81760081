"""Partitioned CG on 2 (or more) GPUs of one box: one process per GPU, NCCL halo exchange + all-reduce inside
libfemocs_b200, against the CPU oracle and the single-GPU path.  Skipped when fewer than 2 GPUs are visible."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q, p2p):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import femocs_b200 as fb
        with np.load(os.path.join(ROOT, "tests", "golden", "mesh_mdsmall.npz")) as z:
            m = {k: z[k] for k in z.files}
        with np.load(os.path.join(ROOT, "tests", "golden", "interp_mdsmall.npz")) as z:
            g = {k: z[k] for k in ("points", "pic_cells", "pic_ok")}
        torch.cuda.set_device(rank)
        ctx = fb.Context(rank)
        ctx.init_comm_torch(dist)
        ctx.set_option("cg_p2p", p2p)
        s = fb.PoissonSolver(ctx, fb.FieldConfig(cg_tolerance=1e-11, mode="transient"))
        assert s.import_mesh(m["nodes"], m["hexs"], m["hex_markers"])
        part = ctx.partition()
        part["comm_mode"] = ctx.comm_mode
        ok = g["pic_ok"]
        cf = -180.9512268 * 0.01
        out = {"part": part}
        # Laplace, then Poisson with the golden particles (global solver cell ids, replicated on every rank)
        s.setup(0.5, 0.0); s.assemble(True)
        out["it_laplace"] = s.solve(); out["phi_laplace"] = s.export_solution()
        bad = s.check_limits(-1.0, 1e4); out["limits"] = (bad, s.stat_sol_min, s.stat_sol_max)
        s.set_particles(g["points"][ok], g["pic_cells"][ok], cf)
        s.setup(0.5, 0.0); s.assemble(True)
        out["it_poisson"] = s.solve(); out["phi_poisson"] = s.export_solution()
        s.assemble(False)                                   # PIC step: warm start from the converged potential
        out["it_warm"] = s.solve(cg_tolerance=1e-7)
        out["launches"] = ctx.kernel_launches
        # the HBM kernel of the benchmark (segmented block-JDS, chosen automatically only for >= 4e6 non-zeros) on the partition
        ctx.set_option("spmv_kernel", 306)
        s.setup(0.5, 0.0); s.assemble(True)
        out["it_jds"] = s.solve(); out["phi_jds"] = s.export_solution(); out["kernel_jds"] = s.solve_kernel()
        ctx.set_option("spmv_kernel", -1)
        if p2p:     # the two-level preconditioner on the partitioned mesh (global Morton aggregates, restriction summed over the peer mappings)
            s.conf.precond = fb.PRECOND_TWOLEVEL
            ctx.set_option("tl_agg", 64)
            s.setup(0.5, 0.0); s.assemble(True)
            out["it_tl"] = s.solve(); out["phi_tl"] = s.export_solution()
            s.conf.precond = fb.PRECOND_JACOBI
        # atoms are independent once the potential is known -- every rank keeps a REPLICA of the mesh
        # tables (MBs) in a second, un-partitioned context, takes the complete potential of the partitioned solve and
        # interpolates ITS shard of the atoms (contiguous blocks: the guess chain restarts per shard); no collective
        rep = fb.Context(rank)
        rs = fb.PoissonSolver(rep); assert rs.import_mesh(m["nodes"], m["hexs"], m["hex_markers"])
        rs.import_solution(out["phi_poisson"])
        it = fb.Interpolator(rep); it.initialize(m); it.extract_solution(rs, True)
        atoms = m["atoms"]; lo = len(atoms) * rank // world; hi = len(atoms) * (rank + 1) // world
        f = fb.FieldReader(it); f.set_preferences(False, 3, 1); f.interpolate(atoms[lo:hi])
        out["shard"] = (lo, hi, f.markers.copy(), f.interpolation.copy())
        rep.close()
        ctx.close()
        q.put((rank, "ok", out))
    except Exception as e:           # noqa: BLE001
        import traceback
        q.put((rank, "FAIL: %s\n%s" % (e, traceback.format_exc()), None))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("p2p", [1, 0], ids=["peer_mapped", "nccl"])
@pytest.mark.parametrize("world", [2, 4])
def test_partitioned_solve_matches_oracle(world, p2p, golden):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import torch.multiprocessing as mp
    from oracle.oracle import Oracle
    mpc = mp.get_context("spawn")
    q = mpc.Queue(); port = _free_port()
    procs = [mpc.Process(target=_worker, args=(r, world, port, q, p2p)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    res.sort(key=lambda r: r[0])
    for r in res:
        assert r[1] == "ok", r[1]
    m = golden("mesh", "mdsmall"); g = golden("interp", "mdsmall")
    o = Oracle(); o.import_mesh(m["nodes"], m["hexs"], m["hex_markers"]); o.interp_initialize(m)
    o.setup(0.5, 0.0, False); o.assemble(True); o.solve(10000, 1e-11, 1.2, 0)
    ref_l = o.export_solution()
    ok = g["pic_ok"]
    o.setup(0.5, 0.0, False); o.assemble(True, g["points"][ok], g["pic_cells"][ok], -180.9512268 * 0.01); o.solve(10000, 1e-11, 1.2, 0)
    ref_p = o.export_solution()
    rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
    outs = [r[2] for r in res]
    assert sum(x["part"]["n_rows"] for x in outs) == o.n_vertices
    # sharded atom interpolation on the replicas: every shard equals the oracle's answer for that block of atoms
    o.extract_solution(True)
    for x in outs:
        lo, hi, cells, sol = x["shard"]
        oc, osol = o.locate_interpolate(3, 1, m["atoms"][lo:hi])
        assert np.array_equal(cells, oc)
        assert np.abs(sol - osol).max() <= 1e-8 * np.abs(osol).max()
    for x in outs:
        # 2 = the iteration runs over peer mappings (CUDA IPC + NVLink stores, CUDA-graph captured), 1 = NCCL inside the iteration
        assert x["part"]["comm_mode"] == (2 if p2p else 1), x["part"]
        assert x["it_laplace"] > 0 and x["it_poisson"] > 0 and x["it_warm"] == 0
        assert x["it_laplace"] == outs[0]["it_laplace"]                      # every rank takes the same decisions
        assert rel(x["phi_laplace"], ref_l) < 1e-8                            # complete potential on every rank
        assert rel(x["phi_poisson"], ref_p) < 1e-8
        assert not x["limits"][0] and x["limits"][1] == 0.0 and abs(x["limits"][2] - ref_l.max()) < 1e-8 * ref_l.max()
        assert x["part"]["n_send"] > 0 and x["launches"] > 0
        assert x["kernel_jds"] == 306 and x["it_jds"] > 0 and rel(x["phi_jds"], ref_p) < 1e-8
        if p2p:
            assert 0 < x["it_tl"] < 0.8 * x["it_poisson"] and x["it_tl"] == outs[0]["it_tl"], (x["it_tl"], x["it_poisson"])
            assert rel(x["phi_tl"], ref_p) < 1e-8
