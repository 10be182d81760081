"""world_size-2 (and 4) CPU tests of the multi-GPU path's host logic over the gloo backend: the partitioner,
the local sparsity, the halo plan and the Dirichlet flags of the product library (fb_plan_*, the same code
fb_import_mesh runs under torchrun) reproduce the single-rank operator when the data path -- halo exchange,
local SpMV, dot-product all-reduce -- is carried out with torch.distributed on CPU tensors."""
import os
import socket
import sys

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, mesh_name, out_q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import femocs_b200 as fb
        from oracle.oracle import Oracle
        with np.load(os.path.join(ROOT, "tests", "golden", "mesh_%s.npz" % mesh_name)) as z:
            m = {k: z[k] for k in ("nodes", "hexs", "hex_markers")}
        plan = fb.PartitionPlan(rank, world)
        bb = plan.phase1(m["nodes"], m["hexs"], m["hex_markers"])
        lo = torch.from_numpy(bb[:3].copy()); hi = torch.from_numpy(bb[3:].copy())
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        plan.phase2(np.concatenate([lo.numpy(), hi.numpy()]))
        nr, nc = plan.n_rows, plan.n_cols
        l2g = plan.local2global.astype(np.int64)

        # single-rank truth from the CPU oracle, re-indexed by global solver vertex
        o = Oracle(); o.import_mesh(m["nodes"], m["hexs"], m["hex_markers"])
        o.setup(0.5, 0.0, False); o.assemble(True)
        rp, col, val, save = o.csr(); _, _, v2d, _ = o.vectors()
        d2v = np.empty_like(v2d); d2v[v2d] = np.arange(len(v2d))
        K = sp.csr_matrix((save, col, rp)); K = K[v2d][:, v2d].tocsr(); K.sort_indices()        # vertex numbering
        nv = o.n_vertices
        assert plan.n_vert_global == nv and plan.n_cells_global == o.n_cells

        # 1. ownership: a partition of the vertices, balanced, ghosts grouped by owner in ascending global id
        counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(counts, torch.tensor([nr]))
        counts = [int(c) for c in counts]
        assert sum(counts) == nv and max(counts) - min(counts) <= world
        assert np.all(plan.owner[:nr] == rank) and np.all(plan.owner[nr:] != rank)
        g_owner = plan.owner[nr:]; g_ids = l2g[nr:]
        assert np.all(np.diff(g_owner) >= 0)
        for p in range(world):
            seg = g_ids[plan.recv_off[p]:plan.recv_off[p + 1]]
            assert np.all(g_owner[plan.recv_off[p]:plan.recv_off[p + 1]] == p) and np.all(np.diff(seg) > 0)

        # 2. local sparsity = the owned rows of the global pattern
        Kl = K[l2g[:nr]].tocsr()
        assert np.array_equal(np.diff(plan.rowptr), np.diff(Kl.indptr))
        mine = np.sort(l2g[plan.col].reshape(-1)); theirs = np.sort(Kl.indices)
        for r in range(0, nr, max(1, nr // 200)):
            a = np.sort(l2g[plan.col[plan.rowptr[r]:plan.rowptr[r + 1]]]); b = np.sort(Kl.indices[Kl.indptr[r]:Kl.indptr[r + 1]])
            assert np.array_equal(a, b)
        assert np.array_equal(mine, theirs)

        # 3. halo exchange with the plan's lists (what ncclSend / ncclRecv do on the GPUs)
        rng = np.random.default_rng(7)
        xg = rng.standard_normal(nv)

        def halo(vec):
            reqs = []; bufs = {}
            for p in range(world):
                if p == rank:
                    continue
                s_ = plan.send_idx[plan.send_off[p]:plan.send_off[p + 1]]
                if len(s_):
                    reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(vec[s_])), p))
                n_r = plan.recv_off[p + 1] - plan.recv_off[p]
                if n_r:
                    bufs[p] = torch.zeros(n_r, dtype=torch.float64)
                    reqs.append(dist.irecv(bufs[p], p))
            for r_ in reqs:
                r_.wait()
            for p, b in bufs.items():
                vec[nr + plan.recv_off[p]: nr + plan.recv_off[p + 1]] = b.numpy()

        xl = np.zeros(nc); xl[:nr] = xg[l2g[:nr]]
        halo(xl)
        assert np.array_equal(xl, xg[l2g])

        # 4. distributed SpMV and dot product = the global ones
        vals = np.asarray(K[l2g[:nr]][:, l2g].todense()) if nr * nc < 4e6 else None
        Al = sp.csr_matrix(K[l2g[:nr]][:, l2g])
        yl = Al @ xl
        ref = (K @ xg)[l2g[:nr]]
        assert np.abs(yl - ref).max() <= 1e-12 * np.abs(ref).max()
        dot = torch.tensor([float(xl[:nr] @ yl)], dtype=torch.float64)
        dist.all_reduce(dot)
        assert abs(float(dot) - float(xg @ (K @ xg))) <= 1e-10 * abs(float(xg @ (K @ xg)))

        # 5. Dirichlet candidates: owned flags are complete, ghost flags come from the owner
        cells = o.cells(); bc, bf, bid = o.bfaces()
        FV = np.array([[0, 2, 4, 6], [1, 3, 5, 7], [0, 1, 4, 5], [2, 3, 6, 7], [0, 1, 2, 3], [4, 5, 6, 7]])
        cu = np.zeros(nv, bool); top = np.zeros(nv, bool)
        for ids, flag in ((2, cu), (8, top)):
            sel = bid == ids
            flag[np.unique(cells[bc[sel]][np.arange(sel.sum())[:, None], FV[bf[sel]]])] = True
        assert np.array_equal(plan.copper[:nr].astype(bool), cu[l2g[:nr]])
        assert np.array_equal(plan.top[:nr].astype(bool), top[l2g[:nr]])
        f = plan.copper.astype(np.float64); f[nr:] = 0
        halo(f)
        assert np.array_equal(f.astype(bool), cu[l2g])
        out_q.put((rank, "ok", nr, nc, plan.n_send))
    except Exception as e:           # noqa: BLE001
        import traceback
        out_q.put((rank, "FAIL: %s\n%s" % (e, traceback.format_exc()), 0, 0, 0))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,mesh", [(2, "mdsmall"), (4, "hemicone"), (3, "mdsmall"), (8, "mdbig")])
def test_partition_plan_gloo(world, mesh):
    import torch.multiprocessing as mp
    from femocs_b200 import build
    build.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, mesh, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for r in sorted(res):
        assert r[1] == "ok", r[1]
    assert all(r[4] > 0 for r in res)           # every rank has a halo
