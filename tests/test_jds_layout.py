"""Structural test (CPU, no CUDA) of the block-JDS tables that the HBM SpMV kernels walk (fb_host_jds_build):
every stored slot, decoded the way k_spmv_jds / k_spmv_sym decode it, is exactly one entry of the CSR pattern,
and every entry that must be stored is stored once."""
import numpy as np
import pytest

from femocs_b200 import synth
from femocs_b200.solver import PartitionPlan


def _plan(golden, name, refine=0):
    m = golden("mesh", name)
    nodes, hexs, mk = m["nodes"], m["hexs"], m["hex_markers"]
    if refine:
        nodes, hexs, mk = synth.refine_vacuum(nodes, hexs, mk, refine)
    p = PartitionPlan(0, 1)
    p.phase2(p.phase1(nodes, hexs, mk))
    return p


def _decode(p, t, sym):
    """(row, column) of every stored slot, walking blocks / slots / diagonals like the kernels do"""
    R = t["R"]; n = p.n_rows
    rows, cols, poss = [], [], []
    for b in range(t["nb"]):
        r0 = b * R; nr = min(R, n - r0)
        jd = t["jd"][t["jdp"][b]:t["jdp"][b + 1]]
        win = t["win_list"][t["win_off"][b]:t["win_off"][b + 1]]
        assert np.all(np.diff(win) > 0)                                      # sorted distinct columns
        assert np.all(jd % 2 == 0) and jd[0] == 0 and np.all(np.diff(jd) >= 0)   # diagonals padded to pairs of slots
        lens = t["len"][b * R:b * R + nr].astype(np.int64)
        assert np.all(np.diff(lens) <= 0)                                    # rows sorted by length inside the block
        perm = t["perm"][b * R:b * R + nr].astype(np.int64)
        assert np.array_equal(np.sort(perm), np.arange(nr))
        assert np.array_equal(t["slot"][r0 + perm], np.arange(nr))
        for tt in range(nr):
            L = lens[tt]
            if L == 0:
                continue
            pos = t["base"][b] + jd[:L] + tt
            w = t["col16"][pos].astype(np.int64)
            if sym:
                c = np.where(w < len(win), win[np.minimum(w, max(len(win) - 1, 0))] if len(win) else 0, r0 + (w - len(win)))
            else:
                c = win[w]
            rows.append(np.full(L, r0 + perm[tt])); cols.append(c); poss.append(pos)
        # a diagonal holds the slots of the rows that are long enough, rounded up to an even count
        active = np.array([(lens > j).sum() for j in range(len(jd) - 1)])
        assert np.array_equal(np.diff(jd), (active + 1) // 2 * 2)
    return np.concatenate(rows), np.concatenate(cols), np.concatenate(poss)


@pytest.mark.parametrize("name,refine,R", [("hemicone", 0, 128), ("mdsmall", 0, 256), ("mdsmall", 1, 512)])
@pytest.mark.parametrize("sym", [False, True])
def test_block_jds_tables_encode_the_csr_pattern(name, refine, R, sym, golden):
    p = _plan(golden, name, refine)
    try:
        t = p.jds(R=R, max_window=8192 if not sym else 8192, sym=sym)
    except Exception as e:                        # a native tet-split mesh in first-touch order can exceed the window cap
        pytest.skip(str(e))
    rows, cols, poss = _decode(p, t, sym)
    assert len(np.unique(poss)) == len(poss) and poss.max() < t["size"]       # no two entries share a slot
    rp, col = p.rowptr, p.col
    csr_rows = np.repeat(np.arange(p.n_rows), np.diff(rp))
    if sym:
        keep = col < csr_rows                                                # strictly lower triangle
        csr_rows, csr_cols = csr_rows[keep], col[keep]
    else:
        csr_cols = col
    assert len(rows) == len(csr_rows)
    a = np.lexsort((cols, rows)); b = np.lexsort((csr_cols, csr_rows))
    assert np.array_equal(rows[a], csr_rows[b]) and np.array_equal(cols[a], csr_cols[b])
    assert t["maxlen"] == np.max(np.bincount(csr_rows, minlength=p.n_rows))


@pytest.mark.parametrize("name,refine", [("mdsmall", 1), ("hemicone", 1)])
@pytest.mark.parametrize("cap", [32, 12])
def test_split_row_layout_multiplies_like_csr(name, refine, cap, golden):
    """spmv_kernel 306: rows longer than `cap` entries are stored as chained segments (jds_link) and a block holds 512 slots
    instead of 512 rows (jds_rowbeg).  Walking the tables the way k_spmv_jds<., SPLIT> does -- slot sums, then the heads add
    their chain in order -- reproduces the CSR product; every row is the head of exactly one chain; no segment exceeds cap."""
    p = _plan(golden, name, refine)
    t = p.jds(R=512, max_window=8192, split=cap)
    R = 512; n = p.n_rows
    rng = np.random.default_rng(5)
    val = rng.standard_normal(p.nnz); x = rng.standard_normal(p.n_cols)
    rp, col = p.rowptr, p.col
    ref = np.add.reduceat(val * x[col], rp[:-1])
    assert t["maxlen"] <= cap and t["rowbeg"][0] == 0 and t["rowbeg"][-1] == n and np.all(np.diff(t["rowbeg"]) > 0)
    out = np.full(n, np.nan); stored = 0
    rowlen = np.diff(rp)
    for b in range(t["nb"]):
        r0 = t["rowbeg"][b]; nr = t["rowbeg"][b + 1] - r0
        perm = t["perm"][b * R:(b + 1) * R].astype(np.int64); lens = t["len"][b * R:(b + 1) * R].astype(np.int64)
        link = t["link"][b * R:(b + 1) * R].astype(np.int64)
        live = perm != 0xFFFF
        nv = int(live.sum())
        assert np.all(live[:nv]) and np.all(np.diff(lens[:nv]) <= 0) and np.all(lens[nv:] == 0)
        jd = t["jd"][t["jdp"][b]:t["jdp"][b + 1]]
        win = t["win_list"][t["win_off"][b]:t["win_off"][b + 1]]
        # the value array the device kernel k_csr_to_jds_split would produce: entry o of row r -> segment o // seglen
        sums = np.zeros(R)
        head = live & ((perm & 0x8000) == 0)
        assert np.array_equal(np.sort(perm[head]), np.arange(nr))
        assert np.array_equal(t["slot"][r0 + perm[head]], np.flatnonzero(head))
        for tt in np.flatnonzero(head):
            r = r0 + perm[tt]; L = rowlen[r]
            k = -(-L // cap) if L > cap else 1
            sl = -(-L // k) if L else 0
            s = tt; q = 0; total = 0.0; first = True
            while s != 0xFFFF:
                assert (perm[s] & 0x7FFF) == perm[tt] and (first or perm[s] & 0x8000)
                seg = lens[s]
                assert seg == max(0, min(sl, L - q * sl))
                pos = t["base"][b] + jd[:seg] + s
                cols = win[t["col16"][pos].astype(np.int64)]
                ks = rp[r] + q * sl + np.arange(seg)
                assert np.array_equal(cols, col[ks])
                sums[s] = float(np.dot(val[ks], x[cols]))
                stored += seg
                total = sums[s] if first else total + sums[s]
                first = False; q += 1; s = link[s]
            assert q == k
            out[r] = total
    assert stored == p.nnz
    assert np.abs(out - ref).max() <= 1e-12 * np.abs(ref).max()
