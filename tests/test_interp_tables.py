"""CPU parity of the HOST product code behind fb_interp_initialize (fb_host_interp_tables, femocs_b200/csrc/host_setup.cpp):
the cell tables it uploads are bit-identical to the oracle's, which tests/test_oracle_vs_ref.py pins to the reference's
own compiled precompute() code (src/InterpolatorCells.cpp:523-629, 1205-1267, 1585-1637, 1151-1173, 1873-1895)."""
import ctypes as C

import numpy as np
import pytest

from femocs_b200 import lib as fblib
from oracle.oracle import Oracle


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("name", ["hemicone", "mdsmall", "mdbig"])
def test_host_tables_bit_identical_to_oracle(name, golden):
    m = golden("mesh", name)
    o = Oracle(); o.import_mesh(m["nodes"], m["hexs"], m["hex_markers"]); o.interp_initialize(m)
    ref = o.tables()
    L = fblib.load()
    plan = C.c_void_p(L.fb_plan_create(0, 1))
    f = lambda a: np.ascontiguousarray(a, np.float64)
    i = lambda a: np.ascontiguousarray(a, np.int32)
    nodes, hexs, hm = f(m["nodes"]), i(m["hexs"]), i(m["hex_markers"])
    nm, tets, nbr, tm = i(m["node_markers"]), i(m["tets"]), i(m["tet_nbrs"]), i(m["tet_markers"])
    tris, tn, quads = i(m["tris"]), f(m["tri_norms"]), i(m["quads"])
    got = {k: np.zeros_like(v) for k, v in ref.items()}
    rc = L.fb_plan_interp_tables(plan, _p(nodes), len(nodes), _p(hexs), _p(hm), len(hexs), _p(nm), _p(tets), _p(nbr), _p(tm), len(tets),
                                 _p(tris), _p(tn), len(tris), _p(quads), len(quads),
                                 _p(got["tet"]), _p(got["tet_cent"]), _p(got["tet_mark"]), _p(got["hex"]), _p(got["tri"]), _p(got["tri_cent"]),
                                 _p(got["qtet"]), _p(got["qtri"]))
    assert rc == 0, L.fb_last_error(plan).decode()
    for k in ("tet", "tet_cent", "tet_mark", "hex", "tri", "tri_cent", "qtet", "qtri"):
        assert np.array_equal(got[k], ref[k]), k              # bit for bit (both sides compiled without FMA contraction)
    L.fb_destroy(plan)
