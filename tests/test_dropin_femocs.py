"""Drop-in test: the reference's OWN Femocs / Femocs_wrap / ProjectRunaway / Interpolator / SolutionReader, compiled
verbatim from /root/reference against the product header include/dropin/PoissonSolver.h and linked with
libfemocs_b200.so (oracle/Makefile.dropin -> oracle/_ref/libfemocs_dropin.so), driven through the untouched C ABI of
include/Femocs_wrap.h:12-50:

    create_femocs -> femocs_import_file -> femocs_run -> femocs_export_data / femocs_interpolate_elfield /
    femocs_interpolate_surface_elfield / femocs_interpolate_phi -> delete_femocs

The atoms are those of in/nanotip_small.xyz (kept in tests/golden/mesh_mdsmall.npz); the mesher that runs inside
femocs_run is the reference's, so the mesh must be the golden one, and the fields must match the CPU oracle run on
that mesh: potential-derived quantities to 1e-8, located cells (flags) exactly.
Without a GPU the same sequence must FAIL LOUDLY (femocs_run returns 1: no CPU fallback) after a correct mesh."""
import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libfemocs_dropin.so")

# the configuration text of the reference's demo driver for this case (src/main/Main.cpp:25-50 write_defaults,
# :93-101 mdsmall), field solver to 1e-11 so that both sides are over-converged; forces are outside the hot path
CONF = """
project = runaway
mesh_quality = 1.8
heat_mode = none
field_mode = laplace
clear_output = false
surface_smooth_factor = 0.1
charge_smooth_factor = 1.0
distance_tol = 0.0
n_write_log = 0
write_period = -1
use_rdf = false
clean_surface = true
surface_thickness = 3.1
coord_cutoff = 3.1
charge_cutoff = 30
latconst = 3.61
femocs_verbose_mode = mute
smooth_steps = 3
smooth_algorithm = laplace
elfield = -0.5
interpolation_rank = 1
force_mode = none
coarse_rate = 0.5
seed = 12345
coarse_factor = 0.3 4 2
radius = 16.0
box_width = 5.0
box_height = 5.0
field_cgtol = 1e-11
"""

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)


def _run_child(code):
    """the reference keeps process-wide globals and chdir()s: run every session in its own interpreter"""
    # Femocs::interpolate_elfield hands export_results an UNINITIALISED new double[3n] (Femocs.cpp:219) and the
    # lower-case label makes SolutionReader::export_vec ADD to it (SolutionReader.cpp:306,330-335): the reference's
    # result is heap garbage + E unless malloc happens to return zero pages.  MALLOC_PERTURB_=255 makes glibc fill
    # every allocation with ~255 = 0x00, which pins that undefined behaviour to the intended value.
    env = dict(os.environ, PYTHONPATH=ROOT, MALLOC_PERTURB_="255")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


SESSION = r'''
import ctypes as C, os, sys, tempfile, json
import numpy as np
ROOT = %(root)r
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_dropin_femocs import CONF, LIB, dp, ip
m = dict(np.load(os.path.join(ROOT, "tests", "golden", "mesh_mdsmall.npz")))
work = tempfile.mkdtemp(prefix="femocs_dropin_")
os.chdir(work)
atoms = m["atoms"]
with open("atoms.xyz", "w") as f:
    f.write("%%d\ncomment\n" %% len(atoms))
    for a in atoms:
        f.write("Cu %%s %%s %%s -1\n" %% (repr(float(a[0])), repr(float(a[1])), repr(float(a[2]))))
open("conf.in", "w").write(CONF)
L = C.CDLL(LIB)
L.create_femocs.restype = C.c_void_p
L.create_femocs.argtypes = [C.c_char_p]
for name in ("femocs_run",):
    getattr(L, name).argtypes = [C.c_void_p, ip, C.c_int, C.c_double]
L.femocs_import_file.argtypes = [C.c_void_p, ip, C.c_char_p]
L.femocs_export_data.argtypes = [C.c_void_p, ip, dp, C.c_int, C.c_char_p]
for name in ("femocs_interpolate_elfield", "femocs_interpolate_surface_elfield"):
    getattr(L, name).argtypes = [C.c_void_p, ip, C.c_int, dp, dp, dp, dp, dp, dp, dp, ip]
L.femocs_interpolate_phi.argtypes = [C.c_void_p, ip, C.c_int, dp, dp, dp, dp, ip]
L.delete_femocs.argtypes = [C.c_void_p]
L.dropin_mesh_count.argtypes = [C.c_void_p, C.c_int]
L.dropin_mesh_copy.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
rv = C.c_int(-1)
fem = L.create_femocs(b"conf.in")
L.femocs_import_file(fem, C.byref(rv), b"atoms.xyz"); assert rv.value == 0
L.femocs_run(fem, C.byref(rv), 0, 0.0)
out = {"run": rv.value}
# the mesh the reference's mesher produced inside femocs_run
mesh = {}
for kind, (key, width, dt) in enumerate([("nodes", 3, np.float64), ("tets", 4, np.int32), ("hexs", 8, np.int32), ("tris", 3, np.int32), ("quads", 4, np.int32)]):
    n = L.dropin_mesh_count(fem, kind)
    a = np.zeros((max(n, 0), width), dt)
    if n > 0: L.dropin_mesh_copy(fem, kind, a.ctypes.data)
    mesh[key] = a
out["mesh_equal"] = {k: bool(mesh[k].shape == m[k].shape and np.array_equal(mesh[k], m[k])) for k in mesh}
if rv.value == 0:
    n = len(atoms)
    def f(a): return np.ascontiguousarray(a, np.float64)
    # femocs_export_data: field on the atoms the run interpolated (surface atoms; zeros elsewhere)
    E = np.zeros(3 * n); L.femocs_export_data(fem, C.byref(rv), E.ctypes.data_as(dp), n, b"elfield"); out["export_rv"] = rv.value
    En = np.zeros(n); L.femocs_export_data(fem, C.byref(rv), En.ctypes.data_as(dp), n, b"elfield_norm")
    ph = np.zeros(n); L.femocs_export_data(fem, C.byref(rv), ph.ctypes.data_as(dp), n, b"potential")
    pts = f(atoms[::10]); k = len(pts)
    x, y, z = f(pts[:, 0]), f(pts[:, 1]), f(pts[:, 2])
    res = {}
    for name in ("femocs_interpolate_elfield", "femocs_interpolate_surface_elfield"):
        Ex, Ey, Ez, Enorm = np.zeros(k), np.zeros(k), np.zeros(k), np.zeros(k); flag = np.zeros(k, np.int32)
        getattr(L, name)(fem, C.byref(rv), k, x.ctypes.data_as(dp), y.ctypes.data_as(dp), z.ctypes.data_as(dp), Ex.ctypes.data_as(dp),
                         Ey.ctypes.data_as(dp), Ez.ctypes.data_as(dp), Enorm.ctypes.data_as(dp), flag.ctypes.data_as(ip))
        res[name] = (rv.value, np.stack([Ex, Ey, Ez], 1), Enorm, flag)
    phi = np.zeros(k); pflag = np.zeros(k, np.int32)
    L.femocs_interpolate_phi(fem, C.byref(rv), k, x.ctypes.data_as(dp), y.ctypes.data_as(dp), z.ctypes.data_as(dp), phi.ctypes.data_as(dp), pflag.ctypes.data_as(ip))
    np.savez(os.path.join(work, "result.npz"), E=E.reshape(n, 3), En=En, ph=ph, pts=pts,
             vol_E=res["femocs_interpolate_elfield"][1], vol_flag=res["femocs_interpolate_elfield"][3], vol_rv=res["femocs_interpolate_elfield"][0],
             surf_E=res["femocs_interpolate_surface_elfield"][1], surf_flag=res["femocs_interpolate_surface_elfield"][3],
             surf_norm=res["femocs_interpolate_surface_elfield"][2], phi=phi, phi_flag=pflag)
    out["result"] = os.path.join(work, "result.npz")
L.delete_femocs(fem)
print("RESULT " + json.dumps(out))
'''


def _session():
    import json
    out = _run_child(SESSION % dict(root=ROOT))
    line = [ln for ln in out.splitlines() if ln.startswith("RESULT ")][-1]
    return json.loads(line[len("RESULT "):])


pytestmark = pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/libfemocs_dropin.so not built (needs /root/reference at build time)")


def test_dropin_fails_loudly_without_gpu_after_meshing():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = _session()
    assert r["run"] == 1                       # import_mesh returned false: no CPU fallback behind the seam
    assert all(r["mesh_equal"].values()), r    # ... but the reference's mesher ran and reproduced the golden mesh


@pytest.mark.gpu
def test_reference_femocs_runs_on_the_gpu_through_its_own_c_abi(golden):
    from oracle.oracle import Oracle
    r = _session()
    assert r["run"] == 0, r
    assert all(r["mesh_equal"].values()), r
    z = dict(np.load(r["result"]))
    m = golden("mesh", "mdsmall")
    o = Oracle(); o.import_mesh(m["nodes"], m["hexs"], m["hex_markers"]); o.interp_initialize(m)
    o.setup(0.5, 0.0, False); o.assemble(True)                      # solve_laplace(conf.field.E0 = -0.5): setup(-E0, V0)
    assert o.solve(10000, 1e-11, 1.2, 0) > 0
    o.extract_solution(True)                                        # smoothen_field defaults to true (Config.cpp:40)
    rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
    # prepare_export: fields.interpolate(dense_surf) with dim 2, rank 1, exported by atom id (ProjectRunaway.cpp:303-304)
    cells, sol = o.locate_interpolate(2, 1, m["surf_atoms"])
    ids = m["surf_ids"]
    E = np.zeros((len(m["atoms"]), 3)); E[ids] = sol[:, :3]
    ph = np.zeros(len(m["atoms"])); ph[ids] = sol[:, 4]
    assert rel(z["E"], E) < 1e-8 and rel(z["ph"], ph) < 1e-8
    assert rel(z["En"][ids], np.sqrt((sol[:, :3] ** 2).sum(1))) < 1e-8
    other = np.ones(len(E), bool); other[ids] = False
    assert np.all(z["E"][other] == 0)
    # femocs_interpolate_elfield / _phi on every 10th atom (dim 3) and femocs_interpolate_surface_elfield (dim 2).
    # The flags the reference hands out are NOT those of the query points: ProjectRunaway::interpolate reads
    # fields.get_marker(i), the cell of the i-th SURFACE atom of the last prepare_export (ProjectRunaway.cpp:655-656),
    # so the query is kept shorter than the surface-atom list and the expectation is that list's located cells.
    c3, s3 = o.locate_interpolate(3, 1, z["pts"])
    assert z["vol_rv"] == 0
    k = len(z["pts"]); assert k <= len(cells)
    assert np.array_equal(z["vol_flag"], (cells[:k] >= 0).astype(np.int32))     # located cells: exact
    assert rel(z["vol_E"], s3[:, :3]) < 1e-8
    assert rel(z["phi"], s3[:, 4]) < 1e-8 and np.array_equal(z["phi_flag"], z["vol_flag"])
    c2, s2 = o.locate_interpolate(2, 1, z["pts"])
    assert np.array_equal(z["surf_flag"], z["vol_flag"])
    assert rel(z["surf_E"], s2[:, :3]) < 1e-8
    assert rel(z["surf_norm"], np.sqrt((s2[:, :3] ** 2).sum(1))) < 1e-8
