"""The C++ seam (include/femocs_b200.hpp: PoissonSolver / Interpolator / FieldReader with the reference's
member names) driven like ProjectRunaway::run by tests/cxx/seam_driver.cpp, checked against the oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "tests", "cxx", "seam_driver")


def _write_blob(path, arrays):
    with open(path, "wb") as f:
        for name, a in arrays.items():
            a = np.ascontiguousarray(a)
            kind = 0 if a.dtype.kind in "iu" else 1
            a = a.astype(np.int32 if kind == 0 else np.float64)
            nm = name.encode()
            f.write(struct.pack("<i", len(nm))); f.write(nm)
            f.write(struct.pack("<iq", kind, a.size)); f.write(a.tobytes())


def _read_blob(path):
    out = {}
    with open(path, "rb") as f:
        while True:
            h = f.read(4)
            if not h:
                break
            (nl,) = struct.unpack("<i", h)
            name = f.read(nl).decode()
            kind, cnt = struct.unpack("<iq", f.read(12))
            out[name] = np.frombuffer(f.read((4 if kind == 0 else 8) * cnt), dtype=np.int32 if kind == 0 else np.float64)
    return out


def build_driver():
    from femocs_b200 import build
    lib = build.build()
    src = os.path.join(ROOT, "tests", "cxx", "seam_driver.cpp")
    deps = [src, os.path.join(ROOT, "include", "femocs_b200.hpp"), os.path.join(ROOT, "include", "femocs_b200.h")]
    if (not os.path.exists(DRIVER)) or any(os.path.getmtime(d) > os.path.getmtime(DRIVER) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), src, "-o", DRIVER,
                               "-L", os.path.dirname(lib), "-lfemocs_b200", "-Wl,-rpath," + os.path.dirname(lib)])
    return DRIVER


def test_seam_compiles_and_links():
    """the header-only seam compiles as C++17 against the C ABI and links with the shared library (no GPU needed)"""
    assert os.path.exists(build_driver())


def test_seam_fails_loudly_without_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    _write_blob(tmp_path / "m.bin", {"nodes": np.zeros(3)})
    r = subprocess.run([build_driver(), str(tmp_path / "m.bin"), str(tmp_path / "o.bin"), "-0.5"], capture_output=True, text=True)
    assert r.returncode == 3 and "no CUDA device" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["mdsmall", "mdbig"])
def test_seam_matches_oracle(name, golden, tmp_path):
    from oracle.oracle import Oracle
    from oracle import pic as opic
    m = golden("mesh", name)
    E0 = -0.5
    o = Oracle(); o.import_mesh(m["nodes"], m["hexs"], m["hex_markers"]); o.interp_initialize(m)
    # super-particles: points 2 A above the first surface atoms that lie in a vacuum cell, one common velocity
    p0 = m["surf_atoms"][:400] + np.array([0.0, 0.0, 2.0])
    c0 = o.particle_cells(p0, np.zeros(len(p0), np.int32))
    p0 = np.ascontiguousarray(p0[c0 >= 0]); c0 = c0[c0 >= 0].astype(np.int32)
    v0 = np.tile(np.array([0.3, -0.2, 0.5]), (len(c0), 1))
    lo = m["nodes"].min(0); hi = m["nodes"].max(0)
    box = np.array([lo[0], hi[0], lo[1], hi[1], lo[2], hi[2]]); dt, qm = 0.5, -17.5882
    blob = {k: m[k] for k in ("nodes", "node_markers", "hexs", "hex_markers", "tets", "tet_nbrs", "tet_markers", "tris",
                              "tri2tet", "tri_norms", "quads", "quad2hex", "edgemax", "voro_off", "voro_list", "surf_atoms")}
    blob.update(pic_pos=p0, pic_vel=v0, pic_cells=c0, pic_box=box, pic_dt=np.array([dt, qm]))
    _write_blob(tmp_path / "m.bin", blob)
    r = subprocess.run([build_driver(), str(tmp_path / "m.bin"), str(tmp_path / "o.bin"), str(E0)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = _read_blob(tmp_path / "o.bin")
    o.setup(-E0, 0.0, False); o.assemble(True)
    assert o.solve(10000, 1e-11, 1.2, 0) > 0
    o.extract_solution(True)
    cells, sol = o.locate_interpolate(2, 1, m["surf_atoms"])
    rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
    assert out["ncg"][0] > 0 and out["ncg"][1] == 0 and out["ncg"][2] > 0
    assert rel(out["phi_vertex"], o.export_solution()) < 1e-8
    assert np.array_equal(out["markers"], cells)                       # bit-exact cell indices
    assert rel(out["E"].reshape(-1, 3), sol[:, :3]) < 1e-8
    assert rel(out["phi"], sol[:, 4]) < 1e-8                           # upper-case label overwrote the 7.0 fill
    assert rel(out["Enorm"], np.sqrt((sol[:, :3] ** 2).sum(1))) < 1e-8
    assert np.array_equal(out["flag"], (cells >= 0).astype(np.int32))
    assert abs(out["stat"][2] - np.sqrt((sol[:, :3] ** 2).sum(1)).max()) < 1e-8 * out["stat"][2]
    # the PIC push through the C++ Pic seam == the oracle restatement (oracle/pic.py) on the oracle's nodal field
    assert len(c0) > 100
    p1, v1, c1, lost = opic.update_positions(o, p0, v0, c0, dt, box, True)
    v2 = opic.update_velocities(o, p1, v1, c1, dt, qm)
    assert out["pic_lost"][0] == lost and np.array_equal(out["pic_cells"], c1)
    assert np.array_equal(out["pic_pos"].reshape(-1, 3), p1)
    # dv = -grad(phi) dt q/m: a difference of nodal potentials over one cell, so the 1e-8 agreement of phi is amplified
    assert rel(out["pic_vel"].reshape(-1, 3), v2) < 1e-6
