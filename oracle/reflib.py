"""TEST INFRASTRUCTURE -- ctypes view of oracle/_ref/libfemocs_ref.so (the reference's own
mesher + interpolator compiled from /root/reference by oracle/Makefile.ref).

Only tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg and the fixture
generator oracle/make_golden.py may import this module.  The product
(femocs_b200/) never does.

The scenario presets restate the *configuration text* the reference's demo driver
writes (src/main/Main.cpp:25-50 write_defaults, :93-101 mdsmall, :103-112 mdbig,
:173-179 extend, :216-220 tip110, :261-284 read_mesh); no reference code is copied.
"""
import ctypes as C
import os
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libfemocs_ref.so")
REF_ROOT = os.environ.get("FEMOCS_REFERENCE", "/root/reference")

DEFAULTS = """
project = runaway
mesh_quality = 1.8
heat_mode = none
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
force_mode = all
coarse_rate = 0.5
field_mode = transient
seed = 12345
"""

PRESETS = {
    # Main.cpp:93-101
    "mdsmall": dict(infile="in/nanotip_small.xyz", text="""
coarse_factor = 0.3 4 2
radius = 16.0
box_width = 5.0
box_height = 5.0
field_mode = laplace
heat_mode = none
"""),
    # Main.cpp:103-112
    "mdbig": dict(infile="in/nanotip_big.xyz", text="""
coarse_factor = 0.3 4 2
radius = 16.0
box_width = 10.0
box_height = 5.0
field_mode = transient
pic_dtmax = 0.51
heat_mode = none
"""),
    # Main.cpp:216-220
    "tip110": dict(infile="in/tip110.ckx", text="""
coarse_factor = 0.4 8 3
radius = 45.0
"""),
    # Main.cpp:173-179 with extended_atoms = in/extension_90nm.xyz
    "extend90": dict(infile="in/apex.ckx", text="""
extended_atoms = %(ref)s/in/extension_90nm.xyz
coarse_factor = 0.3 6 4
femocs_periodic = false
radius = 70.0
"""),
    # Main.cpp:261-284
    "hemicone": dict(infile=None, text="""
mesh_file = %(ref)s/in/hemicone.msh
radius = 10
box_width = 10.0
box_height = 10.0
bulk_height = 10.0
anode_BC = neumann
elfield = -0.2
heat_mode = none
force_mode = none
field_mode = transient
"""),
}


def available():
    return os.path.exists(LIB_PATH)


def reference_inputs_available():
    return os.path.isdir(os.path.join(REF_ROOT, "in"))


_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


class RefLib:
    """One reference session (the reference keeps process-wide globals: one at a time)."""

    def __init__(self):
        if not available():
            raise RuntimeError("oracle/_ref/libfemocs_ref.so missing: run make -f oracle/Makefile.ref")
        self.lib = C.CDLL(LIB_PATH)
        L = self.lib
        L.ref_init.argtypes = [C.c_char_p]
        L.ref_import_atoms.argtypes = [C.c_char_p]
        L.ref_get_atoms.argtypes = [_dp]
        L.ref_get_surface_atoms.argtypes = [_dp, _ip]
        L.ref_mesh_sizes.argtypes = [_ip]
        L.ref_get_nodes.argtypes = [_dp, _ip]
        L.ref_get_tets.argtypes = [_ip, _ip, _ip]
        L.ref_get_hexs.argtypes = [_ip, _ip]
        L.ref_get_tris.argtypes = [_ip, _ip, _dp]
        L.ref_get_quads.argtypes = [_ip, _ip]
        L.ref_get_stats.argtypes = [_dp]
        L.ref_voro_nbors.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.ref_get_maps.argtypes = [_ip, _ip]
        L.ref_set_nodal.argtypes = [_dp]
        L.ref_get_nodal.argtypes = [_dp]
        L.ref_extract_solution.argtypes = [_dp, _dp, C.c_int, C.c_int]
        L.ref_locate_interpolate.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _ip, _dp]
        L.ref_interp_known_cells.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _ip, _dp]
        L.ref_particle_cells.argtypes = [C.c_int, _dp, _ip]
        L.ref_linhex_locate.argtypes = [C.c_int, _dp, _ip]
        L.ref_particle_field.argtypes = [C.c_int, _dp, _ip, _dp]
        L.ref_particle_weights.argtypes = [C.c_int, _dp, _ip, _dp]
        L.ref_nodal_gradient.argtypes = [C.c_int, _ip, _ip, _dp]

    # ---- mesh generation through the reference's own pipeline -------------------------
    def generate(self, preset):
        p = PRESETS[preset]
        text = DEFAULTS + p["text"] % dict(ref=REF_ROOT)
        with tempfile.NamedTemporaryFile("w", suffix=".in", delete=False) as f:
            f.write(text)
            conf = f.name
        cwd = os.getcwd()
        work = tempfile.mkdtemp(prefix="femocs_ref_")
        os.chdir(work)  # the reference writes out/ and in/ relative to cwd
        try:
            self.lib.ref_init(conf.encode())
            if p["infile"]:
                self.lib.ref_import_atoms(os.path.join(REF_ROOT, p["infile"]).encode())
            fail = self.lib.ref_generate_mesh()
        finally:
            os.chdir(cwd)
            os.unlink(conf)
        if fail:
            raise RuntimeError("reference mesh generation failed for " + preset)
        self.lib.ref_interp_init()
        return self.mesh()

    def mesh(self):
        L = self.lib
        sz = np.zeros(5, np.int32)
        L.ref_mesh_sizes(sz)
        nn, nt, nh, ntri, nq = [int(v) for v in sz]
        m = {}
        m["nodes"] = np.zeros((nn, 3)); m["node_markers"] = np.zeros(nn, np.int32)
        L.ref_get_nodes(m["nodes"].reshape(-1), m["node_markers"])
        m["tets"] = np.zeros((nt, 4), np.int32); m["tet_nbrs"] = np.zeros((nt, 4), np.int32)
        m["tet_markers"] = np.zeros(nt, np.int32)
        L.ref_get_tets(m["tets"].reshape(-1), m["tet_nbrs"].reshape(-1), m["tet_markers"])
        m["hexs"] = np.zeros((nh, 8), np.int32); m["hex_markers"] = np.zeros(nh, np.int32)
        L.ref_get_hexs(m["hexs"].reshape(-1), m["hex_markers"])
        m["tris"] = np.zeros((ntri, 3), np.int32); m["tri2tet"] = np.zeros((ntri, 2), np.int32)
        m["tri_norms"] = np.zeros((ntri, 3))
        L.ref_get_tris(m["tris"].reshape(-1), m["tri2tet"].reshape(-1), m["tri_norms"].reshape(-1))
        m["quads"] = np.zeros((nq, 4), np.int32); m["quad2hex"] = np.zeros((nq, 2), np.int32)
        L.ref_get_quads(m["quads"].reshape(-1), m["quad2hex"].reshape(-1))
        st = np.zeros(2)
        L.ref_get_stats(st)
        m["edgemax"] = st
        tot = L.ref_voro_total()
        ncell = L.ref_voro_nbors(1, None, None)  # NULL args: returns the number of cells
        off = np.zeros(ncell + 1, np.int32); lst = np.zeros(max(tot, 1), np.int32)
        L.ref_voro_nbors(1, off.ctypes.data, lst.ctypes.data)
        m["voro_off"] = off; m["voro_list"] = lst[:tot]
        n2d = np.zeros(nn, np.int32); h2d = np.zeros(nh, np.int32)
        L.ref_get_maps(n2d, h2d)
        m["node_femocs2deal"] = n2d; m["hex_femocs2deal"] = h2d
        na = L.ref_n_atoms()
        if na > 0:
            m["atoms"] = np.zeros((na, 3)); L.ref_get_atoms(m["atoms"].reshape(-1))
        ns = L.ref_n_surface_atoms()
        if ns > 0:
            m["surf_atoms"] = np.zeros((ns, 3)); m["surf_ids"] = np.zeros(ns, np.int32)
            L.ref_get_surface_atoms(m["surf_atoms"].reshape(-1), m["surf_ids"])
        return m

    # ---- interpolator --------------------------------------------------------------
    def set_nodal(self, sol5):
        self.lib.ref_set_nodal(np.ascontiguousarray(sol5, np.float64).reshape(-1))

    def get_nodal(self, n_nodes):
        out = np.zeros((n_nodes, 5))
        self.lib.ref_get_nodal(out.reshape(-1))
        return out

    def extract_solution(self, phi_vertex, rho_vertex, smoothen, n_nodes):
        phi = np.ascontiguousarray(phi_vertex, np.float64)
        rho = np.ascontiguousarray(rho_vertex, np.float64)
        self.lib.ref_extract_solution(phi, rho, len(phi), int(smoothen))
        return self.get_nodal(n_nodes)

    def locate_interpolate(self, dim, rank, xyz):
        xyz = np.ascontiguousarray(xyz, np.float64)
        n = len(xyz)
        cells = np.zeros(n, np.int32); sol = np.zeros((n, 5))
        self.lib.ref_locate_interpolate(dim, rank, n, xyz.reshape(-1), cells, sol.reshape(-1))
        return cells, sol

    def interp_known_cells(self, dim, rank, xyz, cells):
        xyz = np.ascontiguousarray(xyz, np.float64)
        n = len(xyz)
        sol = np.zeros((n, 5))
        self.lib.ref_interp_known_cells(dim, rank, n, xyz.reshape(-1), np.ascontiguousarray(cells, np.int32), sol.reshape(-1))
        return sol

    def export_results(self, ids, sol5, n_points, label, data):
        """the reference's own SolutionReader::export_results (label-case append / overwrite rule) on `data` in place"""
        ids = np.ascontiguousarray(ids, np.int32); sol5 = np.ascontiguousarray(sol5, np.float64)
        assert data.dtype == np.float64 and data.flags.c_contiguous
        self.lib.ref_export_results.restype = C.c_int
        return self.lib.ref_export_results(len(ids), ids.ctypes.data_as(C.c_void_p), sol5.ctypes.data_as(C.c_void_p), int(n_points),
                                           label.encode(), data.ctypes.data_as(C.c_void_p))

    def periodic_image(self, p, pmax, pmin):
        p = np.ascontiguousarray(p, np.float64); out = np.zeros_like(p)
        self.lib.ref_periodic_image.argtypes = [C.c_int, C.c_void_p, C.c_double, C.c_double, C.c_void_p]
        self.lib.ref_periodic_image(len(p), p.ctypes.data_as(C.c_void_p), float(pmax), float(pmin), out.ctypes.data_as(C.c_void_p))
        return out

    def clear_lost(self, pos, vel, cells):
        """ParticleSpecies::clear_lost of the reference: returns (pos, vel, cells, n_lost) of the survivors"""
        pos = np.array(pos, np.float64, copy=True); vel = np.array(vel, np.float64, copy=True); cells = np.array(cells, np.int32, copy=True)
        self.lib.ref_clear_lost.restype = C.c_int
        lost = self.lib.ref_clear_lost(len(cells), pos.ctypes.data_as(C.c_void_p), vel.ctypes.data_as(C.c_void_p), cells.ctypes.data_as(C.c_void_p))
        k = len(cells) - lost
        return pos[:k], vel[:k], cells[:k], lost

    def particle_cells(self, xyz, guess):
        xyz = np.ascontiguousarray(xyz, np.float64)
        cells = np.ascontiguousarray(guess, np.int32).copy()
        self.lib.ref_particle_cells(len(xyz), xyz.reshape(-1), cells)
        return cells

    def linhex_locate(self, xyz, guess):
        xyz = np.ascontiguousarray(xyz, np.float64)
        cells = np.ascontiguousarray(guess, np.int32).copy()
        self.lib.ref_linhex_locate(len(xyz), xyz.reshape(-1), cells)
        return cells

    def particle_field(self, xyz, deal_cells):
        xyz = np.ascontiguousarray(xyz, np.float64)
        E = np.zeros((len(xyz), 3))
        self.lib.ref_particle_field(len(xyz), xyz.reshape(-1), np.ascontiguousarray(deal_cells, np.int32), E.reshape(-1))
        return E

    def particle_weights(self, xyz, deal_cells):
        xyz = np.ascontiguousarray(xyz, np.float64)
        w = np.zeros((len(xyz), 8))
        self.lib.ref_particle_weights(len(xyz), xyz.reshape(-1), np.ascontiguousarray(deal_cells, np.int32), w.reshape(-1))
        return w

    def nodal_gradient(self, hexs, nodes):
        hexs = np.ascontiguousarray(hexs, np.int32); nodes = np.ascontiguousarray(nodes, np.int32)
        E = np.zeros((len(hexs), 3))
        self.lib.ref_nodal_gradient(len(hexs), hexs, nodes, E.reshape(-1))
        return E
