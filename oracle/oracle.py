"""TEST INFRASTRUCTURE -- ctypes view of the CPU oracle (oracle/femocs_oracle.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this.  The product (femocs_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_PATH = os.path.join(HERE, "_build", "libfemocs_oracle.so")

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_lp = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


def build():
    src = os.path.join(HERE, "femocs_oracle.cpp")
    if (not os.path.exists(LIB_PATH)) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-f", "oracle/Makefile"], cwd=ROOT)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.fo_create.restype = C.c_void_p
        L.fo_destroy.argtypes = [C.c_void_p]
        L.fo_set_num_threads.argtypes = [C.c_int]
        L.fo_get_max_threads.restype = C.c_int
        L.fo_import_mesh.argtypes = [C.c_void_p, _dp, C.c_int, _ip, _ip, C.c_int]
        L.fo_set_fe_degree.argtypes = [C.c_void_p, C.c_int]
        L.fo_get_cell_dofs27.argtypes = [C.c_void_p, _ip]
        L.fo_setup.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int]
        L.fo_assemble.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_long, C.c_double]
        L.fo_solve.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int, C.POINTER(C.c_double)]
        L.fo_sizes.argtypes = [C.c_void_p, _lp]
        L.fo_get_csr.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.fo_get_vectors.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.fo_set_solution.argtypes = [C.c_void_p, _dp]
        L.fo_get_bfaces.argtypes = [C.c_void_p, _ip, _ip, _ip]
        L.fo_get_cells.argtypes = [C.c_void_p, _ip]
        L.fo_export_solution.argtypes = [C.c_void_p, _dp]
        L.fo_export_charge_dens.argtypes = [C.c_void_p, _dp]
        L.fo_set_write_time.argtypes = [C.c_void_p, C.c_int]
        L.fo_export_solution_grad.argtypes = [C.c_void_p, _dp]
        L.fo_mesh_counts.argtypes = [C.c_void_p, C.POINTER(C.c_long), C.POINTER(C.c_long)]
        L.fo_check_limits.argtypes = [C.c_void_p, C.c_double, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.fo_cell_vol.argtypes = [C.c_void_p, C.c_int]
        L.fo_cell_vol.restype = C.c_double
        L.fo_interp_initialize.argtypes = [C.c_void_p, _ip, _ip, _ip, _ip, C.c_int, _ip, _ip, _dp, C.c_int,
                                           _ip, _ip, C.c_int, C.c_double, _ip, _ip, C.c_int]
        L.fo_extract_solution.argtypes = [C.c_void_p, C.c_int]
        L.fo_set_nodal.argtypes = [C.c_void_p, _dp]
        L.fo_get_nodal.argtypes = [C.c_void_p, _dp]
        L.fo_locate_interpolate.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_long, _dp, _ip, _dp]
        L.fo_interpolate.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_long, _dp, _ip, _dp]
        L.fo_particle_cells.argtypes = [C.c_void_p, C.c_long, _dp, _ip]
        L.fo_particle_field.argtypes = [C.c_void_p, C.c_long, _dp, _ip, _dp]
        L.fo_particle_weights.argtypes = [C.c_void_p, C.c_long, _dp, _ip, _dp]
        L.fo_linhex_locate.argtypes = [C.c_void_p, C.c_long, _dp, _ip]
        L.fo_nodal_gradient.argtypes = [C.c_void_p, C.c_long, _ip, _ip, _dp]
        L.fo_import_mesh_kind.argtypes = [C.c_void_p, _dp, C.c_int, _ip, _ip, C.c_int, C.c_int]
        L.fo_ch_set_physics.argtypes = [C.c_void_p, _dp, _dp, C.c_int, C.c_double]
        L.fo_ch_sigma.argtypes = [C.c_void_p, C.c_double]; L.fo_ch_sigma.restype = C.c_double
        L.fo_ch_kappa.argtypes = [C.c_void_p, C.c_double]; L.fo_ch_kappa.restype = C.c_double
        L.fo_ch_setup.argtypes = [C.c_void_p, C.c_double]
        L.fo_current_assemble.argtypes = [C.c_void_p, _dp]
        L.fo_heat_assemble.argtypes = [C.c_void_p, C.c_double, _dp]
        L.fo_ch_solve.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.POINTER(C.c_double)]
        L.fo_ch_get_solution.argtypes = [C.c_void_p, C.c_int, _dp]
        L.fo_ch_set_solution.argtypes = [C.c_void_p, C.c_int, _dp]
        L.fo_ch_select.argtypes = [C.c_void_p, C.c_int]
        L.fo_surface_centroids.argtypes = [C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _f(a):
    return np.ascontiguousarray(a, np.float64)


def _i(a):
    return np.ascontiguousarray(a, np.int32)


class Oracle:
    """CPU restatement of PoissonSolver<3> + Interpolator + SolutionReader loops."""

    def __init__(self):
        self.L = lib()
        self.h = C.c_void_p(self.L.fo_create())
        self.n_nodes = 0

    def __del__(self):
        try:
            self.L.fo_destroy(self.h)
        except Exception:
            pass

    # ---- solver ------------------------------------------------------------------
    def set_fe_degree(self, degree):
        """1 = FE_Q(1) (the reference build), 2 = FE_Q(2) (DealSolver.h:130 shape_degree = 2); read by the next import_mesh"""
        self.L.fo_set_fe_degree(self.h, int(degree))

    def cell_dofs27(self):
        out = np.zeros((self.n_cells, 27), np.int32)
        self.L.fo_get_cell_dofs27(self.h, out.reshape(-1))
        return out

    def import_mesh(self, nodes, hexs, hex_markers):
        nodes = _f(nodes); hexs = _i(hexs); hex_markers = _i(hex_markers)
        self.n_nodes = len(nodes)
        rc = self.L.fo_import_mesh(self.h, nodes.reshape(-1), len(nodes), hexs.reshape(-1), hex_markers, len(hexs))
        if rc:
            raise RuntimeError("oracle import_mesh failed rc=%d" % rc)
        sz = np.zeros(5, np.int64)
        self.L.fo_sizes(self.h, sz)
        self.n_dofs, self.n_cells, self.nnz, self.n_vertices, self.n_bfaces = [int(v) for v in sz]

    # ---- CurrentHeatSolver on the bulk hexahedra (CurrentHeatSolver.cpp) -----------
    def import_bulk_mesh(self, nodes, hexs, hex_markers):
        nodes = _f(nodes); hexs = _i(hexs); hex_markers = _i(hex_markers)
        self.n_nodes = len(nodes)
        rc = self.L.fo_import_mesh_kind(self.h, nodes.reshape(-1), len(nodes), hexs.reshape(-1), hex_markers, len(hexs), 1)
        if rc:
            raise RuntimeError("oracle import_mesh (bulk) failed rc=%d" % rc)
        sz = np.zeros(5, np.int64)
        self.L.fo_sizes(self.h, sz)
        self.n_dofs, self.n_cells, self.nnz, self.n_vertices, self.n_bfaces = [int(v) for v in sz]

    def ch_set_physics(self, T, rho, lorentz=2.44e-8):
        T = _f(T); rho = _f(rho)
        self.L.fo_ch_set_physics(self.h, T, rho, len(T), lorentz)

    def ch_setup(self, T_ambient):
        self.L.fo_ch_setup(self.h, T_ambient)

    def surface_centroids(self):
        n = self.L.fo_surface_centroids(self.h, None)
        out = np.zeros((n, 3))
        self.L.fo_surface_centroids(self.h, out.ctypes.data)
        return out

    def current_assemble(self, face_bc):
        self.L.fo_current_assemble(self.h, _f(face_bc))

    def heat_assemble(self, delta_time, face_bc):
        self.L.fo_heat_assemble(self.h, delta_time, _f(face_bc))

    def ch_solve(self, which, max_iter=2000, tol=1e-9, ssor=1.2, precond=0):
        res = C.c_double(0)
        it = self.L.fo_ch_solve(self.h, which, max_iter, tol, ssor, precond, C.byref(res))
        self.last_res = res.value
        return it

    def ch_solution(self, which):
        out = np.zeros(self.n_dofs)
        self.L.fo_ch_get_solution(self.h, which, out)
        return out

    def ch_set_solution(self, which, v):
        self.L.fo_ch_set_solution(self.h, which, _f(v))

    def ch_select(self, which):
        self.L.fo_ch_select(self.h, which)

    def sigma(self, T):
        return self.L.fo_ch_sigma(self.h, T)

    def kappa(self, T):
        return self.L.fo_ch_kappa(self.h, T)

    def setup(self, field, potential=0.0, anode_dirichlet=False):
        self.L.fo_setup(self.h, field, potential, int(anode_dirichlet))

    def assemble(self, first_time=True, particles_xyz=None, particle_cells=None, charge_factor=0.0):
        if particles_xyz is None:
            self.L.fo_assemble(self.h, int(first_time), None, None, 0, 0.0)
        else:
            p = _f(particles_xyz); c = _i(particle_cells)
            self.L.fo_assemble(self.h, int(first_time), p.ctypes.data, c.ctypes.data, len(c), charge_factor)

    def solve(self, max_iter=10000, tol=1e-9, ssor=1.2, precond=0):
        res = C.c_double(0)
        it = self.L.fo_solve(self.h, max_iter, tol, ssor, precond, C.byref(res))
        self.last_res = res.value
        return it

    def csr(self):
        rowptr = np.zeros(self.n_dofs + 1, np.int32); col = np.zeros(self.nnz, np.int32)
        val = np.zeros(self.nnz); save = np.zeros(self.nnz)
        self.L.fo_get_csr(self.h, rowptr.ctypes.data, col.ctypes.data, val.ctypes.data, save.ctypes.data)
        return rowptr, col, val, save

    def vectors(self):
        rhs = np.zeros(self.n_dofs); sol = np.zeros(self.n_dofs)
        v2d = np.zeros(self.n_vertices, np.int32); v2n = np.zeros(self.n_vertices, np.int32)
        self.L.fo_get_vectors(self.h, rhs.ctypes.data, sol.ctypes.data, v2d.ctypes.data, v2n.ctypes.data)
        return rhs, sol, v2d, v2n

    def set_solution(self, sol_dof):
        self.L.fo_set_solution(self.h, _f(sol_dof))

    def bfaces(self):
        c = np.zeros(self.n_bfaces, np.int32); f = np.zeros(self.n_bfaces, np.int32); i = np.zeros(self.n_bfaces, np.int32)
        self.L.fo_get_bfaces(self.h, c, f, i)
        return c, f, i

    def cells(self):
        c = np.zeros((self.n_cells, 8), np.int32)
        self.L.fo_get_cells(self.h, c.reshape(-1))
        return c

    def export_solution(self):
        phi = np.zeros(self.n_vertices)
        self.L.fo_export_solution(self.h, phi)
        return phi

    def export_solution_grad(self):
        g = np.zeros((self.n_vertices, 3))
        self.L.fo_export_solution_grad(self.h, g.reshape(-1))
        return g

    def mesh_counts(self):
        a = C.c_long(0); b = C.c_long(0)
        self.L.fo_mesh_counts(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def set_write_time(self, on):
        """FileWriter::write_time() of the solver: the next assemble keeps charge_density = rhs / dof_volume"""
        self.L.fo_set_write_time(self.h, int(on))

    def export_charge_dens(self):
        rho = np.zeros(self.n_vertices)
        self.L.fo_export_charge_dens(self.h, rho)
        return rho

    def check_limits(self, lo, hi):
        a = C.c_double(0); b = C.c_double(0)
        bad = self.L.fo_check_limits(self.h, lo, hi, C.byref(a), C.byref(b))
        return bool(bad), a.value, b.value

    def cell_vol(self, c):
        return self.L.fo_cell_vol(self.h, c)

    # ---- interpolator --------------------------------------------------------------
    def interp_initialize(self, m):
        self.L.fo_interp_initialize(
            self.h, _i(m["node_markers"]), _i(m["tets"]).reshape(-1), _i(m["tet_nbrs"]).reshape(-1), _i(m["tet_markers"]),
            len(m["tets"]), _i(m["tris"]).reshape(-1), _i(m["tri2tet"]).reshape(-1), _f(m["tri_norms"]).reshape(-1),
            len(m["tris"]), _i(m["quads"]).reshape(-1), _i(m["quad2hex"]).reshape(-1), len(m["quads"]),
            float(m["edgemax"][0]), _i(m["voro_off"]), _i(m["voro_list"]) if len(m["voro_list"]) else np.zeros(1, np.int32),
            len(m["voro_off"]) - 1)
        self.n_tet, self.n_tri, self.n_hex = len(m["tets"]), len(m["tris"]), 4 * len(m["tets"])

    def tables(self):
        """precomputed cell tables in the product's record layouts (see fo_get_tables)"""
        nt, nh, nr = self.n_tet, self.n_hex, self.n_tri
        t = dict(tet=np.zeros((nt, 17)), tet_cent=np.zeros((nt, 3)), tet_mark=np.zeros(nt, np.int32), hex=np.zeros((nh, 24)),
                 tri=np.zeros((max(nr, 1), 16)), tri_cent=np.zeros((max(nr, 1), 3)), qtet=np.zeros((nt, 10), np.int32),
                 qtri=np.zeros((max(nr, 1), 6), np.int32))
        self.L.fo_get_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 8
        self.L.fo_get_tables(self.h, *[t[k].ctypes.data_as(C.c_void_p) for k in ("tet", "tet_cent", "tet_mark", "hex", "tri", "tri_cent", "qtet", "qtri")])
        return t

    def extract_solution(self, smoothen=False):
        self.L.fo_extract_solution(self.h, int(smoothen))
        return self.get_nodal()

    def set_nodal(self, sol5):
        self.L.fo_set_nodal(self.h, _f(sol5).reshape(-1))

    def get_nodal(self):
        out = np.zeros((self.n_nodes, 5))
        self.L.fo_get_nodal(self.h, out.reshape(-1))
        return out

    def locate_interpolate(self, dim, rank, xyz):
        xyz = _f(xyz); n = len(xyz)
        cells = np.zeros(n, np.int32); sol = np.zeros((n, 5))
        self.L.fo_locate_interpolate(self.h, dim, rank, n, xyz.reshape(-1), cells, sol.reshape(-1))
        return cells, sol

    def interpolate(self, dim, rank, xyz, cells):
        xyz = _f(xyz); n = len(xyz)
        sol = np.zeros((n, 5))
        self.L.fo_interpolate(self.h, dim, rank, n, xyz.reshape(-1), _i(cells), sol.reshape(-1))
        return sol

    def particle_cells(self, xyz, guess):
        xyz = _f(xyz); cells = _i(guess).copy()
        self.L.fo_particle_cells(self.h, len(xyz), xyz.reshape(-1), cells)
        return cells

    def particle_field(self, xyz, cells):
        xyz = _f(xyz); E = np.zeros((len(xyz), 3))
        self.L.fo_particle_field(self.h, len(xyz), xyz.reshape(-1), _i(cells), E.reshape(-1))
        return E

    def particle_weights(self, xyz, cells):
        xyz = _f(xyz); w = np.zeros((len(xyz), 8))
        self.L.fo_particle_weights(self.h, len(xyz), xyz.reshape(-1), _i(cells), w.reshape(-1))
        return w

    def linhex_locate(self, xyz, guess):
        xyz = _f(xyz); cells = _i(guess).copy()
        self.L.fo_linhex_locate(self.h, len(xyz), xyz.reshape(-1), cells)
        return cells

    def nodal_gradient(self, hexs, nodes):
        hexs = _i(hexs); nodes = _i(nodes); E = np.zeros((len(hexs), 3))
        self.L.fo_nodal_gradient(self.h, len(hexs), hexs, nodes, E.reshape(-1))
        return E
