"""Host-side mirror of the reference's operator interface for the electrostatic hot path.

Class and method names follow the reference (include/PoissonSolver.h, include/DealSolver.h,
include/Interpolator.h, include/SolutionReader.h, include/Pic.h) so that the parity tests
read like the reference's call sites in src/ProjectRunaway.cpp:

    solver = PoissonSolver(ctx, conf)                  # ProjectRunaway.cpp:38
    solver.import_mesh(nodes, hexs, hex_markers)       # :216
    solver.setup(-E0, V0); solver.assemble(True)       # :424-425
    ncg = solver.solve()                               # :431   (+#CG / -#CG)
    interp = Interpolator(ctx); interp.initialize(mesh)          # :435
    interp.extract_solution(solver, smoothen)                    # :436
    fields = FieldReader(interp); fields.set_preferences(False, 2, 1)
    fields.interpolate(points)                                   # :303-304

Everything numerical happens in libfemocs_b200.so (CUDA, sm_100a) behind the C ABI of
include/femocs_b200.h; this module only marshals numpy arrays.  No CPU fallback exists.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import lib as _lib

PRECOND_JACOBI = 1
PRECOND_CHEBYSHEV = 2
PRECOND_TWOLEVEL = 3


class FemocsB200Error(RuntimeError):
    pass


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return None if a is None else a.ctypes.data


class Context:
    """One CUDA device context (fb_ctx): owns the stream, the mesh and all device arrays."""

    def __init__(self, device=0):
        self.L = _lib.load()
        h = self.L.fb_create(device)
        if not h:
            raise FemocsB200Error("fb_create failed: " + self.L.fb_create_error().decode())
        self.h = C.c_void_p(h)

    def close(self):
        if getattr(self, "h", None):
            self.L.fb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc != 0:
            raise FemocsB200Error("libfemocs_b200 error %d: %s" % (rc, self.L.fb_last_error(self.h).decode()))

    def set_option(self, key, value):
        self.check(self.L.fb_set_option(self.h, key.encode(), float(value)))

    @property
    def kernel_launches(self):
        return int(self.L.fb_kernel_launches(self.h))

    def synchronize(self):
        self.check(self.L.fb_synchronize(self.h))

    # ---- multi-GPU: one process per GPU; call before PoissonSolver.import_mesh ----
    def init_comm(self, rank, world, unique_id):
        """unique_id: the 128 bytes drawn by rank 0 with Context.unique_id() and broadcast by the host code."""
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self.check(self.L.fb_comm_init(self.h, int(rank), int(world), C.cast(buf, C.c_void_p)))

    @staticmethod
    def unique_id():
        L = _lib.load()
        buf = C.create_string_buffer(128)
        if L.fb_comm_unique_id(C.cast(buf, C.c_void_p)) != 0:
            raise FemocsB200Error("fb_comm_unique_id failed: " + L.fb_create_error().decode())
        return buf.raw

    def init_comm_torch(self, dist):
        """Convenience for torchrun launches: broadcast the id over an initialised torch.distributed group."""
        import torch
        rank, world = dist.get_rank(), dist.get_world_size()
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        t = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            t = torch.frombuffer(bytearray(self.unique_id()), dtype=torch.uint8).to(dev)
        dist.broadcast(t, 0)
        self.init_comm(rank, world, bytes(t.cpu().numpy().tobytes()))

    def partition(self):
        out = np.zeros(10, np.int64)
        self.check(self.L.fb_get_partition(self.h, _p(out)))
        keys = ("rank", "world", "n_rows", "n_cols", "nnz", "n_cells", "n_send", "n_ghost", "n_vert_global", "n_cells_global")
        return dict(zip(keys, (int(v) for v in out)))

    @property
    def last_import_reused(self):
        """True when the last import_mesh found the connectivity unchanged and only refreshed the geometry"""
        return bool(self.L.fb_last_import_reused(self.h))

    @property
    def comm_mode(self):
        """0 = one GPU, 1 = NCCL inside the CG iteration, 2 = peer-mapped iteration (fb_comm_mode)"""
        return int(self.L.fb_comm_mode(self.h))

    @property
    def stream(self):
        """cudaStream_t of the context (integer address)."""
        return int(self.L.fb_get_stream(self.h))


@dataclass
class FieldConfig:
    """Subset of Config::Field used on the hot path (defaults: src/Config.cpp:63-71)."""
    E0: float = 0.0
    V0: float = 0.0
    ssor_param: float = 1.2         # accepted for config compatibility; the GPU path uses Jacobi
    cg_tolerance: float = 1e-9
    n_cg: int = 10000
    anode_BC: str = "neumann"
    mode: str = "laplace"
    V_min: float = -1.0
    V_max: float = 1e4
    precond: int = PRECOND_JACOBI


class PoissonSolver:
    """femocs::PoissonSolver<3> (+ DealSolver<3>) behind libfemocs_b200."""

    def __init__(self, ctx, conf=None):
        self.ctx = ctx
        self.conf = conf or FieldConfig()
        self.stat_sol_min = 0.0
        self.stat_sol_max = 0.0
        self.last_residual = 0.0
        self._particles = None

    # DealSolver::import_mesh (src/DealSolver.cpp:191-209); returns True on success like the reference
    def import_mesh(self, nodes, hexs, hex_markers):
        nodes = _f(nodes); hexs = _i(hexs); hex_markers = _i(hex_markers)
        rc = self.ctx.L.fb_import_mesh(self.ctx.h, _p(nodes), len(nodes), _p(hexs), _p(hex_markers), len(hexs))
        if rc == 3:
            return False
        self.ctx.check(rc)
        sz = np.zeros(7, np.int64)
        self.ctx.check(self.ctx.L.fb_get_sizes(self.ctx.h, _p(sz)))
        (self.n_dofs, self.n_cells, self.nnz, self.n_vertices, self.n_bfaces, self.n_top_faces, _) = [int(v) for v in sz]
        part = self.ctx.partition()
        if part["world"] > 1:           # partitioned: n_dofs / nnz / n_cells above are this rank's share
            self.n_vertices = part["n_vert_global"]
            self.n_dofs_global = part["n_vert_global"]
        else:
            self.n_dofs_global = self.n_dofs
        return True

    def size(self):
        return self.n_dofs

    def get_n_cells(self):
        return self.n_cells

    # PoissonSolver::set_particles (include/PoissonSolver.h:29): (xyz[n,3], solver cell ids[n], charge factor)
    def set_particles(self, xyz, cells, charge_factor):
        self._particles = None if xyz is None else (_f(xyz), _i(cells), float(charge_factor))

    # PoissonSolver::setup (src/PoissonSolver.cpp:162-167)
    def setup(self, field, potential=0.0):
        self.ctx.check(self.ctx.L.fb_poisson_setup(self.ctx.h, float(field), float(potential),
                                                   int(self.conf.anode_BC.lower() == "dirichlet")))

    # PoissonSolver::assemble (src/PoissonSolver.cpp:170-210)
    def assemble(self, first_time=True):
        if self.conf.mode != "laplace" and self._particles is not None and len(self._particles[1]):
            xyz, cells, cf = self._particles
            self.ctx.check(self.ctx.L.fb_poisson_assemble(self.ctx.h, int(first_time), _p(xyz), _p(cells), len(cells), cf))
        else:
            self.ctx.check(self.ctx.L.fb_poisson_assemble(self.ctx.h, int(first_time), None, None, 0, 0.0))

    # same, with the particle arrays already in HBM (raw device addresses; asynchronous on the context stream)
    def assemble_dev(self, first_time, xyz_dev_ptr, cells_dev_ptr, n_particles, charge_factor):
        self.ctx.check(self.ctx.L.fb_poisson_assemble_dev(self.ctx.h, int(first_time), xyz_dev_ptr, cells_dev_ptr,
                                                          int(n_particles), float(charge_factor)))

    # PoissonSolver::solve (include/PoissonSolver.h:54): +#CG on success, -#CG when n_cg was hit
    def solve(self, n_cg=None, cg_tolerance=None):
        it = C.c_int(0); res = C.c_double(0)
        self.ctx.check(self.ctx.L.fb_poisson_solve(
            self.ctx.h, int(self.conf.n_cg if n_cg is None else n_cg),
            float(self.conf.cg_tolerance if cg_tolerance is None else cg_tolerance),
            int(self.conf.precond), C.byref(it), C.byref(res)))
        self.last_residual = res.value
        return it.value

    def solve_stats(self):
        ms = C.c_double(0); it = C.c_int(0); sp = C.c_long(0)
        self.ctx.L.fb_last_solve_stats(self.ctx.h, C.byref(ms), C.byref(it), C.byref(sp))
        return ms.value, it.value, sp.value

    def solve_kernel(self):
        """SpMV kernel code of the last solve (include/femocs_b200.h: fb_last_solve_kernel)"""
        return int(self.ctx.L.fb_last_solve_kernel(self.ctx.h))

    # (avg ms of the SpMV+dot kernel, avg ms of the vector kernels, samples) of the last solve; needs
    # ctx.set_option("cg_profile", k)
    def solve_profile(self):
        a = C.c_double(0); b = C.c_double(0); n = C.c_int(0)
        self.ctx.L.fb_last_solve_profile(self.ctx.h, C.byref(a), C.byref(b), C.byref(n))
        return a.value, b.value, n.value

    # DealSolver::check_limits (src/DealSolver.cpp:157-167)
    def check_limits(self, low_limit, high_limit):
        bad = C.c_int(0); a = C.c_double(0); b = C.c_double(0)
        self.ctx.check(self.ctx.L.fb_check_limits(self.ctx.h, low_limit, high_limit, C.byref(bad), C.byref(a), C.byref(b)))
        self.stat_sol_min, self.stat_sol_max = a.value, b.value
        return bool(bad.value)

    # DealSolver::export_solution / PoissonSolver::export_charge_dens (vertex order)
    def export_solution(self, out=None):
        if out is None:
            out = np.zeros(self.n_vertices)
        assert out.dtype == np.float64 and out.size >= self.n_vertices
        self.ctx.check(self.ctx.L.fb_export_solution(self.ctx.h, _p(out)))
        return out

    def export_charge_dens(self):
        out = np.zeros(self.n_vertices)
        self.ctx.check(self.ctx.L.fb_export_charge_dens(self.ctx.h, _p(out)))
        return out

    # DealSolver::export_solution_grad (src/DealSolver.cpp:280-301): MINUS the gradient at Gauss point vertex2node[v] of
    # cell vertex2cell[v], one triple per solver vertex
    def export_solution_grad(self):
        out = np.zeros((self.n_vertices, 3))
        self.ctx.check(self.ctx.L.fb_export_solution_grad(self.ctx.h, _p(out)))
        return out

    # (#faces, #edges) of the solver mesh, as printed by operator<< (include/DealSolver.h:107-117)
    def mesh_counts(self):
        a = C.c_long(0); b = C.c_long(0)
        self.ctx.check(self.ctx.L.fb_get_mesh_counts(self.ctx.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    # solver vertices and cells (old-style vertex order) as written by write("*.vtk") / write("*.msh")
    def solver_mesh(self):
        xyz = np.zeros((self.n_vertices, 3)); cells = np.zeros((self.n_cells, 8), np.int32)
        self.ctx.check(self.ctx.L.fb_get_solver_mesh(self.ctx.h, _p(xyz), _p(cells)))
        return xyz, cells

    def import_solution(self, phi_vertex):
        phi = _f(phi_vertex)
        assert len(phi) == self.n_vertices
        self.ctx.check(self.ctx.L.fb_import_solution(self.ctx.h, _p(phi)))

    def get_cell_volumes(self):
        out = np.zeros(self.n_cells)
        self.ctx.check(self.ctx.L.fb_get_cell_volumes(self.ctx.h, _p(out)))
        return out

    def get_cell_vol(self, i):
        return float(self.get_cell_volumes()[i])

    def to_str(self):
        nf, ne = self.mesh_counts()
        return "#elems=%d, #faces=%d, #edges=%d, #nodes=%d, #dofs=%d" % (self.n_cells, nf, ne, self.n_vertices, self.n_dofs)

    # test hook: assembled system in DoF numbering
    def get_system(self):
        rowptr = np.zeros(self.n_dofs + 1, np.int32); col = np.zeros(self.nnz, np.int32)
        val = np.zeros(self.nnz); save = np.zeros(self.nnz)
        rhs = np.zeros(self.n_dofs); sol = np.zeros(self.n_dofs)
        v2d = np.zeros(self.n_vertices, np.int32); v2n = np.zeros(self.n_vertices, np.int32)
        self.ctx.check(self.ctx.L.fb_get_system(self.ctx.h, _p(rowptr), _p(col), _p(val), _p(save), _p(rhs), _p(sol), _p(v2d), _p(v2n)))
        return dict(rowptr=rowptr, col=col, val=val, val_save=save, rhs=rhs, sol=sol, vertex2dof=v2d, vertex2node=v2n)


class Interpolator:
    """femocs::Interpolator: nodal solutions + cell tables on the device."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.n_nodes = 0

    # Interpolator::initialize(mesh, 0, TYPES.VACUUM) (src/Interpolator.cpp:28-77); `mesh` is a dict of
    # the TetgenMesh arrays (tests/golden/mesh_*.npz layout)
    def initialize(self, mesh):
        m = mesh
        self.n_nodes = len(m["nodes"])
        a = dict(nm=_i(m["node_markers"]), tets=_i(m["tets"]), nbr=_i(m["tet_nbrs"]), tm=_i(m["tet_markers"]),
                 tris=_i(m["tris"]), t2t=_i(m["tri2tet"]), tn=_f(m["tri_norms"]), quads=_i(m["quads"]),
                 q2h=_i(m["quad2hex"]), voff=_i(m["voro_off"]), vlist=_i(m["voro_list"]) if len(m["voro_list"]) else np.zeros(1, np.int32))
        self.ctx.check(self.ctx.L.fb_interp_initialize(
            self.ctx.h, _p(a["nm"]), _p(a["tets"]), _p(a["nbr"]), _p(a["tm"]), len(a["tets"]),
            _p(a["tris"]), _p(a["t2t"]), _p(a["tn"]), len(a["tris"]), _p(a["quads"]), _p(a["q2h"]), len(a["quads"]),
            float(m["edgemax"][0]), _p(a["voff"]), _p(a["vlist"]), len(a["voff"]) - 1))

    # Interpolator::extract_solution(PoissonSolver<3>&, smoothen) (src/Interpolator.cpp:172-190)
    def extract_solution(self, fem, smoothen=False):
        assert fem.ctx is self.ctx
        self.ctx.check(self.ctx.L.fb_extract_solution(self.ctx.h, int(smoothen)))

    def get_solutions(self):
        out = np.zeros((self.n_nodes, 5))
        self.ctx.check(self.ctx.L.fb_get_nodal_solutions(self.ctx.h, _p(out)))
        return out

    def set_solutions(self, sol5):
        s = _f(sol5)
        assert s.shape == (self.n_nodes, 5)
        self.ctx.check(self.ctx.L.fb_set_nodal_solutions(self.ctx.h, _p(s)))


class SolutionReader:
    """femocs::SolutionReader: points + interpolated Solutions + located cells (atom markers)."""

    def __init__(self, interpolator):
        self.interpolator = interpolator
        self.ctx = interpolator.ctx
        self.dim, self.rank = 3, 1
        self.points = np.zeros((0, 3)); self.ids = np.zeros(0, np.int64)
        self.markers = np.zeros(0, np.int32); self.interpolation = np.zeros((0, 5))
        self.atoms_mapped_to_cells = False

    # SolutionReader::set_preferences (include/SolutionReader.h:60-67); sorting is not used on the hot path
    def set_preferences(self, sort_atoms, dim, rank, centroid=False):
        if dim not in (2, 3) or rank not in (1, 2, 3):
            raise ValueError("invalid interpolation dimension/rank")
        if sort_atoms or centroid:
            raise NotImplementedError("sort_atoms / interp_centroids are outside the hot path")
        self.dim, self.rank = dim, rank

    def size(self):
        return len(self.points)

    def reserve_points(self, xyz, ids=None):
        self.points = _f(xyz).reshape(-1, 3)
        self.ids = np.arange(len(self.points)) if ids is None else np.asarray(ids)
        self.atoms_mapped_to_cells = False

    # SolutionReader::calc_full_interpolation (src/SolutionReader.cpp:136-165)
    def calc_full_interpolation(self):
        n = len(self.points)
        self.markers = np.zeros(n, np.int32); self.interpolation = np.zeros((n, 5))
        if n:
            x = self.points
            self.ctx.check(self.ctx.L.fb_locate_interpolate(
                self.ctx.h, self.dim, self.rank, n, x.ctypes.data, x.ctypes.data + 8, x.ctypes.data + 16, 3,
                _p(self.markers), _p(self.interpolation)))
        self.atoms_mapped_to_cells = True

    # batched EmissionReader::emission_line look-ups (src/EmissionReader.cpp:51-61): `lines` is (n_lines, n_per_line, 3);
    # every line is an independent guess chain (a fresh phis_on_line reader in the reference)
    def interpolate_lines(self, lines):
        lines = _f(lines)
        assert lines.ndim == 3 and lines.shape[2] == 3
        nl, npl = lines.shape[:2]
        cells = np.zeros((nl, npl), np.int32); sol = np.zeros((nl, npl, 5))
        self.ctx.check(self.ctx.L.fb_locate_interpolate_chains(self.ctx.h, self.dim, self.rank, nl, npl, _p(lines), _p(cells), _p(sol)))
        return cells, sol

    # SolutionReader::calc_interpolation (src/SolutionReader.cpp:167-190)
    def calc_interpolation(self):
        if not self.atoms_mapped_to_cells:
            return self.calc_full_interpolation()
        n = len(self.points)
        if n:
            x = self.points
            self.ctx.check(self.ctx.L.fb_interpolate(
                self.ctx.h, self.dim, self.rank, n, x.ctypes.data, x.ctypes.data + 8, x.ctypes.data + 16, 3,
                _p(self.markers), _p(self.interpolation)))

    # SolutionReader::interpolate(n, x, y, z) (src/SolutionReader.cpp:428-436)
    def interpolate(self, xyz, ids=None):
        self.reserve_points(xyz, ids)
        self.calc_interpolation()

    def update_positions(self, xyz):
        self.points = _f(xyz).reshape(-1, 3)

    # SolutionReader::interpolate_results (src/SolutionReader.cpp:405-421): SoA x,y,z as in Femocs_wrap.h:36
    def interpolate_results(self, x, y, z, data_type):
        x = _f(x); y = _f(y); z = _f(z)
        n = len(x)
        cells = np.zeros(n, np.int32); sol = np.zeros((n, 5))
        self.ctx.check(self.ctx.L.fb_locate_interpolate(self.ctx.h, self.dim, self.rank, n, _p(x), _p(y), _p(z), 1,
                                                        _p(cells), _p(sol)))
        tmp = SolutionReader(self.interpolator)
        tmp.points = np.stack([x, y, z], 1); tmp.ids = np.arange(n); tmp.markers = cells; tmp.interpolation = sol
        n_comp = 3 if data_type.lower() in ("elfield", "vec") else 1
        data = np.zeros(n * n_comp)
        tmp.export_results(n, data_type.upper(), data)
        return data, cells

    # SolutionReader::export_results (src/SolutionReader.cpp:303-398): exact-case label appends,
    # upper-case label overwrites; scatter by atom id
    LABELS = {"elfield": 1, "elfield_norm": 2, "charge_density": 3, "potential": 4}

    def export_results(self, n_points, data_type, data):
        if self.size() == 0:
            return 1
        append = data_type in self.LABELS
        kind = self.LABELS.get(data_type.lower())
        if kind is None:
            raise ValueError("SolutionReader does not contain " + data_type)
        ids = np.asarray(self.ids)
        ok = (ids >= 0) & (ids < n_points)
        ids = ids[ok]; sol = self.interpolation[ok]
        if kind == 1:
            if not append:
                data[:3 * n_points] = 0
            view = data[:3 * n_points].reshape(n_points, 3)
            if append:
                np.add.at(view, ids, sol[:, :3])
            else:
                view[ids] = sol[:, :3]
        else:
            vals = {2: np.sqrt((sol[:, :3] ** 2).sum(1)), 3: sol[:, 3], 4: sol[:, 4]}[kind]
            if not append:
                data[:n_points] = 0
                data[:n_points][ids] = vals
            else:
                np.add.at(data[:n_points], ids, vals)
        return 0


class FieldReader(SolutionReader):
    """femocs::FieldReader: adds field norms and E_max (src/SolutionReader.cpp:473-499)."""

    def __init__(self, interpolator):
        super().__init__(interpolator)
        self.field_norm = np.zeros(0); self.E_max = -1e100

    def calc_interpolation(self):
        super().calc_interpolation()
        self.field_norm = np.sqrt((self.interpolation[:, :3] ** 2).sum(1))
        self.E_max = float(self.field_norm.max()) if len(self.field_norm) else -1e100

    def get_elfield(self, i):
        return self.interpolation[i, :3]

    def get_potential(self, i):
        return self.interpolation[i, 4]


class Pic:
    """The hot-path slice of femocs::Pic<3>: particle cell update and field look-up."""

    def __init__(self, interpolator):
        self.ctx = interpolator.ctx

    # Pic::update_point_cell for all particles (src/Pic.cpp:186-196)
    def update_point_cells(self, xyz, cells):
        xyz = _f(xyz); cells = _i(cells).copy()
        self.ctx.check(self.ctx.L.fb_particle_cells(self.ctx.h, len(cells), _p(xyz), _p(cells)))
        return cells

    # field look-up of Pic::update_velocities (src/Pic.cpp:198-209)
    def fields(self, xyz, cells):
        xyz = _f(xyz); cells = _i(cells)
        E = np.zeros((len(cells), 3))
        self.ctx.check(self.ctx.L.fb_particle_field(self.ctx.h, len(cells), _p(xyz), _p(cells), _p(E)))
        return E


    # Pic::update_positions + ParticleSpecies::clear_lost (src/Pic.cpp:137-184, src/ParticleSpecies.cpp:16-31):
    # returns the surviving (pos, vel, cells) in their original order and the number of lost particles
    def update_positions(self, pos, vel, cells, dt, box, periodic=True):
        pos = _f(pos).copy(); vel = _f(vel).copy(); cells = _i(cells).copy(); box = _f(box)
        lost = C.c_long(0)
        self.ctx.check(self.ctx.L.fb_pic_update_positions(self.ctx.h, len(cells), _p(pos), _p(vel), _p(cells), float(dt), _p(box),
                                                          int(periodic), C.byref(lost)))
        k = len(cells) - lost.value
        return pos[:k], vel[:k], cells[:k], lost.value

    # Pic::update_velocities (src/Pic.cpp:198-209)
    def update_velocities(self, pos, vel, cells, dt, q_over_m):
        pos = _f(pos); vel = _f(vel).copy(); cells = _i(cells)
        self.ctx.check(self.ctx.L.fb_pic_update_velocities(self.ctx.h, len(cells), _p(pos), _p(cells), _p(vel), float(dt), float(q_over_m)))
        return vel


class PartitionPlan:
    """Host-only view of the multi-GPU partition logic (fb_plan_*; no CUDA): what rank `rank` of `world` keeps of
    a mesh, its local sparsity and its halo lists.  Used by the world_size-2 CPU tests."""

    def __init__(self, rank, world):
        self.L = _lib.load()
        self.h = C.c_void_p(self.L.fb_plan_create(int(rank), int(world)))
        self.rank, self.world = rank, world

    def __del__(self):
        try:
            if self.h:
                self.L.fb_destroy(self.h)
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise FemocsB200Error("fb_plan error %d: %s" % (rc, self.L.fb_last_error(self.h).decode()))

    def phase1(self, nodes, hexs, hex_markers):
        nodes = _f(nodes); hexs = _i(hexs); hex_markers = _i(hex_markers)
        bbox = np.zeros(6)
        self._check(self.L.fb_plan_phase1(self.h, _p(nodes), len(nodes), _p(hexs), _p(hex_markers), len(hexs), _p(bbox)))
        return bbox                      # local (min xyz, max xyz) of the boundary-face centres

    def import_whole(self, nodes, hexs, hex_markers, bulk=False, fe_degree=1):
        """the un-partitioned host import of fb_import_mesh (world 1): complete numbering and sparsity;
        bulk=True: the host import of fb_import_bulk_mesh (hexahedra with marker < 0, CurrentHeatSolver::mark_mesh);
        fe_degree=2: the FE_Q(2) system (option "fe_degree")"""
        nodes = _f(nodes); hexs = _i(hexs); hex_markers = _i(hex_markers)
        self._check(self.L.fb_plan_set_kind(self.h, int(bulk)))
        self._check(self.L.fb_set_option(self.h, b"fe_degree", float(fe_degree)))
        self._check(self.L.fb_plan_import(self.h, _p(nodes), len(nodes), _p(hexs), _p(hex_markers), len(hexs)))
        self.reused = bool(self.L.fb_last_import_reused(self.h))
        return self._collect()

    def phase2(self, bbox_global):
        b = _f(bbox_global)
        self._check(self.L.fb_plan_phase2(self.h, _p(b)))
        return self._collect()

    def _collect(self):
        sz = np.zeros(8, np.int64)
        self.L.fb_plan_sizes(self.h, _p(sz))
        (self.n_rows, self.n_cols, self.nnz, self.n_cells, self.n_send, self.n_ghost, self.n_vert_global, self.n_cells_global) = [int(v) for v in sz]
        w = self.world
        a = dict(local2global=np.zeros(self.n_cols, np.int32), owner=np.zeros(self.n_cols, np.int32), send_off=np.zeros(w + 1, np.int32),
                 send_idx=np.zeros(max(1, self.n_send), np.int32), recv_off=np.zeros(w + 1, np.int32), rowptr=np.zeros(self.n_rows + 1, np.int32),
                 col=np.zeros(self.nnz, np.int32), cells_dof=np.zeros((self.n_cells, 8), np.int32), cell2global=np.zeros(self.n_cells, np.int32),
                 copper=np.zeros(self.n_cols, np.int32), top=np.zeros(self.n_cols, np.int32))
        self._check(self.L.fb_plan_get(self.h, _p(a["local2global"]), _p(a["owner"]), _p(a["send_off"]), _p(a["send_idx"]), _p(a["recv_off"]),
                                       _p(a["rowptr"]), _p(a["col"]), _p(a["cells_dof"]), _p(a["cell2global"]), _p(a["copper"]), _p(a["top"])))
        a["send_idx"] = a["send_idx"][:self.n_send]
        self.__dict__.update(a)
        return self

    def cells27(self):
        out = np.zeros((self.n_cells, 27), np.int32)
        self._check(self.L.fb_get_cells27(self.h, _p(out)))
        return out

    def surface_centroids(self):
        """DealSolver::export_surface_centroids of the imported mesh (copper_surface faces, cell / face order)"""
        n = C.c_int(0)
        self._check(self.L.fb_export_surface_centroids(self.h, None, C.byref(n)))
        out = np.zeros((n.value, 3))
        self._check(self.L.fb_export_surface_centroids(self.h, _p(out), C.byref(n)))
        return out

    def jds(self, R=512, max_window=8192, sym=False, split=0):
        """block-JDS tables of the HBM SpMV for this plan's sparsity (fb_host_jds_build), as numpy arrays;
        split > 0: rows longer than that stored as chained segments (spmv_kernel 306): adds rowbeg / link"""
        sz = np.zeros(6, np.int64)
        if split:
            self._check(self.L.fb_plan_jds_split(self.h, int(R), int(max_window), int(split), _p(sz)))
        else:
            self._check(self.L.fb_plan_jds(self.h, int(R), int(max_window), int(sym), _p(sz)))
        nb, size, nwin, maxlen, wmax, njd = [int(v) for v in sz]
        t = dict(R=R, nb=nb, size=size, maxlen=maxlen, win_max=wmax,
                 perm=np.zeros(nb * R, np.uint16), len=np.zeros(nb * R, np.uint16), slot=np.zeros(self.n_rows, np.uint16),
                 jdp=np.zeros(nb + 1, np.int32), jd=np.zeros(njd, np.int32), base=np.zeros(nb + 1, np.int32),
                 col16=np.zeros(size, np.uint16), win_off=np.zeros(nb + 1, np.int32), win_list=np.zeros(max(1, nwin), np.int32))
        self._check(self.L.fb_plan_jds_get(self.h, _p(t["perm"]), _p(t["len"]), _p(t["slot"]), _p(t["jdp"]), _p(t["jd"]), _p(t["base"]),
                                           _p(t["col16"]), _p(t["win_off"]), _p(t["win_list"])))
        t["win_list"] = t["win_list"][:nwin]
        if split:
            t["rowbeg"] = np.zeros(nb + 1, np.int32); t["link"] = np.zeros(nb * R, np.uint16)
            self._check(self.L.fb_plan_jds_get_split(self.h, _p(t["rowbeg"]), _p(t["link"])))
        return t


@dataclass
class HeatingConfig:
    """Subset of Config::Heating used by the current / heat solvers (defaults: src/Config.cpp:73-84)."""
    lorentz: float = 2.44e-8
    t_ambient: float = 300.0
    n_cg: int = 2000
    cg_tolerance: float = 1e-9
    ssor_param: float = 1.2         # accepted for config compatibility; the GPU path uses Jacobi
    T_min: float = 0.0
    T_max: float = 1e5
    precond: int = PRECOND_JACOBI


class _EmissionSolver:
    """femocs::EmissionSolver<3> view (CurrentSolver / HeatSolver) of one of the two systems of a CurrentHeatSolver."""

    def __init__(self, parent, which):
        self.parent, self.which = parent, which
        self.stat_sol_min = 0.0
        self.stat_sol_max = 0.0
        self.last_residual = 0.0
        self.bc_values = None

    # EmissionSolver::set_bcs (include/CurrentHeatSolver.h:49-51): per-face data in export_surface_centroids order
    def set_bcs(self, bc_values):
        self.bc_values = _f(bc_values)

    # CurrentSolver::assemble() (src/CurrentHeatSolver.cpp:420-449) / HeatSolver::assemble(delta_time) (:105-152)
    def assemble(self, delta_time=None):
        ctx = self.parent.ctx
        bc = self.bc_values if self.bc_values is not None else np.zeros(self.parent.n_surface_faces)
        if self.which == 0:
            ctx.check(ctx.L.fb_current_assemble(ctx.h, _p(bc), len(bc)))
        else:
            ctx.check(ctx.L.fb_heat_assemble(ctx.h, float(delta_time), _p(bc), len(bc)))

    # EmissionSolver::solve (include/CurrentHeatSolver.h:41): +#CG / -#CG
    def solve(self, n_cg=None, cg_tolerance=None):
        ctx, conf = self.parent.ctx, self.parent.conf
        it = C.c_int(0); res = C.c_double(0)
        ctx.check(ctx.L.fb_ch_solve(ctx.h, self.which, int(conf.n_cg if n_cg is None else n_cg),
                                    float(conf.cg_tolerance if cg_tolerance is None else cg_tolerance), int(conf.precond),
                                    C.byref(it), C.byref(res)))
        self.last_residual = res.value
        return it.value

    def export_solution(self):
        ctx = self.parent.ctx
        out = np.zeros(self.parent.n_vertices)
        ctx.check(ctx.L.fb_ch_export_solution(ctx.h, self.which, _p(out)))
        return out

    def import_solution(self, vertex_values):
        ctx = self.parent.ctx
        v = _f(vertex_values)
        assert len(v) == self.parent.n_vertices
        ctx.check(ctx.L.fb_ch_import_solution(ctx.h, self.which, _p(v)))

    def export_solution_grad(self):
        ctx = self.parent.ctx
        out = np.zeros((self.parent.n_vertices, 3))
        ctx.check(ctx.L.fb_ch_export_solution_grad(ctx.h, self.which, _p(out)))
        return out

    # DealSolver::check_limits (src/DealSolver.cpp:157-167): True when OUT of limits, like the reference
    def check_limits(self, lo, hi):
        ctx = self.parent.ctx
        bad = C.c_int(0); mn = C.c_double(0); mx = C.c_double(0)
        ctx.check(ctx.L.fb_ch_check_limits(ctx.h, self.which, float(lo), float(hi), C.byref(bad), C.byref(mn), C.byref(mx)))
        self.stat_sol_min, self.stat_sol_max = mn.value, mx.value
        return bool(bad.value)


class CurrentHeatSolver:
    """femocs::CurrentHeatSolver<3> (include/CurrentHeatSolver.h:143-184) behind libfemocs_b200: members `current` and
    `heat` as in the reference (ProjectRunaway.cpp:542-553: ch_solver.current.assemble(); ch_solver.current.solve();
    ch_solver.heat.assemble(dt); ch_solver.heat.solve(); ch_solver.heat.check_limits(T_min, T_max))."""

    def __init__(self, ctx, conf=None):
        self.ctx = ctx
        self.conf = conf or HeatingConfig()
        self.current = _EmissionSolver(self, 0)
        self.heat = _EmissionSolver(self, 1)

    # PhysicalQuantities::resistivity_data rows (T, rho) and Config::Heating::lorentz
    def set_dependencies(self, table_T, table_rho, lorentz=None):
        T = _f(table_T); rho = _f(table_rho)
        self.ctx.check(self.ctx.L.fb_ch_set_physics(self.ctx.h, _p(T), _p(rho), len(T),
                                                    float(self.conf.lorentz if lorentz is None else lorentz)))

    # CurrentHeatSolver::import_mesh(nodes.export_dealii(), hexs.export_bulk()) (ProjectRunaway.cpp:222)
    def import_mesh(self, nodes, hexs, hex_markers):
        nodes = _f(nodes); hexs = _i(hexs); hex_markers = _i(hex_markers)
        rc = self.ctx.L.fb_import_bulk_mesh(self.ctx.h, _p(nodes), len(nodes), _p(hexs), _p(hex_markers), len(hexs))
        if rc == 3:
            return False
        self.ctx.check(rc)
        sz = np.zeros(7, np.int64)
        self.ctx.check(self.ctx.L.fb_get_sizes(self.ctx.h, _p(sz)))
        (self.n_dofs, self.n_cells, self.nnz, self.n_vertices, self.n_bfaces, self.n_surface_faces, _) = [int(v) for v in sz]
        return True

    def size(self):
        return self.n_dofs

    # CurrentHeatSolver::setup(temperature) (src/CurrentHeatSolver.cpp:509-513)
    def setup(self, temperature):
        self.ctx.check(self.ctx.L.fb_ch_setup(self.ctx.h, float(temperature)))

    # DealSolver::export_surface_centroids (src/DealSolver.cpp:229-245)
    def export_surface_centroids(self):
        n = C.c_int(0)
        self.ctx.check(self.ctx.L.fb_export_surface_centroids(self.ctx.h, None, C.byref(n)))
        out = np.zeros((n.value, 3))
        self.ctx.check(self.ctx.L.fb_export_surface_centroids(self.ctx.h, _p(out), C.byref(n)))
        return out

    # test hook: the system assembled last, as the reference holds it after apply_boundary_values (fb_get_system)
    get_system = PoissonSolver.get_system

    # CurrentHeatSolver::export_temp_rho (src/CurrentHeatSolver.cpp:515-523)
    def export_temp_rho(self):
        return self.heat.export_solution(), self.current.export_solution_grad()
