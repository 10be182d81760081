// TEST INFRASTRUCTURE -- not part of the product, never linked into libfemocs_b200.
//
// Thin extern "C" façade over the *verbatim* reference sources (compiled from
// /root/reference by oracle/Makefile.ref into oracle/_ref/libfemocs_ref.so).
// It drives the reference's own mesher and interpolator classes exactly as
// Femocs::import_atoms (Femocs.cpp:116-146) and ProjectRunaway::generate_mesh
// (ProjectRunaway.cpp:147-212) do, and exposes their arrays and their
// locate/interpolate loops so that
//   * tests/golden fixtures can be generated (oracle/make_golden.py), and
//   * oracle/femocs_oracle.cpp can be diffed against the real reference code.
// Nothing here restates reference arithmetic; every number comes out of the
// reference's own functions.

#define MAINFILE
#include "Globals.h"
#include "Macros.h"
#include "Config.h"
#include "AtomReader.h"
#include "Surface.h"
#include "Coarseners.h"
#include "TetgenMesh.h"
#include "Interpolator.h"
#include "SolutionReader.h"
#include "ParticleSpecies.h"

#include <cstring>
#include <memory>

using namespace femocs;
using namespace std;

namespace {

struct RefState {
    Config conf;
    AtomReader reader;
    Coarseners coarseners;
    Surface dense_surf, extended_surf;
    unique_ptr<TetgenMesh> mesh;
    unique_ptr<Interpolator> interp;
    bool first_run = true;
    RefState() : reader(&conf.geometry) {}
};

RefState* S = nullptr;

}  // namespace

extern "C" {

int ref_init(const char* conf_file) {
    delete S;
    S = new RefState();
    S->conf.read_all(conf_file);
    MODES.MUTE = true;
    MODES.VERBOSE = false;
    MODES.WRITELOG = false;
    MODES.WRITE_PERIOD = -1;
    S->dense_surf.set_coarsener(&S->coarseners);
    S->mesh.reset(new TetgenMesh(&S->conf.mesh));
    return 0;
}

// Femocs::import_atoms (Femocs.cpp:116-146)
int ref_import_atoms(const char* file_name) {
    string fname(file_name);
    S->reader.import_file(fname, false);
    if (get_file_type(fname) == "xyz") {
        if (S->conf.run.rdf) S->reader.calc_rdf_coordinations(NULL);
        else S->reader.calc_coordinations(NULL);
        if (S->conf.run.cluster_anal) S->reader.calc_clusters(NULL);
        S->reader.extract_types();
    } else {
        S->reader.calc_pseudo_coordinations();
    }
    return S->reader.size();
}

// overwrite atom positions (for the "wobble" style MD emulation, Main.cpp:374-378)
int ref_n_atoms() { return S->reader.size(); }
void ref_get_atoms(double* xyz) {
    for (int i = 0; i < S->reader.size(); ++i) {
        Point3 p = S->reader.get_point(i);
        xyz[3*i] = p.x; xyz[3*i+1] = p.y; xyz[3*i+2] = p.z;
    }
}

// ProjectRunaway::generate_boundary_nodes + generate_mesh (ProjectRunaway.cpp:147-212)
int ref_generate_mesh() {
    S->mesh.reset(new TetgenMesh(&S->conf.mesh));
    int fail;
    if (S->conf.path.mesh_file != "") {
        fail = S->mesh->read(S->conf.path.mesh_file, "rQnn");
    } else {
        Surface bulk, coarse_surf, vacuum;
        S->reader.extract(S->dense_surf, TYPES.SURFACE);
        S->coarseners.generate(S->dense_surf, S->conf.geometry, S->conf.cfactor);
        if (S->first_run) S->dense_surf.extend(S->extended_surf, S->conf);
        fail = S->dense_surf.generate_boundary_nodes(bulk, coarse_surf, vacuum, S->extended_surf, S->conf);
        if (fail) return 1;
        fail = S->mesh->generate(bulk, coarse_surf, vacuum);
    }
    if (fail) return 1;
    S->interp.reset(new Interpolator(LABELS.elfield, LABELS.charge_density, LABELS.potential));
    if (S->conf.run.surface_cleaner && S->dense_surf.size() > 0)
        S->dense_surf.clean_by_triangles(*S->interp, S->mesh.get(), S->conf.geometry.latconst);
    S->first_run = false;
    return 0;
}

int ref_n_surface_atoms() { return S->dense_surf.size(); }
void ref_get_surface_atoms(double* xyz, int* ids) {
    for (int i = 0; i < S->dense_surf.size(); ++i) {
        Point3 p = S->dense_surf.get_point(i);
        xyz[3*i] = p.x; xyz[3*i+1] = p.y; xyz[3*i+2] = p.z;
        ids[i] = S->dense_surf.get_id(i);
    }
}

// out: n_nodes, n_tets, n_hexs, n_tris, n_quads
void ref_mesh_sizes(int* out) {
    const TetgenMesh& m = *S->mesh;
    out[0] = m.nodes.size(); out[1] = m.tets.size(); out[2] = m.hexs.size();
    out[3] = m.tris.size();  out[4] = m.quads.size();
}

void ref_get_nodes(double* xyz, int* markers) {
    const TetgenMesh& m = *S->mesh;
    memcpy(xyz, m.nodes.get(), sizeof(double) * 3 * m.nodes.size());
    for (int i = 0; i < m.nodes.size(); ++i) markers[i] = m.nodes.get_marker(i);
}

void ref_get_tets(int* tet4, int* nbr4, int* markers) {
    const TetgenMesh& m = *S->mesh;
    for (int i = 0; i < m.tets.size(); ++i) {
        SimpleElement e = m.tets[i];
        vector<int> nb = m.tets.get_neighbours(i);
        for (int k = 0; k < 4; ++k) { tet4[4*i+k] = e[k]; nbr4[4*i+k] = nb[k]; }
        markers[i] = m.tets.get_marker(i);
    }
}

void ref_get_hexs(int* hex8, int* markers) {
    const TetgenMesh& m = *S->mesh;
    memcpy(hex8, m.hexs.get(), sizeof(int) * 8 * m.hexs.size());
    for (int i = 0; i < m.hexs.size(); ++i) markers[i] = m.hexs.get_marker(i);
}

void ref_get_tris(int* tri3, int* tri2tet, double* norm3) {
    const TetgenMesh& m = *S->mesh;
    for (int i = 0; i < m.tris.size(); ++i) {
        SimpleFace f = m.tris[i];
        array<int,2> t = m.tris.to_tets(i);
        Vec3 n = m.tris.get_norm(i);
        for (int k = 0; k < 3; ++k) { tri3[3*i+k] = f[k]; norm3[3*i+k] = n[k]; }
        tri2tet[2*i] = t[0]; tri2tet[2*i+1] = t[1];
    }
}

void ref_get_quads(int* quad4, int* quad2hex) {
    const TetgenMesh& m = *S->mesh;
    for (int i = 0; i < m.quads.size(); ++i) {
        SimpleQuad q = m.quads[i];
        array<int,2> h = m.quads.to_hexs(i);
        for (int k = 0; k < 4; ++k) quad4[4*i+k] = q[k];
        quad2hex[2*i] = h[0]; quad2hex[2*i+1] = h[1];
    }
}

// out: tets.stat.edgemax, tris.stat.edgemax
void ref_get_stats(double* out) {
    out[0] = S->mesh->tets.stat.edgemax;
    out[1] = S->mesh->tris.stat.edgemax;
}

// TetgenMesh::calc_pseudo_3D_vorocells (TetgenMesh.cpp:819-837); call with list==NULL to size
int ref_voro_nbors(int vacuum, int* offsets, int* list) {
    vector<vector<unsigned>> cells;
    S->mesh->calc_pseudo_3D_vorocells(cells, vacuum != 0);
    int k = 0;
    for (size_t i = 0; i < cells.size(); ++i) {
        if (offsets) offsets[i] = k;
        for (unsigned v : cells[i]) { if (list) list[k] = (int) v; ++k; }
    }
    if (offsets) offsets[cells.size()] = k;
    return list || offsets ? k : (int) cells.size();
}
int ref_voro_total() {
    vector<vector<unsigned>> cells;
    S->mesh->calc_pseudo_3D_vorocells(cells, true);
    int k = 0;
    for (auto& c : cells) k += (int) c.size();
    return k;
}

// Interpolator::initialize (Interpolator.cpp:28-77), search region = vacuum as in ProjectRunaway.cpp:435
void ref_interp_init() {
    if (!S->interp) S->interp.reset(new Interpolator(LABELS.elfield, LABELS.charge_density, LABELS.potential));
    S->interp->initialize(S->mesh.get(), 0, TYPES.VACUUM);
}

void ref_get_maps(int* node_femocs2deal, int* hex_femocs2deal) {
    for (int i = 0; i < S->interp->nodes.size(); ++i) node_femocs2deal[i] = S->interp->nodes.femocs2deal(i);
    for (int i = 0; i < S->mesh->hexs.size(); ++i) hex_femocs2deal[i] = S->interp->linhex.femocs2deal(i);
}

// set nodal Solution {Ex,Ey,Ez,scalar1,scalar2} directly
void ref_set_nodal(const double* sol5) {
    for (int i = 0; i < S->interp->nodes.size(); ++i)
        S->interp->nodes.set_solution(i, Solution(Vec3(sol5[5*i], sol5[5*i+1], sol5[5*i+2]), sol5[5*i+3], sol5[5*i+4]));
}

void ref_get_nodal(double* sol5) {
    for (int i = 0; i < S->interp->nodes.size(); ++i) {
        Solution s = S->interp->nodes.get_solution(i);
        sol5[5*i] = s.vector.x; sol5[5*i+1] = s.vector.y; sol5[5*i+2] = s.vector.z;
        sol5[5*i+3] = s.scalar1; sol5[5*i+4] = s.scalar2;
    }
}

// Interpolator::extract_solution(PoissonSolver<3>&, smoothen)  (Interpolator.cpp:172-190)
// phi/rho are in solver-vertex order (what export_solution / export_charge_dens return).
void ref_extract_solution(const double* phi, const double* rho, int n_vert, int smoothen) {
    PoissonSolver<3> fem;
    fem.vertex_solution.assign(phi, phi + n_vert);
    fem.vertex_charge.assign(rho, rho + n_vert);
    S->interp->extract_solution(fem, smoothen != 0);
}

// SolutionReader::interpolate_results body (SolutionReader.cpp:405-421) with the located
// cells (atom markers) and full Solution exported instead of one label.
void ref_locate_interpolate(int dim, int rank, int n, const double* xyz, int* cells, double* sol5) {
    SolutionReader sr(S->interp.get(), LABELS.elfield, LABELS.charge_density, LABELS.potential);
    sr.set_preferences(false, dim, rank);
    sr.reserve(n);
    for (int i = 0; i < n; ++i)
        sr.append(Atom(i, Point3(xyz[3*i], xyz[3*i+1], xyz[3*i+2]), 0));
    sr.calc_interpolation();
    for (int i = 0; i < n; ++i) {
        Solution s = sr.get_interpolation(i);
        cells[i] = sr.get_marker(i);
        sol5[5*i] = s.vector.x; sol5[5*i+1] = s.vector.y; sol5[5*i+2] = s.vector.z;
        sol5[5*i+3] = s.scalar1; sol5[5*i+4] = s.scalar2;
    }
}

// SolutionReader::calc_interpolation with atoms_mapped_to_cells (SolutionReader.cpp:167-190)
void ref_interp_known_cells(int dim, int rank, int n, const double* xyz, const int* cells, double* sol5) {
    Interpolator& I = *S->interp;
    for (int i = 0; i < n; ++i) {
        Point3 p(xyz[3*i], xyz[3*i+1], xyz[3*i+2]);
        int cell = abs(cells[i]);
        Solution s;
        if (dim == 2) {
            if (rank == 1) s = I.lintri.interp_solution(p, cell);
            else if (rank == 2) s = I.quadtri.interp_solution(p, cell);
            else s = I.linquad.interp_solution(p, cell);
        } else {
            if (rank == 1) s = I.lintet.interp_solution(p, cell);
            else if (rank == 2) s = I.quadtet.interp_solution(p, cell);
            else s = I.linhex.interp_solution(p, cell);
        }
        sol5[5*i] = s.vector.x; sol5[5*i+1] = s.vector.y; sol5[5*i+2] = s.vector.z;
        sol5[5*i+3] = s.scalar1; sol5[5*i+4] = s.scalar2;
    }
}

// SolutionReader::export_results (SolutionReader.cpp:303-398) of the reference itself: atoms with the given ids and
// interpolated Solutions, exported under `label` into data (n_points entries, 3 n_points for the vector label).
// Returns the reference's return value (0 ok, 1 nothing to export).
int ref_export_results(int n, const int* ids, const double* sol5, int n_points, const char* label, double* data) {
    SolutionReader sr(S->interp.get(), LABELS.elfield, LABELS.charge_density, LABELS.potential);
    sr.reserve(n);
    for (int i = 0; i < n; ++i) sr.append(Atom(ids[i], Point3(0, 0, 0), 0));
    for (int i = 0; i < n; ++i)          // reserve() has sized the interpolation vector
        sr.set_interpolation(i, Solution(Vec3(sol5[5*i], sol5[5*i+1], sol5[5*i+2]), sol5[5*i+3], sol5[5*i+4]));
    return sr.export_results(n_points, label, data);
}

// periodic_image (Macros.cpp:41-48), element-wise
void ref_periodic_image(int n, const double* p, double pmax, double pmin, double* out) {
    for (int i = 0; i < n; ++i) out[i] = periodic_image(p[i], pmax, pmin);
}

// ParticleSpecies::clear_lost (ParticleSpecies.cpp:16-31) on particles (pos, vel, cell): survivors written back in place,
// returns the number of lost particles
int ref_clear_lost(int n, double* pos3, double* vel3, int* cell) {
    ParticleSpecies ps(-17.5882, -180.9512268, 0.01);
    ps.reserve(n);
    for (int i = 0; i < n; ++i)
        ps.inject_particle(Point3(pos3[3*i], pos3[3*i+1], pos3[3*i+2]), Vec3(vel3[3*i], vel3[3*i+1], vel3[3*i+2]), cell[i]);
    const int lost = ps.clear_lost();
    for (int i = 0; i < ps.size(); ++i) {
        const SuperParticle& sp = ps[i];
        pos3[3*i] = sp.pos.x; pos3[3*i+1] = sp.pos.y; pos3[3*i+2] = sp.pos.z;
        vel3[3*i] = sp.vel.x; vel3[3*i+1] = sp.vel.y; vel3[3*i+2] = sp.vel.z;
        cell[i] = sp.cell;
    }
    return lost;
}

// Pic::update_point_cell (Pic.cpp:186-196): solver-cell guess in, solver-cell (or -1) out
void ref_particle_cells(int n, const double* xyz, int* cell_inout) {
    Interpolator& I = *S->interp;
    for (int i = 0; i < n; ++i) {
        Point3 p(xyz[3*i], xyz[3*i+1], xyz[3*i+2]);
        int fc = I.linhex.deal2femocs(cell_inout[i]);
        fc = I.linhex.locate_cell(p, fc);
        cell_inout[i] = fc < 0 ? -1 : I.linhex.femocs2deal(fc);
    }
}

// raw LinearHexahedra::locate_cell in femocs hex indices (InterpolatorCells.cpp:1536-1562)
void ref_linhex_locate(int n, const double* xyz, int* cell_inout) {
    Interpolator& I = *S->interp;
    for (int i = 0; i < n; ++i)
        cell_inout[i] = I.linhex.locate_cell(Point3(xyz[3*i], xyz[3*i+1], xyz[3*i+2]), cell_inout[i]);
}

// Pic::update_velocities field lookup (Pic.cpp:198-209): E = linhex.interp_gradient(pos, deal2femocs(cell))
void ref_particle_field(int n, const double* xyz, const int* deal_cells, double* E3) {
    Interpolator& I = *S->interp;
    for (int i = 0; i < n; ++i) {
        int cell = I.linhex.deal2femocs(deal_cells[i]);
        Vec3 E = I.linhex.interp_gradient(Point3(xyz[3*i], xyz[3*i+1], xyz[3*i+2]), cell);
        E3[3*i] = E.x; E3[3*i+1] = E.y; E3[3*i+2] = E.z;
    }
}

// LinearHexahedra::shape_funs_dealii (InterpolatorCells.cpp:1355-1358), used by
// PoissonSolver<3>::assemble_space_charge_fast (PoissonSolver.cpp:299-319)
void ref_particle_weights(int n, const double* xyz, const int* deal_cells, double* w8) {
    Interpolator& I = *S->interp;
    for (int i = 0; i < n; ++i) {
        int cell = I.linhex.deal2femocs(deal_cells[i]);
        array<double,8> w = I.linhex.shape_funs_dealii(Vec3(xyz[3*i], xyz[3*i+1], xyz[3*i+2]), cell);
        for (int k = 0; k < 8; ++k) w8[8*i+k] = w[k];
    }
}

// LinearHexahedra::interp_gradient(hex, node) (InterpolatorCells.cpp:425-439,1462-1505)
void ref_nodal_gradient(int n, const int* hex, const int* node, double* E3) {
    Interpolator& I = *S->interp;
    for (int i = 0; i < n; ++i) {
        Vec3 E = I.linhex.interp_gradient(hex[i], node[i]);
        E3[3*i] = E.x; E3[3*i+1] = E.y; E3[3*i+2] = E.z;
    }
}

}  // extern "C"
