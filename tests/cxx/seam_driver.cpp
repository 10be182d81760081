// Drives the C++ seam (include/femocs_b200.hpp) the way ProjectRunaway::run drives the reference classes
// (src/ProjectRunaway.cpp:214-228 import_mesh, :422-447 solve_laplace, :295-309 prepare_export, :644-658 interpolate):
//   seam_driver <mesh.bin> <out.bin> <E0>
// mesh.bin / out.bin: a flat sequence of records {int32 name_len, name, int32 kind (0 = int32, 1 = float64), int64 count, data}
// written / read by tests/test_cxx_seam.py.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <string>
#include <vector>

#include "femocs_b200.hpp"

struct Blob { std::map<std::string, std::vector<int>> i; std::map<std::string, std::vector<double>> d; };

static Blob read_blob(const char* path) {
    Blob b; std::ifstream f(path, std::ios::binary);
    if (!f) { std::cerr << "cannot open " << path << "\n"; std::exit(2); }
    while (true) {
        int32_t nl; if (!f.read((char*) &nl, 4)) break;
        std::string name(nl, ' '); f.read(&name[0], nl);
        int32_t kind; int64_t cnt; f.read((char*) &kind, 4); f.read((char*) &cnt, 8);
        if (kind == 0) { auto& v = b.i[name]; v.resize(cnt); f.read((char*) v.data(), 4 * cnt); }
        else { auto& v = b.d[name]; v.resize(cnt); f.read((char*) v.data(), 8 * cnt); }
    }
    return b;
}
static void put(std::ofstream& f, const std::string& name, const std::vector<int>& v) {
    int32_t nl = (int32_t) name.size(), kind = 0; int64_t cnt = (int64_t) v.size();
    f.write((char*) &nl, 4); f.write(name.data(), nl); f.write((char*) &kind, 4); f.write((char*) &cnt, 8); f.write((char*) v.data(), 4 * cnt);
}
static void put(std::ofstream& f, const std::string& name, const std::vector<double>& v) {
    int32_t nl = (int32_t) name.size(), kind = 1; int64_t cnt = (int64_t) v.size();
    f.write((char*) &nl, 4); f.write(name.data(), nl); f.write((char*) &kind, 4); f.write((char*) &cnt, 8); f.write((char*) v.data(), 8 * cnt);
}

int main(int argc, char** argv) {
    if (argc < 4) { std::cerr << "usage: seam_driver mesh.bin out.bin E0\n"; return 2; }
    using namespace femocs_b200;
    Blob b = read_blob(argv[1]);
    const double E0 = std::atof(argv[3]);
    try {
        Context ctx(0);
        FieldConfig conf; conf.E0 = E0; conf.cg_tolerance = 1e-11;
        MeshArrays m;
        m.nodes = b.d["nodes"].data(); m.n_nodes = (int) b.d["nodes"].size() / 3; m.node_markers = b.i["node_markers"].data();
        m.hexs = b.i["hexs"].data(); m.hex_markers = b.i["hex_markers"].data(); m.n_hexs = (int) b.i["hex_markers"].size();
        m.tets = b.i["tets"].data(); m.tet_nbrs = b.i["tet_nbrs"].data(); m.tet_markers = b.i["tet_markers"].data(); m.n_tets = (int) b.i["tet_markers"].size();
        m.tris = b.i["tris"].data(); m.tri2tet = b.i["tri2tet"].data(); m.tri_norms = b.d["tri_norms"].data(); m.n_tris = (int) b.i["tris"].size() / 3;
        m.quads = b.i["quads"].data(); m.quad2hex = b.i["quad2hex"].data(); m.n_quads = (int) b.i["quads"].size() / 4;
        m.tet_edgemax = b.d["edgemax"][0];
        m.voro_off = b.i["voro_off"].data(); m.voro_list = b.i["voro_list"].data(); m.n_voro = (int) b.i["voro_off"].size() - 1;

        PoissonSolver poisson_solver(ctx, &conf);
        Interpolator vacuum_interpolator(ctx);
        FieldReader fields(&vacuum_interpolator);
        if (!poisson_solver.import_mesh(m)) { std::cerr << "import_mesh failed\n"; return 1; }      // ProjectRunaway.cpp:216
        poisson_solver.setup(-conf.E0, conf.V0);                                                    // :424
        poisson_solver.assemble(true);                                                              // :425
        const int ncg = poisson_solver.solve();                                                     // :431
        if (ncg < 0) { std::cerr << "Field solver did not complete normally, #CG=" << -ncg << "\n"; return 1; }
        const bool out_of_limits = poisson_solver.check_limits(conf.V_min, conf.V_max);
        vacuum_interpolator.initialize(m);                                                          // :435
        vacuum_interpolator.extract_solution(poisson_solver, true);                                 // :436
        const std::vector<double>& a = b.d["surf_atoms"];                                           // AoS from the fixture -> the SoA of Femocs_wrap.h:36
        const int n = (int) a.size() / 3;
        std::vector<double> x(n), y(n), z(n);
        for (int i = 0; i < n; ++i) { x[i] = a[3 * i]; y[i] = a[3 * i + 1]; z[i] = a[3 * i + 2]; }
        fields.set_preferences(false, 2, 1);                                                        // :303
        fields.interpolate(n, x.data(), y.data(), z.data());                                        // :304
        std::vector<double> E(3 * (size_t) n, 0.0), phi(n, 7.0), Enorm(n, 0.0);
        fields.export_results(n, "elfield", E.data());          // exact case: appended to the zeroed array
        fields.export_results(n, "POTENTIAL", phi.data());      // upper case: overwrites the 7.0
        fields.export_results(n, "elfield_norm", Enorm.data());
        std::vector<int> flag(n, 0); fields.export_flags(n, flag.data());
        std::vector<double> sol; poisson_solver.export_solution(sol);
        // one PIC push through the Pic seam (make_pic_step, ProjectRunaway.cpp:492-511): update_positions, then update_velocities
        std::vector<double> ppos = b.d["pic_pos"], pvel = b.d["pic_vel"];
        std::vector<int> pcell = b.i["pic_cells"];
        int n_lost = 0;
        if (!pcell.empty()) {
            Pic pic(&vacuum_interpolator);
            const long np = (long) pcell.size();
            n_lost = pic.update_positions(np, ppos.data(), pvel.data(), pcell.data(), b.d["pic_dt"][0], b.d["pic_box"].data(), true);
            const long kept = np - n_lost;
            ppos.resize(3 * kept); pvel.resize(3 * kept); pcell.resize(kept);
            pic.update_velocities(kept, ppos.data(), pcell.data(), pvel.data(), b.d["pic_dt"][0], b.d["pic_dt"][1]);
        }
        std::ofstream f(argv[2], std::ios::binary);
        put(f, "pic_pos", ppos); put(f, "pic_vel", pvel); put(f, "pic_cells", pcell); put(f, "pic_lost", std::vector<int>{n_lost});
        put(f, "ncg", std::vector<int>{ncg, (int) out_of_limits, (int) ctx.kernel_launches()});
        put(f, "stat", std::vector<double>{poisson_solver.stat.sol_min, poisson_solver.stat.sol_max, fields.E_max});
        put(f, "phi_vertex", sol); put(f, "markers", fields.markers); put(f, "E", E); put(f, "phi", phi); put(f, "Enorm", Enorm); put(f, "flag", flag);
        std::cout << poisson_solver.to_str() << " #CG=" << ncg << std::endl;
    } catch (const std::exception& e) {
        std::cerr << "seam_driver: " << e.what() << "\n";
        return 3;
    }
    return 0;
}
