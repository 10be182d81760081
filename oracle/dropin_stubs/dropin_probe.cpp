// TEST INFRASTRUCTURE (oracle/_ref/libfemocs_dropin.so build only) -- not part of the product.
// The C ABI of include/Femocs_wrap.h has no entry for the zero-copy mesh pointers of the C++ class
// (Femocs::export_data(const double**, ...) / (const int**, ...), include/Femocs.h, src/GeneralProject.cpp:19-55);
// the drop-in test reads the mesh the reference generated through them to make sure it is the mesh of the golden
// fixtures before it compares fields with the oracle.
#include <cstring>
#include <string>

#include "Femocs.h"

extern "C" {

// kind: 0 nodes (3 doubles each) | 1 tetrahedra (4) | 2 hexahedra (8) | 3 triangles (3) | 4 quadrangles (4)
int dropin_mesh_count(femocs::Femocs* f, int kind) {
    static const char* names[5] = {"nodes", "tetrahedra", "hexahedra", "triangles", "quadrangles"};
    if (kind == 0) { const double* p = nullptr; return f->export_data(&p, names[0]); }
    const int* q = nullptr;
    return f->export_data(&q, names[kind]);
}

int dropin_mesh_copy(femocs::Femocs* f, int kind, void* out) {
    static const char* names[5] = {"nodes", "tetrahedra", "hexahedra", "triangles", "quadrangles"};
    static const int width[5] = {3, 4, 8, 3, 4};
    if (kind == 0) {
        const double* p = nullptr;
        const int n = f->export_data(&p, names[0]);
        if (n > 0) memcpy(out, p, sizeof(double) * 3 * (size_t) n);
        return n;
    }
    const int* q = nullptr;
    const int n = f->export_data(&q, names[kind]);
    if (n > 0) memcpy(out, q, sizeof(int) * width[kind] * (size_t) n);
    return n;
}

}
