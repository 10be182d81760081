// TEST INFRASTRUCTURE (oracle/_ref build only) -- not part of the product.
// Minimal stand-in for the three deal.II value types that leak into the
// reference's non-solver code through include/Primitives.h:14
// (dealii::Point<3> at Primitives.h:237, dealii::Tensor<1,3> at :290) and
// TetgenCells.cpp:179-184,673-686 (dealii::CellData<3>).  deal.II 9.2 itself is
// not installed in this image, so the reference's mesher/interpolator sources
// are compiled against this header instead.  It also pulls in the std headers
// the reference relies on deal.II to include transitively.
#pragma once
#include <algorithm>
#include <array>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <map>
#include <numeric>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace dealii {

template <int dim, typename Number = double>
class Point {
public:
    Point() { for (int d = 0; d < dim; ++d) v_[d] = Number(0); }
    Point(Number x, Number y, Number z) { v_[0] = x; v_[1] = y; v_[dim - 1] = z; }
    Number& operator[](unsigned d) { return v_[d]; }
    const Number& operator[](unsigned d) const { return v_[d]; }
    Number& operator()(unsigned d) { return v_[d]; }
    const Number& operator()(unsigned d) const { return v_[d]; }
private:
    Number v_[dim];
};

template <int rank, int dim, typename Number = double>
class Tensor {
public:
    Tensor() { for (int d = 0; d < dim; ++d) v_[d] = Number(0); }
    Number& operator[](unsigned d) { return v_[d]; }
    const Number& operator[](unsigned d) const { return v_[d]; }
private:
    Number v_[dim];
};

template <int dim>
struct CellData {
    unsigned int vertices[1 << dim];
    unsigned int material_id = 0;
};

namespace types { typedef unsigned int global_dof_index; }

}  // namespace dealii
