#pragma once
#include "deal.II/numerics/vector_tools.h"
