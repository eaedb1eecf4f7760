#pragma once
#include <vector>
#include "layout.hpp"

namespace mfc {

// 27 coefficient arrays of `len` doubles (coefficient-major) for the cells lo .. lo+len-1
struct WenoTable {
    int lo = 0, len = 0;
    std::vector<double> data;
};

// cb -> ghosted cell boundaries s_cb(-1-b : N+b) of one direction
WenoTable build_weno_table(const double *cb, int N, int b, int weno_order);

}  // namespace mfc
