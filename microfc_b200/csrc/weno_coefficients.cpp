// weno_coefficients.cpp -- host-side set-up of the grid-dependent WENO coefficients
// (s_initialize_weno_module / s_compute_weno_coefficients, src/simulation/m_weno.fpp:103-363).
//
// The reference evaluates closed-form expressions of the cell-boundary coordinates for every
// cell; on uniform grids they reduce to the classical Jiang-Shu constants only up to rounding,
// so they are evaluated here in the same way (same differences, same products, same order)
// and streamed to the device as 27 coefficient arrays per direction.
#include "weno_coefficients.hpp"

namespace mfc {

// xb[0..5] = s_cb(i-2 .. i+3) for the cell whose left boundary is s_cb(i) (cell index i+1);
// out[27] in the order kernels.cuh::weno5 expects.
static void weno5_cell(const double *xb, double *out) {
    // d(a,c) = s_cb(i+a) - s_cb(i+c)
    auto d = [xb](int a, int c) { return xb[a + 2] - xb[c + 2]; };
    auto sq = [](double x) { return x*x; };
    double *pL = out, *pR = out + 6, *dL = out + 12, *dR = out + 15, *bt = out + 18;
    // candidate polynomials, index [k*2+q] = poly_coef_cb?(cell,k,q)      m_weno.fpp:225-265
    pR[0] = (d(0, 1)*d(1, 2))/(d(0, 3)*d(3, 1));
    pR[1] = (d(0, 2) + d(1, 3))/(d(0, 2)*d(0, 3))*d(0, 1);
    pR[2] = (d(-1, 1)*d(1, 0))/(d(-1, 2)*d(2, 0));
    pR[3] = (d(0, 1)*d(1, 2))/(d(-1, 1)*d(-1, 2));
    pR[4] = (d(-2, 1) + d(-1, 1))/(d(-1, 1)*d(1, -2))*d(1, 0);
    pR[5] = (d(0, 1)*d(1, -1))/(d(-2, 0)*d(-2, 1));
    pL[0] = (d(1, 0)*d(0, 2))/(d(0, 3)*d(3, 1));
    pL[1] = (d(0, 2) + d(0, 3))/(d(0, 2)*d(0, 3))*d(1, 0);
    pL[2] = (d(0, -1)*d(0, 1))/(d(-1, 2)*d(0, 2));
    pL[3] = (d(1, 0)*d(0, 2))/(d(-1, 1)*d(-1, 2));
    pL[4] = (d(-2, 0) + d(-1, 1))/(d(-2, 1)*d(1, -1))*d(0, 1);
    pL[5] = (d(-1, 0)*d(0, 1))/(d(-2, 0)*d(-2, 1));
    // ideal weights                                                       :267-281
    dR[0] = (d(-2, 1)*d(1, -1))/(d(-2, 3)*d(3, -1));
    dR[2] = (d(1, 2)*d(1, 3))/(d(-2, 2)*d(-2, 3));
    dL[0] = (d(-2, 0)*d(0, -1))/(d(-2, 3)*d(3, -1));
    dL[2] = (d(0, 2)*d(0, 3))/(d(-2, 2)*d(-2, 3));
    dR[1] = 1.0 - dR[0] - dR[2];
    dL[1] = 1.0 - dL[0] - dL[2];
    // smoothness indicators, index [k*3+q] = beta_coef(cell,k,q)          :283-345
    const double h2 = 4.0*sq(d(0, 1));
    const double w = d(1, 0);                 // width of the cell itself
    bt[0] = h2*(10.0*sq(w) + w*d(2, 1) + sq(d(2, 1)))/(sq(d(0, 3))*sq(d(1, 3)));
    bt[1] = h2*(19.0*sq(w) - w*d(3, 1) + 2.0*d(2, 0)*(d(2, 0) + d(3, 1)))/(d(0, 2)*sq(d(0, 3))*d(3, 1));
    bt[2] = h2*(10.0*sq(w) + w*(d(2, 0) + d(3, 1)) + sq(d(2, 0) + d(3, 1)))/(sq(d(0, 2))*sq(d(0, 3)));
    bt[3] = h2*(10.0*sq(w) + sq(d(0, -1)) + d(0, -1)*w)/(sq(d(-1, 2))*sq(d(0, 2)));
    bt[4] = h2*(d(0, 1)*(d(0, -1) + 20.0*w) + (2.0*d(0, -1) + w)*d(2, 0))/(d(1, -1)*sq(d(-1, 2))*d(2, 0));
    bt[5] = h2*(10.0*sq(w) + w*d(2, 1) + sq(d(2, 1)))/(sq(d(-1, 1))*sq(d(-1, 2)));
    bt[6] = h2*(12.0*sq(w) + sq(d(0, -2) + d(0, -1)) + 3.0*(d(0, -2) + d(0, -1))*w)/(sq(d(-2, 1))*sq(d(-1, 1)));
    bt[7] = h2*(19.0*sq(w) + (d(0, -2)*d(0, 1)) + 2.0*d(1, -1)*(d(0, -2) + d(1, -1)))/(d(-2, 0)*sq(d(-2, 1))*d(1, -1));
    bt[8] = h2*(10.0*sq(w) + sq(d(0, -1)) + d(0, -1)*w)/(sq(d(-2, 0))*sq(d(-2, 1)));
}

// WENO3, m_weno.fpp:193-217.  xb[0..3] = s_cb(i-1 .. i+2).  The two candidate stencils use the
// slots of k = 0, 1 / q = 0 of the 27-slot layout (poly [k*2], d [k], beta [k*3]); the rest is 0.
static void weno3_cell(const double *xb, double *out) {
    auto d = [xb](int a, int c) { return xb[a + 1] - xb[c + 1]; };
    auto sq = [](double x) { return x*x; };
    for (int c = 0; c < kNumWenoCoef; c++) out[c] = 0.0;
    double *pL = out, *pR = out + 6, *dL = out + 12, *dR = out + 15, *bt = out + 18;
    pR[0] = d(0, 1)/d(0, 2);
    pR[2] = d(0, 1)/d(-1, 1);
    pL[0] = -pR[0];
    pL[2] = -pR[2];
    dR[0] = d(-1, 1)/d(-1, 2);
    dL[0] = d(-1, 0)/d(-1, 2);
    dR[1] = 1.0 - dR[0];
    dL[1] = 1.0 - dL[0];
    bt[0] = 4.0*sq(d(0, 1))/sq(d(0, 2));
    bt[3] = 4.0*sq(d(0, 1))/sq(d(-1, 1));
}

WenoTable build_weno_table(const double *cb, int N, int b, int weno_order) {
    // cb -> s_cb(-1-b : N+b); cells -b+polyn .. N+b-polyn (m_weno.fpp:118-127)
    const int polyn = (weno_order - 1)/2;
    WenoTable t;
    t.lo = -b + polyn;
    t.len = N + 1 + 2*b - 2*polyn;
    t.data.assign((size_t)kNumWenoCoef*t.len, 0.0);
    if (weno_order == 1) return t;                    // :105, first order needs no coefficients
    double tmp[kNumWenoCoef];
    for (int n = 0; n < t.len; n++) {
        const int cell = t.lo + n;
        const int i = cell - 1;                       // left boundary index of the cell
        // element index of s_cb(x) is x + 1 + b
        if (weno_order == 5) weno5_cell(cb + (i - 2) + 1 + b, tmp);
        else weno3_cell(cb + (i - 1) + 1 + b, tmp);
        for (int c = 0; c < kNumWenoCoef; c++) t.data[(size_t)c*t.len + n] = tmp[c];
    }
    return t;
}

}  // namespace mfc
