// kernels.cuh -- hand-written FP64 CUDA kernels (sm_100a) for the s_compute_rhs pipeline and
// the fused TVD-RK update.  No tensor cores: this is a stencil, not a dense contraction.
//
// Compiled twice (kernels_fast.cu / kernels_strict.cu):
//   MFC_STRICT=0  namespace mfc_fast    default FMA contraction, WENO weights in the
//                                       one-division product form (see weno5())
//   MFC_STRICT=1  namespace mfc_strict  -fmad=false and the reference's exact operation
//                                       order: bit-comparable with a strict CPU build
//
// Stage map (reference file:line -> kernel), all citations relative to /root/reference:
//   m_rhs.fpp:686-908 ghost fill (physical BCs)              -> k_bc
//   m_mpi_proxy.fpp:490-499,592-601 (+y) pack / unpack       -> k_halo_pack / k_halo_unpack
//   m_variables_conversion.fpp:313-375 cons -> prim          -> k_prim
//   m_weno.fpp:470-535 + m_riemann_solvers.fpp:132-327 +
//   m_rhs.fpp:565-653 + m_time_steppers.fpp:298-348          -> k_xstream (x: cell stream through smem rings)
//                                                               k_march3 (y / z: marching pencils)
//   m_data_output.fpp:197-258 stability criteria             -> k_stability
#pragma once
#include <cuda_runtime.h>
#include "args.hpp"

#ifndef MFC_STRICT
#define MFC_STRICT 0
#endif
#if MFC_STRICT
#define MFC_NS mfc_strict
#else
#define MFC_NS mfc_fast
#endif

namespace MFC_NS {

using namespace mfc;

// ------------------------------------------------------------------------------------------
// RK update, m_time_steppers.fpp -- the operand order of each statement is kept
//   1: q1 + dt*rhs                      (:167-169, :227-229, :302-304)
//   2: (q1 + q2 + dt*rhs)/2             (:245-248)
//   3: (3*q1 + q2 + dt*rhs)/4           (:322-325)
//   4: (q1 + 2*q2 + 2*dt*rhs)/3         (:342-345)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double rk_apply(int mode, double q1, double qs, double rhs, double dt) {
    switch (mode) {
    case 1: return q1 + dt*rhs;
    case 2: return (q1 + qs + dt*rhs)/2.0;
    case 3: return (3.0*q1 + qs + dt*rhs)/4.0;
    default: return (q1 + 2.0*qs + 2.0*dt*rhs)/3.0;
    }
}

// ------------------------------------------------------------------------------------------
// fast-build arithmetic helpers.  An IEEE-rounded FP64 division costs ~12.5 DFMA issue slots on
// sm_100a (tools/fp64_peak.cu); the MUFU seed + two Newton steps below costs 4 and is accurate
// to ~1 ulp (not correctly rounded) -- well inside the 1e-10 parity budget, re-gated by
// tests/test_gpu_parity.py and the device self-test (mfc_b200_selftest_math).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double rcp_fast(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}
// sqrt(n/d) for n, d > 0 as n*rsqrt(n*d): one MUFU seed and one cubically convergent correction
// y (1 + e/2 + 3 e^2/8), e = 1 - x y^2 (seed error ~2^-23 -> e^3), no division: 7 FP64 instructions
__device__ __forceinline__ double sqrt_ratio_fast(double n, double d) {
    const double x = n*d;
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-(x*y), y, 1.0);
    const double p = fma(0.375, e, 0.5);
    y = fma(y, p*e, y);
    return n*y;
}

// 1/x with a cubically convergent correction: seed error e ~ 2^-23 -> e^3, i.e. ~1 ulp in 3 DFMAs
__device__ __forceinline__ double rcp_fast3(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double e = fma(-x, r, 1.0);
    const double t = fma(e, e, e);
    return fma(r, t, r);
}

// ------------------------------------------------------------------------------------------
// WENO5-JS on a UNIFORM grid (fast build only).  On a uniform grid the coefficient arrays of
// m_weno.fpp:223-347 reduce to the classical rationals
//   poly_L = {1/3,-5/6 | -1/6,-1/3 | -2/3,1/6}   poly_R = {-1/6,2/3 | 1/3,1/6 | 5/6,-1/3}
//   d_L = {1/10,3/5,3/10}   d_R = {3/10,3/5,1/10}
//   beta = {4/3,-11/3,10/3 | 4/3,-5/3,4/3 | 10/3,-11/3,4/3}
// (mfc_b200_init checks the computed tables against these to 1e-12 before selecting this
// path).  The nonlinear weights only depend on RATIOS of d_k/beta_k^2, so beta is scaled by 3
// and d by 10 (small integers: immediate operands), the weights 6 and 3 are folded into the
// candidate increments, and the face value is formed as v_j + N/D so that rounding in the
// weights only touches the (small) increment.  52 FP64 instructions per variable for both
// faces instead of 66; results differ from the table form at the 1e-16 level.
// eps3 = 3*weno_eps.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void weno5_uniform(const double v[5], double eps3, double &vL, double &vR) {
    const double d1 = v[4] - v[3], d0 = v[3] - v[2], dm1 = v[2] - v[1], dm2 = v[1] - v[0];
    const double p00 = d0*d0, pmm = dm1*dm1;
    const double b0 = fma(d1, fma(4.0, d1, -11.0*d0), fma(10.0, p00, eps3));
    const double b1 = fma(4.0, p00 + pmm, fma(-5.0*d0, dm1, eps3));
    const double b2 = fma(dm2, fma(4.0, dm2, -11.0*dm1), fma(10.0, pmm, eps3));
    const double q0 = b0*b0, q1 = b1*b1, q2 = b2*b2;
    const double B0 = q1*q2, B1 = q0*q2, B2 = q0*q1;
    const double DL = fma(6.0, B1, fma(3.0, B2, B0));          // 10*sum_k d_L(k) B_k
    const double DR = fma(6.0, B1, fma(3.0, B0, B2));
    const double m2 = -2.0*dm1, t2 = 2.0*d0;
    const double eL0 = fma(1.0/3.0, d1, (-5.0/6.0)*d0);        // poly_L0 - v_j
    const double eL1 = m2 - d0;                                // 6 (poly_L1 - v_j)
    const double eL2 = fma(0.5, dm2, m2);                      // 3 (poly_L2 - v_j)
    const double eR0 = fma(-0.5, d1, t2);                      // 3 (poly_R0 - v_j)
    const double eR1 = t2 + dm1;                               // 6 (poly_R1 - v_j)
    const double eR2 = fma(5.0/6.0, dm1, (-1.0/3.0)*dm2);      // poly_R2 - v_j
    const double NL = fma(B0, eL0, fma(B1, eL1, B2*eL2));
    const double NR = fma(B0, eR0, fma(B1, eR1, B2*eR2));
    const double inv = rcp_fast3(DL*DR);
    vL = fma(NL, DR*inv, v[2]);
    vR = fma(NR, DL*inv, v[2]);
}

// ------------------------------------------------------------------------------------------
// WENO5-JS on one 5-cell stencil v[0..4] = v(j-2..j+2), m_weno.fpp:476-531.
// c[] = the 27 grid-dependent coefficients of cell j:
//   c[0..5]  poly_coef_cbL(j,k,q) k-major    c[6..11] poly_coef_cbR
//   c[12..14] d_cbL(k,j)   c[15..17] d_cbR(k,j)   c[18..26] beta_coef(j,k,q) k-major
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void weno5(const double v[5], const double c[27], double eps, double &vL, double &vR) {
    const double dvd1 = v[4] - v[3];                                   // :476-483
    const double dvd0 = v[3] - v[2];
    const double dvdm1 = v[2] - v[1];
    const double dvdm2 = v[1] - v[0];
    double pl0 = v[2] + c[0]*dvd1 + c[1]*dvd0;                         // :485-493
    double pl1 = v[2] + c[2]*dvd0 + c[3]*dvdm1;
    double pl2 = v[2] + c[4]*dvdm1 + c[5]*dvdm2;
    double pr0 = v[2] + c[6]*dvd1 + c[7]*dvd0;                         // :516-524
    double pr1 = v[2] + c[8]*dvd0 + c[9]*dvdm1;
    double pr2 = v[2] + c[10]*dvdm1 + c[11]*dvdm2;
#if MFC_STRICT
    const double b0 = c[18]*dvd1*dvd1 + c[19]*dvd1*dvd0 + c[20]*dvd0*dvd0 + eps;          // :495-506
    const double b1 = c[21]*dvd0*dvd0 + c[22]*dvd0*dvdm1 + c[23]*dvdm1*dvdm1 + eps;
    const double b2 = c[24]*dvdm1*dvdm1 + c[25]*dvdm1*dvdm2 + c[26]*dvdm2*dvdm2 + eps;
    double a0 = c[12]/(b0*b0), a1 = c[13]/(b1*b1), a2 = c[14]/(b2*b2);                    // :508
    double s = a0 + a1 + a2;
    double w0 = a0/s, w1 = a1/s, w2 = a2/s;                                                // :510
    vL = w0*pl0 + w1*pl1 + w2*pl2;                                                         // :514
    a0 = c[15]/(b0*b0); a1 = c[16]/(b1*b1); a2 = c[17]/(b2*b2);                            // :526
    s = a0 + a1 + a2;
    w0 = a0/s; w1 = a1/s; w2 = a2/s;                                                       // :528
    vR = w0*pr0 + w1*pr1 + w2*pr2;                                                         // :531
#else
    // Same weights, algebraically: omega_k = (d_k/beta_k^2)/sum_l(d_l/beta_l^2).  Multiplying
    // numerator and denominator by (beta_0 beta_1 beta_2)^2 gives
    //   omega_k = d_k B_k / sum_l d_l B_l,   B_k = prod_{l != k} beta_l^2,
    // and the two face values share ONE division:  vL = NL*DR/(DL*DR), vR = NR*DL/(DL*DR).
    // 12 FP64 divisions per reconstruction become 1; rounding differs at the 1e-16 level.
    const double p11 = dvd1*dvd1, p10 = dvd1*dvd0, p00 = dvd0*dvd0, p0m = dvd0*dvdm1,
                 pmm = dvdm1*dvdm1, pm2 = dvdm1*dvdm2, p22 = dvdm2*dvdm2;
    const double b0 = fma(c[18], p11, fma(c[19], p10, fma(c[20], p00, eps)));
    const double b1 = fma(c[21], p00, fma(c[22], p0m, fma(c[23], pmm, eps)));
    const double b2 = fma(c[24], pmm, fma(c[25], pm2, fma(c[26], p22, eps)));
    const double q0 = b0*b0, q1 = b1*b1, q2 = b2*b2;
    const double B0 = q1*q2, B1 = q0*q2, B2 = q0*q1;
    const double aL0 = c[12]*B0, aL1 = c[13]*B1, aL2 = c[14]*B2;
    const double aR0 = c[15]*B0, aR1 = c[16]*B1, aR2 = c[17]*B2;
    const double DL = aL0 + aL1 + aL2, DR = aR0 + aR1 + aR2;
    const double NL = fma(aL0, pl0, fma(aL1, pl1, aL2*pl2));
    const double NR = fma(aR0, pr0, fma(aR1, pr1, aR2*pr2));
    const double inv = rcp_fast(DL*DR);
    vL = NL*(DR*inv);
    vR = NR*(DL*inv);
#endif
}

// ------------------------------------------------------------------------------------------
// WENO3 on the 3-cell stencil v[1..3] = v(j-1..j+1), m_weno.fpp:416-466; coefficient slots as
// filled by weno3_cell (weno_coefficients.cpp).  WENO1 (:391-414) is vL = vR = v(j).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void weno3(const double v[5], const double c[27], double eps, double &vL, double &vR) {
    const double dvd0 = v[3] - v[2];                                   // dvd(0)  :420
    const double dvdm1 = v[2] - v[1];                                  // dvd(-1) :422
    const double pl0 = v[2] + c[0]*dvd0, pl1 = v[2] + c[2]*dvdm1;      // :425-428
    const double pr0 = v[2] + c[6]*dvd0, pr1 = v[2] + c[8]*dvdm1;      // :446-449
    const double b0 = c[18]*dvd0*dvd0 + eps;                           // :430-433
    const double b1 = c[21]*dvdm1*dvdm1 + eps;
#if MFC_STRICT
    double a0 = c[12]/(b0*b0), a1 = c[13]/(b1*b1);                     // :435
    double s = a0 + a1;
    double w0 = a0/s, w1 = a1/s;                                       // :437
    vL = w0*pl0 + w1*pl1;                                              // :441
    a0 = c[15]/(b0*b0); a1 = c[16]/(b1*b1);                            // :451
    s = a0 + a1;
    w0 = a0/s; w1 = a1/s;                                              // :453
    vR = w0*pr0 + w1*pr1;                                              // :457
#else
    const double q0 = b0*b0, q1 = b1*b1;                               // omega_k = d_k q_l/(d_0 q_1 + d_1 q_0)
    const double aL0 = c[12]*q1, aL1 = c[13]*q0, aR0 = c[15]*q1, aR1 = c[16]*q0;
    const double DL = aL0 + aL1, DR = aR0 + aR1;
    const double inv = rcp_fast3(DL*DR);
    vL = fma(aL0, pl0, aL1*pl1)*(DR*inv);
    vR = fma(aR0, pr0, aR1*pr1)*(DL*inv);
#endif
}

// ------------------------------------------------------------------------------------------
// HLLC flux at one face, m_riemann_solvers.fpp:136-325.  L/R are the reconstructed primitive
// states [alpha_rho(NF) | vel(ND) | pres | alpha(NF)] left and right of the face; NRM is the
// sweep direction (dir_idx(1), :464-470).  dir_flg is folded in exactly: it is 0 or 1, and
// 1*x + 0*y == x for finite y.
//   F[E]  : flux_rs*_vf          uf : vel_src_rs*_vf(normal) = flux_src_rs*_vf(advxb) (:325)
//   vs[ND]: vel_src_rs*_vf of every direction (:314-324), consumed by the viscous source flux
// ------------------------------------------------------------------------------------------
template <int NF, int ND, int NRM>
__device__ __forceinline__ void hllc(const double *L, const double *R, const double *gam, const double *pinf,
                                     double *F, double &uf, double *vs) {
    constexpr int MOM = NF, EN = NF + ND, ADV = NF + ND + 1;
#if !MFC_STRICT
    // Fast build: the same solver with three exact-arithmetic identities applied.
    //  (1) H - v^2/2 = ((Gamma+1) p + Pi)/rho, so c = sqrt(((Gamma+1) p + Pi)/(rho Gamma)) without
    //      forming E or H (2 divisions + the cancellation H - v^2/2 saved per side).
    //  (2) xi_M, xi_P are the 0/1 indicators of the sign of s_S (:254-255): only the upwind
    //      star state contributes, so only that side is evaluated (selects, no branch).
    //  (3) 1/(s_K - s_S) and 1/(s_K - u_K) share one reciprocal of their product.
    double v2L = 0.0, v2R = 0.0;
#pragma unroll
    for (int i = 0; i < ND; i++) {
        v2L = fma(L[MOM + i], L[MOM + i], v2L);
        v2R = fma(R[MOM + i], R[MOM + i], v2R);
    }
    // mixture rules (:159-167); the sums start from their first term instead of 0
    double rho_L = L[0], gamma_L = L[ADV]*gam[0], pi_inf_L = L[ADV]*pinf[0];
    double rho_R = R[0], gamma_R = R[ADV]*gam[0], pi_inf_R = R[ADV]*pinf[0];
#pragma unroll
    for (int i = 1; i < NF; i++) {
        rho_L = rho_L + L[i];
        gamma_L = fma(L[ADV + i], gam[i], gamma_L);
        pi_inf_L = fma(L[ADV + i], pinf[i], pi_inf_L);
        rho_R = rho_R + R[i];
        gamma_R = fma(R[ADV + i], gam[i], gamma_R);
        pi_inf_R = fma(R[ADV + i], pinf[i], pi_inf_R);
    }
    const double pres_L = L[EN], pres_R = R[EN];
    const double uL = L[MOM + NRM], uR = R[MOM + NRM];
    const double c_L = sqrt_ratio_fast(fma(gamma_L + 1.0, pres_L, pi_inf_L), rho_L*gamma_L);
    const double c_R = sqrt_ratio_fast(fma(gamma_R + 1.0, pres_R, pi_inf_R), rho_R*gamma_R);
    const double s_L = fmin(uL - c_L, uR - c_R);
    const double s_R = fmax(uR + c_R, uL + c_L);
    const double mL = rho_L*(s_L - uL), mR = rho_R*(s_R - uR);
    const double s_S = (pres_R - pres_L + mL*uL - mR*uR)*rcp_fast3(mL - mR);
    // The upwind side is taken by a BRANCH, not by selects: the sign of s_S is uniform over large
    // parts of the flow, so most warps run one side only; selecting ~15 doubles per face costs
    // ~60 issue slots (SEL/FSEL/MOV) that the branch does not.
    auto side = [&](const double *K, double rho, double u, double pres, double s_K, double s_MP,
                    double gamma, double pi_inf, double v2) {
        const double E_K = fma(gamma, pres, pi_inf) + 5e-1*rho*v2;
        const double da = s_K - s_S, db = s_K - u;
        const double rab = rcp_fast3(da*db);
        const double xi = db*db*rab;                     // (s_K - u_K)/(s_K - s_S)
        const double p_over = pres*da*rab;               // p_K/(s_K - u_K)
        const double w = fma(s_MP, xi - 1.0, u);         // u_K + s_MP (xi_K - 1)
#pragma unroll
        for (int i = 0; i < NF; i++) {
            F[i] = K[i]*w;
            F[ADV + i] = K[ADV + i]*w;
        }
#pragma unroll
        for (int i = 0; i < ND; i++) {
            if (i == NRM) F[MOM + i] = fma(rho, fma(u, u, s_MP*fma(xi, s_S, -u)), pres);
            else F[MOM + i] = rho*K[MOM + i]*w;
        }
        F[EN] = fma(u, E_K + pres, s_MP*(fma(xi, fma(s_S - u, fma(rho, s_S, p_over), E_K), -E_K)));
        uf = w;
#pragma unroll
        for (int i = 0; i < ND; i++) vs[i] = i == NRM ? w : K[MOM + i];
    };
    if (!signbit(s_S)) side(L, rho_L, uL, pres_L, s_L, fmin(0.0, s_L), gamma_L, pi_inf_L, v2L);   // xi_M = 1 (:254)
    else side(R, rho_R, uR, pres_R, s_R, fmax(0.0, s_R), gamma_R, pi_inf_R, v2R);
#else
    double vel_L_rms = 0.0, vel_R_rms = 0.0;                            // :138-145
#pragma unroll
    for (int i = 0; i < ND; i++) {
        vel_L_rms = vel_L_rms + L[MOM + i]*L[MOM + i];
        vel_R_rms = vel_R_rms + R[MOM + i]*R[MOM + i];
    }
    const double pres_L = L[EN], pres_R = R[EN];                        // :147-148
    double rho_L = 0.0, gamma_L = 0.0, pi_inf_L = 0.0, rho_R = 0.0, gamma_R = 0.0, pi_inf_R = 0.0;
#pragma unroll
    for (int i = 0; i < NF; i++) {                                      // :159-167
        rho_L = rho_L + L[i];
        gamma_L = gamma_L + L[ADV + i]*gam[i];
        pi_inf_L = pi_inf_L + L[ADV + i]*pinf[i];
        rho_R = rho_R + R[i];
        gamma_R = gamma_R + R[ADV + i]*gam[i];
        pi_inf_R = pi_inf_R + R[ADV + i]*pinf[i];
    }
    const double uL = L[MOM + NRM], uR = R[MOM + NRM];
    const double E_L = gamma_L*pres_L + pi_inf_L + 5e-1*rho_L*vel_L_rms;   // :202
    const double E_R = gamma_R*pres_R + pi_inf_R + 5e-1*rho_R*vel_R_rms;   // :204
    const double H_L = (E_L + pres_L)/rho_L;                            // :206-207
    const double H_R = (E_R + pres_R)/rho_R;
    const double c_L = sqrt((H_L - 5e-1*vel_L_rms)/gamma_L);            // :220-223
    const double c_R = sqrt((H_R - 5e-1*vel_R_rms)/gamma_R);
    const double s_L = fmin(uL - c_L, uR - c_R);                        // :232-233
    const double s_R = fmax(uR + c_R, uL + c_L);
    const double s_S = (pres_R - pres_L + rho_L*uL*(s_L - uL) - rho_R*uR*(s_R - uR))   // :235-240
                       /(rho_L*(s_L - uL) - rho_R*(s_R - uR));
    const double s_M = fmin(0.0, s_L), s_P = fmax(0.0, s_R);            // :245
    const double xi_L = (s_L - uL)/(s_L - s_S);                         // :249-250
    const double xi_R = (s_R - uR)/(s_R - s_S);
    const double xi_M = (5e-1 + copysign(5e-1, s_S));                   // :254-255
    const double xi_P = (5e-1 - copysign(5e-1, s_S));
#pragma unroll
    for (int i = 0; i < NF; i++)                                        // :258-264
        F[i] = xi_M*L[i]*(uL + s_M*(xi_L - 1.0)) + xi_P*R[i]*(uR + s_P*(xi_R - 1.0));
#pragma unroll
    for (int i = 0; i < ND; i++) {                                      // :270-286
        if (i == NRM)
            F[MOM + i] = xi_M*(rho_L*(uL*L[MOM + i] + s_M*(xi_L*(s_S) - L[MOM + i])) + (pres_L))
                         + xi_P*(rho_R*(uR*R[MOM + i] + s_P*(xi_R*(s_S) - R[MOM + i])) + (pres_R));
        else
            F[MOM + i] = xi_M*(rho_L*(uL*L[MOM + i] + s_M*(xi_L*(L[MOM + i]) - L[MOM + i])))
                         + xi_P*(rho_R*(uR*R[MOM + i] + s_P*(xi_R*(R[MOM + i]) - R[MOM + i])));
    }
    F[EN] = xi_M*(uL*(E_L + pres_L)                                     // :291-299
                  + s_M*(xi_L*(E_L + (s_S - uL)*(rho_L*s_S + pres_L/(s_L - uL))) - E_L))
            + xi_P*(uR*(E_R + pres_R)
                    + s_P*(xi_R*(E_R + (s_S - uR)*(rho_R*s_S + pres_R/(s_R - uR))) - E_R));
#pragma unroll
    for (int i = 0; i < NF; i++)                                        // :304-310
        F[ADV + i] = xi_M*L[ADV + i]*(uL + s_M*(xi_L - 1.0)) + xi_P*R[ADV + i]*(uR + s_P*(xi_R - 1.0));
    uf = xi_M*(uL + s_M*(xi_L - 1.0)) + xi_P*(uR + s_P*(xi_R - 1.0));   // :316-325
#pragma unroll
    for (int i = 0; i < ND; i++) vs[i] = i == NRM ? uf : xi_M*L[MOM + i] + xi_P*R[MOM + i];
#endif
}

// Face Reynolds numbers for the viscous source flux, m_riemann_solvers.fpp:169-200,225-230:
// Re_K(i) = 1/max(sum_q alpha_K(Re_idx(i,q))/Res(i,q), sgm_eps), Re_avg = 2/(1/Re_L + 1/Re_R).
template <int NF, int ND>
__device__ __forceinline__ void face_reynolds(const double *L, const double *R, const SweepArgs &a, double Re_avg[2]) {
    constexpr int ADV = NF + ND + 1;
#pragma unroll
    for (int i = 0; i < 2; i++) {
        double rl = a.Re_size[i] > 0 ? 0.0 : -1e6, rr = rl;
        for (int q = 0; q < a.Re_size[i]; q++) {
            double al = L[ADV], ar = R[ADV];
#pragma unroll
            for (int f = 1; f < NF; f++)
                if (a.Re_idx[i][q] == f) { al = L[ADV + f]; ar = R[ADV + f]; }
            rl = al/a.Res[i][q] + rl;
            rr = ar/a.Res[i][q] + rr;
        }
        const double Re_L = 1.0/fmax(rl, 1e-16), Re_R = 1.0/fmax(rr, 1e-16);
        Re_avg[i] = 2.0/(1.0/Re_L + 1.0/Re_R);
    }
}
// vel_src (ND planes) and Re_avg (2 planes) of the face whose LEFT cell has in-plane offset off.
// Fast build: 1/Re_avg = (max(sum alpha_L/Res, eps) + max(sum alpha_R/Res, eps))/2 is stored instead
// (the same quantity, :225-230, without its five divisions); k_visc multiplies by it.
template <int NF, int ND>
__device__ __forceinline__ void store_visc_face(const SweepArgs &a, unsigned off, const double *L, const double *R, const double *vs) {
    constexpr int ADV = NF + ND + 1;
    double Re_avg[2];
#if MFC_STRICT
    face_reynolds<NF, ND>(L, R, a, Re_avg);
#else
#pragma unroll
    for (int i = 0; i < 2; i++) {
        double rl = a.Re_size[i] > 0 ? 0.0 : -1e6, rr = rl;
        for (int q = 0; q < a.Re_size[i]; q++) {
            double al = L[ADV], ar = R[ADV];
#pragma unroll
            for (int f = 1; f < NF; f++)
                if (a.Re_idx[i][q] == f) { al = L[ADV + f]; ar = R[ADV + f]; }
            rl = fma(al, a.iRes[i][q], rl);
            rr = fma(ar, a.iRes[i][q], rr);
        }
        Re_avg[i] = 0.5*(fmax(rl, 1e-16) + fmax(rr, 1e-16));
    }
#endif
    const long long fs = a.g.fstride;
#pragma unroll
    for (int i = 0; i < ND; i++) a.visc_face[i*fs + off] = vs[i];
    a.visc_face[ND*fs + off] = Re_avg[0];
    a.visc_face[(ND + 1)*fs + off] = Re_avg[1];
}

// Fast build, weno_Re_flux = F: the viscous source flux of one face of direction ID
// (m_riemann_solvers.fpp:683-902 with the finite-difference gradients of m_viscous.fpp:219-347),
// computed INSIDE the sweep right after the Riemann solve and folded into the flux, so that the
// flux difference of m_rhs.fpp:565-653 carries the viscous terms of :591-604, :639-652 with it: no
// face planes, no separate pass, and the TVD-RK statement stays fused into the last sweep.
//   dn[v] = d vel_v / d x_ID at the face (difference of the two cell centres, m_viscous.fpp:224-248)
//   ct[v] = d vel_v / d x_other at the face = mean of the two cells' averaged one-sided differences
//           (:278-347), which k_vgrad prepared per cell
//   vs    = vel_src of the face (m_riemann_solvers.fpp:314-324), L / R = the face states (alpha for Re)
template <int NF, int ND, int ID>
__device__ __forceinline__ void visc_flux_fd(const SweepArgs &a, const double *L, const double *R, const double *vs,
                                             const double *dn, const double *ct, double *F) {
    constexpr int MOM = NF, EN = NF + ND, ADV = NF + ND + 1, O = ND > 1 ? 1 - ID : 0;
    // 1/Re_avg = (1/Re_L + 1/Re_R)/2 with 1/Re_K = max(sum_f alpha_K(f)/Re(f), sgm_eps), :169-200, :225-230;
    // iRe_f holds 1/Re per FLUID (0 for the fluids without that viscosity): no index lists, no branches
    double iRe[2];
#pragma unroll
    for (int i = 0; i < 2; i++) {
        double rl = L[ADV]*a.iRe_f[i][0], rr = R[ADV]*a.iRe_f[i][0];
#pragma unroll
        for (int f = 1; f < NF; f++) {
            rl = fma(L[ADV + f], a.iRe_f[i][f], rl);
            rr = fma(R[ADV + f], a.iRe_f[i][f], rr);
        }
        iRe[i] = 0.5*(fmax(rl, 1e-16) + fmax(rr, 1e-16));
    }
    double avg[ND > 1 ? 2 : 1][ND];                    // avg[dd][v] = d vel_v / d x_dd
#pragma unroll
    for (int v = 0; v < ND; v++) {
        avg[ND > 1 ? ID : 0][v] = dn[v];
        if (ND > 1) avg[O][v] = ct[v];
    }
    double m[ND], e = 0.0;
#pragma unroll
    for (int i = 0; i < ND; i++) m[i] = 0.0;
    constexpr int Y = ND > 1 ? 1 : 0;
    if (ID == 0) {
        if (a.Re_size[0] > 0) {                        // :714-736
            const double tau = (4.0/3.0)*avg[0][0]*iRe[0];
            m[0] -= tau; e = fma(-vs[0], tau, e);
        }
        if (a.Re_size[1] > 0) {                        // :738-760
            const double tau = avg[0][0]*iRe[1];
            m[0] -= tau; e = fma(-vs[0], tau, e);
        }
        if (ND > 1) {
            if (a.Re_size[0] > 0) {                    // :764-801
                const double t0 = -(2.0/3.0)*avg[Y][Y]*iRe[0];
                const double t1 = (avg[Y][0] + avg[0][Y])*iRe[0];
                m[0] -= t0; e = fma(-vs[0], t0, e);
                m[Y] -= t1; e = fma(-vs[Y], t1, e);
            }
            if (a.Re_size[1] > 0) {                    // :803-825
                const double tau = avg[Y][Y]*iRe[1];
                m[0] -= tau; e = fma(-vs[0], tau, e);
            }
        }
    } else {
        if (a.Re_size[0] > 0) {                        // :831-872
            const double t0 = (avg[Y][0] + avg[0][Y])*iRe[0];
            const double t1 = (4.0*avg[Y][Y] - 2.0*avg[0][0])*((1.0/3.0)*iRe[0]);
            m[0] -= t0; e = fma(-vs[0], t0, e);
            m[Y] -= t1; e = fma(-vs[Y], t1, e);
        }
        if (a.Re_size[1] > 0) {                        // :874-899
            const double tau = (avg[0][0] + avg[Y][Y])*iRe[1];
            m[Y] -= tau; e = fma(-vs[Y], tau, e);
        }
    }
#pragma unroll
    for (int i = 0; i < ND; i++) F[MOM + i] += m[i];
    F[EN] += e;
}

// pull the line of p into L2 (no register, no scoreboard): issued one iteration before the load
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ void load_coef(const SweepArgs &a, int cell, double c[27]) {
    const double *p = a.coef + (cell - a.coef_lo);
#pragma unroll
    for (int i = 0; i < 27; i++) c[i] = __ldg(p + (long long)i*a.clen);
}

// ==========================================================================================
// Sweep kernels: the stage state streams HBM -> shared memory through the TMA engine
// (cp.async.bulk, SASS UBLKCP) into a ring of row slots guarded by mbarriers; every value is
// fetched from HBM exactly once per sweep, the conservative -> primitive conversion
// (m_variables_conversion.fpp:326-373) happens in place in the ring (no q_prim planes in HBM,
// no k_prim launch on the hot path), and the 5-point stencils are read with conflict-free
// LDS.64.  COEF = 0: uniform grid, the 27 WENO coefficients are kernel-parameter constants
// (constant-bank operands of the DFMAs); COEF = 1: per-cell coefficient tables.
// ==========================================================================================
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *b, unsigned n) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *b, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "MFC_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra MFC_DONE;\n"
        "bra MFC_WAIT;\n"
        "MFC_DONE:\n"
        "}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
// one row of all E variables (box {W, 1, 1, E} of the 4-D tensor (x, y, z, variable)) -> smem
__device__ __forceinline__ void tma_load_row(void *dst, const TensorMap *tm, int c0, int c1, int c2, unsigned long long *b) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(0), "r"(smem_u32(b)) : "memory");
}
// one row of all E variables, shared memory -> global (SASS UTMASTG); elements outside the tensor
// (columns beyond the last interior cell of an interior-clipped map) are not written
__device__ __forceinline__ void tma_store_row(const TensorMap *tm, int c0, int c1, int c2, const void *src) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4}], [%5];"
                 ::"l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(0), "r"(smem_u32(src)) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// every bulk store committed by this thread has finished READING its shared-memory source
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// generic-proxy accesses to a slot must be ordered before the async proxy overwrites it
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// tuning knobs (tools/tune_variants.py builds and times alternatives)
#ifndef MFC_RING_Y
#define MFC_RING_Y 7
#define MFC_WARPS_Y 4
#define MFC_CTAS_Y 3
#endif
#ifndef MFC_RING_Z
#define MFC_RING_Z 6
#define MFC_CTAS_Z 3
#endif
#ifndef MFC_RING_X
#define MFC_RING_X 3
#define MFC_WARPS_X 4
#define MFC_CTAS_X 4
#endif
constexpr int kRingY = MFC_RING_Y;   // row slots of the y/z march ring: 5 live rows + the rows in flight
constexpr int kRingX = MFC_RING_X;   // box slots of the x kernel: up to 2 per chunk + the boxes in flight
// doubles per ring slot: E rows of W doubles, padded so that every slot starts on a 128-byte line
// (the TMA destination alignment)
__host__ __device__ constexpr int slot_doubles(int E, int W) { return (E*W + 15)/16*16; }
// Occupancy (measured at 512^3, profiles/r01_tune_occupancy.txt, r01_v3b_tune_pingpong_occupancy.txt,
// r01_v4_tune.txt): x 5 CTAs x 4 warps (<= 96 registers, 40 KB of rings per CTA); y and z 3 CTAs x
// 4 warps (<= 168 registers, 72 KB per CTA: ring + the operand / output slots of k_march3).
constexpr int kWarpsX = MFC_WARPS_X, kCtasX = MFC_CTAS_X;
constexpr int kWarpsY = MFC_WARPS_Y, kCtasY = MFC_CTAS_Y;
// the last direction carries the fused RK update (more live registers): its own occupancy knobs
__host__ __device__ constexpr int march_ring(int dir, int nd) { return dir == nd - 1 ? MFC_RING_Z : kRingY; }
__host__ __device__ constexpr int march_ctas(int dir, int nd) { return dir == nd - 1 ? MFC_CTAS_Z : kCtasY; }
// ring + operand / output slots per warp of k_march3: ring | rhs in | q1 in (last direction) | out
__host__ __device__ constexpr int march_slots(int dir, int nd) { return march_ring(dir, nd) + 2 + (dir == nd - 1 ? 1 : 0); }

// cons -> prim of one cell held in a ring slot (stride LD between variables), in place:
// momenta become velocities, the energy becomes the pressure (:187-227, :353-362, :98-106)
template <int NF, int ND, int LD>
__device__ __forceinline__ void prim_in_place(double *cellp, const double *gam, const double *pinf) {
    constexpr int MOM = NF, EN = NF + ND, ADV = NF + ND + 1;
#if MFC_STRICT
    double rho = 0.0, gamma = 0.0, pi_inf = 0.0;
#pragma unroll
    for (int i = 0; i < NF; i++) {
        const double al = cellp[(ADV + i)*LD];
        rho = rho + cellp[i*LD];
        gamma = gamma + al*gam[i];
        pi_inf = pi_inf + al*pinf[i];
    }
#else
    double rho = cellp[0], gamma = cellp[ADV*LD]*gam[0], pi_inf = cellp[ADV*LD]*pinf[0];
#pragma unroll
    for (int i = 1; i < NF; i++) {
        const double al = cellp[(ADV + i)*LD];
        rho = rho + cellp[i*LD];
        gamma = fma(al, gam[i], gamma);
        pi_inf = fma(al, pinf[i], pi_inf);
    }
#endif
    rho = fmax(rho, 1e-16);
    double dyn = 0.0;
#if MFC_STRICT
#pragma unroll
    for (int i = 0; i < ND; i++) {
        const double mom = cellp[(MOM + i)*LD];
        const double u = mom/rho;
        cellp[(MOM + i)*LD] = u;
        dyn = dyn + 5e-1*mom*u;
    }
    cellp[EN*LD] = (cellp[EN*LD] - dyn - pi_inf)/gamma;
#else
    const double ir = rcp_fast3(rho), ig = rcp_fast3(gamma);
#pragma unroll
    for (int i = 0; i < ND; i++) {
        const double mom = cellp[(MOM + i)*LD];
        const double u = mom*ir;
        cellp[(MOM + i)*LD] = u;
        dyn = fma(5e-1*mom, u, dyn);
    }
    cellp[EN*LD] = (cellp[EN*LD] - dyn - pi_inf)*ig;
#endif
}

// reconstruction of one variable at one cell: uniform-grid constants (COEF = 0, fast build only)
// or the per-cell coefficient tables (m_weno.fpp:223-347)
// WO = weno_order (5, 3 or 1; the lower orders always use the tables)
template <int COEF, int WO>
struct Weno {
    static constexpr bool kUniform = COEF == 0 && !MFC_STRICT && WO == 5;
    double c[(kUniform || WO == 1) ? 1 : 27];
    double eps;
    __device__ __forceinline__ void load(const SweepArgs &a, int cell) {
        if (kUniform) {
            eps = 3.0*a.eps;
        } else {
            if (WO != 1) load_coef(a, cell, c);
            eps = a.eps;
        }
    }
    __device__ __forceinline__ void operator()(const double s[5], double &vL, double &vR) const {
        if (WO == 1) { vL = s[2]; vR = s[2]; }
        else if (WO == 3) weno3(s, c, eps, vL, vR);
        else if (kUniform) weno5_uniform(s, eps, vL, vR);
        else weno5(s, c, eps, vL, vR);
    }
};

// operands of finish_cell2 that stream from HBM, fetched ahead of the Riemann solve.
// ACC: the RHS of earlier directions is accumulated (m_rhs.fpp:610-620); RK: this is the last
// direction, the TVD-RK statement a.rk_mode (1..4, or 0 = store the RHS) is fused in.
// Cells are addressed as (per-variable plane pointer from the constant bank) + (32-bit element
// offset within the plane): one IMAD.WIDE per access.
//
// Strict build: q_cons_ts(1) and the stage state are read back from HBM for the update.
// Fast build: the stage state's momenta and energy are rebuilt from the primitive variables
// held in the ring (mom = rho u, E = Gamma p + Pi + rho |u|^2/2; 1-ulp round trip), so stage 1
// (q1 == stage state) reads nothing and stages 2/3 read only q_cons_ts(1).
template <int E, bool ACC, bool RK>
struct CellIn {
    double r[ACC ? E : 1], q1[RK ? E : 1], qs[RK ? E : 1];
};
template <int NF, int ND, bool ACC, bool RK>
__device__ __forceinline__ void load_cell(const SweepArgs &a, unsigned off, CellIn<2*NF + ND + 1, ACC, RK> &in) {
    constexpr int E = 2*NF + ND + 1, ADV = NF + ND + 1;
    if (ACC) {
#pragma unroll
        for (int v = 0; v < E; v++) in.r[v] = __ldg(a.rhs_v[v] + off);
    }
    if (RK) {
#if MFC_STRICT
        if (a.rk_mode != 0) {
#pragma unroll
            for (int v = 0; v < E; v++) in.q1[v] = __ldg(a.q1_v[v] + off);
        }
        if (a.rk_mode >= 2) {
#pragma unroll
            for (int v = 0; v < ADV; v++) in.qs[v] = __ldg(a.q_v[v] + off);
        }
#else
        if (a.rk_mode >= 2) {
#pragma unroll
            for (int v = 0; v < E; v++) in.q1[v] = __ldg(a.q1_v[v] + off);
        }
#endif
    }
}
// RHS of one cell + fused RK stage, returned in y[] (the RHS itself when !RK or rk_mode == 0,
// else the updated state).  pc = the cell's entry in the ring (stride LD between variables:
// partial densities and volume fractions as stored, velocities and pressure in place of momenta
// and energy); in = the operands fetched by load_cell / from the operand slots.
template <int NF, int ND, int LD, bool ACC, bool RK>
__device__ __forceinline__ void finish_vals(const SweepArgs &a, double rds, const double *pc,
                                            const CellIn<2*NF + ND + 1, ACC, RK> &in,
                                            const double *Fm, double ufm, const double *Fp, double ufp,
                                            double (&y)[2*NF + ND + 1]) {
    constexpr int E = 2*NF + ND + 1, MOM = NF, EN = NF + ND, ADV = NF + ND + 1;
    double x[E], al[NF];
#pragma unroll
    for (int i = 0; i < NF; i++) al[i] = pc[(ADV + i)*LD];
#if MFC_STRICT
#pragma unroll
    for (int v = 0; v < E; v++) {
        x[v] = rds*(Fm[v] - Fp[v]);
        if (ACC) x[v] = in.r[v] + x[v];
        if (v >= ADV) x[v] = x[v] + rds*al[v - ADV]*(ufp - ufm);
    }
#else
    const double du = rds*(ufp - ufm);
#pragma unroll
    for (int v = 0; v < E; v++) {
        const double dF = Fm[v] - Fp[v];
        x[v] = ACC ? fma(rds, dF, in.r[v]) : rds*dF;
        if (v >= ADV) x[v] = fma(al[v - ADV], du, x[v]);
    }
#endif
    if (!RK || a.rk_mode == 0) {
#pragma unroll
        for (int v = 0; v < E; v++) y[v] = x[v];
        return;
    }
#if MFC_STRICT
#pragma unroll
    for (int v = 0; v < E; v++) {
        const double qs = a.rk_mode >= 2 ? (v >= ADV ? al[v - ADV] : in.qs[v]) : 0.0;
        y[v] = rk_apply(a.rk_mode, in.q1[v], qs, x[v], a.dt);
    }
#else
    // stage state rebuilt from the ring
    double qs[E], v2 = 0.0;
#pragma unroll
    for (int i = 0; i < NF; i++) {
        qs[i] = pc[i*LD];
        qs[ADV + i] = al[i];
    }
    double rho = qs[0], gamma = al[0]*a.gammas[0], pi_inf = al[0]*a.pi_infs[0];
#pragma unroll
    for (int i = 1; i < NF; i++) {
        rho += qs[i];
        gamma = fma(al[i], a.gammas[i], gamma);
        pi_inf = fma(al[i], a.pi_infs[i], pi_inf);
    }
    rho = fmax(rho, 1e-16);
#pragma unroll
    for (int i = 0; i < ND; i++) {
        const double u = pc[(MOM + i)*LD];
        qs[MOM + i] = rho*u;
        v2 = fma(u, u, v2);
    }
    qs[EN] = fma(gamma, pc[EN*LD], pi_inf) + 5e-1*rho*v2;
    // every TVD-RK statement is (c1 q1 + c2 qs + c3 dt rhs)*c4   (m_time_steppers.fpp:167,245,322,342),
    // evaluated as k1 q1 + k2 qs + k3 rhs with the constants folded
    const int m = a.rk_mode;
    if (m == 1) {
#pragma unroll
        for (int v = 0; v < E; v++) y[v] = fma(a.dt, x[v], qs[v]);
    } else {
        const double c4 = m == 2 ? 0.5 : (m == 3 ? 0.25 : 1.0/3.0);
        const double k1 = (m == 3 ? 3.0 : 1.0)*c4, k2 = (m == 4 ? 2.0 : 1.0)*c4, k3 = (m == 4 ? 2.0 : 1.0)*a.dt*c4;
#pragma unroll
        for (int v = 0; v < E; v++) y[v] = fma(k1, in.q1[v], fma(k2, qs[v], k3*x[v]));
    }
#endif
}
// the same with direct (masked) stores: store = false masks the stores (lanes beyond the domain,
// halo lanes of the x kernel)
template <int NF, int ND, int LD, bool ACC, bool RK>
__device__ __forceinline__ void finish_cell2(const SweepArgs &a, unsigned off, bool store, double rds, const double *pc,
                                             const CellIn<2*NF + ND + 1, ACC, RK> &in,
                                             const double *Fm, double ufm, const double *Fp, double ufp) {
    constexpr int E = 2*NF + ND + 1;
    double y[E];
    finish_vals<NF, ND, LD, ACC, RK>(a, rds, pc, in, Fm, ufm, Fp, ufp, y);
    if (store) {
        if (!RK || a.rk_mode == 0) {
#pragma unroll
            for (int v = 0; v < E; v++) a.rhsw_v[v][off] = y[v];
        } else {
#pragma unroll
            for (int v = 0; v < E; v++) a.qout_v[v][off] = y[v];
        }
    }
}

// ------------------------------------------------------------------------------------------
// x sweep.  Every WARP is an independent pipeline over a STREAM of cells: its work item is an x
// tile [ja, jb] of `nr` consecutive rows, laid end to end as rows of Ls = max(jb-ja+7, 32)
// stream entries (cells ja-3 .. jb+3 of a row, then the next row).  The warp walks the stream in
// chunks of 32 entries, lane = entry, so that all 32 lanes do useful work whatever the row length
// (a 512-cell row is 518 entries; the first kernel of this file's history finished 30 cells per
// 32 lanes and needed 576 lane slots for it).  The stages of the scheme follow each other through
// small per-warp rings in shared memory, each stage working on the entry its inputs are ready for:
//   entry g     cons -> prim of the lane's own cell, raw values from the TMA box, result into the
//               primitive ring (40 entries per variable)
//   entry g-2   WENO reconstruction from ring entries g-4 .. g; the right-face state goes into an
//               exchange ring (33 entries: the slot the 32 lanes leave untouched in a chunk is
//               lane 31's value of the previous chunk, which lane 0 reads), the lane reads the
//               state of entry g-3 and solves the Riemann problem at the face between g-3 and g-2
//   entry g-3   the flux crosses lanes through a second exchange ring the same way; the lane
//               finishes cell g-3 (flux difference, source term, 1-D: the TVD-RK statement)
// Every shared-memory hand-over is one STS + one LDS per value (a 64-bit warp shuffle is two
// instructions plus a select for the chunk boundary), the conversion runs once per cell, and the
// boxes of consecutive chunks do not overlap.  This matters because the sweeps are ISSUE bound:
// an FP64 warp instruction holds its scheduler's issue port for two cycles, so a kernel's time
// is (2 x FP64 + other instructions) per scheduler -- see DESIGN.md 4.
// A chunk lies in one row or, when it wraps, in two: each part is one tensor-map bulk copy of
// 34 columns x E variables into a ring slot (box start rounded down to an even column, the TMA
// unit wants 16-byte aligned starts; the part of a box before a row's first ghost column is
// zero-filled and read by no lane that stores), and a lane picks the slot of its own row.
// No block-wide barrier exists, so warps never wait for each other.
// BC4: some side of this direction has bc = -4 (Riemann-state extrapolation).  VISC = 1: viscous run,
// vel_src and Re_avg of every face are stored for k_visc; VISC = 2 (fast build, weno_Re_flux = F): the
// viscous source flux is computed here and folded into the flux (visc_flux_fd).
// ------------------------------------------------------------------------------------------
constexpr int kPrimRing = 40;                      // entries of the primitive ring (>= 32 + 4, multiple of 8)
constexpr int kXchRing = 33;                       // entries of the exchange rings (32 lanes + the carried one)
__host__ __device__ constexpr int xstream_warp_doubles(int E) {   // per-warp shared memory besides the TMA ring
    return E*kPrimRing + E*kXchRing + (E + 1)*kXchRing;
}
__host__ __device__ constexpr size_t xstream_smem(int E) {
    return (size_t)kWarpsX*((kRingX*slot_doubles(E, kWX) + xstream_warp_doubles(E))*sizeof(double) + kRingX*sizeof(unsigned long long));
}
// cons -> prim of one cell in registers (m_variables_conversion.fpp:187-227, :353-362, :98-106);
// the same statements as prim_in_place
template <int NF, int ND>
__device__ __forceinline__ void prim_regs(double (&q)[2*NF + ND + 1], const double *gam, const double *pinf) {
    constexpr int MOM = NF, EN = NF + ND, ADV = NF + ND + 1;
#if MFC_STRICT
    double rho = 0.0, gamma = 0.0, pi_inf = 0.0;
#pragma unroll
    for (int i = 0; i < NF; i++) {
        rho = rho + q[i];
        gamma = gamma + q[ADV + i]*gam[i];
        pi_inf = pi_inf + q[ADV + i]*pinf[i];
    }
    rho = fmax(rho, 1e-16);
    double dyn = 0.0;
#pragma unroll
    for (int i = 0; i < ND; i++) {
        const double mom = q[MOM + i];
        const double u = mom/rho;
        q[MOM + i] = u;
        dyn = dyn + 5e-1*mom*u;
    }
    q[EN] = (q[EN] - dyn - pi_inf)/gamma;
#else
    double rho = q[0], gamma = q[ADV]*gam[0], pi_inf = q[ADV]*pinf[0];
#pragma unroll
    for (int i = 1; i < NF; i++) {
        rho = rho + q[i];
        gamma = fma(q[ADV + i], gam[i], gamma);
        pi_inf = fma(q[ADV + i], pinf[i], pi_inf);
    }
    rho = fmax(rho, 1e-16);
    double dyn = 0.0;
    const double ir = rcp_fast3(rho), ig = rcp_fast3(gamma);
#pragma unroll
    for (int i = 0; i < ND; i++) {
        const double mom = q[MOM + i];
        const double u = mom*ir;
        q[MOM + i] = u;
        dyn = fma(5e-1*mom, u, dyn);
    }
    q[EN] = (q[EN] - dyn - pi_inf)*ig;
#endif
}

template <int NF, int ND, int COEF, bool BC4, int VISC, int WO>
__global__ void __launch_bounds__(32*kWarpsX, (COEF == 1 && WO == 5 && !MFC_STRICT) ? 3 : kCtasX) k_xstream(const __grid_constant__ SweepArgs a) {
    constexpr int E = 2*NF + ND + 1, ADV = NF + ND + 1, R = kRingX, SLOT = slot_doubles(E, kWX);
    constexpr int PR = kPrimRing, XR = kXchRing;
    constexpr bool ACC = false, RK = ND == 1;          // x is the first direction, and the last one in 1-D
    static_assert(R >= 3, "a chunk needs up to two boxes, plus at least one in flight");
    static_assert(kWX >= 34 && kWX % 2 == 0, "32 columns + one for the even-start rounding");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const GridDesc &g = a.g;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *ring = reinterpret_cast<double *>(smem_raw) + warp*(R*SLOT);
    double *prim = reinterpret_cast<double *>(smem_raw) + kWarpsX*R*SLOT + warp*xstream_warp_doubles(E);
    double *xv = prim + E*PR, *xf = xv + E*XR;
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(reinterpret_cast<double *>(smem_raw) + kWarpsX*(R*SLOT + xstream_warp_doubles(E))) + warp*R;
    // work item of this warp: x tile tix of row block rb in plane blockIdx.z
    const int wi = (int)blockIdx.x*kWarpsX + warp;
    if (wi >= a.xs_items) return;                      // no block barriers below
    const int tix = wi % a.xs_ntx, rb = wi/a.xs_ntx;
    int ja, jb;
    if (a.xs_strip > 0) {                              // the two boundary strips of a split launch
        ja = tix == 0 ? 0 : g.N[0] - a.xs_strip + 1;
        jb = tix == 0 ? a.xs_strip - 1 : g.N[0];
    } else {                                           // balanced tiles over cells xs_lo .. xs_hi
        const long long W = a.xs_hi - a.xs_lo + 1;
        ja = a.xs_lo + (int)(tix*W/a.xs_ntx);
        jb = a.xs_lo + (int)((tix + 1)*W/a.xs_ntx) - 1;
    }
    const int k0 = a.xs_k0 + rb*a.rows, l = (int)blockIdx.z + a.xs_z0;
    const int nr = min(a.rows, a.xs_k0 + a.xs_nk - k0);
    const int Lt = jb - ja + 7;                        // entries of a row that exist: cells ja-3 .. jb+3
    const int Ls = max(Lt, 32);                        // a chunk touches at most two rows
    const int nload = (nr*Ls + 31) >> 5;               // chunks that bring new cells
    const int nch = (nr*Ls + 3 + 31) >> 5;             // + the pipeline drain (the finish lags 3 entries)
    if (lane == 0) {
        for (int i = 0; i < R; i++) mbar_init(&bar[i], 1);
        mbar_fence_init();
    }
    __syncwarp();
    // issue cursor: the next box to request belongs to the chunk whose lane 0 sits at entry ii of
    // row ir; ipart = 1: its second part (the chunk wraps into row ir+1)
    int ir = 0, ii = 0, ipart = 0, ichunk = 0, islot = 0, inflight = 0;
    const int jbase = ja - 3 + kXoff;                  // tensor column of a row's entry 0
    auto issue_one = [&]() {
        const bool wrap = ii + 31 >= Ls && ir + 1 < nr;
        if (lane == 0) {
            mbar_expect_tx(&bar[islot], (unsigned)(E*kWX*sizeof(double)));
            tma_load_row(ring + islot*SLOT, &a.tm_q, (jbase + ii - (ipart ? Ls : 0)) & ~1, k0 + ir + ipart + g.yoff, l + g.zoff, &bar[islot]);
        }
        if (++islot == R) islot = 0;
        inflight++;
        if (!ipart && wrap) ipart = 1;
        else {
            ipart = 0; ichunk++;
            ii += 32;
            if (ii >= Ls) { ii -= Ls; ir++; }
        }
    };
    while (inflight < R && ichunk < nload) issue_one();

    const bool stab_on = a.stab_out != nullptr;
    double icfl = 0.0;
    const unsigned usy = (unsigned)g.sy;
    const unsigned off_row0 = (unsigned)g.at(0, k0, l);
    int r0 = 0, i0 = 0, cslot = 0;                     // consume cursor: lane 0's row / entry, first slot
    unsigned cphase = 0;
    int il = lane, row = 0;                            // my entry g: position in its row, row
    int ppos = lane;                                   // its slot in the primitive ring
    int wpos = lane;                                   // my slot in the exchange rings
#pragma unroll 1
    for (int c = 0; c < nch; c++) {
        // ---- entry g: conversion ------------------------------------------------------------
        if (c < nload) {
            const bool wrap = i0 + 31 >= Ls && r0 + 1 < nr;
            int slotB = cslot + 1;
            unsigned phaseB = cphase;
            if (slotB == R) { slotB = 0; phaseB ^= 1u; }
            mbar_wait(&bar[cslot], cphase);
            if (wrap) mbar_wait(&bar[slotB], phaseB);
            const int cA = jbase + i0;
            const bool inB = wrap && i0 + lane >= Ls;
            const double *src = (inB ? ring + slotB*SLOT + ((cA - Ls) & 1) : ring + cslot*SLOT + (cA & 1)) + lane;
            double q[E];
#pragma unroll
            for (int v = 0; v < E; v++) q[v] = src[v*kWX];
            prim_regs<NF, ND>(q, a.gammas, a.pi_infs);
#pragma unroll
            for (int v = 0; v < E; v++) prim[v*PR + ppos] = q[v];
            fence_proxy_async();
            __syncwarp();                              // ring entries visible; the warp is done with the boxes
            const int nb = wrap ? 2 : 1;
            inflight -= nb;
            cslot += nb;
            if (cslot >= R) { cslot -= R; cphase ^= 1u; }
            i0 += 32;
            if (i0 >= Ls) { i0 -= Ls; r0++; }
            while (inflight < R && ichunk < nload) issue_one();
        } else {
            __syncwarp();
        }
        // ---- entry g-2: reconstruction, Riemann problem at its left face -------------------------
        int il2 = il - 2, row2 = row;
        if (il2 < 0) { il2 += Ls; row2--; }
        int il3 = il - 3, row3 = row;
        if (il3 < 0) { il3 += Ls; row3--; }
        const int j2 = ja - 3 + il2, j3 = ja - 3 + il3;
        const bool row3_ok = row3 >= 0 && row3 < nr;
        const bool store_on = row3_ok && il3 >= 3 && il3 <= Lt - 4;      // cells ja .. jb
        const int jf = min(max(j3, -1), g.N[0]);       // clamped lanes never store
        const unsigned off = off_row0 + (unsigned)min(max(row3, 0), nr - 1)*usy + (unsigned)jf;   // two's complement for jf = -1
        int tp[5];                                     // ring slots of entries g-4 .. g
#pragma unroll
        for (int t = 0; t < 5; t++) {
            tp[t] = ppos - 4 + t;
            if (tp[t] < 0) tp[t] += PR;
        }
        Weno<COEF, WO> weno;
        weno.load(a, min(max(j2, -1), g.N[0] + 1));
        const double rds = a.rds[max(jf, 0) + g.b];
        CellIn<E, ACC, RK> in;
        if (RK) load_cell<NF, ND, ACC, RK>(a, off, in);
        // VISC == 2: the transverse velocity gradients of the face's two cells (planes [1][v] of
        // k_vgrad), requested here so that the loads fly under the reconstruction
        double vg_l[ND > 1 ? ND : 1], vg_r[ND > 1 ? ND : 1];
        if (VISC == 2 && ND > 1) {
#pragma unroll
            for (int v = 0; v < ND; v++) {
                const double *pl = a.vgrad + (size_t)(ND + v)*g.fstride + off;
                vg_l[v] = __ldg(pl); vg_r[v] = __ldg(pl + 1);
                prefetch_l2(pl + 32);                  // the next chunk's cells (same row, except at a wrap)
            }
        }
        double vL[E], vR[E];
#pragma unroll
        for (int v = 0; v < E; v++) {
            double s[5];
#pragma unroll
            for (int t = 0; t < 5; t++) s[t] = prim[v*PR + tp[t]];
            weno(s, vL[v], vR[v]);
        }
        double pc[E];                                  // primitive variables of the cell I finish (entry g-3)
#pragma unroll
        for (int v = 0; v < E; v++) pc[v] = (RK || v >= ADV || stab_on) ? prim[v*PR + tp[1]] : 0.0;
#if !MFC_STRICT
        if (stab_on) {                                 // ICFL, m_data_output.fpp:215-233 (inviscid)
            double rho = pc[0], gamma = pc[ADV]*a.gammas[0], pi_inf = pc[ADV]*a.pi_infs[0];
#pragma unroll
            for (int i = 1; i < NF; i++) {
                rho += pc[i];
                gamma = fma(pc[ADV + i], a.gammas[i], gamma);
                pi_inf = fma(pc[ADV + i], a.pi_infs[i], pi_inf);
            }
            const double cs = sqrt_ratio_fast(fma(gamma + 1.0, pc[NF + ND], pi_inf), gamma*rho);
            // dt/min_d(ds_d/(|u_d| + c)) = dt max_d((|u_d| + c)/ds_d)
            double m = (fabs(pc[NF]) + cs)*rds;
            if (ND >= 2) m = fmax(m, (fabs(pc[NF + (ND >= 2 ? 1 : 0)]) + cs)*__ldg(a.rds_t[0] + k0 + min(max(row3, 0), nr - 1) + g.b));
            if (ND >= 3) m = fmax(m, (fabs(pc[NF + (ND >= 3 ? 2 : 0)]) + cs)*__ldg(a.rds_t[1] + l + g.b));
            if (store_on) icfl = fmax(icfl, a.dt*m);
        }
#endif
        const int rpos = wpos ? wpos - 1 : XR - 1;     // the slot of the entry below mine
#pragma unroll
        for (int v = 0; v < E; v++) xv[v*XR + wpos] = vR[v];
        __syncwarp();
        double Lst[E];
#pragma unroll
        for (int v = 0; v < E; v++) Lst[v] = xv[v*XR + rpos];
        if (BC4) {
            if (a.bc_beg == -4 && j2 == 0) {           // m_riemann_solvers.fpp:480-487: qL(-1) = qR(0)
#pragma unroll
                for (int v = 0; v < E; v++) Lst[v] = vL[v];
            }
            if (a.bc_end == -4 && j2 == g.N[0] + 1) {  // :515-523: qR(m+1) = qL(m)
#pragma unroll
                for (int v = 0; v < E; v++) vL[v] = Lst[v];
            }
        }
        double F[E], uf;
        double vs[ND];
        hllc<NF, ND, 0>(Lst, vL, a.gammas, a.pi_infs, F, uf, vs);
        // faces ja-1/2 .. jb+1/2, keyed by the left cell (entry g-3; both entries lie in one row there)
        if (VISC == 1 && row3_ok && il3 >= 2 && il3 <= Lt - 4)
            store_visc_face<NF, ND>(a, off, Lst, vL, vs);
        if (VISC == 2) {                               // viscous source flux of this face, folded into F
            double dn[ND], ct[ND];
            const double rd = __ldg(a.rdcc + max(jf, -1) + g.b);   // 1/(x_cc(j3+1) - x_cc(j3))
#pragma unroll
            for (int v = 0; v < ND; v++) {
                dn[v] = (prim[(NF + v)*PR + tp[2]] - prim[(NF + v)*PR + tp[1]])*rd;
                ct[v] = ND > 1 ? 0.5*(vg_l[v] + vg_r[v]) : 0.0;
            }
            visc_flux_fd<NF, ND, 0>(a, Lst, vL, vs, dn, ct, F);
        }
        // ---- entry g-3: finish --------------------------------------------------------------------
#pragma unroll
        for (int v = 0; v < E; v++) xf[v*XR + wpos] = F[v];
        xf[E*XR + wpos] = uf;
        __syncwarp();
        double Fm[E], ufm;
#pragma unroll
        for (int v = 0; v < E; v++) Fm[v] = xf[v*XR + rpos];
        ufm = xf[E*XR + rpos];
        finish_cell2<NF, ND, 1, ACC, RK>(a, off, store_on, rds, pc, in, Fm, ufm, F, uf);
        // my next entry: 32 further down the stream
        il += 32;
        if (il >= Ls) { il -= Ls; row++; }
        ppos += 32;
        if (ppos >= PR) ppos -= PR;
        wpos = rpos;
    }
    if (stab_on) {                                     // all values >= 0: the bit pattern orders like the value
        unsigned long long b = (unsigned long long)__double_as_longlong(icfl);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) b = max(b, __shfl_xor_sync(0xffffffffu, b, o));
        if (lane == 0) atomicMax(a.stab_out, b);
    }
}

// ------------------------------------------------------------------------------------------
// y / z sweep, v3.  Every WARP is an independent pipeline: it owns 32 consecutive x columns at
// one transverse index and marches a pencil segment s0..s1 along the sweep direction.  Rows
// s0-3 .. s1+3 of the stage state stream through the warp's private ring (5 live rows of the
// stencil + the rows in flight, one tensor-map bulk copy per row); every lane converts, reads
// and reconstructs only its own column, carrying the previous cell's right-face state and the
// previous face's flux in registers.  EVERY other HBM operand moves through the TMA engine too:
//   rin   the RHS accumulated by the earlier directions (row of the cell being finished)
//   qin   q_cons_ts(1) of that row (last direction, stages 2/3)
//   outs  the finished row (RHS, or the updated state when the RK statement is fused in),
//         written back with one bulk tensor store; the interior-clipped tensor map drops the
//         columns beyond the domain, so there are no masked per-lane stores
// Each operand row is requested one march iteration (~3 us) before it is read, so the warp
// never waits on a global load (with LDG operands 8-11 % of all issue-stall samples sat on their
// first use, profiles/r01_v3a_stalls.md), and the 16 LDG/STG + 48 address instructions
// per cell become 24 LDS/STS.  The only synchronisation is the mbarrier wait and a __syncwarp
// before a slot is handed back.  Lanes beyond the domain compute on zero-filled columns.
// ------------------------------------------------------------------------------------------
template <int NF, int ND, int DIR, int COEF, bool BC4, int VISC, int WO>
__global__ void __launch_bounds__(32*kWarpsY, (COEF == 1 && WO == 5 && !MFC_STRICT) ? 2 : march_ctas(DIR, ND)) k_march3(const __grid_constant__ SweepArgs a) {
    constexpr int E = 2*NF + ND + 1, ADV = NF + ND + 1, R = march_ring(DIR, ND), SLOT = slot_doubles(E, kWY);
    constexpr bool ACC = true, RK = DIR == ND - 1;
    constexpr int NS = march_slots(DIR, ND);           // slots per warp: ring | rin | qin (RK) | outs
    constexpr int NB = R + 1;                          // mbarriers per warp: ring slots, operand rows (rin + qin)
    constexpr unsigned kRowBytes = (unsigned)(E*kWY*sizeof(double));
    static_assert(R >= 6, "the ring holds the 5 live rows of the stencil plus at least one row in flight");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const GridDesc &g = a.g;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *ring = reinterpret_cast<double *>(smem_raw) + warp*(NS*SLOT);
    double *rin = ring + R*SLOT, *qin = rin + SLOT, *outs = ring + (NS - 1)*SLOT;
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(reinterpret_cast<double *>(smem_raw) + kWarpsY*NS*SLOT) + warp*NB;
    unsigned long long *bar_r = bar + R;
    const int j0 = (blockIdx.x*kWarpsY + warp)*kWY;
    if (j0 > g.N[0]) return;                           // whole warp out of range (no block barriers below)
    const bool on = j0 + lane <= g.N[0];
    const int t = blockIdx.z;
    const int s0 = blockIdx.y*a.seg;
    const int s1 = min(s0 + a.seg - 1, g.N[DIR]);
    const long long ss = DIR == 1 ? g.sy : g.sz;
    const long long base = DIR == 1 ? g.at(j0, 0, t) : g.at(j0, t, 0);   // row 0 of the warp's columns
    // AHEAD (rings of >= 7 rows): the conversion runs one row ahead of the reconstruction -- row
    // s+3 is converted in iteration s, next to the register-only Riemann solve, so that its
    // dependent chain (LDS -> reciprocals -> STS) does not stand in front of the reconstruction.
    // One extra row (s1+4: at most the outermost ghost row, or zero-filled beyond it) is streamed
    // so that the conversion needs no guard.
    constexpr bool AHEAD = R >= 7;
    const int r_first = s0 - 3, r_last = s1 + 3 + (AHEAD ? 1 : 0);       // rows s0-3 .. s1+3 (+1)
    const bool need_q1 = RK && (MFC_STRICT ? a.rk_mode != 0 : a.rk_mode >= 2);
    if (lane == 0) {
        for (int i = 0; i < NB; i++) mbar_init(&bar[i], 1);
        mbar_fence_init();
    }
    __syncwarp();
    // tensor coordinates of row r of the warp's columns: (x, y, z) in the padded box for the
    // stage state; interior coordinates (cell indices) for the operand / output rows
    const int cx = j0 + kXoff, cy = (DIR == 1 ? 0 : t) + g.yoff, cz = (DIR == 1 ? t : 0) + g.zoff;
    // Ghost rows of a PHYSICAL boundary of this direction are not read from memory: the row the
    // boundary condition would have copied there (m_rhs.fpp:807-905: edge row for bc <= -3, mirror row
    // for -2, wrapped row for -1) is streamed in its place, and for -2 the normal velocity changes sign
    // after the conversion -- so k_bc is not launched for this direction at all (map_beg / map_end = the
    // effective boundary code, 0 = the ghosts are real: processor boundary, or mapping switched off).
    // source row of ghost row r: r < 0 -> lo_a + lo_b*r, r > N -> hi_a + hi_b*r (identity when not mapped)
    const int Nd = g.N[DIR];
    const int lo_a = a.map_beg == 0 ? 0 : (a.map_beg <= -3 ? 0 : (a.map_beg == -2 ? -1 : Nd + 1));
    const int lo_b = a.map_beg == 0 ? 1 : (a.map_beg <= -3 ? 0 : (a.map_beg == -2 ? -1 : 1));
    const int hi_a = a.map_end == 0 ? 0 : (a.map_end <= -3 ? Nd : (a.map_end == -2 ? 2*Nd + 1 : -Nd - 1));
    const int hi_b = a.map_end == 0 ? 1 : (a.map_end <= -3 ? 0 : (a.map_end == -2 ? -1 : 1));
    const bool mir_lo = a.map_beg == -2, mir_hi = a.map_end == -2;
    if (lane == 0) {
        for (int r = r_first; r < r_first + R && r <= r_last; r++) {
            const int rs = r < 0 ? lo_a + lo_b*r : (r > Nd ? hi_a + hi_b*r : r);
            mbar_expect_tx(&bar[r - r_first], kRowBytes);
            tma_load_row(ring + (r - r_first)*SLOT, &a.tm_q, cx, DIR == 1 ? cy + rs : cy, DIR == 1 ? cz : cz + rs, &bar[r - r_first]);
        }
        // operands of the first cell finished (s0): both rows complete the same barrier
        mbar_expect_tx(bar_r, need_q1 ? 2*kRowBytes : kRowBytes);
        tma_load_row(rin, &a.tm_rhs_i, j0, DIR == 1 ? s0 : t, DIR == 1 ? t : s0, bar_r);
        if (need_q1) tma_load_row(qin, &a.tm_q1_i, j0, DIR == 1 ? s0 : t, DIR == 1 ? t : s0, bar_r);
    }
    int next_issue = r_first + R;                      // only lane 0 issues, every lane counts

    // ring bookkeeping (R need not be a power of two): slot / phase of the next row to wait for
    // and convert, slot of row s-2 (the one handed back at the end of the iteration), and the
    // lane's pointers to rows s-2 .. s+2
    int slot_cv = 0, slot_lo = 0;
    unsigned phase_cv = 0, phase_op = 0;
#pragma unroll 1
    for (int i = 0; i < 4 + (AHEAD ? 1 : 0); i++) {    // rows s0-3 .. s0 (+ s0+1)
        mbar_wait(&bar[slot_cv], phase_cv);
        prim_in_place<NF, ND, kWY>(ring + slot_cv*SLOT + lane, a.gammas, a.pi_infs);
        if ((mir_lo && r_first + i < 0) || (mir_hi && r_first + i > Nd)) {   // mirrored ghost row: -u_normal
            double *un = ring + slot_cv*SLOT + lane + (NF + DIR)*kWY;
            *un = -*un;
        }
        if (++slot_cv == R) { slot_cv = 0; phase_cv ^= 1u; }
    }
    const double *p0 = ring + lane, *p1 = p0 + SLOT, *p2 = p1 + SLOT, *p3 = p2 + SLOT, *p4 = p3 + SLOT;
    const double *const ring_end = ring + R*SLOT;

    const unsigned uss = (unsigned)ss;
    // plane-relative element offset of cell s-1 of my column (lanes beyond the domain: column N)
    unsigned off = (unsigned)(base + (min(j0 + lane, g.N[0]) - j0) + (long long)(s0 - 3)*ss);
    Weno<COEF, WO> weno;
    // carried from cell s-1 / face s-3/2: right-face state, flux, face velocity.  Two sets that swap
    // roles every iteration (the loop below is unrolled by two), so that the hand-over from one cell to
    // the next is a renaming, not 2E+1 register moves.
    struct Carry { double vR[E], F[E], uf; };
    Carry cA, cB;
#pragma unroll
    for (int v = 0; v < E; v++) { cA.vR[v] = 0.0; cA.F[v] = 0.0; }
    cA.uf = 0.0;
    // One iteration of the march: reconstruct cell s, solve face s-1/2 against the carried
    // right-face state of cell s-1, finish cell s-1 with the carried flux of face s-3/2.
    auto iter = [&](const int s, const Carry &P, Carry &Nx) {
        mbar_wait(&bar[slot_cv], phase_cv);            // row s+2 (AHEAD: s+3) has arrived
        double *const row_cv = ring + slot_cv*SLOT + lane;
        if (++slot_cv == R) { slot_cv = 0; phase_cv ^= 1u; }
        if (!AHEAD) {
            prim_in_place<NF, ND, kWY>(row_cv, a.gammas, a.pi_infs);
            if (mir_hi && s + 2 > Nd) row_cv[(NF + DIR)*kWY] = -row_cv[(NF + DIR)*kWY];    // (rows < 0: prologue)
        }
        off += uss;
        weno.load(a, s);
        // VISC == 2: the x gradients of the velocities (planes [0][v] of k_vgrad) of cells s-1 and s of
        // my column, requested here so that the loads fly under the reconstruction
        double vg_l[ND], vg_r[ND];
        if (VISC == 2) {
#pragma unroll
            for (int v = 0; v < ND; v++) {
                const double *pl = a.vgrad + (size_t)v*g.fstride + off;
                vg_l[v] = __ldg(pl); vg_r[v] = __ldg(pl + uss);
                prefetch_l2(pl + 2*uss);               // row s+1, read by the next iteration: DRAM latency
                                                       // (~0.7 us) is longer than the ~0.5 us between load and use
            }
        }
        double vL[E];
#pragma unroll
        for (int v = 0; v < E; v++) {
            const double st[5] = {p0[v*kWY], p1[v*kWY], p2[v*kWY], p3[v*kWY], p4[v*kWY]};
            weno(st, vL[v], Nx.vR[v]);
        }
        const bool fin = s >= s0 + 1;                  // cell s-1 is finished in this iteration
        {   // face s-1/2.  The warm-up iteration s0-1 solves it too (against the zero state; the
            // result is overwritten before any use): one straight-line body, no second copy
            double Ls[BC4 ? E : 1];
            if (BC4) {
#pragma unroll
                for (int v = 0; v < E; v++) Ls[v] = P.vR[v];
                if (a.bc_beg == -4 && s == 0) {
#pragma unroll
                    for (int v = 0; v < E; v++) Ls[v] = vL[v];
                }
                if (a.bc_end == -4 && s == g.N[DIR] + 1) {
#pragma unroll
                    for (int v = 0; v < E; v++) vL[v] = Ls[v];
                }
            }
            const double *L = BC4 ? Ls : P.vR;
            double vs[ND];
            if (AHEAD) {
                prim_in_place<NF, ND, kWY>(row_cv, a.gammas, a.pi_infs);
                if (mir_hi && s + 3 > Nd) row_cv[(NF + DIR)*kWY] = -row_cv[(NF + DIR)*kWY];
            }
            hllc<NF, ND, DIR>(L, vL, a.gammas, a.pi_infs, Nx.F, Nx.uf, vs);
            if (VISC == 1 && on && s >= s0) store_visc_face<NF, ND>(a, off, L, vL, vs);   // face s-1/2, left cell s-1
            if (VISC == 2) {                           // viscous source flux of face s-1/2, folded into Fn
                double dn[ND], ct[ND];
                const double rd = __ldg(a.rdcc + s - 1 + g.b);       // 1/(s_cc(s) - s_cc(s-1))
#pragma unroll
                for (int v = 0; v < ND; v++) {
                    dn[v] = (p2[(NF + v)*kWY] - p1[(NF + v)*kWY])*rd;   // rows s, s-1
                    ct[v] = 0.5*(vg_l[v] + vg_r[v]);
                }
                visc_flux_fd<NF, ND, DIR>(a, L, vL, vs, dn, ct, Nx.F);
            }
            if (fin) {
                // operand rows (requested one iteration ago), read only now: 2E doubles that need not be
                // live across the Riemann solve
                CellIn<E, ACC, RK> in;
                mbar_wait(bar_r, phase_op);
#pragma unroll
                for (int v = 0; v < E; v++) in.r[v] = rin[v*kWY + lane];
                if (RK) {
                    if (need_q1) {
#pragma unroll
                        for (int v = 0; v < E; v++) in.q1[v] = qin[v*kWY + lane];
                    }
#if MFC_STRICT
                    if (a.rk_mode >= 2) {
#pragma unroll
                        for (int v = 0; v < ADV; v++) in.qs[v] = __ldg(a.q_v[v] + off);
                    }
#endif
                }
                phase_op ^= 1u;
                double y[E];
                finish_vals<NF, ND, kWY, ACC, RK>(a, a.rds[s - 1 + g.b], p1, in, P.F, P.uf, Nx.F, Nx.uf, y);   // p1: row s-1
                if (lane == 0) tma_store_wait_read();  // the previous row has left the output slot
                __syncwarp();
#pragma unroll
                for (int v = 0; v < E; v++) outs[v*kWY + lane] = y[v];
            }
        }
        // One proxy fence per iteration orders, for the whole warp, (a) the output row written
        // above before the bulk store reads it, (b) the reads of the operand slots and of ring
        // row s-2 before the TMA engine overwrites them.
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            if (fin) {
                tma_store_row(&a.tm_out_i, j0, DIR == 1 ? s - 1 : t, DIR == 1 ? t : s - 1, outs);
                if (s <= s1) {                         // operands of cell s, finished by the next iteration
                    mbar_expect_tx(bar_r, need_q1 ? 2*kRowBytes : kRowBytes);
                    tma_load_row(rin, &a.tm_rhs_i, j0, DIR == 1 ? s : t, DIR == 1 ? t : s, bar_r);
                    if (need_q1) tma_load_row(qin, &a.tm_q1_i, j0, DIR == 1 ? s : t, DIR == 1 ? t : s, bar_r);
                }
            }
            if (next_issue <= r_last) {
                const int rs = next_issue > Nd ? hi_a + hi_b*next_issue : next_issue;   // (rows < 0: prologue)
                mbar_expect_tx(&bar[slot_lo], kRowBytes);
                tma_load_row(ring + slot_lo*SLOT, &a.tm_q, cx, DIR == 1 ? cy + rs : cy, DIR == 1 ? cz : cz + rs, &bar[slot_lo]);
            }
        }
        next_issue++;
        if (++slot_lo == R) slot_lo = 0;
        p0 = p1; p1 = p2; p2 = p3; p3 = p4;
        p4 += SLOT;
        if (p4 >= ring_end) p4 -= R*SLOT;
    };
#pragma unroll 1
    for (int s = s0 - 1; s <= s1 + 1; s += 2) {
        iter(s, cA, cB);
        if (s + 1 <= s1 + 1) iter(s + 1, cB, cA);
    }
    if (lane == 0) tma_store_wait_read();              // the output slot must outlive the last store's read
}

// ------------------------------------------------------------------------------------------
// conservative -> primitive over the ghosted box, m_variables_conversion.fpp:326-373.
// Only velocity and pressure are stored (alpha_rho / alpha alias the conservative state).
// ------------------------------------------------------------------------------------------
template <int NF, int ND>
__global__ void __launch_bounds__(256) k_prim(const __grid_constant__ PrimArgs a) {
    constexpr int MOM = NF, EN = NF + ND, ADV = NF + ND + 1;
    const GridDesc &g = a.g;
    const int nx = g.N[0] + 1 + 2*g.b;
    const int jj = blockIdx.y*blockDim.x + threadIdx.x;
    if (jj >= nx) return;
    const int row = blockIdx.x;                    // (k, l) incl. ghosts: k + ey*l (gridDim.x has no 65535 cap)
    const long long cell = (long long)(kXoff - g.b + jj) + (long long)g.pitch*row;
    const long long fs = g.fstride;
    double rho = 0.0, gamma = 0.0, pi_inf = 0.0;   // :187-227
#pragma unroll
    for (int i = 0; i < NF; i++) {
        const double al = a.q[(ADV + i)*fs + cell];
        rho = rho + a.q[i*fs + cell];
        gamma = gamma + al*a.gammas[i];
        pi_inf = pi_inf + al*a.pi_infs[i];
    }
    rho = fmax(rho, 1e-16);                        // :353 (sgm_eps)
    double dyn = 0.0;
#pragma unroll
    for (int i = 0; i < ND; i++) {                 // :357-362
        const double mom = a.q[(MOM + i)*fs + cell];
        const double u = mom/rho;
        a.prim[i*fs + cell] = u;
        dyn = dyn + 5e-1*mom*u;
    }
    a.prim[ND*fs + cell] = (a.q[EN*fs + cell] - dyn - pi_inf)/gamma;   // :98-106
}

// ------------------------------------------------------------------------------------------
// physical boundary conditions on the conservative state, one direction per launch in the
// reference's order x, y, z so that later directions see earlier ghosts (corners),
// m_rhs.fpp:692-797 (x), :807-905 (y).  bc <= -3: extrapolation, -2: symmetry (normal
// momentum negated), -1: periodic.  Sides owned by a neighbour rank (bc >= 0) are left to
// the halo exchange.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void bc_decode(const GridDesc &g, int dir, long long idx, int &t0, int &t1, int &layer) {
    // transverse extents: earlier directions ghosted, later ones interior only
    int n0, n1, o0, o1;
    if (dir == 0) { n0 = g.N[1] + 1; o0 = 0; n1 = g.N[2] + 1; o1 = 0; }
    else if (dir == 1) { n0 = g.N[0] + 1 + 2*g.b; o0 = -g.b; n1 = g.N[2] + 1; o1 = 0; }
    else { n0 = g.N[0] + 1 + 2*g.b; o0 = -g.b; n1 = g.N[1] + 1 + 2*g.b; o1 = -g.b; }
    if (dir == 0) {
        // x slabs: the layer runs fastest, so that the b ghost cells of one row (contiguous in
        // memory, one 32-byte sector for b = 4) are written by neighbouring threads
        layer = (int)(idx % g.b); idx /= g.b;
        t0 = (int)(idx % n0) + o0; idx /= n0;
        t1 = (int)idx + o1;
        return;
    }
    t0 = (int)(idx % n0) + o0; idx /= n0;
    t1 = (int)(idx % n1) + o1; idx /= n1;
    layer = (int)idx;                              // 0 .. b-1
}
__device__ __forceinline__ long long bc_cell(const GridDesc &g, int dir, int s, int t0, int t1) {
    return dir == 0 ? g.at(s, t0, t1) : (dir == 1 ? g.at(t0, s, t1) : g.at(t0, t1, s));
}

#ifndef MFC_NF3_UNIT   // non-template kernels: defined once per build, in the main compile unit
__global__ void __launch_bounds__(256) k_bc(const __grid_constant__ BcArgs a) {
    const GridDesc &g = a.g;
    const long long n = slab_count(g, a.dir);
    const long long idx = (long long)blockIdx.x*blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const int v = blockIdx.y, side = blockIdx.z;
    const int code = side == 0 ? a.bc_beg : a.bc_end;
    if (code >= 0) return;
    int t0, t1, layer;
    bc_decode(g, a.dir, idx, t0, t1, layer);
    const int jj = layer + 1, N = g.N[a.dir];      // ghost index -jj or N+jj
    int src;
    if (code <= -3) src = side == 0 ? 0 : N;                       // :698-699,:750-751
    else if (code == -2) src = side == 0 ? jj - 1 : N - (jj - 1);  // :711-720,:764-774
    else src = side == 0 ? N - (jj - 1) : jj - 1;                  // :731-732,:786-787
    const int dst = side == 0 ? -jj : N + jj;
    double *f = a.q + (long long)v*g.fstride;
    double val = f[bc_cell(g, a.dir, src, t0, t1)];
    if (code == -2 && v == a.mom_normal) val = -val;               // :715-716,:829-830
    f[bc_cell(g, a.dir, dst, t0, t1)] = val;
}

// halo pack / unpack (m_mpi_proxy.fpp:490-499,539-548,592-601 and the y blocks).  The wire
// layout is private to this library (t0 fastest, then t1, layer, variable), chosen so both
// sides stream along x.
__global__ void __launch_bounds__(256) k_halo_pack(const __grid_constant__ HaloArgs a) {
    const GridDesc &g = a.g;
    const long long loc = (long long)blockIdx.x*blockDim.x + threadIdx.x;
    if (loc >= a.cnt) return;
    const long long idx = a.idx0 + loc;
    const int v = blockIdx.y;
    int t0, t1, layer;
    bc_decode(g, a.dir, idx, t0, t1, layer);
    const int N = g.N[a.dir];
    const int src = a.side == 0 ? layer : N - g.b + 1 + layer;     // first b / last b interior layers
    a.buf[(long long)v*a.cnt + loc] = a.q[(long long)v*g.fstride + bc_cell(g, a.dir, src, t0, t1)];
}
__global__ void __launch_bounds__(256) k_halo_unpack(const __grid_constant__ HaloArgs a) {
    const GridDesc &g = a.g;
    const long long loc = (long long)blockIdx.x*blockDim.x + threadIdx.x;
    if (loc >= a.cnt) return;
    const long long idx = a.idx0 + loc;
    const int v = blockIdx.y;
    int t0, t1, layer;
    bc_decode(g, a.dir, idx, t0, t1, layer);
    const int N = g.N[a.dir];
    const int dst = a.side == 0 ? -g.b + layer : N + 1 + layer;    // ghost layers in ascending order
    a.q[(long long)v*g.fstride + bc_cell(g, a.dir, dst, t0, t1)] = a.buf[(long long)v*a.cnt + loc];
}
#endif  // MFC_NF3_UNIT

// ------------------------------------------------------------------------------------------
// stability criteria, m_data_output.fpp:197-258: per-cell ICFL (+VCFL, Rc when viscous), warp
// shuffle reduction, one atomic per block on the ordered bit pattern (all values are >= 0).
// ------------------------------------------------------------------------------------------
template <int NF, int ND>
__global__ void __launch_bounds__(256) k_stability(const __grid_constant__ StabArgs a) {
    constexpr int ADV = NF + ND + 1;
    const GridDesc &g = a.g;
    const int j = blockIdx.y*blockDim.x + threadIdx.x;
    const int k = blockIdx.x % (g.N[1] + 1), l = blockIdx.x/(g.N[1] + 1);   // rows in gridDim.x: no 65535 cap
    double icfl = 0.0, vcfl = 0.0, Rc = 1.0e300;
    const bool visc = a.Re_size[0] > 0 || a.Re_size[1] > 0;
    if (j <= g.N[0]) {
        const long long cell = g.at(j, k, l), fs = g.fstride;
        double rho = 0.0, gamma = 0.0, pi_inf = 0.0, al[NF];
#pragma unroll
        for (int i = 0; i < NF; i++) {
            al[i] = a.q[(ADV + i)*fs + cell];
            rho = rho + a.q[i*fs + cell];
            gamma = gamma + al[i]*a.gammas[i];
            pi_inf = pi_inf + al[i]*a.pi_infs[i];
        }
        double vel[ND];
#pragma unroll
        for (int i = 0; i < ND; i++) vel[i] = a.prim[i*fs + cell];
        const double pres = a.prim[ND*fs + cell];
        const double c = sqrt(((gamma + 1.0)*pres + pi_inf)/(gamma*rho));        // :215-216
        const double dx = a.ds[0][j + g.b];
        double Re[2] = {0.0, 0.0};
        if (visc) {
#pragma unroll
            for (int i = 0; i < 2; i++) {
                double r = a.Re_size[i] > 0 ? 0.0 : -1e6;
                for (int q = 0; q < a.Re_size[i]; q++) r = al[a.Re_idx[i][q]]/a.Res[i][q] + r;
                Re[i] = 1.0/fmax(r, 1e-16);
            }
        }
        if (ND == 1) {                                                           // :231-241
            icfl = (a.dt/dx)*(fabs(vel[0]) + c);
            if (visc) { vcfl = fmax(a.dt/Re[0], a.dt/Re[1])/(dx*dx); Rc = dx*(fabs(vel[0]) + c)/fmax(1.0/Re[0], 1.0/Re[1]); }
        } else {
            const double dy = a.ds[1][k + g.b];
            double mn = fmin(dx/(fabs(vel[0]) + c), dy/(fabs(vel[ND > 1 ? 1 : 0]) + c));   // :220-221
            if (ND == 3) mn = fmin(mn, a.ds[2][l + g.b]/(fabs(vel[ND > 2 ? 2 : 0]) + c));
            icfl = a.dt/mn;
            if (visc) {                                                          // :223-229
                const double md = fmin(dx, dy);
                vcfl = fmax(a.dt/Re[0], a.dt/Re[1])/(md*md);
                Rc = fmin(dx*(fabs(vel[0]) + c), dy*(fabs(vel[ND > 1 ? 1 : 0]) + c))/fmax(1.0/Re[0], 1.0/Re[1]);
            }
        }
    }
    unsigned long long bi = (unsigned long long)__double_as_longlong(icfl);
    unsigned long long bv = (unsigned long long)__double_as_longlong(vcfl);
    unsigned long long br = (unsigned long long)__double_as_longlong(Rc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        bi = max(bi, __shfl_xor_sync(0xffffffffu, bi, o));
        bv = max(bv, __shfl_xor_sync(0xffffffffu, bv, o));
        br = min(br, __shfl_xor_sync(0xffffffffu, br, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(a.out + 0, bi);
        if (visc) { atomicMax(a.out + 1, bv); atomicMin(a.out + 2, br); }
    }
}

// ==========================================================================================
// Viscous path (any fluid_pp(i)%Re > 0; 1-D / 2-D like the reference).  Correctness-first
// kernels beside the fused sweeps: the sweeps store vel_src and Re_avg of every face they solve
// (store_visc_face); k_visc turns them and the velocity gradients into the viscous source flux
// of the cell's two faces and adds its divergence to the RHS, in the reference's order
// (direction by direction, after the inviscid terms of that direction).
// ==========================================================================================
__device__ __forceinline__ void weno_at(const double *f, long long cell, long long stride, const double *coef, int clen,
                                        int coef_lo, int idx, double eps, double &vL, double &vR) {
    double c[27], st[5];
    const double *p = coef + (idx - coef_lo);
#pragma unroll
    for (int i = 0; i < 27; i++) c[i] = p[(long long)i*clen];
#pragma unroll
    for (int q = 0; q < 5; q++) st[q] = f[cell + (long long)(q - 2)*stride];
    weno5(st, c, eps, vL, vR);
}

// Fast build, weno_Re_flux = F: per cell and direction dd the mean of the two one-sided differences
// of every velocity, C[dd][v] = ((u_c - u_{c-1})/(cc_c - cc_{c-1}) + (u_{c+1} - u_c)/(cc_{c+1} - cc_c))/2
// (m_viscous.fpp:278-347: the cross-direction gradient of a face is the mean of its two cells' C).
// Cells -1 .. N+1 of both directions (the faces of direction d reach one cell into its ghosts).
// The velocities come straight from the conservative state (momentum / max(rho, sgm_eps),
// m_variables_conversion.fpp:353-359): no primitive planes are written on this path.
template <int NF, int ND>
__global__ void __launch_bounds__(128) k_vgrad(const __grid_constant__ ViscArgs a) {
    const GridDesc &g = a.g;
    const int j = (int)(blockIdx.y*blockDim.x + threadIdx.x) - 1;
    const int k = ND > 1 ? (int)blockIdx.x + a.k_lo : 0;
    if (j > g.N[0] + 1) return;
    const int c[3] = {j, k, 0};
    const long long cell = g.at(j, k, 0), fs = g.fstride;
    auto vel = [&](long long at, double (&u)[ND]) {
        double rho = a.qs[at];
#pragma unroll
        for (int i = 1; i < NF; i++) rho += a.qs[i*fs + at];
        const double ir = rcp_fast3(fmax(rho, 1e-16));
#pragma unroll
        for (int v = 0; v < ND; v++) u[v] = a.qs[(NF + v)*fs + at]*ir;
    };
    double u0[ND];
    vel(cell, u0);
#pragma unroll
    for (int dd = 0; dd < ND; dd++) {
        const long long sd = g.stride(dd);
        const double rm = a.rdcc[dd][c[dd] - 1 + g.b], rp = a.rdcc[dd][c[dd] + g.b];
        double um[ND], up[ND];
        vel(cell - sd, um);
        vel(cell + sd, up);
#pragma unroll
        for (int v = 0; v < ND; v++)
            a.grad[(dd*ND + v)*fs + cell] = 0.5*((u0[v] - um[v])*rm + (up[v] - u0[v])*rp);
    }
}

// weno_Re_flux branch, m_viscous.fpp:186-217: velocities are WENO-reconstructed along every
// direction i and the divergence theorem gives the cell gradient
//   dq_prim_d<i>_qp(v) = 1/ds_i (vR - vL)                              (:417-422, :444-449)
// over cells -4 .. N+4 of every active direction (what the later reconstruction reads).
template <int ND>
__global__ void __launch_bounds__(128) k_visc_grad(const __grid_constant__ ViscArgs a) {
    const GridDesc &g = a.g;
    const int j = (int)(blockIdx.y*blockDim.x + threadIdx.x) - 4;
    const int k = ND > 1 ? (int)blockIdx.x - 4 : 0;
    if (j > g.N[0] + 4) return;
    const int c[3] = {j, k, 0};
    const long long cell = g.at(j, k, 0), fs = g.fstride;
#pragma unroll
    for (int i = 0; i < ND; i++)
#pragma unroll
        for (int v = 0; v < ND; v++) {
            double vL, vR;
            weno_at(a.prim + v*fs, cell, g.stride(i), a.coef[i], a.clen[i], a.coef_lo[i], c[i], a.eps, vL, vR);
            a.grad[(i*ND + v)*fs + cell] = 1.0/a.ds[i][c[i] + g.b]*(vR - vL);
        }
}

// x / Re_avg: the strict build divides (m_riemann_solvers.fpp:722-899 as written); the fast build
// multiplies by the 1/Re_avg the sweep stored
#if MFC_STRICT
#define VDIV(x, re) ((x)/(re))
#define VDIV3(x, re) ((x)/(3.0*(re)))
#else
#define VDIV(x, re) ((x)*(re))
#define VDIV3(x, re) ((x)*((1.0/3.0)*(re)))
#endif
template <int NF, int ND>
__global__ void __launch_bounds__(128, 4) k_visc(const __grid_constant__ ViscArgs a) {
    const GridDesc &g = a.g;
    const int j = blockIdx.y*blockDim.x + threadIdx.x, k = blockIdx.x;
    if (j > g.N[0]) return;
    const int id = a.dir, b = g.b;
    const int c[3] = {j, k, 0};
    const long long cell = g.at(j, k, 0), fs = g.fstride, sid = g.stride(id);
    constexpr int E = 2*NF + ND + 1, MOM = NF, EN = NF + ND;
    double fm[2][ND], fe[2];                       // flux_src(mom), flux_src(E) of faces c-1/2 and c+1/2
#pragma unroll
    for (int side = 0; side < 2; side++) {
        const int f = c[id] - 1 + side;            // the face's left cell along id
        const long long cl = cell + (long long)(side - 1)*sid, cr = cl + sid;
        double vsrc[ND], Re_avg[2];
#pragma unroll
        for (int i = 0; i < ND; i++) vsrc[i] = a.visc_face[i*fs + cl];
        Re_avg[0] = a.visc_face[ND*fs + cl];
        Re_avg[1] = a.visc_face[(ND + 1)*fs + cl];
        // avg[dd][v] = 5d-1*(dvelL_d<dd>(v)(j,k) + dvelR_d<dd>(v)(j+1,k)), m_riemann_solvers.fpp:716-899
        double avg[ND][ND];
#pragma unroll
        for (int dd = 0; dd < ND; dd++)
#pragma unroll
            for (int v = 0; v < ND; v++) {
                double dR, dL;                     // dqR(left cell), dqL(right cell)
                if (a.weno_Re_flux) {              // m_rhs.fpp:513-529: gradients reconstructed along id
                    const double *G = a.grad + (dd*ND + v)*fs;
                    double vLl, vRl, vLr, vRr;
                    weno_at(G, cl, sid, a.coef[id], a.clen[id], a.coef_lo[id], f, a.eps, vLl, vRl);
                    weno_at(G, cr, sid, a.coef[id], a.clen[id], a.coef_lo[id], f + 1, a.eps, vLr, vRr);
                    dR = vRl; dL = vLr;
                    if (a.bc_beg == -4 && f == -1) dR = dL;          // m_riemann_solvers.fpp:489-511
                    if (a.bc_end == -4 && f == g.N[id]) dL = dR;     // :525-548
                } else {                           // m_viscous.fpp:219-347: finite differences
                    const double *u = a.prim + v*fs;
                    if (dd == id) {
                        const double *cc = a.cc[id] + b;
#if MFC_STRICT
                        dR = (u[cr] - u[cl])/(cc[f + 1] - cc[f]);    // dqR(c) :237-248 == dqL(c+1) :224-235
#else
                        dR = (u[cr] - u[cl])*a.rdcc[id][f + b];
#endif
                        dL = dR;
                    } else {
                        const long long sd = g.stride(dd);
                        const double *cc = a.cc[dd] + b;
                        const int t = c[dd];
#if MFC_STRICT
                        const double sLr = (u[cr] - u[cr - sd])/(cc[t] - cc[t - 1]), sRr = (u[cr + sd] - u[cr])/(cc[t + 1] - cc[t]);
                        const double sLl = (u[cl] - u[cl - sd])/(cc[t] - cc[t - 1]), sRl = (u[cl + sd] - u[cl])/(cc[t + 1] - cc[t]);
#else
                        const double rm = a.rdcc[dd][t - 1 + b], rp = a.rdcc[dd][t + b];
                        const double sLr = (u[cr] - u[cr - sd])*rm, sRr = (u[cr + sd] - u[cr])*rp;
                        const double sLl = (u[cl] - u[cl - sd])*rm, sRl = (u[cl + sd] - u[cl])*rp;
#endif
                        dR = 25e-2*(sLr + sRr + sLl + sRl);          // :300-307 at the left cell == :283-290 at the right cell
                        dL = dR;
                    }
                }
                avg[dd][v] = 5e-1*(dR + dL);
            }
        double m[ND], e = 0.0;                     // s_initialize_riemann_solver zeroes flux_src, :627-670
#pragma unroll
        for (int i = 0; i < ND; i++) m[i] = 0.0;
        if (id == 0) {
            if (a.Re_size[0] > 0) {                // :714-736
                const double tau = VDIV((4.0/3.0)*avg[0][0], Re_avg[0]);
                m[0] = m[0] - tau; e = e - vsrc[0]*tau;
            }
            if (a.Re_size[1] > 0) {                // :738-760
                const double tau = VDIV(avg[0][0], Re_avg[1]);
                m[0] = m[0] - tau; e = e - vsrc[0]*tau;
            }
            if (ND > 1) {
                if (a.Re_size[0] > 0) {            // :764-801
                    double tau[2];
                    tau[0] = VDIV(-(2.0/3.0)*avg[ND > 1 ? 1 : 0][ND > 1 ? 1 : 0], Re_avg[0]);
                    tau[1] = VDIV(avg[ND > 1 ? 1 : 0][0] + avg[0][ND > 1 ? 1 : 0], Re_avg[0]);
#pragma unroll
                    for (int i = 0; i < 2; i++) {
                        m[ND > 1 ? i : 0] = m[ND > 1 ? i : 0] - tau[i];
                        e = e - vsrc[ND > 1 ? i : 0]*tau[i];
                    }
                }
                if (a.Re_size[1] > 0) {            // :803-825
                    const double tau = VDIV(avg[ND > 1 ? 1 : 0][ND > 1 ? 1 : 0], Re_avg[1]);
                    m[0] = m[0] - tau; e = e - vsrc[0]*tau;
                }
            }
        } else if (ND > 1) {
            constexpr int Y = ND > 1 ? 1 : 0;
            if (a.Re_size[0] > 0) {                // :831-872
                double tau[2];
                tau[0] = VDIV(avg[Y][0] + avg[0][Y], Re_avg[0]);
                tau[1] = VDIV3(4.0*avg[Y][Y] - 2.0*avg[0][0], Re_avg[0]);
#pragma unroll
                for (int i = 0; i < 2; i++) {
                    m[ND > 1 ? i : 0] = m[ND > 1 ? i : 0] - tau[i];
                    e = e - vsrc[ND > 1 ? i : 0]*tau[i];
                }
            }
            if (a.Re_size[1] > 0) {                // :874-899
                const double tau = VDIV(avg[0][0] + avg[Y][Y], Re_avg[1]);
                m[Y] = m[Y] - tau; e = e - vsrc[Y]*tau;
            }
        }
#pragma unroll
        for (int i = 0; i < ND; i++) fm[side][i] = m[i];
        fe[side] = e;
    }
    // m_rhs.fpp:591-604, :639-652
#if MFC_STRICT
    const double dsj = a.ds[id][c[id] + b];
#define MFC_VADD(r, d) ((r) + 1.0/dsj*(d))
#else
    const double rdsj = a.rds[id][c[id] + b];
#define MFC_VADD(r, d) fma(rdsj, (d), (r))
#endif
    if (a.rk_mode == 0) {
#pragma unroll
        for (int i = 0; i < ND; i++) {
            double *r = a.rhs + (MOM + i)*fs + cell;
            *r = MFC_VADD(*r, fm[0][i] - fm[1][i]);
        }
        double *r = a.rhs + EN*fs + cell;
        *r = MFC_VADD(*r, fe[0] - fe[1]);
        return;
    }
    // last direction: the RHS is complete here -- apply the TVD-RK statement
    // (m_time_steppers.fpp:167,245,322,342) instead of storing it for a separate pass
    double r[E], q1[E], qs[E];
#pragma unroll
    for (int v = 0; v < E; v++) {                  // all loads first: 3 E independent requests in flight
        const long long o = v*fs + cell;
        r[v] = a.rhs[o]; q1[v] = a.q1[o]; qs[v] = a.qs[o];
    }
#pragma unroll
    for (int i = 0; i < ND; i++) r[MOM + i] = MFC_VADD(r[MOM + i], fm[0][i] - fm[1][i]);
    r[EN] = MFC_VADD(r[EN], fe[0] - fe[1]);
#pragma unroll
    for (int v = 0; v < E; v++) a.qout[v*fs + cell] = rk_apply(a.rk_mode, q1[v], qs[v], r[v], a.dt);
#undef MFC_VADD
}
#undef VDIV
#undef VDIV3

}  // namespace MFC_NS
