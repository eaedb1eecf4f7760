// mfc_oracle.hpp -- CPU restatement of MicroFC's s_compute_rhs pipeline + TVD-RK steppers.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product path (microfc_b200/, libmfc_b200.so)
// may include, link, call or execute anything under oracle/.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it,
// as the checker or the reported CPU baseline -- never as the thing shipped.
//
// PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for this path
// and cannot be compiled in this image (no Fortran compiler, no fypp, no MPI), so this
// restatement is pinned only by construction (same loops, same operand order, same index
// ranges, file:line citations on every block) and by first-principles known answers
// (tests/test_oracle_*.py: exact Sod solution, pressure-equilibrium preservation,
// conservation, symmetry, WENO polynomial exactness, HLLC consistency).
//
// All citations are relative to /root/reference.
#pragma once
#include <cstdint>
#include <vector>
#include <string>
#include "../include/mfc_b200.h"

namespace orc {

struct Bounds { int beg = 0, end = 0; };   // int_bounds_info, m_derived_types.f90:36-43

// scalar_field%sf with arbitrary lower bounds, x fastest (m_derived_types.f90:25-27);
// the third index is the 3-D extension.
struct Field {
    int lo[3] = {0, 0, 0};
    int ext[3] = {1, 1, 1};
    std::vector<double> a;
    void alloc(Bounds bx, Bounds by, Bounds bz) {
        lo[0] = bx.beg; lo[1] = by.beg; lo[2] = bz.beg;
        ext[0] = bx.end - bx.beg + 1; ext[1] = by.end - by.beg + 1; ext[2] = bz.end - bz.beg + 1;
        a.assign((size_t)ext[0]*ext[1]*ext[2], 0.0);
    }
    inline size_t idx(int j, int k, int l) const {
        return (size_t)(j - lo[0]) + (size_t)ext[0]*((size_t)(k - lo[1]) + (size_t)ext[1]*(size_t)(l - lo[2]));
    }
    inline double &operator()(int j, int k, int l) { return a[idx(j, k, l)]; }
    inline double operator()(int j, int k, int l) const { return a[idx(j, k, l)]; }
    // access with the sweep index s along direction d and the two others in natural order
    inline double &at(const int c[3]) { return a[idx(c[0], c[1], c[2])]; }
    inline double at(const int c[3]) const { return a[idx(c[0], c[1], c[2])]; }
};

// 1-D array with a lower bound (x_cb(-1-b:m+b) and friends)
struct Arr1 {
    int lo = 0; std::vector<double> a;
    void alloc(int lo_, int hi_) { lo = lo_; a.assign((size_t)(hi_ - lo_ + 1), 0.0); }
    inline double &operator()(int i) { return a[(size_t)(i - lo)]; }
    inline double operator()(int i) const { return a[(size_t)(i - lo)]; }
};

// grid-dependent WENO coefficients of one direction (m_weno.fpp:41-80)
struct WenoCoef {
    int lo = 0, hi = -1;               // cell range  is%beg+polyn : is%end-polyn
    // poly_coef_cb{L,R}(cell, 0:polyn, 0:polyn-1), d_cb{L,R}(0:polyn, cell), beta_coef(cell,0:polyn,0:2(polyn-1))
    std::vector<double> pL, pR, dL, dR, bt;
    inline double &polyL(int j, int k, int q) { return pL[((size_t)(j - lo)*3 + k)*2 + q]; }
    inline double &polyR(int j, int k, int q) { return pR[((size_t)(j - lo)*3 + k)*2 + q]; }
    inline double &dcbL(int k, int j) { return dL[(size_t)(j - lo)*3 + k]; }
    inline double &dcbR(int k, int j) { return dR[(size_t)(j - lo)*3 + k]; }
    inline double &beta(int j, int k, int q) { return bt[((size_t)(j - lo)*3 + k)*3 + q]; }
};

// One MPI rank of the reference `simulation` executable: the module globals of
// m_global_parameters + the state owned by m_time_steppers / m_rhs / m_weno / m_riemann_solvers.
struct Rank {
    // ---- m_global_parameters.fpp ----
    int N[3] = {0, 0, 0};              // m, n, p (local)
    int nd = 1, nf = 1, E = 0, b = 0;  // num_dims, num_fluids, sys_size, buff_size
    int weno_order = 5, weno_polyn = 2;
    double weno_eps = 1e-16;
    int time_stepper = 3;
    bool weno_Re_flux = false, run_time_info = false;
    int t_step_stop = 0;
    int bc[3][2] = {{-3, -3}, {-3, -3}, {-3, -3}};
    int rank = 0, coords[3] = {0, 0, 0}, start_idx[3] = {0, 0, 0};
    // 0-based equation indices (m_global_parameters.fpp:302-310, :376-381)
    int contxb = 0, contxe = 0, momxb = 0, momxe = 0, E_idx = 0, advxb = 0, advxe = 0;
    double gammas[MFC_B200_MAX_FLUIDS], pi_infs[MFC_B200_MAX_FLUIDS];
    double fluid_Re[MFC_B200_MAX_FLUIDS][2];
    int Re_size[2] = {0, 0};
    int Re_idx[2][MFC_B200_MAX_FLUIDS];
    double Res[2][MFC_B200_MAX_FLUIDS];
    bool viscous = false;               // any(Re_size > 0)
    Arr1 cb[3], cc[3], ds[3];           // x_cb, x_cc, dx (+y,+z)
    Bounds g[3];                        // ghosted range per direction: (-b, N+b) or (0,0)

    // ---- m_weno ----
    WenoCoef wc[3];
    // ---- m_time_steppers ----
    std::vector<Field> q_ts[2];         // q_cons_ts(1:2)%vf
    std::vector<Field> q_prim_vf, rhs_vf;
    // ---- m_rhs ----
    std::vector<Field> q_cons_qp, q_prim_qp;   // q_prim_qp alpha_rho / alpha alias q_cons_qp in the reference
    std::vector<Field> qL_rs[3], qR_rs[3];     // qL_rsx_vf, qL_rsy_vf (+z), stored untransposed
    std::vector<Field> flux, flux_src;         // flux_n(1) (flux_n(2) aliases it), flux_src_n(1)
    Field flux_src_adv;                        // flux_src_n(:)%vf(advxb) -- one array aliased by all adv eqns
    std::vector<Field> vel_src;                // vel_src_rs{x,y}_vf
    Field Re_avg[2];
    // viscous scratch: dq{L,R}_prim_d{x,y,z}_n(dir)%vf(mom)  -> [dir][deriv][comp]
    std::vector<Field> dqL[3][3], dqR[3][3];
    std::vector<Field> dq_prim_d[3];           // dq_prim_dx_qp, dq_prim_dy_qp (+z)
    std::vector<Field> qL_prim[3], qR_prim[3]; // qL_prim(i)%vf(mom)
    std::vector<Field> dqL_rs[3], dqR_rs[3];

    // stability extrema of the last s_write_run_time_information
    double icfl_max_loc = 0, vcfl_max_loc = 0, Rc_min_loc = 0;
};

struct World {
    mfc_b200_params_t gp;               // the GLOBAL case (m_glb etc. in m,n,p; bc = physical codes)
    int num_procs = 1, np[3] = {1, 1, 1};
    std::vector<Rank> ranks;
    std::vector<double> cb_glb[3];      // global cell boundaries (-1 : N_glb)
    double icfl_max_glb = 0, vcfl_max_glb = 0, Rc_min_glb = 0;
    std::string err;
};

// m_mpi_proxy.fpp:134-328 (+ 3-D extension)
bool decompose(int num_procs, int nd, const int Nglb[3], int weno_order, int np_out[3]);

World *world_create(const mfc_b200_params_t *global, const double *const cb_glb[3], int num_procs, std::string &err);
void world_set_q(World &w, const double *const q[]);          // global interior arrays (0:m_glb,0:n_glb,0:p_glb)
void world_get_q(const World &w, double *const q[]);
void world_get_prim(const World &w, double *const q[]);
void world_compute_rhs(World &w, int stage /*0: q_ts(1), 1: q_ts(2)*/, int t_step);
void world_get_rhs(const World &w, double *const rhs[]);
void world_step(World &w, int t_step, double dt, double stab[3]);

}  // namespace orc
