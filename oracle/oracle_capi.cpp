// oracle_capi.cpp -- extern "C" surface of the CPU oracle for ctypes.  TEST INFRASTRUCTURE
// ONLY (see mfc_oracle.hpp): loaded by tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs; never by the product path.
#include "mfc_oracle.hpp"
#include <cstring>
#include <chrono>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace orc;

extern "C" {

// gp: the GLOBAL case (m,n,p = m_glb,n_glb,p_glb; bc = physical codes).  cb[d] -> global
// cell boundaries s_cb(-1:N_glb) (N_glb+2 doubles).  num_procs emulated MPI ranks.
void *orc_create(const mfc_b200_params_t *gp, const double *xcb, const double *ycb, const double *zcb,
                 int num_procs, char *err, int errlen) {
    std::string e;
    const double *cb[3] = {xcb, ycb, zcb};
    World *w = world_create(gp, cb, num_procs, e);
    if (!w && err && errlen > 0) { std::strncpy(err, e.c_str(), (size_t)errlen - 1); err[errlen - 1] = 0; }
    return w;
}
void orc_destroy(void *h) { delete (World *)h; }

void orc_set_q(void *h, const double *const *q) { world_set_q(*(World *)h, q); }
void orc_get_q(void *h, double *const *q) { world_get_q(*(World *)h, q); }
void orc_get_prim(void *h, double *const *q) { world_get_prim(*(World *)h, q); }

// s_compute_rhs on q_cons_ts(1)
void orc_compute_rhs(void *h, int t_step, double *const *rhs) {
    World &w = *(World *)h;
    world_compute_rhs(w, 0, t_step);
    if (rhs) world_get_rhs(w, rhs);
}

void orc_step(void *h, int t_step, double dt, double *stab) { world_step(*(World *)h, t_step, dt, stab); }

// n consecutive steps with constant dt, returns wall seconds (cpu_baseline timing leg)
double orc_run_steps(void *h, int t_step0, int n_steps, double dt) {
    World &w = *(World *)h;
    auto t0 = std::chrono::steady_clock::now();
    for (int s = 0; s < n_steps; s++) world_step(w, t_step0 + s, dt, nullptr);
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

int orc_decompose(int num_procs, int nd, const int *Nglb, int weno_order, int *np_out) {
    return decompose(num_procs, nd, Nglb, weno_order, np_out) ? 0 : -1;
}

// out: N[3], start_idx[3], bc[6], coords[3], buff_size, sys_size  (17 ints)
void orc_rank_info(void *h, int rank, int *out) {
    const Rank &r = ((World *)h)->ranks[(size_t)rank];
    for (int d = 0; d < 3; d++) { out[d] = r.N[d]; out[3 + d] = r.start_idx[d]; out[12 + d] = r.coords[d]; }
    for (int d = 0; d < 3; d++) { out[6 + 2*d] = r.bc[d][0]; out[7 + 2*d] = r.bc[d][1]; }
    out[15] = r.b; out[16] = r.E;
}

// ghosted metrics of one rank/direction: cb (N+2+2b), cc (N+1+2b), ds (N+1+2b)
void orc_rank_metrics(void *h, int rank, int dir, double *cb, double *cc, double *ds) {
    const Rank &r = ((World *)h)->ranks[(size_t)rank];
    std::memcpy(cb, r.cb[dir].a.data(), r.cb[dir].a.size()*sizeof(double));
    std::memcpy(cc, r.cc[dir].a.data(), r.cc[dir].a.size()*sizeof(double));
    std::memcpy(ds, r.ds[dir].a.data(), r.ds[dir].a.size()*sizeof(double));
}

// WENO coefficients of one rank/direction, shapes (ncell,3,2),(ncell,3,2),(ncell,3),(ncell,3),(ncell,3,3)
int orc_weno_coefficients(void *h, int rank, int dir, double *pL, double *pR, double *dL, double *dR, double *bt) {
    const WenoCoef &w = ((World *)h)->ranks[(size_t)rank].wc[dir];
    if (pL) std::memcpy(pL, w.pL.data(), w.pL.size()*sizeof(double));
    if (pR) std::memcpy(pR, w.pR.data(), w.pR.size()*sizeof(double));
    if (dL) std::memcpy(dL, w.dL.data(), w.dL.size()*sizeof(double));
    if (dR) std::memcpy(dR, w.dR.data(), w.dR.size()*sizeof(double));
    if (bt) std::memcpy(bt, w.bt.data(), w.bt.size()*sizeof(double));
    return w.hi - w.lo + 1;
}

// one rank's reconstructed face values of variable v in direction dir over the ghosted box
// (for kernel-level parity tests of the WENO stage); which: 0 = vL, 1 = vR, 2 = flux, 3 = face velocity
void orc_rank_scratch(void *h, int rank, int which, int dir, int v, double *out) {
    const Rank &r = ((World *)h)->ranks[(size_t)rank];
    const Field &f = which == 0 ? r.qL_rs[dir][v] : which == 1 ? r.qR_rs[dir][v] : which == 2 ? r.flux[v] : r.flux_src_adv;
    std::memcpy(out, f.a.data(), f.a.size()*sizeof(double));
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
int orc_is_strict(void) {
#if defined(__FAST_MATH__) || defined(ORC_TIMING_BUILD)
    return 0;
#else
    return 1;
#endif
}

}  // extern "C"
