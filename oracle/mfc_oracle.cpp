// mfc_oracle.cpp -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY,
// see the header of mfc_oracle.hpp (PARITY UNPINNED by reference fixtures: none exist).
//
// Build (oracle/Makefile): strict  g++ -O2 -ffp-contract=off -fno-fast-math   (the oracle)
//                          timing  g++ -O3 -march=native -fopenmp             (cpu_baseline only)
// Every block cites the reference statement(s) it restates (paths relative to /root/reference).
// Loops keep the reference's operand order; `x**2d0` is written x*x (exactly rounded either way).
#include "mfc_oracle.hpp"
#include <cmath>
#include <cstring>
#include <algorithm>
#include <cstdio>

namespace orc {

static const double sgm_eps = 1e-16;     // m_global_parameters.fpp:39
static const double dflt_real = -1e6;    // m_global_parameters.fpp:37
static const int num_stcls_min = 5;      // m_global_parameters.fpp:33

// ---------------------------------------------------------------------------------------
// m_mpi_proxy.fpp:134-328  s_mpi_decompose_computational_domain (processor topology part)
// ---------------------------------------------------------------------------------------
bool decompose(int num_procs, int nd, const int Nglb[3], int weno_order, int np[3]) {
    const int m = Nglb[0], n = Nglb[1], p = Nglb[2];
    const int lim = num_stcls_min*weno_order;
    np[0] = np[1] = np[2] = 1;
    if (nd == 1) {                       // :267-281
        np[0] = num_procs;
        return true;
    }
    if (nd == 2) {                       // :163-212
        int npx = 1, npy = num_procs, ierr = -1;
        double tx = npx, ty = npy;
        double fct_min = 10.0*std::fabs((m + 1)/tx - (n + 1)/ty);
        for (int i = 1; i <= num_procs; i++) {
            if (num_procs % i == 0 && (m + 1)/i >= lim) {
                tx = i; ty = num_procs/i;
                if (fct_min >= std::fabs((m + 1)/tx - (n + 1)/ty) && (n + 1)/ty >= lim) {
                    npx = i; npy = num_procs/i;
                    fct_min = std::fabs((m + 1)/tx - (n + 1)/ty);
                    ierr = 0;
                }
            }
        }
        np[0] = npx; np[1] = npy;
        return ierr == 0 || num_procs == 1;
    }
    // 3-D EXTENSION (no reference): the same rule over three factors.
    int npx = 1, npy = 1, npz = num_procs, ierr = -1;
    double tx = npx, ty = npy, tz = npz;
    double fct_min = 10.0*std::fabs((m + 1)/tx - (n + 1)/ty) + 10.0*std::fabs((n + 1)/ty - (p + 1)/tz);
    for (int i = 1; i <= num_procs; i++) {
        if (num_procs % i == 0 && (m + 1)/i >= lim) {
            for (int j = 1; j <= num_procs/i; j++) {
                if ((num_procs/i) % j == 0 && (n + 1)/j >= lim) {
                    tx = i; ty = j; tz = num_procs/(i*j);
                    double f = std::fabs((m + 1)/tx - (n + 1)/ty) + std::fabs((n + 1)/ty - (p + 1)/tz);
                    if (fct_min >= f && (p + 1)/tz >= lim) {
                        npx = i; npy = j; npz = num_procs/(i*j);
                        fct_min = f; ierr = 0;
                    }
                }
            }
        }
    }
    np[0] = npx; np[1] = npy; np[2] = npz;
    return ierr == 0 || num_procs == 1;
}

static int cart_rank(const int np[3], const int c_in[3], int nd) {
    // MPI_CART_RANK on a fully periodic, row-major Cartesian communicator (:215-222)
    int c[3];
    for (int d = 0; d < 3; d++) { c[d] = ((c_in[d] % np[d]) + np[d]) % np[d]; }
    if (nd == 1) return c[0];
    if (nd == 2) return c[0]*np[1] + c[1];
    return (c[0]*np[1] + c[1])*np[2] + c[2];
}

// ---------------------------------------------------------------------------------------
// m_global_parameters.fpp:285-396  s_initialize_global_parameters_module
// ---------------------------------------------------------------------------------------
static void init_global_parameters(Rank &r) {
    r.weno_polyn = (r.weno_order - 1)/2;                         // :291
    r.contxb = 0; r.contxe = r.nf - 1;                           // :302-310 (0-based here)
    r.momxb = r.nf; r.momxe = r.nf + r.nd - 1;
    r.E_idx = r.momxe + 1;
    r.advxb = r.E_idx + 1; r.advxe = r.E_idx + r.nf;
    r.E = r.advxe + 1;
    r.Re_size[0] = r.Re_size[1] = 0;                             // :300,:314-339
    for (int i = 0; i < r.nf; i++) {
        if (r.fluid_Re[i][0] > 0) { r.Re_idx[0][r.Re_size[0]++] = i; }
        if (r.fluid_Re[i][1] > 0) { r.Re_idx[1][r.Re_size[1]++] = i; }
    }
    r.viscous = (r.Re_size[0] > 0 || r.Re_size[1] > 0);
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < r.Re_size[i]; j++) r.Res[i][j] = r.fluid_Re[r.Re_idx[i][j]][i];   // m_rhs.fpp:385-390
    r.b = r.viscous ? 2*r.weno_polyn + 2 : r.weno_polyn + 2;     // :356-360
    for (int d = 0; d < 3; d++) {
        if (d < r.nd) { r.g[d].beg = -r.b; r.g[d].end = r.N[d] + r.b; }
        else { r.g[d].beg = 0; r.g[d].end = 0; }
    }
}

// ---------------------------------------------------------------------------------------
// m_weno.fpp:168-363  s_compute_weno_coefficients
// ---------------------------------------------------------------------------------------
static void compute_weno_coefficients(Rank &r, int dir) {
    WenoCoef &w = r.wc[dir];
    const int polyn = r.weno_polyn;
    const int isb = -r.b, ise = r.N[dir] + r.b;                  // m_weno.fpp:108,139-140
    w.lo = isb + polyn; w.hi = ise - polyn;
    const size_t nc = (size_t)(w.hi - w.lo + 1);
    w.pL.assign(nc*6, 0.0); w.pR.assign(nc*6, 0.0);
    w.dL.assign(nc*3, 0.0); w.dR.assign(nc*3, 0.0); w.bt.assign(nc*9, 0.0);
    if (r.weno_order == 1) return;                               // :105
    const Arr1 &s = r.cb[dir];
    for (int i = isb - 1 + polyn; i <= ise - 1 - polyn; i++) {   // :194,:223
        // D(a,c) = s_cb(i+a) - s_cb(i+c)
        auto D = [&](int a, int c) { return s(i + a) - s(i + c); };
        auto sq = [](double x) { return x*x; };
        const int c = i + 1;
        if (r.weno_order == 3) {                                 // :193-217
            w.polyR(c, 0, 0) = D(0, 1)/D(0, 2);
            w.polyR(c, 1, 0) = D(0, 1)/D(-1, 1);
            w.polyL(c, 0, 0) = -w.polyR(c, 0, 0);
            w.polyL(c, 1, 0) = -w.polyR(c, 1, 0);
            w.dcbR(0, c) = D(-1, 1)/D(-1, 2);
            w.dcbL(0, c) = D(-1, 0)/D(-1, 2);
            w.dcbR(1, c) = 1.0 - w.dcbR(0, c);
            w.dcbL(1, c) = 1.0 - w.dcbL(0, c);
            w.beta(c, 0, 0) = 4.0*sq(D(0, 1))/sq(D(0, 2));
            w.beta(c, 1, 0) = 4.0*sq(D(0, 1))/sq(D(-1, 1));
        } else {                                                 // :223-347
            w.polyR(c, 0, 0) = (D(0, 1)*D(1, 2))/(D(0, 3)*D(3, 1));          // :225-227
            w.polyR(c, 1, 0) = (D(-1, 1)*D(1, 0))/(D(-1, 2)*D(2, 0));        // :228-230
            w.polyR(c, 1, 1) = (D(0, 1)*D(1, 2))/(D(-1, 1)*D(-1, 2));        // :231-233
            w.polyR(c, 2, 1) = (D(0, 1)*D(1, -1))/(D(-2, 0)*D(-2, 1));       // :234-236
            w.polyL(c, 0, 0) = (D(1, 0)*D(0, 2))/(D(0, 3)*D(3, 1));          // :237-239
            w.polyL(c, 1, 0) = (D(0, -1)*D(0, 1))/(D(-1, 2)*D(0, 2));        // :240-242
            w.polyL(c, 1, 1) = (D(1, 0)*D(0, 2))/(D(-1, 1)*D(-1, 2));        // :243-245
            w.polyL(c, 2, 1) = (D(-1, 0)*D(0, 1))/(D(-2, 0)*D(-2, 1));       // :246-248

            w.polyR(c, 0, 1) = (D(0, 2) + D(1, 3))/(D(0, 2)*D(0, 3))*D(0, 1);        // :250-253
            w.polyR(c, 2, 0) = (D(-2, 1) + D(-1, 1))/(D(-1, 1)*D(1, -2))*D(1, 0);    // :254-257
            w.polyL(c, 0, 1) = (D(0, 2) + D(0, 3))/(D(0, 2)*D(0, 3))*D(1, 0);        // :258-261
            w.polyL(c, 2, 0) = (D(-2, 0) + D(-1, 1))/(D(-2, 1)*D(1, -1))*D(0, 1);    // :262-265

            w.dcbR(0, c) = (D(-2, 1)*D(1, -1))/(D(-2, 3)*D(3, -1));          // :267-269
            w.dcbR(2, c) = (D(1, 2)*D(1, 3))/(D(-2, 2)*D(-2, 3));            // :270-272
            w.dcbL(0, c) = (D(-2, 0)*D(0, -1))/(D(-2, 3)*D(3, -1));          // :273-275
            w.dcbL(2, c) = (D(0, 2)*D(0, 3))/(D(-2, 2)*D(-2, 3));            // :276-278
            w.dcbR(1, c) = 1.0 - w.dcbR(0, c) - w.dcbR(2, c);                // :280
            w.dcbL(1, c) = 1.0 - w.dcbL(0, c) - w.dcbL(2, c);                // :281

            const double f = 4.0*sq(D(0, 1));
            w.beta(c, 0, 0) = f*(10.0*sq(D(1, 0)) + D(1, 0)*D(2, 1) + sq(D(2, 1)))
                              /(sq(D(0, 3))*sq(D(1, 3)));                                        // :283-287
            w.beta(c, 0, 1) = f*(19.0*sq(D(1, 0)) - D(1, 0)*D(3, 1) + 2.0*D(2, 0)*(D(2, 0) + D(3, 1)))
                              /(D(0, 2)*sq(D(0, 3))*D(3, 1));                                    // :289-295
            w.beta(c, 0, 2) = f*(10.0*sq(D(1, 0)) + D(1, 0)*(D(2, 0) + D(3, 1)) + sq(D(2, 0) + D(3, 1)))
                              /(sq(D(0, 2))*sq(D(0, 3)));                                        // :297-302
            w.beta(c, 1, 0) = f*(10.0*sq(D(1, 0)) + sq(D(0, -1)) + D(0, -1)*D(1, 0))
                              /(sq(D(-1, 2))*sq(D(0, 2)));                                       // :304-308
            w.beta(c, 1, 1) = f*(D(0, 1)*(D(0, -1) + 20.0*D(1, 0)) + (2.0*D(0, -1) + D(1, 0))*D(2, 0))
                              /(D(1, -1)*sq(D(-1, 2))*D(2, 0));                                  // :310-316
            w.beta(c, 1, 2) = f*(10.0*sq(D(1, 0)) + D(1, 0)*D(2, 1) + sq(D(2, 1)))
                              /(sq(D(-1, 1))*sq(D(-1, 2)));                                      // :318-323
            w.beta(c, 2, 0) = f*(12.0*sq(D(1, 0)) + sq(D(0, -2) + D(0, -1)) + 3.0*(D(0, -2) + D(0, -1))*D(1, 0))
                              /(sq(D(-2, 1))*sq(D(-1, 1)));                                      // :325-331
            w.beta(c, 2, 1) = f*(19.0*sq(D(1, 0)) + (D(0, -2)*D(0, 1)) + 2.0*D(1, -1)*(D(0, -2) + D(1, -1)))
                              /(D(-2, 0)*sq(D(-2, 1))*D(1, -1));                                 // :333-339
            w.beta(c, 2, 2) = f*(10.0*sq(D(1, 0)) + sq(D(0, -1)) + D(0, -1)*D(1, 0))
                              /(sq(D(-2, 0))*sq(D(-2, 1)));                                      // :341-345
        }
    }
}

// ---------------------------------------------------------------------------------------
// allocation: m_time_steppers.fpp:76-118, m_rhs.fpp:123-399, m_riemann_solvers.fpp:381-424
// ---------------------------------------------------------------------------------------
static void alloc_fields(std::vector<Field> &v, int nvar, const Bounds g[3]) {
    v.resize(nvar);
    for (auto &f : v) f.alloc(g[0], g[1], g[2]);
}

static void allocate_rank(Rank &r) {
    Bounds in[3];
    for (int d = 0; d < 3; d++) { in[d].beg = 0; in[d].end = r.N[d]; }
    alloc_fields(r.q_ts[0], r.E, r.g);
    alloc_fields(r.q_ts[1], r.E, r.g);
    alloc_fields(r.q_prim_vf, r.E, r.g);
    alloc_fields(r.rhs_vf, r.E, in);                             // m_time_steppers.fpp:117
    alloc_fields(r.q_cons_qp, r.E, r.g);
    alloc_fields(r.q_prim_qp, r.E, r.g);
    for (int d = 0; d < r.nd; d++) {
        alloc_fields(r.qL_rs[d], r.E, r.g);
        alloc_fields(r.qR_rs[d], r.E, r.g);
    }
    alloc_fields(r.flux, r.E, r.g);                              // m_rhs.fpp:335-339
    r.flux_src_adv.alloc(r.g[0], r.g[1], r.g[2]);                // m_rhs.fpp:349-357
    alloc_fields(r.vel_src, r.nd, r.g);
    if (r.viscous) {
        alloc_fields(r.flux_src, r.E, r.g);                      // only momxb:E_idx used, m_rhs.fpp:341-347
        r.Re_avg[0].alloc(r.g[0], r.g[1], r.g[2]);
        r.Re_avg[1].alloc(r.g[0], r.g[1], r.g[2]);
        for (int i = 0; i < r.nd; i++) {
            for (int dd = 0; dd < r.nd; dd++) {
                alloc_fields(r.dqL[i][dd], r.nd, r.g);           // m_rhs.fpp:265-297
                alloc_fields(r.dqR[i][dd], r.nd, r.g);
            }
            alloc_fields(r.dq_prim_d[i], r.nd, r.g);             // m_rhs.fpp:227-256
            alloc_fields(r.qL_prim[i], r.nd, r.g);               // m_rhs.fpp:183-191
            alloc_fields(r.qR_prim[i], r.nd, r.g);
            alloc_fields(r.dqL_rs[i], r.nd, r.g);                // m_rhs.fpp:300-322
            alloc_fields(r.dqR_rs[i], r.nd, r.g);
        }
    }
}

// ---------------------------------------------------------------------------------------
// m_start_up.fpp:517-655  s_populate_grid_variables_buffers (needs neighbours' ds)
// ---------------------------------------------------------------------------------------
static void populate_grid_buffers(World &w, int ri, int d) {
    Rank &r = w.ranks[ri];
    const int b = r.b, N = r.N[d];
    Arr1 &ds = r.ds[d], &cb = r.cb[d], &cc = r.cc[d];
    // beginning
    if (r.bc[d][0] <= -3) { for (int i = 1; i <= b; i++) ds(-i) = ds(0); }                 // :527-530
    else if (r.bc[d][0] == -2) { for (int i = 1; i <= b; i++) ds(-i) = ds(i - 1); }        // :531-534
    else if (r.bc[d][0] == -1) { for (int i = 1; i <= b; i++) ds(-i) = ds(N - (i - 1)); }  // :535-538
    else {                                                                                  // :540, m_mpi_proxy.fpp:338-455
        const Rank &nb = w.ranks[r.bc[d][0]];
        for (int i = 0; i < b; i++) ds(-b + i) = nb.ds[d](nb.N[d] - b + 1 + i);
    }
    for (int i = 1; i <= b; i++) cb(-1 - i) = cb(-i) - ds(-i);                              // :545-547
    for (int i = 1; i <= b; i++) cc(-i) = cc(1 - i) - (ds(1 - i) + ds(-i))/2.0;             // :550-552
    // end
    if (r.bc[d][1] <= -3) { for (int i = 1; i <= b; i++) ds(N + i) = ds(N); }              // :558-561
    else if (r.bc[d][1] == -2) { for (int i = 1; i <= b; i++) ds(N + i) = ds(N - (i - 1)); }
    else if (r.bc[d][1] == -1) { for (int i = 1; i <= b; i++) ds(N + i) = ds(i - 1); }
    else {
        const Rank &nb = w.ranks[r.bc[d][1]];
        for (int i = 0; i < b; i++) ds(N + 1 + i) = nb.ds[d](i);
    }
    for (int i = 1; i <= b; i++) cb(N + i) = cb(N + (i - 1)) + ds(N + i);                   // :576-578
    for (int i = 1; i <= b; i++) cc(N + i) = cc(N + (i - 1)) + (ds(N + (i - 1)) + ds(N + i))/2.0;   // :581-583
}

World *world_create(const mfc_b200_params_t *gp, const double *const cb_glb[3], int num_procs, std::string &err) {
    World *w = new World();
    w->gp = *gp;
    w->num_procs = num_procs;
    const int nd = gp->num_dims;
    const int Nglb[3] = {gp->m, nd > 1 ? gp->n : 0, nd > 2 ? gp->p : 0};
    for (int d = 0; d < nd; d++) w->cb_glb[d].assign(cb_glb[d], cb_glb[d] + Nglb[d] + 2);
    if (!decompose(num_procs, nd, Nglb, gp->weno_order, w->np)) {
        err = "Unsupported combination of values of num_procs, m, n and weno_order";
        delete w; return nullptr;
    }
    w->ranks.resize(num_procs);
    for (int ri = 0; ri < num_procs; ri++) {
        Rank &r = w->ranks[ri];
        r.rank = ri; r.nd = nd; r.nf = gp->num_fluids;
        r.weno_order = gp->weno_order; r.weno_eps = gp->weno_eps;
        r.time_stepper = gp->time_stepper; r.weno_Re_flux = gp->weno_Re_flux != 0;
        r.run_time_info = gp->run_time_info != 0; r.t_step_stop = gp->t_step_stop;
        for (int i = 0; i < MFC_B200_MAX_FLUIDS; i++) {
            r.gammas[i] = gp->gammas[i]; r.pi_infs[i] = gp->pi_infs[i];
            r.fluid_Re[i][0] = gp->Re[i][0]; r.fluid_Re[i][1] = gp->Re[i][1];
        }
        // MPI_CART_COORDS, row-major (m_mpi_proxy.fpp:221,278)
        if (nd == 1) { r.coords[0] = ri; }
        else if (nd == 2) { r.coords[0] = ri/w->np[1]; r.coords[1] = ri % w->np[1]; }
        else { r.coords[0] = ri/(w->np[1]*w->np[2]); r.coords[1] = (ri/w->np[2]) % w->np[1]; r.coords[2] = ri % w->np[2]; }
        for (int d = 0; d < 3; d++) { r.bc[d][0] = gp->bc[2*d]; r.bc[d][1] = gp->bc[2*d + 1]; r.N[d] = 0; }
        for (int d = 0; d < nd; d++) {
            const int rem = (Nglb[d] + 1) % w->np[d];                                   // :229,:287
            int N = (Nglb[d] + 1)/w->np[d] - 1;                                         // :232,:290
            if (r.coords[d] < rem) N += 1;                                              // :235-239
            r.N[d] = N;
            if (num_procs > 1) {
                int c[3] = {r.coords[0], r.coords[1], r.coords[2]};
                if (r.coords[d] > 0 || gp->bc[2*d] == -1) {                             // :242-247,:300-304
                    c[d] = r.coords[d] - 1; r.bc[d][0] = cart_rank(w->np, c, nd);
                }
                if (r.coords[d] < w->np[d] - 1 || gp->bc[2*d + 1] == -1) {              // :250-255,:307-311
                    c[d] = r.coords[d] + 1; r.bc[d][1] = cart_rank(w->np, c, nd);
                }
            }
            r.start_idx[d] = (r.coords[d] < rem) ? (N + 1)*r.coords[d] : (N + 1)*r.coords[d] + rem;   // :257-263
        }
        init_global_parameters(r);
        allocate_rank(r);
        for (int d = 0; d < nd; d++) {
            const int N = r.N[d], b = r.b;
            r.cb[d].alloc(-1 - b, N + b); r.cc[d].alloc(-b, N + b); r.ds[d].alloc(-b, N + b);   // m_global_parameters.fpp:386-394
            for (int i = -1; i <= N; i++) r.cb[d](i) = w->cb_glb[d][(size_t)(r.start_idx[d] + i + 1)];   // m_start_up.fpp:434
            for (int i = 0; i <= N; i++) r.ds[d](i) = r.cb[d](i) - r.cb[d](i - 1);              // :436
            for (int i = 0; i <= N; i++) r.cc[d](i) = r.cb[d](i - 1) + r.ds[d](i)/2.0;          // :438
        }
    }
    for (int ri = 0; ri < num_procs; ri++)
        for (int d = 0; d < nd; d++) populate_grid_buffers(*w, ri, d);
    for (int ri = 0; ri < num_procs; ri++)
        for (int d = 0; d < nd; d++) compute_weno_coefficients(w->ranks[ri], d);
    return w;
}

// global interior arrays <-> per-rank ghosted fields ------------------------------------
void world_set_q(World &w, const double *const q[]) {
    const int nd = w.gp.num_dims;
    const size_t Mx = (size_t)w.gp.m + 1, Ny = nd > 1 ? (size_t)w.gp.n + 1 : 1;
    for (auto &r : w.ranks)
        for (int i = 0; i < r.E; i++)
            for (int l = 0; l <= r.N[2]; l++)
                for (int k = 0; k <= r.N[1]; k++)
                    for (int j = 0; j <= r.N[0]; j++)
                        r.q_ts[0][i](j, k, l) = q[i][(size_t)(r.start_idx[0] + j) + Mx*((size_t)(r.start_idx[1] + k) + Ny*(size_t)(r.start_idx[2] + l))];
}

static void gather(const World &w, int which, double *const q[]) {
    const int nd = w.gp.num_dims;
    const size_t Mx = (size_t)w.gp.m + 1, Ny = nd > 1 ? (size_t)w.gp.n + 1 : 1;
    for (auto &r : w.ranks)
        for (int i = 0; i < r.E; i++) {
            const Field &f = which == 0 ? r.q_ts[0][i] : which == 1 ? r.q_prim_qp[i] : r.rhs_vf[i];
            for (int l = 0; l <= r.N[2]; l++)
                for (int k = 0; k <= r.N[1]; k++)
                    for (int j = 0; j <= r.N[0]; j++)
                        q[i][(size_t)(r.start_idx[0] + j) + Mx*((size_t)(r.start_idx[1] + k) + Ny*(size_t)(r.start_idx[2] + l))] = f(j, k, l);
        }
}
void world_get_q(const World &w, double *const q[]) { gather(w, 0, q); }
void world_get_prim(const World &w, double *const q[]) { gather(w, 1, q); }
void world_get_rhs(const World &w, double *const rhs[]) { gather(w, 2, rhs); }

// ---------------------------------------------------------------------------------------
// m_rhs.fpp:686-908  s_populate_conservative_variables_buffers, one direction at a time so
// that the emulated ranks can run in lockstep like the blocking MPI_SENDRECVs do
// (m_mpi_proxy.fpp:468-979).  Direction d covers the interior of later directions and the
// ghosted extent of earlier ones (x: rows 0:n; y: columns -b:m+b; z: both ghosted).
// ---------------------------------------------------------------------------------------
static void populate_cons_buffers_dir(World &w, int ri, int d) {
    Rank &r = w.ranks[ri];
    const int b = r.b, N = r.N[d];
    Bounds tr[3];
    for (int e = 0; e < 3; e++) {
        if (e < d) tr[e] = r.g[e];                        // already-filled directions: full ghosted extent (:812)
        else { tr[e].beg = 0; tr[e].end = r.N[e]; }       // not yet filled: interior only (:696)
    }
    const int mom_n = r.momxb + d;                        // momxb (x, :715), momxb+1 (y, :829)
    tr[d].beg = tr[d].end = 0;                            // the (d) loop is the j = 1..buff_size loop below
    for (int side = 0; side < 2; side++) {
        const int bcv = r.bc[d][side];
        const Rank *nb = bcv >= 0 ? &w.ranks[bcv] : nullptr;
        for (int i = 0; i < r.E; i++) {
            Field &q = r.q_cons_qp[i];
            for (int c2 = tr[2].beg; c2 <= tr[2].end; c2++)
            for (int c1 = tr[1].beg; c1 <= tr[1].end; c1++)
            for (int c0 = tr[0].beg; c0 <= tr[0].end; c0++) {
                for (int j = 1; j <= b; j++) {
                    int s[3] = {c0, c1, c2}, g[3] = {c0, c1, c2};
                    g[d] = side == 0 ? -j : N + j;
                    double v;
                    if (bcv <= -3) {                      // ghost-cell extrapolation (:692-702,:744-754)
                        s[d] = side == 0 ? 0 : N;
                        v = q.at(s);
                    } else if (bcv == -2) {               // symmetry (:704-723,:756-778)
                        s[d] = side == 0 ? j - 1 : N - (j - 1);
                        v = q.at(s);
                        if (i == mom_n) v = -v;
                    } else if (bcv == -1) {               // periodic (:725-735,:780-790)
                        s[d] = side == 0 ? N - (j - 1) : j - 1;
                        v = q.at(s);
                    } else {                              // processor boundary (:739,:794 -> m_mpi_proxy.fpp:468)
                        s[d] = side == 0 ? nb->N[d] - (j - 1) : j - 1;
                        v = nb->q_cons_qp[i].at(s);
                    }
                    q.at(g) = v;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// m_variables_conversion.fpp:313-375  s_convert_conservative_to_primitive_variables
// ---------------------------------------------------------------------------------------
static void convert_cons_to_prim(Rank &r) {
    #pragma omp parallel for collapse(3) schedule(static)
    for (int l = r.g[2].beg; l <= r.g[2].end; l++)
    for (int k = r.g[1].beg; k <= r.g[1].end; k++)
    for (int j = r.g[0].beg; j <= r.g[0].end; j++) {
        double alpha_K[MFC_B200_MAX_FLUIDS], alpha_rho_K[MFC_B200_MAX_FLUIDS];
        double dyn_pres_K = 0.0;                                                     // :329
        for (int i = 0; i < r.nf; i++) {                                             // :332-335
            alpha_rho_K[i] = r.q_cons_qp[i](j, k, l);
            alpha_K[i] = r.q_cons_qp[r.advxb + i](j, k, l);
        }
        for (int i = 0; i <= r.contxe; i++) r.q_prim_qp[i](j, k, l) = r.q_cons_qp[i](j, k, l);   // :337-339
        // s_convert_species_to_mixture_variables_acc, :187-227
        double rho_K = 0.0, gamma_K = 0.0, pi_inf_K = 0.0;
        for (int i = 0; i < r.nf; i++) {
            rho_K = rho_K + alpha_rho_K[i];
            gamma_K = gamma_K + alpha_K[i]*r.gammas[i];
            pi_inf_K = pi_inf_K + alpha_K[i]*r.pi_infs[i];
        }
        rho_K = std::max(rho_K, sgm_eps);                                            // :353
        for (int i = r.momxb; i <= r.momxe; i++) {                                   // :357-362
            r.q_prim_qp[i](j, k, l) = r.q_cons_qp[i](j, k, l)/rho_K;
            dyn_pres_K = dyn_pres_K + 5e-1*r.q_cons_qp[i](j, k, l)*r.q_prim_qp[i](j, k, l);
        }
        const double pres = (r.q_cons_qp[r.E_idx](j, k, l) - dyn_pres_K - pi_inf_K)/gamma_K;   // :98-106
        r.q_prim_qp[r.E_idx](j, k, l) = pres;                                        // :366
        for (int i = r.advxb; i <= r.advxe; i++) r.q_prim_qp[i](j, k, l) = r.q_cons_qp[i](j, k, l);   // :368-370
    }
}

// ---------------------------------------------------------------------------------------
// m_weno.fpp:365-542  s_weno  (the reshape copies of s_initialize_weno :554-598 are
// pure data movement and are folded into the strided access here)
//   v: cell averages; vL/vR: left/right cell-boundary values; sweep direction dir over
//   is1 = (g.beg+polyn, g.end-polyn) and the full ghosted transverse extent
//   (m_rhs.fpp:928-937).
// ---------------------------------------------------------------------------------------
static void weno(Rank &r, const Field &v, Field &vL, Field &vR, int dir) {
    const int polyn = r.weno_polyn;
    const int s_beg = r.g[dir].beg + polyn, s_end = r.g[dir].end - polyn;
    WenoCoef &w = r.wc[dir];
    const double eps = r.weno_eps;
    Bounds t[3] = {r.g[0], r.g[1], r.g[2]};
    t[dir].beg = 0; t[dir].end = 0;
    #pragma omp parallel for collapse(3) schedule(static)
    for (int c2 = t[2].beg; c2 <= t[2].end; c2++)
    for (int c1 = t[1].beg; c1 <= t[1].end; c1++)
    for (int c0 = t[0].beg; c0 <= t[0].end; c0++) {
        int c[3] = {c0, c1, c2};
        auto V = [&](int j) { int cc[3] = {c[0], c[1], c[2]}; cc[dir] = j; return v.at(cc); };
        for (int j = s_beg; j <= s_end; j++) {
            int o[3] = {c[0], c[1], c[2]}; o[dir] = j;
            if (r.weno_order == 1) {                                                 // :391-414
                vL.at(o) = V(j); vR.at(o) = V(j);
            } else if (r.weno_order == 3) {                                          // :416-466
                double dvd[2], poly[2], beta[2], alpha[2], omega[2];                 // dvd(-1:0) -> [0]=-1, [1]=0
                dvd[1] = V(j + 1) - V(j);
                dvd[0] = V(j) - V(j - 1);
                poly[0] = V(j) + w.polyL(j, 0, 0)*dvd[1];
                poly[1] = V(j) + w.polyL(j, 1, 0)*dvd[0];
                beta[0] = w.beta(j, 0, 0)*dvd[1]*dvd[1] + eps;
                beta[1] = w.beta(j, 1, 0)*dvd[0]*dvd[0] + eps;
                for (int q = 0; q < 2; q++) alpha[q] = w.dcbL(q, j)/(beta[q]*beta[q]);
                double sum = alpha[0] + alpha[1];
                for (int q = 0; q < 2; q++) omega[q] = alpha[q]/sum;
                vL.at(o) = omega[0]*poly[0] + omega[1]*poly[1];
                poly[0] = V(j) + w.polyR(j, 0, 0)*dvd[1];
                poly[1] = V(j) + w.polyR(j, 1, 0)*dvd[0];
                for (int q = 0; q < 2; q++) alpha[q] = w.dcbR(q, j)/(beta[q]*beta[q]);
                sum = alpha[0] + alpha[1];
                for (int q = 0; q < 2; q++) omega[q] = alpha[q]/sum;
                vR.at(o) = omega[0]*poly[0] + omega[1]*poly[1];
            } else {                                                                 // :468-539
                double dvd1, dvd0, dvdm1, dvdm2, poly[3], beta[3], alpha[3], omega[3];
                dvd1 = V(j + 2) - V(j + 1);                                          // :476-483
                dvd0 = V(j + 1) - V(j);
                dvdm1 = V(j) - V(j - 1);
                dvdm2 = V(j - 1) - V(j - 2);
                poly[0] = V(j) + w.polyL(j, 0, 0)*dvd1 + w.polyL(j, 0, 1)*dvd0;      // :485-493
                poly[1] = V(j) + w.polyL(j, 1, 0)*dvd0 + w.polyL(j, 1, 1)*dvdm1;
                poly[2] = V(j) + w.polyL(j, 2, 0)*dvdm1 + w.polyL(j, 2, 1)*dvdm2;
                beta[0] = w.beta(j, 0, 0)*dvd1*dvd1 + w.beta(j, 0, 1)*dvd1*dvd0      // :495-506
                          + w.beta(j, 0, 2)*dvd0*dvd0 + eps;
                beta[1] = w.beta(j, 1, 0)*dvd0*dvd0 + w.beta(j, 1, 1)*dvd0*dvdm1
                          + w.beta(j, 1, 2)*dvdm1*dvdm1 + eps;
                beta[2] = w.beta(j, 2, 0)*dvdm1*dvdm1 + w.beta(j, 2, 1)*dvdm1*dvdm2
                          + w.beta(j, 2, 2)*dvdm2*dvdm2 + eps;
                for (int q = 0; q < 3; q++) alpha[q] = w.dcbL(q, j)/(beta[q]*beta[q]);   // :508
                double sum = alpha[0] + alpha[1] + alpha[2];
                for (int q = 0; q < 3; q++) omega[q] = alpha[q]/sum;                 // :510
                vL.at(o) = omega[0]*poly[0] + omega[1]*poly[1] + omega[2]*poly[2];   // :514
                poly[0] = V(j) + w.polyR(j, 0, 0)*dvd1 + w.polyR(j, 0, 1)*dvd0;      // :516-524
                poly[1] = V(j) + w.polyR(j, 1, 0)*dvd0 + w.polyR(j, 1, 1)*dvdm1;
                poly[2] = V(j) + w.polyR(j, 2, 0)*dvdm1 + w.polyR(j, 2, 1)*dvdm2;
                for (int q = 0; q < 3; q++) alpha[q] = w.dcbR(q, j)/(beta[q]*beta[q]);   // :526
                sum = alpha[0] + alpha[1] + alpha[2];
                for (int q = 0; q < 3; q++) omega[q] = alpha[q]/sum;                 // :528
                vR.at(o) = omega[0]*poly[0] + omega[1]*poly[1] + omega[2]*poly[2];   // :531
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// m_viscous.fpp:159-370  s_get_viscous
// ---------------------------------------------------------------------------------------
static void get_viscous(Rank &r) {
    const int nd = r.nd;
    // velocities are WENO-reconstructed up front in every direction (:186-198); the results
    // live in qL_rs*/qR_rs* (momentum slots) and, if weno_Re_flux, are copied to qL_prim(i)
    // (s_reconstruct_cell_boundary_values_visc :457-523)
    for (int i = 0; i < nd; i++) {
        for (int v = r.momxb; v <= r.momxe; v++) {
            weno(r, r.q_prim_qp[v], r.qL_rs[i][v], r.qR_rs[i][v], i);
            if (r.weno_Re_flux) {
                r.qL_prim[i][v - r.momxb] = r.qL_rs[i][v];       // :498-517 (copy over is1 x is2; the rest is never read)
                r.qR_prim[i][v - r.momxb] = r.qR_rs[i][v];
            }
        }
    }
    const Bounds *g = r.g;
    if (r.weno_Re_flux) {
        // s_apply_scalar_divergence_theorem :381-455
        for (int i = 0; i < nd; i++) {
            Bounds t[3] = {g[0], g[1], g[2]};
            t[i].beg += 1; t[i].end -= 1;                                            // :413,:440
            for (int v = 0; v < nd; v++) {
                Field &dv = r.dq_prim_d[i][v];
                const Field &vL = r.qL_prim[i][v], &vR = r.qR_prim[i][v];
                for (int l = t[2].beg; l <= t[2].end; l++)
                for (int k = t[1].beg; k <= t[1].end; k++)
                for (int j = t[0].beg; j <= t[0].end; j++) {
                    const int c[3] = {j, k, l};
                    dv(j, k, l) = 1.0/r.ds[i](c[i])*(vR(j, k, l) - vL(j, k, l));     // :417-422,:444-449
                }
            }
        }
    } else {
        // finite-difference face gradients (:219-347); 2-D only in the reference, the 3-D
        // extension applies the same normal/cross pattern to every pair of directions.
        for (int dn = 0; dn < nd; dn++) {
            for (int v = 0; v < nd; v++) {
                const Field &u = r.q_prim_qp[r.momxb + v];
                Field &L = r.dqL[dn][dn][v], &R = r.dqR[dn][dn][v];
                for (int l = g[2].beg; l <= g[2].end; l++)
                for (int k = g[1].beg; k <= g[1].end; k++)
                for (int j = g[0].beg; j <= g[0].end; j++) {
                    int c[3] = {j, k, l}, cm[3] = {j, k, l}, cp[3] = {j, k, l};
                    cm[dn] -= 1; cp[dn] += 1;
                    if (c[dn] >= g[dn].beg + 1)                                      // :224-235,:252-263
                        L.at(c) = (u.at(c) - u.at(cm))/(r.cc[dn](c[dn]) - r.cc[dn](c[dn] - 1));
                    if (c[dn] <= g[dn].end - 1)                                      // :237-248,:265-276
                        R.at(c) = (u.at(cp) - u.at(c))/(r.cc[dn](c[dn] + 1) - r.cc[dn](c[dn]));
                }
            }
        }
        // cross derivatives: derivative along dd seen from the faces normal to dn (:278-347)
        for (int dn = 0; dn < nd; dn++)
        for (int dd = 0; dd < nd; dd++) {
            if (dd == dn) continue;
            for (int v = 0; v < nd; v++) {
                const Field &sL = r.dqL[dd][dd][v], &sR = r.dqR[dd][dd][v];
                Field &L = r.dqL[dn][dd][v], &R = r.dqR[dn][dd][v];
                Bounds t[3] = {g[0], g[1], g[2]};
                t[dd].beg += 1; t[dd].end -= 1;                                      // :280,:297,:314,:332
                for (int l = t[2].beg; l <= t[2].end; l++)
                for (int k = t[1].beg; k <= t[1].end; k++)
                for (int j = t[0].beg; j <= t[0].end; j++) {
                    int c[3] = {j, k, l}, cm[3] = {j, k, l}, cp[3] = {j, k, l};
                    cm[dn] -= 1; cp[dn] += 1;
                    if (c[dn] >= g[dn].beg + 1) {                                    // :283-290
                        double a = (sL.at(c) + sR.at(c) + sL.at(cm) + sR.at(cm));
                        L.at(c) = 25e-2*a;
                    }
                    if (c[dn] <= g[dn].end - 1) {                                    // :300-307
                        double a = (sL.at(cp) + sR.at(cp) + sL.at(c) + sR.at(c));
                        R.at(c) = 25e-2*a;
                    }
                }
            }
        }
        // s_compute_fd_gradient (:52-153, called :350-364) fills dq_prim_d*_qp, which only the
        // weno_Re_flux branch reads (m_rhs.fpp:514-529) -- dead in this branch, not restated.
    }
}

// ---------------------------------------------------------------------------------------
// m_riemann_solvers.fpp:71-351  s_hllc_riemann_solver for sweep direction id.
// The caller passes the RIGHT-face values first (m_rhs.fpp:545-556): the solver's "L" state
// at face j+1/2 is qR_rs(j), its "R" state is qL_rs(j+1).
// ---------------------------------------------------------------------------------------
static void hllc_riemann_solver(Rank &r, int id) {
    std::vector<Field> &qLs = r.qR_rs[id];   // dummy qL_prim_rs*_vf
    std::vector<Field> &qRs = r.qL_rs[id];   // dummy qR_prim_rs*_vf
    const int nd = r.nd, nf = r.nf, N = r.N[id];
    // s_populate_riemann_states_variables_buffers :444-619 (bc == -4 only)
    Bounds in[3];
    for (int d = 0; d < 3; d++) { in[d].beg = 0; in[d].end = r.N[d]; }
    for (int side = 0; side < 2; side++) {
        if (r.bc[id][side] != -4) continue;
        Bounds t[3] = {in[0], in[1], in[2]};
        t[id].beg = t[id].end = 0;
        for (int l = t[2].beg; l <= t[2].end; l++)
        for (int k = t[1].beg; k <= t[1].end; k++)
        for (int j = t[0].beg; j <= t[0].end; j++) {
            int a[3] = {j, k, l}, c[3] = {j, k, l};
            if (side == 0) { a[id] = -1; c[id] = 0; }
            else { a[id] = N + 1; c[id] = N; }
            for (int i = 0; i < r.E; i++) {
                if (side == 0) qLs[i].at(a) = qRs[i].at(c);                          // :480-487,:556-563
                else qRs[i].at(a) = qLs[i].at(c);                                    // :515-523,:588-596
            }
            if (r.viscous) {
                // the derivative arrays get the same treatment (:489-511,:525-548,:565-584,:598-614);
                // dummy dqL = caller's dqR (R-first call)
                for (int dd = 0; dd < nd; dd++)
                    for (int v = 0; v < nd; v++) {
                        if (side == 0) r.dqR[id][dd][v].at(a) = r.dqL[id][dd][v].at(c);
                        else r.dqL[id][dd][v].at(a) = r.dqR[id][dd][v].at(c);
                    }
            }
        }
    }
    // dir_idx / dir_flg :464-470 (0-based; z is the 3-D extension following the same pattern)
    int dir_idx[3]; double dir_flg[3] = {0.0, 0.0, 0.0};
    if (id == 0) { dir_idx[0] = 0; dir_idx[1] = 1; dir_idx[2] = 2; }
    else if (id == 1) { dir_idx[0] = 1; dir_idx[1] = 0; dir_idx[2] = 2; }
    else { dir_idx[0] = 2; dir_idx[1] = 0; dir_idx[2] = 1; }
    dir_flg[id] = 1.0;
    const int idx1 = dir_idx[0];                                                     // :136
    const int contxe = r.contxe, E_idx = r.E_idx;

    Bounds f[3] = {in[0], in[1], in[2]};
    f[id].beg = -1;                                                                  // m_rhs.fpp:534-539
    // s_initialize_riemann_solver :627-670: zero the viscous source flux
    if (r.viscous) {
        for (int i = r.momxb; i <= E_idx; i++)
            for (int l = f[2].beg; l <= f[2].end; l++)
            for (int k = f[1].beg; k <= f[1].end; k++)
            for (int j = f[0].beg; j <= f[0].end; j++) r.flux_src[i](j, k, l) = 0.0;
    }
    #pragma omp parallel for collapse(3) schedule(static)
    for (int l = f[2].beg; l <= f[2].end; l++)
    for (int k = f[1].beg; k <= f[1].end; k++)
    for (int j = f[0].beg; j <= f[0].end; j++) {
        const int cL[3] = {j, k, l};
        int cR[3] = {j, k, l}; cR[id] += 1;
        double vel_L[3], vel_R[3];
        double vel_L_rms = 0.0, vel_R_rms = 0.0;                                     // :138-145
        for (int i = 0; i < nd; i++) {
            vel_L[i] = qLs[contxe + 1 + i].at(cL);
            vel_R[i] = qRs[contxe + 1 + i].at(cR);
            vel_L_rms = vel_L_rms + vel_L[i]*vel_L[i];
            vel_R_rms = vel_R_rms + vel_R[i]*vel_R[i];
        }
        const double pres_L = qLs[E_idx].at(cL), pres_R = qRs[E_idx].at(cR);         // :147-148
        double rho_L = 0.0, gamma_L = 0.0, pi_inf_L = 0.0, rho_R = 0.0, gamma_R = 0.0, pi_inf_R = 0.0;
        for (int i = 0; i < nf; i++) {                                               // :159-167
            rho_L = rho_L + qLs[i].at(cL);
            gamma_L = gamma_L + qLs[E_idx + 1 + i].at(cL)*r.gammas[i];
            pi_inf_L = pi_inf_L + qLs[E_idx + 1 + i].at(cL)*r.pi_infs[i];
            rho_R = rho_R + qRs[i].at(cR);
            gamma_R = gamma_R + qRs[E_idx + 1 + i].at(cR)*r.gammas[i];
            pi_inf_R = pi_inf_R + qRs[E_idx + 1 + i].at(cR)*r.pi_infs[i];
        }
        double Re_L[2], Re_R[2];
        if (r.viscous) {                                                             // :169-200
            for (int i = 0; i < 2; i++) {
                Re_L[i] = dflt_real;
                if (r.Re_size[i] > 0) Re_L[i] = 0.0;
                for (int q = 0; q < r.Re_size[i]; q++)
                    Re_L[i] = qLs[E_idx + 1 + r.Re_idx[i][q]].at(cL)/r.Res[i][q] + Re_L[i];
                Re_L[i] = 1.0/std::max(Re_L[i], sgm_eps);
            }
            for (int i = 0; i < 2; i++) {
                Re_R[i] = dflt_real;
                if (r.Re_size[i] > 0) Re_R[i] = 0.0;
                for (int q = 0; q < r.Re_size[i]; q++)
                    Re_R[i] = qRs[E_idx + 1 + r.Re_idx[i][q]].at(cR)/r.Res[i][q] + Re_R[i];
                Re_R[i] = 1.0/std::max(Re_R[i], sgm_eps);
            }
        }
        const double E_L = gamma_L*pres_L + pi_inf_L + 5e-1*rho_L*vel_L_rms;         // :202
        const double E_R = gamma_R*pres_R + pi_inf_R + 5e-1*rho_R*vel_R_rms;         // :204
        const double H_L = (E_L + pres_L)/rho_L;                                     // :206-207
        const double H_R = (E_R + pres_R)/rho_R;
        // rho_avg, H_avg, gamma_avg, c_avg, vel_avg_rms (:209-218) are computed and never used
        double c_L = ((H_L - 5e-1*vel_L_rms)/gamma_L);                               // :220-223
        double c_R = ((H_R - 5e-1*vel_R_rms)/gamma_R);
        c_L = std::sqrt(c_L);
        c_R = std::sqrt(c_R);
        if (r.viscous) {                                                             // :225-230
            for (int i = 0; i < 2; i++) r.Re_avg[i].at(cL) = 2.0/(1.0/Re_L[i] + 1.0/Re_R[i]);
        }
        const double s_L = std::min(vel_L[idx1] - c_L, vel_R[idx1] - c_R);           // :232-233
        const double s_R = std::max(vel_R[idx1] + c_R, vel_L[idx1] + c_L);
        const double s_S = (pres_R - pres_L + rho_L*vel_L[idx1]*(s_L - vel_L[idx1])  // :235-240
                            - rho_R*vel_R[idx1]*(s_R - vel_R[idx1]))
                           /(rho_L*(s_L - vel_L[idx1]) - rho_R*(s_R - vel_R[idx1]));
        const double s_M = std::min(0.0, s_L), s_P = std::max(0.0, s_R);             // :245
        const double xi_L = (s_L - vel_L[idx1])/(s_L - s_S);                         // :249-250
        const double xi_R = (s_R - vel_R[idx1])/(s_R - s_S);
        const double xi_M = (5e-1 + std::copysign(5e-1, s_S));                       // :254-255
        const double xi_P = (5e-1 - std::copysign(5e-1, s_S));

        for (int i = 0; i <= contxe; i++)                                            // :258-264
            r.flux[i].at(cL) = xi_M*qLs[i].at(cL)*(vel_L[idx1] + s_M*(xi_L - 1.0))
                               + xi_P*qRs[i].at(cR)*(vel_R[idx1] + s_P*(xi_R - 1.0));
        for (int i = 0; i < nd; i++) {                                               // :270-286
            const int idxi = dir_idx[i];
            r.flux[contxe + 1 + idxi].at(cL) =
                xi_M*(rho_L*(vel_L[idx1]*vel_L[idxi]
                             + s_M*(xi_L*(dir_flg[idxi]*s_S + (1.0 - dir_flg[idxi])*vel_L[idxi]) - vel_L[idxi]))
                      + dir_flg[idxi]*(pres_L))
                + xi_P*(rho_R*(vel_R[idx1]*vel_R[idxi]
                               + s_P*(xi_R*(dir_flg[idxi]*s_S + (1.0 - dir_flg[idxi])*vel_R[idxi]) - vel_R[idxi]))
                        + dir_flg[idxi]*(pres_R));
        }
        r.flux[E_idx].at(cL) =                                                       // :291-299
            xi_M*(vel_L[idx1]*(E_L + pres_L)
                  + s_M*(xi_L*(E_L + (s_S - vel_L[idx1])*(rho_L*s_S + pres_L/(s_L - vel_L[idx1]))) - E_L))
            + xi_P*(vel_R[idx1]*(E_R + pres_R)
                    + s_P*(xi_R*(E_R + (s_S - vel_R[idx1])*(rho_R*s_S + pres_R/(s_R - vel_R[idx1]))) - E_R));
        for (int i = r.advxb; i <= r.advxe; i++)                                     // :304-310
            r.flux[i].at(cL) = xi_M*qLs[i].at(cL)*(vel_L[idx1] + s_M*(xi_L - 1.0))
                               + xi_P*qRs[i].at(cR)*(vel_R[idx1] + s_P*(xi_R - 1.0));
        for (int i = 0; i < nd; i++) {                                               // :314-324
            const int idxi = dir_idx[i];
            r.vel_src[idxi].at(cL) = xi_M*(vel_L[idxi] + dir_flg[idxi]*s_M*(xi_L - 1.0))
                                     + xi_P*(vel_R[idxi] + dir_flg[idxi]*s_P*(xi_R - 1.0));
        }
        r.flux_src_adv.at(cL) = r.vel_src[idx1].at(cL);                              // :325
    }

    // s_compute_viscous_source_flux :683-902.  dvelL_* = caller's dqR_prim_*_n(id), dvelR_* =
    // caller's dqL_prim_*_n(id) (R-first call, m_rhs.fpp:545-556 -> :333-345).
    if (r.viscous) {
        if (nd == 3) return;   // no reference for 3-D viscous stresses; the extension is inviscid only
        for (int l = f[2].beg; l <= f[2].end; l++)
        for (int k = f[1].beg; k <= f[1].end; k++)
        for (int j = f[0].beg; j <= f[0].end; j++) {
            const int c[3] = {j, k, l};
            int cp[3] = {j, k, l}; cp[id] += 1;
            auto avg = [&](int dd, int v) {        // 5d-1*(dvelL_d?(v)(j,k) + dvelR_d?(v)(j+1,k))
                return 5e-1*(r.dqR[id][dd][v].at(c) + r.dqL[id][dd][v].at(cp));
            };
            Field *fs = r.flux_src.data();
            const int momxb = r.momxb;
            if (id == 0) {
                if (r.Re_size[0] > 0) {                                              // :714-736
                    const double tau = (4.0/3.0)*avg(0, 0)/r.Re_avg[0].at(c);
                    fs[momxb].at(c) = fs[momxb].at(c) - tau;
                    fs[E_idx].at(c) = fs[E_idx].at(c) - r.vel_src[0].at(c)*tau;
                }
                if (r.Re_size[1] > 0) {                                              // :738-760
                    const double tau = avg(0, 0)/r.Re_avg[1].at(c);
                    fs[momxb].at(c) = fs[momxb].at(c) - tau;
                    fs[E_idx].at(c) = fs[E_idx].at(c) - r.vel_src[0].at(c)*tau;
                }
                if (nd == 1) continue;                                               // :762
                if (r.Re_size[0] > 0) {                                              // :764-801
                    const double dy0 = avg(1, 0), dy1 = avg(1, 1), dx1 = avg(0, 1);
                    double tau[2];
                    tau[0] = -(2.0/3.0)*dy1/r.Re_avg[0].at(c);
                    tau[1] = (dy0 + dx1)/r.Re_avg[0].at(c);
                    for (int i = 0; i < 2; i++) {
                        fs[contxe + 1 + i].at(c) = fs[contxe + 1 + i].at(c) - tau[i];
                        fs[E_idx].at(c) = fs[E_idx].at(c) - r.vel_src[i].at(c)*tau[i];
                    }
                }
                if (r.Re_size[1] > 0) {                                              // :803-825
                    const double tau = avg(1, 1)/r.Re_avg[1].at(c);
                    fs[momxb].at(c) = fs[momxb].at(c) - tau;
                    fs[E_idx].at(c) = fs[E_idx].at(c) - r.vel_src[0].at(c)*tau;
                }
            } else {
                if (r.Re_size[0] > 0) {                                              // :831-872
                    const double dx0 = avg(0, 0), dx1 = avg(0, 1), dy0 = avg(1, 0), dy1 = avg(1, 1);
                    double tau[2];
                    tau[0] = (dy0 + dx1)/r.Re_avg[0].at(c);
                    tau[1] = (4.0*dy1 - 2.0*dx0)/(3.0*r.Re_avg[0].at(c));
                    for (int i = 0; i < 2; i++) {
                        fs[contxe + 1 + i].at(c) = fs[contxe + 1 + i].at(c) - tau[i];
                        fs[E_idx].at(c) = fs[E_idx].at(c) - r.vel_src[i].at(c)*tau[i];
                    }
                }
                if (r.Re_size[1] > 0) {                                              // :874-899
                    const double tau = (avg(0, 0) + avg(1, 1))/r.Re_avg[1].at(c);
                    fs[momxb + 1].at(c) = fs[momxb + 1].at(c) - tau;
                    fs[E_idx].at(c) = fs[E_idx].at(c) - r.vel_src[1].at(c)*tau;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// m_rhs.fpp:405-679  s_compute_rhs, split where the MPI halo exchange sits
// ---------------------------------------------------------------------------------------
static void rhs_part1_copy(Rank &r, int stage) {
    for (int i = 0; i < r.E; i++) r.q_cons_qp[i].a = r.q_ts[stage][i].a;            // :425-432
}

static void rhs_part2(Rank &r, int t_step) {
    convert_cons_to_prim(r);                                                         // :445-447
    if (t_step == r.t_step_stop) return;                                             // :452
    const int nd = r.nd;
    if (r.viscous) get_viscous(r);                                                   // :456-464

    for (int id = 0; id < nd; id++) {                                                // :469
        // reconstruction (:480-531).  In the viscous branch the velocities were already
        // reconstructed by s_get_viscous, the remaining variables are done here in three
        // groups -- the arithmetic per variable is identical.
        for (int v = 0; v < r.E; v++) {
            if (r.viscous && v >= r.momxb && v <= r.momxe) continue;
            weno(r, r.q_prim_qp[v], r.qL_rs[id][v], r.qR_rs[id][v], id);
        }
        if (r.viscous && r.weno_Re_flux) {                                           // :513-529
            for (int dd = 0; dd < nd; dd++)
                for (int v = 0; v < nd; v++) {
                    weno(r, r.dq_prim_d[dd][v], r.dqL_rs[id][v], r.dqR_rs[id][v], id);
                    r.dqL[id][dd][v] = r.dqL_rs[id][v];          // m_viscous.fpp:563-587
                    r.dqR[id][dd][v] = r.dqR_rs[id][v];
                }
        }
        hllc_riemann_solver(r, id);                                                  // :545-556

        const Arr1 &ds = r.ds[id];
        const int N0 = r.N[0], N1 = r.N[1], N2 = r.N[2];
        #pragma omp parallel for collapse(3) schedule(static)
        for (int l = 0; l <= N2; l++)
        for (int k = 0; k <= N1; k++)
        for (int j = 0; j <= N0; j++) {
            const int c[3] = {j, k, l};
            int cm[3] = {j, k, l}; cm[id] -= 1;
            const double dsj = ds(c[id]);
            for (int i = 0; i < r.E; i++) {
                if (id == 0)                                                         // :567-576
                    r.rhs_vf[i](j, k, l) = 1.0/dsj*(r.flux[i].at(cm) - r.flux[i].at(c));
                else                                                                 // :610-620
                    r.rhs_vf[i](j, k, l) = r.rhs_vf[i](j, k, l) + 1.0/dsj*(r.flux[i].at(cm) - r.flux[i].at(c));
            }
            for (int i = r.advxb; i <= r.advxe; i++)                                 // :578-589,:624-635
                r.rhs_vf[i](j, k, l) = r.rhs_vf[i](j, k, l)
                                       + 1.0/dsj*r.q_cons_qp[i](j, k, l)*(r.flux_src_adv.at(c) - r.flux_src_adv.at(cm));
            if (r.viscous)                                                           // :591-604,:639-652
                for (int i = r.momxb; i <= r.E_idx; i++)
                    r.rhs_vf[i](j, k, l) = r.rhs_vf[i](j, k, l)
                                           + 1.0/dsj*(r.flux_src[i].at(cm) - r.flux_src[i].at(c));
        }
    }
    if (r.run_time_info)                                                             // :659-675
        for (int i = 0; i < r.E; i++) r.q_prim_vf[i].a = r.q_prim_qp[i].a;
}

void world_compute_rhs(World &w, int stage, int t_step) {
    for (auto &r : w.ranks) rhs_part1_copy(r, stage);
    for (int d = 0; d < w.gp.num_dims; d++)                                          // :435 -> :686-908
        for (int ri = 0; ri < w.num_procs; ri++) populate_cons_buffers_dir(w, ri, d);
    for (auto &r : w.ranks) rhs_part2(r, t_step);
}

// ---------------------------------------------------------------------------------------
// m_data_output.fpp:178-310  s_write_run_time_information (stability criteria only)
// ---------------------------------------------------------------------------------------
static void run_time_information(Rank &r, double dt) {
    double icfl = -1e300, vcfl = -1e300, Rc = 1e300;
    const int nd = r.nd;
    for (int l = 0; l <= r.N[2]; l++)
    for (int k = 0; k <= r.N[1]; k++)
    for (int j = 0; j <= r.N[0]; j++) {
        double alpha_rho[MFC_B200_MAX_FLUIDS], alpha[MFC_B200_MAX_FLUIDS], vel[3], Re[2];
        for (int i = 0; i < r.nf; i++) {                                             // :201-204
            alpha_rho[i] = r.q_prim_vf[i](j, k, l);
            alpha[i] = r.q_prim_vf[r.E_idx + 1 + i](j, k, l);
        }
        double rho = 0.0, gamma = 0.0, pi_inf = 0.0;                                 // :206 -> m_variables_conversion.fpp:187-227
        for (int i = 0; i < r.nf; i++) {
            rho = rho + alpha_rho[i];
            gamma = gamma + alpha[i]*r.gammas[i];
            pi_inf = pi_inf + alpha[i]*r.pi_infs[i];
        }
        if (r.viscous) {
            for (int i = 0; i < 2; i++) {
                Re[i] = dflt_real;
                if (r.Re_size[i] > 0) Re[i] = 0.0;
                for (int q = 0; q < r.Re_size[i]; q++) Re[i] = alpha[r.Re_idx[i][q]]/r.Res[i][q] + Re[i];
                Re[i] = 1.0/std::max(Re[i], sgm_eps);
            }
        }
        for (int i = 0; i < nd; i++) vel[i] = r.q_prim_vf[r.contxe + 1 + i](j, k, l);   // :208-210
        const double pres = r.q_prim_vf[r.E_idx](j, k, l);
        double c = (((gamma + 1.0)*pres + pi_inf)/(gamma*rho));                      // :215-216
        c = std::sqrt(c);
        double icfl_c, vcfl_c = 0, Rc_c = 0;
        if (nd == 3) {       // extension: same form as 2-D with the third direction added
            icfl_c = dt/std::min(std::min(r.ds[0](j)/(std::fabs(vel[0]) + c), r.ds[1](k)/(std::fabs(vel[1]) + c)),
                                 r.ds[2](l)/(std::fabs(vel[2]) + c));
        } else if (nd == 2) {                                                        // :218-230
            icfl_c = dt/std::min(r.ds[0](j)/(std::fabs(vel[0]) + c), r.ds[1](k)/(std::fabs(vel[1]) + c));
            if (r.viscous) {
                const double mn = std::min(r.ds[0](j), r.ds[1](k));
                vcfl_c = std::max(dt/Re[0], dt/Re[1])/(mn*mn);
                Rc_c = std::min(r.ds[0](j)*(std::fabs(vel[0]) + c), r.ds[1](k)*(std::fabs(vel[1]) + c))
                       /std::max(1.0/Re[0], 1.0/Re[1]);
            }
        } else {                                                                     // :231-241
            icfl_c = (dt/r.ds[0](j))*(std::fabs(vel[0]) + c);
            if (r.viscous) {
                vcfl_c = std::max(dt/Re[0], dt/Re[1])/(r.ds[0](j)*r.ds[0](j));
                Rc_c = r.ds[0](j)*(std::fabs(vel[0]) + c)/std::max(1.0/Re[0], 1.0/Re[1]);
            }
        }
        icfl = std::max(icfl, icfl_c);                                               // :249-258
        if (r.viscous) { vcfl = std::max(vcfl, vcfl_c); Rc = std::min(Rc, Rc_c); }
    }
    r.icfl_max_loc = icfl; r.vcfl_max_loc = vcfl; r.Rc_min_loc = Rc;
}

// ---------------------------------------------------------------------------------------
// m_time_steppers.fpp:129-362  s_1st/2nd/3rd_order_tvd_rk
// ---------------------------------------------------------------------------------------
template <class F>
static void update(World &w, F f) {
    for (auto &r : w.ranks)
        for (int i = 0; i < r.E; i++) {
            Field &q1 = r.q_ts[0][i], &q2 = r.q_ts[1][i];
            const Field &rhs = r.rhs_vf[i];
            #pragma omp parallel for collapse(3) schedule(static)
            for (int l = 0; l <= r.N[2]; l++)
            for (int k = 0; k <= r.N[1]; k++)
            for (int j = 0; j <= r.N[0]; j++) f(q1(j, k, l), q2(j, k, l), rhs(j, k, l));
        }
}

void world_step(World &w, int t_step, double dt, double stab[3]) {
    const int ts = w.gp.time_stepper;
    world_compute_rhs(w, 0, t_step);                                                 // :143,:211,:285
    if (w.gp.run_time_info) {                                                        // :149-151,:213-215,:288-290
        double icfl = -1e300, vcfl = -1e300, Rc = 1e300;
        for (auto &r : w.ranks) {
            run_time_information(r, dt);
            icfl = std::max(icfl, r.icfl_max_loc);                                   // m_mpi_common.fpp:135-171
            vcfl = std::max(vcfl, r.vcfl_max_loc);
            Rc = std::min(Rc, r.Rc_min_loc);
        }
        w.icfl_max_glb = icfl; w.vcfl_max_glb = vcfl; w.Rc_min_glb = Rc;
        if (stab) { stab[0] = icfl; if (w.ranks[0].viscous) { stab[1] = vcfl; stab[2] = Rc; } }
    }
    if (t_step == w.gp.t_step_stop) return;                                          // :161,:221,:296
    if (ts == 1) {
        update(w, [dt](double &q1, double &, double rhs) { q1 = q1 + dt*rhs; });     // :163-172
    } else if (ts == 2) {
        update(w, [dt](double &q1, double &q2, double rhs) { q2 = q1 + dt*rhs; });   // :223-232
        world_compute_rhs(w, 1, t_step);                                             // :239
        update(w, [dt](double &q1, double &q2, double rhs) { q1 = (q1 + q2 + dt*rhs)/2.0; });   // :241-251
    } else {
        update(w, [dt](double &q1, double &q2, double rhs) { q2 = q1 + dt*rhs; });   // :298-307
        world_compute_rhs(w, 1, t_step);                                             // :315
        update(w, [dt](double &q1, double &q2, double rhs) { q2 = (3.0*q1 + q2 + dt*rhs)/4.0; });   // :318-328
        world_compute_rhs(w, 1, t_step);                                             // :335
        update(w, [dt](double &q1, double &q2, double rhs) { q1 = (q1 + 2.0*q2 + 2.0*dt*rhs)/3.0; });   // :338-348
    }
}

}  // namespace orc
