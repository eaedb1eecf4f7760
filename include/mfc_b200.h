/*
 * mfc_b200.h -- C ABI of libmfc_b200.so, the B200-native (sm_100a) replacement for
 * MicroFC's right-hand-side evaluation and time stepping.
 *
 * The reference has no FFI/plugin interface: the hot path is reached by plain Fortran
 * module calls from the unchanged host program (src/simulation/p_main.fpp).  Each entry
 * point below replaces one of those call sites and is bound from Fortran with
 * ISO_C_BINDING (fortran/m_b200_bindings.f90).  Citations are
 * relative to the reference tree.
 *
 * Conventions
 *   - every entry returns 0 on success or a negative MFC_B200_E* code; nothing throws
 *     across the ABI.  mfc_b200_last_error() returns a message for the last failure
 *     (the Fortran wrapper prints it and calls s_mpi_abort(), m_mpi_common.fpp:307-320).
 *   - the host owns host memory, the library owns all device memory and never keeps a
 *     host pointer past the call that received it.
 *   - one host thread per rank drives one GPU (p_main.fpp:84-100); entries are not
 *     re-entrant; library state is a per-process singleton, like the Fortran module
 *     globals it replaces.
 *   - a "field" is one contiguous Fortran array sf(-b:m+b, -b:n+b [, -b:p+b]) with x
 *     fastest (m_time_steppers.fpp:84-85); inactive dimensions have extent 1.  Field
 *     arguments are arrays of sys_size base pointers (c_loc(q_cons_ts(1)%vf(i)%sf)).
 *   - the library FAILS (MFC_B200_ENODEVICE) when no CUDA device is usable: there is
 *     no CPU fallback.
 */
#ifndef MFC_B200_H
#define MFC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MFC_B200_MAX_FLUIDS 4     /* array extents of the ABI structs (num_fluids_max is 10 in the reference) */
#define MFC_B200_BUILT_FLUIDS 3   /* sweep kernels are instantiated for num_fluids = 1, 2, 3 (the shipped examples use 1 and 2);
                                     mfc_b200_init fails with MFC_B200_EUNSUPPORTED above that */
#define MFC_B200_ABI_VERSION 1

enum {
    MFC_B200_OK = 0,
    MFC_B200_EINVAL = -1,      /* bad parameter combination (mirrors s_check_input_file) */
    MFC_B200_ENODEVICE = -2,   /* no CUDA device / kernels missing: no CPU fallback */
    MFC_B200_ECUDA = -3,       /* CUDA runtime error */
    MFC_B200_ESTATE = -4,      /* call order violated (e.g. step before init/upload) */
    MFC_B200_ENCCL = -5,       /* NCCL error */
    MFC_B200_ENOMEM = -6,
    MFC_B200_EUNSUPPORTED = -7 /* feature present in the reference but not built yet */
};

/*
 * Everything the hot path reads implicitly from module globals in the reference
 * (m_global_parameters.fpp:32-215), AFTER domain decomposition
 * (m_mpi_proxy.fpp:134-328): m/n/p are LOCAL, bc_* < 0 are physical boundary
 * codes, bc_* >= 0 are neighbour ranks.
 */
typedef struct mfc_b200_params {
    int32_t abi_version;       /* MFC_B200_ABI_VERSION */

    /* local grid: cells 0..m, 0..n, 0..p  (m_global_parameters.fpp:50); n = p = 0 when unused.
       p > 0 (3D) is an EXTENSION beyond the reference (it is 1D/2D only). */
    int32_t m, n, p;
    int32_t m_glb, n_glb, p_glb;
    int32_t num_dims;          /* 1 + min(1,n) [+ min(1,p)]   (m_global_parameters.fpp:403) */
    int32_t num_fluids;
    int32_t sys_size;          /* 2*num_fluids + num_dims + 1 (m_global_parameters.fpp:302-310) */
    int32_t buff_size;         /* weno_polyn+2, or 2*weno_polyn+2 if viscous (:356-360) */

    int32_t weno_order;        /* 1, 3 or 5 */
    double  weno_eps;
    int32_t time_stepper;      /* 1, 2 or 3 */
    int32_t weno_Re_flux;      /* logical */
    int32_t run_time_info;     /* logical: compute ICFL (VCFL, Rc) after stage 1 */
    int32_t t_step_start;
    int32_t t_step_stop;       /* the step at which s_*_tvd_rk returns before updating */

    /* boundary conditions / neighbours, order: x_beg, x_end, y_beg, y_end, z_beg, z_end */
    int32_t bc[6];
    int32_t proc_rank, num_procs;
    int32_t proc_coords[3];
    int32_t num_procs_dir[3];

    /* stiffened-gas parameters (m_variables_conversion.fpp:253-260): gamma = 1/(g-1),
       pi_inf = g*pinf/(g-1); Re(i,1:2) shear/bulk, <= 0 (dflt_real) means inviscid */
    double gammas[MFC_B200_MAX_FLUIDS];
    double pi_infs[MFC_B200_MAX_FLUIDS];
    double Re[MFC_B200_MAX_FLUIDS][2];

    /* ghosted metric arrays exactly as allocated in m_global_parameters.fpp:386-394 and
       filled by s_populate_grid_variables_buffers (m_start_up.fpp:517-655):
         cb[d] -> s_cb(-1-b : N_d+b)   (N_d + 2 + 2b doubles)
         cc[d] -> s_cc(-b   : N_d+b)   (N_d + 1 + 2b doubles)
         ds[d] -> ds  (-b   : N_d+b)   (N_d + 1 + 2b doubles)
       the pointer is the address of the FIRST element (c_loc(x_cb(-1-buff_size))).
       Unused directions may be NULL. */
    const double *cb[3];
    const double *cc[3];
    const double *ds[3];

    /* extensions (0 = reference behaviour) */
    int32_t strict_math;       /* 1: no FMA contraction and the reference's exact operation
                                  order (bit-comparable with a strict CPU build); 0: fast */
    int32_t device;            /* CUDA device ordinal; -1 = local_rank mod device count */
    int32_t reserved[6];
} mfc_b200_params_t;

/* p_main.fpp:131-151,175 -- module initialisers (scratch, WENO coefficients).  Call once,
   after s_populate_grid_variables_buffers (p_main.fpp:170) because the WENO coefficients
   need the ghosted x_cb/y_cb (m_weno.fpp:103-159,168-363). */
int mfc_b200_init(const mfc_b200_params_t *params);

/* Multi-GPU bootstrap, replaces MPI_INIT + MPI_CART_CREATE (m_mpi_common.fpp:47-59,
   m_mpi_proxy.fpp:215-222).  Rank 0 obtains a 128-byte NCCL unique id, the host
   broadcasts it with whatever it already has (MPI_BCAST in the Fortran host;
   torch.distributed in the Python driver), every rank then calls mfc_b200_comm_init.
   Must follow mfc_b200_init.  Not needed when num_procs == 1. */
int mfc_b200_get_unique_id(unsigned char id[128]);
int mfc_b200_comm_init(const unsigned char id[128], int rank, int nranks);

/* p_main.fpp:188-193 -- "!$acc update device(q_cons_ts(1)%vf(i)%sf)": H2D of the
   conservative state (ghost cells are ignored; they are rebuilt every RHS,
   m_rhs.fpp:425-435). */
int mfc_b200_upload(const double *const q_cons[/*sys_size*/]);

/* p_main.fpp:229-235 -> s_{1st,2nd,3rd}_order_tvd_rk(t_step, time_avg)
   (m_time_steppers.fpp:129,197,271).  dt is passed BY VALUE every step because the host
   mutates it (p_main.fpp:287).  When t_step == t_step_stop the RHS is evaluated up to the
   primitive conversion and NO update is done (m_time_steppers.fpp:296, m_rhs.fpp:452).
   stab (nullable): receives {ICFL max, VCFL max, Rc min} reduced over all ranks when
   run_time_info is set (m_data_output.fpp:197-274); entries not computed are left as is.
   step_seconds (nullable): device time of this call, the quantity time_avg averages
   (m_time_steppers.fpp:281,352-358). */
int mfc_b200_step(int t_step, double dt, double stab[3], double *step_seconds);

/* Asynchronous variant for hosts that do not need the diagnostics every step: enqueues
   n_steps consecutive RK steps starting at t_step with a constant dt and returns without
   synchronising.  mfc_b200_sync() waits for them. */
int mfc_b200_step_async(int t_step, double dt, int n_steps);
int mfc_b200_sync(void);

/* m_time_steppers.fpp:285 -> s_compute_rhs(q_cons_vf, q_prim_vf, rhs_vf, t_step)
   (m_rhs.fpp:405).  Stateless with respect to the stepper: q (ghosted fields, ghosts
   ignored) is uploaded to a scratch state, rhs (fields of shape (0:m,0:n[,0:p])) comes
   back.  Exported for parity tests and for hosts that keep their own stepper. */
int mfc_b200_compute_rhs(const double *const q_cons[], double *const rhs[]);

/* p_main.fpp:218,296 and m_time_steppers.fpp:374 -- "!$acc update host(...)". */
int mfc_b200_download(double *const q_cons[/*sys_size*/]);
int mfc_b200_download_prim(double *const q_prim[/*sys_size*/]);

/* ---- initial condition on the device -------------------------------------
   The reference's pre_process executable lays patch_icpp(1..num_patches) over the grid
   (src/pre_process/m_initial_condition.fpp:42-113, m_create_patches.fpp, m_assign_patches.fpp:
   54-165), converts to conservative variables (src/common/m_variables_conversion.fpp:385-443)
   and writes restart files that the simulation reads back and uploads.  This entry does the
   same on the device, straight into the library's state, and replaces that file round trip +
   mfc_b200_upload for hosts that start from patches: 512^3 x 8 variables per GPU never exist on
   the host.  Fields mirror patch_icpp (src/pre_process/m_global_parameters.fpp:196-230).
   Geometries: 1 line segment, 15 1-D analytical; 2 circle, 3 rectangle, 4 sweep line, 5 ellipse,
   6 isentropic vortex, 7 2-D analytical, 18 varcircle (everything m_initial_condition.fpp:50-100
   dispatches); 8 sphere, 9 cuboid, 10 z-invariant cylinder are the 3-D extension.  Anything else fails with
   MFC_B200_EUNSUPPORTED. */
#define MFC_B200_MAX_PATCHES 10   /* num_patches_max */
typedef struct mfc_b200_patch {
    int32_t geometry;
    int32_t smoothen;          /* logical */
    int32_t smooth_patch_id;   /* 1-based id of the patch this one is smeared against */
    int32_t alter_patch[MFC_B200_MAX_PATCHES + 1];   /* alter_patch(0:num_patches), logical */
    double x_centroid, y_centroid, z_centroid;
    double length_x, length_y, length_z;
    double radius;
    double radii[3];
    double normal[3];
    double epsilon;
    double smooth_coeff;
    double vel[3];
    double pres;
    double alpha_rho[MFC_B200_MAX_FLUIDS];
    double alpha[MFC_B200_MAX_FLUIDS];
} mfc_b200_patch_t;

/* cc[d]: pre_process' cell centres (x_cb(i-1) + x_cb(i))/2 (m_start_up.fpp:717,743) of THIS
   rank's interior cells, N_d + 1 doubles per active direction; ds_min: the smallest cell width
   of the GLOBAL grid over all active directions (the smoothing length scale,
   m_create_patches.fpp:123).  Must follow mfc_b200_init; leaves the library in the same state
   as mfc_b200_upload of the generated fields. */
int mfc_b200_generate_initial_condition(int32_t num_patches, const mfc_b200_patch_t *patches,
                                        const double *const cc[3], double ds_min);
/* The same with cb[d] = the RIGHT boundaries x_cb(0:m), y_cb(0:n) of this rank's interior cells,
   which the analytical patches (geometry 7 s_2D_analytical, 15 s_1D_analytical,
   m_create_patches.fpp:424-534) evaluate their pressure bump at.  cb may be NULL when no such
   patch is present. */
int mfc_b200_generate_initial_condition2(int32_t num_patches, const mfc_b200_patch_t *patches,
                                         const double *const cc[3], const double *const cb[3], double ds_min);

/* p_main.fpp:329-341 -- module finalisers. */
int mfc_b200_finalize(void);

const char *mfc_b200_last_error(void);

/* ---- introspection used by tests and bench.py (not part of the Fortran binding) ---- */

/* The grid-dependent WENO coefficients the library computed for direction dir (0,1,2),
   in the reference's array shapes flattened C-order over (cell, k, q):
   poly_L/poly_R: (ncell,3,2), d_L/d_R: (ncell,3), beta: (ncell,3,3) for WENO5, where
   cell runs over -b+polyn .. N+b-polyn (m_weno.fpp:118-127). Any pointer may be NULL. */
int mfc_b200_get_weno_coefficients(int dir, double *poly_L, double *poly_R,
                                   double *d_L, double *d_R, double *beta);

/* Number of kernels this library launched since init (bench.py's gpu_launches). */
int64_t mfc_b200_kernel_launches(void);

/* Device-resident helpers for benchmarking with inputs already in HBM: snapshot the
   current state on the device / restore it, without host traffic. */
int mfc_b200_state_snapshot(void);
int mfc_b200_state_restore(void);

/* CUDA-event stopwatch on the library's own launching stream (torch.cuda.Event only sees
   torch's current stream): _start records an event, _stop records another, synchronises and
   returns the device time between them. */
int mfc_b200_timer_start(void);
int mfc_b200_timer_stop(double *seconds);

/* Time of the dominant kernels accumulated with CUDA events on the launching stream
   since the last reset: out[0..n) seconds per kernel class, names via _kernel_name. */
int mfc_b200_profile_enable(int on);
int mfc_b200_profile_get(int kernel_class, double *seconds, int64_t *launches);
const char *mfc_b200_kernel_name(int kernel_class);

/* Test hook: rebuild the ghost cells of q_cons_ts(1) in place (a following mfc_b200_download
   returns them).  mode 0: the production ghost fill (m_rhs.fpp:686-908 / the NCCL halo
   exchange).  mode 1: every periodic direction is filled by the halo-exchange kernels
   (pack -> device copy in place of ncclSend/ncclRecv to myself -> unpack,
   m_mpi_proxy.fpp:490-601,733-969) instead of the periodic boundary kernel; the two must agree
   bit for bit, which checks the x / y / z pack and unpack index maps on a single GPU. */
int mfc_b200_debug_fill_ghosts(int mode);

#ifdef __cplusplus
}
#endif
#endif /* MFC_B200_H */
