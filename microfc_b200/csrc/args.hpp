// args.hpp -- kernel argument blocks and the launcher table shared by the API and the two
// kernel builds (fast / strict).
#pragma once
#include <cuda_runtime.h>
#include "layout.hpp"

namespace mfc {

// CUtensorMap (128 opaque bytes, 64-byte aligned) without pulling <cuda.h> into device code
struct alignas(64) TensorMap {
    unsigned long long opaque[16];
};

struct SweepArgs {
    GridDesc g;
    const double *q;       // stage state, conservative (alpha_rho, mom, E, alpha), ghosts filled
    const double *prim;    // velocity (nd fields) + pressure of the stage state (viscous runs: k_prim)
    double *rhs;           // accumulated RHS (same padded layout)
    const double *q1;      // q_cons_ts(1) for the fused RK update
    double *qout;          // updated state
    const double *coef;    // kNumWenoCoef arrays of clen doubles for this direction
    const double *rds;     // 1/ds(-b:N+b) of this direction
    int clen, coef_lo;     // coefficient arrays cover cells coef_lo .. coef_lo+clen-1
    double eps, dt;
    double gammas[kMaxFluids], pi_infs[kMaxFluids];
    int bc_beg, bc_end;    // boundary codes of this direction (Riemann-state extrapolation, -4)
    // y / z march: effective code (-1, -2, <= -3) of a PHYSICAL boundary whose ghost rows the march
    // synthesises by streaming the source row instead (no k_bc launch for this direction); 0 = read
    // the ghost rows from memory
    int map_beg, map_end;
    int first_dir;         // 1: RHS assigned (m_rhs.fpp:567-576), 0: accumulated (:610-620)
    int rk_mode;           // 0: store RHS; 1..4: fused update, see rk_apply()
    int seg;               // cells per thread along the sweep (march kernels)
    int rows;              // rows per block (x kernel)
    // x kernel: warp work items = xs_ntx x tiles x row blocks of `rows` rows (xs_items per plane).
    // Tiles are balanced over cells xs_lo .. xs_hi, or (xs_strip > 0) the two boundary strips
    // [0, xs_strip-1] and [N-xs_strip+1, N].  xsplit (set by the API for multi-rank runs whose x
    // halo is still in flight) = 1: the cells that read no x ghost column, 2: the two boundary
    // strips, 0: everything in one launch
    int xsplit, xs_ntx, xs_lo, xs_hi, xs_strip, xs_items;
    // rows / planes this x launch covers (multi-rank runs sweep x piece by piece as the x halo arrives,
    // viscous runs sweep the rows that need no y ghosts while the y halo is in flight): rows
    // xs_k0 .. xs_k0 + xs_nk - 1 of planes xs_z0 .. xs_z0 + xs_nz - 1; xs_nk = 0 means everything
    int xs_k0, xs_nk, xs_z0, xs_nz;
    // fused stability criterion (x kernel of stage 1, inviscid fast build): ICFL max of the
    // cells this kernel finishes, m_data_output.fpp:215-233; nullptr = off
    unsigned long long *stab_out;
    const double *rds_t[2];      // 1/ds of the two transverse directions (y, z)
    int weno_order;        // 5, 3 or 1
    int coef_uniform;      // 1: cuni[] holds the coefficients of every cell of this direction
    double cuni[kNumWenoCoef];   // uniform-grid WENO coefficients (COEF = 0 kernels)
    // per-variable plane pointers (filled by the launcher from rhs / q1 / qout + v*fstride): the
    // sweep kernels address a cell as plane pointer (constant bank) + 32-bit element offset, one
    // IMAD.WIDE per access instead of 64-bit pointer arithmetic per variable
    const double *q_v[kMaxE], *rhs_v[kMaxE], *q1_v[kMaxE];
    double *rhsw_v[kMaxE], *qout_v[kMaxE];
    // TMA descriptors (cuTensorMapEncodeTiled) of the stage state, the RHS accumulator and
    // q_cons_ts(1) seen as 4-D tensors (x, y, z, variable): one cp.async.bulk.tensor copy
    // brings a row of ALL variables into a ring slot, box = {kWX or kWY, 1, 1, E}
    TensorMap tm_q, tm_rhs, tm_q1;
    // interior-clipped maps (cell indices as coordinates, extents N+1: out-of-range columns are
    // zero-filled on load and dropped on store) of the RHS accumulator, q_cons_ts(1) and the
    // destination of this sweep (RHS accumulator, or the updated state when rk_mode != 0)
    TensorMap tm_rhs_i, tm_q1_i, tm_out_i;
    // viscous runs only (else nullptr): the sweep stores vel_src (nd planes) and Re_avg (2 planes)
    // of every face it solves, for k_visc (m_riemann_solvers.fpp:225-230,314-324)
    double *visc_face;
    // fast build, weno_Re_flux = F (visc_mode 2): the viscous source flux is computed inside the sweeps
    // from rdcc = 1/(s_cc(i+1) - s_cc(i)) of this direction and the per-cell gradient planes of k_vgrad
    int visc_mode;               // 0 inviscid, 1 face planes for k_visc, 2 in-sweep
    const double *rdcc, *vgrad;
    double Res[2][kMaxFluids];
    double iRes[2][kMaxFluids];  // 1/Res (fast build: the face stores 1/Re_avg, no divisions)
    int Re_idx[2][kMaxFluids], Re_size[2];
    double iRe_f[2][kMaxFluids]; // 1/Re(i) of every FLUID, 0 where that fluid has none (in-sweep viscous path)
};

// viscous source flux + its RHS contribution for one direction (m_riemann_solvers.fpp:683-902,
// m_rhs.fpp:591-604,639-652), and the velocity gradients of the weno_Re_flux branch
struct ViscArgs {
    GridDesc g;
    const double *prim;        // velocity (nd planes) + pressure
    const double *visc_face;   // vel_src (nd) + Re_avg (2) per face of direction dir
    double *grad;              // dq_prim_d{x,y}_qp: nd*nd planes, [dd*nd + v] (weno_Re_flux only)
    double *rhs;
    const double *coef[3];     // WENO coefficient tables per direction
    int clen[3], coef_lo[3];
    const double *cc[3];       // cell centres s_cc(-b:N+b)
    const double *ds[3];       // cell widths  ds(-b:N+b)
    double eps;
    int dir, nf, weno_Re_flux;
    int bc_beg, bc_end;
    int Re_size[2];
    // fast build: reciprocals of the centre-to-centre distances, rdcc(i) = 1/(s_cc(i+1) - s_cc(i)) for
    // i = -b .. N+b-1, and of the cell widths
    const double *rdcc[3], *rds[3];
    // last direction: the TVD-RK statement rk_mode (1..4, 0 = store the RHS) is applied here
    // instead of by a separate pass; E variables, q1 = q_cons_ts(1), qs = stage state
    int rk_mode, E;
    const double *q1, *qs;
    double *qout;
    double dt;
    int k_lo, k_hi;            // k_vgrad: cell rows k_lo .. k_hi (within -1 .. N+1)
};

struct BcArgs {
    GridDesc g;
    double *q;
    int dir, E, mom_normal;
    int bc_beg, bc_end;    // only sides with a negative (physical) code are filled
};

struct HaloArgs {
    GridDesc g;
    double *q;
    double *buf;           // the piece's own contiguous message: E runs of cnt doubles
    int dir, side, E;      // pack: side 0 = first b interior layers, 1 = last b;  unpack: ghosts at beg / end
    long long idx0, cnt;   // the piece: slab elements idx0 .. idx0 + cnt - 1 of every variable (cnt = 0: the whole slab)
};

struct PrimArgs {
    GridDesc g;
    const double *q;
    double *prim;
    double gammas[kMaxFluids], pi_infs[kMaxFluids];
};

struct StabArgs {
    GridDesc g;
    const double *q, *prim;
    const double *ds[3];   // ds(-b:N+b) per direction
    double gammas[kMaxFluids], pi_infs[kMaxFluids];
    double Res[2][kMaxFluids];
    int Re_idx[2][kMaxFluids], Re_size[2];
    double dt;
    unsigned long long *out;   // [0] icfl max, [1] vcfl max, [2] Rc min  as ordered bit patterns
};

// initial condition on the device (patches.cu): patch_icpp(i) as pre_process reads it
// (src/pre_process/m_global_parameters.fpp:196-230); layout = mfc_b200_patch_t of the C ABI
constexpr int kMaxPatches = 10;    // num_patches_max, src/pre_process/m_global_parameters.fpp
struct PatchDesc {
    int geometry, smoothen, smooth_patch_id;
    int alter_patch[kMaxPatches + 1];
    double x_centroid, y_centroid, z_centroid, length_x, length_y, length_z, radius;
    double radii[3], normal[3], epsilon, smooth_coeff;
    double vel[3], pres, alpha_rho[kMaxFluids], alpha[kMaxFluids];
};
struct PatchArgs {
    GridDesc g;
    double *q;                 // state planes (conservative variables are written to the interior)
    const double *cc[3];       // pre_process cell centres of this rank's interior cells, N_d + 1 doubles
    const double *cb[3];       // their right boundaries s_cb(0:N_d) (analytical patches 7, 15), or nullptr
    const PatchDesc *patches;  // device copy, num_patches entries in patch order
    int num_patches;
    double ds_min;             // min(dx, dy[, dz]) over the GLOBAL grid (s_mpi_reduce_min, m_start_up.fpp:720)
    double gammas[kMaxFluids], pi_infs[kMaxFluids];
};
int launch_patches(int nf, int nd, const PatchArgs &a, cudaStream_t st);

// number of (transverse x layer) elements of one variable in a ghost slab of direction dir:
// earlier directions ghosted, later ones interior only (m_rhs.fpp:696-697 vs :811-812)
MFC_HD long long slab_count(const GridDesc &g, int dir) {
    long long n0, n1;
    if (dir == 0) { n0 = g.N[1] + 1; n1 = g.N[2] + 1; }
    else if (dir == 1) { n0 = g.N[0] + 1 + 2*g.b; n1 = g.N[2] + 1; }
    else { n0 = g.N[0] + 1 + 2*g.b; n1 = g.N[1] + 1 + 2*g.b; }
    return n0*n1*g.b;
}

// every launcher returns the number of kernels it launched (0 = unsupported combination)
struct Launchers {
    int (*sweep)(int nf, int nd, int dir, const SweepArgs &, cudaStream_t);
    int (*prim)(int nf, int nd, const PrimArgs &, cudaStream_t);
    int (*bc)(const BcArgs &, cudaStream_t);
    int (*halo_pack)(const HaloArgs &, cudaStream_t);
    int (*halo_unpack)(const HaloArgs &, cudaStream_t);
    int (*stability)(int nf, int nd, const StabArgs &, cudaStream_t);
    int (*visc_grad)(int nd, const ViscArgs &, cudaStream_t);
    int (*visc)(int nd, const ViscArgs &, cudaStream_t);
    int (*vgrad)(int nd, const ViscArgs &, cudaStream_t);
};
const Launchers &launchers_fast();
const Launchers &launchers_strict();

}  // namespace mfc
