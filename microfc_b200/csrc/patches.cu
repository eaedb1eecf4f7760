// patches.cu -- initial condition on the device: the reference's pre_process
// lays the patches patch_icpp(1..num_patches) over the grid one after the other
// (src/pre_process/m_initial_condition.fpp:42-113), every cell independently of every other
// cell.  One thread therefore owns one cell and walks the patch list in order, holding the
// cell's primitive variables and its patch_id_fp entry in registers, and finally stores the
// CONSERVATIVE variables (m_variables_conversion.fpp:385-443) straight into the padded state
// planes -- 512^3 x 8 variables never exist on the host.
//
// Built with -fmad=false: every expression is evaluated in the order the Fortran statements
// are written.  tanh / exp come from the CUDA math library (<= 2 ulp from glibc / numpy), so cells in
// a smoothed patch boundary or an analytical patch (geometry 7, 15) may differ from a host
// pre_process in the last bits; hard-edged patches are bit-identical.  Citations relative to /root/reference/src/pre_process.
#include <cuda_runtime.h>
#include "args.hpp"

namespace mfc {

// geometry test + smoothing function of one patch at one cell
__device__ __forceinline__ void patch_geometry(const PatchDesc &pt, double X, double Y, double Z, double dmin,
                                               bool &inside, bool &smoothable, double &eta) {
    eta = 1.0;
    smoothable = false;
    inside = false;
    switch (pt.geometry) {
    case 6: {                                        // s_isentropic_vortex, m_create_patches.fpp:379-419: a hard circle
        const double dx = X - pt.x_centroid, dy = Y - pt.y_centroid;
        inside = dx*dx + dy*dy <= pt.radius*pt.radius;
        break;
    }
    case 15:                                         // s_1D_analytical, :424-473 (+ the pressure bump, see the caller)
    case 1: {                                        // s_line_segment, m_create_patches.fpp:47-88
        const double xb = pt.x_centroid - 0.5*pt.length_x, xe = pt.x_centroid + 0.5*pt.length_x;
        inside = xb <= X && xe >= X;
        break;
    }
    case 2: case 10: {                               // s_circle, :96-146 (10: z-invariant cylinder, extension)
        smoothable = true;
        const double dx = X - pt.x_centroid, dy = Y - pt.y_centroid;
        const double r2 = dx*dx + dy*dy;
        if (pt.smoothen) eta = tanh(pt.smooth_coeff/dmin*(sqrt(r2) - pt.radius))*(-0.5) + 0.5;
        inside = r2 <= pt.radius*pt.radius;
        break;
    }
    case 18: {                                       // s_varcircle, :148-192
        const double dx = X - pt.x_centroid, dy = Y - pt.y_centroid;
        const double myr = sqrt(dx*dx + dy*dy);
        inside = myr <= pt.radius + pt.epsilon/2.0 && myr >= pt.radius - pt.epsilon/2.0;
        break;
    }
    case 5: {                                        // s_ellipse, :200-251
        smoothable = true;
        const double ex = (X - pt.x_centroid)/pt.radii[0], ey = (Y - pt.y_centroid)/pt.radii[1];
        const double r2 = ex*ex + ey*ey;
        if (pt.smoothen) eta = tanh(pt.smooth_coeff/dmin*(sqrt(r2) - 1.0))*(-0.5) + 0.5;
        inside = r2 <= 1.0;
        break;
    }
    case 7:                                          // s_2D_analytical, :479-534 (+ the pressure bump, see the caller)
    case 3: {                                        // s_rectangle, :262-313
        const double xb = pt.x_centroid - 0.5*pt.length_x, xe = pt.x_centroid + 0.5*pt.length_x;
        const double yb = pt.y_centroid - 0.5*pt.length_y, ye = pt.y_centroid + 0.5*pt.length_y;
        inside = xb <= X && xe >= X && yb <= Y && ye >= Y;
        break;
    }
    case 4: {                                        // s_sweep_line, :324-370
        smoothable = true;
        const double a = pt.normal[0], b = pt.normal[1];
        const double c = -a*pt.x_centroid - b*pt.y_centroid;
        const double lin = a*X + b*Y + c;
        if (pt.smoothen) eta = 5e-1 + 5e-1*tanh(pt.smooth_coeff/dmin*lin/sqrt(a*a + b*b));
        inside = lin >= 0.0;
        break;
    }
    case 8: {                                        // EXTENSION: sphere
        smoothable = true;
        const double dx = X - pt.x_centroid, dy = Y - pt.y_centroid, dz = Z - pt.z_centroid;
        const double r2 = dx*dx + dy*dy + dz*dz;
        if (pt.smoothen) eta = tanh(pt.smooth_coeff/dmin*(sqrt(r2) - pt.radius))*(-0.5) + 0.5;
        inside = r2 <= pt.radius*pt.radius;
        break;
    }
    case 9: {                                        // EXTENSION: cuboid
        const double xb = pt.x_centroid - 0.5*pt.length_x, xe = pt.x_centroid + 0.5*pt.length_x;
        const double yb = pt.y_centroid - 0.5*pt.length_y, ye = pt.y_centroid + 0.5*pt.length_y;
        const double zb = pt.z_centroid - 0.5*pt.length_z, ze = pt.z_centroid + 0.5*pt.length_z;
        inside = xb <= X && xe >= X && yb <= Y && ye >= Y && zb <= Z && ze >= Z;
        break;
    }
    default: break;                                  // rejected on the host (MFC_B200_EUNSUPPORTED)
    }
}

template <int NF, int ND>
__global__ void __launch_bounds__(128) k_patches(const __grid_constant__ PatchArgs a) {
    constexpr int E = 2*NF + ND + 1, MOM = NF, EN = NF + ND, ADV = NF + ND + 1;
    const GridDesc &g = a.g;
    const int j = blockIdx.y*blockDim.x + threadIdx.x;
    if (j > g.N[0]) return;
    const int k = blockIdx.x % (g.N[1] + 1), l = blockIdx.x/(g.N[1] + 1);   // rows in gridDim.x: no 65535 cap
    const double X = a.cc[0][j], Y = ND > 1 ? a.cc[1][k] : 0.0, Z = ND > 2 ? a.cc[2][l] : 0.0;
    double q[E];
#pragma unroll
    for (int v = 0; v < E; v++) q[v] = 0.0;
    int patch_id = 0;                                // patch_id_fp, m_assign_patches.fpp:200
    for (int i = 0; i < a.num_patches; i++) {        // m_initial_condition.fpp:74-105, in order
        const PatchDesc &pt = a.patches[i];
        bool inside, smoothable;
        double eta;
        patch_geometry(pt, X, Y, Z, a.ds_min, inside, smoothable, eta);
        // "covers the cell AND may overwrite what is there, OR the cell belongs to the patch this
        // one is smeared against" (m_create_patches.fpp:131-137)
        bool mask = inside && pt.alter_patch[patch_id] != 0;
        if (smoothable) mask = mask || patch_id == pt.smooth_patch_id;
        if (!mask) continue;
        // s_assign_patch_species_primitive_variables, m_assign_patches.fpp:54-165
        const double ome = 1.0 - eta;
#pragma unroll
        for (int f = 0; f < NF; f++) {
            q[f] = eta*pt.alpha_rho[f] + ome*q[f];                   // :132-136
            q[ADV + f] = eta*pt.alpha[f] + ome*q[ADV + f];           // :137-141
        }
#pragma unroll
        for (int d = 0; d < ND; d++) q[MOM + d] = eta*pt.vel[d] + ome*q[MOM + d];   // :149-153
        q[EN] = eta*pt.pres + ome*q[EN];                             // :156-158
        if (pt.geometry == 15 || pt.geometry == 7) {
            // the analytical patches multiply the pressure by a Gaussian bump evaluated at the RIGHT
            // cell boundaries x_cb(i), y_cb(j) (m_create_patches.fpp:466-467, :527-528)
            const double tx = a.cb[0][j] - pt.x_centroid;
            double r2 = tx*tx;
            if (pt.geometry == 7) { const double ty = a.cb[ND > 1 ? 1 : 0][k] - pt.y_centroid; r2 = r2 + ty*ty; }
            q[EN] = q[EN]*(1.0 + 0.2*exp(-1.0*r2/(2.0*0.005)));
        }
        if (ome < 1e-16) patch_id = i + 1;                           // :163
    }
    // s_convert_primitive_to_conservative_variables, src/common/m_variables_conversion.fpp:385-443
    double rho = 0.0, gamma = 0.0, pi_inf = 0.0;
#pragma unroll
    for (int f = 0; f < NF; f++) {
        rho = rho + q[f];
        gamma = gamma + q[ADV + f]*a.gammas[f];
        pi_inf = pi_inf + q[ADV + f]*a.pi_infs[f];
    }
    double dyn = 0.0;
#pragma unroll
    for (int d = 0; d < ND; d++) {
        const double mom = rho*q[MOM + d];                           // :426-430
        dyn = dyn + mom*q[MOM + d]/2.0;
        q[MOM + d] = mom;
    }
    q[EN] = gamma*q[EN] + dyn + pi_inf;                              // :434-435
    const long long cell = g.at(j, k, l);
#pragma unroll
    for (int v = 0; v < E; v++) a.q[(long long)v*g.fstride + cell] = q[v];
}

int launch_patches(int nf, int nd, const PatchArgs &a, cudaStream_t st) {
    const GridDesc &g = a.g;
    dim3 grid((g.N[1] + 1)*(g.N[2] + 1), (g.N[0] + 128)/128, 1);
    switch (nf*10 + nd) {
    case 11: k_patches<1, 1><<<grid, 128, 0, st>>>(a); return 1;
    case 12: k_patches<1, 2><<<grid, 128, 0, st>>>(a); return 1;
    case 13: k_patches<1, 3><<<grid, 128, 0, st>>>(a); return 1;
    case 21: k_patches<2, 1><<<grid, 128, 0, st>>>(a); return 1;
    case 22: k_patches<2, 2><<<grid, 128, 0, st>>>(a); return 1;
    case 23: k_patches<2, 3><<<grid, 128, 0, st>>>(a); return 1;
    case 31: k_patches<3, 1><<<grid, 128, 0, st>>>(a); return 1;
    case 32: k_patches<3, 2><<<grid, 128, 0, st>>>(a); return 1;
    case 33: k_patches<3, 3><<<grid, 128, 0, st>>>(a); return 1;
    default: return 0;
    }
}

}  // namespace mfc
