// layout.hpp -- how the reference's structure-of-arrays fields live in HBM, shared by host
// and device code.
//
// The reference allocates every scalar field separately as sf(-b:m+b, -b:n+b), x fastest
// (m_time_steppers.fpp:84-85).  Here one "state" is a single allocation of E padded planes:
//
//   element (j,k,l) of field v  ->  base + v*fstride + (j+xoff) + pitch*((k+yoff) + ey*(l+zoff))
//
//   xoff  = 16 doubles: interior column 0 starts on a 128-byte line, the b ghost columns sit
//           just before it
//   pitch = row length rounded up to 16 doubles (128 B) so every row starts line-aligned
//   yoff/zoff = b for active directions, 0 otherwise; ey/ez = N+1+2b or 1
//
// Sweeps read along x, y or z directly from this layout (coalesced along x in every case);
// there are no per-direction transposed copies (the reference's v_rs_ws_{x,y} / flux_rs*
// buffers, m_weno.fpp:570-595, m_riemann_solvers.fpp:923-957).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define MFC_HD __host__ __device__ __forceinline__
#else
#define MFC_HD inline
#endif

namespace mfc {

constexpr int kXoff = 16;
constexpr int kMaxFluids = 4;
constexpr int kMaxE = 2*kMaxFluids + 3 + 1;
constexpr int kWX = 34;        // columns per box and variable of the x kernel: 32 + the even-start rounding
constexpr int kWY = 32;        // columns per warp in the y/z march
constexpr int kNumWenoCoef = 27;   // 6 poly_L + 6 poly_R + 3 d_L + 3 d_R + 9 beta per cell

struct GridDesc {
    int N[3];          // local m, n, p  (last cell index per direction)
    int nd;            // num_dims
    int b;             // buff_size
    int pitch;         // doubles per x row
    int ey, ez;        // allocated rows / planes
    int yoff, zoff;    // ghost offsets of the y / z directions
    long long fstride; // doubles per field plane-set
    long long sy, sz;  // element strides of the y and z directions (pitch, pitch*ey)

    MFC_HD long long at(int j, int k, int l) const {
        return (long long)(j + kXoff) + (long long)pitch*((long long)(k + yoff) + (long long)ey*(long long)(l + zoff));
    }
    MFC_HD long long stride(int dir) const { return dir == 0 ? 1 : (dir == 1 ? sy : sz); }
};

inline GridDesc make_grid(int m, int n, int p, int nd, int b) {
    GridDesc g{};
    g.N[0] = m; g.N[1] = nd > 1 ? n : 0; g.N[2] = nd > 2 ? p : 0;
    g.nd = nd; g.b = b;
    int row = kXoff + (m + 1) + b;
    g.pitch = (row + 15)/16*16;
    g.yoff = nd > 1 ? b : 0; g.zoff = nd > 2 ? b : 0;
    g.ey = nd > 1 ? g.N[1] + 1 + 2*b : 1;
    g.ez = nd > 2 ? g.N[2] + 1 + 2*b : 1;
    g.sy = g.pitch; g.sz = (long long)g.pitch*g.ey;
    g.fstride = (long long)g.pitch*g.ey*g.ez;
    return g;
}

}  // namespace mfc
