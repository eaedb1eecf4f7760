// mfc_api.cu -- the C ABI of libmfc_b200.so (include/mfc_b200.h): device state, stage
// orchestration (s_compute_rhs, m_rhs.fpp:405-679, and the TVD-RK steppers,
// m_time_steppers.fpp:129-362), NCCL halo exchange (replacing m_mpi_proxy.fpp:468-979) and
// the stability reduction (m_data_output.fpp:249-274, m_mpi_common.fpp:135-171).
//
// There is NO CPU fallback: without a usable CUDA device every entry fails with
// MFC_B200_ENODEVICE.
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mfc_b200.h"
#include "args.hpp"
#include "weno_coefficients.hpp"

using namespace mfc;

namespace {

enum KernelClass { KC_BC = 0, KC_PRIM, KC_SWEEP_X, KC_SWEEP_Y, KC_SWEEP_Z, KC_STAB, KC_PACK, KC_UNPACK, KC_VISC, KC_PATCH, KC_COUNT };
const char *kKernelNames[KC_COUNT] = {"k_bc", "k_prim", "k_xstream", "k_march3<y>", "k_march3<z>",
                                      "k_stability", "k_halo_pack", "k_halo_unpack", "k_visc", "k_patches"};

// NCCL is resolved at run time so the library loads (and every symbol is exported) on hosts
// without it; only mfc_b200_comm_init needs it.  In a process that already imported torch
// the loader hands back torch's bundled libnccl.so.2.
struct NcclApi {
    void *h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool load(std::string &err) {
        if (h) return true;
        h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) { err = std::string("cannot load libnccl.so.2: ") + dlerror(); return false; }
#define MFC_SYM(name) *(void **)(&name) = dlsym(h, "nccl" #name); if (!name) { err = "missing NCCL symbol nccl" #name; return false; }
        MFC_SYM(GetUniqueId) MFC_SYM(CommInitRank) MFC_SYM(CommDestroy) MFC_SYM(Send) MFC_SYM(Recv)
        MFC_SYM(AllReduce) MFC_SYM(GroupStart) MFC_SYM(GroupEnd) MFC_SYM(GetErrorString)
#undef MFC_SYM
        return true;
    }
};

struct ProfRec { int kc; cudaEvent_t a, b; };

struct Sim {
    bool inited = false, uploaded = false;
    mfc_b200_params_t p{};
    GridDesc g{};
    int nf = 0, nd = 0, E = 0, b = 0;
    bool viscous = false;
    bool visc_fused = false;           // fast build, weno_Re_flux = F: viscous fluxes inside the sweeps
    const Launchers *L = nullptr;
    cudaStream_t st = nullptr;
    // multi-rank: the halo exchange runs on its own stream so that the y / z exchanges overlap
    // with the x sweep (inviscid runs: the three directions are independent, see ghosts_begin)
    cudaStream_t cs = nullptr;
    cudaEvent_t ev_q = nullptr, ev_halo[3] = {nullptr, nullptr, nullptr};
    bool halo_pending[3] = {false, false, false};
    bool overlap = false, xsplit_on = false;
    // the x halo travels in xpieces consecutive pieces (plane ranges in 3-D, row ranges in 2-D), each
    // with its own event, and the x sweep follows piece by piece: only the first piece is exposed
    static constexpr int kMaxPieces = 8;
    int xpieces = 1;
    cudaEvent_t ev_xp[kMaxPieces] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    // pieces 1.. of the x sweep run on streams of their own, so that the tail of one piece is backfilled by
    // the CTAs of the next (on one stream every piece would drain before the next starts)
    cudaStream_t xstream[kMaxPieces] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_xdone[kMaxPieces] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_pre = nullptr;
    bool xp_pending = false;
    bool bc_map = true;                // y / z ghost rows of physical boundaries synthesised by the march (no k_bc)
    bool visc_overlap = false;         // viscous fused path: y halo in flight under the x sweep of the inner rows
    double *state[3] = {nullptr, nullptr, nullptr};
    int cur = 0;                       // which buffer holds q_cons_ts(1)
    const double *last_q = nullptr;    // state of the most recent RHS evaluation (what q_prim_vf reflects)
    double *prim = nullptr, *rhs = nullptr, *snap = nullptr;
    // TMA descriptors of the three state buffers and the RHS accumulator, [0]: box kWX wide (x
    // sweep), [1]: box kWY wide (y/z march), [2]: interior cells only, box kWY wide (operand loads
    // and output stores of the march)
    TensorMap tm_state[3][3], tm_rhs[3];
    double *coef[3] = {nullptr, nullptr, nullptr};
    int clen[3] = {0, 0, 0}, coef_lo[3] = {0, 0, 0};
    int coef_uniform[3] = {0, 0, 0};   // every cell of the direction has the same 27 coefficients (to 1e-12)
    double cuni[3][kNumWenoCoef];
    double *rds[3] = {nullptr, nullptr, nullptr}, *ds[3] = {nullptr, nullptr, nullptr}, *cc[3] = {nullptr, nullptr, nullptr};
    double *rdcc[3] = {nullptr, nullptr, nullptr};   // 1/(s_cc(i+1) - s_cc(i)), viscous fast build
    // viscous runs: vel_src + Re_avg per face (nd+2 planes), dq_prim_d (nd*nd planes, weno_Re_flux)
    double *visc_face = nullptr, *visc_grad = nullptr;
    double Res[2][kMaxFluids];
    int Re_idx[2][kMaxFluids], Re_size[2] = {0, 0};
    std::vector<double> h_coef[3];
    unsigned long long *stab_dev = nullptr, *stab_host = nullptr, *stab_init = nullptr;
    bool stab_pending = false;
    int bc[3][2];                      // effective codes: self-neighbours folded into periodic
    double *sendbuf[3][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
    double *recvbuf[3][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
    NcclApi nccl;
    ncclComm_t comm = nullptr;
    int64_t launches = 0;
    bool prof = false;
    std::vector<ProfRec> prof_recs;
    double prof_s[KC_COUNT] = {0};
    int64_t prof_n[KC_COUNT] = {0};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, sv0 = nullptr, sv1 = nullptr;
    // one RK step captured as a CUDA graph (single rank): a step of a small grid is a dozen
    // launches of a few microseconds each and is bound by launch latency, not by the kernels
    struct StepGraph {
        unsigned long long dt_bits; int stab, cur_before;          // key
        cudaGraphExec_t exec;
        int cur_after; const double *last_q_after; bool stab_pending_after; int64_t launches;
    };
    std::vector<StepGraph> graphs;
    bool graph_on = false;
    std::string err;
};
Sim S;

int fail(int code, const std::string &msg) { S.err = msg; return code; }

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(e_ == cudaErrorMemoryAllocation ? MFC_B200_ENOMEM : MFC_B200_ECUDA,        \
                        std::string(#call) + ": " + cudaGetErrorString(e_));                       \
    } while (0)
#define NK(call)                                                                                   \
    do {                                                                                           \
        ncclResult_t r_ = (call);                                                                  \
        if (r_ != ncclSuccess) return fail(MFC_B200_ENCCL, std::string(#call) + ": " + S.nccl.GetErrorString(r_)); \
    } while (0)

// launch bookkeeping: count, optional per-class CUDA-event timing on the launching stream
struct Scope {
    int kc; ProfRec r{}; cudaStream_t s;
    explicit Scope(int kc_, cudaStream_t s_ = nullptr) : kc(kc_), s(s_ ? s_ : S.st) {
        if (S.prof) { r.kc = kc; cudaEventCreate(&r.a); cudaEventCreate(&r.b); cudaEventRecord(r.a, s); }
    }
    void done(int n) {
        S.launches += n;
        if (S.prof) { cudaEventRecord(r.b, s); S.prof_recs.push_back(r); }
    }
};

void prof_collect() {
    if (S.prof_recs.empty()) return;
    cudaStreamSynchronize(S.st);
    if (S.cs) cudaStreamSynchronize(S.cs);
    for (auto &r : S.prof_recs) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.a, r.b);
        S.prof_s[r.kc] += ms*1e-3; S.prof_n[r.kc] += 1;
        cudaEventDestroy(r.a); cudaEventDestroy(r.b);
    }
    S.prof_recs.clear();
}

void free_all() {
    auto fr = [](double *&p) { if (p) cudaFree(p); p = nullptr; };
    for (auto &s : S.state) fr(s);
    fr(S.prim); fr(S.rhs); fr(S.snap); fr(S.visc_face); fr(S.visc_grad);
    for (int d = 0; d < 3; d++) {
        fr(S.coef[d]); fr(S.rds[d]); fr(S.ds[d]); fr(S.cc[d]); fr(S.rdcc[d]);
        for (int s = 0; s < 2; s++) { fr(S.sendbuf[d][s]); fr(S.recvbuf[d][s]); }
    }
    if (S.stab_dev) { cudaFree(S.stab_dev); S.stab_dev = nullptr; }
    if (S.stab_host) { cudaFreeHost(S.stab_host); S.stab_host = nullptr; }
    if (S.stab_init) { cudaFreeHost(S.stab_init); S.stab_init = nullptr; }
    if (S.ev0) { cudaEventDestroy(S.ev0); S.ev0 = nullptr; }
    if (S.ev1) { cudaEventDestroy(S.ev1); S.ev1 = nullptr; }
    if (S.sv0) { cudaEventDestroy(S.sv0); S.sv0 = nullptr; }
    if (S.sv1) { cudaEventDestroy(S.sv1); S.sv1 = nullptr; }
    if (S.comm && S.nccl.CommDestroy) { S.nccl.CommDestroy(S.comm); S.comm = nullptr; }
    if (S.ev_q) { cudaEventDestroy(S.ev_q); S.ev_q = nullptr; }
    for (auto &e : S.ev_halo) if (e) { cudaEventDestroy(e); e = nullptr; }
    for (auto &e : S.ev_xp) if (e) { cudaEventDestroy(e); e = nullptr; }
    for (auto &e : S.ev_xdone) if (e) { cudaEventDestroy(e); e = nullptr; }
    for (auto &x : S.xstream) if (x) { cudaStreamDestroy(x); x = nullptr; }
    if (S.ev_pre) { cudaEventDestroy(S.ev_pre); S.ev_pre = nullptr; }
    if (S.cs) { cudaStreamDestroy(S.cs); S.cs = nullptr; }
    S.overlap = false;
    if (S.st) { cudaStreamDestroy(S.st); S.st = nullptr; }
    for (auto &r : S.prof_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    S.prof_recs.clear();
    for (auto &gr : S.graphs) cudaGraphExecDestroy(gr.exec);
    S.graphs.clear();
    S.inited = S.uploaded = false;
}

size_t field_bytes() { return (size_t)S.g.fstride*sizeof(double); }

// E padded planes seen as the 4-D tensor (x, y, z, variable); box = one row of boxw columns of
// every variable, which lands in shared memory as E consecutive runs of boxw doubles
// interior = true: only the interior cells (coordinates = cell indices, extents N+1), so that a box
// hanging over the last interior column is zero-filled on load and clipped on store
int make_tmap(TensorMap &out, double *base, int boxw, bool interior = false) {
    typedef CUresult (*Encode)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                               const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static Encode encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn)
            return fail(MFC_B200_ENODEVICE, "cuTensorMapEncodeTiled is not available from this driver");
        encode = (Encode)fn;
    }
    const GridDesc &g = S.g;
    cuuint64_t dims[4] = {(cuuint64_t)g.pitch, (cuuint64_t)g.ey, (cuuint64_t)g.ez, (cuuint64_t)S.E};
    if (interior) {
        base += g.at(0, 0, 0);                           // 128-byte aligned: kXoff and the pitch are multiples of 16 doubles
        dims[0] = (cuuint64_t)g.N[0] + 1; dims[1] = (cuuint64_t)g.N[1] + 1; dims[2] = (cuuint64_t)g.N[2] + 1;
    }
    const cuuint64_t strides[3] = {(cuuint64_t)g.sy*8, (cuuint64_t)g.sz*8, (cuuint64_t)g.fstride*8};
    const cuuint32_t box[4] = {(cuuint32_t)boxw, 1, 1, (cuuint32_t)S.E};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUtensorMap m;
    const CUresult r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(MFC_B200_ECUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
    static_assert(sizeof(CUtensorMap) == sizeof(TensorMap), "CUtensorMap is 128 bytes");
    std::memcpy(&out, &m, sizeof(m));
    return 0;
}
const TensorMap *state_tmap(const double *q, int which) {
    for (int i = 0; i < 3; i++)
        if (q == S.state[i]) return &S.tm_state[i][which];
    return nullptr;
}

// ---- ghost cells: physical BCs (k_bc) and processor boundaries (pack / NCCL / unpack) --------
// processor boundaries of direction d: pack, ncclSend/ncclRecv with the +-d neighbours, unpack
// (m_mpi_proxy.fpp:468-979), all on stream st
int exchange_dir(double *q, int d, cudaStream_t st, long long idx0 = 0, long long cnt = 0) {
    if (!S.comm) return fail(MFC_B200_ESTATE, "processor boundary present but mfc_b200_comm_init was not called");
    if (cnt <= 0) { idx0 = 0; cnt = slab_count(S.g, d); }
    const long long n = cnt*S.E;                         // doubles per message; piece p lives at E*idx0 of the buffers
    for (int s = 0; s < 2; s++) {
        if (S.bc[d][s] < 0) continue;
        HaloArgs h{S.g, q, S.sendbuf[d][s] + S.E*idx0, d, s, S.E, idx0, cnt};
        Scope sc(KC_PACK, st); sc.done(S.L->halo_pack(h, st));
    }
    // sends: [to beg: my first layers] [to end: my last layers];  receives in the opposite
    // order so that two exchanges with the SAME peer (2 ranks, periodic) pair up correctly.
    NK(S.nccl.GroupStart());
    for (int s = 0; s < 2; s++)
        if (S.bc[d][s] >= 0) NK(S.nccl.Send(S.sendbuf[d][s] + S.E*idx0, (size_t)n, ncclDouble, S.bc[d][s], S.comm, st));
    for (int s = 1; s >= 0; s--)
        if (S.bc[d][s] >= 0) NK(S.nccl.Recv(S.recvbuf[d][s] + S.E*idx0, (size_t)n, ncclDouble, S.bc[d][s], S.comm, st));
    NK(S.nccl.GroupEnd());
    for (int s = 0; s < 2; s++) {
        if (S.bc[d][s] < 0) continue;
        HaloArgs h{S.g, q, S.recvbuf[d][s] + S.E*idx0, d, s, S.E, idx0, cnt};
        Scope sc(KC_UNPACK, st); sc.done(S.L->halo_unpack(h, st));
    }
    return 0;
}
// piece p of the x slab: a range of planes (3-D) or rows (2-D); the slab index runs layer fastest,
// then row, then plane (bc_decode), so a piece is a contiguous index range
void xpiece_range(int p, int &k0, int &nk, int &z0, int &nz, long long &idx0, long long &cnt) {
    const GridDesc &g = S.g;
    const int ny = g.N[1] + 1, nzz = g.N[2] + 1, P = S.xpieces;
    if (S.nd == 3) {
        z0 = (int)((long long)nzz*p/P); nz = (int)((long long)nzz*(p + 1)/P) - z0;
        k0 = 0; nk = ny;
        idx0 = (long long)g.b*ny*z0; cnt = (long long)g.b*ny*nz;
    } else {
        k0 = (int)((long long)ny*p/P); nk = (int)((long long)ny*(p + 1)/P) - k0;
        z0 = 0; nz = 1;
        idx0 = (long long)g.b*k0; cnt = (long long)g.b*nk;
    }
}
// mapped = this call is on behalf of an inviscid sweep along d: for d >= 1 the march synthesises the
// ghost rows of physical boundaries itself (SweepArgs::map_beg / map_end) and no kernel is launched
int physical_bc_dir(double *q, int d, bool mapped = false) {
    if (S.bc[d][0] >= 0 && S.bc[d][1] >= 0) return 0;
    if (mapped && d >= 1 && S.bc_map && !S.viscous) return 0;
    BcArgs a{S.g, q, d, S.E, S.nf + d, S.bc[d][0], S.bc[d][1]};
    Scope sc(KC_BC); sc.done(S.L->bc(a, S.st));
    return 0;
}

// Ghost fill of the stage state, m_rhs.fpp:686-908.
//   sequential (single rank, viscous): one direction after the other on the compute
//     stream, later directions covering the ghosts of the earlier ones (corners), like the reference.
//   overlapped (multi-rank inviscid): a sweep along d reads ghosts of direction d only,
//     at interior transverse indices -- corner ghosts are never read -- so the three exchanges
//     are independent.  They are enqueued on the communication stream in the order x, y, z; the
//     sweep along d waits for exchange d only (ghosts_ready), i.e. the y and z exchanges run
//     under the x sweep.  Interior results are identical to the sequential order bit for bit.
int ghosts_begin(double *q, bool mapped = false) {
    for (int d = 0; d < 3; d++) S.halo_pending[d] = false;
    if (!S.overlap) {
        for (int d = 0; d < S.nd; d++) {
            int rc;
            if ((S.bc[d][0] >= 0 || S.bc[d][1] >= 0) && (rc = exchange_dir(q, d, S.st))) return rc;
            if ((rc = physical_bc_dir(q, d, mapped))) return rc;
        }
        return 0;
    }
    CK(cudaEventRecord(S.ev_q, S.st));
    CK(cudaStreamWaitEvent(S.cs, S.ev_q, 0));
    S.xp_pending = false;
    for (int d = 0; d < S.nd; d++) {
        if (S.bc[d][0] < 0 && S.bc[d][1] < 0) continue;
        int rc;
        if (d == 0 && S.xpieces > 1 && !S.xsplit_on) {
            for (int p = 0; p < S.xpieces; p++) {
                int k0, nk, z0, nz; long long idx0, cnt;
                xpiece_range(p, k0, nk, z0, nz, idx0, cnt);
                if ((rc = exchange_dir(q, 0, S.cs, idx0, cnt))) return rc;
                CK(cudaEventRecord(S.ev_xp[p], S.cs));
            }
            S.xp_pending = true;
        } else if ((rc = exchange_dir(q, d, S.cs))) return rc;
        CK(cudaEventRecord(S.ev_halo[d], S.cs));
        S.halo_pending[d] = true;
    }
    return 0;
}
// ghosts of direction d complete on the compute stream (overlapped mode; no-op otherwise)
int ghosts_ready(double *q, int d, bool mapped = false) {
    if (!S.overlap) return 0;
    if (S.halo_pending[d]) { CK(cudaStreamWaitEvent(S.st, S.ev_halo[d], 0)); S.halo_pending[d] = false; }
    return physical_bc_dir(q, d, mapped);
}
int fill_ghosts(double *q) {
    int rc;
    if ((rc = ghosts_begin(q))) return rc;
    for (int d = 0; d < S.nd; d++)
        if ((rc = ghosts_ready(q, d))) return rc;
    return 0;
}

int run_prim(const double *q) {
    PrimArgs a{};
    a.g = S.g; a.q = q; a.prim = S.prim;
    for (int i = 0; i < kMaxFluids; i++) { a.gammas[i] = S.p.gammas[i]; a.pi_infs[i] = S.p.pi_infs[i]; }
    Scope sc(KC_PRIM);
    const int n = S.L->prim(S.nf, S.nd, a, S.st);
    if (!n) return fail(MFC_B200_EUNSUPPORTED, "no kernel instantiated for this (num_fluids, num_dims)");
    sc.done(n);
    return 0;
}

// stability criteria (m_data_output.fpp:197-274): everything is ENQUEUED on the stream; the
// host reads the result in stab_fetch() after the step's single synchronisation.
int stab_reset() {
    const double inf = INFINITY;
    unsigned long long init[3] = {0ull, 0ull, 0ull};
    std::memcpy(&init[2], &inf, sizeof(double));
    std::memcpy(S.stab_init, init, sizeof(init));
    CK(cudaMemcpyAsync(S.stab_dev, S.stab_init, sizeof(init), cudaMemcpyHostToDevice, S.st));
    return 0;
}
int stab_reduce_and_copy() {
    if (S.comm) {   // m_mpi_common.fpp:155-165: MAX / MIN over ranks (every rank gets the result)
        NK(S.nccl.AllReduce(S.stab_dev, S.stab_dev, 2, ncclUint64, ncclMax, S.comm, S.st));
        NK(S.nccl.AllReduce(S.stab_dev + 2, S.stab_dev + 2, 1, ncclUint64, ncclMin, S.comm, S.st));
    }
    CK(cudaMemcpyAsync(S.stab_host, S.stab_dev, 3*sizeof(unsigned long long), cudaMemcpyDeviceToHost, S.st));
    S.stab_pending = true;
    return 0;
}
int run_stability(const double *q, double dt) {
    int rc;
    // the sweeps convert in shared memory; q_prim_vf is materialised only for this diagnostic
    if ((rc = run_prim(q))) return rc;
    if ((rc = stab_reset())) return rc;
    StabArgs a{};
    a.g = S.g; a.q = q; a.prim = S.prim; a.dt = dt; a.out = S.stab_dev;
    for (int d = 0; d < 3; d++) a.ds[d] = S.ds[d];
    for (int i = 0; i < kMaxFluids; i++) { a.gammas[i] = S.p.gammas[i]; a.pi_infs[i] = S.p.pi_infs[i]; }
    for (int i = 0; i < 2; i++) {
        a.Re_size[i] = S.Re_size[i];
        for (int q = 0; q < kMaxFluids; q++) { a.Res[i][q] = S.Res[i][q]; a.Re_idx[i][q] = S.Re_idx[i][q]; }
    }
    {
        Scope sc(KC_STAB); sc.done(S.L->stability(S.nf, S.nd, a, S.st));
    }
    return stab_reduce_and_copy();
}
// after the stream has been synchronised
void stab_fetch(double stab[3]) {
    if (!S.stab_pending) return;
    S.stab_pending = false;
    if (!stab) return;
    std::memcpy(&stab[0], &S.stab_host[0], sizeof(double));
    if (S.viscous) { std::memcpy(&stab[1], &S.stab_host[1], sizeof(double)); std::memcpy(&stab[2], &S.stab_host[2], sizeof(double)); }
}

// one s_compute_rhs on the stage state q (ghosts rebuilt in place), followed -- fused into the
// last sweep -- by the RK statement rk_mode writes into qout
int rhs_stage(double *q, int rk_mode, const double *q1, double *qout, double dt, int t_step,
              bool first_stage, bool want_stab) {
    int rc;
    const bool stop = first_stage && t_step == S.p.t_step_stop;
    // Reference quirk, kept: at t_step == t_step_stop s_compute_rhs returns (m_rhs.fpp:452)
    // BEFORE it refreshes q_prim_vf (:659-675), so the last run_time.inf row is computed from
    // the primitive variables of the previous RHS evaluation (m_time_steppers.fpp:288-290).
    if (stop && S.p.run_time_info && want_stab && S.last_q)
        if ((rc = run_stability(S.last_q, dt))) return rc;
    // m_rhs.fpp:435.  Anything that reads the whole ghosted box needs every direction complete.
    const bool need_all = stop || S.viscous;
    // viscous, in-sweep path, y neighbours: x ghosts first (whole, on the compute stream), then the y
    // exchange -- whose messages span the x ghosts and so carry the corners, m_mpi_proxy.fpp:736-739 --
    // goes to the communication stream and flies under k_vgrad / the x sweep of rows 1 .. n-1, which
    // read no y ghost; rows 0 and n and the y sweep follow its arrival
    const bool vo = S.visc_overlap && !stop && S.nd == 2 && (S.bc[1][0] >= 0 || S.bc[1][1] >= 0) && S.g.N[1] >= 8;
    if (vo) {
        if ((S.bc[0][0] >= 0 || S.bc[0][1] >= 0) && (rc = exchange_dir(q, 0, S.st))) return rc;
        if ((rc = physical_bc_dir(q, 0))) return rc;
        CK(cudaEventRecord(S.ev_q, S.st));
        CK(cudaStreamWaitEvent(S.cs, S.ev_q, 0));
        if ((rc = exchange_dir(q, 1, S.cs))) return rc;
        CK(cudaEventRecord(S.ev_halo[1], S.cs));
    } else if ((rc = need_all ? fill_ghosts(q) : ghosts_begin(q, true))) return rc;
    // :445-447 (fused into the sweeps; the viscous kernels read the velocity planes -- the in-sweep
    // viscous path only needs them for the cross-direction gradients, i.e. not in 1-D)
    if (S.viscous && (!S.visc_fused || stop) && (rc = run_prim(q))) return rc;
    if (stop) return 0;                                      // m_rhs.fpp:452, m_time_steppers.fpp:296
    S.last_q = q;
    // m_time_steppers.fpp:288-290.  Inviscid fast build: the ICFL maximum is taken by the x sweep
    // itself from the primitive variables it already holds (no extra pass over the state).
    const bool do_stab = first_stage && S.p.run_time_info && want_stab;
    const bool fuse_stab = do_stab && !S.p.strict_math && !S.viscous;
    if (do_stab && !fuse_stab) {
        if (!need_all)
            for (int d = 0; d < S.nd; d++)
                if ((rc = ghosts_ready(q, d))) return rc;
        if ((rc = run_stability(q, dt))) return rc;
    }
    if (fuse_stab && (rc = stab_reset())) return rc;
    ViscArgs va{};
    if (S.viscous) {                                         // :456-464 s_get_viscous
        va.g = S.g; va.prim = S.prim; va.visc_face = S.visc_face; va.grad = S.visc_grad; va.rhs = S.rhs;
        for (int d = 0; d < 3; d++) {
            va.coef[d] = S.coef[d]; va.clen[d] = S.clen[d]; va.coef_lo[d] = S.coef_lo[d]; va.cc[d] = S.cc[d]; va.ds[d] = S.ds[d];
            va.rdcc[d] = S.rdcc[d]; va.rds[d] = S.rds[d];
        }
        va.rk_mode = 0; va.E = S.E; va.q1 = q1; va.qs = q; va.qout = qout; va.dt = dt;
        va.eps = S.p.weno_eps; va.nf = S.nf; va.weno_Re_flux = S.p.weno_Re_flux;
        va.Re_size[0] = S.Re_size[0]; va.Re_size[1] = S.Re_size[1];
        if (S.p.weno_Re_flux) { Scope sc(KC_VISC); sc.done(S.L->visc_grad(S.nd, va, S.st)); }
        if (S.visc_fused && S.nd > 1) {
            va.k_lo = vo ? 1 : -1; va.k_hi = vo ? S.g.N[1] - 1 : S.g.N[1] + 1;
            Scope sc(KC_VISC); sc.done(S.L->vgrad(S.nd, va, S.st));
        }
    }
    for (int d = 0; d < S.nd; d++) {                         // :469
        SweepArgs a{};
        a.g = S.g; a.q = q; a.prim = S.prim; a.rhs = S.rhs; a.q1 = q1; a.qout = qout;
        a.coef = S.coef[d]; a.clen = S.clen[d]; a.coef_lo = S.coef_lo[d]; a.rds = S.rds[d];
        a.eps = S.p.weno_eps; a.dt = dt;
        for (int i = 0; i < kMaxFluids; i++) { a.gammas[i] = S.p.gammas[i]; a.pi_infs[i] = S.p.pi_infs[i]; }
        a.bc_beg = S.p.bc[2*d]; a.bc_end = S.p.bc[2*d + 1];
        a.first_dir = d == 0;
        a.rk_mode = (d == S.nd - 1 && (!S.viscous || S.visc_fused)) ? rk_mode : 0;
        a.visc_mode = S.viscous ? (S.visc_fused ? 2 : 1) : 0;
        a.visc_face = a.visc_mode == 1 ? S.visc_face : nullptr;
        a.rdcc = S.rdcc[d]; a.vgrad = S.visc_grad;
        for (int i = 0; i < 2; i++) {
            a.Re_size[i] = S.Re_size[i];
            for (int k = 0; k < kMaxFluids; k++) {
                a.Res[i][k] = S.Res[i][k]; a.Re_idx[i][k] = S.Re_idx[i][k];
                a.iRes[i][k] = k < S.Re_size[i] ? 1.0/S.Res[i][k] : 0.0;
            }
        }
        for (int i = 0; i < 2; i++)
            for (int k = 0; k < kMaxFluids; k++) a.iRe_f[i][k] = (k < S.nf && S.p.Re[k][i] > 0.0) ? 1.0/S.p.Re[k][i] : 0.0;
        a.coef_uniform = S.coef_uniform[d]; a.weno_order = S.p.weno_order;
        a.stab_out = (fuse_stab && d == 0) ? S.stab_dev : nullptr;
        a.rds_t[0] = S.rds[1]; a.rds_t[1] = S.rds[2];
        for (int i = 0; i < kNumWenoCoef; i++) a.cuni[i] = S.cuni[d][i];
        // multi-rank, x halo still in flight: sweep the tiles that read no x ghost column first
        // (optional, MFC_B200_XSPLIT=1: the two boundary strips cost 2 x 32 lane slots per row whatever
        // their width, ~12 % of a 512-cell row, against ~0.2 ms of exposed x exchange at 512^3)
        const bool split_x = d == 0 && S.overlap && S.xsplit_on && S.halo_pending[0] && S.nd >= 2;
        const bool pieces_x = d == 0 && (S.xp_pending || vo);
        const bool mapped = !need_all;
        a.map_beg = a.map_end = 0;
        if (mapped && d >= 1 && S.bc_map && !S.viscous) {
            a.map_beg = S.bc[d][0] < 0 ? S.bc[d][0] : 0;
            a.map_end = S.bc[d][1] < 0 ? S.bc[d][1] : 0;
        }
        if (!split_x && !pieces_x && !(vo && d == 1) && (rc = ghosts_ready(q, d, mapped))) return rc;
        {
            const TensorMap *tq = state_tmap(q, d == 0 ? 0 : 1), *t1 = state_tmap(q1, d == 0 ? 0 : 1);
            if (!tq || !t1) return fail(MFC_B200_ESTATE, "stage state is not one of the library's state buffers");
            a.tm_q = *tq; a.tm_q1 = *t1; a.tm_rhs = S.tm_rhs[d == 0 ? 0 : 1];
            // interior-clipped maps of the march kernels' operand rows and of this sweep's destination
            const TensorMap *t1i = state_tmap(q1, 2), *toi = a.rk_mode != 0 ? state_tmap(qout, 2) : &S.tm_rhs[2];
            if (!t1i || !toi) return fail(MFC_B200_ESTATE, "stage state is not one of the library's state buffers");
            a.tm_rhs_i = S.tm_rhs[2]; a.tm_q1_i = *t1i; a.tm_out_i = *toi;
        }
        Scope sc(KC_SWEEP_X + d);
        int n = 0;
        if (d == 0 && vo) {
            // rows 1 .. n-1 now; then, with the y halo in place: the gradients of rows -1, 0, n, n+1 and
            // the x sweep of rows 0 and n
            a.xs_k0 = 1; a.xs_nk = S.g.N[1] - 1; a.xs_z0 = 0; a.xs_nz = 1;
            n = S.L->sweep(S.nf, S.nd, 0, a, S.st);
            if (!n) return fail(MFC_B200_EUNSUPPORTED, "no sweep kernel instantiated for this (num_fluids, num_dims)");
            CK(cudaStreamWaitEvent(S.st, S.ev_halo[1], 0));
            if ((rc = physical_bc_dir(q, 1))) return rc;
            for (int side = 0; side < 2; side++) {
                va.k_lo = side == 0 ? -1 : S.g.N[1]; va.k_hi = va.k_lo + 1;
                S.launches += S.L->vgrad(S.nd, va, S.st);
            }
            for (int side = 0; side < 2; side++) {
                a.xs_k0 = side == 0 ? 0 : S.g.N[1]; a.xs_nk = 1;
                n += S.L->sweep(S.nf, S.nd, 0, a, S.st);
            }
            a.xs_nk = 0;
        } else if (d == 0 && S.xp_pending) {
            // the x halo arrives piece by piece (ghosts_begin): sweep each piece as soon as it is there
            if ((rc = physical_bc_dir(q, 0))) return rc;
            CK(cudaEventRecord(S.ev_pre, S.st));
            for (int p = 0; p < S.xpieces; p++) {
                long long idx0, cnt;
                xpiece_range(p, a.xs_k0, a.xs_nk, a.xs_z0, a.xs_nz, idx0, cnt);
                cudaStream_t xs = p == 0 ? S.st : S.xstream[p];
                if (p > 0) CK(cudaStreamWaitEvent(xs, S.ev_pre, 0));
                CK(cudaStreamWaitEvent(xs, S.ev_xp[p], 0));
                const int np = S.L->sweep(S.nf, S.nd, 0, a, xs);
                if (!np) return fail(MFC_B200_EUNSUPPORTED, "no sweep kernel instantiated for this (num_fluids, num_dims)");
                n += np;
                if (p > 0) CK(cudaEventRecord(S.ev_xdone[p], xs));
            }
            for (int p = 1; p < S.xpieces; p++) CK(cudaStreamWaitEvent(S.st, S.ev_xdone[p], 0));
            a.xs_nk = 0;
            S.xp_pending = false; S.halo_pending[0] = false;
        } else if (split_x) {
            a.xsplit = 1;
            const int n1 = S.L->sweep(S.nf, S.nd, d, a, S.st);
            if (!n1) return fail(MFC_B200_EUNSUPPORTED, "no sweep kernel instantiated for this (num_fluids, num_dims)");
            if ((rc = ghosts_ready(q, d))) return rc;
            a.xsplit = 2;
            const int n2 = S.L->sweep(S.nf, S.nd, d, a, S.st);
            n = (n1 > 0 ? n1 : 0) + (n2 > 0 ? n2 : 0);
        } else {
            n = S.L->sweep(S.nf, S.nd, d, a, S.st);
            if (!n) return fail(MFC_B200_EUNSUPPORTED, "no sweep kernel instantiated for this (num_fluids, num_dims)");
        }
        sc.done(n);
        if (fuse_stab && d == 0 && (rc = stab_reduce_and_copy())) return rc;
        if (S.viscous && !S.visc_fused) {                    // m_rhs.fpp:591-604, :639-652
            va.dir = d; va.bc_beg = S.p.bc[2*d]; va.bc_end = S.p.bc[2*d + 1];
            // the RK statement cannot be fused into the last sweep (the viscous terms come after
            // it); it is applied by the last direction's k_visc, which completes the RHS
            va.rk_mode = d == S.nd - 1 ? rk_mode : 0;
            Scope sv(KC_VISC); sv.done(S.L->visc(S.nd, va, S.st));
        }
    }
    return 0;
}

int do_step(int t_step, double dt, bool stab) {
    double *q1 = S.state[S.cur], *A = S.state[(S.cur + 1) % 3], *B = S.state[(S.cur + 2) % 3];
    int rc;
    const int ts = S.p.time_stepper;
    if (ts == 1) {                                                         // m_time_steppers.fpp:129-193
        if ((rc = rhs_stage(q1, 1, q1, A, dt, t_step, true, stab))) return rc;
        if (t_step != S.p.t_step_stop) S.cur = (S.cur + 1) % 3;            // A becomes q_cons_ts(1)
    } else if (ts == 2) {                                                  // :197-267
        if ((rc = rhs_stage(q1, 1, q1, A, dt, t_step, true, stab))) return rc;
        if (t_step == S.p.t_step_stop) return 0;
        if ((rc = rhs_stage(A, 2, q1, q1, dt, t_step, false, false))) return rc;
    } else {                                                               // :271-362
        if ((rc = rhs_stage(q1, 1, q1, A, dt, t_step, true, stab))) return rc;
        if (t_step == S.p.t_step_stop) return 0;
        if ((rc = rhs_stage(A, 3, q1, B, dt, t_step, false, false))) return rc;
        if ((rc = rhs_stage(B, 4, q1, q1, dt, t_step, false, false))) return rc;
    }
    return 0;
}

// One RK step through a CUDA graph: captured on first use for this (dt, diagnostics, buffer
// rotation) and replayed afterwards.  The host-side bookkeeping do_step performs (buffer rotation,
// which state q_prim_vf reflects, a pending stability read-back) is recorded with the graph and
// re-applied on replay.  Falls back to direct launches when capture is not possible.
int do_step_graphed(int t_step, double dt, bool stab) {
    unsigned long long bits;
    std::memcpy(&bits, &dt, sizeof(bits));
    for (auto &gr : S.graphs)
        if (gr.dt_bits == bits && gr.stab == (int)stab && gr.cur_before == S.cur) {
            CK(cudaGraphLaunch(gr.exec, S.st));
            S.cur = gr.cur_after; S.last_q = gr.last_q_after; S.stab_pending = gr.stab_pending_after;
            S.launches += gr.launches;
            return 0;
        }
    if (S.graphs.size() >= 8) {                          // dt keeps changing: not worth caching
        for (auto &gr : S.graphs) cudaGraphExecDestroy(gr.exec);
        S.graphs.clear();
    }
    Sim::StepGraph gr{};
    gr.dt_bits = bits; gr.stab = (int)stab; gr.cur_before = S.cur;
    const int64_t l0 = S.launches;
    if (cudaStreamBeginCapture(S.st, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
        cudaGetLastError();
        return do_step(t_step, dt, stab);
    }
    const int rc = do_step(t_step, dt, stab);
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(S.st, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (ce != cudaSuccess || !graph) return fail(MFC_B200_ECUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce));
    const cudaError_t ie = cudaGraphInstantiate(&gr.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) return fail(MFC_B200_ECUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ie));
    gr.cur_after = S.cur; gr.last_q_after = S.last_q; gr.stab_pending_after = S.stab_pending;
    gr.launches = S.launches - l0;
    S.graphs.push_back(gr);
    CK(cudaGraphLaunch(gr.exec, S.st));                  // the capture only recorded the work
    return 0;
}
// graphs: single rank (no NCCL calls inside a capture), no per-kernel profiling events, and not the
// final "convert only" call
bool graph_ok(int t_step) { return S.graph_on && !S.comm && !S.prof && t_step != S.p.t_step_stop; }

// host field (Fortran sf(-b:m+b, ...), x fastest, contiguous) <-> padded device plane.  The
// host array crosses PCIe as ONE contiguous copy into a staging plane (the RHS accumulator, which
// is scratch between steps) and is re-pitched by a kernel: a strided 2-D copy of 4 KB rows
// straight from host memory reaches only ~8 GB/s, and the copy engine's device-to-device 2-D copy
// of 270 k short rows is no faster (~10 GB/s end to end, measured); the kernel runs at HBM speed,
// so the transfer is bound by the PCIe link (~55 GB/s measured with pinned host memory).
__global__ void __launch_bounds__(256) k_repitch(double *pitched, double *packed, int w, int pitch, bool to_pitched) {
    const size_t r = blockIdx.x;
    double *p = pitched + r*(size_t)pitch, *c = packed + r*(size_t)w;
    for (int j = threadIdx.x; j < w; j += 256) {
        if (to_pitched) p[j] = c[j];
        else c[j] = p[j];
    }
}
int copy_field(double *dev_plane, const double *host, bool to_device, cudaStream_t st, int v) {
    const GridDesc &g = S.g;
    const int wd = g.N[0] + 1 + 2*g.b;
    const size_t w = (size_t)wd*sizeof(double);
    double *d0 = dev_plane + (kXoff - g.b);
    const size_t rows = (size_t)g.ey*g.ez;
    double *stage = S.rhs + (size_t)v*g.fstride;         // w*rows <= fstride*8 bytes
    if (to_device) {
        CK(cudaMemcpyAsync(stage, host, w*rows, cudaMemcpyHostToDevice, st));
        k_repitch<<<(unsigned)rows, 256, 0, st>>>(d0, stage, wd, g.pitch, true);
    } else {
        k_repitch<<<(unsigned)rows, 256, 0, st>>>(d0, stage, wd, g.pitch, false);
        CK(cudaMemcpyAsync(const_cast<double *>(host), stage, w*rows, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaGetLastError());
    return 0;
}

}  // namespace

extern "C" {

const char *mfc_b200_last_error(void) { return S.err.c_str(); }

int mfc_b200_init(const mfc_b200_params_t *p) {
    if (!p) return fail(MFC_B200_EINVAL, "params is NULL");
    if (p->abi_version != MFC_B200_ABI_VERSION) return fail(MFC_B200_EINVAL, "ABI version mismatch");
    if (S.inited) free_all();
    // the checks of s_check_input_file that guard this path (m_start_up.fpp:147-229)
    const int nd = p->num_dims, nf = p->num_fluids;
    if (nd < 1 || nd > 3) return fail(MFC_B200_EINVAL, "Unsupported value of num_dims");
    if (nf < 1 || nf > MFC_B200_MAX_FLUIDS)
        return fail(MFC_B200_EINVAL, "Unsupported value of num_fluids (kernels are instantiated for 1.." + std::to_string(MFC_B200_MAX_FLUIDS) + " fluids). Exiting ...");
    if (nf > MFC_B200_BUILT_FLUIDS)
        return fail(MFC_B200_EUNSUPPORTED, "num_fluids = " + std::to_string(nf) + ": the sweep kernels are instantiated for 1.." +
                                           std::to_string(MFC_B200_BUILT_FLUIDS) + " fluids (MFC_DISPATCH in kernels_inst.inc)");
    if (p->sys_size != 2*nf + nd + 1) return fail(MFC_B200_EINVAL, "sys_size /= 2*num_fluids + num_dims + 1");
    if (p->m <= 0) return fail(MFC_B200_EINVAL, "Unsupported value of m. Exiting ...");
    if (p->n < 0 || (nd > 1) != (p->n > 0)) return fail(MFC_B200_EINVAL, "Unsupported value of n. Exiting ...");
    if (p->p < 0 || (nd > 2) != (p->p > 0)) return fail(MFC_B200_EINVAL, "Unsupported value of p. Exiting ...");
    if (p->weno_order != 1 && p->weno_order != 3 && p->weno_order != 5)
        return fail(MFC_B200_EINVAL, "Unsupported value of weno_order. Exiting ...");
    if (!(p->weno_eps > 0.0) || p->weno_eps > 1e-6) return fail(MFC_B200_EINVAL, "Unsupported value of weno_eps. Exiting ...");
    if (p->time_stepper < 1 || p->time_stepper > 3) return fail(MFC_B200_EINVAL, "Unsupported value of time_stepper. Exiting ...");
    bool visc = false;
    for (int i = 0; i < nf; i++) {
        if (!(p->gammas[i] > 0.0)) return fail(MFC_B200_EINVAL, "Unsupported value of fluid_pp(i)%gamma. Exiting ...");
        if (p->pi_infs[i] < 0.0) return fail(MFC_B200_EINVAL, "Unsupported value of fluid_pp(i)%pi_inf. Exiting ...");
        if (p->Re[i][0] > 0.0 || p->Re[i][1] > 0.0) visc = true;
    }
    const int polyn = (p->weno_order - 1)/2;
    const int b_expect = visc ? 2*polyn + 2 : polyn + 2;
    if (p->buff_size != b_expect) return fail(MFC_B200_EINVAL, "buff_size does not match weno_order / viscosity (m_global_parameters.fpp:356-360)");
    for (int d = 0; d < nd; d++) {
        if (!p->cb[d] || !p->ds[d] || !p->cc[d]) return fail(MFC_B200_EINVAL, "grid metric pointer is NULL");
        for (int s = 0; s < 2; s++) {
            const int c = p->bc[2*d + s];
            if (c < -12 || c >= p->num_procs) return fail(MFC_B200_EINVAL, "Unsupported value of bc. Exiting ...");
        }
    }
    if (p->weno_order != 5 && visc) return fail(MFC_B200_EUNSUPPORTED, "viscous fluxes are built for weno_order = 5 only");
    if (visc && nd == 3) return fail(MFC_B200_EUNSUPPORTED, "viscous fluxes exist in 1D/2D only (the reference has no 3D; the 3D extension is inviscid)");

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(MFC_B200_ENODEVICE, "no CUDA device: libmfc_b200 has no CPU fallback");
    int dev = p->device >= 0 ? p->device : p->proc_rank % ndev;      // p_main.fpp:95-99
    if (dev >= ndev) return fail(MFC_B200_EINVAL, "device ordinal out of range");
    CK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    if (prop.major < 10) return fail(MFC_B200_ENODEVICE, "kernels are built for sm_100a only");

    S.p = *p;
    for (int d = 0; d < 3; d++) S.p.cb[d] = S.p.cc[d] = S.p.ds[d] = nullptr;    // never keep host pointers
    S.nf = nf; S.nd = nd; S.E = p->sys_size; S.b = p->buff_size; S.viscous = visc;
    S.L = p->strict_math ? &launchers_strict() : &launchers_fast();
    {
        const char *e = std::getenv("MFC_B200_VISC_FUSED");    // 0: keep the separate k_visc pass (A/B measurements)
        S.visc_fused = visc && !p->strict_math && !p->weno_Re_flux && !(e && e[0] == '0');
    }
    S.Re_size[0] = S.Re_size[1] = 0;                                   // m_global_parameters.fpp:314-339, m_rhs.fpp:385-390
    for (int i = 0; i < nf; i++)
        for (int k = 0; k < 2; k++)
            if (p->Re[i][k] > 0.0) { S.Re_idx[k][S.Re_size[k]] = i; S.Res[k][S.Re_size[k]] = p->Re[i][k]; S.Re_size[k]++; }
    S.g = make_grid(p->m, p->n, p->p, nd, S.b);
    if (S.g.fstride >= (1LL << 32)) return fail(MFC_B200_EUNSUPPORTED, "more than 2^32 elements per field (kernels use 32-bit in-plane offsets)");
    for (int d = 0; d < 3; d++)
        for (int s = 0; s < 2; s++) {
            int c = p->bc[2*d + s];
            if (c >= 0 && c == p->proc_rank) c = -1;     // my own periodic neighbour: plain periodic fill
            S.bc[d][s] = d < nd ? c : -3;
        }
    CK(cudaStreamCreateWithFlags(&S.st, cudaStreamNonBlocking));
    CK(cudaEventCreate(&S.ev0)); CK(cudaEventCreate(&S.ev1)); CK(cudaEventCreate(&S.sv0)); CK(cudaEventCreate(&S.sv1));
    const size_t sb = field_bytes()*S.E;
    for (auto &s : S.state) { CK(cudaMalloc(&s, sb)); CK(cudaMemsetAsync(s, 0, sb, S.st)); }
    CK(cudaMalloc(&S.prim, field_bytes()*(nd + 1))); CK(cudaMemsetAsync(S.prim, 0, field_bytes()*(nd + 1), S.st));
    CK(cudaMalloc(&S.rhs, sb)); CK(cudaMemsetAsync(S.rhs, 0, sb, S.st));
    for (int w = 0; w < 2; w++) {
        int rc;
        for (int i = 0; i < 3; i++)
            if ((rc = make_tmap(S.tm_state[i][w], S.state[i], w == 0 ? kWX : kWY))) return rc;
        if ((rc = make_tmap(S.tm_rhs[w], S.rhs, w == 0 ? kWX : kWY))) return rc;
    }
    {
        int rc;
        for (int i = 0; i < 3; i++)
            if ((rc = make_tmap(S.tm_state[i][2], S.state[i], kWY, true))) return rc;
        if ((rc = make_tmap(S.tm_rhs[2], S.rhs, kWY, true))) return rc;
    }
    if (visc) {
        CK(cudaMalloc(&S.visc_face, field_bytes()*(nd + 2))); CK(cudaMemsetAsync(S.visc_face, 0, field_bytes()*(nd + 2), S.st));
        CK(cudaMalloc(&S.visc_grad, field_bytes()*nd*nd)); CK(cudaMemsetAsync(S.visc_grad, 0, field_bytes()*nd*nd, S.st));
    }
    CK(cudaMalloc(&S.stab_dev, 3*sizeof(unsigned long long)));
    CK(cudaMallocHost(&S.stab_host, 3*sizeof(unsigned long long)));
    CK(cudaMallocHost(&S.stab_init, 3*sizeof(unsigned long long)));
    for (int d = 0; d < nd; d++) {
        const int N = S.g.N[d], b = S.b;
        WenoTable t = build_weno_table(p->cb[d], N, b, p->weno_order);
        S.clen[d] = t.len; S.coef_lo[d] = t.lo; S.h_coef[d] = t.data;
        {   // uniform grid?  then every cell's coefficients equal the classical WENO5-JS rationals
            // to rounding, and the fast kernels use those as compile-time constants
            // (weno5_uniform in kernels.cuh); otherwise they read the per-cell tables
            static const double classic[kNumWenoCoef] = {
                1.0/3, -5.0/6, -1.0/6, -1.0/3, -2.0/3, 1.0/6,            // poly_coef_cbL
                -1.0/6, 2.0/3, 1.0/3, 1.0/6, 5.0/6, -1.0/3,              // poly_coef_cbR
                0.1, 0.6, 0.3, 0.3, 0.6, 0.1,                            // d_cbL, d_cbR
                4.0/3, -11.0/3, 10.0/3, 4.0/3, -5.0/3, 4.0/3, 10.0/3, -11.0/3, 4.0/3};   // beta_coef
            // Tolerance: the tables are computed from the cell-boundary coordinates, whose own
            // rounding (eps |s| against a width ds) reappears in the coefficients amplified by
            // |s|/ds; a deviation explainable by that noise is the uniform grid the user asked for.
            double smax = 0.0, dsmin = 1e300;
            for (int i = 0; i < N + 2 + 2*b; i++) smax = std::fmax(smax, std::fabs(p->cb[d][i]));
            for (int i = 0; i < N + 1 + 2*b; i++) dsmin = std::fmin(dsmin, p->ds[d][i]);
            const double tol = std::fmax(1e-12, 16.0*2.220446049250313e-16*smax/dsmin);
            bool uni = true;
            for (int c = 0; c < kNumWenoCoef; c++) {
                S.cuni[d][c] = classic[c];
                for (int i = 0; i < t.len && uni; i++)
                    if (std::fabs(t.data[(size_t)c*t.len + i] - classic[c]) > tol) uni = false;
            }
            S.coef_uniform[d] = (uni && p->weno_order == 5) ? 1 : 0;
        }
        CK(cudaMalloc(&S.coef[d], t.data.size()*sizeof(double)));
        CK(cudaMemcpyAsync(S.coef[d], S.h_coef[d].data(), t.data.size()*sizeof(double), cudaMemcpyHostToDevice, S.st));
        std::vector<double> r((size_t)N + 1 + 2*b);
        for (size_t i = 0; i < r.size(); i++) r[i] = 1.0/p->ds[d][i];          // "1d0/dx(k)" of m_rhs.fpp:571
        CK(cudaMalloc(&S.rds[d], r.size()*sizeof(double)));
        CK(cudaMalloc(&S.ds[d], r.size()*sizeof(double)));
        CK(cudaMemcpy(S.rds[d], r.data(), r.size()*sizeof(double), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(S.ds[d], p->ds[d], r.size()*sizeof(double), cudaMemcpyHostToDevice));
        CK(cudaMalloc(&S.cc[d], r.size()*sizeof(double)));
        CK(cudaMemcpy(S.cc[d], p->cc[d], r.size()*sizeof(double), cudaMemcpyHostToDevice));
        for (size_t i = 0; i + 1 < r.size(); i++) r[i] = 1.0/(p->cc[d][i + 1] - p->cc[d][i]);
        r[r.size() - 1] = 0.0;
        CK(cudaMalloc(&S.rdcc[d], r.size()*sizeof(double)));
        CK(cudaMemcpy(S.rdcc[d], r.data(), r.size()*sizeof(double), cudaMemcpyHostToDevice));
        for (int s = 0; s < 2; s++)
            if (S.bc[d][s] >= 0) {
                const size_t n = (size_t)slab_count(S.g, d)*S.E*sizeof(double);
                CK(cudaMalloc(&S.sendbuf[d][s], n)); CK(cudaMalloc(&S.recvbuf[d][s], n));
            }
    }
    CK(cudaStreamSynchronize(S.st));
    {   // CUDA graphs per RK step: on by default for grids whose step is launch-bound (< 4 M cells);
        // MFC_B200_GRAPH=0 / 1 forces it off / on
        const char *e = std::getenv("MFC_B200_GRAPH");
        const long long cells = (long long)(S.g.N[0] + 1)*(S.g.N[1] + 1)*(S.g.N[2] + 1);
        S.graph_on = e ? e[0] != '0' : cells < (4LL << 20);
    }
    {
        const char *e = std::getenv("MFC_B200_BCMAP");         // 0: fill the y / z ghost rows with k_bc (A/B measurements)
        S.bc_map = !(e && e[0] == '0');
    }
    S.cur = 0; S.launches = 0; S.inited = true; S.uploaded = false; S.last_q = nullptr;
    S.err.clear();
    return 0;
}

int mfc_b200_get_unique_id(unsigned char id[128]) {
    if (!S.nccl.load(S.err)) return MFC_B200_ENCCL;
    ncclUniqueId u;
    NK(S.nccl.GetUniqueId(&u));
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    std::memcpy(id, &u, 128);
    return 0;
}

int mfc_b200_comm_init(const unsigned char id[128], int rank, int nranks) {
    if (!S.inited) return fail(MFC_B200_ESTATE, "mfc_b200_comm_init before mfc_b200_init");
    if (rank != S.p.proc_rank || nranks != S.p.num_procs) return fail(MFC_B200_EINVAL, "rank / nranks disagree with params");
    if (!S.nccl.load(S.err)) return MFC_B200_ENCCL;
    ncclUniqueId u;
    std::memcpy(&u, id, 128);
    NK(S.nccl.CommInitRank(&S.comm, nranks, u, rank));
    {   // halo traffic ahead of the sweeps' CTAs whenever an SM slot frees up
        int lo = 0, hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CK(cudaStreamCreateWithPriority(&S.cs, cudaStreamNonBlocking, hi));
    }
    CK(cudaEventCreateWithFlags(&S.ev_q, cudaEventDisableTiming));
    for (auto &e : S.ev_halo) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    {
        const char *e = std::getenv("MFC_B200_OVERLAP");       // 0 disables (for A/B measurements)
        S.overlap = !S.viscous && !(e && e[0] == '0');
        const char *x = std::getenv("MFC_B200_XSPLIT");
        S.xsplit_on = x && x[0] == '1';
        for (auto &ev : S.ev_xp) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        for (auto &ev : S.ev_xdone) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&S.ev_pre, cudaEventDisableTiming));
        for (auto &x : S.xstream) CK(cudaStreamCreateWithFlags(&x, cudaStreamNonBlocking));
        // x halo pieces of >= 24 MB per message (smaller ones are latency-bound and gain nothing; measured at
        // 512^3, 67 MB per face: 2 pieces 53.95 ms/step, 4 pieces 54.16, 8 pieces 54.21, unpieced 55.10)
        const long long bytes = slab_count(S.g, 0)*S.E*(long long)sizeof(double);
        const char *pc = std::getenv("MFC_B200_XPIECES");
        long long P = pc ? std::atoll(pc) : bytes/(24LL << 20);
        const int lim = S.nd == 3 ? S.g.N[2] + 1 : S.g.N[1] + 1;
        if (P > Sim::kMaxPieces) P = Sim::kMaxPieces;
        if (P > lim/8) P = lim/8;
        S.xpieces = (S.nd >= 2 && S.overlap && P >= 2) ? (int)P : 1;
        S.visc_overlap = S.visc_fused && !(e && e[0] == '0');
    }
    return 0;
}

int mfc_b200_upload(const double *const q_cons[]) {
    if (!S.inited) return fail(MFC_B200_ESTATE, "mfc_b200_upload before mfc_b200_init");
    for (int v = 0; v < S.E; v++) {
        int rc = copy_field(S.state[S.cur] + (size_t)v*S.g.fstride, q_cons[v], true, S.st, v);
        if (rc) return rc;
    }
    CK(cudaStreamSynchronize(S.st));
    S.uploaded = true;
    return 0;
}

int mfc_b200_generate_initial_condition(int32_t num_patches, const mfc_b200_patch_t *patches,
                                        const double *const cc[3], double ds_min) {
    return mfc_b200_generate_initial_condition2(num_patches, patches, cc, nullptr, ds_min);
}

int mfc_b200_generate_initial_condition2(int32_t num_patches, const mfc_b200_patch_t *patches,
                                         const double *const cc[3], const double *const cb[3], double ds_min) {
    if (!S.inited) return fail(MFC_B200_ESTATE, "mfc_b200_generate_initial_condition before mfc_b200_init");
    if (num_patches < 1 || num_patches > MFC_B200_MAX_PATCHES || !patches || !cc)
        return fail(MFC_B200_EINVAL, "num_patches must be 1..MFC_B200_MAX_PATCHES with non-NULL arrays");
    static_assert(sizeof(PatchDesc) == sizeof(mfc_b200_patch_t), "PatchDesc mirrors mfc_b200_patch_t");
    for (int i = 0; i < num_patches; i++) {
        const int geo = patches[i].geometry;
        // m_initial_condition.fpp:50-100: 1, 15 in 1-D; 2..7, 18 in 2-D; 8, 9, 10 are the 3-D extension
        const bool ok = ((geo == 1 || geo == 15) && S.nd == 1) || (geo == 1 && S.nd > 1) ||
                        ((geo == 2 || geo == 3 || geo == 4 || geo == 5 || geo == 6 || geo == 7 || geo == 18) && S.nd >= 2) ||
                        ((geo == 8 || geo == 9 || geo == 10) && S.nd == 3);
        if (!ok) return fail(MFC_B200_EUNSUPPORTED, "patch geometry " + std::to_string(geo) + " is not built for this num_dims");
        if ((geo == 7 || geo == 15) && !cb)
            return fail(MFC_B200_EINVAL, "analytical patches (geometry 7, 15) need the cell boundaries: call mfc_b200_generate_initial_condition2 with cb");
        if (patches[i].smooth_patch_id < 0 || patches[i].smooth_patch_id > num_patches)
            return fail(MFC_B200_EINVAL, "smooth_patch_id out of range");
    }
    const GridDesc &g = S.g;
    PatchArgs a{};
    a.g = g; a.q = S.state[S.cur]; a.num_patches = num_patches; a.ds_min = ds_min;
    for (int i = 0; i < kMaxFluids; i++) { a.gammas[i] = S.p.gammas[i]; a.pi_infs[i] = S.p.pi_infs[i]; }
    // scratch for the cell centres and the patch table: the RHS accumulator (idle between steps)
    double *scr = S.rhs;
    size_t used = 0;
    for (int d = 0; d < S.nd; d++) {
        if (!cc[d]) return fail(MFC_B200_EINVAL, "cc[d] is NULL for an active direction");
        used += 2*(((size_t)g.N[d] + 1 + 15)/16*16);
    }
    if (used*sizeof(double) + (size_t)num_patches*sizeof(PatchDesc) > field_bytes()*S.E)
        return fail(MFC_B200_ENOMEM, "grid too small to stage the patch table");
    PatchDesc *pd = reinterpret_cast<PatchDesc *>(scr + used);
    used = 0;
    for (int d = 0; d < S.nd; d++) {
        const size_t n = (size_t)g.N[d] + 1;
        CK(cudaMemcpyAsync(scr + used, cc[d], n*sizeof(double), cudaMemcpyHostToDevice, S.st));
        a.cc[d] = scr + used;
        used += (n + 15)/16*16;
        a.cb[d] = nullptr;
        if (cb && cb[d]) {
            CK(cudaMemcpyAsync(scr + used, cb[d], n*sizeof(double), cudaMemcpyHostToDevice, S.st));
            a.cb[d] = scr + used;
        }
        used += (n + 15)/16*16;
    }
    CK(cudaMemcpyAsync(pd, patches, (size_t)num_patches*sizeof(PatchDesc), cudaMemcpyHostToDevice, S.st));
    a.patches = pd;
    CK(cudaMemsetAsync(S.state[S.cur], 0, field_bytes()*S.E, S.st));   // ghosts: rebuilt by every RHS evaluation
    {
        Scope sc(KC_PATCH);
        const int n = launch_patches(S.nf, S.nd, a, S.st);
        if (!n) return fail(MFC_B200_EUNSUPPORTED, "no patch kernel instantiated for this (num_fluids, num_dims)");
        sc.done(n);
    }
    CK(cudaStreamSynchronize(S.st));
    CK(cudaGetLastError());
    S.uploaded = true;
    return 0;
}

int mfc_b200_download(double *const q_cons[]) {
    if (!S.uploaded) return fail(MFC_B200_ESTATE, "mfc_b200_download before mfc_b200_upload");
    for (int v = 0; v < S.E; v++) {
        int rc = copy_field(S.state[S.cur] + (size_t)v*S.g.fstride, q_cons[v], false, S.st, v);
        if (rc) return rc;
    }
    CK(cudaStreamSynchronize(S.st));
    return 0;
}

int mfc_b200_download_prim(double *const q_prim[]) {
    if (!S.uploaded) return fail(MFC_B200_ESTATE, "mfc_b200_download_prim before mfc_b200_upload");
    // primitive variables of the most recent RHS evaluation's state (q_prim_qp): partial
    // densities and volume fractions alias the conservative state (m_rhs.fpp:154-164)
    int rc;
    if ((rc = fill_ghosts(S.state[S.cur]))) return rc;
    if ((rc = run_prim(S.state[S.cur]))) return rc;
    for (int v = 0; v < S.E; v++) {
        const bool is_prim = v >= S.nf && v <= S.nf + S.nd;
        double *plane = is_prim ? S.prim + (size_t)(v - S.nf)*S.g.fstride : S.state[S.cur] + (size_t)v*S.g.fstride;
        if ((rc = copy_field(plane, q_prim[v], false, S.st, v))) return rc;
    }
    CK(cudaStreamSynchronize(S.st));
    return 0;
}

int mfc_b200_step(int t_step, double dt, double stab[3], double *step_seconds) {
    if (!S.uploaded) return fail(MFC_B200_ESTATE, "mfc_b200_step before mfc_b200_upload");
    CK(cudaEventRecord(S.sv0, S.st));
    int rc = graph_ok(t_step) ? do_step_graphed(t_step, dt, S.p.run_time_info != 0) : do_step(t_step, dt, S.p.run_time_info != 0);
    if (rc) return rc;
    CK(cudaEventRecord(S.sv1, S.st));
    CK(cudaStreamSynchronize(S.st));
    CK(cudaGetLastError());
    stab_fetch(stab);
    if (step_seconds) { float ms = 0.f; CK(cudaEventElapsedTime(&ms, S.sv0, S.sv1)); *step_seconds = ms*1e-3; }
    return 0;
}

int mfc_b200_step_async(int t_step, double dt, int n_steps) {
    if (!S.uploaded) return fail(MFC_B200_ESTATE, "mfc_b200_step_async before mfc_b200_upload");
    for (int s = 0; s < n_steps; s++) {
        int rc = graph_ok(t_step + s) ? do_step_graphed(t_step + s, dt, false) : do_step(t_step + s, dt, false);
        if (rc) return rc;
    }
    return 0;
}

int mfc_b200_sync(void) {
    if (!S.inited) return fail(MFC_B200_ESTATE, "mfc_b200_sync before mfc_b200_init");
    CK(cudaStreamSynchronize(S.st));
    CK(cudaGetLastError());
    return 0;
}

int mfc_b200_compute_rhs(const double *const q_cons[], double *const rhs[]) {
    if (!S.inited) return fail(MFC_B200_ESTATE, "mfc_b200_compute_rhs before mfc_b200_init");
    double *scratch = S.state[(S.cur + 1) % 3];
    int rc;
    for (int v = 0; v < S.E; v++)
        if ((rc = copy_field(scratch + (size_t)v*S.g.fstride, q_cons[v], true, S.st, v))) return rc;
    if ((rc = rhs_stage(scratch, 0, scratch, scratch, 0.0, S.p.t_step_stop - 1, false, false))) return rc;
    const GridDesc &g = S.g;
    const size_t w = (size_t)(g.N[0] + 1)*sizeof(double);
    for (int v = 0; v < S.E; v++)
        for (int l = 0; l <= g.N[2]; l++) {
            const double *src = S.rhs + (size_t)v*g.fstride + g.at(0, 0, l);
            double *dst = rhs[v] + (size_t)l*(g.N[0] + 1)*(g.N[1] + 1);
            CK(cudaMemcpy2DAsync(dst, w, src, (size_t)g.pitch*sizeof(double), w, (size_t)g.N[1] + 1, cudaMemcpyDeviceToHost, S.st));
        }
    CK(cudaStreamSynchronize(S.st));
    CK(cudaGetLastError());
    return 0;
}

int mfc_b200_finalize(void) {
    prof_collect();
    free_all();
    return 0;
}

int mfc_b200_get_weno_coefficients(int dir, double *poly_L, double *poly_R, double *d_L, double *d_R, double *beta) {
    if (!S.inited) return fail(MFC_B200_ESTATE, "mfc_b200_get_weno_coefficients before mfc_b200_init");
    if (dir < 0 || dir >= S.nd) return fail(MFC_B200_EINVAL, "dir out of range");
    const int n = S.clen[dir];
    const double *c = S.h_coef[dir].data();
    for (int i = 0; i < n; i++) {
        for (int k = 0; k < 6; k++) {
            if (poly_L) poly_L[(size_t)i*6 + k] = c[(size_t)k*n + i];
            if (poly_R) poly_R[(size_t)i*6 + k] = c[(size_t)(6 + k)*n + i];
        }
        for (int k = 0; k < 3; k++) {
            if (d_L) d_L[(size_t)i*3 + k] = c[(size_t)(12 + k)*n + i];
            if (d_R) d_R[(size_t)i*3 + k] = c[(size_t)(15 + k)*n + i];
        }
        for (int k = 0; k < 9; k++)
            if (beta) beta[(size_t)i*9 + k] = c[(size_t)(18 + k)*n + i];
    }
    return 0;
}

int64_t mfc_b200_kernel_launches(void) { return S.launches; }

int mfc_b200_state_snapshot(void) {
    if (!S.uploaded) return fail(MFC_B200_ESTATE, "mfc_b200_state_snapshot before mfc_b200_upload");
    const size_t sb = field_bytes()*S.E;
    if (!S.snap) CK(cudaMalloc(&S.snap, sb));
    CK(cudaMemcpyAsync(S.snap, S.state[S.cur], sb, cudaMemcpyDeviceToDevice, S.st));
    CK(cudaStreamSynchronize(S.st));
    return 0;
}

int mfc_b200_state_restore(void) {
    if (!S.snap) return fail(MFC_B200_ESTATE, "mfc_b200_state_restore without a snapshot");
    CK(cudaMemcpyAsync(S.state[S.cur], S.snap, field_bytes()*S.E, cudaMemcpyDeviceToDevice, S.st));
    CK(cudaStreamSynchronize(S.st));
    return 0;
}

int mfc_b200_timer_start(void) {
    if (!S.inited) return fail(MFC_B200_ESTATE, "mfc_b200_timer_start before mfc_b200_init");
    CK(cudaEventRecord(S.ev0, S.st));
    return 0;
}

int mfc_b200_timer_stop(double *seconds) {
    if (!S.inited) return fail(MFC_B200_ESTATE, "mfc_b200_timer_stop before mfc_b200_init");
    CK(cudaEventRecord(S.ev1, S.st));
    CK(cudaEventSynchronize(S.ev1));
    CK(cudaGetLastError());
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, S.ev0, S.ev1));
    if (seconds) *seconds = ms*1e-3;
    return 0;
}

int mfc_b200_profile_enable(int on) {
    prof_collect();
    S.prof = on != 0;
    if (on) for (int i = 0; i < KC_COUNT; i++) { S.prof_s[i] = 0; S.prof_n[i] = 0; }
    return 0;
}

int mfc_b200_profile_get(int kc, double *seconds, int64_t *launches) {
    if (kc < 0 || kc >= KC_COUNT) return fail(MFC_B200_EINVAL, "kernel class out of range");
    prof_collect();
    if (seconds) *seconds = S.prof_s[kc];
    if (launches) *launches = S.prof_n[kc];
    return 0;
}

const char *mfc_b200_kernel_name(int kc) { return kc >= 0 && kc < KC_COUNT ? kKernelNames[kc] : nullptr; }

// Test hook: rebuild the ghost cells of q_cons_ts(1) in place.
//   mode 0: the production path (fill_ghosts: k_bc for physical sides, pack / NCCL / unpack for
//           processor boundaries)
//   mode 1: every PERIODIC direction is filled by the halo-exchange kernels instead of k_bc --
//           k_halo_pack of my first / last b layers, a device-to-device copy standing in for
//           ncclSend/ncclRecv to myself, k_halo_unpack into the opposite ghost layers -- in the
//           reference's direction order (m_rhs.fpp:692-905).  On one GPU this exercises the
//           x, y and z pack / unpack index maps (m_mpi_proxy.fpp:490-601, 733-969) against the
//           periodic ghost fill, which must agree bit for bit.
int mfc_b200_debug_fill_ghosts(int mode) {
    if (!S.uploaded) return fail(MFC_B200_ESTATE, "mfc_b200_debug_fill_ghosts before mfc_b200_upload");
    double *q = S.state[S.cur];
    int rc;
    if (mode == 0) {
        if ((rc = fill_ghosts(q))) return rc;
    } else {
        for (int d = 0; d < S.nd; d++) {
            if (S.bc[d][0] == -1 && S.bc[d][1] == -1) {
                const size_t n = (size_t)slab_count(S.g, d)*S.E;
                double *buf = nullptr;
                CK(cudaMalloc(&buf, 4*n*sizeof(double)));
                for (int s = 0; s < 2; s++) {
                    HaloArgs h{S.g, q, buf + s*n, d, s, S.E};
                    Scope sc(KC_PACK); sc.done(S.L->halo_pack(h, S.st));
                }
                // my first layers are my end neighbour's (= my own) ... beg-side message and vice versa
                CK(cudaMemcpyAsync(buf + 2*n, buf + n, n*sizeof(double), cudaMemcpyDeviceToDevice, S.st));   // last layers -> beg ghosts
                CK(cudaMemcpyAsync(buf + 3*n, buf, n*sizeof(double), cudaMemcpyDeviceToDevice, S.st));       // first layers -> end ghosts
                for (int s = 0; s < 2; s++) {
                    HaloArgs h{S.g, q, buf + (2 + s)*n, d, s, S.E};
                    Scope sc(KC_UNPACK); sc.done(S.L->halo_unpack(h, S.st));
                }
                CK(cudaStreamSynchronize(S.st));
                CK(cudaFree(buf));
            } else if ((rc = physical_bc_dir(q, d))) {
                return rc;
            }
        }
    }
    CK(cudaStreamSynchronize(S.st));
    CK(cudaGetLastError());
    return 0;
}

}  // extern "C"
