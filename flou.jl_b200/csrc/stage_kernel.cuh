// One fused kernel per RK stage:  k = rhs(u_in);  tmp = A*tmp + dt*k;  u_out = u_in + B*tmp.
//
// It replaces, in a single pass over the state, the reference's seven sweeps
// (src/FlouSpatial/Equations/Hyperbolic.jl:31-69):
//   project2faces!          Interfaces.jl:51-109      -> traces built on the fly (own element
//                                                        from shared memory, neighbour from L2)
//   volume_contribution!    OpDivergence.jl:105-160 (strong), :184-282 (split)
//   applyBCs!               Interfaces.jl:25-49       -> evaluated inside the face task
//   interface_fluxes!       Interfaces.jl:111-136     -> recomputed by both neighbours with the
//                                                        same (master, slave) argument order,
//                                                        hence bitwise equal on both sides
//   surface_contribution!   OpDivergence.jl:42-100
//   apply_massmatrix!       MultielementDiscontinuous.jl:132-137 (true division by jac)
// plus OrdinaryDiffEq's LowStorageRK2N stage update (call site FlouTime.jl:34-38).
// `u` is ping-ponged (u_in != u_out) because neighbours read the same-stage state.
//
// Thread mapping: a CTA owns EPB consecutive elements; thread t <-> (element t / NPTS,
// node t % NPTS), node = ix + NP*iy + NP^2*iz (StdQuad.jl:46-50, StdHex.jl:48-54).  Face
// work is a flat task list (element, local face, face dof) strided over the CTA.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "physics.cuh"

#ifndef FLOU_L2_PREFETCH
#define FLOU_L2_PREFETCH 592   // groups ahead (= resident CTAs on a B200: 4 per SM x 148); 0 disables
#endif
#ifndef FLOU_TPB
#define FLOU_TPB 128      // target threads per CTA (measured: 128 beats 256 on B200)
#endif

namespace flou {

// 8-byte asynchronous global -> shared copy (LDGSTS); completion via cp_async_wait_all()
__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gmem_src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
}
// 16-byte copy that bypasses L1 (LDGSTS.BYPASS.128); both addresses 16-byte aligned
__device__ __forceinline__ void cp_async16(double *smem_dst, const double *gmem_src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
// 4-byte variant (one int of a connectivity record)
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gmem_src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__host__ __device__ constexpr int ipow_c(int b, int e) { return e <= 0 ? 1 : b * ipow_c(b, e - 1); }
// Stride (in doubles) between the flux blocks Fn[slot][v][i] of consecutive face slots: nv*nfp
// rounded up to an even count, so that every block starts on a 16-byte boundary and the element
// kernel can fetch it with ONE TMA bulk copy (cp.async.bulk needs 16-byte addresses and sizes).
__host__ __device__ constexpr int fn_block(int nv, int nfp) { return (nv * nfp + 1) & ~1; }

enum : int { MODE_RHS = 0, MODE_STAGE = 1, MODE_STAGE_FIRST = 2 };
enum : int { FK_INTERIOR = 0, FK_GHOST = 1, FK_BOUNDARY = 2 };

// conn.info bit layout
//   [0:3)  neighbour's local face (0-based)         [3:6) orientation
//   [6]    this element is the face's master        [7:9) kind (FK_*)
//   [9:..) boundary index (FK_BOUNDARY)
__host__ __device__ inline int conn_pack(int nbrface, int orient, int master, int kind, int bc)
{
    return (nbrface & 7) | ((orient & 7) << 3) | ((master & 1) << 6) | ((kind & 3) << 7) | (bc << 9);
}

struct Conn {
    int nbr;    // interior: local neighbour element; ghost: ghost slot; boundary: ordinal in bc_faces
    int info;
};

// One record per local face (two-kernel path).  Side A = master, side B = slave.
//   em / es : local element id, ghost slot, or (slave only) ordinal of the boundary face
//   info    : [0:3) master local face  [3:6) slave local face  [6:9) orientation
//             [9:11) master kind (0 local, 1 ghost)  [11:13) slave kind (0 local, 1 ghost, 2 boundary)
//             [13:..) boundary index
struct FaceRec {
    int em, es, info, pad;
};
__host__ __device__ inline int facerec_pack(int lfm, int lfs, int orient, int mkind, int skind, int bc)
{
    return (lfm & 7) | ((lfs & 7) << 3) | ((orient & 7) << 6) | ((mkind & 3) << 9) | ((skind & 3) << 11) | (bc << 13);
}

struct KParams {
    // operators, column-major NP x NP : Dvol = Ds (strong) or Dsharp (split)
    double Dvol[64];
    double lm[8], lp[8], dgl[8], dgr[8];
    int colloc;                 // GLL: l(-1) = e_1, l(+1) = e_np up to round-off
    FluxParams fp;
    // Cartesian geometry (PhysicalRegions.jl:370-408, 541-696)
    double cjac;                // prod(dx)/2^nd
    double crjac;               // 1/cjac
    double cmet[3];             // metric diagonal  Ja^d_d
    double cfjac[3];            // face jac by direction
    double rcmet[3];            // 1/cmet (line kernel: lift pre-division in the folded split form)
    int diag_mask;              // bit j: |Dvol[j,j]| is not round-off (line kernel, split form)
    // HybridDivOperator (VOL_HYBRID): Dvol = std.D, 1-D quadrature weights, op.blend, op.tpflux
    double w1d[8];
    double blend;
    int tpflux;
    // sub-grid frames / Jacobians of unstructured elements (hybrid operator, Gauss-node split form):
    // [(((e*ND + d)*NFP + k)*(NP+1) + ii)*3*ND + r*ND + c] and [((e*ND + d)*NFP + k)*(NP+1) + ii], e local
    const double *sub_frames;
    const double *sub_jac;
    int prefetch_groups;        // line kernel: CTAs resident on the device (L2 prefetch distance)
    // general geometry, device SoA
    const double *jac;          // [dof]
    const double *metric;       // [(c + nd*d)][dof]  (plane-major)
    const double *fjac;         // [lface*NFP + i]
    const double *frames;       // [(r*nd + c)][lface*NFP + i], r = n,t,b
    const int *faceid;          // [e*2nd + lf] -> local face slot
    int64_t nfacedofs;          // plane stride of frames
    // connectivity
    const Conn *conn;           // [e*2nd + lf]
    // boundary conditions
    const int *bc_kind;
    const double *bc_state;     // [ib*nv + v]
    const double *bc_table;     // [(m*NFP + i)*nv + v]
    // traces of the two x-faces of every element (the only faces whose nodes are strided in
    // u), one contiguous block per (element, side): [(e*2 + side)*nv + v]*NFP + k
    const double *tr_in;        // traces of u_in  (read by the neighbours)
    double *tr_out;             // traces of u_out (written together with u_out)
    // Gauss nodes only: interpolated traces of the remaining faces, [(e*2nd + lf)*nv + v]*NFP + k
    const double *tr_hi;
    // two-kernel path: face table, per (element, face) flux slot, flux array
    const FaceRec *faces;       // [local face slot]
    const int2 *econn;          // [e*2nd + lf] = {flux slot, info: bit0 master, bits 1..3 orientation}
    double *Fn;                 // [(slot*nv + v)*NFP + i]: master-outward normal flux * face jac
    int face_first, face_count;
    int face_reverse;           // face kernel walks the slots from the last to the first
    int split_faces;            // 1: stage kernel reads Fn instead of evaluating Riemann fluxes
    // halo
    const double *ghost;        // [(slot*nv + v)*NFP + k] in the sender's face-dof order
    // state
    const double *u_in;
    double *u_out;
    double *tmp;
    double *k_out;
    // apply_sourceterm! (MultielementDiscontinuous.jl:139-146): S(x, t[, Q]) tabulated per node,
    // added to dQ after the mass matrix; nullptr = the reference's default no-op source
    const double *source;       // [dof + ndof*v]
    int64_t ndof;               // local dofs = plane stride of the state
    int elem_first, elem_count;
    const int *elem_list;       // optional indirection
    int mode;
    double rkA, rkB, dt;
    int *status;
};

// ------------------------------------------------------------------ index helpers
template <int ND, int NP>
__device__ __forceinline__ void line_of(int d, int k, int &base, int &stride)
{   // tpdofs order: StdQuad.jl:116-124, StdHex.jl:135-145
    if (ND == 1) { base = 0; stride = 1; }
    else if (ND == 2) {
        if (d == 0) { base = NP * k; stride = 1; } else { base = k; stride = NP; }
    } else {
        if (d == 0) { base = NP * k; stride = 1; }
        else if (d == 1) { base = (k % NP) + NP * NP * (k / NP); stride = NP; }
        else { base = k; stride = NP * NP; }
    }
}

// face-dof (line number) of `node` in direction d, and its position along the line
template <int ND, int NP>
__device__ __forceinline__ void node_line(int node, int d, int &k, int &ii)
{
    if (ND == 1) { k = 0; ii = node; }
    else if (ND == 2) {
        const int ix = node % NP, iy = node / NP;
        if (d == 0) { k = iy; ii = ix; } else { k = ix; ii = iy; }
    } else {
        const int ix = node % NP, iy = (node / NP) % NP, iz = node / (NP * NP);
        if (d == 0) { k = iy + NP * iz; ii = ix; }
        else if (d == 1) { k = ix + NP * iz; ii = iy; }
        else { k = ix + NP * iy; ii = iz; }
    }
}

// master2slave / slave2master (0-based): StdSegment.jl:145-167, StdQuad.jl:139-181
template <int ND, int NP>
__device__ __forceinline__ int master2slave(int i, int o)
{
    if (ND <= 1 || o == 0) return i;
    if (ND == 2) return NP - 1 - i;
    const int m1 = i % NP, m2 = i / NP;     // 0-based; n - m + 1 (1-based) == NP-1-m (0-based)
    int a, b;
    switch (o) {
    case 1: a = m2; b = NP - 1 - m1; break;
    case 2: a = NP - 1 - m1; b = NP - 1 - m2; break;
    case 3: a = NP - 1 - m2; b = m1; break;
    case 4: a = m2; b = m1; break;
    case 5: a = NP - 1 - m1; b = m2; break;
    case 6: a = NP - 1 - m2; b = NP - 1 - m1; break;
    default: a = m1; b = NP - 1 - m2; break;
    }
    return a + NP * b;
}

template <int ND, int NP>
__device__ __forceinline__ int slave2master(int i, int o)
{
    if (ND <= 1 || o == 0) return i;
    if (ND == 2) return NP - 1 - i;
    const int s1 = i % NP, s2 = i / NP;
    int a, b;
    switch (o) {
    case 1: a = NP - 1 - s2; b = s1; break;
    case 2: a = NP - 1 - s1; b = NP - 1 - s2; break;
    case 3: a = s2; b = NP - 1 - s1; break;
    case 4: a = s2; b = s1; break;
    case 5: a = NP - 1 - s1; b = s2; break;
    case 6: a = NP - 1 - s2; b = NP - 1 - s1; break;
    default: a = s1; b = NP - 1 - s2; break;
    }
    return a + NP * b;
}

// Cartesian face frame chosen by the master's element-local face position `pm` (0-based):
// PhysicalRegions.jl:541-696.  fr = n[ND], t[ND], b[ND].
template <int ND>
__device__ __forceinline__ void cart_frame(int pm, double *fr)
{
    const int dm = pm >> 1;
    const double s = (pm & 1) ? 1.0 : -1.0;
    // selects instead of dynamic indexing keep `fr` in registers
#pragma unroll
    for (int c = 0; c < ND; c++) fr[c] = (c == dm) ? s : 0.0;
    if (ND == 2) {
        // pos1: t=(0,-1)  pos2: t=(0,1)  pos3: t=(1,0)  pos4: t=(-1,0)
#pragma unroll
        for (int c = 0; c < ND; c++) fr[ND + c] = (c == 1 - dm) ? ((dm == 0) ? s : -s) : 0.0;
#pragma unroll
        for (int c = 0; c < ND; c++) fr[2 * ND + c] = 0.0;
    } else if (ND == 3) {
        const int tm = (dm == 2) ? 0 : dm + 1, bm = (dm == 0) ? 2 : dm - 1;
#pragma unroll
        for (int c = 0; c < ND; c++) fr[ND + c] = (c == tm) ? s : 0.0;
#pragma unroll
        for (int c = 0; c < ND; c++) fr[2 * ND + c] = (c == bm) ? 1.0 : 0.0;
    } else {
        fr[1] = 0.0; fr[2] = 0.0;
    }
}

template <int ND, int EQ>
__device__ __forceinline__ void rotate2face(const double *Q, const double *fr, double *R)
{
    if (EQ == EQ_ADV) { R[0] = Q[0]; return; }
    R[0] = Q[0];
    if (ND == 1) { R[1] = Q[1] * fr[0]; }
    else {
#pragma unroll
        for (int r = 0; r < ND; r++) {
            double s = Q[1] * fr[r * ND];
#pragma unroll
            for (int c = 1; c < ND; c++) s += Q[1 + c] * fr[r * ND + c];
            R[1 + r] = s;
        }
    }
    R[ND + 1] = Q[ND + 1];
}

template <int ND, int EQ>
__device__ __forceinline__ void rotate2phys(const double *R, const double *fr, double *Q)
{
    if (EQ == EQ_ADV) { Q[0] = R[0]; return; }
    Q[0] = R[0];
    if (ND == 1) { Q[1] = R[1] * fr[0]; }
    else {
#pragma unroll
        for (int c = 0; c < ND; c++) {
            double s = R[1] * fr[c];
#pragma unroll
            for (int r = 1; r < ND; r++) s += R[1 + r] * fr[r * ND + c];
            Q[1 + c] = s;
        }
    }
    Q[ND + 1] = R[ND + 1];
}

// Cartesian frames are signed axis permutations (PhysicalRegions.jl:541-696): rotating is
// a component select and a sign, which reproduces the reference's dot products with
// (+-1, 0, 0)-type vectors exactly.
template <int ND>
__device__ __forceinline__ double pick(const double *m, int c)
{   // m[c] without dynamic register indexing
    if (ND == 1) return m[0];
    if (ND == 2) return c == 0 ? m[0] : m[1];
    return c == 0 ? m[0] : (c == 1 ? m[1] : m[2]);
}

template <int ND, int EQ>
__device__ __forceinline__ void rotate2face_cart(const double *Q, int dm, double s, double *R)
{
    R[0] = Q[0];
    if (EQ == EQ_ADV) return;
    R[1] = s * pick<ND>(Q + 1, dm);
    if (ND == 2) R[2] = ((dm == 0) ? s : -s) * pick<ND>(Q + 1, 1 - dm);
    if (ND == 3) {
        const int tm = (dm == 2) ? 0 : dm + 1, bm = (dm == 0) ? 2 : dm - 1;
        R[2] = s * pick<ND>(Q + 1, tm);
        R[3] = pick<ND>(Q + 1, bm);
    }
    R[ND + 1] = Q[ND + 1];
}

template <int ND, int EQ>
__device__ __forceinline__ void rotate2phys_cart(const double *R, int dm, double s, double *Q)
{
    Q[0] = R[0];
    if (EQ == EQ_ADV) return;
    if (ND == 1) Q[1] = s * R[1];
    if (ND == 2) {
        const double n = s * R[1], t = ((dm == 0) ? s : -s) * R[2];
        Q[1] = dm == 0 ? n : t;
        Q[2] = dm == 0 ? t : n;
    }
    if (ND == 3) {
        const int tm = (dm == 2) ? 0 : dm + 1;
        const double n = s * R[1], t = s * R[2], b = R[3];
#pragma unroll
        for (int c = 0; c < 3; c++) Q[1 + c] = (c == dm) ? n : (c == tm ? t : b);
    }
    Q[ND + 1] = R[ND + 1];
}

// ------------------------------------------------------------------ kernel configuration
template <int ND_, int NP_, int EQ_, int VOL_, bool CART_>
struct KCfg {
    static constexpr int ND = ND_, NP = NP_, EQ = EQ_, VOL = VOL_;
    static constexpr bool CART = CART_;
    static constexpr bool NB = false;                  // line kernels: see LCfg
    static constexpr int NV = (EQ == EQ_ADV) ? 1 : ND + 2;
    static constexpr int NPTS = ipow_c(NP, ND);
    static constexpr int NFP = ipow_c(NP, ND - 1);
    static constexpr int NFACES = 2 * ND;
    static constexpr int NFT = NFACES * NFP;           // face tasks per element
    static constexpr int EPB = (NPTS >= FLOU_TPB) ? 1 : FLOU_TPB / NPTS;   // elements per group
    static constexpr int THREADS = ((EPB * NPTS + 31) / 32) * 32;
    static constexpr int TPT = (EPB * NFT + THREADS - 1) / THREADS;         // face tasks per thread
    static constexpr bool SPLIT = (VOL != VOL_STRONG);
    // split form: node pair (i, i+s) is evaluated once, by node i, for s = 1..NP/2; pairs
    // with s <= (NP-1)/2 are handed to the partner through the exchange buffer
    static constexpr int ROUNDS = NP / 2;
    static constexpr int XROUNDS = (NP - 1) / 2;
    // shared-memory layout (doubles), per element slot
    static constexpr int QOFF = 0;                                           // sQ[NV][NPTS]
    static constexpr int TOFF = QOFF + NV * NPTS;                            // sT[NV][NPTS]
    static constexpr int AOFF = TOFF + NV * NPTS;
    // Euler split form: Chandrasekhar keeps (v/2, |v|^2, beta) per node, StdAverage (v, p)
    static constexpr int NAUX = (EQ == EQ_EULER && SPLIT) ? (VOL == VOL_SPLIT_CHA ? ND + 2 : ND + 1) : 0;
    static constexpr int NFT_VOL = SPLIT ? 0 : ND * NV;                      // contravariant fluxes
    static constexpr int NMET = (CART || !SPLIT) ? 0 : ND * ND;
    static constexpr int NXCH = SPLIT ? 2 * XROUNDS * NV : 0;               // two exchange buffers
    // parked volume accumulators reuse the exchange buffer that is idle after the last
    // direction (buffer ND&1) when there is one
    static constexpr bool ACC_ALIAS = SPLIT && XROUNDS >= 1;
    static constexpr int NACC = ACC_ALIAS ? 0 : NV;
    static constexpr int FTOFF = AOFF + NAUX * NPTS;
    static constexpr int MOFF = FTOFF + NFT_VOL * NPTS;
    static constexpr int XOFF = MOFF + NMET * NPTS;
    static constexpr int ACCOFF = ACC_ALIAS ? XOFF + (ND & 1) * (XROUNDS * NV * NPTS)
                                            : XOFF + NXCH * NPTS;
    // neighbour traces are prefetched into the slots the face fluxes later overwrite
    static constexpr int FOFF = XOFF + (NXCH + NACC) * NPTS;
    static constexpr int PER_ELEM = FOFF + NFACES * NV * NFP;
    static constexpr int OPS = NP * NP + 4 * NP + EPB * NFACES;
    static constexpr size_t SMEM_BYTES = sizeof(double) * (size_t)(OPS + EPB * PER_ELEM);
    static constexpr int MIN_BLOCKS_WANTED =
#ifdef FLOU_MIN_BLOCKS
        FLOU_MIN_BLOCKS;
#else
        (THREADS > 256) ? 1 : (THREADS > 128 ? 2 : 4);
#endif
    static constexpr int SMEM_BLOCKS = (int)((227 * 1024) / (SMEM_BYTES + 1024));
    static constexpr int MIN_BLOCKS =
        SMEM_BLOCKS < 1 ? 1 : (SMEM_BLOCKS < MIN_BLOCKS_WANTED ? SMEM_BLOCKS : MIN_BLOCKS_WANTED);
    // element kernel of the two-kernel path (no Riemann solver inside: fewer registers)
    static constexpr int MIN_BLOCKS_E_WANTED =
#ifdef FLOU_MIN_BLOCKS_E
        FLOU_MIN_BLOCKS_E;
#else
        (THREADS > 256) ? 1 : (THREADS > 128 ? 2 : 5);
#endif
    static constexpr int MIN_BLOCKS_E =
        SMEM_BLOCKS < 1 ? 1 : (SMEM_BLOCKS < MIN_BLOCKS_E_WANTED ? SMEM_BLOCKS : MIN_BLOCKS_E_WANTED);
};

template <int N>
__device__ __forceinline__ Conn pick_conn(const Conn *cn, int j)
{   // cn[j] for a register array (no dynamic indexing)
    Conn c = cn[0];
#pragma unroll
    for (int q = 1; q < N; q++) if (j == q) c = cn[q];
    return c;
}

// One CTA = one group of EPB consecutive elements.  Phase 0 starts the asynchronous copies
// (cp.async) of `tmp` and of the neighbours' face traces -- contiguous blocks of the trace
// array, so they coalesce -- and they are only waited for right before the face / update
// phases, i.e. behind the volume work.
// SPLITF = true: the Riemann fluxes come from face_flux_kernel through P.Fn (each face
// evaluated once); phase 3 disappears and phase 0 copies the six flux blocks instead of the
// neighbours' traces.  SPLITF = false: the fully fused single-kernel stage.
// Programmatic dependent launch (FLOU_B200_PDL=1: the stage kernels are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization): the next kernel of the stream may be set up
// and its CTAs placed while this grid is still running; it then waits here until every grid it
// depends on has completed and flushed.  Both instructions are no-ops in a normal launch.
__device__ __forceinline__ void pdl_prologue()
{
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

template <class C, bool SPLITF>
__global__ void __launch_bounds__(C::THREADS, SPLITF ? C::MIN_BLOCKS_E : C::MIN_BLOCKS)
stage_kernel(const __grid_constant__ KParams P)
{
    pdl_prologue();
    constexpr int ND = C::ND, NP = C::NP, EQ = C::EQ, VOL = C::VOL, NV = C::NV;
    constexpr int NPTS = C::NPTS, NFP = C::NFP, NFACES = C::NFACES, NFT = C::NFT, EPB = C::EPB;
    constexpr int TPT = C::TPT;
    constexpr bool CART = C::CART, SPLIT = C::SPLIT;

    extern __shared__ double smem[];
    double *sD = smem;                       // [ii + NP*jj]
    double *sLm = sD + NP * NP, *sLp = sLm + NP, *sGl = sLp + NP, *sGr = sGl + NP;
    double *sSign = sGr + NP;                // [el][lf]: +1 master, -1 slave (two-kernel path)
    double *sElem = sSign + EPB * NFACES;

    const int tid = threadIdx.x;
    for (int i = tid; i < NP * NP; i += C::THREADS) sD[i] = P.Dvol[i];
    if (tid < NP) { sLm[tid] = P.lm[tid]; sLp[tid] = P.lp[tid]; sGl[tid] = P.dgl[tid]; sGr[tid] = P.dgr[tid]; }

    const int el = tid / NPTS, node = tid - el * NPTS;
    const bool node_thread = el < EPB;
    double *sMine = sElem + (size_t)(node_thread ? el : 0) * C::PER_ELEM;
    double *sQ = sMine + C::QOFF;
    double *sT = sMine + C::TOFF;
    double *sA = sMine + C::AOFF;
    double *sFt = sMine + C::FTOFF;
    double *sM = sMine + C::MOFF;
    double *sX = sMine + C::XOFF;
    double *sAcc = sMine + C::ACCOFF;
    const int64_t ndof = P.ndof;
    const bool need_tmp = (P.mode == MODE_STAGE);

    const int g = blockIdx.x;
    const int nact = min(EPB, P.elem_count - g * EPB);
    const bool active = node_thread && el < nact;
    auto elem_of = [&](int idx) { return P.elem_list ? P.elem_list[idx] : P.elem_first + idx; };
    const int e = active ? elem_of(g * EPB + el) : 0;
    const int64_t dof = (int64_t)e * NPTS + node;

    // ---------------- phase 0: connectivity of this thread's face tasks; start the copies
    Conn cn_cur[TPT];
#pragma unroll
    for (int j = 0; j < TPT; j++) {
        const int task = tid + j * C::THREADS;
        cn_cur[j].nbr = 0; cn_cur[j].info = conn_pack(0, 0, 1, FK_BOUNDARY, 0);
        if (task < nact * NFT) {
            const int tel = task / NFT, r = task - tel * NFT;
            const int lf = r / NFP, k = r - lf * NFP;
            const int te = elem_of(g * EPB + tel);
            if (SPLITF) {
                // flux block of this face, permuted into my face-dof order while copying
                const int2 ec = __ldg(P.econn + ((int64_t)te * NFACES + lf));
                const bool master = ec.y & 1;
                const int i = master ? k : slave2master<ND, NP>(k, (ec.y >> 1) & 7);
                double *dst = sElem + (size_t)tel * C::PER_ELEM + C::FOFF + lf * NV * NFP + k;
                const double *src = P.Fn + (int64_t)ec.x * fn_block(NV, NFP) + i;
#pragma unroll
                for (int v = 0; v < NV; v++) cp_async8(dst + v * NFP, src + v * NFP);
                if (k == 0) sSign[tel * NFACES + lf] = master ? 1.0 : -1.0;
                continue;
            }
            const int2 c = __ldg(reinterpret_cast<const int2 *>(P.conn) + ((int64_t)te * NFACES + lf));
            cn_cur[j].nbr = c.x; cn_cur[j].info = c.y;
            const int kind = (c.y >> 7) & 3;
#ifdef FLOU_EXPERIMENT_SKIP_TRACE_READ
            if (kind != FK_BOUNDARY && P.elem_count < 0) {
#else
            if (kind != FK_BOUNDARY) {
#endif
                const int nlf = c.y & 7, orient = (c.y >> 3) & 7;
                const bool master = (c.y >> 6) & 1;
                const int kn = master ? master2slave<ND, NP>(k, orient) : slave2master<ND, NP>(k, orient);
                double *dst = sElem + (size_t)tel * C::PER_ELEM + C::FOFF + lf * NV * NFP + k;
                if (kind == FK_GHOST) {
                    const double *src = P.ghost + (int64_t)c.x * (NV * NFP) + kn;
#pragma unroll
                    for (int v = 0; v < NV; v++) cp_async8(dst + v * NFP, src + v * NFP);
                } else if (nlf < 2) {
                    // the neighbour's x-faces are strided in u: read its trace block instead
                    const double *src = P.tr_in + ((int64_t)c.x * 2 + nlf) * (NV * NFP) + kn;
#pragma unroll
                    for (int v = 0; v < NV; v++) cp_async8(dst + v * NFP, src + v * NFP);
                } else if (P.colloc) {
                    // y-/z-faces of the neighbour are (runs of) contiguous nodes of u
                    int nb, ns;
                    line_of<ND, NP>(nlf >> 1, kn, nb, ns);
                    const double *src = P.u_in + (int64_t)c.x * NPTS + nb + ((nlf & 1) ? (NP - 1) * ns : 0);
#pragma unroll
                    for (int v = 0; v < NV; v++) cp_async8(dst + v * NFP, src + ndof * v);
                } else {
                    const double *src = P.tr_hi + ((int64_t)c.x * NFACES + nlf) * (NV * NFP) + kn;
#pragma unroll
                    for (int v = 0; v < NV; v++) cp_async8(dst + v * NFP, src + v * NFP);
                }
            }
        }
    }
    if (need_tmp && active) {
#pragma unroll
        for (int v = 0; v < NV; v++) cp_async8(sT + v * NPTS + node, P.tmp + dof + ndof * v);
    }
    cp_async_commit();
#if FLOU_L2_PREFETCH > 0
    // warm L2 for the group that will occupy this SM slot one wave later: its first action is
    // a dependent load of its own state, which then costs an L2 hit instead of a DRAM access
    if (active && (node & 15) == 0) {
        const int idx = (g + FLOU_L2_PREFETCH) * EPB + el;
        if (idx < P.elem_count) {
            const int64_t pd = (int64_t)elem_of(idx) * NPTS + node;
#pragma unroll
            for (int v = 0; v < NV; v++) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(P.u_in + pd + ndof * v));
                if (need_tmp) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.tmp + pd + ndof * v));
            }
        }
    }
#endif

    {
        // ---------------- phase 1: node primitives, contravariant fluxes
        double Q[NV];
        double met[CART ? 1 : ND * ND];
        double vi[ND], hvi[ND], pi = 0.0, bi = 0.0, qi = 0.0;
        if (active) {
#pragma unroll
            for (int v = 0; v < NV; v++) { Q[v] = __ldg(P.u_in + dof + ndof * v); sQ[v * NPTS + node] = Q[v]; }
            if (!CART) {
#pragma unroll
                for (int m = 0; m < ND * ND; m++) {
                    met[m] = __ldg(P.metric + dof + ndof * m);
                    if (SPLIT) sM[m * NPTS + node] = met[m];
                }
            }
            if (EQ == EQ_EULER) {
                NodeAux<ND> A;
                node_aux<ND>(Q, P.fp.gamma, A);
                if (!(Q[0] > 0.0) || !(A.p > 0.0)) atomicOr(P.status, 1);
#pragma unroll
                for (int d = 0; d < ND; d++) vi[d] = A.vel[d];
                pi = A.p; bi = A.beta;
                if (VOL == VOL_SPLIT_CHA) {
                    double q = 0.0;
#pragma unroll
                    for (int d = 0; d < ND; d++) {
                        q = fma(A.vel[d], A.vel[d], q);
                        hvi[d] = 0.5 * A.vel[d];
                        sA[d * NPTS + node] = hvi[d];
                    }
                    qi = q;
                    sA[ND * NPTS + node] = q;
                    sA[(ND + 1) * NPTS + node] = A.beta;
                } else if (VOL == VOL_SPLIT_STD) {
#pragma unroll
                    for (int d = 0; d < ND; d++) sA[d * NPTS + node] = A.vel[d];
                    sA[ND * NPTS + node] = A.p;
                } else {
#pragma unroll
                    for (int d = 0; d < ND; d++) {
                        double Fc[NV], Ft[NV];
#pragma unroll
                        for (int v = 0; v < NV; v++) Ft[v] = 0.0;
#pragma unroll
                        for (int c = 0; c < ND; c++) {
                            if (CART && c != d) continue;
                            const double m = CART ? P.cmet[d] : met[c + ND * d];
                            euler_flux_dir<ND>(Q, A.vel, A.p, c, Fc);
#pragma unroll
                            for (int v = 0; v < NV; v++) Ft[v] += Fc[v] * m;
                        }
#pragma unroll
                        for (int v = 0; v < NV; v++) sFt[(d * NV + v) * NPTS + node] = Ft[v];
                    }
                }
            } else if (!SPLIT) {
#pragma unroll
                for (int d = 0; d < ND; d++) {
                    double an = 0.0;
#pragma unroll
                    for (int c = 0; c < ND; c++)
                        an += P.fp.a[c] * (CART ? (c == d ? P.cmet[d] : 0.0) : met[c + ND * d]);
                    sFt[d * NPTS + node] = an * Q[0];
                }
            }
        }
        __syncthreads();     // every node's state and primitives are now visible

        // ---------------- phase 2: volume term
        double acc[NV];
#pragma unroll
        for (int v = 0; v < NV; v++) acc[v] = 0.0;
        if (!SPLIT) {
            if (active) {
#pragma unroll
                for (int d = 0; d < ND; d++) {
                    int k, ii, base, stride;
                    node_line<ND, NP>(node, d, k, ii);
                    line_of<ND, NP>(d, k, base, stride);
#pragma unroll
                    for (int jj = 0; jj < NP; jj++) {
                        const double dij = sD[ii + NP * jj];
                        const int l = base + jj * stride;
#pragma unroll
                        for (int v = 0; v < NV; v++) acc[v] -= dij * sFt[(d * NV + v) * NPTS + l];
                    }
                }
            }
        }
#ifdef FLOU_EXPERIMENT_SKIP_VOLUME
        else if (P.elem_count < 0) {
#else
        else {
#endif
            // split form  dQ_i -= sum_j D#[i,j] F#(i,j)   (OpDivergence.jl:248-282).  F# is
            // symmetric: node i evaluates the pairs (i, i+s mod NP), s = 1..NP/2, keeps them
            // for itself and leaves those with s <= (NP-1)/2 in shared memory for node i+s.
#pragma unroll
            for (int d = 0; d < ND; d++) {
                int k, ii, base, stride;
                node_line<ND, NP>(node, d, k, ii);
                line_of<ND, NP>(d, k, base, stride);
                double *xb = sX + (size_t)(d & 1) * (C::XROUNDS * NV * NPTS);
                if (active) {
                    double ni[ND];
#pragma unroll
                    for (int c = 0; c < ND; c++) ni[c] = CART ? (c == d ? P.cmet[d] : 0.0) : met[c + ND * d];
                    // diagonal entry: the node's own contravariant flux (OpDivergence.jl:252)
                    {
                        double F[NV];
                        if (EQ == EQ_EULER) {
#pragma unroll
                            for (int v = 0; v < NV; v++) F[v] = 0.0;
#pragma unroll
                            for (int c = 0; c < ND; c++) {
                                if (CART && c != d) continue;
                                double Fc[NV];
                                euler_flux_dir<ND>(Q, vi, pi, c, Fc);
#pragma unroll
                                for (int v = 0; v < NV; v++) F[v] += Fc[v] * ni[c];
                            }
                        } else {
                            double an = 0.0;
#pragma unroll
                            for (int c = 0; c < ND; c++) an += P.fp.a[c] * ni[c];
                            F[0] = an * Q[0];
                        }
                        const double dii = sD[ii + NP * ii];
#pragma unroll
                        for (int v = 0; v < NV; v++) acc[v] -= dii * F[v];
                    }
#pragma unroll
                    for (int s = 1; s <= C::ROUNDS; s++) {
                        int jj = ii + s;
                        if (jj >= NP) jj -= NP;
                        const int l = base + jj * stride;
                        const double dij = sD[ii + NP * jj];
                        double n[ND], F[NV];
#pragma unroll
                        for (int c = 0; c < ND; c++)
                            n[c] = CART ? ni[c] : 0.5 * (ni[c] + sM[(c + ND * d) * NPTS + l]);
                        if (EQ == EQ_EULER) {
                            double vl[ND];
#pragma unroll
                            for (int c = 0; c < ND; c++) vl[c] = sA[c * NPTS + l];
                            if (VOL == VOL_SPLIT_CHA) {
                                tp_chandrasekhar<ND>(Q[0], hvi, qi, bi, sQ[l], vl, sA[ND * NPTS + l],
                                                     sA[(ND + 1) * NPTS + l], P.fp.inv_gm1, n, F);
                            } else {
                                double Ql[NV];
#pragma unroll
                                for (int v = 0; v < NV; v++) Ql[v] = sQ[v * NPTS + l];
                                tp_stdavg<ND>(Q, vi, pi, Ql, vl, sA[ND * NPTS + l], n, F);
                            }
                        } else {
                            double an = 0.0;
#pragma unroll
                            for (int c = 0; c < ND; c++) an += P.fp.a[c] * n[c];
                            F[0] = an * (Q[0] + sQ[l]) * 0.5;
                        }
#pragma unroll
                        for (int v = 0; v < NV; v++) acc[v] -= dij * F[v];
                        if (s <= C::XROUNDS) {
#pragma unroll
                            for (int v = 0; v < NV; v++) xb[((s - 1) * NV + v) * NPTS + node] = F[v];
                        }
                    }
                }
                if (C::XROUNDS > 0) {
                    __syncthreads();
                    if (active) {
#pragma unroll
                        for (int s = 1; s <= C::XROUNDS; s++) {
                            int jp = ii - s;
                            if (jp < 0) jp += NP;
                            const int lp = base + jp * stride;
                            const double dij = sD[ii + NP * jp];
#pragma unroll
                            for (int v = 0; v < NV; v++) acc[v] -= dij * xb[((s - 1) * NV + v) * NPTS + lp];
                        }
                    }
                }
            }
        }
        // fused kernel only: park the volume accumulators, the face phase needs the registers
        if (!SPLITF && active) {
#pragma unroll
            for (int v = 0; v < NV; v++) sAcc[v * NPTS + node] = acc[v];
        }

        // ---------------- phase 3: face tasks (traces, BCs, Riemann flux) -> shared memory
        cp_async_wait<0>();                 // tmp and the neighbours' traces have landed
#ifdef FLOU_EXPERIMENT_SKIP_FACES
        if (P.elem_count < 0)
#endif
        if (!SPLITF)
#pragma unroll 1
        for (int j = 0; j < TPT; j++) {
            const int task = tid + j * C::THREADS;
            if (task >= nact * NFT) break;
            const int tel = task / NFT, r = task - tel * NFT;
            const int lf = r / NFP, k = r - lf * NFP;
            const int d = lf >> 1, side = lf & 1;
            const double *tQ = sElem + (size_t)tel * C::PER_ELEM + C::QOFF;
            double *tF = sElem + (size_t)tel * C::PER_ELEM + C::FOFF;
            const Conn cn = pick_conn<TPT>(cn_cur, j);

            // own trace
            double Qown[NV];
            int base, stride;
            line_of<ND, NP>(d, k, base, stride);
            if (P.colloc) {
                const int n0 = base + (side ? (NP - 1) * stride : 0);
#pragma unroll
                for (int v = 0; v < NV; v++) Qown[v] = tQ[v * NPTS + n0];
            } else {
                const double *lv = side ? sLp : sLm;
#pragma unroll
                for (int v = 0; v < NV; v++) {
                    double s = 0.0;
#pragma unroll
                    for (int ii = 0; ii < NP; ii++) s += lv[ii] * tQ[v * NPTS + base + ii * stride];
                    Qown[v] = s;
                }
            }

            const int kind = (cn.info >> 7) & 3;
            const int nlf = cn.info & 7, orient = (cn.info >> 3) & 7;
            const bool master = (cn.info >> 6) & 1;
            // face dof seen from the other side / from the master
            const int kn = master ? master2slave<ND, NP>(k, orient) : slave2master<ND, NP>(k, orient);
            const int im = master ? k : kn;          // master face dof: frame and jac index

            double fr[CART ? 1 : 3 * ND], fj;
            int dm = 0;
            double sn_ = 1.0;                      // Cartesian: normal = sn_ * e_dm
            if (CART) {
                const int pm = master ? lf : nlf;
                dm = pm >> 1;
                sn_ = (pm & 1) ? 1.0 : -1.0;
                fj = dm == 0 ? P.cfjac[0] : (dm == 1 ? P.cfjac[1] : P.cfjac[2]);
            } else {
                const int te = elem_of(g * EPB + tel);
                const int64_t fi = (int64_t)P.faceid[(int64_t)te * NFACES + lf] * NFP + im;
#pragma unroll
                for (int c = 0; c < 3 * ND; c++)
                    fr[c] = (c < ND * ND || ND == 3) ? __ldg(P.frames + fi + P.nfacedofs * c) : 0.0;
                fj = __ldg(P.fjac + fi);
            }

            double Qnb[NV];
            if (kind != FK_BOUNDARY) {
#pragma unroll
                for (int v = 0; v < NV; v++) Qnb[v] = tF[(lf * NV + v) * NFP + k];
            } else {
                // boundary face: exterior state from the BC functor (Interfaces.jl:44-48)
                const int ib = cn.info >> 9;
                const int bk = P.bc_kind[ib];
                if (bk == FLOU_B200_BC_INFLOW) {
#pragma unroll
                    for (int v = 0; v < NV; v++) Qnb[v] = P.bc_state[ib * NV + v];
                } else if (bk == FLOU_B200_BC_OUTFLOW) {
#pragma unroll
                    for (int v = 0; v < NV; v++) Qnb[v] = Qown[v];
                } else if (bk == FLOU_B200_BC_SLIP) {
                    double R[NV];
                    if (CART) rotate2face_cart<ND, EQ>(Qown, dm, sn_, R);
                    else rotate2face<ND, EQ>(Qown, fr, R);
                    if (NV > 1) R[NV > 1 ? 1 : 0] = -R[NV > 1 ? 1 : 0];
                    if (CART) rotate2phys_cart<ND, EQ>(R, dm, sn_, Qnb);
                    else rotate2phys<ND, EQ>(R, fr, Qnb);
                } else {
#pragma unroll
                    for (int v = 0; v < NV; v++) Qnb[v] = P.bc_table[((int64_t)cn.nbr * NFP + k) * NV + v];
                }
            }

            // Riemann flux with (left = master, right = slave) exactly like the reference
            double Ql[NV], Qr[NV], Fn[NV], Fp[NV];
            {
                double Ro[NV], Rn[NV];
                if (CART) {
                    rotate2face_cart<ND, EQ>(Qown, dm, sn_, Ro);
                    rotate2face_cart<ND, EQ>(Qnb, dm, sn_, Rn);
                } else {
                    rotate2face<ND, EQ>(Qown, fr, Ro);
                    rotate2face<ND, EQ>(Qnb, fr, Rn);
                }
#pragma unroll
                for (int v = 0; v < NV; v++) { Ql[v] = master ? Ro[v] : Rn[v]; Qr[v] = master ? Rn[v] : Ro[v]; }
            }
            if (EQ == EQ_EULER) {
                euler_numflux<ND>(P.fp, Ql, Qr, Fn);
            } else {
                double an = 0.0;
                if (CART) an = sn_ * pick<ND>(P.fp.a, dm);
                else {
#pragma unroll
                    for (int c = 0; c < ND; c++) an += P.fp.a[c] * fr[c];
                }
                Fn[0] = an * (Ql[0] + Qr[0]) * 0.5;
                if (P.fp.numflux == FX_LXF) Fn[0] += fabs(an) * (Ql[0] - Qr[0]) * 0.5 * P.fp.intensity;
            }
            if (CART) rotate2phys_cart<ND, EQ>(Fn, dm, sn_, Fp);
            else rotate2phys<ND, EQ>(Fn, fr, Fp);
            const double sgn = master ? fj : -fj;
#pragma unroll
            for (int v = 0; v < NV; v++) tF[(lf * NV + v) * NFP + k] = Fp[v] * sgn;
        }
        __syncthreads();

        // ---------------- phase 4: lift, mass matrix, RK stage update
        if (active) {
            const double *sF = sMine + C::FOFF;
            if (!SPLITF) {
#pragma unroll
                for (int v = 0; v < NV; v++) acc[v] = sAcc[v * NPTS + node];
            }
#ifdef FLOU_EXPERIMENT_SKIP_LIFT
            if (P.elem_count < 0)
#endif
#pragma unroll
            for (int d = 0; d < ND; d++) {
                int k, ii;
                node_line<ND, NP>(node, d, k, ii);
                const double gl = sGl[ii], gr = sGr[ii];
                // collocated nodes: the lifting weights vanish away from the two end nodes
                if (gl != 0.0) {
                    const double w = SPLITF ? gl * sSign[el * NFACES + 2 * d] : gl;
#pragma unroll
                    for (int v = 0; v < NV; v++) acc[v] -= w * sF[((2 * d) * NV + v) * NFP + k];
                }
                if (gr != 0.0) {
                    const double w = SPLITF ? gr * sSign[el * NFACES + 2 * d + 1] : gr;
#pragma unroll
                    for (int v = 0; v < NV; v++) acc[v] -= w * sF[((2 * d + 1) * NV + v) * NFP + k];
                }
            }
            // mass matrix: dQ / jac (Diagonal ldiv!, MultielementDiscontinuous.jl:132-137)
            const double rjac = CART ? P.crjac : fast_rcp(__ldg(P.jac + dof));
            double src[NV];
#pragma unroll
            for (int v = 0; v < NV; v++) src[v] = P.source ? __ldg(P.source + dof + ndof * v) : 0.0;
            if (P.mode == MODE_RHS) {
#pragma unroll
                for (int v = 0; v < NV; v++) P.k_out[dof + ndof * v] = fma(acc[v], rjac, src[v]);
            } else {
#pragma unroll
                for (int v = 0; v < NV; v++) {
                    const double kv = fma(acc[v], rjac, src[v]);
                    double t;
                    if (P.mode == MODE_STAGE_FIRST) t = P.dt * kv;
                    else t = fma(P.dt, kv, P.rkA * sT[v * NPTS + node]);
                    P.tmp[dof + ndof * v] = t;
                    const double un = fma(P.rkB, t, sQ[v * NPTS + node]);
                    P.u_out[dof + ndof * v] = un;
                    acc[v] = un;
                }
                // traces of the new state for the next stage (collocated nodes: the boundary
                // node values; Gauss nodes are handled by emit_traces_kernel)
#ifdef FLOU_EXPERIMENT_SKIP_TRACE_WRITE
                if (P.colloc && P.elem_count < 0) {
#else
                if (P.colloc && P.tr_out) {
#endif
                    {
                        int k, ii;
                        node_line<ND, NP>(node, 0, k, ii);
                        if (ii == 0 || ii == NP - 1) {
                            double *dst = P.tr_out + ((int64_t)e * 2 + (ii == 0 ? 0 : 1)) * (NV * NFP) + k;
#pragma unroll
                            for (int v = 0; v < NV; v++) dst[v * NFP] = acc[v];
                        }
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------ halo trace emit
// Face traces of a state, in the owning side's face-dof order: out[(slot*NV + v)*NFP + k].
// list[slot] = local element * 2nd + local face (partition-boundary faces packed for
// ncclSend), or list == nullptr for the first `faces_per_elem` faces of every element.
template <int ND, int NP, int NV>
__global__ void emit_traces_kernel(const double *__restrict__ u, int64_t ndof,
                                   const int *__restrict__ list, int nslots, int faces_per_elem, int colloc,
                                   const double *__restrict__ lm, const double *__restrict__ lp,
                                   double *__restrict__ out)
{
    constexpr int NPTS = ipow_c(NP, ND), NFP = ipow_c(NP, ND - 1), NFACES = 2 * ND;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)nslots * NFP) return;
    const int slot = (int)(t / NFP), k = (int)(t - (int64_t)slot * NFP);
    // no list: the first `faces_per_elem` local faces of every element (2 = the x-faces)
    const int ef = list ? list[slot] : (slot / faces_per_elem) * NFACES + (slot % faces_per_elem);
    const int e = ef / NFACES, lf = ef - e * NFACES;
    const int d = lf >> 1, side = lf & 1;
    int base, stride;
    line_of<ND, NP>(d, k, base, stride);
    const int64_t b0 = (int64_t)e * NPTS + base;
#pragma unroll
    for (int v = 0; v < NV; v++) {
        double s;
        if (colloc) {
            s = u[b0 + (side ? (NP - 1) * stride : 0) + ndof * v];
        } else {
            s = 0.0;
            for (int ii = 0; ii < NP; ii++) s += (side ? lp[ii] : lm[ii]) * u[b0 + ii * stride + ndof * v];
        }
        out[((int64_t)slot * NV + v) * NFP + k] = s;
    }
}

}  // namespace flou
