// C-ABI implementation (include/flou_b200.h): handle, table upload, partition/halo
// bookkeeping, stage scheduling (streams, CUDA graph, NCCL halo exchange).
//
// Reference behaviour mirrored here (paths relative to the reference root):
//   MultielementDisc ctor        src/FlouSpatial/MultielementDiscontinuous.jl:29-92
//   rhs! orchestration           src/FlouSpatial/Equations/Hyperbolic.jl:31-69
//   timeintegrate                src/FlouTime/FlouTime.jl:34-54
//   Cartesian metrics            src/FlouSpatial/PhysicalRegions.jl:370-408, 541-696
// No CPU compute path exists in this file: without a usable CUDA device every entry point
// returns FLOU_B200_ECUDA.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <map>
#include <string>
#include <vector>

#include "flou_b200.h"
#include "launch.h"
#include "cfl_kernel.cuh"
#include "monitor_kernel.cuh"
#include "project_kernel.cuh"

using namespace flou;

namespace {

// FLOU_B200_SLAB = n > 0: the two-kernel stage of a single-GPU handle runs slab by slab -- face kernel
// over the faces mastered by n consecutive elements, then the element kernel over those elements --
// so that a slab's flux blocks (and the state the face kernel has just read) are still in L2 when
// the element kernel asks for them.  0 (default): one face launch and one element launch per stage.
// a state of 4 GB and more on one GPU (config 4): no x-face trace array and face slots in blocks of
// 4096 master elements unless FLOU_B200_XTRACE / FLOU_B200_FACE_CHUNK say otherwise
bool large_single_gpu_state(const flou_b200_desc *d)
{
    if (d->nranks > 1) return false;
    double npts = 1.0;
    for (int i = 0; i < d->nd; i++) npts *= d->np;
    return (double)d->ne * npts * d->nv * sizeof(double) >= 4.0 * (1u << 30);
}

long slab_elements()
{
    static const long v = [] { const char *e = std::getenv("FLOU_B200_SLAB"); return e ? std::atol(e) : 0L; }();
    return v;
}

thread_local std::string g_last_error;

int32_t fail(int32_t code, const std::string &msg)
{
    g_last_error = msg;
    return code;
}

#define CUDA_TRY(expr)                                                                       \
    do {                                                                                     \
        cudaError_t err__ = (expr);                                                          \
        if (err__ != cudaSuccess)                                                            \
            return fail(FLOU_B200_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(err__)); \
    } while (0)

// ---------------------------------------------------------------- NCCL through dlopen
// (keeps single-GPU use free of any NCCL dependency; with torch in the process this
// resolves to the libnccl.so.2 torch already loaded)
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat64 = 8 };
struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool load()
    {
        if (lib) return true;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) return false;
#define SYM(field, name) field = (decltype(field))dlsym(lib, name); if (!field) return false;
        SYM(GetUniqueId, "ncclGetUniqueId")
        SYM(CommInitRank, "ncclCommInitRank")
        SYM(CommDestroy, "ncclCommDestroy")
        SYM(GroupStart, "ncclGroupStart")
        SYM(GroupEnd, "ncclGroupEnd")
        SYM(Send, "ncclSend")
        SYM(Recv, "ncclRecv")
        SYM(AllReduce, "ncclAllReduce")
        SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
        return true;
    }
} g_nccl;

#define NCCL_TRY(expr)                                                                        \
    do {                                                                                      \
        ncclResult_t r__ = (expr);                                                            \
        if (r__ != 0)                                                                         \
            return fail(FLOU_B200_ENCCL, std::string(#expr) + ": " + g_nccl.GetErrorString(r__)); \
    } while (0)

template <class T>
cudaError_t upload(T **dst, const std::vector<T> &src)
{
    *dst = nullptr;
    const size_t bytes = std::max<size_t>(src.size(), 1) * sizeof(T);
    cudaError_t e = cudaMalloc((void **)dst, bytes);
    if (e != cudaSuccess) return e;
    if (!src.empty()) e = cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice);
    return e;
}

struct Peer {
    int rank;
    int64_t offset;   // first slot in the ghost / send buffers
    int64_t nslots;
};

// Host-only partition / halo plan: element-face records, ghost slots ordered by
// (peer rank, global face id) -- both sides of an interface sort the same way, so slot s of
// my send buffer is slot s of the peer's ghost buffer.
struct Ghost { int peer; int64_t gf; int le, lf; };
struct Plan {
    int64_t ne_local = 0;
    std::vector<Conn> conn;
    std::vector<Ghost> ghosts;
    std::vector<Peer> peers;
    std::vector<int> send_list, interior, boundary, faceid;
    std::vector<int64_t> slot_face;       // local face slot -> global face (0-based)
    // two-kernel path: one record per local face slot (faces with a ghost side last)
    std::vector<FaceRec> faces;
    std::vector<int2> econn;              // [le*2nd + lf] = {slot, master | orient << 1}
    int n_faces_local_only = 0;           // slots [0, n) touch no ghost
};

int32_t build_plan(const flou_b200_desc *d, bool cart, Plan &pl)
{
    const int nd = d->nd, NF = 2 * nd;
    const int nranks = d->nranks <= 0 ? 1 : d->nranks;
    pl.ne_local = d->elem_end - d->elem_begin;
    if (pl.ne_local * (int64_t)NF > INT32_MAX) return fail(FLOU_B200_EINVAL, "too many elements for one handle");
    // boundary-face lookup: global face -> (boundary index, ordinal in bc_faces)
    std::map<int64_t, std::pair<int, int64_t>> bdface;
    for (int ib = 0; ib < d->nbound; ib++)
        for (int64_t m = d->bc_offsets[ib]; m < d->bc_offsets[ib + 1]; m++)
            bdface[d->bc_faces[m] - 1] = {ib, m};
    pl.conn.assign((size_t)pl.ne_local * NF, Conn{0, 0});
    pl.faceid.assign((size_t)pl.ne_local * NF, 0);
    std::map<int64_t, int> face_slot;            // global face -> local slot (general geometry)
    auto owner_of = [&](int64_t ge) {
        return (int)(std::upper_bound(d->part_offsets, d->part_offsets + nranks + 1, ge) -
                     d->part_offsets) - 1;
    };
    for (int64_t le = 0; le < pl.ne_local; le++) {
        const int64_t ge = d->elem_begin + le;
        for (int lf = 0; lf < NF; lf++) {
            const int64_t gf = d->faceinds[ge * NF + lf] - 1;
            const int64_t pos = d->facepos[ge * NF + lf];
            if (gf < 0 || gf >= d->nf || (pos != 1 && pos != 2))
                return fail(FLOU_B200_EINVAL, "faceinds/facepos out of range");
            const int master = pos == 1;
            if (d->eleminds[gf * 2 + (master ? 0 : 1)] - 1 != ge ||
                d->elempos[gf * 2 + (master ? 0 : 1)] - 1 != lf)
                return fail(FLOU_B200_EINVAL, "element and face connectivity disagree");
            const int64_t gn = d->eleminds[gf * 2 + (master ? 1 : 0)] - 1;
            const int nlf = (int)d->elempos[gf * 2 + (master ? 1 : 0)] - 1;
            const int orient = d->orientation[gf];
            Conn c;
            if (gn < 0) {
                auto it = bdface.find(gf);
                if (it == bdface.end())
                    return fail(FLOU_B200_EINVAL, "boundary face without a boundary condition");
                c.nbr = (int)it->second.second;
                c.info = conn_pack(0, 0, 1, FK_BOUNDARY, it->second.first);
            } else if (gn >= d->elem_begin && gn < d->elem_end) {
                if (nlf < 0 || nlf >= NF) return fail(FLOU_B200_EINVAL, "elempos out of range");
                c.nbr = (int)(gn - d->elem_begin);
                c.info = conn_pack(nlf, orient, master, FK_INTERIOR, 0);
            } else {
                if (gn >= d->ne || nranks == 1) return fail(FLOU_B200_EINVAL, "eleminds out of range");
                c.nbr = -1;   // slot assigned after sorting
                c.info = conn_pack(nlf, orient, master, FK_GHOST, 0);
                pl.ghosts.push_back({owner_of(gn), gf, (int)le, lf});
            }
            pl.conn[(size_t)le * NF + lf] = c;
            {
                auto it = face_slot.find(gf);
                if (it == face_slot.end()) {
                    it = face_slot.emplace(gf, (int)pl.slot_face.size()).first;
                    pl.slot_face.push_back(gf);
                }
                pl.faceid[(size_t)le * NF + lf] = it->second;
            }
        }
    }
    std::sort(pl.ghosts.begin(), pl.ghosts.end(), [](const Ghost &a, const Ghost &b) {
        return a.peer != b.peer ? a.peer < b.peer : a.gf < b.gf;
    });
    pl.send_list.resize(pl.ghosts.size());
    std::vector<char> is_boundary_elem((size_t)pl.ne_local, 0);
    for (size_t s = 0; s < pl.ghosts.size(); s++) {
        const Ghost &g = pl.ghosts[s];
        pl.conn[(size_t)g.le * NF + g.lf].nbr = (int)s;
        pl.send_list[s] = g.le * NF + g.lf;
        is_boundary_elem[g.le] = 1;
        if (pl.peers.empty() || pl.peers.back().rank != g.peer)
            pl.peers.push_back({g.peer, (int64_t)s, 0});
        pl.peers.back().nslots++;
    }
    for (int64_t le = 0; le < pl.ne_local; le++)
        (is_boundary_elem[le] ? pl.boundary : pl.interior).push_back((int)le);

    // ---- face table of the two-kernel path: slots without a ghost side first
    (void)cart;
    const int nslots = (int)pl.slot_face.size();
    std::vector<char> has_ghost((size_t)nslots, 0);
    for (const Ghost &g : pl.ghosts) has_ghost[pl.faceid[(size_t)g.le * NF + g.lf]] = 1;
    // ... and, inside each class, grouped by the master's local face: face_flux_kernel is
    // specialised on it and wants warps that do not mix directions (Flou's own face order
    // interleaves them)
    std::vector<int> newslot((size_t)nslots);
    {
        std::vector<int> order((size_t)nslots);
        for (int sidx = 0; sidx < nslots; sidx++) order[sidx] = sidx;
        // FLOU_B200_FACE_CHUNK = c > 0: inside each class, blocks of c consecutive master elements,
        // and the master's local face inside a block (runs of up to c faces per direction keep the
        // switch warp-uniform): the traces of a block's x-, y- and z-faces are then read close in
        // time and the element kernel finds the fluxes of an element's own faces side by side.
        // 0 (default): one run per direction over the whole mesh.  Measured (profiles/r2j): blocks of
        // 256 cost the face kernel 3-4 % at config 4 and gain nothing elsewhere.
        // Together with NO x-face trace array (profiles/r2u, config 4): blocks of 4096 let the y / z node
        // layers hit in L2 behind the x-face reads of the same elements -- face kernel 20.8 -> 14.1 GB
        // read, stage 18.5 -> 18.1 ms: the default of a single-GPU state of 4 GB and more.
        static const long chunk_env = [] { const char *e = std::getenv("FLOU_B200_FACE_CHUNK"); return e ? std::atol(e) : -1L; }();
        // slab-wise stage (FLOU_B200_SLAB, single GPU): the face slots are blocked by slab
        const long slab = slab_elements();
        const long chunk = (slab > 0 && nranks == 1) ? slab
                           : chunk_env >= 0 ? chunk_env
                           : large_single_gpu_state(d) ? 4096L : 0L;
        auto key = [&](int sidx) -> int64_t {
            const int64_t gf = pl.slot_face[sidx];
            const int lfm_ = (int)d->elempos[gf * 2 + 0] - 1;
            const int64_t blk = chunk > 0 ? (d->eleminds[gf * 2 + 0] - 1) / chunk : 0;
            return ((has_ghost[sidx] ? (int64_t)1 << 40 : 0) + blk) * 8 + (lfm_ & 7);
        };
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return key(a) < key(b); });
        int n0 = 0;
        for (int q = 0; q < nslots; q++) {
            newslot[order[q]] = q;
            if (!has_ghost[order[q]]) n0 = q + 1;
        }
        pl.n_faces_local_only = n0;
    }
    {
        std::vector<int64_t> sf((size_t)nslots);
        for (int sidx = 0; sidx < nslots; sidx++) sf[newslot[sidx]] = pl.slot_face[sidx];
        pl.slot_face.swap(sf);
        for (int &f : pl.faceid) f = newslot[f];
    }
    pl.faces.assign((size_t)nslots, FaceRec{-1, -1, 0, 0});
    pl.econn.assign((size_t)pl.ne_local * NF, int2{0, 0});
    std::vector<int> lfm((size_t)nslots, 0), lfs((size_t)nslots, 0), mk((size_t)nslots, 0),
        sk((size_t)nslots, 0), bcidx((size_t)nslots, 0), ori((size_t)nslots, 0);
    for (int64_t le = 0; le < pl.ne_local; le++)
        for (int lf = 0; lf < NF; lf++) {
            const Conn &c = pl.conn[(size_t)le * NF + lf];
            const int slot = pl.faceid[(size_t)le * NF + lf];
            const int kind = (c.info >> 7) & 3, nlf = c.info & 7, orient = (c.info >> 3) & 7;
            const int master = (c.info >> 6) & 1;
            FaceRec &r = pl.faces[slot];
            pl.econn[(size_t)le * NF + lf] = int2{slot, master | (orient << 1)};
            ori[slot] = orient;
            if (master) { r.em = (int)le; lfm[slot] = lf; mk[slot] = 0; }
            else { r.es = (int)le; lfs[slot] = lf; sk[slot] = 0; }
            if (kind == FK_BOUNDARY) { r.es = c.nbr; sk[slot] = 2; bcidx[slot] = c.info >> 9; }
            else if (kind == FK_GHOST) {
                if (master) { r.es = c.nbr; lfs[slot] = nlf; sk[slot] = 1; }
                else { r.em = c.nbr; lfm[slot] = nlf; mk[slot] = 1; }
            } else {
                if (master) lfs[slot] = nlf; else lfm[slot] = nlf;
            }
        }
    for (int sidx = 0; sidx < nslots; sidx++)
        pl.faces[sidx].info = facerec_pack(lfm[sidx], lfs[sidx], ori[sidx], mk[sidx], sk[sidx], bcidx[sidx]);
    return FLOU_B200_OK;
}

int32_t validate_desc(const flou_b200_desc *d)
{
    if (!d) return fail(FLOU_B200_EINVAL, "null argument");
    if (d->struct_size != (int32_t)sizeof(flou_b200_desc))
        return fail(FLOU_B200_EINVAL, "flou_b200_desc size mismatch (ABI)");
    if (d->nd < 1 || d->nd > 3) return fail(FLOU_B200_EINVAL, "nd must be 1, 2 or 3");
    if (d->np < 2 || d->np > 8) return fail(FLOU_B200_EINVAL, "np must be between 2 and 8");
    const int nv_expected = d->equation == FLOU_B200_EQ_EULER ? d->nd + 2 : 1;
    if (d->equation != FLOU_B200_EQ_EULER && d->equation != FLOU_B200_EQ_LINEAR_ADVECTION)
        return fail(FLOU_B200_EINVAL, "unknown equation");
    if (d->nv != nv_expected) return fail(FLOU_B200_EINVAL, "nv does not match the equation");
    if (d->divop != FLOU_B200_OP_STRONG && d->divop != FLOU_B200_OP_SPLIT && d->divop != FLOU_B200_OP_HYBRID)
        return fail(FLOU_B200_EINVAL, "unknown divergence operator");
    if (d->divop == FLOU_B200_OP_HYBRID) {
        // OpDivergence.jl:452-612; vars_cons2entropy exists for the Euler equations only
        if (d->equation != FLOU_B200_EQ_EULER)
            return fail(FLOU_B200_EINVAL, "HybridDivOperator needs the Euler equations");
        if (d->flags & (FLOU_B200_FLAG_FUSED | FLOU_B200_FLAG_NODE_KERNEL))
            return fail(FLOU_B200_EINVAL, "HybridDivOperator exists in the line-per-thread kernel only");
        if (!(d->blend >= 0.0)) return fail(FLOU_B200_EINVAL, "HybridDivOperator blend must be >= 0");
    }
    if (d->numflux < 0 || d->numflux > FLOU_B200_FLUX_MATRIXDISSIPATION)
        return fail(FLOU_B200_EINVAL, "unknown numerical flux");
    if (d->equation == FLOU_B200_EQ_LINEAR_ADVECTION &&
        (d->numflux != FLOU_B200_FLUX_STDAVERAGE && d->numflux != FLOU_B200_FLUX_LXF))
        return fail(FLOU_B200_EINVAL, "linear advection supports StdAverage and LxF fluxes only");
    if (d->divop != FLOU_B200_OP_STRONG && d->tpflux != FLOU_B200_FLUX_STDAVERAGE &&
        d->tpflux != FLOU_B200_FLUX_CHANDRASEKHAR)
        return fail(FLOU_B200_EINVAL, "the two-point flux must be StdAverage or ChandrasekharAverage");
    if (d->equation == FLOU_B200_EQ_LINEAR_ADVECTION && d->divop == FLOU_B200_OP_SPLIT &&
        d->tpflux != FLOU_B200_FLUX_STDAVERAGE)
        return fail(FLOU_B200_EINVAL, "linear advection split form needs the StdAverage two-point flux");
    const bool has_avg = d->numflux == FLOU_B200_FLUX_LXF ||
                         d->numflux == FLOU_B200_FLUX_SCALARDISSIPATION ||
                         d->numflux == FLOU_B200_FLUX_MATRIXDISSIPATION;
    if (has_avg && d->numflux_avg != FLOU_B200_FLUX_STDAVERAGE &&
        d->numflux_avg != FLOU_B200_FLUX_CHANDRASEKHAR)
        return fail(FLOU_B200_EINVAL, "numflux.avg must be StdAverage or ChandrasekharAverage");
    if (d->geometry != FLOU_B200_GEOM_CARTESIAN && d->geometry != FLOU_B200_GEOM_GENERAL)
        return fail(FLOU_B200_EINVAL, "unknown geometry kind");
    if (d->ne <= 0 || d->nf <= 0 || !d->faceinds || !d->facepos || !d->eleminds || !d->elempos ||
        !d->orientation)
        return fail(FLOU_B200_EINVAL, "connectivity tables missing");
    if (d->nbound < 0 || (d->nbound > 0 && (!d->bc_kind || !d->bc_offsets || !d->bc_faces)))
        return fail(FLOU_B200_EINVAL, "boundary tables missing");
    const int nranks = d->nranks <= 0 ? 1 : d->nranks;
    if (d->elem_begin < 0 || d->elem_end > d->ne || d->elem_begin >= d->elem_end)
        return fail(FLOU_B200_EINVAL, "bad owned element range");
    if (nranks > 1 && !d->part_offsets) return fail(FLOU_B200_EINVAL, "part_offsets missing");
    if (nranks > 1 && (d->rank < 0 || d->rank >= nranks ||
                       d->part_offsets[d->rank] != d->elem_begin ||
                       d->part_offsets[d->rank + 1] != d->elem_end))
        return fail(FLOU_B200_EINVAL, "owned range does not match part_offsets[rank]");
    if (nranks == 1 && (d->elem_begin != 0 || d->elem_end != d->ne))
        return fail(FLOU_B200_EINVAL, "single-rank handle must own every element");
    // value ranges of the byte / enum tables: an unknown bc_kind would fall into the TABLE branch of
    // the face kernel and read past bc_table, a bad orientation into master2slave's default case
    const int max_orient = d->nd == 1 ? 0 : (d->nd == 2 ? 1 : 7);
    for (int64_t f = 0; f < d->nf; f++)
        if (d->orientation[f] > max_orient)
            return fail(FLOU_B200_EINVAL, "face orientation out of range (0 in 1-D, 0..1 in 2-D, 0..7 in 3-D)");
    if (d->nbound > 0 && d->bc_offsets[0] != 0)
        return fail(FLOU_B200_EINVAL, "bc_offsets must start at 0");
    for (int ib = 0; ib < d->nbound; ib++) {
        if (d->bc_kind[ib] < FLOU_B200_BC_INFLOW || d->bc_kind[ib] > FLOU_B200_BC_TABLE)
            return fail(FLOU_B200_EINVAL, "unknown boundary-condition kind");
        if (d->bc_offsets[ib + 1] < d->bc_offsets[ib])
            return fail(FLOU_B200_EINVAL, "bc_offsets must not decrease");
        for (int64_t m = d->bc_offsets[ib]; m < d->bc_offsets[ib + 1]; m++)
            if (d->bc_faces[m] < 1 || d->bc_faces[m] > d->nf)
                return fail(FLOU_B200_EINVAL, "bc_faces entry outside [1, nf]");
    }
    return FLOU_B200_OK;
}

}  // namespace

struct flou_b200_handle {
    int nd = 0, nv = 0, np = 0, npts = 0, nfp = 0, nfaces = 0;
    int device = 0, flags = 0;
    int rank = 0, nranks = 1;
    int64_t ne_local = 0, ndof = 0;
    const StageLauncher *stage = nullptr;
    const EmitLauncher *emit = nullptr;
    KParams base;
    // device memory
    double *u[2] = {nullptr, nullptr}, *tmp = nullptr, *k = nullptr;
    double *tr[2] = {nullptr, nullptr};   // x-face traces of u[0], u[1]
    double *tr_all = nullptr;             // Gauss nodes: interpolated traces of every face
    bool traces_valid = false;            // tr[cur] matches u[cur]
    bool colloc = false;
    int cur = 0;
    Conn *conn = nullptr;
    int *faceid = nullptr;
    double *jac = nullptr, *metric = nullptr, *fjac = nullptr, *frames = nullptr;
    double *sub_frames = nullptr, *sub_jac = nullptr;     // sub-grid tables of the owned elements
    int *bc_kind = nullptr;
    double *bc_state = nullptr, *bc_table = nullptr;
    int *status = nullptr;
    double *d_lm = nullptr, *d_lp = nullptr;
    // halo
    std::vector<Peer> peers;
    int64_t nghost = 0;
    double *ghost = nullptr, *sendbuf = nullptr;
    int *send_list = nullptr;          // [slot] = local element*2nd + local face
    int *interior_list = nullptr, *boundary_list = nullptr;
    int n_interior = 0, n_boundary = 0;
    int interior_first = -1;          // >= 0: the interior elements are the contiguous range starting here
    ncclComm_t comm = nullptr;
    // streams / events
    cudaStream_t stream = nullptr, comm_stream = nullptr;
    cudaEvent_t ev_emit = nullptr, ev_recv = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
    // CUDA graph of two consecutive RK steps (returns to the same ping-pong buffer)
    cudaGraphExec_t graph[2] = {nullptr, nullptr};      // one per starting ping-pong buffer
    std::vector<double> graph_key[2];
    int64_t launches = 0;
    int64_t graph_launches_per_replay = 0;
    // per-kernel CUDA-event timing (flou_b200_profile): {before faces, after faces, after elements}
    bool profile = false;
    std::vector<cudaEvent_t> prof_events;
    // get_max_dt
    double *elem_dx = nullptr;           // general geometry: (volume_e/npts)^(1/nd) per element
    double cart_dx = 0.0;
    unsigned long long *dt_bits = nullptr;
    double gamma = 0.0, anorm = 0.0;
    int equation = 0;
    // monitors / Zhang-Shu limiter (row f3)
    double *w_nodes = nullptr;           // tensor-product quadrature weights of one element [npts]
    double *mon_partial = nullptr;       // per-block partial sums + the result in the last slot
    int mon_blocks = 0;
    bool stage_limiter = false;          // advance(): limiter after every RK stage (stage_limiter!)
    double limiter_minval = 0.0;
    // two-kernel path
    bool split_faces = true;
    bool line_kernel = true;             // element kernel of the two-kernel stage: line per thread
    bool keep_xtraces = true;            // FLOU_B200_XTRACE=0 / 1 overrides the size rule (see uses_xtraces)
    FaceRec *faces = nullptr;
    int2 *econn = nullptr;
    double *Fn = nullptr;
    int n_faces = 0, n_faces_local_only = 0;
    // slab-wise stage: launch list {kind (0 face / 1 element), first, count}
    struct SlabOp { int kind, first, count; };
    std::vector<SlabOp> slab_ops;
    // source term (row a13) and boundary data that change between stages
    double *source = nullptr;            // [dof + ndof*v], allocated by the first flou_b200_set_source
    int64_t bc_rows = 0;                 // rows (boundary-face nodes) of bc_table
    int *bd_list = nullptr;              // owned boundary faces: local element*2nd + local face
    std::vector<int64_t> bd_ordinal;     // ... and their ordinal m in the concatenated bc_faces
    double *bd_traces = nullptr;         // [owned boundary face][v][k]
};

namespace {

// The x-face trace array (written by the element kernel with the new state, read by the face
// kernel / the fused kernel) is needed by the fused kernel and with Gauss nodes.  In the two-kernel
// stage with collocated nodes the face kernel can read the x-face node layers of u like those of the
// other directions (FLOU_B200_XTRACE=0).  Measured (profiles/r2m): at config 4 the element kernel
// writes 4.2 GB less and loses its trace pass (ncu 12.70 -> 11.96 ms at the burst clock, no change
// under the board's power cap), the face kernel reads u once more (20.8 -> 27.2 GB, 4.49 -> 5.05 ms):
// stage +2.7 %; config 2 (L2-resident) -4.6 %.  With the face slots in blocks of 4096 master elements
// the face kernel reads 14.1 GB instead and the stage is 2.5 % FASTER (profiles/r2u).  Default: no
// array for states up to 48 MB and for single-GPU states of 4 GB and more (then with the blocks).
bool uses_xtraces(const flou_b200_handle *h)
{
    return !(h->colloc && h->split_faces) || h->keep_xtraces;
}

// one RHS / stage pass over the owned elements (halo exchange included when partitioned)
int32_t run_pass(flou_b200_handle *h, int mode, double A, double B, double dt,
                 const double *u_in, double *u_out)
{
    KParams P = h->base;
    const int iin = (u_in == h->u[0]) ? 0 : 1;
    const bool xtr = uses_xtraces(h);
    if (xtr && !h->traces_valid) {
        // traces of u_in (first pass after an upload, and every pass with Gauss nodes)
        CUDA_TRY(h->emit->launch(u_in, h->ndof, nullptr, (int)(h->ne_local * 2), 2,
                                 h->base.colloc, h->d_lm, h->d_lp, h->tr[iin], h->stream));
        h->launches += 1;
        if (!h->colloc) {
            CUDA_TRY(h->emit->launch(u_in, h->ndof, nullptr, (int)(h->ne_local * h->nfaces), h->nfaces,
                                     0, h->d_lm, h->d_lp, h->tr_all, h->stream));
            h->launches += 1;
        }
        h->traces_valid = true;
    }
    P.tr_hi = h->tr_all;
    P.tr_in = xtr ? h->tr[iin] : nullptr;
    P.tr_out = xtr ? h->tr[iin ^ 1] : nullptr;
    P.u_in = u_in;
    P.u_out = u_out;
    P.tmp = h->tmp;
    P.k_out = h->k;
    P.mode = mode;
    P.rkA = A;
    P.rkB = B;
    P.dt = dt;
    // after a stage pass the kernel has written the traces of u_out (collocated nodes only)
    const bool out_traces = (mode != MODE_RHS) && h->colloc;
    if (h->nranks == 1 || h->nghost == 0) {
        P.elem_first = 0;
        P.elem_count = (int)h->ne_local;
        P.elem_list = nullptr;
        if (h->split_faces) {
            // two-kernel stage: every face flux once, then volume + lift + RK update
            P.face_first = 0;
            P.face_count = h->n_faces;
            cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
            if (h->profile) {
                for (auto &e : ev) { CUDA_TRY(cudaEventCreate(&e)); h->prof_events.push_back(e); }
                CUDA_TRY(cudaEventRecord(ev[0], h->stream));
            }
            if (!h->slab_ops.empty() && !h->profile) {
                for (const auto &op : h->slab_ops) {
                    if (op.kind == 0) {
                        P.face_first = op.first; P.face_count = op.count;
                        CUDA_TRY(h->stage->launch_faces(P, h->stream));
                    } else {
                        P.elem_first = op.first; P.elem_count = op.count;
                        CUDA_TRY(h->stage->launch_lines(P, h->stream));
                    }
                }
                h->launches += (int64_t)h->slab_ops.size();
                if (mode != MODE_RHS) h->traces_valid = out_traces;
                return FLOU_B200_OK;
            }
            CUDA_TRY(h->stage->launch_faces(P, h->stream));
            if (h->profile) CUDA_TRY(cudaEventRecord(ev[1], h->stream));
            CUDA_TRY((h->line_kernel ? h->stage->launch_lines : h->stage->launch_elements)(P, h->stream));
            if (h->profile) CUDA_TRY(cudaEventRecord(ev[2], h->stream));
            h->launches += 2;
        } else {
            CUDA_TRY(h->stage->launch(P, h->stream));
            h->launches += 1;
        }
        if (mode != MODE_RHS) h->traces_valid = out_traces;
        return FLOU_B200_OK;
    }
    if (!h->comm) return fail(FLOU_B200_EINVAL, "partitioned handle used before flou_b200_comm_init");
    // 1. pack the traces the neighbours need, 2. exchange on the comm stream,
    // 3. interior elements meanwhile, 4. partition-boundary elements after the receive
    CUDA_TRY(h->emit->launch(u_in, h->ndof, h->send_list, (int)h->nghost, h->nfaces, h->base.colloc,
                             h->d_lm, h->d_lp, h->sendbuf, h->stream));
    h->launches += 1;
    CUDA_TRY(cudaEventRecord(h->ev_emit, h->stream));
    CUDA_TRY(cudaStreamWaitEvent(h->comm_stream, h->ev_emit, 0));
    const size_t per_slot = (size_t)h->nv * h->nfp;
    NCCL_TRY(g_nccl.GroupStart());
    for (const Peer &p : h->peers) {
        NCCL_TRY(g_nccl.Send(h->sendbuf + p.offset * per_slot, p.nslots * per_slot, ncclFloat64,
                             p.rank, h->comm, h->comm_stream));
        NCCL_TRY(g_nccl.Recv(h->ghost + p.offset * per_slot, p.nslots * per_slot, ncclFloat64,
                             p.rank, h->comm, h->comm_stream));
    }
    NCCL_TRY(g_nccl.GroupEnd());
    CUDA_TRY(cudaEventRecord(h->ev_recv, h->comm_stream));
    if (h->split_faces) {
        // faces without a ghost side while the halo is in flight, the rest after the receive
        P.face_first = 0;
        P.face_count = h->n_faces_local_only;
        CUDA_TRY(h->stage->launch_faces(P, h->stream));
        CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_recv, 0));
        P.face_first = h->n_faces_local_only;
        P.face_count = h->n_faces - h->n_faces_local_only;
        CUDA_TRY(h->stage->launch_faces(P, h->stream));
        P.elem_first = 0;
        P.elem_count = (int)h->ne_local;
        P.elem_list = nullptr;
        CUDA_TRY((h->line_kernel ? h->stage->launch_lines : h->stage->launch_elements)(P, h->stream));
        h->launches += 3;
        if (mode != MODE_RHS) h->traces_valid = out_traces;
        return FLOU_B200_OK;
    }
    // interior elements: a contiguous range on slab partitions (no indirection), else a list
    P.elem_first = h->interior_first >= 0 ? h->interior_first : 0;
    P.elem_list = h->interior_first >= 0 ? nullptr : h->interior_list;
    P.elem_count = h->n_interior;
    CUDA_TRY(h->stage->launch(P, h->stream));
    P.elem_first = 0;
    CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_recv, 0));
    P.elem_list = h->boundary_list;
    P.elem_count = h->n_boundary;
    CUDA_TRY(h->stage->launch(P, h->stream));
    h->launches += 2;
    if (mode != MODE_RHS) h->traces_valid = out_traces;
    return FLOU_B200_OK;
}

// zhang_shu_limiter (Equations/Euler.jl:616-660) on a device state, in place; the x-face traces
// of that state are stale afterwards
int32_t launch_zhang_shu(flou_b200_handle *h, double *u, double minval)
{
    const int threads = 256, warps_per_block = threads / 32;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
    const int64_t want = (h->ne_local + warps_per_block - 1) / warps_per_block;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)sms * 8));
    const double *jac = h->base.jac;     // nullptr on Cartesian meshes
    switch (h->nv) {
    case 3: zhang_shu_kernel<3><<<grid, threads, 0, h->stream>>>(u, h->ndof, h->ne_local, h->npts, h->gamma, minval, h->w_nodes, jac, h->base.cjac); break;
    case 4: zhang_shu_kernel<4><<<grid, threads, 0, h->stream>>>(u, h->ndof, h->ne_local, h->npts, h->gamma, minval, h->w_nodes, jac, h->base.cjac); break;
    case 5: zhang_shu_kernel<5><<<grid, threads, 0, h->stream>>>(u, h->ndof, h->ne_local, h->npts, h->gamma, minval, h->w_nodes, jac, h->base.cjac); break;
    default: return fail(FLOU_B200_EINVAL, "the Zhang-Shu limiter is defined for the Euler equations");
    }
    CUDA_TRY(cudaGetLastError());
    h->launches += 1;
    h->traces_valid = false;
    return FLOU_B200_OK;
}

int32_t run_steps_direct(flou_b200_handle *h, int nstages, const double *A, const double *B,
                         double dt, int64_t nsteps)
{
    for (int64_t it = 0; it < nsteps; it++)
        for (int s = 0; s < nstages; s++) {
            const int32_t rc = run_pass(h, s == 0 ? MODE_STAGE_FIRST : MODE_STAGE, A[s], B[s], dt,
                                        h->u[h->cur], h->u[h->cur ^ 1]);
            if (rc) return rc;
            h->cur ^= 1;
            if (h->stage_limiter) {
                // OrdinaryDiffEq's `stage_limiter!(u, integrator, p, t)` after every stage update
                // (examples/src/3D_Euler.jl:76-80)
                const int32_t rl = launch_zhang_shu(h, h->u[h->cur], h->limiter_minval);
                if (rl) return rl;
            }
        }
    return FLOU_B200_OK;
}

// State a query (get_max_dt, monitor, limiter call) works on: the device-resident state, or -- when
// the caller passes an explicit host state -- a copy of it in the scratch ping-pong buffer u[cur^1],
// which nothing uses between passes, so that the state of an ongoing advance() stays untouched.
int32_t query_state(flou_b200_handle *h, const double *Q, double **state)
{
    *state = h->u[h->cur];
    if (Q) {
        *state = h->u[h->cur ^ 1];
        CUDA_TRY(cudaMemcpyAsync(*state, Q, sizeof(double) * (size_t)h->ndof * h->nv,
                                 cudaMemcpyHostToDevice, h->stream));
    }
    return FLOU_B200_OK;
}

void destroy_graph(flou_b200_handle *h)
{
    for (int i = 0; i < 2; i++) {
        if (h->graph[i]) cudaGraphExecDestroy(h->graph[i]);
        h->graph[i] = nullptr;
        h->graph_key[i].clear();
    }
}

}  // namespace

extern "C" {

const char *flou_b200_last_error(void) { return g_last_error.c_str(); }

int32_t flou_b200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

static int vol_kind(int32_t divop, int32_t tpflux)
{
    if (divop == FLOU_B200_OP_STRONG) return VOL_STRONG;
    if (divop == FLOU_B200_OP_HYBRID) return VOL_HYBRID;
    return tpflux == FLOU_B200_FLUX_CHANDRASEKHAR ? VOL_SPLIT_CHA : VOL_SPLIT_STD;
}

int32_t flou_b200_supported(int32_t nd, int32_t np, int32_t equation, int32_t divop,
                            int32_t tpflux, int32_t geometry)
{
    if (equation != FLOU_B200_EQ_LINEAR_ADVECTION && equation != FLOU_B200_EQ_EULER) return 0;
    if (divop != FLOU_B200_OP_STRONG && divop != FLOU_B200_OP_SPLIT && divop != FLOU_B200_OP_HYBRID) return 0;
    if (divop != FLOU_B200_OP_STRONG && tpflux != FLOU_B200_FLUX_STDAVERAGE &&
        tpflux != FLOU_B200_FLUX_CHANDRASEKHAR) return 0;
    return get_stage_launcher(nd, np, equation, vol_kind(divop, tpflux),
                              geometry == FLOU_B200_GEOM_CARTESIAN) != nullptr;
}

int32_t flou_b200_partition_plan(const flou_b200_desc *d, int64_t *nghost, int32_t *npeers,
                                 int32_t *peer_ranks, int64_t *peer_nslots,
                                 int64_t *ghost_faces, int32_t *ghost_elemfaces,
                                 int64_t *n_interior, int64_t *n_boundary)
{
    if (int32_t rc = validate_desc(d)) return rc;
    Plan pl;
    if (int32_t rc = build_plan(d, d->geometry == FLOU_B200_GEOM_CARTESIAN, pl)) return rc;
    if (nghost) *nghost = (int64_t)pl.ghosts.size();
    if (npeers) *npeers = (int32_t)pl.peers.size();
    for (size_t i = 0; i < pl.peers.size(); i++) {
        if (peer_ranks) peer_ranks[i] = pl.peers[i].rank;
        if (peer_nslots) peer_nslots[i] = pl.peers[i].nslots;
    }
    for (size_t s = 0; s < pl.ghosts.size(); s++) {
        if (ghost_faces) ghost_faces[s] = pl.ghosts[s].gf + 1;
        if (ghost_elemfaces) ghost_elemfaces[s] = pl.send_list[s];
    }
    if (n_interior) *n_interior = (int64_t)pl.interior.size();
    if (n_boundary) *n_boundary = (int64_t)pl.boundary.size();
    return FLOU_B200_OK;
}

int32_t flou_b200_create(const flou_b200_desc *d, flou_b200_handle **out)
{
    if (!d || !out) return fail(FLOU_B200_EINVAL, "null argument");
    *out = nullptr;
    if (int32_t rc = validate_desc(d)) return rc;
    const int nd = d->nd, np = d->np;
    if (!d->Ds || !d->Dsharp || !d->lminus || !d->lplus || !d->dgminus || !d->dgplus)
        return fail(FLOU_B200_EINVAL, "operator tables missing");
    if (d->divop == FLOU_B200_OP_HYBRID && (!d->D || !d->weights))
        return fail(FLOU_B200_EINVAL, "HybridDivOperator needs std.D and the 1-D weights");
    if (d->geometry == FLOU_B200_GEOM_GENERAL && (!d->jac || !d->metric || !d->fjac || !d->frames))
        return fail(FLOU_B200_EINVAL, "general geometry tables missing");
    if (!flou_b200_supported(nd, np, d->equation, d->divop, d->tpflux, d->geometry))
        return fail(FLOU_B200_EUNSUPPORTED, "no kernel compiled for this (nd, np, equation, operator)");
    const int nranks = d->nranks <= 0 ? 1 : d->nranks;
    const bool cart = d->geometry == FLOU_B200_GEOM_CARTESIAN;

    if (flou_b200_device_count() <= d->device || d->device < 0)
        return fail(FLOU_B200_ECUDA, "no usable CUDA device (this library has no CPU path)");
    CUDA_TRY(cudaSetDevice(d->device));
    Plan pl;
    if (int32_t rc = build_plan(d, cart, pl)) return rc;

    flou_b200_handle *h = new flou_b200_handle();
    h->nd = nd; h->nv = d->nv; h->np = np;
    h->npts = 1; for (int i = 0; i < nd; i++) h->npts *= np;
    h->nfp = h->npts / np;
    h->nfaces = 2 * nd;
    h->device = d->device; h->flags = d->flags;
    h->rank = d->rank; h->nranks = nranks;
    h->ne_local = pl.ne_local;
    h->ndof = h->ne_local * h->npts;
    h->stage = get_stage_launcher(nd, np, d->equation, vol_kind(d->divop, d->tpflux), cart);
    h->emit = get_emit_launcher(nd, np, d->nv);
    h->peers = pl.peers;
    h->nghost = (int64_t)pl.ghosts.size();
    h->n_interior = (int)pl.interior.size();
    h->n_boundary = (int)pl.boundary.size();
    if (!pl.interior.empty() &&
        pl.interior.back() - pl.interior.front() + 1 == (int)pl.interior.size())
        h->interior_first = pl.interior.front();
    const std::vector<Conn> &conn = pl.conn;
    const std::vector<int> &send_list = pl.send_list, &interior = pl.interior,
                           &boundary = pl.boundary, &faceid = pl.faceid;
    const std::vector<int64_t> &slot_face = pl.slot_face;

    // ---- kernel parameters
    KParams &P = h->base;
    std::memset(&P, 0, sizeof(P));
    const bool hybrid = d->divop == FLOU_B200_OP_HYBRID;
    const double *Dvol = d->divop == FLOU_B200_OP_STRONG ? d->Ds : (hybrid ? d->D : d->Dsharp);
    for (int i = 0; i < np * np; i++) P.Dvol[i] = Dvol[i];
    P.tpflux = d->tpflux;
    P.blend = hybrid ? d->blend : 0.0;
    for (int i = 0; i < np; i++) P.w1d[i] = d->weights ? d->weights[i] : 0.0;
    bool colloc = true;
    for (int i = 0; i < np; i++) {
        P.lm[i] = d->lminus[i]; P.lp[i] = d->lplus[i];
        P.dgl[i] = d->dgminus[i]; P.dgr[i] = d->dgplus[i];
        const double em = (i == 0) ? 1.0 : 0.0, ep = (i == np - 1) ? 1.0 : 0.0;
        if (std::fabs(d->lminus[i] - em) > 1e-13 || std::fabs(d->lplus[i] - ep) > 1e-13) colloc = false;
    }
    P.colloc = colloc ? 1 : 0;
    {
        // split form: the diagonal of D# is analytically zero on GLL nodes; the reference carries
        // round-off there (<= 4e-13 at np = 8, 1e-14 relative to max|D#|).  Entries below
        // 1e-12 max|D| are skipped by the line kernel.
        double dmax = 0.0;
        for (int i = 0; i < np * np; i++) dmax = std::max(dmax, std::fabs(Dvol[i]));
        P.diag_mask = 0;
        for (int j = 0; j < np; j++)
            if (std::fabs(Dvol[j + np * j]) > 1e-12 * dmax) P.diag_mask |= 1 << j;
    }
    h->colloc = colloc;
    // split form on nodes without boundaries (Gauss): entropy-projected surface term
    // (_splitdiv_nb_surface_contribution!, OpDivergence.jl:300-437), line kernels only
    const bool split_nb = d->divop == FLOU_B200_OP_SPLIT && !colloc;
    if (split_nb) {
        const char *why = nullptr;
        if (d->equation != FLOU_B200_EQ_EULER) why = "SplitDivOperator on Gauss nodes needs entropy variables (Euler equations)";
        else if (!cart && (!d->sub_frames || !d->sub_jac)) why = "SplitDivOperator on Gauss nodes: sub-grid tables (sub_frames, sub_jac) missing";
        else if (d->flags & (FLOU_B200_FLAG_FUSED | FLOU_B200_FLAG_NODE_KERNEL))
            why = "SplitDivOperator on Gauss nodes exists in the line-per-thread kernel only";
        if (why) { flou_b200_destroy(h); return fail(FLOU_B200_EUNSUPPORTED, why); }
        // the instances built for such nodes (dispatch indices 4 / 5, inst.cu)
        h->stage = get_stage_launcher(nd, np, d->equation,
                                      d->tpflux == FLOU_B200_FLUX_CHANDRASEKHAR ? 5 : 4, cart);
        if (!h->stage) { flou_b200_destroy(h); return fail(FLOU_B200_EUNSUPPORTED, "no kernel compiled for this (nd, np)"); }
    }
    if (hybrid && !cart && (!d->sub_frames || !d->sub_jac)) {
        flou_b200_destroy(h);
        return fail(FLOU_B200_EINVAL, "HybridDivOperator on a general mesh: sub-grid tables (sub_frames, sub_jac) missing");
    }
    if (hybrid && !colloc) {
        // nodes without boundaries: everything is a surface contribution
        // (_hybrid_nb_surface_contribution!, OpDivergence.jl:629-779): instances of dispatch index 6,
        // which use D# (the GLL form uses D)
        h->stage = get_stage_launcher(nd, np, d->equation, 6, cart);
        if (!h->stage) { flou_b200_destroy(h); return fail(FLOU_B200_EUNSUPPORTED, "no kernel compiled for this (nd, np)"); }
        for (int i = 0; i < np * np; i++) P.Dvol[i] = d->Dsharp[i];
    }
    if (colloc) {
        // GLL: l(-1) = e_1 and l(+1) = e_np up to the reference's monomial round-off
        // (O(1e-16) off-entries); the lifting weights away from the end nodes are dropped
        for (int i = 0; i < np; i++) {
            if (i != 0) P.dgl[i] = 0.0;
            if (i != np - 1) P.dgr[i] = 0.0;
        }
    }
    P.fp.gamma = d->gamma; P.fp.intensity = d->intensity;
    P.fp.gm1 = d->gamma - 1.0; P.fp.inv_gm1 = 1.0 / (d->gamma - 1.0); P.fp.inv_gamma = 1.0 / d->gamma;
    for (int c = 0; c < 3; c++) P.fp.a[c] = d->a[c];
    P.fp.numflux = d->numflux; P.fp.numflux_avg = d->numflux_avg;
    if (cart) {
        // PhysicalRegions.jl:370-408 (element) and :541-696 (faces)
        double prod = 1.0;
        for (int c = 0; c < nd; c++) prod *= d->dx[c];
        P.cjac = prod / (double)(1 << nd);
        P.crjac = 1.0 / P.cjac;
        if (nd == 1) { P.cmet[0] = 1.0; P.cfjac[0] = 1.0; }
        else if (nd == 2) {
            P.cmet[0] = d->dx[1] / 2; P.cmet[1] = d->dx[0] / 2;
            P.cfjac[0] = d->dx[1] / 2; P.cfjac[1] = d->dx[0] / 2;
        } else {
            P.cmet[0] = d->dx[1] * d->dx[2] / 4; P.cmet[1] = d->dx[0] * d->dx[2] / 4;
            P.cmet[2] = d->dx[0] * d->dx[1] / 4;
            for (int c = 0; c < 3; c++) P.cfjac[c] = P.cmet[c];
        }
        for (int c = 0; c < nd; c++) P.rcmet[c] = 1.0 / P.cmet[c];
    }

#define H_TRY(expr)                                                                            \
    do {                                                                                       \
        cudaError_t err__ = (expr);                                                            \
        if (err__ != cudaSuccess) {                                                            \
            std::string m__ = std::string(#expr) + ": " + cudaGetErrorString(err__);           \
            flou_b200_destroy(h);                                                              \
            return fail(FLOU_B200_ECUDA, m__);                                                 \
        }                                                                                      \
    } while (0)

    H_TRY(upload(&h->conn, conn));
    h->split_faces = !(d->flags & FLOU_B200_FLAG_FUSED);
    {
        // meshes too small to fill the device with element groups are launch-latency-bound: one
        // fused launch per stage beats two (config 1: 1024 elements = 64 groups on 148 SMs)
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
        const int64_t groups = (h->ne_local + h->stage->line_e - 1) / std::max(1, h->stage->line_e);
        if (h->nranks == 1 && groups < sms && !hybrid && !split_nb &&
            !(d->flags & (FLOU_B200_FLAG_NODE_KERNEL | FLOU_B200_FLAG_LINE_KERNEL)))
            h->split_faces = false;
    }
    h->line_kernel = h->split_faces && !(d->flags & FLOU_B200_FLAG_NODE_KERNEL);
    {
        // the fused / node-per-thread kernels of this instance may not fit an SM's shared memory
        // either (3-D np = 8, general geometry, split form): use the line kernel, or refuse when the
        // caller insisted on one of them
        int optin = 0;
        cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device);
        if (optin > 0 && h->stage->launch && h->stage->smem > (size_t)optin && !h->line_kernel) {
            if (d->flags & (FLOU_B200_FLAG_FUSED | FLOU_B200_FLAG_NODE_KERNEL)) {
                flou_b200_destroy(h);
                return fail(FLOU_B200_EUNSUPPORTED, "the fused / node kernels of this (nd, np, operator) need more shared memory than the device offers");
            }
            h->split_faces = true;
            h->line_kernel = true;
        }
    }
    if (h->line_kernel) {
        // the line kernel of this instance may need more shared memory than an SM has (one element
        // per group at 3-D np = 8): fall back to the node-per-thread element kernel where one
        // exists, refuse otherwise (hybrid operator, Gauss-node split form: line kernel only)
        int optin = 0;
        cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device);
        if (optin > 0 && h->stage->line_smem > (size_t)optin) {
            if (!h->stage->launch_elements || h->stage->smem > (size_t)optin || (d->flags & FLOU_B200_FLAG_LINE_KERNEL)) {
                flou_b200_destroy(h);
                return fail(FLOU_B200_EUNSUPPORTED, "the kernels of this (nd, np, operator) need more shared memory than the device offers");
            }
            h->line_kernel = false;
        }
    }
    h->n_faces = (int)pl.faces.size();
    h->n_faces_local_only = pl.n_faces_local_only;
    H_TRY(upload(&h->faces, pl.faces));
    H_TRY(upload(&h->econn, pl.econn));
    if (slab_elements() > 0 && h->nranks == 1 && h->split_faces && h->line_kernel && h->ne_local > slab_elements()) {
        const int64_t slab = slab_elements(), ne = h->ne_local;
        const int nslab = (int)((ne + slab - 1) / slab), NFc = 2 * nd;
        // face slots are sorted by the slab of their master element: first slot of every slab
        std::vector<int> ffirst((size_t)nslab + 1, h->n_faces), blk_of((size_t)h->n_faces, 0);
        for (int sl = h->n_faces - 1; sl >= 0; sl--) {
            const int b = (int)(pl.faces[sl].em / slab);
            blk_of[sl] = b;
            ffirst[b] = sl;
        }
        for (int b = nslab - 1; b >= 0; b--) ffirst[b] = std::min(ffirst[b], ffirst[b + 1]);
        ffirst[0] = 0;
        bool sorted = true;
        for (int sl = 1; sl < h->n_faces; sl++) sorted = sorted && blk_of[sl - 1] <= blk_of[sl];
        // an element slab can run once the last face slab any of its faces lives in is done
        std::vector<int> need((size_t)nslab, 0);
        for (int64_t e = 0; e < ne; e++)
            for (int lf = 0; lf < NFc; lf++)
                need[e / slab] = std::max(need[e / slab], blk_of[pl.econn[(size_t)e * NFc + lf].x]);
        for (int c = 0; c < nslab && sorted; c++) {
            if (ffirst[c + 1] > ffirst[c]) h->slab_ops.push_back({0, ffirst[c], ffirst[c + 1] - ffirst[c]});
            for (int b = 0; b < nslab; b++)
                if (need[b] == c)
                    h->slab_ops.push_back({1, (int)(b * slab), (int)std::min<int64_t>(slab, ne - b * slab)});
        }
    }
    if (h->split_faces)
        H_TRY(cudaMalloc((void **)&h->Fn, sizeof(double) * std::max<size_t>((size_t)h->n_faces * fn_block(h->nv, h->nfp), 2)));
    if (!cart) {
        // re-lay geometry as plane-major SoA restricted to owned elements / touched faces
        const int64_t ndof = h->ndof, npts = h->npts, nfp = h->nfp;
        std::vector<double> jac((size_t)ndof), met((size_t)ndof * nd * nd);
        for (int64_t i = 0; i < ndof; i++) {
            const int64_t gi = d->elem_begin * npts + i;
            jac[i] = d->jac[gi];
            for (int m = 0; m < nd * nd; m++) met[(size_t)m * ndof + i] = d->metric[gi * nd * nd + m];
        }
        const int64_t nfd = (int64_t)slot_face.size() * nfp;
        std::vector<double> fj((size_t)nfd), fr((size_t)nfd * 3 * nd);
        for (size_t s = 0; s < slot_face.size(); s++)
            for (int64_t i = 0; i < nfp; i++) {
                const int64_t gi = slot_face[s] * nfp + i, li = (int64_t)s * nfp + i;
                fj[li] = d->fjac[gi];
                for (int c = 0; c < 3 * nd; c++) fr[(size_t)c * nfd + li] = d->frames[gi * 3 * nd + c];
            }
        H_TRY(upload(&h->jac, jac));
        H_TRY(upload(&h->metric, met));
        H_TRY(upload(&h->fjac, fj));
        H_TRY(upload(&h->frames, fr));
        if ((hybrid || split_nb) && d->sub_frames && d->sub_jac) {
            // geometry.subgrids of the owned elements, layout of the descriptor kept
            const size_t per_elem = (size_t)nd * h->nfp * (np + 1);
            std::vector<double> sf(d->sub_frames + (size_t)d->elem_begin * per_elem * 3 * nd,
                                   d->sub_frames + (size_t)d->elem_end * per_elem * 3 * nd);
            std::vector<double> sj(d->sub_jac + (size_t)d->elem_begin * per_elem,
                                   d->sub_jac + (size_t)d->elem_end * per_elem);
            H_TRY(upload(&h->sub_frames, sf));
            H_TRY(upload(&h->sub_jac, sj));
        }
        H_TRY(upload(&h->faceid, faceid));
        P.nfacedofs = nfd;
    }
    {
        std::vector<int> kinds(d->bc_kind, d->bc_kind + d->nbound);
        std::vector<double> state, table;
        if (d->nbound > 0 && d->bc_state) state.assign(d->bc_state, d->bc_state + (size_t)d->nbound * d->nv);
        if (d->nbound > 0 && d->bc_table)
            table.assign(d->bc_table, d->bc_table + (size_t)d->bc_offsets[d->nbound] * h->nfp * d->nv);
        bool need_table = false, need_state = false;
        for (int k : kinds) { need_table |= k == FLOU_B200_BC_TABLE; need_state |= k == FLOU_B200_BC_INFLOW; }
        if ((need_table && table.empty()) || (need_state && state.empty()))
            { flou_b200_destroy(h); return fail(FLOU_B200_EINVAL, "boundary-condition data missing"); }
        // full-size table even when no boundary is tabulated yet (flou_b200_set_bc_table)
        if (d->nbound > 0 && table.empty())
            table.assign((size_t)d->bc_offsets[d->nbound] * h->nfp * d->nv, 0.0);
        H_TRY(upload(&h->bc_kind, kinds));
        H_TRY(upload(&h->bc_state, state));
        H_TRY(upload(&h->bc_table, table));
    }
    {
        std::vector<double> lm(P.lm, P.lm + 8), lp(P.lp, P.lp + 8);
        H_TRY(upload(&h->d_lm, lm));
        H_TRY(upload(&h->d_lp, lp));
    }
    {
        // dx_e = (volume_e / npts)^(1/nd), volume_e = sum_i jac_i * w_i with the tensor-product
        // weights w = wx*wy*wz (StdQuad.jl:49, StdHex.jl:54; PhysicalRegions.jl:403-405)
        if (!d->weights) { flou_b200_destroy(h); return fail(FLOU_B200_EINVAL, "weights (std.w) missing"); }
        const int npts = h->npts;
        std::vector<double> w((size_t)npts);
        for (int i = 0; i < npts; i++) {
            double x = d->weights[i % np];
            if (nd > 1) x *= d->weights[(i / np) % np];
            if (nd > 2) x *= d->weights[i / (np * np)];
            w[i] = x;
        }
        auto to_dx = [&](double vol) {
            const double r = vol / npts;
            return nd == 1 ? r : (nd == 2 ? std::sqrt(r) : std::cbrt(r));
        };
        if (cart) {
            double vol = 0.0;
            for (int i = 0; i < npts; i++) vol += P.cjac * w[i];
            h->cart_dx = to_dx(vol);
        } else {
            std::vector<double> edx((size_t)h->ne_local);
            for (int64_t le = 0; le < h->ne_local; le++) {
                double vol = 0.0;
                const double *j = d->jac + (d->elem_begin + le) * npts;
                for (int i = 0; i < npts; i++) vol += j[i] * w[i];
                edx[le] = to_dx(vol);
            }
            H_TRY(upload(&h->elem_dx, edx));
        }
        h->gamma = d->gamma;
        h->equation = d->equation;
        double a2 = 0.0;
        for (int c = 0; c < nd; c++) a2 += d->a[c] * d->a[c];
        h->anorm = std::sqrt(a2);
        H_TRY(cudaMalloc((void **)&h->dt_bits, sizeof(unsigned long long)));
        H_TRY(upload(&h->w_nodes, w));
        int sms_ = 148;
        cudaDeviceGetAttribute(&sms_, cudaDevAttrMultiProcessorCount, h->device);
        h->mon_blocks = sms_ * 4;
        H_TRY(cudaMalloc((void **)&h->mon_partial, sizeof(double) * (size_t)(h->mon_blocks + 1)));
    }
    {
        // owned boundary faces in bc_faces order (flou_b200_boundary_traces)
        std::vector<int> bl;
        for (int64_t le = 0; le < pl.ne_local; le++)
            for (int lf = 0; lf < 2 * nd; lf++) {
                const Conn &c = conn[(size_t)le * 2 * nd + lf];
                if (((c.info >> 7) & 3) == FK_BOUNDARY) { bl.push_back((int)(le * 2 * nd + lf)); h->bd_ordinal.push_back(c.nbr); }
            }
        // sort by ordinal
        std::vector<size_t> ord(bl.size());
        for (size_t i = 0; i < ord.size(); i++) ord[i] = i;
        std::sort(ord.begin(), ord.end(), [&](size_t a, size_t b) { return h->bd_ordinal[a] < h->bd_ordinal[b]; });
        std::vector<int> bl2(bl.size());
        std::vector<int64_t> bo2(bl.size());
        for (size_t i = 0; i < ord.size(); i++) { bl2[i] = bl[ord[i]]; bo2[i] = h->bd_ordinal[ord[i]]; }
        h->bd_ordinal.swap(bo2);
        H_TRY(upload(&h->bd_list, bl2));
        H_TRY(cudaMalloc((void **)&h->bd_traces, sizeof(double) * std::max<size_t>(bl2.size() * h->nfp * h->nv, 1)));
        h->bc_rows = d->nbound > 0 ? d->bc_offsets[d->nbound] * (int64_t)h->nfp : 0;
    }
    H_TRY(upload(&h->send_list, send_list));
    H_TRY(upload(&h->interior_list, interior));
    H_TRY(upload(&h->boundary_list, boundary));
    const size_t state_bytes = sizeof(double) * (size_t)h->ndof * h->nv;
    H_TRY(cudaMalloc((void **)&h->u[0], state_bytes));
    H_TRY(cudaMalloc((void **)&h->u[1], state_bytes));
    H_TRY(cudaMalloc((void **)&h->tmp, state_bytes));
    H_TRY(cudaMalloc((void **)&h->k, state_bytes));
    const size_t trace_bytes = sizeof(double) * (size_t)h->ne_local * 2 * h->nfp * h->nv;
    if (!h->colloc) {
        const size_t all_bytes = sizeof(double) * (size_t)h->ne_local * h->nfaces * h->nfp * h->nv;
        H_TRY(cudaMalloc((void **)&h->tr_all, all_bytes));
        H_TRY(cudaMemset(h->tr_all, 0, all_bytes));
    }
    {
        // default: no x-face trace array while the state fits in L2 (the face kernel's extra reads of u
        // are L2 hits there: config 2 -4.6 % per stage), the array beyond (config 4: +2.7 % without it)
        const char *x = std::getenv("FLOU_B200_XTRACE");
        h->keep_xtraces = x ? x[0] != '0' : (state_bytes > ((size_t)48 << 20) && !large_single_gpu_state(d));
    }
    if (uses_xtraces(h)) {
        H_TRY(cudaMalloc((void **)&h->tr[0], trace_bytes));
        H_TRY(cudaMalloc((void **)&h->tr[1], trace_bytes));
        H_TRY(cudaMemset(h->tr[0], 0, trace_bytes));
        H_TRY(cudaMemset(h->tr[1], 0, trace_bytes));
    }
    H_TRY(cudaMemset(h->u[0], 0, state_bytes));
    H_TRY(cudaMemset(h->u[1], 0, state_bytes));
    H_TRY(cudaMemset(h->tmp, 0, state_bytes));
    H_TRY(cudaMemset(h->k, 0, state_bytes));
    const size_t halo_bytes = sizeof(double) * std::max<size_t>((size_t)h->nghost * h->nfp * h->nv, 1);
    H_TRY(cudaMalloc((void **)&h->ghost, halo_bytes));
    H_TRY(cudaMalloc((void **)&h->sendbuf, halo_bytes));
    H_TRY(cudaMalloc((void **)&h->status, sizeof(int)));
    H_TRY(cudaMemset(h->status, 0, sizeof(int)));
    H_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    H_TRY(cudaStreamCreateWithFlags(&h->comm_stream, cudaStreamNonBlocking));
    H_TRY(cudaEventCreateWithFlags(&h->ev_emit, cudaEventDisableTiming));
    H_TRY(cudaEventCreateWithFlags(&h->ev_recv, cudaEventDisableTiming));
    H_TRY(cudaEventCreate(&h->ev_t0));
    H_TRY(cudaEventCreate(&h->ev_t1));
    H_TRY(h->stage->prepare());
#undef H_TRY
    P.jac = h->jac; P.metric = h->metric; P.fjac = h->fjac; P.frames = h->frames;
    P.sub_frames = h->sub_frames; P.sub_jac = h->sub_jac;
    P.faceid = h->faceid; P.conn = h->conn;
    P.faces = h->faces; P.econn = h->econn; P.Fn = h->Fn; P.split_faces = h->split_faces ? 1 : 0;
    P.bc_kind = h->bc_kind; P.bc_state = h->bc_state; P.bc_table = h->bc_table;
    P.ghost = h->ghost;
    P.ndof = h->ndof;
    P.status = h->status;
    {
        const char *e = std::getenv("FLOU_B200_FACE_REVERSE");
        P.face_reverse = (e && e[0] == '0') ? 0 : 1;
    }
    // the memsets and table uploads above ran on the legacy default stream, which the handle's
    // non-blocking streams never synchronise with: finish them before the first upload / launch
    {
        const cudaError_t es = cudaDeviceSynchronize();
        if (es != cudaSuccess) {
            const std::string m = std::string("cudaDeviceSynchronize: ") + cudaGetErrorString(es);
            flou_b200_destroy(h);
            return fail(FLOU_B200_ECUDA, m);
        }
    }
    *out = h;
    return FLOU_B200_OK;
}

int32_t flou_b200_destroy(flou_b200_handle *h)
{
    if (!h) return FLOU_B200_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->comm_stream) cudaStreamSynchronize(h->comm_stream);
    destroy_graph(h);
    if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
    void *ptrs[] = {h->u[0], h->u[1], h->tr[0], h->tr[1], h->tr_all, h->tmp, h->k, h->conn, h->faceid, h->jac, h->metric, h->fjac,
                    h->frames, h->faces, h->econn, h->Fn, h->elem_dx, h->dt_bits, h->bc_kind, h->bc_state, h->bc_table, h->status, h->d_lm, h->d_lp,
                    h->ghost, h->sendbuf, h->send_list, h->interior_list, h->boundary_list, h->w_nodes, h->mon_partial,
                    h->sub_frames, h->sub_jac, h->source, h->bd_list, h->bd_traces};
    for (void *p : ptrs) if (p) cudaFree(p);
    for (cudaEvent_t e : h->prof_events) cudaEventDestroy(e);
    if (h->ev_emit) cudaEventDestroy(h->ev_emit);
    if (h->ev_recv) cudaEventDestroy(h->ev_recv);
    if (h->ev_t0) cudaEventDestroy(h->ev_t0);
    if (h->ev_t1) cudaEventDestroy(h->ev_t1);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->comm_stream) cudaStreamDestroy(h->comm_stream);
    delete h;
    return FLOU_B200_OK;
}

int32_t flou_b200_upload_state(flou_b200_handle *h, const double *Q)
{
    if (!h || !Q) return fail(FLOU_B200_EINVAL, "null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaMemsetAsync(h->status, 0, sizeof(int), h->stream));     // a new state: flags start over
    CUDA_TRY(cudaMemcpyAsync(h->u[h->cur], Q, sizeof(double) * (size_t)h->ndof * h->nv,
                             cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->traces_valid = false;
    return FLOU_B200_OK;
}

int32_t flou_b200_download_state(flou_b200_handle *h, double *Q)
{
    if (!h || !Q) return fail(FLOU_B200_EINVAL, "null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaMemcpyAsync(Q, h->u[h->cur], sizeof(double) * (size_t)h->ndof * h->nv,
                             cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return FLOU_B200_OK;
}

int32_t flou_b200_max_dt(flou_b200_handle *h, const double *Q, double cfl, double *dt)
{
    if (!h || !dt) return fail(FLOU_B200_EINVAL, "null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    double *state = nullptr;
    if (int32_t rc = query_state(h, Q, &state)) return rc;
    const unsigned long long inf_bits = 0x7ff0000000000000ULL;
    CUDA_TRY(cudaMemcpyAsync(h->dt_bits, &inf_bits, sizeof(inf_bits), cudaMemcpyHostToDevice, h->stream));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
    const int threads = 256;
    const int64_t want = (h->ndof + threads - 1) / threads;
    const int grid = (int)std::min<int64_t>(want, (int64_t)sms * 8);
    max_dt_kernel<<<grid, threads, 0, h->stream>>>(state, h->ndof, h->npts, h->nd,
                                                   h->equation == FLOU_B200_EQ_EULER, h->gamma, h->anorm,
                                                   cfl, h->elem_dx, h->cart_dx, h->dt_bits);
    CUDA_TRY(cudaGetLastError());
    h->launches += 1;
    if (h->nranks > 1) {
        if (!h->comm) return fail(FLOU_B200_EINVAL, "partitioned handle used before flou_b200_comm_init");
        // the bit pattern of a positive double is monotone: reduce it as float64 with ncclMin
        NCCL_TRY(g_nccl.AllReduce(h->dt_bits, h->dt_bits, 1, ncclFloat64, 3 /* ncclMin */, h->comm, h->stream));
    }
    double out = 0.0;
    CUDA_TRY(cudaMemcpyAsync(&out, h->dt_bits, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    *dt = out;
    return FLOU_B200_OK;
}

int32_t flou_b200_monitor(flou_b200_handle *h, int32_t kind, const double *Q, double *value)
{
    if (!h || !value) return fail(FLOU_B200_EINVAL, "null argument");
    if (h->equation != FLOU_B200_EQ_EULER)
        return fail(FLOU_B200_EINVAL, "monitors are defined for the Euler equations (list_monitors)");
    if (kind != FLOU_B200_MONITOR_KINETIC_ENERGY && kind != FLOU_B200_MONITOR_ENTROPY)
        return fail(FLOU_B200_EINVAL, "unknown monitor");
    CUDA_TRY(cudaSetDevice(h->device));
    double *state = nullptr;
    if (int32_t rc = query_state(h, Q, &state)) return rc;
    const int threads = 256;
    const int64_t want = (h->ndof + threads - 1) / threads;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, h->mon_blocks));
    double *out = h->mon_partial + h->mon_blocks;
    monitor_kernel<<<grid, threads, 0, h->stream>>>(state, h->ndof, h->npts, h->nd, kind, h->gamma,
                                                    h->w_nodes, h->base.jac, h->base.cjac, h->mon_partial);
    CUDA_TRY(cudaGetLastError());
    monitor_reduce_kernel<<<1, 256, 0, h->stream>>>(h->mon_partial, grid, out);
    CUDA_TRY(cudaGetLastError());
    h->launches += 2;
    if (h->nranks > 1) {
        if (!h->comm) return fail(FLOU_B200_EINVAL, "partitioned handle used before flou_b200_comm_init");
        NCCL_TRY(g_nccl.AllReduce(out, out, 1, ncclFloat64, 0 /* ncclSum */, h->comm, h->stream));
    }
    double v = 0.0;
    CUDA_TRY(cudaMemcpyAsync(&v, out, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    *value = v;
    return FLOU_B200_OK;
}

int32_t flou_b200_zhang_shu(flou_b200_handle *h, double *Q, double minval)
{
    if (!h) return fail(FLOU_B200_EINVAL, "null handle");
    if (h->equation != FLOU_B200_EQ_EULER)
        return fail(FLOU_B200_EINVAL, "the Zhang-Shu limiter is defined for the Euler equations (list_limiters)");
    CUDA_TRY(cudaSetDevice(h->device));
    const size_t bytes = sizeof(double) * (size_t)h->ndof * h->nv;
    // explicit host state: limited in the scratch buffer and copied back; the device-resident state
    // (and its traces) are untouched.  Q = NULL: the device-resident state is limited in place.
    double *state = nullptr;
    if (int32_t rc = query_state(h, Q, &state)) return rc;
    const bool tv = h->traces_valid;
    if (int32_t rc = launch_zhang_shu(h, state, minval)) return rc;
    if (Q) {
        h->traces_valid = tv;
        CUDA_TRY(cudaMemcpyAsync(Q, state, bytes, cudaMemcpyDeviceToHost, h->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return FLOU_B200_OK;
}

int32_t flou_b200_set_stage_limiter(flou_b200_handle *h, int32_t enable, double minval)
{
    if (!h) return fail(FLOU_B200_EINVAL, "null handle");
    if (enable && h->equation != FLOU_B200_EQ_EULER)
        return fail(FLOU_B200_EINVAL, "the Zhang-Shu limiter is defined for the Euler equations (list_limiters)");
    h->stage_limiter = enable != 0;
    h->limiter_minval = enable ? minval : 0.0;
    return FLOU_B200_OK;
}

int32_t flou_b200_synchronize(flou_b200_handle *h)
{
    if (!h) return fail(FLOU_B200_EINVAL, "null handle");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->comm_stream));
    return FLOU_B200_OK;
}

int32_t flou_b200_status(flou_b200_handle *h, int32_t *flags)
{
    if (!h || !flags) return fail(FLOU_B200_EINVAL, "null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    int f = 0;
    CUDA_TRY(cudaMemcpy(&f, h->status, sizeof(int), cudaMemcpyDeviceToHost));
    *flags = f;
    return FLOU_B200_OK;
}

int32_t flou_b200_rhs(flou_b200_handle *h, const double *Q, double *dQ, double t)
{
    (void)t;   // no time-dependent boundary data or sources live on the device
    if (!h) return fail(FLOU_B200_EINVAL, "null handle");
    CUDA_TRY(cudaSetDevice(h->device));
    const size_t bytes = sizeof(double) * (size_t)h->ndof * h->nv;
    if (Q) {
        CUDA_TRY(cudaMemcpyAsync(h->u[h->cur], Q, bytes, cudaMemcpyHostToDevice, h->stream));
        h->traces_valid = false;
    }
    const int32_t rc = run_pass(h, MODE_RHS, 0.0, 0.0, 0.0, h->u[h->cur], h->u[h->cur ^ 1]);
    if (rc) return rc;
    if (dQ) CUDA_TRY(cudaMemcpyAsync(dQ, h->k, bytes, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return FLOU_B200_OK;
}

int32_t flou_b200_lsrk2n_advance(flou_b200_handle *h, int32_t nstages, const double *A,
                                 const double *B, const double *c, double dt, double t0,
                                 int64_t nsteps)
{
    (void)c; (void)t0;
    if (!h || !A || !B || nstages < 1 || nstages > 16 || nsteps < 0)
        return fail(FLOU_B200_EINVAL, "bad RK arguments");
    CUDA_TRY(cudaSetDevice(h->device));
    // CUDA graph of two steps (2*nstages passes bring u back to the same ping-pong buffer).
    // Partitioned handles capture the halo exchange with it: the comm stream forks from the compute
    // stream at the pack event and joins it at the receive event inside every pass, and NCCL's
    // send/recv are capturable (every rank captures and replays the same sequence).
    // Measured (profiles/r2mg, r2mg8): direct launches are as fast at 2 GPUs (27.6 vs 27.3 GDOF/s) and
    // 2.5 % faster at 8 (106.2 vs 103.6), so the captured exchange is opt-in: FLOU_B200_MG_GRAPH=1.
    static const bool mg_graph = [] { const char *e = std::getenv("FLOU_B200_MG_GRAPH"); return e && e[0] == '1'; }();
    const bool use_graph = !(h->flags & FLOU_B200_FLAG_NO_GRAPH) && nsteps >= 2 && !h->profile &&
                           (h->nranks == 1 || h->nghost == 0 || (h->comm && mg_graph));
    int64_t done = 0;
    if (use_graph && !h->traces_valid && uses_xtraces(h)) {
        // bring the traces of the current state up to date outside the captured region
        CUDA_TRY(h->emit->launch(h->u[h->cur], h->ndof, nullptr, (int)(h->ne_local * 2), 2,
                                 h->base.colloc, h->d_lm, h->d_lp, h->tr[h->cur], h->stream));
        h->launches += 1;
        h->traces_valid = h->colloc;     // Gauss nodes: the captured passes emit their own
    }
    if (use_graph) {
        std::vector<double> key;
        key.push_back((double)nstages); key.push_back(dt);
        key.insert(key.end(), A, A + nstages);
        key.insert(key.end(), B, B + nstages);
        key.push_back(h->stage_limiter ? 1.0 : 0.0); key.push_back(h->limiter_minval);
        const int gi = h->cur;           // two steps bring u back to the buffer they started in
        // The graphs of BOTH ping-pong parities are built at the first use of a tableau / dt (a run
        // whose leftover odd step flips the parity would otherwise instantiate the second one in the
        // middle of a later, possibly timed, call) and uploaded to the device right away.
        const int cur0 = h->cur;
        for (int gj : {gi, gi ^ 1}) {
            if (h->graph[gj] && key == h->graph_key[gj]) continue;
            if (h->graph[gj]) cudaGraphExecDestroy(h->graph[gj]);
            h->graph[gj] = nullptr;
            cudaGraph_t g = nullptr;
            const int64_t l0 = h->launches;
            const bool tv0 = h->traces_valid;
            // Gauss nodes / stage limiter: every captured pass starts with its own trace emit
            if (!h->colloc || h->stage_limiter) h->traces_valid = false;
            h->cur = gj;
            CUDA_TRY(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
            const int32_t rc = run_steps_direct(h, nstages, A, B, dt, 2);
            cudaError_t e = cudaStreamEndCapture(h->stream, &g);
            h->cur = cur0;
            h->graph_launches_per_replay = h->launches - l0;
            h->launches = l0;
            if (gj != gi) h->traces_valid = tv0;
            if (rc) { if (g) cudaGraphDestroy(g); return rc; }
            CUDA_TRY(e);
            e = cudaGraphInstantiate(&h->graph[gj], g, 0);
            cudaGraphDestroy(g);
            CUDA_TRY(e);
            CUDA_TRY(cudaGraphUpload(h->graph[gj], h->stream));
            h->graph_key[gj] = key;
        }
        while (nsteps - done >= 2) {
            CUDA_TRY(cudaGraphLaunch(h->graph[gi], h->stream));
            h->launches += h->graph_launches_per_replay;
            done += 2;
            h->traces_valid = h->colloc && !h->stage_limiter;
        }
    }
    return run_steps_direct(h, nstages, A, B, dt, nsteps - done);
}

int32_t flou_b200_timeintegrate(flou_b200_handle *h, double *Q, int32_t nstages, const double *A,
                                const double *B, const double *c, double dt, double t0,
                                int64_t nsteps)
{
    if (!h || !Q) return fail(FLOU_B200_EINVAL, "null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    const size_t bytes = sizeof(double) * (size_t)h->ndof * h->nv;
    CUDA_TRY(cudaMemsetAsync(h->status, 0, sizeof(int), h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->u[h->cur], Q, bytes, cudaMemcpyHostToDevice, h->stream));
    h->traces_valid = false;
    const int32_t rc = flou_b200_lsrk2n_advance(h, nstages, A, B, c, dt, t0, nsteps);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(Q, h->u[h->cur], bytes, cudaMemcpyDeviceToHost, h->stream));
    int f = 0;
    CUDA_TRY(cudaMemcpyAsync(&f, h->status, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (f & 1) return fail(FLOU_B200_EDOMAIN, "non-positive density/pressure or NaN (Simulation crashed!)");
    return FLOU_B200_OK;
}

int32_t flou_b200_set_source(flou_b200_handle *h, const double *S)
{
    if (!h) return fail(FLOU_B200_EINVAL, "null handle");
    CUDA_TRY(cudaSetDevice(h->device));
    const size_t bytes = sizeof(double) * (size_t)h->ndof * h->nv;
    const double *before = h->base.source;
    if (S) {
        if (!h->source) CUDA_TRY(cudaMalloc((void **)&h->source, bytes));
        CUDA_TRY(cudaMemcpyAsync(h->source, S, bytes, cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));      // the host buffer is borrowed for this call only
        h->base.source = h->source;
    } else {
        h->base.source = nullptr;
    }
    if (h->base.source != before) destroy_graph(h);      // captured kernel parameters hold the pointer
    return FLOU_B200_OK;
}

int32_t flou_b200_set_bc_table(flou_b200_handle *h, const double *table)
{
    if (!h || !table) return fail(FLOU_B200_EINVAL, "null argument");
    if (h->bc_rows <= 0) return fail(FLOU_B200_EINVAL, "the discretisation has no boundary faces");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaMemcpyAsync(h->bc_table, table, sizeof(double) * (size_t)h->bc_rows * h->nv,
                             cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return FLOU_B200_OK;
}

int32_t flou_b200_boundary_traces(flou_b200_handle *h, double *Qin, int64_t *ordinals, int64_t *count)
{
    if (!h) return fail(FLOU_B200_EINVAL, "null handle");
    const int64_t n = (int64_t)h->bd_ordinal.size();
    if (count) *count = n;
    if (ordinals) std::copy(h->bd_ordinal.begin(), h->bd_ordinal.end(), ordinals);
    if (!Qin || n == 0) return FLOU_B200_OK;
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(h->emit->launch(h->u[h->cur], h->ndof, h->bd_list, (int)n, h->nfaces, h->base.colloc,
                             h->d_lm, h->d_lp, h->bd_traces, h->stream));
    h->launches += 1;
    CUDA_TRY(cudaMemcpyAsync(Qin, h->bd_traces, sizeof(double) * (size_t)n * h->nfp * h->nv,
                             cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return FLOU_B200_OK;
}

int32_t flou_b200_project_equispaced(flou_b200_handle *h, const double *Q, int32_t neq, const double *node2eq,
                                     double *Qe)
{
    if (!h || !node2eq || !Qe) return fail(FLOU_B200_EINVAL, "null argument");
    if (neq < 1 || neq > 64) return fail(FLOU_B200_EINVAL, "number of equispaced nodes per direction out of range (1..64)");
    CUDA_TRY(cudaSetDevice(h->device));
    double *state = nullptr;
    if (int32_t rc = query_state(h, Q, &state)) return rc;
    int64_t nout = neq;
    for (int d = 1; d < h->nd; d++) nout *= neq;
    const size_t out_elems = (size_t)h->ne_local * nout * h->nv;
    // the RHS buffer doubles as the output when it is large enough (neq == np: same size as a state)
    double *out = h->k, *own = nullptr;
    if (out_elems > (size_t)h->ndof * h->nv) {
        CUDA_TRY(cudaMalloc((void **)&own, sizeof(double) * out_elems));
        out = own;
    }
    double *dM = nullptr;
    cudaError_t e = cudaMalloc((void **)&dM, sizeof(double) * (size_t)neq * h->np);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(dM, node2eq, sizeof(double) * (size_t)neq * h->np, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(h->ne_local, (int64_t)sms * 8));
        const size_t smem = sizeof(double) * ((size_t)neq * h->np + (size_t)h->npts);
        project_equispaced_kernel<<<grid, 256, smem, h->stream>>>(state, h->ndof, h->nv, h->nd, h->np, neq, dM,
                                                                  h->ne_local, out);
        e = cudaGetLastError();
        h->launches += 1;
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(Qe, out, sizeof(double) * out_elems, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(dM);
    cudaFree(own);
    CUDA_TRY(e);
    return FLOU_B200_OK;
}

int32_t flou_b200_lsrk2n_stage(flou_b200_handle *h, double A, double B, double dt, int32_t first)
{
    if (!h) return fail(FLOU_B200_EINVAL, "null handle");
    CUDA_TRY(cudaSetDevice(h->device));
    const int32_t rc = run_pass(h, first ? MODE_STAGE_FIRST : MODE_STAGE, A, B, dt, h->u[h->cur], h->u[h->cur ^ 1]);
    if (rc) return rc;
    h->cur ^= 1;
    if (h->stage_limiter) return launch_zhang_shu(h, h->u[h->cur], h->limiter_minval);
    return FLOU_B200_OK;
}

int64_t flou_b200_ndofs_local(const flou_b200_handle *h) { return h ? h->ndof : 0; }
void *flou_b200_stream(flou_b200_handle *h) { return h ? (void *)h->stream : nullptr; }
void *flou_b200_device_state(flou_b200_handle *h) { return h ? (void *)h->u[h->cur] : nullptr; }
int64_t flou_b200_kernel_launches(const flou_b200_handle *h) { return h ? h->launches : 0; }

int32_t flou_b200_kernel_info(flou_b200_handle *h, int32_t *grid_ctas, int32_t *threads,
                              int32_t *smem_bytes, int32_t *elems_per_cta_iter)
{
    if (!h) return fail(FLOU_B200_EINVAL, "null handle");
    CUDA_TRY(cudaSetDevice(h->device));
    if (grid_ctas) *grid_ctas = (h->line_kernel && h->stage->line_resident) ? h->stage->line_resident() : h->stage->resident();
    if (threads) *threads = h->line_kernel ? h->stage->line_t : h->stage->threads;
    if (smem_bytes) *smem_bytes = (int32_t)(h->line_kernel ? h->stage->line_smem : h->stage->smem);
    if (elems_per_cta_iter) *elems_per_cta_iter = h->line_kernel ? h->stage->line_e : h->stage->epb;
    return FLOU_B200_OK;
}

int32_t flou_b200_profile(flou_b200_handle *h, int32_t enable, float *ms_faces, float *ms_elements,
                          int64_t *npasses)
{
    if (!h) return fail(FLOU_B200_EINVAL, "null handle");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    double tf = 0.0, te = 0.0;
    const size_t n = h->prof_events.size() / 3;
    for (size_t i = 0; i < n; i++) {
        float a = 0.f, b = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&a, h->prof_events[3 * i], h->prof_events[3 * i + 1]));
        CUDA_TRY(cudaEventElapsedTime(&b, h->prof_events[3 * i + 1], h->prof_events[3 * i + 2]));
        tf += a; te += b;
    }
    for (cudaEvent_t e : h->prof_events) cudaEventDestroy(e);
    h->prof_events.clear();
    if (ms_faces) *ms_faces = (float)tf;
    if (ms_elements) *ms_elements = (float)te;
    if (npasses) *npasses = (int64_t)n;
    h->profile = enable != 0;
    return FLOU_B200_OK;
}

int32_t flou_b200_timer_start(flou_b200_handle *h)
{
    if (!h) return fail(FLOU_B200_EINVAL, "null handle");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaEventRecord(h->ev_t0, h->stream));
    return FLOU_B200_OK;
}

int32_t flou_b200_timer_stop(flou_b200_handle *h, float *ms)
{
    if (!h || !ms) return fail(FLOU_B200_EINVAL, "null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaEventRecord(h->ev_t1, h->stream));
    CUDA_TRY(cudaEventSynchronize(h->ev_t1));
    CUDA_TRY(cudaEventElapsedTime(ms, h->ev_t0, h->ev_t1));
    return FLOU_B200_OK;
}

int32_t flou_b200_pin_host(void *ptr, uint64_t bytes)
{
    if (!ptr) return fail(FLOU_B200_EINVAL, "null argument");
    CUDA_TRY(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterDefault));
    return FLOU_B200_OK;
}

int32_t flou_b200_unpin_host(void *ptr)
{
    if (!ptr) return fail(FLOU_B200_EINVAL, "null argument");
    CUDA_TRY(cudaHostUnregister(ptr));
    return FLOU_B200_OK;
}

int32_t flou_b200_nccl_unique_id(char id[128])
{
    if (!id) return fail(FLOU_B200_EINVAL, "null argument");
    if (!g_nccl.load()) return fail(FLOU_B200_ENCCL, "libnccl.so.2 not found");
    ncclUniqueId uid;
    NCCL_TRY(g_nccl.GetUniqueId(&uid));
    std::memcpy(id, uid.internal, 128);
    return FLOU_B200_OK;
}

int32_t flou_b200_comm_init(flou_b200_handle *h, const char id[128])
{
    if (!h || !id) return fail(FLOU_B200_EINVAL, "null argument");
    if (!g_nccl.load()) return fail(FLOU_B200_ENCCL, "libnccl.so.2 not found");
    CUDA_TRY(cudaSetDevice(h->device));
    ncclUniqueId uid;
    std::memcpy(uid.internal, id, 128);
    NCCL_TRY(g_nccl.CommInitRank(&h->comm, h->nranks, uid, h->rank));
    return FLOU_B200_OK;
}

}  // extern "C"
