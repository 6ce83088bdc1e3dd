/* flou_b200.h -- C ABI of the B200-native DGSEM right-hand side + low-storage RK library.
 *
 * Drop-in boundary for ONE hot path of Andres-MG/Flou.jl (reference paths are relative to
 * the reference repository root):
 *
 *   rhs!(dQ, Q, ::EquationConfig, t)          src/FlouSpatial/Equations/Hyperbolic.jl:31-69
 *   timeintegrate(Q0, disc, eq, solver, tf)   src/FlouTime/FlouTime.jl:34-54  (the 2N RK loop
 *                                             itself lives in OrdinaryDiffEq v6.49.1)
 *   MultielementDisc(mesh, std, eq, op, bcs)  src/FlouSpatial/MultielementDiscontinuous.jl:29-92
 *
 * The host (Julia through `ccall`; Python/ctypes in this repository's
 * tests and benchmark) keeps building meshes, standard regions, operators and boundary
 * conditions with Flou's own types and hands their plain arrays to `flou_b200_create`.
 * Everything after that runs in hand-written sm_100a CUDA kernels; there is no CPU
 * fallback -- every entry point fails with FLOU_B200_ECUDA when no device is usable.
 *
 * Conventions
 *   - all entry points return int32 status (0 = OK) and never abort/exit;
 *   - ids in connectivity tables are 1-BASED exactly as Flou stores them
 *     (src/FlouCommon/Mesh.jl:26-150); 0 in `eleminds` means "no element";
 *   - matrices are column-major (Julia `Matrix`): state Q is (ndofs, nv) so variable v of
 *     dof i is Q[i + ndofs*v]  (src/FlouSpatial/GlobalContainers.jl:19-29,92-105);
 *   - host pointers are borrowed for the duration of the call only;
 *   - one handle = one GPU, one compute stream, one communication stream; calls on a
 *     handle are stream-ordered and not thread-safe.
 */
#ifndef FLOU_B200_H
#define FLOU_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status codes */
#define FLOU_B200_OK       0
#define FLOU_B200_EINVAL   1   /* bad descriptor / argument (Julia: ArgumentError)          */
#define FLOU_B200_ECUDA    2   /* CUDA failure or no usable device                           */
#define FLOU_B200_ENCCL    3   /* NCCL failure                                               */
#define FLOU_B200_EDOMAIN  4   /* rho <= 0, p <= 0 or NaN met on device (Julia: DomainError,
                                  cf. FlouTime.jl:37-52 "Simulation crashed!")               */
#define FLOU_B200_EUNSUPPORTED 5 /* combination not compiled into the library               */

/* equations: src/FlouCommon/LinearAdvection.jl:16-18, src/FlouCommon/Euler.jl:16-24 */
#define FLOU_B200_EQ_LINEAR_ADVECTION 0
#define FLOU_B200_EQ_EULER            1

/* divergence operators: src/FlouSpatial/Equations/OpDivergence.jl:105 (Strong), :184 (Split) */
#define FLOU_B200_OP_STRONG 0
/* SplitDivOperator: on nodes with boundaries (GLL) any geometry; on Gauss nodes the entropy-
 * projected surface term (OpDivergence.jl:300-437) for the Euler equations (general geometry
 * needs the sub_frames / sub_jac tables) */
#define FLOU_B200_OP_SPLIT  1
/* HybridDivOperator(tpflux, numflux, blend) (OpDivergence.jl:452-477, volume term :557-612):
 * telescopic split form blended with sub-cell finite-volume fluxes (fvflux = numflux, as both
 * convenience constructors set it).  Euler; GLL nodes (volume form) or Gauss nodes (all-surface
 * form, :629-779); Cartesian sub-grids or, with the sub_frames / sub_jac tables, unstructured
 * elements. */
#define FLOU_B200_OP_HYBRID 2

/* numerical-flux structs: src/FlouSpatial/Interfaces.jl:16-23,
 * src/FlouSpatial/Equations/Euler.jl:167,228-231,303-306 */
#define FLOU_B200_FLUX_STDAVERAGE        0
#define FLOU_B200_FLUX_LXF               1
#define FLOU_B200_FLUX_CHANDRASEKHAR     2
#define FLOU_B200_FLUX_SCALARDISSIPATION 3
#define FLOU_B200_FLUX_MATRIXDISSIPATION 4

/* boundary-condition functors: src/FlouSpatial/Equations/Euler.jl:69-94,
 * src/FlouSpatial/FlouSpatial.jl:85-91 (GenericBC).  A GenericBC whose closure depends only
 * on the face-node coordinates is passed as a TABLE evaluated by the host at create time. */
#define FLOU_B200_BC_INFLOW  0   /* EulerInflowBC(Qext): bc_state row                        */
#define FLOU_B200_BC_OUTFLOW 1   /* EulerOutflowBC                                           */
#define FLOU_B200_BC_SLIP    2   /* EulerSlipBC                                              */
#define FLOU_B200_BC_TABLE   3   /* GenericBC(x -> Qext) tabulated per boundary-face node    */

/* geometry kinds */
#define FLOU_B200_GEOM_CARTESIAN 0  /* CartesianMesh: constant metrics from dx, frames chosen
                                       by the master's elempos (PhysicalRegions.jl:370-408,
                                       541-696)                                              */
#define FLOU_B200_GEOM_GENERAL   1  /* per-node jac/metric, per-face-node frames/jac
                                       (PhysicalRegions.jl:438-472, 797-971)                 */

typedef struct flou_b200_handle flou_b200_handle;

/* Everything `MultielementDisc` + `EquationConfig` hold that the hot path reads
 * (SURVEY.md section 8(a) row a15). */
typedef struct flou_b200_desc {
    int32_t struct_size;      /* sizeof(flou_b200_desc), ABI check                           */
    int32_t nd;               /* spatial dimension 1..3                                      */
    int32_t nv;               /* variables: 1 (advection) or nd+2 (Euler)                    */
    int32_t np;               /* nodes per direction (p+1), 2..8                             */
    int32_t equation;         /* FLOU_B200_EQ_*                                              */
    int32_t divop;            /* FLOU_B200_OP_*        (disc.operators[1])                   */
    int32_t tpflux;           /* Split-/HybridDivOperator.tpflux: STDAVERAGE | CHANDRASEKHAR */
    int32_t numflux;          /* operator.numflux kind                                       */
    int32_t numflux_avg;      /* numflux.avg for LxF / Scalar- / MatrixDissipation           */
    int32_t geometry;         /* FLOU_B200_GEOM_*                                            */
    double intensity;         /* numflux.intensity                                           */
    double gamma;             /* EulerEquation.gamma                                         */
    double a[3];              /* LinearAdvection.a                                           */
    double dx[3];             /* CartesianMesh.dx (GEOM_CARTESIAN)                           */

    /* mesh connectivity, GLOBAL tables in the reference's layout and numbering */
    int64_t ne;               /* nelements(mesh)                                             */
    int64_t nf;               /* nfaces(mesh)                                                */
    const int64_t *faceinds;  /* ne*2nd : mesh.elements.faceinds  (Mesh.jl:26-33)            */
    const int64_t *facepos;   /* ne*2nd : mesh.elements.facepos   1 = master, 2 = slave      */
    const int64_t *eleminds;  /* nf*2   : mesh.faces.eleminds     (Mesh.jl:91-99)            */
    const int64_t *elempos;   /* nf*2   : mesh.faces.elempos                                 */
    const uint8_t *orientation; /* nf   : mesh.faces.orientation                             */

    /* 1-D operators of the standard region (StdSegment.jl:76-89), np x np column-major */
    const double *D;          /* std.D   (HybridDivOperator; otherwise unused)               */
    const double *Ds;         /* std.Ds  = D - B                                             */
    const double *Dsharp;     /* std.D♯  = 2D - B                                            */
    const double *lminus;     /* std.l[1]                                                    */
    const double *lplus;      /* std.l[2]                                                    */
    const double *dgminus;    /* std.∂g[1]                                                   */
    const double *dgplus;     /* std.∂g[2]                                                   */
    const double *weights;    /* std.ω of the 1-D region (np): element volume for get_max_dt */

    /* GEOM_GENERAL only (NULL otherwise); indexed by GLOBAL dof / face-dof */
    const double *jac;        /* ne*npts            geometry.elements.jac                    */
    const double *metric;     /* ne*npts*nd*nd      geometry.elements.metric, each SMatrix
                                                    column-major: [c + nd*d] = Ja^d_c        */
    const double *fjac;       /* nf*nfp             geometry.faces.jac                       */
    const double *frames;     /* nf*nfp*3*nd        geometry.faces.frames: n, t, b           */

    /* boundary conditions in mesh.bdfaces order (MultielementDiscontinuous.jl:45-51) */
    int32_t nbound;
    const int32_t *bc_kind;   /* nbound                                                      */
    const int64_t *bc_offsets;/* nbound+1 offsets into bc_faces                              */
    const int64_t *bc_faces;  /* concatenated mesh.bdfaces (1-based face ids)                */
    const double *bc_state;   /* nbound*nv  row per boundary (INFLOW)                        */
    const double *bc_table;   /* (len(bc_faces)*nfp)*nv, row per boundary-face node (TABLE)  */

    /* element partition (new; the reference is shared-memory only): this handle owns the
     * contiguous range [elem_begin, elem_end) of 0-based global element indices, and
     * part_offsets[0..nranks] lists every rank's range start (part_offsets[nranks] = ne). */
    int64_t elem_begin, elem_end;
    int32_t rank, nranks;
    const int64_t *part_offsets;  /* NULL when nranks == 1 */

    int32_t device;           /* CUDA device ordinal                                         */
    int32_t flags;            /* FLOU_B200_FLAG_* */
    double blend;             /* HybridDivOperator.blend (OpDivergence.jl:464)               */
    /* GEOM_GENERAL with HybridDivOperator or SplitDivOperator on Gauss nodes (NULL otherwise):
     * geometry.subgrids (PhysicalRegions.jl:179-292) by GLOBAL element, direction, tensor-product
     * line and position along the line (np+1 sub-cell interfaces, tpdofs_subgrid order):
     *   sub_frames[((((e*nd + d)*nlines + k)*(np+1) + ii)*3 + r)*nd + c], r = n, t, b
     *   sub_jac   [ ((e*nd + d)*nlines + k)*(np+1) + ii]                                          */
    const double *sub_frames;
    const double *sub_jac;
} flou_b200_desc;

#define FLOU_B200_FLAG_NO_GRAPH 1  /* launch stages directly instead of replaying a CUDA graph */
#define FLOU_B200_FLAG_FUSED    2  /* single fused stage kernel (both neighbours recompute each
                                      face flux) instead of face-flux kernel + element kernel */
#define FLOU_B200_FLAG_NODE_KERNEL 4 /* two-kernel stage with the node-per-thread element kernel
                                      instead of the default line-per-thread one (A/B testing) */
#define FLOU_B200_FLAG_LINE_KERNEL 8 /* two-kernel stage with the line-per-thread element kernel even on
                                      meshes too small to fill the GPU (default there: fused kernel) */

/* ---- lifetime -------------------------------------------------------------------------- */
/* Replaces MultielementDisc(...) + construct_cache (Hyperbolic.jl:21-29): uploads tables,
 * allocates u (x2, ping-pong), tmp and k on the device. */
int32_t flou_b200_create(const flou_b200_desc *desc, flou_b200_handle **out);
int32_t flou_b200_destroy(flou_b200_handle *h);

/* ---- state ----------------------------------------------------------------------------- */
/* Q is the (ndofs_local, nv) column-major matrix of the OWNED elements (Q.data of
 * GlobalStateVector restricted to the rank's rows). */
int32_t flou_b200_upload_state(flou_b200_handle *h, const double *Q);
int32_t flou_b200_download_state(flou_b200_handle *h, double *Q);

/* ---- hot path -------------------------------------------------------------------------- */
/* rhs!(dQ, Q, p, t)  (Hyperbolic.jl:31-69).  Q == NULL: use the device-resident state;
 * dQ == NULL: leave the result in the device `k` buffer. Synchronises before returning. */
int32_t flou_b200_rhs(flou_b200_handle *h, const double *Q, double *dQ, double t);

/* OrdinaryDiffEq `LowStorageRK2N` steps on the device-resident state (a14 in SURVEY 8a):
 *   stage 1: tmp = dt*k;              u += B[0]*tmp
 *   stage s: tmp = A[s]*tmp + dt*k;   u += B[s]*tmp          (A[0] ignored)
 * each stage is ONE fused kernel pass (RHS + update).  Asynchronous: returns after
 * enqueueing; use flou_b200_synchronize / download / status to wait. */
int32_t flou_b200_lsrk2n_advance(flou_b200_handle *h, int32_t nstages, const double *A,
                                 const double *B, const double *c, double dt, double t0,
                                 int64_t nsteps);

/* timeintegrate(Q0, disc, eq, solver, tf; adaptive=false, dt, alias_u0=true)
 * (FlouTime.jl:34-54): upload Q, advance nsteps, download into the same buffer, report
 * FLOU_B200_EDOMAIN if the state left the admissible set. */
int32_t flou_b200_timeintegrate(flou_b200_handle *h, double *Q, int32_t nstages,
                                const double *A, const double *B, const double *c,
                                double dt, double t0, int64_t nsteps);

/* get_max_dt(q, disc, equation, cfl) (MultielementDiscontinuous.jl:162-178 with the
 * per-node rule of FlouCommon/Euler.jl:116-135 / LinearAdvection.jl:46-48): the global minimum
 * of cfl*dx/(|v| + c), dx = (volume/npts)^(1/nd), over the device-resident state (Q == NULL) or
 * an explicit host state, which is staged in a scratch buffer: the device-resident state of an
 * ongoing advance() is never replaced by a query (the same holds for _monitor and _zhang_shu);
 * an ncclAllReduce(min) joins the ranks of a partitioned run.  Row f1. */
int32_t flou_b200_max_dt(flou_b200_handle *h, const double *Q, double cfl, double *dt);

/* ---- monitors and limiters ------------------------------------------ */
/* get_monitor(disc, eq, :kinetic_energy | :entropy) (src/FlouSpatial/Equations/Euler.jl:541-593):
 * sum over the elements of integrate(f(Q_i), geom) with f = kinetic_energy
 * (FlouCommon/Euler.jl:162-175) or math_entropy (:213-217), as a device reduction over the
 * device-resident state (Q == NULL) or an uploaded one; ncclAllReduce(sum) across ranks. */
#define FLOU_B200_MONITOR_KINETIC_ENERGY 0
#define FLOU_B200_MONITOR_ENTROPY        1
int32_t flou_b200_monitor(flou_b200_handle *h, int32_t kind, const double *Q, double *value);
/* get_limiter(disc, eq, :zhang_shu, minval) (Equations/Euler.jl:597-660): positivity limiter of
 * Zhang & Shu, element by element.  Q == NULL: the device-resident state, in place; otherwise Q
 * (host, owned rows) is staged in a scratch buffer, limited and written back (the device-resident
 * state is not touched). */
int32_t flou_b200_zhang_shu(flou_b200_handle *h, double *Q, double minval);
/* ORK256(stage_limiter! = get_limiter_callback(dg, eq, :zhang_shu, minval)) as in
 * examples/src/3D_Euler.jl:76-80: flou_b200_lsrk2n_advance / _timeintegrate apply the limiter to
 * the state after every RK stage, on the device. */
int32_t flou_b200_set_stage_limiter(flou_b200_handle *h, int32_t enable, double minval);

/* ---- source term and boundary data that change between stages ---
 * apply_sourceterm! (MultielementDiscontinuous.jl:139-146; default no-op :75-79): the source
 * S(Q_i, x_i, t) is TABULATED per node by the host -- it is a user closure in the reference, which
 * cannot run on the device -- and added to dQ after the mass matrix by the stage kernels.
 * S: host (ndofs_local, nv) column-major, copied before the call returns; NULL removes the source.
 * A position-only source is set once; one that depends on t or Q is re-tabulated by the host
 * between stages (flou_b200_lsrk2n_stage below; Q through flou_b200_download_state). */
int32_t flou_b200_set_source(flou_b200_handle *h, const double *S);
/* GenericBC closures bc(Qin, x, frame, t, eq) (Interfaces.jl:44-48, FlouSpatial.jl:85-91) that
 * read Qin, the frame or the time: the host re-tabulates the exterior state of every boundary-face
 * node between stages.  table: same layout as flou_b200_desc.bc_table (all rows). */
int32_t flou_b200_set_bc_table(flou_b200_handle *h, const double *table);
/* Interior traces Qin of the device-resident state at the nodes of the boundary faces this handle
 * owns: Qin[(j*nv + v)*nfp + k] for the j-th owned boundary face (k in the element's face-dof
 * order), ordinals[j] = its position m in the concatenated bc_faces (the row block m*nfp.. of
 * bc_table); *count = number of owned boundary faces.  Any of Qin / ordinals / count may be NULL. */
int32_t flou_b200_boundary_traces(flou_b200_handle *h, double *Qin, int64_t *ordinals, int64_t *count);
/* ONE low-storage stage  tmp = A*tmp + dt*k(u);  u += B*tmp  (first != 0: tmp = dt*k) on the
 * device-resident state, followed by the stage limiter when one is set: the unit the host loops
 * over when source or boundary data change from stage to stage. */
int32_t flou_b200_lsrk2n_stage(flou_b200_handle *h, double A, double B, double dt, int32_t first);
/* Snapshot path: the state projected to equispaced nodes, element by element, what
 * pointdata2VTKHDF (src/FlouSpatial/IO.jl:78-97) computes with project2equispaced!
 * (StdRegions/StdRegions.jl:232-238) before FlouBiz.add_solution! writes it.
 * Q: host state (ndofs_local, nv) column-major, or NULL = the device-resident state.
 * node2eq: the 1-D interpolation matrix std.node2eq of the segment, row-major [neq][np]
 * (interp_matrix(range(-1, 1, neq), basis), StdSegment.jl:56-58); quads and hexes use its
 * Kronecker products (StdQuad.jl:52-53, StdHex.jl:57-61), applied here as nested 1-D sums.
 * Qe: host, Qe[p + npoints*v] with npoints = nelements_local * neq^nd and p = element*neq^nd +
 * equispaced node (x fastest): one contiguous vector per variable, as the reference appends them. */
int32_t flou_b200_project_equispaced(flou_b200_handle *h, const double *Q, int32_t neq, const double *node2eq,
                                     double *Qe);

int32_t flou_b200_synchronize(flou_b200_handle *h);
/* sticky device flags: bit 0 = non-positive density/pressure or NaN seen since the last
 * flou_b200_upload_state / flou_b200_timeintegrate (both clear them) */
int32_t flou_b200_status(flou_b200_handle *h, int32_t *flags);
const char *flou_b200_last_error(void);

/* ---- introspection (benchmarks, tests) --------------------------------------------------- */
int64_t flou_b200_ndofs_local(const flou_b200_handle *h);
void *flou_b200_stream(flou_b200_handle *h);         /* cudaStream_t of the compute stream   */
void *flou_b200_device_state(flou_b200_handle *h);   /* device pointer of the current u      */
int64_t flou_b200_kernel_launches(const flou_b200_handle *h); /* stage/rhs/halo kernels so far */
/* launch geometry of the fused stage kernel: persistent grid size (resident CTAs), threads
 * per CTA, dynamic shared memory per CTA, elements a CTA processes per loop iteration */
int32_t flou_b200_kernel_info(flou_b200_handle *h, int32_t *grid_ctas, int32_t *threads,
                              int32_t *smem_bytes, int32_t *elems_per_cta_iter);
/* Per-kernel CUDA-event timing of the two-kernel stage (single-rank handles): returns the time
 * accumulated since the previous call in the face-flux kernel and in the element kernel over
 * `npasses` passes, then switches the instrumentation on/off (graph replay is off while on). */
int32_t flou_b200_profile(flou_b200_handle *h, int32_t enable, float *ms_faces, float *ms_elements,
                          int64_t *npasses);
/* CUDA-event stopwatch on the compute stream (kernel timing for the roofline figure) */
int32_t flou_b200_timer_start(flou_b200_handle *h);
int32_t flou_b200_timer_stop(flou_b200_handle *h, float *ms);
/* page-lock a host buffer the caller owns (e.g. Julia's Q.data) so uploads/downloads run at
 * full PCIe rate; optional */
int32_t flou_b200_pin_host(void *ptr, uint64_t bytes);
int32_t flou_b200_unpin_host(void *ptr);
int32_t flou_b200_device_count(void);
int32_t flou_b200_supported(int32_t nd, int32_t np, int32_t equation, int32_t divop,
                            int32_t tpflux, int32_t geometry);

/* ---- multi-GPU (one process per GPU) ----------------------------------------------------- */
/* Halo face traces travel with ncclSend/ncclRecv over NVLink on the communication stream,
 * overlapped with the interior-element kernel.  `id` is an ncclUniqueId (128 bytes). */
/* Host-only (no GPU needed): the halo plan flou_b200_create would build for this descriptor.
 * Ghost slots are ordered by (peer rank, global face id), so slot s of this rank's send
 * buffer towards a peer is slot s of that peer's ghost buffer.  Output arrays may be NULL;
 * call once for the counts, then with buffers: peer_* need *npeers entries (<= nranks-1),
 * ghost_faces (1-based global face ids) and ghost_elemfaces (local element*2nd + local face)
 * need *nghost entries. */
int32_t flou_b200_partition_plan(const flou_b200_desc *desc, int64_t *nghost, int32_t *npeers,
                                 int32_t *peer_ranks, int64_t *peer_nslots,
                                 int64_t *ghost_faces, int32_t *ghost_elemfaces,
                                 int64_t *n_interior, int64_t *n_boundary);
int32_t flou_b200_nccl_unique_id(char id[128]);
int32_t flou_b200_comm_init(flou_b200_handle *h, const char id[128]);

#ifdef __cplusplus
}
#endif
#endif /* FLOU_B200_H */
