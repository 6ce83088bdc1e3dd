# FlouB200.jl -- drop-in shim: Flou.jl's own objects in, B200 kernels underneath.
#
# NOT runnable in the build container (no `julia` there); it is the binding a Flou
# maintainer adds next to the package (see INTEGRATION.md).  Everything above the `ccall`s is
# plain Flou API: the mesh, standard region, operators and boundary conditions are built by
# Flou's own constructors and only their arrays cross the C ABI (include/flou_b200.h).
#
#   rhs!(dQ, Q, p::EquationConfig{<:B200Disc}, t)      replaces Hyperbolic.jl:31-69
#   timeintegrate(Q0, ::B200Disc, eq, solver, tf; dt)  replaces FlouTime.jl:34-54 for the
#                                                      LowStorageRK2N solvers (ORK256,
#                                                      CarpenterKennedy2N54)
module FlouB200

using Flou
using Flou.FlouCommon: AbstractSpatialDiscretization, EquationConfig, CartesianMesh,
    LinearAdvection, EulerEquation, nvariables, spatialdim, nelements, nfaces
using Flou.FlouSpatial: MultielementDisc, StrongDivOperator, SplitDivOperator, HybridDivOperator,
    StdAverage, LxF,
    ChandrasekharAverage, ScalarDissipation, MatrixDissipation, EulerInflowBC, EulerOutflowBC,
    EulerSlipBC, GenericBC, ndofs
import Flou.FlouCommon: rhs!
import Flou.FlouTime: timeintegrate
using OrdinaryDiffEq: ORK256, CarpenterKennedy2N54

const lib = get(ENV, "FLOU_B200_LIB", "libflou_b200.so")

# mirrors `flou_b200_desc` field by field (include/flou_b200.h)
struct Desc
    struct_size::Int32; nd::Int32; nv::Int32; np::Int32
    equation::Int32; divop::Int32; tpflux::Int32; numflux::Int32; numflux_avg::Int32
    geometry::Int32
    intensity::Float64; gamma::Float64
    a::NTuple{3,Float64}; dx::NTuple{3,Float64}
    ne::Int64; nf::Int64
    faceinds::Ptr{Int64}; facepos::Ptr{Int64}; eleminds::Ptr{Int64}; elempos::Ptr{Int64}
    orientation::Ptr{UInt8}
    D::Ptr{Float64}; Ds::Ptr{Float64}; Dsharp::Ptr{Float64}
    lminus::Ptr{Float64}; lplus::Ptr{Float64}; dgminus::Ptr{Float64}; dgplus::Ptr{Float64}
    weights::Ptr{Float64}
    jac::Ptr{Float64}; metric::Ptr{Float64}; fjac::Ptr{Float64}; frames::Ptr{Float64}
    nbound::Int32
    bc_kind::Ptr{Int32}; bc_offsets::Ptr{Int64}; bc_faces::Ptr{Int64}
    bc_state::Ptr{Float64}; bc_table::Ptr{Float64}
    elem_begin::Int64; elem_end::Int64
    rank::Int32; nranks::Int32; part_offsets::Ptr{Int64}
    device::Int32; flags::Int32
    blend::Float64
    sub_frames::Ptr{Float64}; sub_jac::Ptr{Float64}
end

# the hand-written mirror above against the header's offsetof table (generated, desc_offsets.jl):
# a field added to, or reordered in, include/flou_b200.h without the same change here fails at
# module load instead of corrupting a descriptor
include("desc_offsets.jl")
@assert sizeof(Desc) == DESC_SIZEOF "FlouB200.Desc does not match flou_b200_desc (sizeof)"
@assert fieldcount(Desc) == length(DESC_OFFSETS) "FlouB200.Desc does not match flou_b200_desc (field count)"
for (i, (name, off)) in enumerate(DESC_OFFSETS)
    @assert fieldname(Desc, i) == name "FlouB200.Desc field $i is $(fieldname(Desc, i)), header has $name"
    @assert fieldoffset(Desc, i) == off "FlouB200.Desc.$name at offset $(fieldoffset(Desc, i)), header has $off"
end

fluxkind(::StdAverage) = Int32(0)
fluxkind(::LxF) = Int32(1)
fluxkind(::ChandrasekharAverage) = Int32(2)
fluxkind(::ScalarDissipation) = Int32(3)
fluxkind(::MatrixDissipation) = Int32(4)
avgkind(f) = hasproperty(f, :avg) ? fluxkind(f.avg) : Int32(0)
intensity(f) = hasproperty(f, :intensity) ? Float64(f.intensity) : 0.0

bckind(::EulerInflowBC) = Int32(0)
bckind(::EulerOutflowBC) = Int32(1)
bckind(::EulerSlipBC) = Int32(2)
bckind(::GenericBC) = Int32(3)      # tabulated at the boundary-face nodes below

function check(rc)
    rc == 0 && return
    msg = unsafe_string(ccall((:flou_b200_last_error, lib), Cstring, ()))
    rc == 1 || rc == 5 ? throw(ArgumentError(msg)) :
    rc == 4 ? throw(DomainError(NaN, msg)) : error("flou_b200 error $rc: $msg")
end

"""
    B200Disc(disc::MultielementDisc, equation; device=0)

Wraps a discretisation built by Flou itself and uploads its tables to the GPU.
"""
mutable struct B200Disc{ND,RT,D<:MultielementDisc{ND,RT}} <: AbstractSpatialDiscretization{ND,RT}
    disc::D
    handle::Ptr{Cvoid}
end

function B200Disc(disc::MultielementDisc{ND,RT}, equation; device=0) where {ND,RT}
    RT === Float64 || throw(ArgumentError("the B200 path is fp64 only"))
    (; mesh, std, geometry, bcs) = disc
    op = disc.operators[1]
    nv = nvariables(equation)
    cart = mesh isa CartesianMesh
    # connectivity exactly as Flou stores it (1-based, Mesh.jl:26-150)
    faceinds = Int64.(mesh.elements.faceinds); facepos = Int64.(mesh.elements.facepos)
    eleminds = Int64.(mesh.faces.eleminds);    elempos = Int64.(mesh.faces.elempos)
    orientation = mesh.faces.orientation
    # operators: Julia matrices are column-major, which is what the ABI expects
    D, Ds, Dsharp = std.D, std.Ds, std.D♯
    lm, lp = std.l; dgm, dgp = std.∂g
    w1d = ND == 1 ? std.ω : (ND == 2 ? std.face.ω : std.edge.ω)
    # general geometry tables (unused for CartesianMesh)
    jac = cart ? Float64[] : geometry.elements.jac
    metric = cart ? Float64[] : collect(reinterpret(Float64, geometry.elements.metric))
    fjac = cart ? Float64[] : geometry.faces.jac
    frames = cart ? Float64[] : collect(reinterpret(Float64, geometry.faces.frames))
    # geometry.subgrids (PhysicalRegions.jl:179-292) re-laid by (element, direction, tensor-product
    # line, position along the line) for the hybrid operator / the Gauss-node split form on
    # unstructured meshes: frames rows n, t, b (C order), then the sub-cell face Jacobians
    needs_sub = !cart && Flou.FlouSpatial.requires_subgrid(op, std)
    np1 = size(D, 1) + 1
    nlines = ND == 1 ? 1 : (ND == 2 ? size(D, 1) : size(D, 1)^2)
    sub_frames = needs_sub ? zeros(Float64, ND, 3, np1, nlines, ND, nelements(mesh)) : Float64[]
    sub_jac = needs_sub ? zeros(Float64, np1, nlines, ND, nelements(mesh)) : Float64[]
    if needs_sub
        dirs = ND == 1 ? (Val(1),) : (ND == 2 ? (Val(1), Val(2)) : (Val(1), Val(2), Val(3)))
        for ie in 1:nelements(mesh), (dir, vdir) in enumerate(dirs)
            sg = geometry.subgrids[ie]
            for (k, sinds) in enumerate(Flou.FlouSpatial.tpdofs_subgrid(std, vdir)), (ii, is) in enumerate(sinds)
                fr = sg.frames[dir][is]
                sub_frames[:, 1, ii, k, dir, ie] .= fr.n
                sub_frames[:, 2, ii, k, dir, ie] .= fr.t
                sub_frames[:, 3, ii, k, dir, ie] .= fr.b
                sub_jac[ii, k, dir, ie] = sg.jac[dir][is]
            end
        end
    end
    # boundary conditions in mesh.bdfaces order (disc.bcs is already ordered by bdmap)
    kinds = Int32[bckind(bc) for bc in bcs]
    offsets = Int64[0; cumsum(length.(mesh.bdfaces))]
    bcfaces = Int64.(reduce(vcat, mesh.bdfaces; init=Int[]))
    nfp = Flou.FlouSpatial.ndofs(std.face)
    state = zeros(Float64, nv, max(length(bcs), 1))          # row per boundary (C order)
    table = zeros(Float64, nv, max(length(bcfaces), 1) * nfp)
    for (ib, bc) in enumerate(bcs)
        if bc isa EulerInflowBC
            state[:, ib] .= bc.Qext
        elseif bc isa GenericBC
            for m in (offsets[ib] + 1):offsets[ib + 1], i in 1:nfp
                f = bcfaces[m]
                x = geometry.faces[f].coords[i]
                # only position-dependent closures are supported on the device
                table[:, (m - 1) * nfp + i] .= bc.Qext(nothing, x, nothing, 0.0, equation)
            end
        end
    end
    a = equation isa LinearAdvection ? ntuple(i -> i <= ND ? Float64(equation.a[i]) : 0.0, 3) : (0.0, 0.0, 0.0)
    dx = cart ? ntuple(i -> i <= ND ? Float64(mesh.Δx[i]) : 0.0, 3) : (0.0, 0.0, 0.0)
    ne = nelements(mesh)
    if op isa HybridDivOperator && op.fvflux !== op.numflux
        throw(ArgumentError("flou_b200: HybridDivOperator needs fvflux === numflux (both convenience constructors)"))
    end
    handle = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve faceinds facepos eleminds elempos orientation D Ds Dsharp lm lp dgm dgp w1d jac metric fjac frames sub_frames sub_jac kinds offsets bcfaces state table begin
        desc = Desc(
            Int32(sizeof(Desc)), Int32(ND), Int32(nv), Int32(size(D, 1)),
            equation isa EulerEquation ? Int32(1) : Int32(0),
            op isa HybridDivOperator ? Int32(2) : (op isa SplitDivOperator ? Int32(1) : Int32(0)),
            op isa Union{SplitDivOperator,HybridDivOperator} ? fluxkind(op.tpflux) : Int32(0),
            fluxkind(op.numflux), avgkind(op.numflux), cart ? Int32(0) : Int32(1),
            intensity(op.numflux), equation isa EulerEquation ? Float64(equation.γ) : 0.0, a, dx,
            Int64(ne), Int64(nfaces(mesh)),
            pointer(faceinds), pointer(facepos), pointer(eleminds), pointer(elempos), pointer(orientation),
            pointer(D), pointer(Ds), pointer(Dsharp), pointer(lm), pointer(lp), pointer(dgm), pointer(dgp),
            pointer(w1d),
            cart ? C_NULL : pointer(jac), cart ? C_NULL : pointer(metric),
            cart ? C_NULL : pointer(fjac), cart ? C_NULL : pointer(frames),
            Int32(length(bcs)), pointer(kinds), pointer(offsets), pointer(bcfaces),
            pointer(state), pointer(table),
            Int64(0), Int64(ne), Int32(0), Int32(1), C_NULL, Int32(device), Int32(0),
            op isa HybridDivOperator ? Float64(op.blend) : 0.0,
            needs_sub ? pointer(sub_frames) : C_NULL, needs_sub ? pointer(sub_jac) : C_NULL)
        check(ccall((:flou_b200_create, lib), Int32, (Ref{Desc}, Ref{Ptr{Cvoid}}), desc, handle))
    end
    b = B200Disc{ND,RT,typeof(disc)}(disc, handle[])
    finalizer(x -> ccall((:flou_b200_destroy, lib), Int32, (Ptr{Cvoid},), x.handle), b)
    return b
end

# rhs!(dQ, Q, p, t): host in, host out -- what OrdinaryDiffEq calls as f(du, u, p, t)
function rhs!(dQ::Matrix{Float64}, Q::Matrix{Float64}, p::EquationConfig{<:B200Disc}, time::Real)
    check(ccall((:flou_b200_rhs, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Float64),
                p.disc.handle, Q, dQ, Float64(time)))
    return nothing
end

# get_max_dt(q, disc, equation, cfl): used by get_cfl_callback (FlouTime.jl:92-105)
function Flou.FlouCommon.get_max_dt(q::Matrix{Float64}, disc::B200Disc, ::Any, cfl)
    dt = Ref{Float64}(0.0)
    check(ccall((:flou_b200_max_dt, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}, Float64, Ref{Float64}),
                disc.handle, q, Float64(cfl), dt))
    return dt[]
end

# get_monitor(disc, equation, name) (FlouSpatial/Equations/Euler.jl:541-557): the returned closure
# has the reference's signature (_Q, disc, equation); `_Q === nothing` reads the device state.
# (The reference lists :kinetic_energy but dispatches on :energy: both are accepted.)
function Flou.FlouCommon.get_monitor(disc::B200Disc, ::EulerEquation, name::Symbol, _=nothing)
    kind = name in (:kinetic_energy, :energy) ? Int32(0) :
           name == :entropy ? Int32(1) : error("Unknown monitor '$(name)'.")
    return (_Q, d, _eq) -> begin
        v = Ref{Float64}(0.0)
        check(ccall((:flou_b200_monitor, lib), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ref{Float64}),
                    d.handle, kind, _Q === nothing ? Ptr{Float64}(C_NULL) : pointer(_Q), v))
        v[]
    end
end

# get_limiter(disc, equation, :zhang_shu, minval) (Euler.jl:597-660): limits _Q in place
function Flou.FlouCommon.get_limiter(disc::B200Disc, ::EulerEquation, name::Symbol, minval=nothing)
    name == :zhang_shu || error("Unknown limiter '$(name)'.")
    minval === nothing && throw(ArgumentError(
        "The minimum value must be specified when using the limiter of Zhang & Shu."))
    return (_Q, d, _eq) -> begin
        check(ccall((:flou_b200_zhang_shu, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}, Float64),
                    d.handle, _Q === nothing ? Ptr{Float64}(C_NULL) : pointer(_Q), Float64(minval)))
        nothing
    end
end

# ORK256(stage_limiter! = get_limiter_callback(dg, eq, :zhang_shu, minval)) as in
# examples/src/3D_Euler.jl:76-80: the fused RK loop applies the limiter after every stage
struct B200StageLimiter
    minval::Float64
end
stage_limiter(disc::B200Disc, equation, name::Symbol, minval) =
    (name == :zhang_shu || error("Unknown limiter '$(name)'."); B200StageLimiter(Float64(minval)))
function set_stage_limiter!(disc::B200Disc, lim::Union{B200StageLimiter,Nothing})
    check(ccall((:flou_b200_set_stage_limiter, lib), Int32, (Ptr{Cvoid}, Int32, Float64),
                disc.handle, lim === nothing ? Int32(0) : Int32(1), lim === nothing ? 0.0 : lim.minval))
end

# 2N tableaus of OrdinaryDiffEq's LowStorageRK2N solvers (A_1 = 0; tmp = A_s tmp + dt k, u += B_s tmp),
# written out here so that the shim does not depend on the signature of the internal `alg_cache`:
# ORK256 (Bernardini & Pirozzoli 2009, 5 stages, as tabulated by OrdinaryDiffEq v6.49.1) and
# Carpenter & Kennedy (1994) 2N(5,4).  The same numbers as flou_b200/time.py, pinned by
# tests/test_oracle_kat.py (Sod / Shockwave2D KATs, order conditions).
tableau(::ORK256) = (
    Float64[0.0, -1.0, -1.55798, -1.0, -0.45031],
    Float64[0.2, 0.83204, 0.6, 0.35394, 0.2],
    Float64[0.0, 0.2, 0.2, 0.8, 0.8])
tableau(::CarpenterKennedy2N54) = (
    Float64[0.0, -567301805773 / 1357537059087, -2404267990393 / 2016746695238,
            -3550918686646 / 2091501179385, -1275806237668 / 842570457699],
    Float64[1432997174477 / 9575080441755, 5161836677717 / 13612068292357,
            1720146321549 / 2090206949498, 3134564353537 / 4481467310338,
            2277821191437 / 14882151754819],
    Float64[0.0, 1432997174477 / 9575080441755, 2526269341429 / 6820363962896,
            2006345519317 / 3224310063776, 2802321613138 / 2924317926251])

# timeintegrate(Q0, disc, equation, solver, tfinal; adaptive=false, dt, alias_u0=true)
function timeintegrate(Q0::Matrix{Float64}, disc::B200Disc, equation,
                       solver::Union{ORK256,CarpenterKennedy2N54}, tfinal;
                       dt, adaptive=false, alias_u0=true,
                       stage_limiter::Union{B200StageLimiter,Nothing}=nothing, kwargs...)
    adaptive && throw(ArgumentError("adaptive=true is not supported on the B200 path"))
    set_stage_limiter!(disc, stage_limiter)
    A, B, c = tableau(solver)
    Q = alias_u0 ? Q0 : copy(Q0)
    # adaptive=false: full steps of dt, the last one shortened so that it lands on tfinal (tstops)
    r = tfinal / dt
    nsteps = round(Int64, r)
    last = 0.0
    if abs(r - nsteps) > 1e-9 * max(1.0, abs(r))
        nsteps = floor(Int64, r)
        last = tfinal - nsteps * dt
    end
    advance(h, n) = ccall((:flou_b200_lsrk2n_advance, lib), Int32,
                          (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                           Float64, Float64, Int64),
                          disc.handle, Int32(length(B)), A, B, c, Float64(h), 0.0, Int64(n))
    rc = Int32(0)
    exetime = @elapsed begin
        if last == 0.0
            rc = ccall((:flou_b200_timeintegrate, lib), Int32,
                       (Ptr{Cvoid}, Ptr{Float64}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                        Float64, Float64, Int64),
                       disc.handle, Q, Int32(length(B)), A, B, c, Float64(dt), 0.0, nsteps)
        else
            rc = ccall((:flou_b200_upload_state, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}), disc.handle, Q)
            rc == 0 && (rc = advance(dt, nsteps))
            rc == 0 && (rc = advance(last, 1))
            rc == 0 && (rc = ccall((:flou_b200_download_state, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}), disc.handle, Q))
            flags = Ref{Int32}(0)
            rc == 0 && (rc = ccall((:flou_b200_status, lib), Int32, (Ptr{Cvoid}, Ref{Int32}), disc.handle, flags))
            rc == 0 && (flags[] & 1) != 0 && (rc = Int32(4))
        end
    end
    if rc == 4                       # FLOU_B200_EDOMAIN, cf. FlouTime.jl:39-51
        @error "Simulation crashed!"
        return (nothing, exetime)
    end
    check(rc)
    return ((u=[Q], t=[tfinal]), exetime)
end

# ---- source term and boundary data that change between stages ---------------------------------
# apply_sourceterm! (MultielementDiscontinuous.jl:139-146) and GenericBC closures that read Qin,
# frame or time (Interfaces.jl:44-48) are host closures: the shim tabulates them and the device reads
# the tables.  `S`: (ndofs, nv) increments of dQ after the mass matrix; `nothing` removes the source.
function set_source!(disc::B200Disc, S::Union{Matrix{Float64},Nothing})
    check(ccall((:flou_b200_set_source, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}),
                disc.handle, S === nothing ? Ptr{Float64}(C_NULL) : pointer(S)))
end
# interior traces at the nodes of the owned boundary faces: (nfp, nv, count) and the ordinal of
# every face in the concatenated `bdfaces` (row block ordinal*nfp .. of the BC table)
function boundary_traces(disc::B200Disc, nfp::Integer, nv::Integer)
    n = Ref{Int64}(0)
    check(ccall((:flou_b200_boundary_traces, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Int64}, Ref{Int64}),
                disc.handle, C_NULL, C_NULL, n))
    Qin = Array{Float64}(undef, nfp, nv, n[]); ord = Vector{Int64}(undef, n[])
    check(ccall((:flou_b200_boundary_traces, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Int64}, Ref{Int64}),
                disc.handle, Qin, ord, n))
    return Qin, ord
end
set_bc_table!(disc::B200Disc, table::Matrix{Float64}) =       # (nv, rows): row-major rows of nv values
    check(ccall((:flou_b200_set_bc_table, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}), disc.handle, table))
lsrk2n_stage!(disc::B200Disc, A, B, dt, first::Bool) =
    check(ccall((:flou_b200_lsrk2n_stage, lib), Int32, (Ptr{Cvoid}, Float64, Float64, Float64, Int32),
                disc.handle, Float64(A), Float64(B), Float64(dt), Int32(first)))

"""
    pointdata2VTKHDF(Q, disc, node2eq1d, nv)

Device version of `FlouBiz.pointdata2VTKHDF(Q, disc::MultielementDisc)` (src/FlouSpatial/IO.jl:78-97):
the state at the equispaced nodes of every element, one vector per variable, ready for
`write(file.handler, "VTKHDF/PointData/name", data)`.  `node2eq1d` is the `node2eq` of the SEGMENT
the element region is built from (`std.face` of a quad, `std.face.face` of a hex: StdQuad.jl:43,
StdHex.jl:44-45): the 1-D
interpolation matrix whose Kronecker products the quads and hexes use.  `Q === nothing` projects
the device-resident state (what a save callback wants between steps).
"""
function pointdata2VTKHDF(Q::Union{Nothing,Matrix{Float64}}, disc::B200Disc{ND}, node2eq1d::Matrix{Float64},
                          nv::Integer) where {ND}
    neq, np = size(node2eq1d)
    M = permutedims(node2eq1d)                       # the library reads [neq][np] row-major
    out = Matrix{Float64}(undef, nelements(disc.disc.mesh) * neq^ND, nv)
    check(ccall((:flou_b200_project_equispaced, lib), Int32,
                (Ptr{Cvoid}, Ptr{Float64}, Int32, Ptr{Float64}, Ptr{Float64}),
                disc.handle, Q === nothing ? C_NULL : Q, Int32(neq), M, out))
    return [out[:, v] for v in 1:nv]
end

"""
    timeintegrate_stagewise(Q0, disc, solver, tf, refresh!; dt)

The RK loop for discretisations whose source term or `GenericBC` closures depend on the state or
the time: `refresh!(disc, t)` re-tabulates them (with `set_source!`, `boundary_traces`,
`set_bc_table!`) before every stage at `t + c_s dt`, as OrdinaryDiffEq evaluates `f(u, p, t + c_s dt)`.
"""
function timeintegrate_stagewise(Q0::Matrix{Float64}, disc::B200Disc,
                                 solver::Union{ORK256,CarpenterKennedy2N54}, tf, refresh!; dt)
    A, B, c = tableau(solver)
    check(ccall((:flou_b200_upload_state, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}), disc.handle, Q0))
    t = 0.0
    while t < tf - 1e-14 * max(1.0, abs(tf))
        h = min(dt, tf - t)
        for s in eachindex(B)
            refresh!(disc, t + c[s] * h)
            lsrk2n_stage!(disc, A[s], B[s], h, s == 1)
        end
        t += h
    end
    check(ccall((:flou_b200_download_state, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}), disc.handle, Q0))
    return Q0
end

end # module
