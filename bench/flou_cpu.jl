# Plan A of the CPU baseline (BASELINE.md section 4): Flou.jl's OWN multithreaded path, timed on
# the host cores of the GPU box, and the Q / dQ dump for the 1e-12 parity check against it.
#
#   julia -t <threads> --project=<Flou.jl checkout> bench/flou_cpu.jl <cfg> <n> <np> <steps> [dumpdir]
#
#   cfg     cfg1 | cfg2 | cfg3 | cfg4        (SURVEY.md 8(d): operator, fluxes and IC of that config)
#   n       elements per direction of the (bounded) sample mesh
#   np      nodes per direction (p + 1)
#   steps   timed ORK256 steps (after one untimed step that pays for the JIT)
#   dumpdir optional: writes Q.bin / dQ.bin (Float64, column-major (ndofs, nv)) of the initial
#           state and `rhs!` of it, for tests/test_julia_dump.py
#
# Prints ONE JSON line: {"rate": DOF-updates/s per RK stage, "rhs_rate": ..., "threads": ...,
# "ndofs": ..., "ms_per_step": ...}.  bench.py probes `julia` (PATH, then baseline/_ref/julia/bin)
# and runs this script when it finds one; otherwise the C/OpenMP restatement (oracle/) is timed
# and labelled `"kind": "port"`.
#
# NOT executed in the build container (no Julia there).  Uses only the reference's public API as
# its own examples do: examples/src/3D_Euler.jl:29-91, examples/src/2D_LinearAdvection.jl:25-72,
# test/tests.jl:136-185; timed functions: rhs! (src/FlouSpatial/Equations/Hyperbolic.jl:31-69) and
# timeintegrate (src/FlouTime/FlouTime.jl:34-54).
using Flou
using OrdinaryDiffEq
using LinearAlgebra: BLAS
using Printf

Threads.nthreads() > 1 && BLAS.set_num_threads(1)       # examples/src/3D_Euler.jl:29-31

cfg = length(ARGS) >= 1 ? ARGS[1] : "cfg3"
n = length(ARGS) >= 2 ? parse(Int, ARGS[2]) : 16
np = length(ARGS) >= 3 ? parse(Int, ARGS[3]) : 4
steps = length(ARGS) >= 4 ? parse(Int, ARGS[4]) : 3
dumpdir = length(ARGS) >= 5 ? ARGS[5] : ""

const γ = 1.4
function setup(cfg, n, np)
    if cfg == "cfg1"
        eq = LinearAdvection(2.0, -1.0)
        mesh = CartesianMesh{2,Float64}((0, 0), (1, 1), (n, n))
        apply_periodicBCs!(mesh, "1" => "2", "3" => "4")
        op = StrongDivOperator(LxF(StdAverage(), 1.0))
        std = StdQuad(LagrangeBasis(:GLL, np), LagrangeBasis(:GLL, np) |> DGSEMrec, nvariables(eq))
        ic = x -> (Flou.gaussian_bump(x[1], x[2], 0.5, 0.5, 0.1, 0.1, 1.0),)
        return eq, mesh, op, std, ic, 1e-3
    end
    op = SplitDivOperator(MatrixDissipation(ChandrasekharAverage(), 1.0))
    if cfg == "cfg2"
        eq = EulerEquation{2}(γ)
        mesh = CartesianMesh{2,Float64}((0, 0), (10, 10), (n, n))
        apply_periodicBCs!(mesh, "1" => "2", "3" => "4")
        basis = LagrangeBasis(:GLL, np)
        std = StdQuad(basis, basis |> DGSEMrec, nvariables(eq))
        β = 5.0
        ic = x -> begin                                  # isentropic vortex, SURVEY.md 8(d)
            xr, yr = x[1] - 5, x[2] - 5
            r2 = xr^2 + yr^2
            e = exp(0.5 * (1 - r2))
            u = 1 - β / 2π * yr * e
            v = 1 + β / 2π * xr * e
            T = 1 - (γ - 1) * β^2 / (8γ * π^2) * exp(1 - r2)
            ρ = T^(1 / (γ - 1))
            Flou.vars_prim2cons((ρ, u, v, ρ * T), eq)
        end
        return eq, mesh, op, std, ic, 1e-3
    end
    eq = EulerEquation{3}(γ)                             # cfg3 / cfg4: Taylor-Green vortex
    mesh = CartesianMesh{3,Float64}((0, 0, 0), (2π, 2π, 2π), (n, n, n))
    apply_periodicBCs!(mesh, "1" => "2", "3" => "4", "5" => "6")
    basis = LagrangeBasis(:GLL, np)
    std = StdHex(basis, basis |> DGSEMrec, nvariables(eq))
    M0 = 0.1
    ic = x -> begin
        u = sin(x[1]) * cos(x[2]) * cos(x[3])
        v = -cos(x[1]) * sin(x[2]) * cos(x[3])
        p = 1 / (γ * M0^2) + (cos(2x[1]) + cos(2x[2])) * (cos(2x[3]) + 2) / 16
        Flou.vars_prim2cons((1.0, u, v, 0.0, p), eq)
    end
    return eq, mesh, op, std, ic, cfg == "cfg4" ? 1e-4 : 1e-3
end

eq, mesh, op, std, ic, Δt = setup(cfg, n, np)
dg = MultielementDisc(mesh, std, eq, op, ())
Q = GlobalStateVector{nvariables(eq)}(undef, dg.dofhandler)
for i in eachdof(dg)
    q = ic(dg.geometry.elements.coords[i])
    for v in 1:nvariables(eq)
        Q.data[i, v] = q[v]                              # Q.data::Matrix (ndofs, nv), GlobalContainers.jl:19-29
    end
end
dQ = similar(Q.data)
p = Flou.EquationConfig(dg, eq)

if !isempty(dumpdir)
    mkpath(dumpdir)
    Flou.FlouCommon.rhs!(dQ, Q.data, p, 0.0)
    write(joinpath(dumpdir, "Q.bin"), Q.data)
    write(joinpath(dumpdir, "dQ.bin"), dQ)
    write(joinpath(dumpdir, "shape.txt"), "$(size(Q.data, 1)) $(size(Q.data, 2))\n")
end

# (i) rhs! alone
Flou.FlouCommon.rhs!(dQ, Q.data, p, 0.0)                            # JIT
nrhs = max(steps, 1) * 5
t_rhs = @elapsed for _ in 1:nrhs
    Flou.FlouCommon.rhs!(dQ, Q.data, p, 0.0)
end

# (ii) the RK loop as the reference runs it
solver = ORK256(williamson_condition=false)
timeintegrate(Q.data, dg, eq, solver, Δt; save_everystep=false, alias_u0=true, adaptive=false, dt=Δt)   # JIT
_, exetime = timeintegrate(Q.data, dg, eq, solver, steps * Δt;
                           save_everystep=false, alias_u0=true, adaptive=false, dt=Δt)

nd = ndofs(dg)
@printf("{\"rate\": %.6e, \"rhs_rate\": %.6e, \"threads\": %d, \"ndofs\": %d, \"ms_per_step\": %.6f, \"steps\": %d}\n",
        nd * 5 * steps / exetime, nd * nrhs / t_rhs, Threads.nthreads(), nd, exetime / steps * 1e3, steps)
