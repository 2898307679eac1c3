# QPB200.jl -- thin Julia shim over libqpb200.so (include/qpb200.h).
#
# Drop-in for the hot path of the reference (RoyiAvital/QuadraticProgramSolver):
#
#     SolveQuadraticProgram!(vX, mP, vQ, mA, vL, vU, LinSysSolInit, LinSysSol!; kw...)      SolveQuadraticProgram.jl:14-17
#
# A call site changes only the two plugin handles:
#
#     include("SolveQuadraticProgram.jl"); include("LinearSystemSolvers.jl");   # the reference, unchanged
#     include("QPB200.jl");                                                        # adds one method + the plugin pair
#     convFlag = SolveQuadraticProgram!(vX, mP, vQ, mA, vL, vU, B200Init, B200Sol!; numIterations = 50000, ρ = 0.1, adptΡ = true);
#
# Everything below the `ccall`s runs on the GPU; there is no CUDA.jl array path and no CPU fallback
# (on a machine without a B200 the create call returns QPB200_ERR_DEVICE and this shim throws).
#
# NOTE: Julia is not installed in the build/test environment of this repository, so this file is not
# executed by the test-suite; quadraticprogramsolver_b200/solver.py is the tested mirror and calls the
# same C symbols in the same order with the same struct layouts (tests/test_host.py checks the layouts).

# Like every file of the reference this one is `include`d into Main (the reference has no module
# structure: RunTests.jl:15-18); it adds a METHOD to `SolveQuadraticProgram!` that dispatches on the
# plugin singletons, so the reference's own methods and plugins keep working next to it.

using SparseArrays

const libqpb200 = get(ENV, "QPB200_LIB", joinpath(@__DIR__, "..", "quadraticprogramsolver_b200", "libqpb200.so"))

# same numbering as the reference enum (SolveQuadraticProgram.jl:12); defined here only when the
# reference file has not been included
if !@isdefined(ConvergenceFlag)
    @enum ConvergenceFlag convNumItr = 1 convAdmm convPrimDual
end

# the plugin pair that selects the B200 path (cf. FacLdlInit / FacLdl!, LinearSystemSolvers.jl:78,91)
struct B200InitT end
struct B200SolT end
const B200Init = B200InitT()
const B200Sol! = B200SolT()

# qpb200_settings (include/qpb200.h) -- field order and types must match exactly (200 bytes)
mutable struct Settings
    max_iter::Int64
    eps_abs::Float64
    eps_rel::Float64
    rho::Float64
    sigma::Float64
    alpha::Float64
    delta::Float64
    adaptive_rho::Int32
    lin_solver::Int32
    rho_factor::Float64
    check_every::Int64
    polish_iter::Int64
    minres_eps::Float64
    minres_iter::Int64
    pcg_eps::Float64
    pcg_max_iter::Int64
    pcg_rel_eps::Float64
    precond::Int32
    device::Int32
    spmv_loader::Int32
    reserved_i::NTuple{7, Int32}
    reserved_d::NTuple{4, Float64}
    Settings() = new()
end

# qpb200_info (104 bytes)
mutable struct Info
    conv_flag::Int32
    polish_status::Int32
    iterations::Int64
    rho_final::Float64
    res_prim::Float64
    res_dual::Float64
    rho_updates::Int64
    pcg_iters_total::Int64
    pcg_maxed::Int64
    solve_ms::Float64
    setup_ms::Float64
    kernel_launches::Int64
    polish_minres_iters::Int64
    polish_active::Int64
    Info() = new()
end

function check(rc::Cint)
    rc == 0 && return
    msg = unsafe_string(ccall((:qpb200_last_error, libqpb200), Cstring, ()))
    error("libqpb200 error $(rc): $(msg)")
end

"""
    SolveQuadraticProgram!(vX, mP, vQ, mA, vL, vU, ::B200InitT, ::B200SolT; kw...) -> ConvergenceFlag

Same positional arguments, keyword names, defaults and return value as the reference method
(SolveQuadraticProgram.jl:14-17); `vX` is the start point and is overwritten with the solution.
New keywords: `ϵPcg`, `numItrPcg` (the plugin kwargs of LinearSystemSolvers.jl:125, which the reference
driver never forwards), `precond ∈ (:jacobi, :none)`, `device`, `numItrScaling` (Ruiz equilibration, default off),
`rhoScale` (per-constraint step size ρᵢ = ρ·rhoScale[i], default `nothing` = the scalar ρ; `EqualityRhoScale(vL, vU)`
builds OSQP's choice).
"""
function SolveQuadraticProgram!(vX::Vector{Float64}, mP::SparseMatrixCSC{Float64, Int64}, vQ::Vector{Float64},
        mA::SparseMatrixCSC{Float64, Int64}, vL::Vector{Float64}, vU::Vector{Float64}, ::B200InitT, ::B200SolT;
        numIterations = 5000, ϵAbs = 1e-6, ϵRel = 1e-6, ρ = 1, σ = 1e-6, α = 1.6, δ = 1e-6, adptΡ::Bool = false,
        fctrΡ = 5, numItrConv = 25, numItrPolish = 10, ϵMinres = 1e-6, numItrMinres = 500,
        ϵPcg = 1e-6, numItrPcg = 1000, precond::Symbol = :jacobi, device::Integer = -1, numItrScaling::Integer = 0,
        rhoScale::Union{Vector{Float64}, Nothing} = nothing, info::Union{Info, Nothing} = nothing)

    numElements, numConstraints = length(vX), size(mA, 1);
    (size(mP) == (numElements, numElements) && size(mA, 2) == numElements && length(vQ) == numElements &&
        length(vL) == numConstraints && length(vU) == numConstraints) || throw(DimensionMismatch("QP dimensions"));

    s = Settings();
    ccall((:qpb200_default_settings, libqpb200), Cvoid, (Ref{Settings},), s);
    s.max_iter = numIterations; s.eps_abs = ϵAbs; s.eps_rel = ϵRel; s.rho = ρ; s.sigma = σ; s.alpha = α; s.delta = δ;
    s.adaptive_rho = adptΡ; s.rho_factor = fctrΡ; s.check_every = numItrConv; s.polish_iter = numItrPolish;
    s.minres_eps = ϵMinres; s.minres_iter = numItrMinres; s.pcg_eps = ϵPcg; s.pcg_max_iter = numItrPcg;
    s.precond = precond == :none ? 0 : 1; s.device = device;
    # reserved_i[QPB200_RSV_SCALING_ITERS = 2] (0-based): Ruiz equilibration iterations, 0 = off (README.md:71 TODO)
    s.reserved_i = ntuple(k -> k == 3 ? Int32(numItrScaling) : s.reserved_i[k], 7);

    hRef = Ref{Ptr{Cvoid}}(C_NULL);
    # SparseMatrixCSC fields are passed as they are: colptr/rowval are 1-based Int64 -> index_base = 1, zero copies
    GC.@preserve mP mA vQ vL vU begin
        check(ccall((:qpb200_create, libqpb200), Cint,
            (Ref{Ptr{Cvoid}}, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64},
             Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Settings}, Int32),
            hRef, numElements, numConstraints, mP.colptr, mP.rowval, mP.nzval, mA.colptr, mA.rowval, mA.nzval,
            vQ, vL, vU, s, Int32(1)));
    end
    sInfo = info === nothing ? Info() : info;
    try
        if rhoScale !== nothing
            length(rhoScale) == numConstraints || throw(DimensionMismatch("rhoScale"));
            check(ccall((:qpb200_set_rho_scale, libqpb200), Cint, (Ptr{Cvoid}, Ptr{Float64}), hRef[], rhoScale));
        end
        check(ccall((:qpb200_solve, libqpb200), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Info}),
            hRef[], vX, C_NULL, C_NULL, sInfo));
    finally
        ccall((:qpb200_destroy, libqpb200), Cvoid, (Ptr{Cvoid},), hRef[]);
    end
    return ConvergenceFlag(sInfo.conv_flag);
end

"""
    EqualityRhoScale(vL, vU, factor = 1e3)

OSQP's rho vector as a scale on the scalar ρ: `factor` on the equality rows (`l == u`), 1 elsewhere.
"""
EqualityRhoScale(vL, vU, factor = 1e3) = [l == u ? Float64(factor) : 1.0 for (l, u) in zip(vL, vU)];

"""
    SolveQuadraticProgram(P, q, A, l, u; kw...) -> (x, convFlag, info)

The convenience form named in BASELINE.json (start point zero).
"""
function SolveQuadraticProgram(mP, vQ, mA, vL, vU; kw...)
    vX = zeros(length(vQ));
    sInfo = Info();
    convFlag = SolveQuadraticProgram!(vX, SparseMatrixCSC{Float64, Int64}(mP), Vector{Float64}(vQ),
        SparseMatrixCSC{Float64, Int64}(mA), Vector{Float64}(vL), Vector{Float64}(vU), B200Init, B200Sol!; info = sInfo, kw...);
    return vX, convFlag, sInfo;
end

