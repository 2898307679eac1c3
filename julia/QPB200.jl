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


# qpb200_proxqp_report (72 bytes)
mutable struct ProxReport
    converged::Int32
    reserved::Int32
    iterations::Int64
    rho::Float64
    sigma::Float64
    res_prim::Float64
    res_dual::Float64
    rho_updates::Int64
    solve_ms::Float64
    kernel_launches::Int64
    ProxReport() = new()
end

# settings.reserved_i slots (0-based in the header, 1-based here): QPB200_RSV_*
const RSV_CHOL_UNBLOCKED = 1; const RSV_DIST_MODE = 2; const RSV_SCALING_ITERS = 3; const RSV_DENSE_VARIANT = 4;
const RSV_CG_RECURRENCE = 5; const RSV_POLISH = 6; const RSV_BATCH_CHUNK = 7;

function _SetReserved!(s::Settings, slot::Integer, val::Integer)
    s.reserved_i = ntuple(k -> k == slot ? Int32(val) : s.reserved_i[k], 7);
end

"""
    B200Settings(; kw...) -> Settings

Keyword arguments of `SolveQuadraticProgram!` (SolveQuadraticProgram.jl:15-17, same names and defaults) -> `qpb200_settings`.
New keywords (all default to the reference's behaviour): `ϵPcg`, `numItrPcg` (the plugin kwargs of
LinearSystemSolvers.jl:125, which the reference driver never forwards), `linSolver ∈ (:pcg, :cholesky)` (`:cholesky` =
exact solve like `FacLdlInit, FacLdl!`), `precond ∈ (:jacobi, :none)`, `device`, `numItrScaling` (Ruiz equilibration),
`polish` (run the polish the reference reserves `numItrPolish, δ, ϵMinres, numItrMinres` for), `distMode`.
"""
function B200Settings(; numIterations = 5000, ϵAbs = 1e-6, ϵRel = 1e-6, ρ = 1, σ = 1e-6, α = 1.6, δ = 1e-6, adptΡ::Bool = false,
        fctrΡ = 5, numItrConv = 25, numItrPolish = 10, ϵMinres = 1e-6, numItrMinres = 500,
        ϵPcg = 1e-6, numItrPcg = 1000, linSolver::Symbol = :pcg, precond::Symbol = :jacobi, device::Integer = -1,
        numItrScaling::Integer = 0, polish::Bool = false, distMode::Symbol = :auto, batchChunk::Integer = 0)
    s = Settings();
    ccall((:qpb200_default_settings, libqpb200), Cvoid, (Ref{Settings},), s);
    s.max_iter = numIterations; s.eps_abs = ϵAbs; s.eps_rel = ϵRel; s.rho = ρ; s.sigma = σ; s.alpha = α; s.delta = δ;
    s.adaptive_rho = adptΡ; s.rho_factor = fctrΡ; s.check_every = numItrConv; s.polish_iter = numItrPolish;
    s.minres_eps = ϵMinres; s.minres_iter = numItrMinres; s.pcg_eps = ϵPcg; s.pcg_max_iter = numItrPcg;
    s.lin_solver = linSolver == :cholesky ? 1 : 0;
    s.precond = precond == :none ? 0 : 1; s.device = device;
    _SetReserved!(s, RSV_SCALING_ITERS, numItrScaling);        # README.md:71 TODO of the reference; 0 = off
    _SetReserved!(s, RSV_POLISH, polish ? 1 : 0);
    _SetReserved!(s, RSV_DIST_MODE, distMode == :nccl ? 1 : (distMode == :peer ? 2 : 0));
    _SetReserved!(s, RSV_BATCH_CHUNK, batchChunk);
    return s;
end

"""
    SolveQuadraticProgram!(vX, mP, vQ, mA, vL, vU, ::B200InitT, ::B200SolT; kw...) -> ConvergenceFlag

Same positional arguments, keyword names, defaults and return value as the reference method
(SolveQuadraticProgram.jl:14-17); `vX` is the start point and is overwritten with the solution.
Keywords: see `B200Settings`; plus `rhoScale` (per-constraint step size ρᵢ = ρ·rhoScale[i], default `nothing` = the
scalar ρ; `EqualityRhoScale(vL, vU)` builds OSQP's choice) and `info` (an `Info()` to fill).
"""
function SolveQuadraticProgram!(vX::Vector{Float64}, mP::SparseMatrixCSC{Float64, Int64}, vQ::Vector{Float64},
        mA::SparseMatrixCSC{Float64, Int64}, vL::Vector{Float64}, vU::Vector{Float64}, ::B200InitT, ::B200SolT;
        rhoScale::Union{Vector{Float64}, Nothing} = nothing, info::Union{Info, Nothing} = nothing, kw...)

    numElements, numConstraints = length(vX), size(mA, 1);
    (size(mP) == (numElements, numElements) && size(mA, 2) == numElements && length(vQ) == numElements &&
        length(vL) == numConstraints && length(vU) == numConstraints) || throw(DimensionMismatch("QP dimensions"));
    s = B200Settings(; kw...);
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

The convenience form named in BASELINE.json (start point zero).  `numGpus = R > 1` is not a keyword of a single
process: one Julia process drives one GPU, see `SolveQuadraticProgramDist!`.
"""
function SolveQuadraticProgram(mP, vQ, mA, vL, vU; kw...)
    vX = zeros(length(vQ));
    sInfo = Info();
    convFlag = SolveQuadraticProgram!(vX, SparseMatrixCSC{Float64, Int64}(mP), Vector{Float64}(vQ),
        SparseMatrixCSC{Float64, Int64}(mA), Vector{Float64}(vL), Vector{Float64}(vU), B200Init, B200Sol!; info = sInfo, kw...);
    return vX, convFlag, sInfo;
end

# ---------------------------------------------------------------------------------------------------------------
# One large sparse QP over R GPUs (qpb200_dist_*): one Julia process (or task pinned to a device) per GPU, e.g. under
# MPI.jl / Distributed.jl.  Rank 0 calls DistUniqueId() and ships the 128 bytes to the others by whatever transport
# the launcher has; every rank then calls SolveQuadraticProgramDist! with the WHOLE problem -- the library partitions
# it (rows of A / columns of P per rank) and combines the partial sums in-kernel over NVLink peer memory.
# ---------------------------------------------------------------------------------------------------------------
function DistUniqueId()
    vId = zeros(UInt8, 128);
    check(ccall((:qpb200_dist_unique_id, libqpb200), Cint, (Ptr{UInt8},), vId));
    return vId;
end

"""
    SolveQuadraticProgramDist!(vX, mP, vQ, mA, vL, vU, rank, numGpus, vId; kw...) -> (ConvergenceFlag, rowRange, vZ, vY)

Collective over the `numGpus` ranks (`rank` is 0-based).  `vX` is replicated; `vZ, vY` are this rank's rows `rowRange`.
"""
function SolveQuadraticProgramDist!(vX::Vector{Float64}, mP::SparseMatrixCSC{Float64, Int64}, vQ::Vector{Float64},
        mA::SparseMatrixCSC{Float64, Int64}, vL::Vector{Float64}, vU::Vector{Float64}, rank::Integer, numGpus::Integer,
        vId::Vector{UInt8}; info::Union{Info, Nothing} = nothing, kw...)
    numElements, numConstraints = length(vX), size(mA, 1);
    s = B200Settings(; kw...);
    hRef = Ref{Ptr{Cvoid}}(C_NULL);
    GC.@preserve mP mA vQ vL vU vId begin
        check(ccall((:qpb200_dist_create_full, libqpb200), Cint,
            (Ref{Ptr{Cvoid}}, Int32, Int32, Ptr{UInt8}, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int64},
             Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Settings}, Int32),
            hRef, Int32(rank), Int32(numGpus), vId, numElements, numConstraints, mP.colptr, mP.rowval, mP.nzval,
            mA.colptr, mA.rowval, mA.nzval, vQ, vL, vU, s, Int32(1)));
    end
    sInfo = info === nothing ? Info() : info;
    r0 = Ref{Int64}(0); r1 = Ref{Int64}(0);
    vZ = Float64[]; vY = Float64[];
    try
        check(ccall((:qpb200_dist_rows, libqpb200), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}), hRef[], r0, r1));
        vZ = zeros(r1[] - r0[]); vY = zeros(r1[] - r0[]);
        check(ccall((:qpb200_dist_solve, libqpb200), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Info}),
            hRef[], vX, vZ, vY, sInfo));
    finally
        ccall((:qpb200_destroy, libqpb200), Cvoid, (Ptr{Cvoid},), hRef[]);
    end
    return ConvergenceFlag(sInfo.conv_flag), (r0[] + 1):r1[], vZ, vY;
end

# ---------------------------------------------------------------------------------------------------------------
# Batches of small dense QPs (qpb200_batch_*): tP is n x n x batch, tA is m x n x batch (Julia's column-major layout is
# exactly the library's: one column-major block per problem), mQ n x batch, mL / mU m x batch, mX n x batch.
# A MATRIX pair (mP n x n, mA m x n) instead of the 3-D arrays selects the shared-matrix (MPC-style) engine.
# ---------------------------------------------------------------------------------------------------------------
"""
    SolveQuadraticProgramBatch!(mX, tP, mQ, tA, mL, mU; kw...) -> (vFlags, vIters, info)

Every problem is solved as `SolveQuadraticProgram!` with a direct plugin would (LinearSystemSolvers.jl:16-107).
"""
function SolveQuadraticProgramBatch!(mX::Matrix{Float64}, tP::Array{Float64, 3}, mQ::Matrix{Float64}, tA::Array{Float64, 3},
        mL::Matrix{Float64}, mU::Matrix{Float64}; kw...)
    n, batch = size(mQ); m = size(mL, 1);
    (size(tP) == (n, n, batch) && size(tA) == (m, n, batch) && size(mU) == (m, batch) && size(mX) == (n, batch)) ||
        throw(DimensionMismatch("batch dimensions"));
    s = B200Settings(; linSolver = :cholesky, kw...);
    vFlags = zeros(Int32, batch); vIters = zeros(Int64, batch); sInfo = Info();
    check(ccall((:qpb200_batch_solve_once, libqpb200), Cint,
        (Int64, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Settings}, Ptr{Float64},
         Ptr{Int32}, Ptr{Int64}, Ref{Info}), batch, n, m, tP, tA, mQ, mL, mU, s, mX, vFlags, vIters, sInfo));
    return vFlags, vIters, sInfo;
end

"""
    SolveQuadraticProgramBatch!(mX, mP::Matrix, mQ, mA::Matrix, mL, mU; kw...)

MPC-style batch: ONE `mP` (n x n) and ONE `mA` (m x n) for all columns of `mQ, mL, mU` (qpb200_batch_create_shared).
"""
function SolveQuadraticProgramBatch!(mX::Matrix{Float64}, mP::Matrix{Float64}, mQ::Matrix{Float64}, mA::Matrix{Float64},
        mL::Matrix{Float64}, mU::Matrix{Float64}; kw...)
    n, batch = size(mQ); m = size(mL, 1);
    (size(mP) == (n, n) && size(mA) == (m, n) && size(mU) == (m, batch) && size(mX) == (n, batch)) ||
        throw(DimensionMismatch("batch dimensions"));
    s = B200Settings(; linSolver = :cholesky, kw...);
    hRef = Ref{Ptr{Cvoid}}(C_NULL);
    check(ccall((:qpb200_batch_create_shared, libqpb200), Cint,
        (Ref{Ptr{Cvoid}}, Int64, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Settings}),
        hRef, batch, n, m, mP, mA, mQ, mL, mU, s));
    vFlags = zeros(Int32, batch); vIters = zeros(Int64, batch); sInfo = Info();
    try
        check(ccall((:qpb200_batch_solve, libqpb200), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Int32}, Ptr{Int64}, Ref{Info}),
            hRef[], mX, vFlags, vIters, sInfo));
    finally
        ccall((:qpb200_batch_destroy, libqpb200), Cvoid, (Ptr{Cvoid},), hRef[]);
    end
    return vFlags, vIters, sInfo;
end

# ---------------------------------------------------------------------------------------------------------------
# The reference's second solver (ProxQP.jl:118-173) on the GPU: a method for a sparse problem given by its pieces.
# (ProxQP.jl's own `struct ProxQP` keeps a CPU Cholesky factor; here the factor lives on the device, so the method takes
#  the problem data and the iterates -- the fields vX, vY, vZ, vS of that struct -- directly.)
# ---------------------------------------------------------------------------------------------------------------
"""
    SolveProxQP!(vX, vY, vZ, vS, mP, vQ, mA, vB, mC, vD; numIterations = 2000, ϵAbs = 1e-7, ϵRel = 1e-6, numItrConv = 50,
                 ρ = 1e2, σ = 1e-2, adptΡ = true, τ = 10.0, initSlack = false) -> Dict (the reference's dReport)
"""
function SolveProxQP!(vX::Vector{Float64}, vY::Vector{Float64}, vZ::Vector{Float64}, vS::Vector{Float64},
        mP::SparseMatrixCSC{Float64, Int64}, vQ::Vector{Float64}, mA::SparseMatrixCSC{Float64, Int64}, vB::Vector{Float64},
        mC::SparseMatrixCSC{Float64, Int64}, vD::Vector{Float64}; numIterations = 2000, ϵAbs = 1e-7, ϵRel = 1e-6,
        numItrConv = 50, ρ = 1e2, σ = 1e-2, adptΡ::Bool = true, τ = 10.0, initSlack::Bool = false, device::Integer = -1)
    s = Settings();
    ccall((:qpb200_proxqp_default_settings, libqpb200), Cvoid, (Ref{Settings},), s);
    s.max_iter = numIterations; s.eps_abs = ϵAbs; s.eps_rel = ϵRel; s.check_every = numItrConv; s.rho = ρ; s.sigma = σ;
    s.adaptive_rho = adptΡ; s.rho_factor = τ; s.device = device;
    mAC = SparseMatrixCSC{Float64, Int64}([mA; mC]);
    vLo = [vB; fill(-Inf, length(vD))]; vUp = [vB; vD];
    hRef = Ref{Ptr{Cvoid}}(C_NULL);
    GC.@preserve mP mAC vQ vLo vUp begin
        check(ccall((:qpb200_create, libqpb200), Cint,
            (Ref{Ptr{Cvoid}}, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64},
             Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Settings}, Int32),
            hRef, length(vX), size(mAC, 1), mP.colptr, mP.rowval, mP.nzval, mAC.colptr, mAC.rowval, mAC.nzval,
            vQ, vLo, vUp, s, Int32(1)));
    end
    sRep = ProxReport();
    try
        check(ccall((:qpb200_proxqp_solve, libqpb200), Cint,
            (Ptr{Cvoid}, Int64, Ref{Settings}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32, Ref{ProxReport}),
            hRef[], length(vB), s, vX, vY, vZ, vS, Int32(initSlack), sRep));
    finally
        ccall((:qpb200_destroy, libqpb200), Cvoid, (Ptr{Cvoid},), hRef[]);
    end
    return Dict{String, Real}("Converged" => sRep.converged != 0, "Iterations" => sRep.iterations, "ρ" => sRep.rho, "σ" => sRep.sigma,
        "PrimalResidual" => sRep.res_prim, "DualResidual" => sRep.res_dual);
end
