# B200AMG.jl — the reference-side binding of libb200amg.so (include/b200amg.h).
#
# NOT RUNNABLE IN THIS IMAGE (no Julia toolchain here or on the GPU box); it is the shim a maintainer of
# AlgebraicMultigrid.jl would add so that `_solve!`, `ldiv!` and `smooth!` of an existing `MultiLevel`
# (built by the package's own `ruge_stuben` / `smoothed_aggregation` on the host) run on a B200.
# The Python mirror in algebraicmultigrid.jl_b200/ calls exactly the same entry points through ctypes
# and is what the tests and benchmarks of this repository exercise.
module B200AMG

using AlgebraicMultigrid, SparseArrays, LinearAlgebra
import AlgebraicMultigrid: MultiLevel, Level, Cycle, V, W, F, Smoother, GaussSeidel, Jacobi, SOR,
                           ForwardSweep, BackwardSweep, SymmetricSweep, Pinv, QRSolver, Preconditioner
import LinearAlgebra: ldiv!

const lib = get(ENV, "B200AMG_LIB", "libb200amg")

# ---- b200amg_csc_t / b200amg_smoother_t (include/b200amg.h) -------------------------------------
struct CscDesc
    m::Int64; n::Int64
    colptr::Ptr{Cvoid}; rowval::Ptr{Cvoid}; nzval::Ptr{Float64}
    index_bits::Int32; index_base::Int32; adjoint::Int32; reserved::Int32
end
struct SmootherDesc
    kind::Int32; sweep::Int32; iter::Int32; reserved::Int32; omega::Float64
end

# a SparseMatrixCSC{Float64,Int64} (or its lazy Adjoint: RS stores P = R', SA stores R = P') crosses as is
csc(A::SparseMatrixCSC{Float64,Int64}; adj = false) =
    CscDesc(size(A, 1), size(A, 2), pointer(A.colptr), pointer(A.rowval), pointer(A.nzval), 64, 1, adj ? 1 : 0, 0)
csc(A::Adjoint{Float64,<:SparseMatrixCSC{Float64,Int64}}) = csc(parent(A); adj = true)

sweepcode(::ForwardSweep) = Int32(1); sweepcode(::BackwardSweep) = Int32(2); sweepcode(::SymmetricSweep) = Int32(3)
desc(s::GaussSeidel) = SmootherDesc(1, sweepcode(s.sweep), s.iter, 0, 1.0)
desc(s::Jacobi)      = SmootherDesc(2, 3, s.iter, 0, s.ω)
desc(s::SOR)         = SmootherDesc(3, sweepcode(s.sweep), s.iter, 0, s.ω)
cyclecode(::V) = Int32(0); cyclecode(::W) = Int32(1); cyclecode(::F) = Int32(2)

check(rc) = rc == 0 || error("b200amg error $rc: " * unsafe_string(ccall((:b200amg_last_error, lib), Cstring, ())))

const DENSE_COARSE_LIMIT = 16384
const COARSE_SOLVERS = Dict{Ptr{Cvoid},Any}()      # handle => the hierarchy's coarse solver (kept alive for the callback)

"""C callback of `b200amg_set_coarse_callback`: `coarse_solver(x, b)` (src/multilevel.jl:180,228) on the staging vectors."""
function coarse_trampoline(user::Ptr{Cvoid}, n::Int64, ncols::Int64, x::Ptr{Float64}, b::Ptr{Float64})::Int32
    try
        COARSE_SOLVERS[user](unsafe_wrap(Array, x, n), unsafe_wrap(Array, b, n))
        return Int32(0)
    catch
        return Int32(1)
    end
end

"""Device-resident copy of a `MultiLevel`: upload once, reuse for every solve."""
mutable struct DeviceMultiLevel
    handle::Ptr{Cvoid}
    ml::MultiLevel
    n::Int
end

function DeviceMultiLevel(ml::MultiLevel, pre::Smoother = GaussSeidel(), post::Smoother = GaussSeidel();
                          device::Integer = 0, symmetry::Integer = 0)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:b200amg_create, lib), Int32, (Ref{Ptr{Cvoid}}, Int32), h, device))
    GC.@preserve ml begin
        for l in ml.levels      # push!(levels, Level(A, P, R, pre, post))  classical.jl:48-52, aggregation.jl:147-151
            a, p, r = Ref(csc(l.A)), Ref(csc(l.P)), Ref(csc(l.R))
            check(ccall((:b200amg_add_level, lib), Int32,
                        (Ptr{Cvoid}, Ref{CscDesc}, Ref{CscDesc}, Ref{CscDesc}, Ref{SmootherDesc}, Ref{SmootherDesc}, Int32),
                        h[], a, p, r, Ref(desc(pre)), Ref(desc(post)), symmetry))
        end
        n = size(ml.final_A, 1)
        if n <= DENSE_COARSE_LIMIT
            Minv = Matrix(pinv(Matrix(ml.final_A)))  # Pinv (coarse_solver.jl:9-16); inv(A) for the QR/LU solvers
            check(ccall((:b200amg_set_coarse, lib), Int32, (Ptr{Cvoid}, Ref{CscDesc}, Int64, Ptr{Float64}),
                        h[], Ref(csc(ml.final_A)), n, Minv))
        else
            # a coarsest level too large for a dense operator: the reference's own callable (sparse QR / LinearSolve
            # factorisation, coarse_solver.jl:24-58,66-81) runs on the host from inside the cycle
            COARSE_SOLVERS[h[]] = ml.coarse_solver
            check(ccall((:b200amg_set_coarse_callback, lib), Int32, (Ptr{Cvoid}, Ref{CscDesc}, Int64, Ptr{Cvoid}, Ptr{Cvoid}),
                        h[], Ref(csc(ml.final_A)), n,
                        @cfunction(coarse_trampoline, Int32, (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}, Ptr{Float64})), h[]))
        end
    end
    check(ccall((:b200amg_finalize, lib), Int32, (Ptr{Cvoid},), h[]))
    d = DeviceMultiLevel(h[], ml, isempty(ml.levels) ? size(ml.final_A, 1) : size(ml.levels[1].A, 1))
    finalizer(x -> (delete!(COARSE_SOLVERS, x.handle); ccall((:b200amg_destroy, lib), Int32, (Ptr{Cvoid},), x.handle)), d)
    return d
end

"""`_solve!(x, ml, b, cycle; ...)` (src/multilevel.jl:158-198) on the device."""
function AlgebraicMultigrid._solve!(x::Vector{Float64}, d::DeviceMultiLevel, b::Vector{Float64}, cycle::Cycle = V();
                                    maxiter::Int = 100, abstol::Real = zero(Float64), reltol::Real = sqrt(eps(Float64)),
                                    verbose::Bool = false, log::Bool = false, calculate_residual = true, kwargs...)
    residuals = Vector{Float64}(undef, maxiter + 2)
    nres, iters = Ref{Int32}(0), Ref{Int32}(0)
    check(ccall((:b200amg_solve, lib), Int32,
                (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int32, Int32, Float64, Float64, Int32, Ptr{Float64}, Int32,
                 Ref{Int32}, Ref{Int32}, Int32),
                d.handle, x, b, cyclecode(cycle), maxiter, abstol, reltol, calculate_residual ? 1 : 0, residuals,
                length(residuals), nres, iters, 0))
    resize!(residuals, nres[])
    if verbose && calculate_residual          # the reference prints the PREVIOUS residual (multilevel.jl:185-187)
        for itr in 1:iters[]
            AlgebraicMultigrid.Printf.@printf "Norm of residual at iteration %6d is %.4e\n" itr residuals[itr]
        end
    end
    return log ? (x, residuals) : x
end
AlgebraicMultigrid._solve(d::DeviceMultiLevel, b::Vector{Float64}, args...; kwargs...) =
    AlgebraicMultigrid._solve!(zeros(Float64, size(b)), d, b, args...; kwargs...)

"""Block right-hand sides (`Val{bs}` workspaces, src/multilevel.jl:28-59; one Frobenius norm, :170,190): every column stays
on the device for the whole call (`b200amg_solve_block`)."""
function AlgebraicMultigrid._solve!(x::Matrix{Float64}, d::DeviceMultiLevel, b::Matrix{Float64}, cycle::Cycle = V();
                                    maxiter::Int = 100, abstol::Real = zero(Float64), reltol::Real = sqrt(eps(Float64)),
                                    log::Bool = false, calculate_residual = true, kwargs...)
    size(x) == size(b) || throw(DimensionMismatch("x is $(size(x)), b is $(size(b))"))
    residuals = Vector{Float64}(undef, maxiter + 2)
    nres, iters = Ref{Int32}(0), Ref{Int32}(0)
    check(ccall((:b200amg_solve_block, lib), Int32,
                (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Int64, Int32, Int32, Float64, Float64, Int32, Ptr{Float64}, Int32,
                 Ref{Int32}, Ref{Int32}, Int32),
                d.handle, x, b, size(b, 2), stride(b, 2), cyclecode(cycle), maxiter, abstol, reltol, calculate_residual ? 1 : 0,
                residuals, length(residuals), nres, iters, 0))
    resize!(residuals, nres[])
    return log ? (x, residuals) : x
end
AlgebraicMultigrid._solve(d::DeviceMultiLevel, b::Matrix{Float64}, args...; kwargs...) =
    AlgebraicMultigrid._solve!(zeros(Float64, size(b)), d, b, args...; kwargs...)

"""`ldiv!(x, p, b)` (src/preconditioner.jl:12-19): x .= 0 (or b), one cycle, no residual."""
struct DevicePreconditioner{C<:Cycle}
    d::DeviceMultiLevel
    init::Symbol
    cycle::C
end
AlgebraicMultigrid.aspreconditioner(d::DeviceMultiLevel, cycle::Cycle = V()) = DevicePreconditioner(d, :zero, cycle)
function ldiv!(x::Vector{Float64}, p::DevicePreconditioner, b::Vector{Float64})
    check(ccall((:b200amg_precond, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int32, Int32, Int32),
                p.d.handle, x, b, cyclecode(p.cycle), p.init == :zero ? 1 : 0, 0))
    return x
end
ldiv!(p::DevicePreconditioner, b) = copyto!(b, ldiv!(similar(b), p, b))
Base.:\(p::DevicePreconditioner, b) = ldiv!(similar(b), p, b)

"""Device-resident preconditioned CG (what the reference's tests get from IterativeSolvers.cg(A, b; Pl = p))."""
function cg(p::DevicePreconditioner, b::Vector{Float64}; abstol = 0.0, reltol = sqrt(eps(Float64)), maxiter = length(b))
    x = zeros(length(b)); res = Vector{Float64}(undef, maxiter + 2); nres, iters = Ref{Int32}(0), Ref{Int32}(0)
    check(ccall((:b200amg_pcg, lib), Int32,
                (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int32, Int32, Float64, Float64, Ptr{Float64}, Int32, Ref{Int32}, Ref{Int32}, Int32),
                p.d.handle, x, b, cyclecode(p.cycle), maxiter, abstol, reltol, res, length(res), nres, iters, 0))
    return x
end

"""`setup_smoother(config, A, symmetry)` / `smooth!(x, s, b)` (src/smoother.jl:1-49) as a device object."""
mutable struct DeviceSmoother
    handle::Ptr{Cvoid}
end
function AlgebraicMultigrid.setup_smoother(config::Smoother, A::SparseMatrixCSC{Float64,Int64}, symmetry, ::Val{:b200})
    s = Ref{Ptr{Cvoid}}(C_NULL)
    sym = symmetry isa AlgebraicMultigrid.NoSymmetry ? 1 : 0
    rc = GC.@preserve A ccall((:b200amg_smoother_create, lib), Int32, (Ref{Ptr{Cvoid}}, Int32, Ref{CscDesc}, Ref{SmootherDesc}, Int32),
                              s, 0, Ref(csc(A)), Ref(desc(config)), sym)
    rc == -3 && throw(SingularException(0))      # DiagonalIndices, smoother.jl:239-241
    check(rc)
    d = DeviceSmoother(s[])
    finalizer(x -> ccall((:b200amg_smoother_destroy, lib), Int32, (Ptr{Cvoid},), x.handle), d)
    return d
end
function AlgebraicMultigrid.smooth!(x::Vector{Float64}, s::DeviceSmoother, b::Vector{Float64})
    check(ccall((:b200amg_smoother_apply, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int32), s.handle, x, b, 0))
    return nothing
end

"""
    galerkin(R, A, P) -> R*A*P

The Galerkin product of `extend_hierarchy!` (src/classical.jl:46, src/aggregation.jl:145) on the device
(`b200amg_spgemm_begin` / `_fetch`): stdlib semantics (sorted rows per column, structural zeros kept) and the
sequential accumulation order, so the result is bit-identical to `(R*A)*P` on the host.  The library takes 0-based
`Int32` compressed-sparse-column arrays.
"""
function spgemm(A::SparseMatrixCSC{Float64,Int64}, B::SparseMatrixCSC{Float64,Int64}; device::Integer = 0)
    size(A, 2) == size(B, 1) || throw(DimensionMismatch("A has $(size(A, 2)) columns, B has $(size(B, 1)) rows"))
    ap, aj = Int32.(A.colptr .- 1), Int32.(A.rowval .- 1)
    bp, bj = Int32.(B.colptr .- 1), Int32.(B.rowval .- 1)
    nnzc = Ref{Int64}(0)
    check(ccall((:b200amg_spgemm_begin, lib), Int32,
                (Int32, Int64, Int64, Int64, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ref{Int64}),
                device, size(A, 1), size(A, 2), size(B, 2), ap, aj, A.nzval, bp, bj, B.nzval, nnzc))
    cp, cj, cx = Vector{Int32}(undef, size(B, 2) + 1), Vector{Int32}(undef, nnzc[]), Vector{Float64}(undef, nnzc[])
    check(ccall((:b200amg_spgemm_fetch, lib), Int32, (Ptr{Int32}, Ptr{Int32}, Ptr{Float64}), cp, cj, cx))
    return SparseMatrixCSC(size(A, 1), size(B, 2), Int64.(cp) .+ 1, Int64.(cj) .+ 1, cx)
end
galerkin(R, A, P) = spgemm(spgemm(R, A), P)

"""
    partition!(h, rank, world, nccl_id; levels = 1)

One process per GPU (`b200amg_set_partition`): call on a fresh handle before the first `add_level`; `levels` finest
levels are split by rows over the ranks (`B200AMG_OPT_PART_LEVELS` = option 12).  `nccl_id` is the 128-byte id rank 0
obtained from `b200amg_nccl_unique_id` and broadcast by the host (MPI.jl, a file, ...).
"""
function partition!(h::Ptr{Cvoid}, rank::Integer, world::Integer, nccl_id::Vector{UInt8}; levels::Integer = 1)
    check(ccall((:b200amg_set_partition, lib), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{UInt8}, Int64), h, rank, world, nccl_id, length(nccl_id)))
    levels == 1 || check(ccall((:b200amg_set_option, lib), Int32, (Ptr{Cvoid}, Int32, Float64), h, 12, Float64(levels)))
    return h
end

end # module
