# DREB200.jl -- drop-in glue that routes the LRSIF-ADI hot path of DifferentialRiccatiEquations.jl
# through libdre_b200.so (C ABI in include/dre_b200.h) with plain `ccall`s: no CUDA.jl array
# dispatch, no multi-backend layer, no CPU fallback.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: neither the build container nor the GPU box has a Julia
# binary.  The Python mirror (differentialriccatiequations.jl_b200/api.py) implements the same
# methods over the same symbols and is what the tests exercise.
#
# Usage:  using DifferentialRiccatiEquations, DREB200
#         prob = GDREProblem(E, A, B, C, DREB200.lowrank_device(E, A, L0, D0), tspan)
#         sol  = solve(prob, Ros1(); dt = -100)          # K(t) as usual
module DREB200

using DifferentialRiccatiEquations
const DRE = DifferentialRiccatiEquations
using LinearAlgebra, SparseArrays
import CommonSolve

const LIB = get(ENV, "DRE_B200_LIB", joinpath(@__DIR__, "..", "differentialriccatiequations.jl_b200", "libdre_b200.so"))

struct View            # mirrors dre_view (isbits, passed by value)
    id::Int32
    col0::Int32
    ncols::Int32
end

mutable struct Context
    h::Ptr{Cvoid}
    n::Int
end

function check(ctx, rc::Int32)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:dre_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx === nothing ? C_NULL : ctx.h))
    # numerical collapse maps onto the reference's "Increment is zero" path (src/lyapunov/adi.jl:161-165)
    error("libdre_b200 error $rc: $msg")
end

function Context(device::Integer = 0)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:dre_create, LIB), Int32, (Int32, Ref{Ptr{Cvoid}}), device, h)
    rc == 0 || error(unsafe_string(ccall((:dre_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
    ctx = Context(h[], 0)
    finalizer(c -> ccall((:dre_destroy, LIB), Int32, (Ptr{Cvoid},), c.h), ctx)
end

const CTX = Ref{Union{Nothing,Context}}(nothing)
context() = (CTX[] === nothing && (CTX[] = Context()); CTX[])

# SparseMatrixCSC{Float64,Int64} is passed zero-copy (1-based indices, index_base = 1)
function set_pencil!(ctx::Context, E::SparseMatrixCSC{Float64,Int64}, A::SparseMatrixCSC{Float64,Int64})
    n = size(E, 1)
    GC.@preserve E A check(ctx, ccall((:dre_set_pencil, LIB), Int32,
        (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Int32),
        ctx.h, n, E.colptr, E.rowval, E.nzval, A.colptr, A.rowval, A.nzval, 1))
    ctx.n = n
end

# ---- device panel type: the TL of LDLᵀ{T,TL,TD} (src/LDLt.jl:29-33) ----
mutable struct Panel
    ctx::Context
    id::Int32
    cols::Int
end
struct DeviceMatrix <: AbstractMatrix{Float64}
    p::Panel
    col0::Int
    ncols::Int
end
Base.size(M::DeviceMatrix) = (M.p.ctx.n, M.ncols)
view_of(M::DeviceMatrix) = View(M.p.id, M.col0, M.ncols)

function DeviceMatrix(ctx::Context, cols::Integer)
    id = Ref{Int32}(-1)
    check(ctx, ccall((:dre_mat_create, LIB), Int32, (Ptr{Cvoid}, Int32, Ref{Int32}), ctx.h, cols, id))
    p = Panel(ctx, id[], cols)
    finalizer(q -> ccall((:dre_mat_free, LIB), Int32, (Ptr{Cvoid}, Int32), q.ctx.h, q.id), p)
    DeviceMatrix(p, 0, cols)
end
function DeviceMatrix(ctx::Context, M::Matrix{Float64})
    D = DeviceMatrix(ctx, size(M, 2))
    GC.@preserve M check(ctx, ccall((:dre_mat_upload, LIB), Int32, (Ptr{Cvoid}, View, Ptr{Float64}, Int64),
                                    ctx.h, view_of(D), M, stride(M, 2)))
    D
end
function Base.Matrix(D::DeviceMatrix)
    M = Matrix{Float64}(undef, size(D)...)
    check(D.p.ctx, ccall((:dre_mat_download, LIB), Int32, (Ptr{Cvoid}, View, Ptr{Float64}, Int64),
                         D.p.ctx.h, view_of(D), M, stride(M, 2)))
    M
end
Base.similar(D::DeviceMatrix) = DeviceMatrix(D.p.ctx, D.ncols)
view_cols(D::DeviceMatrix, r::UnitRange{Int}) = DeviceMatrix(D.p, D.col0 + first(r) - 1, length(r))   # no copy

lowrank_device(E, A, L::Matrix, D::Matrix) = (set_pencil!(context(), E, A); DRE.lowrank(DeviceMatrix(context(), L), D))

# ---- specialised methods of the reference's generic functions ----

# orthf/compress!/norm (src/LDLt.jl:77-89, 204-245)
function DRE.compress!(X::DRE.LDLᵀ{Float64,DeviceMatrix,Matrix{Float64}})
    ctx = first(X.Ls).p.ctx
    nt = length(X.Ls)
    views = [view_of(L) for L in X.Ls]
    Dp = [pointer(D) for D in X.Ds]
    ldds = Int64[stride(D, 2) for D in X.Ds]
    cap = min(sum(L -> L.ncols, X.Ls), ctx.n)
    out = DeviceMatrix(ctx, cap)
    lam = Vector{Float64}(undef, cap)
    newrank = Ref{Int32}(0)
    GC.@preserve X check(ctx, ccall((:dre_ldlt_compress, LIB), Int32,
        (Ptr{Cvoid}, Int32, Ptr{View}, Ptr{Ptr{Float64}}, Ptr{Int64}, Ptr{Float64}, Float64, View, Ptr{Float64}, Ref{Int32}),
        ctx.h, nt, views, Dp, ldds, X.alphas, 100.0, view_of(out), lam, newrank))
    k = Int(newrank[])
    resize!(X.alphas, 1); resize!(X.Ls, 1); resize!(X.Ds, 1)
    X.alphas[1] = 1.0
    X.Ls[1] = DeviceMatrix(out.p, 0, k)
    X.Ds[1] = diagm(lam[1:k])
    X
end

function LinearAlgebra.norm(X::DRE.LDLᵀ{Float64,DeviceMatrix,Matrix{Float64}})
    DRE.concatenate!(X)
    L, D, a = only(X.Ls), only(X.Ds), only(X.alphas)
    out = Ref{Float64}(0.0)
    GC.@preserve D check(L.p.ctx, ccall((:dre_ldlt_norm, LIB), Int32,
        (Ptr{Cvoid}, View, Ptr{Float64}, Int64, Float64, Ref{Float64}),
        L.p.ctx.h, view_of(L), D, stride(D, 2), a, out))
    out[]
end

# the closed-loop operator handed to ADI: (a, e) of a*A + e*E plus the low-rank factors
struct DeviceOperator
    a::Float64
    e::Float64
    alpha::Float64
    U::Union{Nothing,DeviceMatrix}    # B
    Vt::Union{Nothing,DeviceMatrix}   # K'
end

# perform_single_step! / perform_double_step! (src/lyapunov/adi.jl:149-225): ONE ccall per ADI step
function adi_step!(ctx::Context, F::DeviceOperator, μ::Complex, R::DeviceMatrix)
    z = View(-1, 0, 0)
    check(ctx, ccall((:dre_set_operator, LIB), Int32, (Ptr{Cvoid}, Float64, Float64, Float64, View, View),
                     ctx.h, F.a, F.e, F.alpha, F.U === nothing ? z : view_of(F.U), F.Vt === nothing ? z : view_of(F.Vt)))
    V1 = similar(R)
    V2 = iszero(imag(μ)) ? nothing : similar(R)
    check(ctx, ccall((:dre_adi_step, LIB), Int32, (Ptr{Cvoid}, Float64, Float64, View, View, View),
                     ctx.h, real(μ), imag(μ), view_of(R), view_of(V1), V2 === nothing ? z : view_of(V2)))
    V1, V2
end

# After compress! the outer factor has orthonormal columns; telling the library lets the next compress! adopt
# them as basis vectors without re-orthogonalisation (call right before compress! on X + increments).
hint_orthonormal!(L::DeviceMatrix) =
    check(L.p.ctx, ccall((:dre_hint_orthonormal, LIB), Int32, (Ptr{Cvoid}, View), L.p.ctx.h, view_of(L)))

# The context's CUDA stream (cudaStream_t) for stream-ordered collectives on the raw panel pointers.
function cuda_stream(ctx::Context)
    s = Ref{Ptr{Cvoid}}(C_NULL)
    check(ctx, ccall((:dre_get_stream, LIB), Int32, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}), ctx.h, s))
    s[]
end

# Performance hint (results unchanged): the next shifts of the buffered iterator (src/shifts/helpers.jl:106-113)
# are factored ahead on the library's side streams while the current step's sweeps / Gram / compression run.
prefactor!(ctx::Context, μ::Complex) =
    check(ctx, ccall((:dre_prefactor, LIB), Int32, (Ptr{Cvoid}, Float64, Float64), ctx.h, real(μ), imag(μ)))

# Multi-GPU (one Julia process per GPU, e.g. MPI.jl ranks): every rank solves its own block of R's columns with
# the replicated factorization, then the solved blocks are exchanged.  `dre_mat_devptr` hands the raw device
# address of a view (row-major, leading dimension ld) to the collective library (NCCL.jl / CUDA-aware MPI);
# the residual update runs afterwards on the full panel through `spmm!`.
function adi_solve_block!(ctx::Context, μ::Complex, R::DeviceMatrix, V1::DeviceMatrix, V2, cols::UnitRange{Int})
    sub(M) = View(M.p.id, M.col0 + first(cols) - 1, length(cols))
    z = View(-1, 0, 0)
    check(ctx, ccall((:dre_adi_solve, LIB), Int32, (Ptr{Cvoid}, Float64, Float64, View, View, View),
                     ctx.h, real(μ), imag(μ), sub(R), sub(V1), V2 === nothing ? z : sub(V2)))
end
function devptr(M::DeviceMatrix)
    p = Ref{Ptr{Cvoid}}(C_NULL); ld = Ref{Int64}(0)
    check(M.p.ctx, ccall((:dre_mat_devptr, LIB), Int32, (Ptr{Cvoid}, View, Ref{Ptr{Cvoid}}, Ref{Int64}),
                         M.p.ctx.h, view_of(M), p, ld))
    check(M.p.ctx, ccall((:dre_sync, LIB), Int32, (Ptr{Cvoid},), M.p.ctx.h))
    p[], ld[]
end

# E'L, A'L  (src/lyapunov/residual.jl:18, src/riccati/lowrank_ros1.jl:42)
function spmm!(Y::DeviceMatrix, op::Char, X::DeviceMatrix, α::Real, β::Real)
    check(X.p.ctx, ccall((:dre_spmm, LIB), Int32, (Ptr{Cvoid}, Int32, Float64, View, Float64, View),
                         X.p.ctx.h, Int32(op), α, view_of(X), β, view_of(Y)))
    Y
end

# B'L etc.: X'Y to a host Matrix
function gemm_tn(X::DeviceMatrix, Y::DeviceMatrix)
    out = Matrix{Float64}(undef, X.ncols, Y.ncols)
    check(X.p.ctx, ccall((:dre_gemm_tn, LIB), Int32, (Ptr{Cvoid}, View, View, Ptr{Float64}, Int64),
                         X.p.ctx.h, view_of(X), view_of(Y), out, stride(out, 2)))
    out
end

# dot(::LDLᵀ, ::LDLᵀ) (src/LDLt.jl:91-108), the inner product of the low-rank (F)GMRES (src/lyapunov/gmres.jl:52-55):
# the reference loops over the n rows of the outer factors; here one Gram product A'C on the device and
# k1 x k2 algebra on the host.  With this method (and compress!/norm above, spmm! for LyapunovOperator * X) the
# reference's generic GMRES driver runs unchanged on DeviceMatrix-backed LDLᵀ objects.
function LinearAlgebra.dot(X1::DRE.LDLᵀ{Float64,DeviceMatrix,Matrix{Float64}},
                           X2::DRE.LDLᵀ{Float64,DeviceMatrix,Matrix{Float64}})
    DRE.concatenate!(X1)
    DRE.concatenate!(X2)
    α, A, B = X1.alphas[1], X1.Ls[1], X1.Ds[1]
    β, C, D = X2.alphas[1], X2.Ls[1], X2.Ds[1]
    AtC = gemm_tn(A, C)
    M = (B' * AtC * D) .* (α * β)
    sum(AtC .* M)
end

function Base.:(*)(L::DRE.LyapunovOperator, X::DRE.LDLᵀ{Float64,DeviceMatrix,Matrix{Float64}})   # gmres.jl:105-117
    a, Z, Y = X
    k = size(Z, 2)
    Z2 = DeviceMatrix(Z.p.ctx, 2k)
    spmm!(view_cols(Z2, 1:k), 'E', Z, 1.0, 0.0)
    spmm!(view_cols(Z2, k+1:2k), 'A', Z, 1.0, 0.0)          # symmetric pencil: A'Z == AZ (closed loop: see api._adj_matmul)
    O = zero(Y)
    a * DRE.lowrank(Z2, [O Y; Y O])
end

# The remaining specialisations (residual, take_many!(::ProjectionShiftIterator), the K update and RHS
# assembly of lowrank_ros1/2.jl and newton.jl) follow api.py function by function:
#   api.residual          -> DRE.residual(::GALEProblem{<:LDLᵀ{…DeviceMatrix…}}, ::LDLᵀ)
#   api.orth_restrict     -> Shifts.take_many!(::ProjectionShiftIterator)   (dre_rrqr + host svd/eigvals)
#   api._feedback         -> K = (B'L D)(L'E)                               (dre_gemm_tn, dre_spmm, dre_gemm_nn)
end # module
