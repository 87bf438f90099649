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

# compress! in three phases for a process that receives the terms of X one by one (the compression lane of the
# multi-GPU pipeline mode: `dre_b200.dist.serve` is the executable pattern).  compress_begin! reserves room,
# compress_add! orthogonalises terms against the basis built so far, compress_finish! returns (L, lambda);
# compress!(X) above is the three in one call.  compress_scale_hint! is for a lane that adds X last.
compress_begin!(ctx::Context, max_cols::Integer; tol_factor = 100.0) =
    check(ctx, ccall((:dre_compress_begin, LIB), Int32, (Ptr{Cvoid}, Int32, Float64), ctx.h, max_cols, tol_factor))
compress_scale_hint!(ctx::Context, scale::Real) =
    check(ctx, ccall((:dre_compress_scale_hint, LIB), Int32, (Ptr{Cvoid}, Float64), ctx.h, scale))
function compress_add!(ctx::Context, alphas::Vector{Float64}, Ls::Vector{DeviceMatrix}, Ds::Vector{Matrix{Float64}})
    nt = length(Ls)
    views = [view_of(L) for L in Ls]
    Dp = [pointer(D) for D in Ds]
    ldds = Int64[max(stride(D, 2), 1) for D in Ds]
    GC.@preserve Ds check(ctx, ccall((:dre_compress_add, LIB), Int32,
        (Ptr{Cvoid}, Int32, Ptr{View}, Ptr{Ptr{Float64}}, Ptr{Int64}, Ptr{Float64}),
        ctx.h, nt, views, Dp, ldds, alphas))
end
function compress_finish!(ctx::Context, cap::Integer)
    out = DeviceMatrix(ctx, cap)
    lam = Vector{Float64}(undef, cap)
    newrank = Ref{Int32}(0)
    check(ctx, ccall((:dre_compress_finish, LIB), Int32, (Ptr{Cvoid}, View, Ptr{Float64}, Ref{Int32}),
                     ctx.h, view_of(out), lam, newrank))
    k = Int(newrank[])
    DeviceMatrix(out.p, 0, k), lam[1:k]
end

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

# ==============================================================================================
# The hooks that make the reference's OWN drivers run on DeviceMatrix-backed problems.
#
#   using DifferentialRiccatiEquations, DREB200
#   ctx  = DREB200.context();  Ed, Ad = DREB200.pencil!(ctx, E, A)          # uploads once, symbolic analysis once
#   X0   = lowrank(DREB200.DeviceMatrix(ctx, L0), D0)
#   prob = GDREProblem(Ed, Ad, B, C, X0, tspan)                             # B, C stay host matrices
#   sol  = solve(prob, Ros1(ADI(inner_alg = DREB200.Solver())); dt = -100)  # reference driver, unmodified
#
# Everything below is a METHOD OF A REFERENCE FUNCTION (or of Base/LinearAlgebra functions the reference's generic
# code calls) specialised on the three types of this file:
#   DeviceMatrix  n x k panel           (TL of LDLᵀ{T,TL,TD};  Adjoint{Float64,DeviceMatrix} = k x n, e.g. K)
#   PencilOp      a*A + e*E, lazy       (<: AbstractSparseMatrix, so lr_update/`+` keep it lazy: LowRankUpdate.jl:38-39,66-70)
#   Solver        <: BlockLinearSolver  (the `inner_alg` seam, blocklinear/types.jl:15-30)
# ==============================================================================================

# ---- PencilOp: the sparse part of every operator on the path is a combination of the two uploaded matrices ----
# Ros1: A - E/(2τ) (lowrank_ros1.jl:39), Ros2: γτA - E/2 (lowrank_ros2.jl), ADI: A' + (μE)' (adi.jl:156,195).
struct PencilOp{T<:Number} <: SparseArrays.AbstractSparseMatrix{T,Int64}
    ctx::Context
    a::T        # coefficient of A
    e::T        # coefficient of E
end
function pencil!(ctx::Context, E::SparseMatrixCSC{Float64,Int64}, A::SparseMatrixCSC{Float64,Int64})
    set_pencil!(ctx, E, A)
    PencilOp(ctx, 0.0, 1.0), PencilOp(ctx, 1.0, 0.0)          # (E, A)
end
Base.size(P::PencilOp) = (P.ctx.n, P.ctx.n)
Base.size(P::PencilOp, i::Integer) = i <= 2 ? P.ctx.n : 1
SparseArrays.issparse(::PencilOp) = true                        # LowRankUpdate.jl:67 asserts it
Base.adjoint(P::PencilOp{<:Real}) = P                           # symmetric pencil (dre_symbolic_create checks it)
Base.adjoint(P::PencilOp{<:Complex}) = PencilOp(P.ctx, conj(P.a), conj(P.e))   # (conj(μ)E)' == μE', adi.jl:195
Base.transpose(P::PencilOp) = P
Base.:+(P::PencilOp, Q::PencilOp) = PencilOp(P.ctx, P.a + Q.a, P.e + Q.e)
Base.:-(P::PencilOp, Q::PencilOp) = PencilOp(P.ctx, P.a - Q.a, P.e - Q.e)
Base.:-(P::PencilOp) = PencilOp(P.ctx, -P.a, -P.e)
Base.:*(s::Number, P::PencilOp) = PencilOp(P.ctx, s * P.a, s * P.e)
Base.:*(P::PencilOp, s::Number) = s * P
Base.:/(P::PencilOp, s::Number) = PencilOp(P.ctx, P.a / s, P.e / s)

# P*L, P'L  (E'L, A'L: residual.jl:18, lowrank_ros1.jl:42, adi.jl:171,217)
function Base.:*(P::PencilOp{<:Real}, L::DeviceMatrix)
    Y = similar(L)
    iszero(P.e) || spmm!(Y, 'E', L, P.e, 0.0)
    iszero(P.a) || spmm!(Y, 'A', L, P.a, iszero(P.e) ? 0.0 : 1.0)
    iszero(P.e) && iszero(P.a) && axpby!(Y, 0.0, Y, 0.0)
    Y
end
# mul!(R, E', V, -2μ, true): the residual update of adi.jl:171,217 when the generic step methods are used
function LinearAlgebra.mul!(Y::DeviceMatrix, P::PencilOp{<:Real}, X::DeviceMatrix, α::Number, β::Number)
    iszero(P.e) || spmm!(Y, 'E', X, α * P.e, β)
    iszero(P.a) || spmm!(Y, 'A', X, α * P.a, iszero(P.e) ? β : 1.0)
    Y
end
LinearAlgebra.mul!(Y::DeviceMatrix, P::PencilOp{<:Real}, X::DeviceMatrix) = mul!(Y, P, X, true, false)
# L'E as the lazy adjoint of E'L (k x n "row panel", lowrank_ros1.jl:28,56)
Base.:*(Lt::Adjoint{Float64,DeviceMatrix}, P::PencilOp{<:Real}) = (P' * parent(Lt))'

# ---- small-matrix products with panels ----
axpby!(Y::DeviceMatrix, α, X::DeviceMatrix, β) =
    (check(Y.p.ctx, ccall((:dre_mat_axpby, LIB), Int32, (Ptr{Cvoid}, Float64, View, Float64, View),
                          Y.p.ctx.h, α, view_of(X), β, view_of(Y))); Y)
function gemm_nn!(Y::DeviceMatrix, X::DeviceMatrix, W::Matrix{Float64}, α::Real, β::Real)
    GC.@preserve W check(X.p.ctx, ccall((:dre_gemm_nn, LIB), Int32,
        (Ptr{Cvoid}, Float64, View, Ptr{Float64}, Int64, Float64, View),
        X.p.ctx.h, α, view_of(X), W, stride(W, 2), β, view_of(Y)))
    Y
end
# host matrices that multiply panels (B, C') are uploaded once and cached by identity
# (B is n x m, C is q x n: whichever orientation has n rows is the panel -- B, B' -> B;  C, C' -> C')
const HOSTCACHE = IdDict{Any,DeviceMatrix}()
function device_of(ctx::Context, M::Union{Matrix{Float64},Adjoint{Float64,Matrix{Float64}}})
    key = M isa Adjoint ? parent(M) : M
    get!(HOSTCACHE, key) do
        DeviceMatrix(ctx, size(M, 1) == ctx.n ? Matrix(M) : Matrix(M'))
    end
end
device_of(::Context, M::DeviceMatrix) = M
device_of(::Context, M::Adjoint{Float64,DeviceMatrix}) = parent(M)

# B'L -> host m x k  (lowrank_ros1.jl:26,54; newton.jl; smw)
Base.:*(Bt::Adjoint{Float64,Matrix{Float64}}, L::DeviceMatrix) = gemm_tn(device_of(L.p.ctx, parent(Bt)), L)
Base.:*(Lt::Adjoint{Float64,DeviceMatrix}, B::Matrix{Float64}) = gemm_tn(parent(Lt), device_of(parent(Lt).p.ctx, B))
Base.:*(Xt::Adjoint{Float64,DeviceMatrix}, Y::DeviceMatrix) = gemm_tn(parent(Xt), Y)          # Q'(EQ), V*Q with V = K
# L*W (W small host) and W*(L'E): K = (B'L D) * (L'E) stays a lazy adjoint of an n x m panel
Base.:*(L::DeviceMatrix, W::Matrix{Float64}) = gemm_nn!(DeviceMatrix(L.p.ctx, size(W, 2)), L, W, 1.0, 0.0)
Base.:*(W::Matrix{Float64}, Xt::Adjoint{Float64,DeviceMatrix}) = (parent(Xt) * Matrix(W'))'
LinearAlgebra.mul!(Y::DeviceMatrix, U::DeviceMatrix, W::Matrix{Float64}, α::Number, β::Number) = gemm_nn!(Y, U, W, α, β)
Base.Matrix(Kt::Adjoint{Float64,DeviceMatrix}) = Matrix(Matrix(parent(Kt))')                    # K(t) for the user
# adapt(TL, BᵀLD) / adapt(TD, B'L) of lowrank_ros1.jl:26-28: small matrices stay on the host
import Adapt
Adapt.adapt_storage(::Type{DeviceMatrix}, M::Matrix{Float64}) = M
Adapt.adapt_storage(::Type{<:Matrix}, M::Matrix{Float64}) = M

# ---- _hcat / similar / column assignment (util/_hcat.jl:5-18; LDLt.jl:174-191) ----
Base.similar(::Type{DeviceMatrix}, m::Int, k::Int) = (ctx = context(); @assert m == ctx.n; DeviceMatrix(ctx, k))
function Base.setindex!(L::DeviceMatrix, X::DeviceMatrix, ::Colon, span::UnitRange{Int})
    check(L.p.ctx, ccall((:dre_mat_copy, LIB), Int32, (Ptr{Cvoid}, View, View), L.p.ctx.h, view_of(view_cols(L, span)), view_of(X)))
    X
end
Base.setindex!(L::DeviceMatrix, X::Union{Matrix{Float64},Adjoint{Float64,Matrix{Float64}}}, c::Colon, span::UnitRange{Int}) =
    setindex!(L, device_of(L.p.ctx, X), c, span)                                               # C' in _hcat(TL, C', E'L)
Base.hcat(Xs::DeviceMatrix...) = DRE._hcat(DeviceMatrix, Xs)                                   # projection.jl:58
Base.copy(L::DeviceMatrix) = (Y = similar(L); Y[:, 1:L.ncols] = L; Y)
Base.deepcopy_internal(L::DeviceMatrix, ::IdDict) = copy(L)                                     # residual.jl:9

# ---- residual(::GALEProblem{<:LDLᵀ}, ::LDLᵀ)  (lyapunov/residual.jl:3-31) ----
# The generic method already works through the products above; this one avoids the temporary A'L panel of the
# closed-loop operator by writing the three blocks straight into column views of R.
function DRE.residual(prob::DRE.GALEProblem{<:DRE.LDLᵀ{Float64,DeviceMatrix,Matrix{Float64}}},
                      val::DRE.LDLᵀ{Float64,DeviceMatrix,Matrix{Float64}})
    E, A, C = prob.E, prob.A, prob.C
    iszero(val) && return deepcopy(C)
    alpha, G, S = C
    beta, L, D = val
    n_G, n_0 = size(G, 2), size(L, 2)
    dim = n_G + 2n_0
    R = DeviceMatrix(L.p.ctx, dim)
    R[:, 1:n_G] = G
    spmm!(view_cols(R, n_G+1:n_G+n_0), 'E', L, 1.0, 0.0)
    AtL = view_cols(R, n_G+n_0+1:dim)
    if A isa DRE.LowRankUpdate                        # A'L = A_s'L + inv(α) V'(U'L), LowRankUpdate.jl:51-54,77-86
        As, α, U, V = A
        mul!(AtL, As', L)
        mul!(AtL, parent(V), U' * L, inv(α), true)    # V = K is an Adjoint{DeviceMatrix}; U = B (host)
    else
        mul!(AtL, A', L)
    end
    T = zeros(dim, dim)
    T[1:n_G, 1:n_G] .= alpha .* S
    T[n_G+1:n_G+n_0, n_G+n_0+1:dim] .= beta .* D
    T[n_G+n_0+1:dim, n_G+1:n_G+n_0] .= beta .* D
    DRE.compress!(DRE.lowrank(R, T))
end

# ---- the BlockLinearSolver seam (blocklinear/types.jl:15-30, backslash.jl:8-21, sherman-morrison-woodbury.jl) ----
# ADI(inner_alg = DREB200.Solver()).  prob.A is a PencilOp (plain GALE) or a LowRankUpdate over one (Ros1/Ros2/
# Newton); prob.B the real n x r residual factor.  Real shifts return a DeviceMatrix, complex shifts a ComplexPanel.
struct Solver <: DRE.BlockLinearSolver end
struct ComplexPanel            # V = re + im*i of the complex double step (adi.jl:196-199)
    re::DeviceMatrix
    im::DeviceMatrix
end
Base.real(V::ComplexPanel) = V.re
Base.imag(V::ComplexPanel) = V.im
Base.iszero(::ComplexPanel) = false
mutable struct SolverState     # init/solve!/rhs protocol: rhs(solver) is modified in place by the caller (heuristic.jl:51-60)
    F
    B::DeviceMatrix
end
CommonSolve.init(prob::DRE.BlockLinearProblem, ::Solver) = SolverState(prob.A, prob.B isa DeviceMatrix ? prob.B :
                                                                       DeviceMatrix(context(), reshape(Vector{Float64}(prob.B), :, 1)))
DRE.rhs(s::SolverState) = s.B
split_operator(F::PencilOp) = (F, 1.0, nothing, nothing)
split_operator(F::DRE.LowRankUpdate) = (F.A, Float64(F.α), F.U, F.V)
function CommonSolve.solve!(s::SolverState)
    P, α, U, V = split_operator(s.F)
    ctx = P.ctx
    z = View(-1, 0, 0)
    # System handed over by the ADI step (adi.jl:156,195):  (A_s' + μE' + inv(α) U V) X = B  where, after
    # adjoint(::LowRankUpdate) (LowRankUpdate.jl:51-54), U = K' is an n x m panel and V = B' a host adjoint.
    # Library convention (include/dre_b200.h, dre_set_operator / dre_shift_solve): with operator panels (U_op, Vt_op)
    # it solves (a A + (e + μ) E + inv(α) Vt_op U_op') X = B.  Hence Vt_op = U, U_op = V'.
    a, e = real(P.a), P.e
    Vt_op = U === nothing ? nothing : device_of(ctx, U)
    U_op = V === nothing ? nothing : device_of(ctx, V)
    check(ctx, ccall((:dre_set_operator, LIB), Int32, (Ptr{Cvoid}, Float64, Float64, Float64, View, View),
                     ctx.h, a, 0.0, α, U_op === nothing ? z : view_of(U_op), Vt_op === nothing ? z : view_of(Vt_op)))
    R = s.B
    if iszero(imag(e))
        X = similar(R)
        check(ctx, ccall((:dre_shift_solve, LIB), Int32, (Ptr{Cvoid}, Float64, Float64, View, View, View),
                         ctx.h, real(e), 0.0, view_of(R), view_of(X), z))
        return X
    else
        Xr, Xi = similar(R), similar(R)
        check(ctx, ccall((:dre_shift_solve, LIB), Int32, (Ptr{Cvoid}, Float64, Float64, View, View, View),
                         ctx.h, real(e), imag(e), view_of(R), view_of(Xr), view_of(Xi)))
        return ComplexPanel(Xr, Xi)
    end
end
CommonSolve.solve(prob::DRE.BlockLinearProblem, alg::Solver) = CommonSolve.solve!(CommonSolve.init(prob, alg))

# ---- perform_single_step! / perform_double_step!  (lyapunov/adi.jl:149-225): ONE ccall per ADI step ----
const DeviceADICache = DRE.ADICache{Float64,DeviceMatrix,DeviceMatrix,Matrix{Float64}}
operator_of(A::PencilOp) = DeviceOperator(real(A.a), real(A.e), 1.0, nothing, nothing)
function operator_of(F::DRE.LowRankUpdate)        # F = A_s + inv(α) U V with U = B (host), V = K (lazy adjoint panel)
    ctx = F.A.ctx
    DeviceOperator(real(F.A.a), real(F.A.e), Float64(F.α), device_of(ctx, F.U), parent(F.V))
end
function prefetch!(cache::DeviceADICache)
    # the buffered shifts are known ahead (shifts/helpers.jl:106-113): factor the next ones on side streams
    it = cache.shifts_oracle
    hasproperty(it, :buffer) || return
    ctx = only(cache.residual.Ls).p.ctx
    left = cache.alg.maxiters - length(cache.shifts)
    queued, skip = 0, false
    for μ in Iterators.take(it.buffer, 6)
        (left <= 0 || queued >= 3) && break
        left -= 1
        skip && (skip = false; continue)             # conjugate partner: one factorization serves the pair
        skip = !isreal(μ)
        prefactor!(ctx, Complex(μ)); queued += 1
    end
end
function DRE.perform_single_step!(cache::DeviceADICache, μ)
    alpha, R, T = cache.residual
    ctx = R.p.ctx
    V, _ = adi_step!(ctx, operator_of(cache.prob.A), Complex(μ), R)      # V = (F'+μE')⁻¹R;  R -= 2μ E'V
    prefetch!(cache)
    cache.increment = (-2real(μ) * alpha) * DRE.lowrank(V, T)
    cache.X += cache.increment
    cache.last_compression += 1
    DRE.Shifts.update!(cache.shifts_oracle, cache.X, R, V)
    nothing
end
function DRE.perform_double_step!(cache::DeviceADICache, μ)
    alpha, R, T = cache.residual
    ctx = R.p.ctx
    μ_next = DRE.Shifts.take!(cache.shifts_oracle)
    @assert μ_next ≈ conj(μ)
    push!(cache.shifts, μ_next)
    DRE.Callbacks.observe_gale_metadata!(cache.observer, "ADI shifts", μ_next)
    # V₁ = √2(Re V + δ Im V), V₂ = √(2δ²+2) Im V are formed inside the kernel epilogue; R -= 2√2 Re(μ) E'V₁
    V₁, V₂ = adi_step!(ctx, operator_of(cache.prob.A), Complex(μ), R)
    prefetch!(cache)
    cache.increment = (-2real(μ) * alpha) * (DRE.lowrank(V₁, T) + DRE.lowrank(V₂, T))
    cache.X += cache.increment
    cache.last_compression += 2
    DRE.Shifts.update!(cache.shifts_oracle, cache.X, R, V₁, V₂)
    nothing
end

# ---- Projection shifts: orth / restrict  (Stuff.jl:9-18, util/restrict.jl:5-8, shifts/projection.jl:54-73) ----
# take_many! itself stays the reference's: hcat(Vs...) -> orth -> restrict(E,Q), restrict(A,Q) -> eigvals.
function DRE.Stuff.orth(N::DeviceMatrix)
    ctx = N.p.ctx
    n, k = size(N)
    cap = min(k, n)
    Q0 = DeviceMatrix(ctx, cap)
    Rt = zeros(k, cap)
    rho = Ref{Int32}(0)
    views = [view_of(N)]
    ε = n * eps()
    GC.@preserve views Rt check(ctx, ccall((:dre_rrqr, LIB), Int32,
        (Ptr{Cvoid}, Int32, Ptr{View}, Float64, Float64, View, Ptr{Float64}, Int64, Ref{Int32}),
        ctx.h, 1, views, 1e-15, 1e-3 * ε, view_of(Q0), Rt, k, rho))
    r = Int(rho[])
    r == 0 && return DeviceMatrix(ctx, 0)
    F = svd(Matrix(Rt[:, 1:r]'))                     # N = Q0 (U S W'): the singular values of N (Stuff.jl:14-16)
    ids = findall(s -> abs(s) > ε, F.S)
    view_cols(Q0, 1:r) * F.U[:, ids]                 # n x length(ids) panel
end
DRE.Stuff.restrict(P::PencilOp{<:Real}, Q::DeviceMatrix) = Q' * (P * Q)
# restrict(::LowRankUpdate, Q) (util/restrict.jl:5-8) is generic: Q'U -> (U'Q)' and V*Q are the products above
Base.:*(Qt::Adjoint{Float64,DeviceMatrix}, U::Adjoint{Float64,Matrix{Float64}}) = Qt * Matrix(U)

# ---- Heuristic shifts with device-resident Arnoldi vectors  (shifts/heuristic.jl:39-130) ----
# init(::Heuristic, prob) stays the reference's.  arnoldi_b0 puts the start vector on the device, the operator closures
# (mul!(rhs(solver), A, x); solve!(solver)) go through spmm! / SolverState above, and compute_ritz_values keeps the
# Krylov basis in ONE n x (k+1) panel: each step's twice-repeated MGS sweep and the normalisation are one
# dre_arnoldi_orth call (chain of dot/update launches on the stream, k+2 doubles back); H stays on the host.
DRE.Shifts.arnoldi_b0(E::PencilOp) = DeviceMatrix(E.ctx, reshape(ones(Float64, size(E, 2)), :, 1))
function DRE.Shifts.compute_ritz_values(A, b0::DeviceMatrix, k::Int, desc::String)
    ctx = b0.p.ctx
    H = zeros(k + 1, k)
    V = DeviceMatrix(ctx, k + 1)
    V[:, 1:1] = b0
    lmul_cols!(view_cols(V, 1:1), 1.0 / sqrt(size(b0, 1)))            # (1/norm(b0)) * b0 with b0 = ones(n)
    h = zeros(k + 2)
    for j in 1:k
        w = A(view_cols(V, j:j))                                       # device column -> device column
        GC.@preserve h check(ctx, ccall((:dre_arnoldi_orth, LIB), Int32, (Ptr{Cvoid}, View, View, View, Ptr{Float64}),
                                        ctx.h, view_of(view_cols(V, 1:j)), view_of(w), view_of(view_cols(V, j+1:j+1)), h))
        H[1:j+1, j] .= @view h[1:j+1]
    end
    ritz = eigvals(@view H[1:k, 1:k])
    DRE.Shifts.stabilize_ritz_values!(ritz, desc)
end
lmul_cols!(Y::DeviceMatrix, β::Float64) = (check(Y.p.ctx, ccall((:dre_mat_axpby, LIB), Int32,
    (Ptr{Cvoid}, Float64, View, Float64, View), Y.p.ctx.h, 0.0, View(-1, 0, 0), β, view_of(Y))); Y)

# ---- concatenate! (LDLt.jl:174-191) works through _hcat above; _dcat is generic (host cores) ----
# ---- zero / iszero / rank are generic (LDLt.jl:112-121) given size(::DeviceMatrix) ----
end # module
