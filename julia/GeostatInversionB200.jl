# GeostatInversionB200.jl -- `ccall` shim over libgsi_b200.so (include/gsi_b200.h).
#
# UNTESTED HERE: Julia is not installed in the build image; this file is a mechanical
# transcription of the header (the same table the tested Python `ctypes` host uses,
# geostatinversion.jl_b200/_lib.py).  It keeps the reference's call surface:
#   RandMatFact.randsvd(A, K, p, q), rangefinder(A, l, q), getxis(Q, numxis, p, q, seed)
# for A::Matrix{Float64}, A::KernelCovMatrix (new) and A::LowRankCovMatrix, and
#   pcgalsqr / pcgadirect / pcga / rga with the reference's positional and keyword arguments.
module GeostatInversionB200

import Random
import LinearAlgebra

const LIB = get(ENV, "GSI_B200_LIB", joinpath(@__DIR__, "..", "geostatinversion.jl_b200", "lib", "libgsi_b200.so"))

const GSI_LAYOUT_TALL = Int32(0)
const GSI_LAYOUT_COLMAJOR = Int32(1)
const GSI_NORMALISER_LU_REF = Int32(0)

lasterror() = unsafe_string(ccall((:gsi_last_error_string, LIB), Cstring, ()))

function check(status::Int32)
	status == 0 && return nothing
	msg = lasterror()
	status == 2 && throw(DimensionMismatch(msg))
	status == 3 && throw(LinearAlgebra.SingularException(0))      # lu(...; check=true), RandMatFact.jl:60,68,72
	status == 4 && throw(LinearAlgebra.PosDefException(0))        # eig_nystrom, RandMatFact.jl:95
	error(msg)                                                     # incl. "parameter numiterations should be positive, ..."
end

mutable struct Context
	h::Ptr{Cvoid}
	function Context(device::Integer=0; rank::Integer=0, world::Integer=1, uid::Vector{UInt8}=UInt8[])
		out = Ref{Ptr{Cvoid}}(C_NULL)
		check(ccall((:gsi_ctx_create, LIB), Int32, (Int32, Int32, Int32, Ptr{UInt8}, Ref{Ptr{Cvoid}}),
			device, rank, world, world > 1 ? uid : C_NULL, out))
		ctx = new(out[])
		finalizer(c->ccall((:gsi_ctx_destroy, LIB), Int32, (Ptr{Cvoid},), c.h), ctx)
		return ctx
	end
end

# tuning knobs of include/gsi_b200.h (gsi_ctx_set_option), e.g. setoption!(ctx, "kcov.window", 8)
setoption!(ctx::Context, name::AbstractString, value::Integer) =
	check(ccall((:gsi_ctx_set_option, LIB), Int32, (Ptr{Cvoid}, Cstring, Int64), ctx.h, name, value))
function getoption(ctx::Context, name::AbstractString)
	v = Ref{Int64}(0)
	check(ccall((:gsi_ctx_get_option, LIB), Int32, (Ptr{Cvoid}, Cstring, Ref{Int64}), ctx.h, name, v))
	return v[]
end

const defaultctx = Ref{Union{Nothing, Context}}(nothing)
context() = (defaultctx[] === nothing && (defaultctx[] = Context()); defaultctx[])

mutable struct DeviceMatrix
	h::Ptr{Cvoid}
	rows::Int
	cols::Int
	function DeviceMatrix(ctx::Context, layout::Int32, rows::Integer, cols::Integer)
		out = Ref{Ptr{Cvoid}}(C_NULL)
		check(ccall((:gsi_buf_alloc, LIB), Int32, (Ptr{Cvoid}, Int32, Int64, Int64, Ref{Ptr{Cvoid}}), ctx.h, layout, rows, cols, out))
		b = new(out[], rows, cols)
		finalizer(x->ccall((:gsi_buf_free, LIB), Int32, (Ptr{Cvoid},), x.h), b)    # idempotent, never throws
		return b
	end
end

function upload!(b::DeviceMatrix, A::Matrix{Float64})
	GC.@preserve A check(ccall((:gsi_buf_upload, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64), b.h, A, max(1, size(A, 1))))
	return b
end

function download(b::DeviceMatrix)
	A = Matrix{Float64}(undef, b.rows, b.cols)
	GC.@preserve A check(ccall((:gsi_buf_download, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64), b.h, A, max(1, b.rows)))
	return A
end

abstract type Operator end
mutable struct DenseOperator <: Operator; h::Ptr{Cvoid}; buf::DeviceMatrix; ctx::Context; end
mutable struct KernelCovMatrix <: Operator; h::Ptr{Cvoid}; n::Int; ctx::Context; end
mutable struct LowRankCovMatrix <: Operator; h::Ptr{Cvoid}; buf::DeviceMatrix; ctx::Context; end

function DenseOperator(A::Matrix{Float64}; ctx::Context=context())
	buf = upload!(DeviceMatrix(ctx, GSI_LAYOUT_COLMAJOR, size(A)...), A)
	out = Ref{Ptr{Cvoid}}(C_NULL)
	check(ccall((:gsi_op_dense, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Ref{Ptr{Cvoid}}), ctx.h, buf.h, 0, size(A, 1), out))
	op = DenseOperator(out[], buf, ctx)
	finalizer(o->ccall((:gsi_op_free, LIB), Int32, (Ptr{Cvoid},), o.h), op)
	return op
end

"kind: 0 exponential, 1 gaussian, 2 powerlaw; coords d x n; ell d"
function KernelCovMatrix(kind::Integer, coords::Matrix{Float64}, ell::Vector{Float64}; sigma2=1.0, nugget=0.0, beta=1.0, ctx::Context=context())
	d, n = size(coords)
	out = Ref{Ptr{Cvoid}}(C_NULL)
	GC.@preserve coords ell check(ccall((:gsi_op_kernelcov, LIB), Int32,
		(Ptr{Cvoid}, Int32, Int32, Int64, Ptr{Float64}, Ptr{Float64}, Float64, Float64, Float64, Int64, Int64, Ref{Ptr{Cvoid}}),
		ctx.h, kind, d, n, coords, ell, sigma2, nugget, beta, 0, n, out))
	op = KernelCovMatrix(out[], n, ctx)
	finalizer(o->ccall((:gsi_op_free, LIB), Int32, (Ptr{Cvoid},), o.h), op)
	return op
end

"structured grid variant: dims[1] fastest (Julia linear index), coordinates idx .* spacing"
function KernelCovGrid(kind::Integer, dims::Vector{Int}, spacing::Vector{Float64}, ell::Vector{Float64}; sigma2=1.0, nugget=0.0, beta=1.0, ctx::Context=context())
	d = length(dims)
	n = prod(dims)
	out = Ref{Ptr{Cvoid}}(C_NULL)
	dims64 = Int64.(dims)
	GC.@preserve dims64 spacing ell check(ccall((:gsi_op_kernelcov_grid, LIB), Int32,
		(Ptr{Cvoid}, Int32, Int32, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Float64, Float64, Float64, Int64, Int64, Ref{Ptr{Cvoid}}),
		ctx.h, kind, d, dims64, spacing, ell, sigma2, nugget, beta, 0, n, out))
	op = KernelCovMatrix(out[], n, ctx)
	finalizer(o->ccall((:gsi_op_free, LIB), Int32, (Ptr{Cvoid},), o.h), op)
	return op
end

function LowRankCovMatrix(samples::Vector{Vector{Float64}}; ctx::Context=context())
	S = reduce(hcat, samples)
	buf = upload!(DeviceMatrix(ctx, GSI_LAYOUT_COLMAJOR, size(S)...), S)
	out = Ref{Ptr{Cvoid}}(C_NULL)
	check(ccall((:gsi_op_lowrankcov, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ref{Ptr{Cvoid}}), ctx.h, buf.h, 1, out))
	op = LowRankCovMatrix(out[], buf, ctx)
	finalizer(o->ccall((:gsi_op_free, LIB), Int32, (Ptr{Cvoid},), o.h), op)
	return op
end

function Base.size(op::Operator)
	m = Ref{Int64}(0); n = Ref{Int64}(0)
	check(ccall((:gsi_op_size, LIB), Int32, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}), op.h, m, n))
	return (Int(m[]), Int(n[]))
end
Base.size(op::Operator, i::Int) = (i == 1 || i == 2) ? size(op)[i] : error("there is no $i-th dimension in a $(typeof(op))")
Base.adjoint(op::Union{KernelCovMatrix, LowRankCovMatrix}) = op      # symmetric (src/lowrank.jl:38-44)

function Base.:*(op::Operator, X::Matrix{Float64})
	ctx = op.ctx
	Xd = upload!(DeviceMatrix(ctx, GSI_LAYOUT_TALL, size(X)...), X)
	Yd = DeviceMatrix(ctx, GSI_LAYOUT_TALL, size(op, 1), size(X, 2))
	check(ccall((:gsi_op_apply, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Ptr{Cvoid}), op.h, 0, Xd.h, Yd.h))
	return download(Yd)
end

module RandMatFact
import ..GeostatInversionB200: LIB, check, context, Operator, DenseOperator, DeviceMatrix, upload!, download,
	GSI_LAYOUT_TALL, GSI_NORMALISER_LU_REF

asoperator(A::Operator) = A
asoperator(A::Matrix{Float64}) = DenseOperator(A)

"randsvd(A, K, p, q) -- reference src/RandMatFact.jl:83-90 (Omega drawn on the host, :54)"
function randsvd(A, K::Int, p::Int, q::Int)
	q < 0 && error("parameter numiterations should be positive, but numiterations=$q")
	op = asoperator(A)
	n = size(op, 2)
	Omega = randn(n, K + p)                              # same host RNG stream as the reference
	Om = upload!(DeviceMatrix(op.ctx, GSI_LAYOUT_TALL, n, K + p), Omega)
	Z = DeviceMatrix(op.ctx, GSI_LAYOUT_TALL, n, K + p)
	check(ccall((:gsi_randsvd, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Int32, Ptr{Cvoid}, Ptr{Float64}),
		op.h, Om.h, K, p, q, GSI_NORMALISER_LU_REF, Z.h, C_NULL))
	return download(Z)
end

"rangefinder(A, l, numiterations) -- reference src/RandMatFact.jl:50-80"
function rangefinder(A, l::Int64, numiterations::Int64)
	numiterations < 0 && error("parameter numiterations should be positive, but numiterations=$numiterations")
	op = asoperator(A)
	n = size(op, 2)
	Om = upload!(DeviceMatrix(op.ctx, GSI_LAYOUT_TALL, n, l), randn(n, l))
	Q = DeviceMatrix(op.ctx, GSI_LAYOUT_TALL, size(op, 1), l)
	check(ccall((:gsi_rangefinder_fixed, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int32, Ptr{Cvoid}), op.h, Om.h, numiterations, GSI_NORMALISER_LU_REF, Q.h))
	return download(Q)
end
end # RandMatFact

function randsvdwithseed(Q, numxis, p, q, seed::Nothing)
	return RandMatFact.randsvd(Q, numxis, p, q)
end
function randsvdwithseed(Q, numxis, p, q, seed::Int)
	Random.seed!(seed)                                   # src/GeostatInversion.jl:24-27
	return RandMatFact.randsvd(Q, numxis, p, q)
end

"getxis(Q, numxis, p, q=3, seed=nothing) -- src/GeostatInversion.jl:63-70; Q::Matrix or any Operator"
function getxis(Q, numxis::Int, p::Int, q::Int=3, seed=nothing)
	Z = randsvdwithseed(Q, numxis, p, q, seed)
	return [Z[:, i] for i = 1:numxis]
end

# ---------------------------------------------------------------------------------------------
# PCGA / RGA drivers behind the reference's keyword surface (src/lsqr.jl:20-63,
# src/direct.jl:21-67, src/GeostatInversion.jl:101-105).  Transcription of the tested Python
# host geostatinversion.jl_b200/pcga.py.  The user's forward model stays a Julia function run
# under `pmap`, exactly as in the reference; the batch of K+3 parameter vectors, the
# saddle-point solve and the update run on the device.
import Distributed
import SparseArrays

splitR(R::Number, nobs) = (fill(Float64(R), nobs), nothing)
splitR(R::AbstractVector, nobs) = (Vector{Float64}(R), nothing)
splitR(R::LinearAlgebra.Diagonal, nobs) = (Vector{Float64}(R.diag), nothing)
function splitR(R::SparseArrays.SparseMatrixCSC, nobs)
	d = Vector{Float64}(LinearAlgebra.diag(R))
	return SparseArrays.nnz(SparseArrays.dropzeros(R - SparseArrays.spdiagm(0=>d))) == 0 ? (d, nothing) : (nothing, Matrix{Float64}(R))
end
splitR(R::AbstractMatrix, nobs) = (nothing, Matrix{Float64}(R))

function xistodevice(ctx::Context, xis::Vector{Vector{Float64}})
	Zk = reduce(hcat, xis)
	return upload!(DeviceMatrix(ctx, GSI_LAYOUT_TALL, size(Zk)...), Zk)
end

"`paramstorun` batch (src/lsqr.jl:37-43 == src/direct.jl:39-45) as the columns of an n x (K+3) matrix"
function paramstorun(ctx::Context, Zk::DeviceMatrix, K::Int, s::Vector{Float64}, X::Vector{Float64}, delta::Float64)
	P = DeviceMatrix(ctx, GSI_LAYOUT_TALL, Zk.rows, K + 3)
	GC.@preserve s X check(ccall((:gsi_pcga_paramstorun, LIB), Int32,
		(Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Float64, Ptr{Cvoid}), ctx.h, Zk.h, K, s, X, delta, P.h))
	return download(P)
end

# x = lsqr(PCGALowRankMatrix(etas, HX, R), b) (src/lsqr.jl:53-54) or pinv(bigA) * b (src/direct.jl:49-58)
function saddlesolve(ctx::Context, E::Matrix{Float64}, HX::Vector{Float64}, R, b::Vector{Float64}, direct::Bool)
	nobs, K = size(E)
	rd, rD = splitR(R, nobs)
	rdp = rd === nothing ? Ptr{Float64}(C_NULL) : pointer(rd)
	rDp = rD === nothing ? Ptr{Float64}(C_NULL) : pointer(rD)
	x = Vector{Float64}(undef, nobs + 1)
	GC.@preserve E HX rd rD b x begin
		if direct
			check(ccall((:gsi_pcga_direct_solve, LIB), Int32,
				(Ptr{Cvoid}, Int64, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}),
				ctx.h, nobs, K, E, nobs, HX, rdp, rDp, nobs, b, x, C_NULL))
		else
			check(ccall((:gsi_pcga_lsqr_solve, LIB), Int32,
				(Ptr{Cvoid}, Int64, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Float64, Float64, Float64, Int64, Ptr{Float64}, Ptr{Int64}, Ptr{Int32}),
				ctx.h, nobs, K, E, nobs, HX, rdp, rDp, nobs, b, 0.0, 0.0, 0.0, 0, x, C_NULL, C_NULL))
		end
	end
	return x
end

function pcgaiteration(forwardmodel::Function, s::Vector{Float64}, X::Vector{Float64}, Zk::DeviceMatrix, K::Int, R, y::Vector{Float64},
		delta::Float64, callback::Function, direct::Bool, ctx::Context)
	P = paramstorun(ctx, Zk, K, s, X, delta)
	results = Distributed.pmap(forwardmodel, [P[:, i] for i = 1:K + 3])          # src/lsqr.jl:44
	hs = results[K + 3]
	callback(s, hs)                                                             # src/direct.jl:47
	E = Matrix{Float64}(undef, length(hs), K)
	for i = 1:K
		E[:, i] = (results[i] - hs) / delta                                     # etas, :46-49
	end
	HX = (results[K + 1] - hs) / delta                                          # :50
	Hs = (results[K + 2] - hs) / delta                                          # :51
	b = [y - hs + Hs; 0.0]                                                      # :52
	x = saddlesolve(ctx, E, HX, R, b, direct)
	snew = Vector{Float64}(undef, Zk.rows)
	GC.@preserve X E x snew check(ccall((:gsi_pcga_update, LIB), Int32,
		(Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Int64, Int64, Ptr{Float64}, Ptr{Float64}),
		ctx.h, Zk.h, K, X, E, size(E, 1), size(E, 1), x, snew))                 # s = X*beta + sum xi_i (eta_i . xi_bar), :55-61
	return snew
end

function pcgaouter(forwardmodel, s0, X, xis, R, y, maxiters, delta, xtol, callback, direct, ctx)
	Zk = xistodevice(ctx, xis)
	K = length(xis)
	converged = false
	s = Vector{Float64}(s0)
	itercount = 0
	while !converged && itercount < maxiters                                    # src/lsqr.jl:24-31
		olds = s
		s = pcgaiteration(forwardmodel, s, Vector{Float64}(X), Zk, K, R, Vector{Float64}(y), Float64(delta), callback, direct, ctx)
		if LinearAlgebra.norm(s - olds) < xtol
			converged = true
		end
		itercount += 1
	end
	return s
end

"pcgalsqr(forwardmodel, s0, X, xis, R, y; maxiters=5, delta=sqrt(eps(Float64)), xtol=1e-6) -- src/lsqr.jl:20-33 (+ callback, SURVEY F5)"
pcgalsqr(forwardmodel::Function, s0::Vector, X::Vector, xis::Array{Array{Float64, 1}, 1}, R, y::Vector;
		maxiters::Int=5, delta::Float64=sqrt(eps(Float64)), xtol::Float64=1e-6, callback=(s, obs_cal)->nothing, ctx::Context=context()) =
	pcgaouter(forwardmodel, s0, X, xis, R, y, maxiters, delta, xtol, callback, false, ctx)

"pcgadirect(forwardmodel, s0, X, xis, R, y; maxiters=5, delta=sqrt(eps(Float64)), xtol=1e-6, callback=(s, obs_cal)->nothing) -- src/direct.jl:21-35"
pcgadirect(forwardmodel::Function, s0::Vector, X::Vector, xis::Array{Array{Float64, 1}, 1}, R, y::Vector;
		maxiters::Int=5, delta::Float64=sqrt(eps(Float64)), xtol::Float64=1e-6, callback=(s, obs_cal)->nothing, ctx::Context=context()) =
	pcgaouter(forwardmodel, s0, X, xis, R, y, maxiters, delta, xtol, callback, true, ctx)

const pcga = pcgadirect                                                         # src/GeostatInversion.jl:105

"rga(forwardmodel, s0, X, xis, R, y, S; maxiters, delta, xtol, pcgafunc=pcgadirect, callback) -- src/GeostatInversion.jl:101-103"
function rga(forwardmodel::Function, s0::Vector, X::Vector, xis::Array{Array{Float64, 1}, 1}, R, y::Vector, S::Matrix{Float64};
		maxiters::Int=5, delta::Float64=sqrt(eps(Float64)), xtol::Float64=1e-6, pcgafunc=pcgadirect, callback=(s, obs_cal)->nothing,
		ctx::Context=context())
	Nred, nobs = size(S)
	Sd = upload!(DeviceMatrix(ctx, GSI_LAYOUT_COLMAJOR, Nred, nobs), S)
	function sketch(V::Matrix{Float64})                                         # S * V on the tensor-core GEMM
		Vd = upload!(DeviceMatrix(ctx, GSI_LAYOUT_TALL, size(V)...), V)
		out = DeviceMatrix(ctx, GSI_LAYOUT_TALL, Nred, size(V, 2))
		check(ccall((:gsi_sketch_apply, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), ctx.h, Sd.h, Vd.h, out.h))
		return download(out)
	end
	rd, rD = splitR(R, nobs)
	if rd === nothing
		SRS = sketch(rD * S')                                                   # dense R: S * (R * S')
	else
		SRS = Matrix{Float64}(undef, Nred, Nred)
		GC.@preserve rd SRS check(ccall((:gsi_sketch_cov, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64),
			ctx.h, Sd.h, rd, SRS, Nred))
	end
	return pcgafunc(x->vec(sketch(reshape(forwardmodel(x), :, 1))), s0, X, xis, SRS, vec(sketch(reshape(Vector{Float64}(y), :, 1)));
		maxiters=maxiters, delta=delta, xtol=xtol, callback=callback, ctx=ctx)
end

end
