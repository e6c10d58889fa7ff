#
# HyperelasticB200.jl -- thin `ccall` shim over libhyperelastic_b200.so.
#
# Drop-in for the hot path of BlackSiberian/HyperelasticSolver: it exports the SAME names as the
# reference modules it replaces
#     HyperelasticityMPh: prim2cons_mph, cons2prim_mph, flux_mph, noncons_flux, get_eigvals   (HyperelasticityMPh.jl:13)
#     NumFluxes:          lxf, hll                                                             (NumFluxes.jl:15)
#     EquationsOfState:   Barton2009, Hank2016, energy, pressure, stress (Hank2016 methods)        (EquationsOfState.jl:71,305,366)
# with the same argument meaning (`eos::Tuple{Barton2009,Barton2009}`, `Vector{Float64}` states),
# plus batched methods on `Matrix{Float64}(nvar, n)` and a device-resident `Solver` that replaces
# the two `Threads.@threads` loops of main.jl:204-227 with one call per step.
#
# NOTE: Julia is not installed in the build image, so this file could not be executed there.  It
# is deliberately trivial: every function is one `ccall` on the C ABI that the Python ctypes layer
# (hyperelasticsolver_b200/_lib.py) binds and the GPU tests exercise.  No Julia GPU packages.
#
module HyperelasticB200

export Barton2009, Hank2016, energy, pressure, stress, prim2cons_mph, cons2prim_mph, flux_mph, noncons_flux, get_eigvals, lxf, hll,
       Solver, upload!, download!, step!, advance!, wave_speeds, destroy!

const LIB = get(ENV, "HYPERELASTIC_B200_LIB", joinpath(@__DIR__, "..", "hyperelasticsolver_b200", "libhyperelastic_b200.so"))

const HS_OK, HS_ERR_ARG, HS_ERR_CUDA, HS_ERR_DOMAIN = 0, 1, 2, 3
const HS_MODEL_SP13, HS_MODEL_MPH30 = 0, 1
const HS_FLUX_LXF, HS_FLUX_HLL = 0, 1

# EquationsOfState.jl:71-116 -- same fields, same keyword constructor, but concretely typed so
# that a Tuple of them is a contiguous C array of hs_barton2009_t.
struct Barton2009
  rho0::Float64; c0::Float64; cv::Float64; t0::Float64; b0::Float64
  alpha::Float64; beta::Float64; gamma::Float64
  b0sq::Float64; k0::Float64
  function Barton2009(; _rho0=8.93, _c0=4.6, _cv=3.9e-4, _t0=300, _b0=2.1, _alpha=1, _beta=3, _gamma=2)
    new(_rho0, _c0, _cv, _t0, _b0, _alpha, _beta, _gamma, _b0^2, _c0^2 - (4 / 3) * _b0^2)
  end
end

# EquationsOfState.jl:305-319 -- same fields and defaults, concretely typed (== hs_hank2016_t)
struct Hank2016
  rho0::Float64; mu::Float64; gamma::Float64; pres_inf::Float64; a::Float64
  Hank2016(rho0=2.7, mu=26e9, gamma=3.4, pres_inf=21.5e9, a=0.5) = new(rho0, mu, gamma, pres_inf, a)
end

function check(rc::Cint)
  rc == HS_OK && return nothing
  msg = unsafe_string(ccall((:hs_last_error, LIB), Cstring, ()))
  rc == HS_ERR_DOMAIN && throw(DomainError(msg))   # where the reference throws: HyperelasticityMPh.jl:114
  error("hyperelastic_b200 error $rc: $msg")
end

eosvec(eos::Tuple) = collect(eos)::Vector{Barton2009}
ncols(Q::AbstractVecOrMat) = Q isa AbstractVector ? 1 : size(Q, 2)

# --- per-cell functions (HyperelasticityMPh.jl) ---------------------------------------------------
for (jl, c) in ((:prim2cons_mph, :hs_prim2cons), (:cons2prim_mph, :hs_cons2prim), (:flux_mph, :hs_flux))
  @eval function $jl(eos::Tuple{Barton2009,Barton2009}, X::VecOrMat{Float64}; device::Integer=0)
    Y = similar(X); e = eosvec(eos)
    GC.@preserve X Y e check(ccall(($(QuoteNode(c)), LIB), Cint,
        (Cint, Ptr{Barton2009}, Cint, Ptr{Float64}, Ptr{Float64}, Int64, Cint),
        HS_MODEL_MPH30, e, 2, X, Y, ncols(X), device))
    return Y
  end
end

# noncons_flux returns the dense 30x30 matrix like the reference (HyperelasticityMPh.jl:178-250)
function noncons_flux(eos::Tuple{Barton2009,Barton2009}, Q::Vector{Float64}; device::Integer=0)
  col = similar(Q); B = zeros(30, 30); e = eosvec(eos)
  GC.@preserve Q col B e check(ccall((:hs_noncons_flux, LIB), Cint,
      (Ptr{Barton2009}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Cint), e, Q, col, B, 1, device))
  return B
end

function get_eigvals(eos::Tuple{Barton2009,Barton2009}, Q::VecOrMat{Float64}, n::Array{<:Any,1}; device::Integer=0)
  nn = Vector{Float64}(n)   # unit normal; main.jl:208 passes [1, 0, 0]
  E = Q isa AbstractVector ? zeros(12) : zeros(12, size(Q, 2)); e = eosvec(eos)
  GC.@preserve Q E e nn check(ccall((:hs_get_eigvals, LIB), Cint,
      (Cint, Ptr{Barton2009}, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Cint), HS_MODEL_MPH30, e, 2, Q, nn, E, ncols(Q), device))
  return E
end

# --- numerical fluxes (NumFluxes.jl) ---------------------------------------------------------------
# hll(eos, Q_l, Q_r, eigvals) -> (zeros(30), D-, D+);  eigvals = [eig_l, eig_r] as in main.jl:56-57
function hll(eos::Tuple{Barton2009,Barton2009}, Q_l::VecOrMat{Float64}, Q_r::VecOrMat{Float64}, eigvals; device::Integer=0)
  cons = similar(Q_l); dm = similar(Q_l); dp = similar(Q_l); e = eosvec(eos)
  el = Q_l isa AbstractVector ? Vector{Float64}(eigvals[1]) : Matrix{Float64}(eigvals[1])
  er = Q_l isa AbstractVector ? Vector{Float64}(eigvals[2]) : Matrix{Float64}(eigvals[2])
  GC.@preserve Q_l Q_r el er cons dm dp e check(ccall((:hs_hll, LIB), Cint,
      (Cint, Ptr{Barton2009}, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Cint),
      HS_MODEL_MPH30, e, 2, Q_l, Q_r, el, er, cons, dm, dp, C_NULL, ncols(Q_l), device))
  return cons, dm, dp
end

function lxf(eos::Tuple{Barton2009,Barton2009}, Q_l::VecOrMat{Float64}, Q_r::VecOrMat{Float64}, lambda; device::Integer=0)
  cons = similar(Q_l); dm = similar(Q_l); dp = similar(Q_l); e = eosvec(eos)
  GC.@preserve Q_l Q_r cons dm dp e check(ccall((:hs_lxf, LIB), Cint,
      (Cint, Ptr{Barton2009}, Cint, Ptr{Float64}, Ptr{Float64}, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Cint),
      HS_MODEL_MPH30, e, 2, Q_l, Q_r, Float64(lambda), cons, dm, dp, ncols(Q_l), device))
  return cons, dm, dp
end

# --- device-resident time loop (main.jl:202-227) ---------------------------------------------------
mutable struct Solver
  ctx::Ptr{Cvoid}
  nvar::Int; ncells::Int; nprob::Int
end

# devices = [0, 1, ..., 7]: several GPUs of this process (hs_create_multi): one grid is slab-decomposed,
# an ensemble (nprob > 1) is shared out by problems
function Solver(eos::Tuple{Barton2009,Barton2009}, ncells::Integer; nprob::Integer=1, device::Integer=0, devices::Vector{<:Integer}=Int[])
  ref = Ref{Ptr{Cvoid}}(C_NULL); e = eosvec(eos)
  if length(devices) > 1
    d = Vector{Cint}(devices)
    GC.@preserve e d check(ccall((:hs_create_multi, LIB), Cint, (Ref{Ptr{Cvoid}}, Cint, Ptr{Barton2009}, Cint, Int64, Int64, Ptr{Cint}, Cint),
        ref, HS_MODEL_MPH30, e, 2, ncells, nprob, d, length(d)))
  else
    GC.@preserve e check(ccall((:hs_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Cint, Ptr{Barton2009}, Cint, Int64, Int64, Cint),
        ref, HS_MODEL_MPH30, e, 2, ncells, nprob, device))
  end
  s = Solver(ref[], 30, ncells, nprob)
  finalizer(destroy!, s)
  return s
end

destroy!(s::Solver) = (s.ctx != C_NULL && ccall((:hs_destroy, LIB), Cint, (Ptr{Cvoid},), s.ctx); s.ctx = C_NULL; nothing)
upload!(s::Solver, Q0::Array{Float64}) = GC.@preserve Q0 check(ccall((:hs_upload, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), s.ctx, Q0))
download!(s::Solver, Q::Array{Float64}) = (GC.@preserve Q check(ccall((:hs_download, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), s.ctx, Q)); Q)

# lambda_max of the CFL sweep, main.jl:204-212
function wave_speeds(s::Solver)
  lam = zeros(s.nprob)
  GC.@preserve lam check(ccall((:hs_wave_speeds, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), s.ctx, C_NULL, lam))
  return lam
end

# one pass of main.jl:204-227; returns dt
function step!(s::Solver, flux::Function, cfl, dx)
  dt = zeros(s.nprob)
  GC.@preserve dt check(ccall((:hs_step, LIB), Cint, (Ptr{Cvoid}, Cint, Float64, Float64, Ptr{Float64}),
      s.ctx, flux === hll ? HS_FLUX_HLL : HS_FLUX_LXF, cfl, dx, dt))
  return s.nprob == 1 ? dt[1] : dt
end

# `while t < T` without returning to the host between steps; returns (t, step_num)
function advance!(s::Solver, flux::Function, cfl, dx, T; t=0.0, step_num=0, max_steps=typemax(Int32))
  tv = fill(Float64(t), s.nprob); sv = fill(Int64(step_num), s.nprob)
  GC.@preserve tv sv check(ccall((:hs_advance, LIB), Cint,
      (Ptr{Cvoid}, Cint, Float64, Float64, Float64, Int64, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}),
      s.ctx, flux === hll ? HS_FLUX_HLL : HS_FLUX_LXF, cfl, dx, T, max_steps, tv, sv, C_NULL))
  return s.nprob == 1 ? (tv[1], sv[1]) : (tv, sv)
end

# --- Hank2016 (EquationsOfState.jl:317-356): scalar methods like the reference, batched over columns ------------
# A 3x3 tensor is a 3x3 Matrix or its 9 column-major entries (the reference's own methods mix the two and
# cannot run as written); batches are (9, n) / (3, n) matrices with den, pres, e_int as length-n vectors.
_vecf(x) = x isa Real ? Float64[x] : Vector{Float64}(x)
function energy(eos::Hank2016, den, pres, G::Array{Float64}; device::Integer=0)
  d = _vecf(den); p = _vecf(pres); n = length(d); out = Vector{Float64}(undef, n); e = Ref(eos)
  GC.@preserve d p G out e check(ccall((:hs_hank2016_energy, LIB), Cint,
      (Ptr{Hank2016}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Cint), e, d, p, G, out, n, device))
  return den isa Real ? out[1] : out
end
function pressure(eos::Hank2016, den, e_int, i::Array{Float64}; device::Integer=0)
  d = _vecf(den); ei = _vecf(e_int); n = length(d); out = Vector{Float64}(undef, n); e = Ref(eos)
  GC.@preserve d ei i out e check(ccall((:hs_hank2016_pressure, LIB), Cint,
      (Ptr{Hank2016}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Cint), e, d, ei, i, out, n, device))
  return den isa Real ? out[1] : out
end
function stress(eos::Hank2016, den, pressure, distortion::Array{Float64}; device::Integer=0)
  d = _vecf(den); p = _vecf(pressure); n = length(d); out = similar(distortion); e = Ref(eos)
  GC.@preserve d p distortion out e check(ccall((:hs_hank2016_stress, LIB), Cint,
      (Ptr{Hank2016}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Cint), e, d, p, distortion, out, n, device))
  return out
end

end # module HyperelasticB200
