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
       initial_states, initial_condition, initial_condition_tanh, update_cell,
       Solver, upload!, download!, step!, step_host!, advance!, set_time!, wave_speeds, destroy!, host_register!, host_unregister!,
       SinglePhase, Solver2D

const LIB = get(ENV, "HYPERELASTIC_B200_LIB", joinpath(@__DIR__, "..", "hyperelasticsolver_b200", "libhyperelastic_b200.so"))

const HS_OK, HS_ERR_ARG, HS_ERR_CUDA, HS_ERR_DOMAIN, HS_ERR_EXCHANGE = 0, 1, 2, 3, 4
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

# --- host-side setup of main.jl that does not touch physics -------------------------------------------
# initial_states(eos, testcase) -> (Ql, Qr): the tables of HyperelasticityMPh.jl:275-426, primitives assembled as at
# :412-420 (nominal density den/det(F), F splatted column-major) and converted with prim2cons_mph (:422-423).
const _I3 = [1.0 0 0; 0 1 0; 0 0 1]
const _F3 = [1.0 0 0; -0.01 0.95 0.02; -0.015 0 0.9]
const _F4L = [0.98 0 0; 0.02 1 0.1; 0 0 1]
const _F4R = [1.0 0 0; 0 1 0.1; 0 0 1]
const _F5R = [1.0 0 0; 0.015 0.95 0; -0.01 0 0.9]
# testcase => (alpha_l_1, alpha_l_2, alpha_r_1, alpha_r_2, den, u_l, S_l, F_l, u_r, S_r, F_r); the alpha values are the
# reference's own literals (0.1 is a literal, not 1 - 0.9: HyperelasticityMPh.jl:349-352)
const _MPH_CASES = Dict(
  1 => (0.5, 0.5, 0.5, 0.5, 5.0, [0.0, 0, 0], 0.0, _I3, [0.0, 0, 0], 0.0, _I3),
  2 => (0.5, 0.5, 0.5, 0.5, 5.0, [1.0, 0, 0], 0.0, _I3, [1.0, 0, 0], 0.0, _I3),
  3 => (0.5, 0.5, 0.5, 0.5, 8.9, [2.0, 0.0, 0.1], 0.0, _F3, [2.0, 0.0, 0.1], 0.0, _F3),
  4 => (0.5, 0.5, 0.5, 0.5, 8.9, [0.0, 0.5, 1.0], 1e-3, _F4L, [0.0, 0.0, 0.0], 0.0, _F4R),
  5 => (0.5, 0.5, 0.5, 0.5, 8.9, [2.0, 0.0, 0.1], 0.0, _F3, [0.0, -0.03, -0.01], 0.0, _F5R),
  6 => (0.1, 0.9, 0.9, 0.1, 8.9, [0.0, 0.5, 1.0], 1.0e-3, _F4L, [0.0, 0.0, 0.0], 0.0, _F4R),
  7 => (0.1, 0.9, 0.9, 0.1, 8.9, [2.0, 0.0, 0.1], 0.0, _F3, [0.0, -0.03, -0.01], 0.0, _F5R),
  10 => (0.4, 0.6, 0.6, 0.4, 8.9, [2.0, 0.0, 0.1], 0.0, _F3, [2.0, 0.0, 0.1], 0.0, _F3))
_det3(F) = F[1,1]*(F[2,2]*F[3,3]-F[2,3]*F[3,2]) - F[1,2]*(F[2,1]*F[3,3]-F[2,3]*F[3,1]) + F[1,3]*(F[2,1]*F[3,2]-F[2,2]*F[3,1])
function initial_states(eos::Tuple{Barton2009,Barton2009}, testcase::Integer; device::Integer=0)
  haskey(_MPH_CASES, testcase) || error("unknown multiphase test case $testcase")
  (al1, al2, ar1, ar2, den, ul, Sl, Fl, ur, Sr, Fr) = _MPH_CASES[testcase]
  dl = den / _det3(Fl); dr = den / _det3(Fr)
  Pl = Float64[al1, dl, ul..., Sl, Fl..., al2, dl, ul..., Sl, Fl...]     # F... splats column-major (:417-420)
  Pr = Float64[ar1, dr, ur..., Sr, Fr..., ar2, dr, ur..., Sr, Fr...]
  return prim2cons_mph(eos, Pl; device=device), prim2cons_mph(eos, Pr; device=device)
end

# main.jl:99-106
function initial_condition(Ql::Vector{Float64}, Qr::Vector{Float64}, nx::Integer)
  Q0 = Array{Float64}(undef, length(Ql), nx)
  for i in 1:nx
    Q0[:, i] = (i - 1) < nx / 2 ? Ql : Qr
  end
  return Q0
end

# main.jl:110-123: volume fraction smoothed with tanh across `width` around x0; the other primitives of the left state
function initial_condition_tanh(eos::Tuple{Barton2009,Barton2009}, Ql::Vector{Float64}, Qr::Vector{Float64}, nx::Integer; x0=0.5, width=0.05, device::Integer=0)
  Pl = cons2prim_mph(eos, Ql; device=device); Pr = cons2prim_mph(eos, Qr; device=device)
  P = Array{Float64}(undef, 30, nx)
  for i in 1:nx
    x = (i - 0.5) / nx
    w = 0.5 * (1 + tanh((x - x0) / width))
    P[:, i] = Pl
    P[1, i] = (1 - w) * Pl[1] + w * Pr[1]
    P[16, i] = 1 - P[1, i]
  end
  return prim2cons_mph(eos, P; device=device)
end

# update_cell, main.jl:30-41 (LxF) and :43-60 (HLL), on one 3-cell stencil Q (nvar x 3), as the reference defines them
function update_cell(Q::Array{Float64,2}, flux_num::Function, lambda, eos::Tuple{Barton2009,Barton2009})
  Q_l, Qc, Q_r = Q[:, 1], Q[:, 2], Q[:, 3]
  F_l, _, NF_l = flux_num(eos, Q_l, Qc, lambda)
  F_r, NF_r, _ = flux_num(eos, Qc, Q_r, lambda)
  return Qc - 1.0 / lambda * ((F_r - F_l) + (NF_r + NF_l))
end
function update_cell(Q::Array{Float64,2}, flux_num::Function, eigvals, dtdx, eos::Tuple{Barton2009,Barton2009})
  Q_l, Qc, Q_r = Q[:, 1], Q[:, 2], Q[:, 3]
  F_l, _, NF_l = flux_num(eos, Q_l, Qc, eigvals[1:2])
  F_r, NF_r, _ = flux_num(eos, Qc, Q_r, eigvals[2:3])
  return Qc - dtdx * ((F_r - F_l) + (NF_r + NF_l))
end

# page-lock a Julia Array so that upload! / download! / step_host! copy at the link rate and the chunks of step_host!
# overlap (a Julia Array is pageable memory); call once per array, release before the array is freed
host_register!(A::Array{Float64}) = (GC.@preserve A check(ccall((:hs_host_register, LIB), Cint, (Ptr{Cvoid}, Csize_t), A, sizeof(A))); A)
host_unregister!(A::Array{Float64}) = (GC.@preserve A check(ccall((:hs_host_unregister, LIB), Cint, (Ptr{Cvoid},), A)); A)

# --- single-phase 13-variable model (Hyperelasticity.jl; SURVEY.md A.6) ----------------------------------
# Q = [rho*u(3), rho*F(9, row-major), rho*E], primitives P = [u(3), F(9 row-major), S] = the arguments of
# prim2cons (Hyperelasticity.jl:70).  Same names as the reference module, inside a sub-module because the
# two-phase methods above already own `get_eigvals` / `initial_states` for Tuple{Barton2009,Barton2009}.
module SinglePhase
  import ..Barton2009, ..check, ..LIB, ..HS_MODEL_SP13, ..ncols
  export prim2cons, cons2prim, flux, get_eigvals, initial_states
  for (jl, c) in ((:prim2cons, :hs_prim2cons), (:cons2prim, :hs_cons2prim), (:flux, :hs_flux))
    @eval function $jl(eos::Barton2009, X::VecOrMat{Float64}; device::Integer=0)
      Y = similar(X); e = [eos]
      GC.@preserve X Y e check(ccall(($(QuoteNode(c)), LIB), Cint,
          (Cint, Ptr{Barton2009}, Cint, Ptr{Float64}, Ptr{Float64}, Int64, Cint),
          HS_MODEL_SP13, e, 1, X, Y, ncols(X), device))
      return Y
    end
  end
  function get_eigvals(eos::Barton2009, Q::VecOrMat{Float64}, n::Array{<:Any,1}=[1, 0, 0]; device::Integer=0)
    nn = Vector{Float64}(n); e = [eos]
    E = Q isa AbstractVector ? zeros(6) : zeros(6, size(Q, 2))
    GC.@preserve Q E e nn check(ccall((:hs_get_eigvals, LIB), Cint,
        (Cint, Ptr{Barton2009}, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Cint), HS_MODEL_SP13, e, 1, Q, nn, E, ncols(Q), device))
    return E
  end
  const _R3 = sqrt(3.0)
  const _ROT = [0.5 -0.5*_R3 0; 0.5*_R3 0.5 0; 0 0 1.0]
  # Hyperelasticity.jl:124-172 (F is written row-major into P, :81-86)
  function initial_states(eos::Barton2009, testcase::Integer; device::Integer=0)
    rowmajor(F) = vec(permutedims(F))
    (ul, Fl, Sl, ur, Fr, Sr) =
      testcase == 1 ? ([0.0, 0.5, 1.0], [0.98 0 0; 0.02 1 0.1; 0 0 1], 1e-3, [0.0, 0.0, 0.0], [1.0 0 0; 0 1 0.1; 0 0 1], 0.0) :
      testcase == 2 ? ([2.0, 0.0, 0.1], [1.0 0 0; -0.01 0.95 0.02; -0.015 0 0.9], 0.0, [0.0, -0.03, -0.01], [1.0 0 0; 0.015 0.95 0; -0.01 0 0.9], 0.0) :
      testcase == 3 ? ([1.0, 0.0, 0.0], _ROT, 0.0, [1.0, 0.0, 0.0], _ROT, 0.0) :
                      ([0.0, 0.0, 0.0], [1.0 0 0; 0 1 0; 0 0 1], 0.0, [0.0, 0.0, 0.0], [1.0 0 0; 0 1 0; 0 0 1], 0.0)
    Pl = Float64[ul..., rowmajor(Fl)..., Sl]; Pr = Float64[ur..., rowmajor(Fr)..., Sr]
    return prim2cons(eos, Pl; device=device), prim2cons(eos, Pr; device=device)
  end
end # module SinglePhase

# --- device-resident time loop (main.jl:202-227) ---------------------------------------------------
mutable struct Solver
  ctx::Ptr{Cvoid}
  nvar::Int; ncells::Int; nprob::Int
end

# single-phase solver: Solver(eos::Barton2009, ncells; ...)
function Solver(eos::Barton2009, ncells::Integer; nprob::Integer=1, device::Integer=0)
  ref = Ref{Ptr{Cvoid}}(C_NULL); e = [eos]
  GC.@preserve e check(ccall((:hs_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Cint, Ptr{Barton2009}, Cint, Int64, Int64, Cint),
      ref, HS_MODEL_SP13, e, 1, ncells, nprob, device))
  s = Solver(ref[], 13, ncells, nprob)
  finalizer(destroy!, s)
  return s
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

# one pass of main.jl:204-227 on HOST arrays (Q1 may be Q0): the literal drop-in with the state in Julia memory.  The library
# overlaps the H2D copy, the kernels and the D2H copy chunk by chunk (hs_step_host); host_register!(Q0), host_register!(Q1) once
# beforehand makes the copies asynchronous and link-rate.  Returns dt.
function step_host!(s::Solver, flux::Function, cfl, dx, Q0::Array{Float64}, Q1::Array{Float64})
  dt = zeros(s.nprob)
  GC.@preserve Q0 Q1 dt check(ccall((:hs_step_host, LIB), Cint, (Ptr{Cvoid}, Cint, Float64, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
      s.ctx, flux === hll ? HS_FLUX_HLL : HS_FLUX_LXF, cfl, dx, Q0, Q1, dt))
  return s.nprob == 1 ? dt[1] : dt
end

# restart (main.jl:185-186): set the clock of every problem
set_time!(s::Solver, t, step_num) = check(ccall((:hs_set_time, LIB), Cint, (Ptr{Cvoid}, Float64, Int64), s.ctx, t, step_num))

# `while t < T` without returning to the host between steps; returns (t, step_num)
function advance!(s::Solver, flux::Function, cfl, dx, T; t=0.0, step_num=0, max_steps=typemax(Int32))
  tv = fill(Float64(t), s.nprob); sv = fill(Int64(step_num), s.nprob)
  GC.@preserve tv sv check(ccall((:hs_advance, LIB), Cint,
      (Ptr{Cvoid}, Cint, Float64, Float64, Float64, Int64, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}),
      s.ctx, flux === hll ? HS_FLUX_HLL : HS_FLUX_LXF, cfl, dx, T, max_steps, tv, sv, C_NULL))
  return s.nprob == 1 ? (tv[1], sv[1]) : (tv, sv)
end

# --- dimension-split 2-D solver (hs2d_*; the reference driver is 1-D, its physics takes a normal) ----------------
# Q::Array{Float64,3}(nvar, nx, ny); every sweep is the 1-D step of main.jl:204-227 along the grid lines
mutable struct Solver2D
  ctx::Ptr{Cvoid}
  nvar::Int; nx::Int; ny::Int
end
function Solver2D(eos::Tuple{Barton2009,Barton2009}, nx::Integer, ny::Integer; device::Integer=0)
  ref = Ref{Ptr{Cvoid}}(C_NULL); e = eosvec(eos)
  GC.@preserve e check(ccall((:hs2d_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Cint, Ptr{Barton2009}, Cint, Int64, Int64, Cint),
      ref, HS_MODEL_MPH30, e, 2, nx, ny, device))
  s = Solver2D(ref[], 30, nx, ny)
  finalizer(x -> (x.ctx != C_NULL && ccall((:hs2d_destroy, LIB), Cint, (Ptr{Cvoid},), x.ctx); x.ctx = C_NULL), s)
  return s
end
upload!(s::Solver2D, Q::Array{Float64,3}) = GC.@preserve Q check(ccall((:hs2d_upload, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), s.ctx, Q))
download!(s::Solver2D, Q::Array{Float64,3}) = (GC.@preserve Q check(ccall((:hs2d_download, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), s.ctx, Q)); Q)
function step!(s::Solver2D, flux::Function, cfl, dx, dy)
  dt = Ref{Float64}(0.0)
  check(ccall((:hs2d_step, LIB), Cint, (Ptr{Cvoid}, Cint, Float64, Float64, Float64, Ref{Float64}), s.ctx, flux === hll ? HS_FLUX_HLL : HS_FLUX_LXF, cfl, dx, dy, dt))
  return dt[]
end
function advance!(s::Solver2D, flux::Function, cfl, dx, dy, T; max_steps=typemax(Int32))
  t = Ref{Float64}(0.0); n = Ref{Int64}(0)
  check(ccall((:hs2d_advance, LIB), Cint, (Ptr{Cvoid}, Cint, Float64, Float64, Float64, Float64, Int64, Ref{Float64}, Ref{Int64}),
      s.ctx, flux === hll ? HS_FLUX_HLL : HS_FLUX_LXF, cfl, dx, dy, T, max_steps, t, n))
  return t[], n[]
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
