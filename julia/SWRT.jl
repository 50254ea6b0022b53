# SWRT.jl -- thin ccall layer over libswrt.so (include/swrt.h), mirroring the call surface the reference's
# drivers use: Problem / set_solution! / stepforward! / updatevars! / raytrace! / interpolate_*!.
#
# UNVERIFIED IN THIS REPOSITORY'S CI: no Julia toolchain exists in the build image; this file is the binding a
# maintainer of ndefilippis/JuliaRaytracingSW would add (see INTEGRATION.md).  It uses only Base + Libdl.
module SWRT

using Libdl

const libswrt = get(ENV, "SWRT_LIB", joinpath(@__DIR__, "..", "juliaraytracingsw_b200", "lib", "libswrt.so"))

struct SwrtError <: Exception
    code::Cint
    msg::String
end
check(rc::Cint) = rc == 0 ? nothing : throw(SwrtError(rc, unsafe_string(ccall((:swrt_last_error, libswrt), Cstring, ()))))

# --- descriptors: field order == include/swrt.h ---------------------------------------------------------------
Base.@kwdef struct FlowDesc
    model::Cint = 0;  stepper::Cint = 0
    nx::Cint = 128;   ny::Cint = 128
    nnu::Cint = 4;    use_filter::Cint = 0;  filter_order::Cint = 4;  device::Cint = 0
    Lx::Cdouble = 2π; Ly::Cdouble = 2π; dt::Cdouble = 5e-2; nu::Cdouble = 1e-16; f::Cdouble = 1.0; Cg::Cdouble = 1.0
    aliased_fraction::Cdouble = 1/3
    filter_innerK::Cdouble = 2/3; filter_outerK::Cdouble = 1.0; filter_tol::Cdouble = 1e-15
    U::Cdouble = 0.0; mu::Cdouble = 0.0; F::Cdouble = 0.0; Ro::Cdouble = 0.0; Kd2::Cdouble = 0.0
    U2::Cdouble = 0.0; beta::Cdouble = 0.0
    slab_rank::Cint = 0; slab_size::Cint = 0
end

Base.@kwdef struct PacketsDesc
    n::Clonglong
    interp::Cint = 0; nsub::Cint = 1; time_lerp::Cint = 0; sort_every::Cint = 16   # interp: 0 bilinear, 1 Hermite bicubic, 2 quadratic B-spline, 3 bilinear fp32, 4 cubic B-spline
    integrator::Cint = 0                                                           # 0 RK4, 1 implicit midpoint
    f::Cdouble = 1.0; Cg::Cdouble = 1.0
    band_first::Clonglong = 0; band_capacity::Clonglong = 0                        # team mode (slab-decomposed flow): see include/swrt.h
end

const MODELS = Dict("RotatingShallowWater" => 0, "ModifiedShallowWater" => 1, "LinborgShallowWater" => 2,
                    "QuadHeightModifiedShallowWater" => 3, "SWQG" => 4, "TwoLayerQG" => 5, "ThomasYamada" => 6, "MultiLayerQG" => 7)
const STEPPERS = Dict("IFMAB3" => 0, "FilteredAB3" => 1, "ETDRK4" => 2, "FilteredRK4" => 3, "FilteredETDRK4" => 4)
const NVAR = Dict(0 => 3, 1 => 3, 2 => 3, 3 => 3, 4 => 1, 5 => 2, 6 => 4, 7 => 2)

# --- flow: RotatingShallowWater.Problem and friends (rsw/RotatingShallowWater.jl:70-133, 309-336) ---------------
mutable struct Problem
    h::Ptr{Cvoid}
    nx::Int; ny::Int; nkr::Int; dt::Float64; nvar::Int
    function Problem(; model = "RotatingShallowWater", stepper = "IFMAB3", nx = 128, ny = nx, Lx = 2π, Ly = Lx, ν = 1e-16, nν = 4,
                     f = 1.0, Cg = 1.0, dt = 5e-2, aliased_fraction = 1/3, use_filter = false, order = 4, dev = 0,
                     U = 0.5, μ = 1e-2, f0 = f, δρρ0 = 0.2, Ro = 0.2, H = [0.5, 0.5], b = [2.0, 1.0], β = 0.0,
                     slab_rank = 0, slab_size = 0)
        # MultiLayerQG.Problem(2, dev; nx, Lx, f₀, H, b, U, μ, β, dt, stepper, aliased_fraction): two equal layers, U = [U₁, U₂]
        F = model == "MultiLayerQG" ? f0^2 / ((b[1] - b[2]) * H[1]) : 2 * f0^2 / Cg^2 / δρρ0
        U1, U2 = U isa Number ? (U, -U) : (U[1], U[2])
        d = FlowDesc(model = MODELS[model], stepper = STEPPERS[stepper], nx = nx, ny = ny, Lx = Lx, Ly = Ly, nu = ν, nnu = nν, f = f,
                     Cg = Cg, dt = dt, aliased_fraction = aliased_fraction, use_filter = use_filter, filter_order = order, device = dev,
                     U = U1, mu = μ, F = F, Ro = Ro, U2 = U2, beta = β, slab_rank = slab_rank, slab_size = slab_size)
        out = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:swrt_flow_create, libswrt), Cint, (Ref{FlowDesc}, Ref{Ptr{Cvoid}}), d, out))
        p = new(out[], nx, ny, nx ÷ 2 + 1, dt, NVAR[MODELS[model]])
        finalizer(q -> ccall((:swrt_flow_destroy, libswrt), Cint, (Ptr{Cvoid},), q.h), p)
        return p
    end
end

"set_solution!(prob, fields...): one (nkr, nl) ComplexF64 array per state variable (u0h, v0h, η0h | q0h layers | ζ0h, u0h, v0h, p0h)"
function set_solution!(prob::Problem, fields...)
    sol = Array{ComplexF64}(undef, prob.nkr, prob.ny, prob.nvar)
    for (i, f) in enumerate(fields)
        sol[:, :, i] .= f
    end
    check(ccall((:swrt_flow_set_solution, libswrt), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}), prob.h, sol))
end
"mul!(varh, grid.rfftplan, field) into state variable `var` (1-based)"
set_field_physical!(prob::Problem, var::Integer, field::Matrix{Float64}) =
    check(ccall((:swrt_flow_set_field_physical, libswrt), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}), prob.h, var - 1, field))
"Array(prob.sol)"
function solution(prob::Problem)
    sol = Array{ComplexF64}(undef, prob.nkr, prob.ny, prob.nvar)
    check(ccall((:swrt_flow_get_solution, libswrt), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}), prob.h, sol))
    return sol
end
"vars.Fh of the forcing hook (addforcing!, rsw/RotatingShallowWater.jl:228-240): added to every component of N at each calcN!; `nothing` clears"
set_forcing!(prob::Problem, Fh::Matrix{ComplexF64}) =
    check(ccall((:swrt_flow_set_forcing, libswrt), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}), prob.h, Fh))
set_forcing!(prob::Problem, ::Nothing) = check(ccall((:swrt_flow_set_forcing, libswrt), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), prob.h, C_NULL))
enforce_reality_condition!(prob::Problem) = check(ccall((:swrt_flow_enforce_reality, libswrt), Cint, (Ptr{Cvoid},), prob.h))
"stepforward!(prob, diags, nsteps) -- diags: objects with .freq and increment!(d, prob)"
function stepforward!(prob::Problem, diags = [], nsteps::Integer = 1)
    if isempty(diags)
        check(ccall((:swrt_flow_step, libswrt), Cint, (Ptr{Cvoid}, Cint), prob.h, nsteps))
    else
        for _ in 1:nsteps
            check(ccall((:swrt_flow_step, libswrt), Cint, (Ptr{Cvoid}, Cint), prob.h, 1))
            s = clock(prob).step
            for d in diags
                s % d.freq == 0 && d.increment!(d, prob)
            end
        end
    end
end
function clock(prob::Problem)
    t = Ref{Cdouble}(0); s = Ref{Clonglong}(0)
    check(ccall((:swrt_flow_clock, libswrt), Cint, (Ptr{Cvoid}, Ref{Cdouble}, Ref{Clonglong}), prob.h, t, s))
    return (t = t[], step = Int(s[]), dt = prob.dt)
end
"updatevars!(prob) + Array(vars.<field>): which = 0 u, 1 v, 2 η, 16 ζ"
function field(prob::Problem, which::Integer)
    a = Array{Float64}(undef, prob.nx, prob.ny)
    check(ccall((:swrt_flow_get_field, libswrt), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}), prob.h, which, a))
    return a
end
function energies(prob::Problem)
    ke = Ref{Cdouble}(0); pe = Ref{Cdouble}(0)
    check(ccall((:swrt_flow_energies, libswrt), Cint, (Ptr{Cvoid}, Ref{Cdouble}, Ref{Cdouble}), prob.h, ke, pe))
    return (kinetic_energy = ke[], potential_energy = pe[])
end
kinetic_energy(prob::Problem) = energies(prob).kinetic_energy
potential_energy(prob::Problem) = energies(prob).potential_energy

# --- wave / balanced projections on the device: rsw/RSWUtils.jl:5-57, thomasyamada/TYUtils.jl:40-51 ---------------
"wave_balanced_decomposition(prob) / decompose_balanced_wave(sol, grid): (balanced, wave) as (nkr, nl, 3) arrays"
function wave_balanced_decomposition(prob::Problem)
    bal = Array{ComplexF64}(undef, prob.nkr, prob.ny, 3); wav = similar(bal)
    check(ccall((:swrt_flow_wave_balanced_decomposition, libswrt), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{ComplexF64}), prob.h, bal, wav))
    return bal, wav
end
"compute_balanced_wave_weights with compute_balanced_wave_bases: (c₀, c₊, c₋)"
function compute_balanced_wave_weights(prob::Problem)
    c = [Array{ComplexF64}(undef, prob.nkr, prob.ny) for _ in 1:3]
    check(ccall((:swrt_flow_wave_balanced_weights, libswrt), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{ComplexF64}), prob.h, c[1], c[2], c[3]))
    return Tuple(c)
end
"wave_geostrophic_energy(prob) (thomasyamada/ThomasYamada.jl:355-367): ((KE_w, PE_w), (KE_g, PE_g))"
function wave_geostrophic_energy(prob::Problem)
    out = zeros(Cdouble, 4)
    check(ccall((:swrt_flow_wave_balanced_energies, libswrt), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), prob.h, out))
    return ((out[1], out[2]), (out[3], out[4]))
end
function barotropic_energy(prob::Problem)
    e = Ref{Cdouble}(0)
    check(ccall((:swrt_flow_barotropic_energy, libswrt), Cint, (Ptr{Cvoid}, Ref{Cdouble}), prob.h, e))
    return e[]
end

# --- k-omega accumulator: thomasyamada/TY_k_omega.jl:46-110, rsw/fourier-analysis/mrsw/FourierRSW.jl:76-160 ------
mutable struct KOmega
    h::Ptr{Cvoid}
    prob::Problem
    "k_idx is 1-based like the reference's; kind 0 = Thomas-Yamada series, 1 = RSW series"
    function KOmega(prob::Problem, k_idx::Integer, max_frames::Integer; kind = 0)
        out = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:swrt_series_create, libswrt), Cint, (Ptr{Cvoid}, Cint, Cint, Clonglong, Ref{Ptr{Cvoid}}), prob.h, kind, k_idx - 1, max_frames, out))
        s = new(out[], prob)
        finalizer(q -> ccall((:swrt_series_destroy, libswrt), Cint, (Ptr{Cvoid},), q.h), s)
        return s
    end
end
append!(s::KOmega) = check(ccall((:swrt_series_append, libswrt), Cint, (Ptr{Cvoid},), s.h))
function nframes(s::KOmega)
    n = Ref{Clonglong}(0)
    check(ccall((:swrt_series_frames, libswrt), Cint, (Ptr{Cvoid}, Ref{Clonglong}), s.h, n))
    return Int(n[])
end
"fft(window .* series, 1) / clean_fft(t, series, window) of series `which` (0-based, see include/swrt.h): (nframes, nl)"
function spectrum(s::KOmega, which::Integer)
    out = Array{ComplexF64}(undef, nframes(s), s.prob.ny)
    check(ccall((:swrt_series_spectrum, libswrt), Cint, (Ptr{Cvoid}, Cint, Ptr{ComplexF64}), s.h, which, out))
    return out
end

# --- packets: raytracing/GPURaytracing.jl -----------------------------------------------------------------------
"get_streamfunction! + get_velocity_info into snapshot slot (0 = old, 1 = new)"
get_velocity_info!(prob::Problem, slot::Integer; psi_kind = 0) =
    check(ccall((:swrt_flow_velocity_snapshot, libswrt), Cint, (Ptr{Cvoid}, Cint, Cint), prob.h, psi_kind, slot))
"node data of the snapshots: 0 bilinear (u,v,ux,uy,vx), 1 Hermite bicubic (+ uxy, vxy)"
set_interpolation!(prob::Problem, interp::Integer) = check(ccall((:swrt_flow_set_interp, libswrt), Cint, (Ptr{Cvoid}, Cint), prob.h, interp))
"FFT interpolation: snapshots on a node grid `refine` (1 or 2) times finer than the flow's, by spectral zero padding"
set_snapshot_refinement!(prob::Problem, refine::Integer) =
    check(ccall((:swrt_flow_set_snapshot_refinement, libswrt), Cint, (Ptr{Cvoid}, Cint), prob.h, refine))
"old_velocity = new_velocity; old_grad_v = new_grad_v"
swap_snapshots!(prob::Problem; alias = false) =
    check(ccall((:swrt_flow_swap_snapshots, libswrt), Cint, (Ptr{Cvoid}, Cint), prob.h, alias))

mutable struct Packets
    h::Ptr{Cvoid}
    n::Int
    prob::Problem
    function Packets(prob::Problem, n; f, Cg, nsub = 1, time_lerp = 0, sort_every = 16, interp = 0, integrator = 0, band_first = 0, band_capacity = 0)
        d = PacketsDesc(n = n, interp = interp, integrator = integrator, nsub = nsub, time_lerp = time_lerp, sort_every = sort_every, f = f, Cg = Cg,
                        band_first = band_first, band_capacity = band_capacity)
        out = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:swrt_packets_create, libswrt), Cint, (Ref{PacketsDesc}, Ptr{Cvoid}, Ref{Ptr{Cvoid}}), d, prob.h, out))
        p = new(out[], n, prob)
        finalizer(q -> ccall((:swrt_packets_destroy, libswrt), Cint, (Ptr{Cvoid},), q.h), p)
        return p
    end
end
"generate_initial_wavepackets(dev, L, k0, Npackets, sqrtNpackets) on the device; `first` = 0-based global row of this shard"
function generate_initial_wavepackets(prob, L, k0, Npackets, sqrtNpackets; f, Cg, first = 0, kw...)
    p = Packets(prob, Npackets; f, Cg, kw...)
    check(ccall((:swrt_packets_generate, libswrt), Cint, (Ptr{Cvoid}, Cdouble, Cdouble, Clonglong, Clonglong), p.h, L, k0, sqrtNpackets, first))
    return p
end
generate!(p::Packets, L, k0, sqrtNpackets, first = 0) =
    check(ccall((:swrt_packets_generate, libswrt), Cint, (Ptr{Cvoid}, Cdouble, Cdouble, Clonglong, Clonglong), p.h, L, k0, sqrtNpackets, first))
set_packets!(p::Packets, xk::Matrix{Float64}, ωsign::Vector{Float64}) =
    check(ccall((:swrt_packets_set, libswrt), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}), p.h, xk, ωsign))
function Base.Array(p::Packets)
    xk = Matrix{Float64}(undef, p.n, 4)
    check(ccall((:swrt_packets_get, libswrt), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), p.h, xk))
    return xk
end
create_template_ode(p::Packets) = p
"raytrace!(tmpl, v_old, v_new, g_old, g_new, grid, packets, dt, (t0, t1), params) -- snapshot slots 0/1 of the flow"
raytrace!(tmpl, v1, v2, g1, g2, grid, p::Packets, dt, tspan, params = nothing) =
    check(ccall((:swrt_packets_raytrace, libswrt), Cint, (Ptr{Cvoid}, Cdouble, Cdouble), p.h, tspan[1], tspan[2]))
"interpolate_velocity!/interpolate_gradients! + Array: returns (U (N,2), G (N,4))"
function interpolate_velocity_and_gradients(p::Packets, slot::Integer)
    U = Matrix{Float64}(undef, p.n, 2); G = Matrix{Float64}(undef, p.n, 4)
    check(ccall((:swrt_packets_sample, libswrt), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ptr{Cdouble}), p.h, slot, U, G))
    return U, G
end
# overlapped packet I/O: the handle's own stream, asynchronous copies of row blocks (ld = rows of the parent (N, ncol) array)
use_own_stream!(p::Packets) = check(ccall((:swrt_packets_use_own_stream, libswrt), Cint, (Ptr{Cvoid},), p.h))
set_packets_async!(p::Packets, xk::Ptr{Cdouble}, ld::Integer, ωsign::Ptr{Cdouble}) =
    check(ccall((:swrt_packets_set_async, libswrt), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Clonglong, Ptr{Cdouble}), p.h, xk, ld, ωsign))
get_packets_async!(p::Packets, xk::Ptr{Cdouble}, ld::Integer) =
    check(ccall((:swrt_packets_get_async, libswrt), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Clonglong), p.h, xk, ld))
sample_async!(p::Packets, slot::Integer, U::Ptr{Cdouble}, G::Ptr{Cdouble}, ld::Integer) =
    check(ccall((:swrt_packets_sample_async, libswrt), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Clonglong), p.h, slot, U, G, ld))
"RAYKERNEL_AUTO = -1, CACHED = 0, TILE = 1 (two staged levels), TILE3 = 2 (three staged levels, nsub == 1), PIPE = 3 (persistent variant of TILE3)"
set_kernel!(p::Packets, kernel::Integer) = check(ccall((:swrt_packets_set_kernel, libswrt), Cint, (Ptr{Cvoid}, Cint), p.h, kernel))
sync!(p::Packets) = check(ccall((:swrt_packets_sync, libswrt), Cint, (Ptr{Cvoid},), p.h))
"the hot loop (stepforward!; get_velocity_info; raytrace!; old = new) nsteps times in one ccall"
coupled_steps!(p::Packets, nsteps::Integer; psi_kind = 0, k_cutoff = 0.0, k0 = 0.0) =
    check(ccall((:swrt_packets_coupled_steps, libswrt), Cint, (Ptr{Cvoid}, Cint, Cint, Cdouble, Cdouble), p.h, psi_kind, nsteps, k_cutoff, k0))
function kcutoff_reset!(p::Packets, k_cutoff, k0)
    n = Ref{Clonglong}(0)
    check(ccall((:swrt_packets_kcutoff_reset, libswrt), Cint, (Ptr{Cvoid}, Cdouble, Cdouble, Ref{Clonglong}), p.h, k_cutoff, k0, n))
    return Int(n[])
end

# --- team mode: slab-decomposed flow + y-band-sharded packets, one Julia process per GPU (include/swrt.h "team mode") ----------
# The host only distributes 64-byte CUDA-IPC handles once (here: any `allgather(bytes)::Vector` the caller supplies, e.g.
# MPI.Allgather); every step after that is native.  which: 1 A_RECV, 3 B_RECV, 6 FLAGS, 7 BAND.
const TEAM_SHARED = (1, 3, 6, 7)
function open_team!(prob::Problem, rank::Integer, world::Integer, allgather)
    for which in TEAM_SHARED
        mine = Vector{UInt8}(undef, 64)
        check(ccall((:swrt_slab_ipc_handle, libswrt), Cint, (Ptr{Cvoid}, Cint, Ptr{UInt8}), prob.h, which, mine))
        for (r, h) in enumerate(allgather(mine))
            check(ccall((:swrt_slab_ipc_open, libswrt), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}), prob.h, which, r - 1, h))
        end
    end
    check(ccall((:swrt_slab_set_mode, libswrt), Cint, (Ptr{Cvoid}, Cint), prob.h, world >= 4 ? 2 : 0))
end
"stepforward!(prob, [], n) of the slab-decomposed problem"
slab_stepforward!(prob::Problem, n::Integer = 1) = check(ccall((:swrt_slab_step, libswrt), Cint, (Ptr{Cvoid}, Cint), prob.h, n))
"get_streamfunction! + get_velocity_info for this rank's band of rows (+ halo)"
slab_velocity_snapshot!(prob::Problem, slot::Integer; psi_kind = 0) =
    check(ccall((:swrt_slab_band_snapshot, libswrt), Cint, (Ptr{Cvoid}, Cint, Cint), prob.h, psi_kind, slot))
team_barrier!(prob::Problem) = check(ccall((:swrt_slab_barrier, libswrt), Cint, (Ptr{Cvoid},), prob.h))
function open_team!(p::Packets, allgather)
    mine = Vector{UInt8}(undef, 64)
    check(ccall((:swrt_packets_ipc_handle, libswrt), Cint, (Ptr{Cvoid}, Ptr{UInt8}), p.h, mine))
    for (r, h) in enumerate(allgather(mine))
        check(ccall((:swrt_packets_ipc_open, libswrt), Cint, (Ptr{Cvoid}, Cint, Ptr{UInt8}), p.h, r - 1, h))
    end
end
function resident(p::Packets)
    n = Ref{Clonglong}(0)
    check(ccall((:swrt_packets_resident, libswrt), Cint, (Ptr{Cvoid}, Ref{Clonglong}), p.h, n))
    return Int(n[])
end
"set_initial_condition! (rsw/RSWRaytracingDriver.jl:15-54) on the device from phase = 2π rand(nkr, nl), sgn = sign.(rand(nkr, nl) .- 0.5)"
function set_rsw_initial_condition!(prob::Problem, phase::Matrix{Float64}, sgn::Matrix{Float64}, Kg, ag, Kw, aw)
    scales = zeros(2)
    check(ccall((:swrt_flow_set_rsw_initial_condition, libswrt), Cint,
                (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Ptr{Cdouble}),
                prob.h, phase, sgn, Kg[1], Kg[2], ag, Kw[1], Kw[2], aw, scales))
    return scales
end
set_nufft_width!(prob::Problem, nw::Integer) = check(ccall((:swrt_flow_set_nufft_width, libswrt), Cint, (Ptr{Cvoid}, Cint), prob.h, nw))

end # module
