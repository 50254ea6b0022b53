/* libswrt -- C ABI of the B200-native pseudo-spectral flow step and wave-packet ray tracer.
 *
 * Drop-in boundary for the two hot paths of ndefilippis/JuliaRaytracingSW.  The reference has no
 * FFI of its own: the seam is the Julia call surface its drivers use.  Every entry
 * point below cites the reference function it stands in for (paths relative to the reference
 * checkout).  Conventions:
 *   - every function returns 0 on success, a negative SWRT_ERR_* otherwise; swrt_last_error()
 *     returns a thread-local message; nothing throws or calls back;
 *   - host buffers are caller owned, column-major like the Julia arrays they mirror, and are only
 *     touched during the call (calls are synchronous with respect to host buffers);
 *   - device memory is owned by the handles; one CUDA stream per flow handle, packets attached to a
 *     flow run on that flow's stream; a handle is not thread safe;
 *   - there is no CPU fallback: creating a handle without a usable CUDA device fails.
 */
#ifndef SWRT_H
#define SWRT_H

#ifdef __cplusplus
extern "C" {
#endif

#define SWRT_VERSION 100

enum { SWRT_OK = 0, SWRT_ERR_ARG = -1, SWRT_ERR_CUDA = -2, SWRT_ERR_UNSUPPORTED = -3, SWRT_ERR_STATE = -4 };

/* models: rsw/RotatingShallowWater.jl, rsw/ModifiedShallowWater.jl, rsw/LinborgShallowWater.jl,
 * swqg/SWQG.jl, swqg/TwoLayerQG.jl, thomasyamada/ThomasYamada.jl */
enum { SWRT_RSW = 0, SWRT_RSW_MODIFIED = 1, SWRT_RSW_LINDBORG = 2, SWRT_RSW_QUADHEIGHT = 3 /* rsw/QuadHeightModifiedShallowWater.jl */, SWRT_SWQG = 4, SWRT_TWOLAYERQG = 5, SWRT_THOMASYAMADA = 6,
       /* GeophysicalFlows MultiLayerQG with two equal layers (raytracing/TwoLayerRaytracing.jl:174; simulation/TwoLayerSimulation.jl:37-47):
          diagonal L, mean flow / PV gradient / bottom drag inside calcN!; state q_1, q_2 */
       SWRT_MULTILAYERQG2 = 7 };
/* steppers: utils/IFMAB3.jl; FourierFlows FilteredAB3 / ETDRK4 / FilteredRK4 (raytracing/CPUParameters.jl:7) / FilteredETDRK4
 * (raytracing/TestParameters.jl:6: ETDRK4 followed by sol *= filter) */
enum { SWRT_IFMAB3 = 0, SWRT_FILTEREDAB3 = 1, SWRT_ETDRK4 = 2, SWRT_FILTEREDRK4 = 3, SWRT_FILTEREDETDRK4 = 4 };

typedef struct swrt_flow swrt_flow;
typedef struct swrt_packets swrt_packets;

/* = keyword arguments of RotatingShallowWater.Problem (rsw/RotatingShallowWater.jl:70-85) and of
 * IFMAB3TimeStepper / makefilter (utils/IFMAB3.jl:68-88) */
typedef struct swrt_flow_desc {
    int model, stepper;
    int nx, ny;
    int nnu;               /* order of the hyperviscous operator */
    int use_filter;        /* utils/IFMAB3.jl:80-85 */
    int filter_order;      /* makefilter(order=...) */
    int device;            /* CUDA device ordinal */
    double Lx, Ly, dt, nu, f, Cg;
    double aliased_fraction;
    double filter_innerK, filter_outerK, filter_tol;   /* <=0: FourierFlows defaults 2/3, 1, 1e-15 */
    double U, mu, F, Ro, Kd2;                            /* model specific (two-layer, TY, SWQG) */
    double U2, beta;                                     /* MultiLayerQG-2: layer flows (U, U2), planetary PV gradient; F = f0^2/(g' H_j) */
    int slab_rank, slab_size;                            /* slab-decomposed flow over slab_size GPUs (0 or 1: off), see swrt_slab_* */
} swrt_flow_desc;

const char* swrt_last_error(void);
int swrt_version(void);
int swrt_device_count(int* n);

/* Problem(dev; ...)  rsw/RotatingShallowWater.jl:70-99 (grid, params, L, exp(L dt), work arrays) */
int swrt_flow_create(const swrt_flow_desc* desc, swrt_flow** out);
int swrt_flow_destroy(swrt_flow* h);
/* set_solution!(prob, u0h, v0h, eta0h)  :309-321; sol_host is complex128 (nkr, nl, nvar) column-major.
 * The state is stored dealiased (dealias!(sol) is the first statement of every calcN!, :141). */
int swrt_flow_set_solution(swrt_flow* h, const void* sol_host);
/* Array(prob.sol) as left by updatevars!/calcN! (aliased modes are zero) */
int swrt_flow_get_solution(swrt_flow* h, void* sol_host);
/* set_initial_condition! rsw/RSWRaytracingDriver.jl:15-54: random-phase geostrophic band [Kg0, Kg1] scaled to max|u_g| = ag plus wave band
 * [Kw0, Kw1] scaled to max|u_w| = aw (K15), built on the device from the host's random numbers phase = 2 pi rand(nkr, nl) and
 * sgn = sign(rand - 0.5) (column-major (nkr, nl)); scales_out (may be NULL) receives the two normalisation factors */
int swrt_flow_set_rsw_initial_condition(swrt_flow* h, const double* phase_host, const double* sgn_host, double Kg0, double Kg1, double ag,
                                        double Kw0, double Kw1, double aw, double* scales_out);
/* enforce_reality_condition!(prob) :118-133 -- in the reference this leaves sol dealiased and refreshes vars */
int swrt_flow_enforce_reality(swrt_flow* h);
/* stepforward!(prob, [], nsteps): utils/IFMAB3.jl:157-169 looped by FourierFlows.stepforward!(prob, diags, n) */
int swrt_flow_step(swrt_flow* h, int nsteps);
/* addforcing!(N, sol, t, clock, vars::StochasticVars, params, grid) = params.calcF!(vars.Fh, ...); @. N += vars.Fh
 * (rsw/RotatingShallowWater.jl:228-240, ModifiedShallowWater.jl:246-258, QuadHeightModifiedShallowWater.jl:252-264,
 * LinborgShallowWater.jl:239-251).  Fh_host = the (nkr, nl) complex128 field the caller's calcF! produced; it is added to every
 * component of N (the reference's broadcast of the 2-D Fh over the three equations) at every calcN! until replaced; NULL clears.
 * SWRT_ERR_UNSUPPORTED for the QG / Thomas-Yamada models, whose calcN! never calls the hook. */
int swrt_flow_set_forcing(swrt_flow* h, const void* Fh_host);
/* prob.clock.t / prob.clock.step */
int swrt_flow_clock(swrt_flow* h, double* t, long long* step);
int swrt_flow_set_clock(swrt_flow* h, double t, long long step);
/* updatevars!(prob) + Array(vars.<field>) :101-116; real_host is float64 (nx, ny) column-major */
enum { SWRT_FIELD_U = 0, SWRT_FIELD_V = 1, SWRT_FIELD_ETA = 2, SWRT_FIELD_ZETA = 16,
       /* QG models (swqg/SWQG.jl:109-125, swqg/TwoLayerQG.jl:113-129): state variable j = q_j; add the layer index */
       SWRT_FIELD_QG_PSI = 32, SWRT_FIELD_QG_U = 40, SWRT_FIELD_QG_V = 48, SWRT_FIELD_QG_ZETA = 56 };
int swrt_flow_get_field(swrt_flow* h, int which, double* real_host);
/* mul!(varh, grid.rfftplan, field): forward transform of a physical (nx, ny) field into state variable `var` (dealiased),
 * e.g. m0h = rfft(1/(1+eta0)) of rsw/QuadHeightModifiedShallowWater.jl:333-347 or initial conditions given on the grid */
int swrt_flow_set_field_physical(swrt_flow* h, int var, const double* real_host);
/* kinetic_energy(prob), potential_energy(prob): rsw/RotatingShallowWater.jl:323-336, swqg/SWQG.jl:205-222,
 * swqg/TwoLayerQG.jl:221-250 (two-layer: ke = KE_1 + KE_2; the per-layer values through swrt_flow_layer_kinetic_energy) */
int swrt_flow_energies(swrt_flow* h, double* ke, double* pe);
int swrt_flow_layer_kinetic_energy(swrt_flow* h, int layer, double* ke);
/* maximum(abs.(vars.u)), maximum(abs.(vars.v)) (CFL log, raytracing/RaytracingDriver.jl:244) and
 * any(isnan.(vars.uh)) (:282) */
int swrt_flow_max_abs_uv(swrt_flow* h, double* umax, double* vmax);
int swrt_flow_has_nan(swrt_flow* h, int* flag);
/* get_streamfunction! (rsw/RSWRaytracingDriver.jl:56-67) + get_velocity_info
 * (raytracing/RaytracingDriver.jl:132-154) into snapshot slot 0 (old) or 1 (new); stays on device */
/* RSW balanced psi; SWQG psi (swqg/RaytracingDriver.jl); two-layer baroclinic 0.5(psi1-psi2) (swqg/TwoLayerRaytracingDriver.jl:232)
 * and layer mean (psi1+psi2)/2 (raytracing/TwoLayerRaytracing.jl:122) */
enum { SWRT_PSI_RSW_BALANCED = 0, SWRT_PSI_SWQG = 1, SWRT_PSI_TWOLAYER_BAROCLINIC = 2, SWRT_PSI_TWOLAYER_MEAN = 3 };
int swrt_flow_velocity_snapshot(swrt_flow* h, int psi_kind, int slot);
/* which node data velocity_snapshot produces (SWRT_INTERP_*): 5 fields u,v,ux,uy,vx or 7 fields (+ uxy, vxy); packets created with
 * the same interpolant read them.  snapshot_fields returns 5 or 7 (the third extent of get/set_snapshot arrays). */
int swrt_flow_set_interp(swrt_flow* h, int interp);
/* "FFT interpolation": the snapshots' node grid is `refine` (1 or 2) times finer than the flow's, by spectral zero padding
 * (the exact trigonometric interpolant sampled on the finer grid; raytracing/NUFFTRaytracing.jl:68-84 aims at that interpolant,
 * Notebooks/FFTInterpTest.ipynb refines the same way).  Set before creating packet handles; any interpolant then acts on the finer grid. */
int swrt_flow_set_snapshot_refinement(swrt_flow* h, int refine);
int swrt_flow_snapshot_dims(swrt_flow* h, int* nx, int* ny);
int swrt_flow_snapshot_fields(swrt_flow* h, int* nfields);
/* kernel width of the NUFFT mode in nodes of the oversampled grid, 4 <= nw <= 16 (default 8; FINUFFT: nw = ceil(log10(1/tol)) + 1) */
int swrt_flow_set_nufft_width(swrt_flow* h, int nw);
/* old_velocity = new_velocity; old_grad_v = new_grad_v (raytracing/RaytracingDriver.jl:269-270).
 * alias != 0 reproduces the reference's rebinding (both names then refer to the same buffers);
 * alias == 0 swaps the two slots. */
int swrt_flow_swap_snapshots(swrt_flow* h, int alias);
/* Array(u), Array(v), ... of a snapshot: out is float64 (nx, ny, nfields) column-major = u, v, ux, uy, vx[, uxy, vxy] */
int swrt_flow_get_snapshot(swrt_flow* h, int slot, double* out_host);
/* load a snapshot from host fields (same layout) -- used for steady/analytic background flows
 * (raytracing/SteadyRaytracing.jl) and by tests */
int swrt_flow_set_snapshot(swrt_flow* h, int slot, const double* in_host);

/* ---- wave / balanced projections on the device ------------------------------------------------------------
 * wave_balanced_decomposition(prob)  rsw/RSWUtils.jl:5-22 for the eta-based RSW models, decompose_balanced_wave(sol, grid)
 * thomasyamada/TYUtils.jl:40-51 for Thomas-Yamada: complex128 (nkr, nl, 3) arrays; either pointer may be NULL. */
int swrt_flow_wave_balanced_decomposition(swrt_flow* h, void* balanced_host, void* wave_host);
/* compute_balanced_wave_weights with compute_balanced_wave_bases, rsw/RSWUtils.jl:24-57: c0, c+, c- as complex128 (nkr, nl) */
int swrt_flow_wave_balanced_weights(swrt_flow* h, void* c0_host, void* cp_host, void* cm_host);
/* wave_geostrophic_energy(prob)  thomasyamada/ThomasYamada.jl:355-367: out[4] = {KE_wave, PE_wave, KE_balanced, PE_balanced};
 * for RSW the parts go through kinetic_energy / potential_energy of rsw/RotatingShallowWater.jl:323-336.  Reduced on the device. */
int swrt_flow_wave_balanced_energies(swrt_flow* h, double* out);
/* barotropic_energy(prob)  thomasyamada/ThomasYamada.jl:343-350 */
int swrt_flow_barotropic_energy(swrt_flow* h, double* e);

/* ---- k-omega accumulator ----------------------------------------------------------------------------------
 * The reference post-processes stored snapshots with one SLURM task per kr index (thomasyamada/TY_k_omega.jl:46-110,
 * rsw/fourier-analysis/mrsw/FourierRSW.jl:76-160).  Here the series are appended on the device while the flow runs and the
 * windowed transforms in time run on the device at the end.  `kr_index` is 0-based (the reference's k_idx - 1).
 * Series of SWRT_SERIES_TY: 0 ut, 1 vt, 2 ug, 3 vg, 4 uw, 5 vw; spectra 0..5 of those (Hann window, no detrend) and
 * 6 U_balanced, 7 U_wave, 8 U_total.  Series of SWRT_SERIES_RSW: 0..2 u, v, eta; 3..5 balanced; 6..8 wave; 9..11 c0, c+, c-;
 * spectra = clean_fft (detrend + Hann) of each.  Arrays are complex128 (nframes, nl) column-major. */
typedef struct swrt_series swrt_series;
enum { SWRT_SERIES_TY = 0, SWRT_SERIES_RSW = 1 };
int swrt_series_create(swrt_flow* flow, int kind, int kr_index, long long max_frames, swrt_series** out);
int swrt_series_destroy(swrt_series* s);
int swrt_series_append(swrt_series* s);                       /* one frame from the flow's current state and clock */
int swrt_series_frames(swrt_series* s, long long* nframes);
int swrt_series_times(swrt_series* s, double* t_host);
int swrt_series_get(swrt_series* s, int which, void* series_host);
int swrt_series_spectrum(swrt_series* s, int which, void* spectrum_host);

/* ---- slab-decomposed flow step (grids >= 4096^2; one process per GPU) ------------------------------------
 * Rank r owns retained kr columns [r*chunk, (r+1)*chunk) in spectral space and ny/P rows in physical space.  A step is
 *   slab_stage_a  (y-transforms of the local columns)      -> buffer A_SEND, laid out [dest][job][row][chunk]
 *   all-to-all A_SEND -> A_RECV                              (the caller: ncclAllToAll / torch.distributed.all_to_all_single)
 *   slab_stage_b  (x-pass on the local rows)                 A_RECV -> B_SEND
 *   all-to-all B_SEND -> B_RECV
 *   slab_stage_c  (y-transforms back, IFMAB3 update, clock)  from B_RECV
 * (the one-calcN!-per-step steppers IFMAB3 / FilteredAB3; ETDRK4 / FilteredRK4 evaluate calcN! four times per step, each time as the
 * three passes above with a device barrier after A and B: swrt_slab_step, which needs the peers mapped, handles every model and stepper)
 * and a velocity snapshot is slab_psi_a, all-to-all, slab_snap_b (writes this rank's rows of the full snapshot), all-gather.
 * The exchange buffers are owned by the handle; swrt_slab_buffer returns their device pointers so that the caller's
 * communication library can work on them in place.  swrt_flow_set_stream makes the handle launch on the caller's stream
 * (e.g. torch's current stream) so that kernels and collectives are ordered without host synchronisation. */
enum { SWRT_SLAB_A_SEND = 0, SWRT_SLAB_A_RECV = 1, SWRT_SLAB_B_SEND = 2, SWRT_SLAB_B_RECV = 3, SWRT_SLAB_SNAP0 = 4, SWRT_SLAB_SNAP1 = 5,
       SWRT_SLAB_FLAGS = 6, SWRT_SLAB_BAND = 7 };
int swrt_flow_set_stream(swrt_flow* h, void* cuda_stream);
int swrt_slab_buffer(swrt_flow* h, int which, void** device_ptr, long long* nbytes);
/* elements (complex128) per destination of the two all-to-alls for njobs jobs, rows per rank, columns per rank */
int swrt_slab_info(swrt_flow* h, int* yrows, int* chunk, int* njobs_a, int* njobs_b);
/* Direct NVLink transposes: every rank exports the IPC handles (64 bytes each) of its two RECEIVE buffers, the caller distributes
 * them, and each rank opens its peers'.  Once all are open (slab_p2p -> 1) stage_a / stage_b / psi_a store their output straight
 * into the destination ranks' receive buffers; the caller then only needs a stream-ordered barrier (e.g. a one-element all-reduce)
 * where the all-to-all used to be. */
int swrt_slab_ipc_handle(swrt_flow* h, int which, void* handle64);
int swrt_slab_ipc_open(swrt_flow* h, int which, int peer_rank, const void* handle64);
/* first transpose of the slab step: 0 = the y-pass stores its 32-64 byte pieces straight into the peers' receive buffers,
 * 1 = the y-pass stores locally and the x-pass pulls its input segments from the peers' send buffers (needs SWRT_SLAB_A_SEND
 * mapped too), 2 = local stores followed by a block-copy kernel that ships whole lines to the peers */
int swrt_slab_set_mode(swrt_flow* h, int mode);
int swrt_slab_p2p(swrt_flow* h, int* enabled);
int swrt_slab_stage_a(swrt_flow* h);
int swrt_slab_stage_b(swrt_flow* h);
int swrt_slab_stage_c(swrt_flow* h);
int swrt_slab_psi_a(swrt_flow* h, int psi_kind);
int swrt_slab_snap_b(swrt_flow* h, int slot);

/* ---- team mode: the slab step and the coupled loop driven natively, no communication library on the data path ------------
 * With the peers' receive buffers, barrier flags (SWRT_SLAB_FLAGS) and band snapshots (SWRT_SLAB_BAND) mapped through
 * swrt_slab_ipc_handle / swrt_slab_ipc_open, the host language only distributes 64-byte handles once (MPI.jl, sockets, files,
 * torch.distributed ...); after that
 *   swrt_slab_barrier        stream-ordered barrier: a one-CTA kernel stores this rank's epoch into every peer's flag word over
 *                            NVLink and spins until all peers' epochs have arrived (mode 1: stream synchronise + the caller's
 *                            host barrier -- for processes sharing one GPU, where a spinning kernel would starve the peers)
 *   swrt_slab_step           = stepforward!(prob, [], n) of the slab-decomposed problem (stage a, barrier, stage b, barrier, stage c)
 *   swrt_slab_band_snapshot  = get_streamfunction! + get_velocity_info (rsw/RSWRaytracingDriver.jl:56-67,
 *                            raytracing/RaytracingDriver.jl:132-154) for THIS RANK'S BAND of ny/P rows plus `halo` rows of each
 *                            neighbour: packets are sharded by y-band (swrt_packets_desc.band_*), so no rank ever needs the
 *                            whole field and the all-gather of the round-1 design is gone.
 * Packets of such a flow live on the rank whose band holds them and are handed over at every re-sort (sort_every); set / get /
 * generate / sample / raytrace / coupled_steps keep their meaning on the caller-order block of each rank and become COLLECTIVE
 * calls (every rank of the team must make them in the same order).  swrt_flow_get_snapshot / set_snapshot move the rank's own
 * band, an (nx, ny/P, 5) array. */
int swrt_slab_set_barrier(swrt_flow* h, int mode, void (*callback)(void*), void* arg);
int swrt_slab_barrier(swrt_flow* h);
int swrt_slab_step(swrt_flow* h, int nsteps);
int swrt_slab_band_snapshot(swrt_flow* h, int psi_kind, int slot);
int swrt_slab_band_info(swrt_flow* h, int* row0, int* rows, int* halo);

/* timing helpers on the handle's stream (CUDA events) */
int swrt_flow_timer_start(swrt_flow* h);
int swrt_flow_timer_stop(swrt_flow* h, float* ms);
int swrt_flow_sync(swrt_flow* h);
/* per-kernel device timing: enable = 1 brackets every launch with CUDA events on the handle's stream, 2 also clears the
 * accumulators, 0 switches it off.  profile_get(id) -> accumulated ms, launch count and kernel name for id in [0, 13) */
int swrt_flow_profile(swrt_flow* h, int enable);
int swrt_flow_profile_get(swrt_flow* h, int id, double* ms_total, long long* count, const char** name);
/* number of kernels this handle has launched so far */
int swrt_flow_launch_count(swrt_flow* h, long long* n);

/* BILINEAR = the reference's texture sampling (raytracing/GPURaytracing.jl:118-127); HERMITE_BICUBIC = u, v from (f, f_x, f_y, f_xy)
 * node data (utils/CUDAInterpolations.jl:71-108) with the analytic gradient of the interpolant in dk/dt */
enum { SWRT_INTERP_BILINEAR = 0, SWRT_INTERP_HERMITE_BICUBIC = 1,
       /* quadratic B-spline of the CPU tracer (raytracing/Raytracing.jl:161-170); coefficients are prefiltered in the snapshot */
       SWRT_INTERP_BSPLINE2 = 2,
       /* fp32 packet mode: bilinear sampling of Float32 node data with an fp32 right-hand side (the reference's texture path,
          raytracing/GPURaytracing.jl:118-127); packet state and RK4 combination stay fp64.  Reported separately from the fp64 numbers. */
       SWRT_INTERP_BILINEAR_F32 = 3,
       /* cubic B-spline of the CPU tracer's steady-flow mode (raytracing/Raytracing.jl:152-159); prefiltered like BSPLINE2 */
       SWRT_INTERP_BSPLINE3 = 4,
       /* type-2 NUFFT: spectrally exact values of u, v, ux, uy, vx at the packet positions (the intent of raytracing/NUFFTRaytracing.jl:68-84,
          nufft2d2 with tol 1e-5): 2x oversampled node grid (swrt_flow_set_snapshot_refinement(h, 2) first) of the spectrum deconvolved by the
          kernel's transform, sampled with an nw x nw "exponential of semicircle" kernel; error ~ 10^(1 - nw), swrt_flow_set_nufft_width */
       SWRT_INTERP_NUFFT = 5 };
/* classical RK4 (north star) or the CPU tracer's implicit midpoint (raytracing/Raytracing.jl:106-109), 12 fixed-point sweeps */
enum { SWRT_INTEG_RK4 = 0, SWRT_INTEG_IMPLICIT_MIDPOINT = 1 };
enum { SWRT_LERP_PHYSICAL = 0, SWRT_LERP_REFERENCE_GPU = 1 };
typedef struct swrt_packets_desc {
    long long n;            /* Npackets */
    int interp;             /* SWRT_INTERP_* */
    int nsub;               /* RK4 sub-steps per raytrace call */
    int time_lerp;          /* SWRT_LERP_*   */
    int sort_every;         /* re-sort the device copy by grid cell every this many raytrace calls (0 = never);
                               host-visible arrays always keep the caller's row order */
    int integrator;         /* SWRT_INTEG_* */
    double f, Cg;           /* packet_params.f, packet_params.Cg */
    /* Packets attached to a slab-decomposed flow are sharded by y-band (team mode, below): `n` is this rank's caller-order block
     * of rows [band_first, band_first + n) of the ensemble, band_capacity >= n the number of packets the rank can host (the same
     * value on every rank; packets move between ranks as they are advected).  Ignored for a single-GPU flow. */
    long long band_first, band_capacity;
} swrt_packets_desc;

/* create_template_ode(packets) raytracing/GPURaytracing.jl:111-113 -- device state for N packets */
int swrt_packets_create(const swrt_packets_desc* desc, swrt_flow* flow, swrt_packets** out);
int swrt_packets_destroy(swrt_packets* p);
/* packets (N,4) column-major = x, y, k, l ; omega_sign (N)  (raytracing/RaytracingDriver.jl:27-47) */
int swrt_packets_set(swrt_packets* p, const double* xk_host, const double* omega_sign_host);
int swrt_packets_get(swrt_packets* p, double* xk_host);
/* generate_initial_wavepackets on the device (raytracing/RaytracingDriver.jl:27-47); first = global index of
 * this shard's first packet (0-based), ntotal = sqrtN^2 */
int swrt_packets_generate(swrt_packets* p, double L, double k0, long long sqrtN, long long first);
/* raytrace!(tmpl, v_old, v_new, g_old, g_new, grid, packets, dt, (t0,t1), params) raytracing/GPURaytracing.jl:115-142,
 * reading the flow's snapshot slots 0 (old) and 1 (new) */
int swrt_packets_raytrace(swrt_packets* p, double t0, double t1);
/* Which ray kernel integrates the fp64 bilinear RK4 mode (raytracing/GPURaytracing.jl:32-65 dxkdt + :137 solve): SWRT_RAYKERNEL_AUTO
 * (default; the environment knobs of DESIGN.md apply), SWRT_RAYKERNEL_CACHED (per-thread stencil cache, gathers through L1/L2) or
 * SWRT_RAYKERNEL_TILE (one CTA per sort tile, node records of the two levels staged in shared memory by TMA; same arithmetic as
 * CACHED, bit-identical results), SWRT_RAYKERNEL_TILE3 (nsub == 1 only, what AUTO picks then: three staged patches -- first level,
 * mean of the levels, last level -- one per RK4 stage time; agrees with the other two to rounding, ~1e-16 relative per step) or
 * SWRT_RAYKERNEL_PIPE (nsub == 1 only: the TILE3 arithmetic, bit-identical to it, in one persistent CTA per SM whose producer warp
 * stages the next tile while the consumer warps integrate the current one; measured slower than TILE3, kept for comparison). */
enum { SWRT_RAYKERNEL_AUTO = -1, SWRT_RAYKERNEL_CACHED = 0, SWRT_RAYKERNEL_TILE = 1, SWRT_RAYKERNEL_TILE3 = 2, SWRT_RAYKERNEL_PIPE = 3 };
int swrt_packets_set_kernel(swrt_packets* p, int kernel);
/* band-sharded packets (team mode): export / map the 64-byte IPC handle of the handle's arena; packets resident on this rank */
int swrt_packets_ipc_handle(swrt_packets* p, void* handle64);
int swrt_packets_ipc_open(swrt_packets* p, int peer_rank, const void* handle64);
int swrt_packets_resident(swrt_packets* p, long long* n);
/* interpolate_velocity! / interpolate_gradients! :67-109 + Array: u_host (N,2), g_host (N,4) or NULL */
int swrt_packets_sample(swrt_packets* p, int slot, double* u_host, double* g_host);
/* k-cutoff reset raytracing/GPUTwoLayerRaytracing.jl:136-138 */
int swrt_packets_kcutoff_reset(swrt_packets* p, double kcut, double k0, long long* nreset);
/* the hot loop of start_raytracing! (raytracing/RaytracingDriver.jl:256-270): stepforward!(prob, [], 1); get_velocity_info(new);
 * raytrace!(old -> new, (old_t, new_t)); [k-cutoff reset when kcut > 0, raytracing/TwoLayerRaytracing.jl:136-141]; old = new --
 * nsteps times in one call (one ccall per output period instead of five per step; matters on launch-bound grid sizes) */
int swrt_packets_coupled_steps(swrt_packets* p, int psi_kind, int nsteps, double kcut, double k0);

/* ---- overlapped packet I/O -----------------------------------------------------------------------------
 * Give the handle its own CUDA stream; its copies and kernels then overlap the flow's stream and other packet handles
 * (events order them against the snapshots they read).  Split an ensemble over several handles and issue, per handle,
 * set_async -> raytrace -> get_async / sample_async: uploads, kernels and downloads of different handles pipeline on the two
 * copy engines and the SMs.  Host buffers must be page-locked for the copies to be asynchronous; `ld` is the host arrays'
 * leading dimension in elements (>= n), so handles can address row blocks of one (N, ncol) column-major array.
 * Results are valid after swrt_packets_sync. */
int swrt_packets_use_own_stream(swrt_packets* p);
int swrt_packets_set_async(swrt_packets* p, const double* xk_host, long long ld, const double* sign_host);
int swrt_packets_get_async(swrt_packets* p, double* xk_host, long long ld);
int swrt_packets_sample_async(swrt_packets* p, int slot, double* u_host, double* g_host, long long ld);
int swrt_packets_sync(swrt_packets* p);

/* Output roll-over arithmetic (host, integer only): utils/SequencedOutputs.jl:37-63, utils/Collated.jl:40-60 */
typedef struct swrt_seqout { long long max_writes, current_writes, file_index; } swrt_seqout;
int swrt_seqout_init(swrt_seqout* s, long long max_writes);
/* one `out[key] = val`; returns in *file_index the file the key went to */
int swrt_seqout_write(swrt_seqout* s, long long nwrites, long long* file_index);
int swrt_seqout_filename(const char* base, long long idx, char* buf, int buflen);     /* "%s.%06d.jld2" */
int swrt_collated_filename(const char* base, long long idx, char* buf, int buflen);   /* "%s_%08d.out"  */

#ifdef __cplusplus
}
#endif
#endif
