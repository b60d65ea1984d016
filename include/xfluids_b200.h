/* xfluids_b200.h -- C ABI of the B200-native XFluids inviscid-RHS engine (libxfluids_b200.so).
 *
 * The reference has no plugin/FFI layer; its de-facto seam is the set of host "block" functions
 * that class Fluid forwards to (reference src/Fluids.cpp:585-588,897-961), which take POD structs
 * and raw device pointers.  Every entry point below replaces one of them and cites it.  All
 * functions return 0 on success or a negative xf_status; no exceptions cross this boundary, no
 * torch / C++ types appear in it.
 *
 * Data layout on the device (DESIGN.md "Data layout"): structure-of-arrays, FP64,
 *     field[n][k][j][i]  at  d_field[n * xf_field_stride(ctx) + (k*Ymax + j)*xf_pitch(ctx) + i]
 * with the x pitch padded to a multiple of 16 doubles (128 B rows).  The reference's AoS layout
 * A[Emax*id+n], id = Xmax*Ymax*k + Xmax*j + i (src/solver_UpdateStates/Update_kernels.hpp:14-15, also the
 * payload of its checkpoint file, src/XFLUIDS.cpp:658-687) is what xf_upload_aos/xf_download_aos speak.
 *
 * There is no CPU fallback: every function fails with XF_ERR_CUDA when no device is usable.
 */
#ifndef XFLUIDS_B200_H
#define XFLUIDS_B200_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef enum xf_status {
	XF_OK = 0,
	XF_ERR_ARG = -1,      /* bad argument / unsupported configuration */
	XF_ERR_CUDA = -2,     /* CUDA runtime error (xf_last_error() has the text) */
	XF_ERR_NUMERIC = -3,  /* a guard kernel found NaN/Inf/negative state (reference "error captured") */
	XF_ERR_COMM = -4      /* multi-GPU exchange error */
} xf_status;

/* reference enum BConditions, src/include/global_setup.h:96-110 */
enum { XF_BC_INFLOW = 0, XF_BC_OUTFLOW = 1, XF_BC_SYMMETRY = 2, XF_BC_PERIODIC = 3, XF_BC_NSLIPWALL = 4,
       XF_BC_VISCWALL = 5, XF_BC_SLIPWALL = 6, XF_BC_INNERBLOCK = 7, XF_BC_COPY = 99 /* neighbour rank: halo */ };

/* the fields of reference struct Block / MeshSize (global_setup.h:155-190) that the path uses;
 * metric formulas are the caller's (reference src/read_ini/src/iniset.cpp:290-336). */
typedef struct xf_block {
	int X_inner, Y_inner, Z_inner;
	int Bwidth_X, Bwidth_Y, Bwidth_Z;   /* 0 in an inactive dimension */
	int Xmax, Ymax, Zmax;               /* inner + 2*Bwidth, 1 in an inactive dimension */
	int DimX, DimY, DimZ;
	double dx, dy, dz, _dx, _dy, _dz;   /* _dx = 1/dx exactly as the host computed it */
	double CFLnumber;
} xf_block;

/* reference struct Thermal (global_setup.h:192-199) + the compile-time mixture macros made run-time */
typedef struct xf_thermal {
	int num_species;       /* NUM_SPECIES (1 when cop == 0) */
	int cop;               /* 1: multi-component (COP), NASA-9 thermo; 0: single gamma-law gas */
	int ghost_species;     /* GhostSpecies: last species is inert filler, Y renormalised (Update_device.hpp:15-22) */
	double ncop_gamma;     /* NCOP_Gamma (cop == 0) */
	const double *Hia;     /* host, [num_species*7*3]  Hia[n*21 + m*3 + range]   (src/read_ini/src/thermal.cpp:50-88) */
	const double *Hib;     /* host, [num_species*2*3] */
	const double *Ri;      /* host, [num_species]  Ru/Wi */
	const double *_Wi;     /* host, [num_species]  1/Wi  */
} xf_thermal;

/* reference compile-time scheme macros (cmake/init_options.cmake) made run-time */
typedef struct xf_scheme {
	int weno_order;        /* SCHEME_ORDER: 5 (WENO5-JS, weno5old), 6 (WENO-CU6, WENO6s_schemes.hpp:5-78) or 7 (WENO7-JS) */
	int artificial_type;   /* Artificial_type: 1 ROE, 2 LLF, 3 GLF */
	int fp_mode;           /* 0 strict: no FMA contraction, reference summation order (parity mode);
	                          1 fast:   FMA contraction allowed (same formulas, same order) */
	int positivity;        /* equations.PositivityPreserving (read_json.cpp:68): the flux limiter of
	                          PositivityPreserving_kernels.hpp:5-76, fused into the tail of each sweep */
} xf_scheme;

/* viscous / heat-conduction / species-diffusion terms.  Compile-time in the reference (Visc, Visc_Heat, Visc_Diffu,
 * cmake/init_options.cmake:81-92), run-time here; the transport fits are the host's (Setup::GetFitCoefficient, viscfit.cpp:148-190). */
typedef struct xf_transport {
	int visc, visc_heat, visc_diffu;
	const double *fit_visc;    /* host, [num_species*4]  ln(mu_k)     = sum_m c[m] (ln T)^m   (fitted_coefficients_visc)  */
	const double *fit_therm;   /* host, [num_species*4]  ln(lambda_k)                          (fitted_coefficients_therm) */
	const double *fit_Dkj;     /* host, [num_species*num_species*4]  ln(p D_kj)                (Dkj_matrix)                */
	const double *Wi;          /* host, [num_species]  molar masses, kg/mol                    (species_chara[.. + Wi])    */
	double Yil_limiter, Dim_limiter;   /* Block::Yil_limiter, Block::Dim_limiter (src/read_ini/src/iniset.cpp:358-359) */
	double dim_max0;           /* Dim_max before scaling: 0.0 = the reference's single-process build (its Dkm reduction is commented
	                              out, ConVenction_block.hpp:478-479, so the diffusion fluxes are clipped to zero), 1.0 = its MPI build */
} xf_transport;

typedef struct xf_ctx xf_ctx;

/* ---- life cycle ---------------------------------------------------------------------------- */
/* replaces Fluid::AllocateFluidMemory + Setup::CpyToGPU (Fluids.cpp:270-583, src/read_ini/src/gpucopy.cpp:5-36):
 * allocates the primitive / wall-flux work arrays and uploads the thermo tables. */
int xf_create(const xf_block *bl, const xf_thermal *th, const xf_scheme *sc, int device, xf_ctx **out);
int xf_destroy(xf_ctx *ctx);
const char *xf_last_error(void);
/* switches the viscous terms on (visc != 0) for every later xf_get_lu / xf_rk_stage / xf_run: allocates the velocity-derivative and
 * transport-coefficient work arrays (21 + 2 num_species scalars per cell) and copies the fits.  The wall fluxes then are
 * inviscid (+ limiter) - viscous, as in GetLU (ConVenction_block.hpp:424-575). */
int xf_set_transport(xf_ctx *ctx, const xf_transport *tr);
int xf_set_stream(xf_ctx *ctx, void *cuda_stream);   /* all later launches go to this stream (default: 0) */
int xf_synchronize(xf_ctx *ctx);

/* ---- geometry of the device layout ------------------------------------------------------------ */
size_t xf_pitch(const xf_ctx *ctx);          /* padded x extent, in doubles */
size_t xf_field_stride(const xf_ctx *ctx);   /* doubles between consecutive components n of one field */
size_t xf_field_doubles(const xf_ctx *ctx);  /* Emax * field_stride: size of U, U1 or LU */
int xf_emax(const xf_ctx *ctx);

/* ---- conserved-variable fields owned by the caller (reference Fluid::d_U, d_U1, d_LU) ----------- */
int xf_field_alloc(xf_ctx *ctx, double **d_field);           /* zero-filled */
int xf_field_free(xf_ctx *ctx, double *d_field);
int xf_upload_aos(xf_ctx *ctx, double *d_field, const double *h_aos);        /* host AoS [N*Emax] -> device SoA */
int xf_download_aos(xf_ctx *ctx, const double *d_field, double *h_aos);      /* device SoA -> host AoS */
/* scalar work arrays by reference name: "T" "p" "u" "v" "w" "H" "c" "rho" "gamma" "e" ; y as "y<k>".
 * set: host [N] in reference cell order -> device; get: device -> host.  T must be set before the first
 * xf_update_states: it is the Newton warm start and a state variable. */
int xf_set_scalar(xf_ctx *ctx, const char *name, const double *h);
int xf_get_scalar(xf_ctx *ctx, const char *name, double *h);
int xf_get_wallflux_aos(xf_ctx *ctx, int dir, double *h_aos);                /* FluxFw/Gw/Hw of the last xf_get_lu */

/* ---- the block-level entry points (all asynchronous on the ctx stream unless they return a host value) */
/* float FluidBoundaryCondition(queue&, Setup, BConditions[6], real_t*)   src/solver_BCs/BCs_block.cpp:4-218 */
int xf_boundary(xf_ctx *ctx, double *d_UI, const int bc[6]);
/* pair<bool,...> UpdateFluidStateFlux(queue&, Setup, Thermal, real_t *UI, FlowData&, ...)
 *                                     src/solver_UpdateStates/UpdateStates_block.cpp:7-191 (K1..K4; mutates UI under GhostSpecies).
 * error (may be NULL): receives 1 when a guard fired (synchronises); NULL keeps the call asynchronous. */
int xf_update_states(xf_ctx *ctx, double *d_UI, int *error);
/* vector<float> GetLU(queue&, Setup&, Block, BConditions[6], Thermal, real_t *UI, real_t *LU, ...)
 *                                     src/solver_Reconstruction/FDM_Method/ConVenction_block.hpp:10-617 (inviscid branch) */
int xf_get_lu(xf_ctx *ctx, const double *d_UI, double *d_LU);
/* bool Fluid::EstimateFluidNAN(queue&, int flag)   src/Fluids.cpp:963-1040 */
int xf_estimate_nan(xf_ctx *ctx, const double *d_UI, const double *d_LU, int *error);
/* void UpdateURK3rd(queue&, Block, real_t *U, real_t *U1, real_t *LU, real_t dt, int flag)
 *                                     src/solver_UpdateStates/UpdateStates_block.cpp:193-211 */
int xf_update_u_rk3(xf_ctx *ctx, double *d_U, double *d_U1, const double *d_LU, double dt, int flag);
/* real_t GetDt(queue&, Block, Thermal&, FlowData&, real_t *uvw_c_max)   src/solver_GetDt/GlobalDt_block.hpp:6-114
 * uvw_c_max (may be NULL) receives the three directional maxima (for the cross-rank MAX reduction). */
int xf_get_dt(xf_ctx *ctx, double *dt, double uvw_c_max[3]);

/* ---- fused path: one SSP-RK3 stage / whole steps, state stays on the device ----------------------- */
/* XFLUIDS::RungeKuttaSP3rd (src/XFLUIDS.cpp:441-525): BC -> UpdateStates -> GetLU -> NaN guard -> UpdateU.
 * dt is read from the device-resident value written by xf_dt_device. */
int xf_rk_stage(xf_ctx *ctx, double *d_U, double *d_U1, double *d_LU, const int bc[6], int flag);
/* device-resident ComputeTimeStep (src/XFLUIDS.cpp:196-199,527-544): dt = CFL / sum(max_d * _dd), clipped so that
 * time + dt <= t_end; time += dt.  Uses the maxima gathered by the last xf_update_states. */
int xf_dt_device(xf_ctx *ctx, double t_end);
/* nsteps time steps of XFLUIDS::Evolution's inner loop (src/XFLUIDS.cpp:172-294) without host round trips
 * (captured once into a CUDA graph and replayed).  Stops early at t_end.  Synchronises at the end. */
int xf_run(xf_ctx *ctx, double *d_U, double *d_U1, double *d_LU, const int bc[6], int nsteps, double t_end,
           int *steps_done, double *time_out, int *error);
int xf_get_time(xf_ctx *ctx, double *time, double *last_dt);
int xf_set_time(xf_ctx *ctx, double time);
int xf_error_flags(xf_ctx *ctx, int flags[4]);      /* [0] rho/Yi guard [1] primitive guard [2] U/LU NaN guard; synchronises */
int xf_clear_errors(xf_ctx *ctx);
/* device addresses for the multi-GPU glue (one process per GPU; the exchange itself is the caller's NCCL/P2P):
 * the 3 directional dt maxima (double[3]) and the error word (int[4]). */
double *xf_device_dtmax(xf_ctx *ctx);
int *xf_device_errors(xf_ctx *ctx);
/* 1 when the species-diffusion limiter of the viscous block reads domain-wide extrema of the mass fractions (xf_transport.dim_max0 != 0, the
 * reference's MPI build: ConVenction_block.hpp:489-503 MPI-reduces yi_min / yi_max); the built-in slab driver refuses such a setup on N > 1 ranks
 * instead of limiting with per-slab extrema */
int xf_transport_needs_global_extrema(const xf_ctx *ctx);
double *xf_device_glfmax(xf_ctx *ctx);   /* 9 device doubles: the GLF running maxima of |lambda| (dir*3 + {u-c, u, u+c}), eigen_block_{x,y,z} of ConVenction_block.hpp:115-215 */

/* ---- z-slab halo (replaces FluidMpiCopyKernelZ pack/unpack, src/solver_BCs/BCs_kernels.hpp:304-324 and
 *      MpiTrans::MpiTransBuf, src/mpiPacks/mpiPacks.cpp:357-505): Bwidth_Z planes x Emax, contiguous per component */
size_t xf_halo_doubles(const xf_ctx *ctx);                               /* Emax * Bwidth_Z * Ymax * pitch */
int xf_halo_pack(xf_ctx *ctx, const double *d_UI, int face, double *d_buf);   /* face 4 = zmin inner planes, 5 = zmax */
int xf_halo_unpack(xf_ctx *ctx, double *d_UI, int face, const double *d_buf); /* into the ghost planes of that face */

int xf_halo_pack_on(xf_ctx *ctx, const double *d_UI, int face, double *d_buf, void *cuda_stream);   /* same, on an explicit stream */
int xf_halo_unpack_on(xf_ctx *ctx, double *d_UI, int face, const double *d_buf, void *cuda_stream);
/* One RK stage split around the exchange so that it overlaps with compute (the reference's exchange is blocking and
 * barrier-fenced, BCs_block.cpp:50-217 / mpiPacks.cpp:405-494):
 *   xf_boundary -> xf_halo_pack -> [send/recv + xf_halo_unpack_on a second stream]
 *   xf_stage_interior: primitive recovery of the planes that do not depend on the incoming halo (all but the z ghosts) and
 *                      the x and y sweeps (they only read inner z planes), each adding its part of the flux divergence to d_LU
 *   [wait for the unpack]
 *   xf_stage_finish:   primitive recovery of the z ghost planes, the z sweep with its part of the divergence, NaN guard + RK update
 * Results are bit-identical to xf_rk_stage. */
int xf_stage_interior(xf_ctx *ctx, double *d_U, double *d_U1, double *d_LU, int flag);
/* the same stage split between primitive recovery and sweeps instead: with global Lax-Friedrichs splitting on N > 1 GPUs the caller
 * MAX-reduces xf_device_glfmax over the ranks between the two calls (the reference's MPI build reduces eigen_block the same way) */
int xf_stage_states(xf_ctx *ctx, double *d_U, double *d_U1, int flag);
int xf_stage_fluxes(xf_ctx *ctx, double *d_U, double *d_U1, double *d_LU, int flag);
int xf_stage_finish(xf_ctx *ctx, double *d_U, double *d_U1, double *d_LU, int flag);

/* ---- multi-GPU: z-slab decomposition over the GPUs of one box (csrc/xf_slab.cu) -------------------------------------------------
 * Replaces the reference's MPI layer for this path: MpiTrans::MpiTransBuf / FluidMpiCopyKernelZ (src/mpiPacks/mpiPacks.cpp:357-505,
 * src/solver_BCs/BCs_block.cpp:50-217) and the MPI_Allreduce(MAX) of Fluid::GetFluidDt (src/Fluids.cpp:902-913).  Transport: NCCL
 * send/recv over NVLink + all-reduce(MAX), loaded at run time (libnccl.so.2).  One xf_comm per rank; ranks are processes
 * (torchrun) or threads of one process (`xfluids -mpi=1,1,N`).  Rank r owns Z_inner planes; its z faces carry XF_BC_COPY towards
 * a neighbour rank (the host Setup computes that list, mpiPacks.cpp:44-72; periodic z: also the two outer faces). */
typedef struct xf_comm xf_comm;
typedef struct xf_slab xf_slab;
const char *xf_slab_last_error(void);
int xf_comm_unique_id(char id[128]);                                     /* rank 0: ncclGetUniqueId; hand the bytes to every rank */
int xf_comm_create(const char id[128], int rank, int world, int device, xf_comm **out);   /* collective over the ranks */
int xf_comm_destroy(xf_comm *comm);
int xf_comm_rank(const xf_comm *comm);
int xf_comm_world(const xf_comm *comm);
int xf_comm_allreduce_max(xf_comm *comm, double *d_values, int n, void *cuda_stream);      /* in place, device memory */
int xf_comm_allreduce_max_int(xf_comm *comm, int *d_values, int n, void *cuda_stream);
/* the slab stepper of one rank: bc = this rank's six face conditions; all launches go to compute_stream (xf_set_stream is called) */
int xf_slab_create(xf_ctx *ctx, xf_comm *comm, const int bc[6], int artificial_type, int weno_order, void *compute_stream, xf_slab **out);
int xf_slab_destroy(xf_slab *slab);
int xf_slab_set_overlap(xf_slab *slab, int on);                          /* 0: blocking exchange (the reference's order); default on */
int xf_slab_neighbours(const xf_slab *slab, int lo_hi[2]);               /* neighbour ranks across zmin / zmax, -1 = physical boundary */
int xf_slab_halo(xf_slab *slab, double *d_field);                        /* pack -> send/recv -> unpack on the compute stream */
int xf_slab_startup(xf_slab *slab, double *d_U, int *error);             /* main.cpp:44-48: ghost fill + exchange + UpdateStates */
int xf_slab_stage(xf_slab *slab, double *d_U, double *d_U1, double *d_LU, int flag);   /* one RK stage, exchange overlapped with interior work */
int xf_slab_step(xf_slab *slab, double *d_U, double *d_U1, double *d_LU, double t_end);   /* dt MAX-reduction + device dt + 3 stages; asynchronous */
int xf_slab_run(xf_slab *slab, double *d_U, double *d_U1, double *d_LU, int nsteps, double t_end, int *steps_done, double *time_out, int *error);
int xf_slab_any_error(xf_slab *slab, int *error);                        /* guard flags MAX-reduced over the ranks; synchronises */
int xf_slab_step_host(xf_slab *slab, double *h_U_aos_pinned, double t_end, double *d_U, double *d_U1, double *d_LU, int *error);
int xf_slab_allreduce_max_host(xf_slab *slab, double *h_values, int n);  /* n <= 8 host doubles (block-level GetFluidDt) */

/* ---- host-buffer convenience used for end-to-end timing: upload AoS U, run nsteps, download AoS U ---- */
int xf_step_host(xf_ctx *ctx, double *h_U_aos_pinned, const int bc[6], int nsteps, double t_end,
                 double *d_U, double *d_U1, double *d_LU, int *steps_done, int *error);
/* nsteps == 1 on a 3-D block: the upload goes up in `chunks` z-chunks on a copy stream while the plane-local part of stage 1
 * (AoS->SoA, x / y ghost fill, primitive recovery, x and y sweeps) runs on each arrived chunk, and stage 3 runs in z-chunks
 * (sweeps, update, SoA->AoS) with the download of each finished chunk behind it; bit-identical to the plain sequence.
 * chunks <= 1 switches the overlap off (default 32). */
int xf_set_host_overlap(xf_ctx *ctx, int chunks);
/* the same overlapped step in three pieces, for a z-slab caller that has halo exchanges to put in between (xfluids_b200/slab.py):
 *   [MAX-reduce xf_device_dtmax over the ranks]
 *   xf_host_begin          chunked upload + plane-local part of stage 1 on the planes that do not feed the z ghost fill / halo
 *   [z-halo exchange of U]
 *   xf_host_stage1_finish  z ghost fill (physical faces), the waiting planes, z sweep, update
 *   [stage 2: xf_boundary + halo + xf_rk_stage / the split stage; then xf_boundary + halo on U1]
 *   xf_host_stage3         primitive recovery of U1, then sweeps / update / SoA->AoS / download per z-chunk; returns when the host buffer is complete
 * Not available (XF_ERR_ARG) for 1-D / 2-D blocks, chunks <= 1 or global Lax-Friedrichs splitting. */
int xf_host_begin(xf_ctx *ctx, double *h_U_aos_pinned, const int bc[6], double t_end, double *d_U, double *d_U1, double *d_LU);
int xf_host_stage1_finish(xf_ctx *ctx, const int bc[6], double *d_U, double *d_U1, double *d_LU);
int xf_host_stage3(xf_ctx *ctx, double *h_U_aos_pinned, double *d_U, double *d_U1, double *d_LU, int *error);
void *xf_host_alloc_pinned(size_t bytes);
void xf_host_free_pinned(void *p);

/* ---- measurement support (bench.py): no reference counterpart; the reference's wall-clock timers are
 *      ConVenction_block.hpp:592-614 / UpdateStates_block.cpp:186-190 ------------------------------------------ */
/* one eager time step with CUDA events between the kernels of every stage.  ms[0] dt, [1] boundary fill, [2] primitive
 * recovery, [3..5] sweep x/y/z, [6] flux divergence + RK update (each summed over the 3 stages), [7] whole step. */
int xf_profile_step(xf_ctx *ctx, double *d_U, double *d_U1, double *d_LU, const int bc[6], double t_end, float ms[8]);
/* roofline denominators measured on `device`: FP64 FMA rate (TFLOP/s, FMA = 2 flop), copy bandwidth (GB/s, read+write) */
int xf_measure_peaks(int device, double *dfma_tflops, double *copy_gbs);

/* FP64 issue rates of `device`, 1e12 thread-instructions per second: [0] DADD [1] DMUL [2] DFMA -- the strict (no-FMA) build issues DADD /
 * DMUL where a contracted build issues DFMA, so its ceiling is an ISSUE rate, not a flop rate */
int xf_measure_fp64_issue(int device, double tinst[3]);
/* copy rates of this context's link for `bytes` of pinned host memory (GB/s each way, nothing else running): the e2e leg's denominators.
 * NOTE: the d2h pass overwrites the host buffer with the staging buffer's content -- call it on a scratch / about-to-be-refilled buffer. */
int xf_measure_pcie(xf_ctx *ctx, void *h_pinned, size_t bytes, double *h2d_gbs, double *d2h_gbs);

/* y[i] = the device logarithm of x[i] (host arrays, evaluated on `device`).  The NASA-9 enthalpy's log(T) (reference
 * src/solver_Ini/Thermo_device.h:62-80) is the one non-IEEE-basic operation of the path; the device version replays glibc's
 * table-driven algorithm so that it matches the reference CPU path bit for bit (csrc/xf_log.cuh); this entry point lets the
 * tests prove it against the host libm. */
int xf_log_eval(int device, const double *h_x, double *h_y, size_t n);
/* the same for the three transcendental functions of the viscous block's transport fits (reference
 * src/solver_Reconstruction/viscosity/Visc_device.h:10-41): which = 0 log(x), 1 exp(x), 2 pow(x, y2) -- the device versions replay
 * glibc's table-driven algorithms (csrc/xf_log.cuh, csrc/xf_exp.cuh); h_y2 may be null unless which = 2. */
int xf_math_eval(int device, int which, const double *h_x, const double *h_y2, double *h_out, size_t n);

/* kernel launch counter (bench.py "gpu_launches") */
long long xf_launch_count(const xf_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif
