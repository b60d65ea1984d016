// xf_types.h -- device-side parameter blocks shared by the kernels and the C-ABI layer.
#pragma once
#include <cstddef>

#define XF_MAXS 6            // max NUM_SPECIES carried in kernel-parameter thermo tables
#define XF_PITCH_ALIGN 16    // x pitch in doubles (128-byte rows)

// slots of XfDev::red (device doubles)
#define XF_RED_DTMAX 0       // [0..2]  max(|u_d| + sqrt(1.4 p/rho)) per direction        (GlobalDt_block.hpp:34-66)
#define XF_RED_GLF 3         // [3..11] running max |lambda|: dir*3 + {u-c, u, u+c}       (ConVenction_block.hpp:115-170)
#define XF_RED_DT 12         // device-resident dt
#define XF_RED_TIME 13       // device-resident physical time
#define XF_RED_STEPS 14      // time steps taken with dt > 0 (k_dt_final): what xf_run reports, even if a replayed batch overshoots t_end
#define XF_RED_PPL 16        // [16..18] uvw_c_max of the last GetDt, kept for the positivity-preserving limiter
#define XF_RED_COUNT 20

// tiled y / z sweeps (k_sweep): tile = XF_TW_ cells in x by XF_TF_ faces along the sweep; the TMA box of a tile's conserved pencil has
// XF_TF_ + NST - 1 rows (NST = 6 stencil cells for WENO5 / CU6, 8 for WENO7)
#ifndef XF_TW_
#define XF_TW_ 32
#endif
#ifndef XF_TF_
#define XF_TF_ 8
#endif
#define XF_TILE_ROWS(weno) (XF_TF_ + ((weno) == 7 ? 8 : 6) - 1)

// marching y / z sweeps (k_march, xf_march.cuh): tile = XF_MW cells in x by TF faces along the sweep per iteration
#ifndef XF_MW
#define XF_MW 16
#endif
#ifndef XF_MTF_
#define XF_MTF_ 16
#endif
#define XF_MTF(weno) ((weno) == 7 && XF_MTF_ == 16 ? 15 : XF_MTF_)   // WENO7's 8-cell stencil needs two more ring rows: 15 faces keep two blocks per SM at Emax = 9
#ifndef XF_MARCH_MINB
#define XF_MARCH_MINB 2     // resident blocks per SM k_march is compiled for
#endif
#ifndef XF_MARCH_ALIAS
#define XF_MARCH_ALIAS 0    // 1: the flux-exchange rows live in the (consumed) landing buffer: 2 more barriers per iteration, 19 KB less shared memory per block
#endif
#ifndef XF_MARCH_SIDEPF
#define XF_MARCH_SIDEPF 1   // 1: the face-side scalars of the next iteration are loaded before the closing barrier of this one
#endif
// what a sweep does with the wall flux it has formed
#define XF_MODE_FW 0    // store it (block-level API: FluxFw / FluxGw / FluxHw of the reference)
#define XF_MODE_ACC 1   // LU (+)= (F_{f-1} - F_f) * _dl   (UpdateFluidLU, one direction at a time in the reference's x -> y -> z order)
#define XF_MODE_RK 2    // the same for the last direction, then NaN guard + SSP-RK3 stage update; LU never leaves the registers

struct XfDev
{
	int Xmax, Ymax, Zmax, Xp;          // Xp: padded x pitch
	int Xi, Yi, Zi, Bx, By, Bz;        // inner sizes, ghost widths
	int DimX, DimY, DimZ;
	int weno, alpha, ghost;            // scheme + GhostSpecies
	int positivity;                    // equations.PositivityPreserving (read_json.cpp:68)
	int nsm;                           // SM count of the context's device
	long long N;                       // field stride = Xp*Ymax*Zmax (doubles between components)
	long long sY, sZ;                  // cell strides along y and z
	double _dx, _dy, _dz, CFL, gamma0;
	double dx, dy, dz;                 // mesh widths (WENO-CU6 epsilon = 1e-8 dl dl)
	// scalar work arrays [N] (reference FlowData, global_setup.h:218-250) + per-cell pieces of the
	// Roe-averaged pressure derivatives (Utils_device.hpp:14-35), which are pure functions of one cell
	double *u, *v, *w, *p, *H, *c, *T, *g3, *dpdrho, *e, *prho; // u, v, w, p, c are ONE allocation [5][N] in that order (one TMA tensor map)
	double *y;                         // [NS][N]
	double *dpdrhoi;                   // [NC][N]
	double *Fw[3];                     // wall fluxes [E][N] per direction (allocated by the first xf_get_lu; the fused path never stores them)
	double *red;                       // XF_RED_* slots
	int *err;                          // [4]
	unsigned *hard_ids, *hard_count;   // cells whose Newton iteration needs more than XF_NEWTON_FAST steps (k_prim -> k_prim_hard)
};

// viscous / heat-conduction / species-diffusion terms (reference: solver_Reconstruction/viscosity/**; Visc, Visc_Heat, Visc_Diffu of
// cmake/init_options.cmake:81-92 made run-time).  Transport fits of Setup::GetFitCoefficient (viscfit.cpp:148-190): cubic polynomials in ln T.
struct XfVisc
{
	int on, heat, diffu;
	double fit_visc[XF_MAXS][4], fit_therm[XF_MAXS][4]; // ln(mu_k), ln(lambda_k)
	double fit_Dkj[XF_MAXS * XF_MAXS][4];               // ln(p D_kj)
	double Wi[XF_MAXS];                                 // kg/mol
	// the two factors of PHI(k, i) (Visc_device.h:34-41) that depend on the molar masses only, from the host's pow():
	// phiW[k * NS + i] = pow(W_i / W_k, 0.25), phiS[k * NS + i] = pow(1 + W_k / W_i, -0.5)
	double phiW[XF_MAXS * XF_MAXS], phiS[XF_MAXS * XF_MAXS];
	int dkj_sym;                                        // fit_Dkj[i][k] == fit_Dkj[k][i] bit for bit: each pair's coefficient is evaluated once
	double Yil_limiter, Dim_limiter, dim_max0;          // Block::Yil_limiter / Dim_limiter (iniset.cpp:358-359); Dim_max before scaling (0: the single-process build's value)
	// work arrays
	double *Vde;   // [9][N] velocity derivatives, order ducx dvcx dwcx ducy dvcy dwcy ducz dvcz dwcz (global_setup.h VdeType)
	double *va, *tca; // [N] mixture viscosity, thermal conductivity
	double *Dkm;   // [NS][N] mixture-averaged diffusion coefficients
	double *hi;    // [NS][N] species enthalpies
	double *lim;   // device doubles: [0..NS) yi_min, [NS..2NS) yi_max -> after k_visc_limits: [2NS..3NS) Yil_limiter, [3NS..4NS) Diffu_limiter
};

// what the wall-flux part of the viscous block reads (the tail of a sweep carries it by value)
struct XfViscF
{
	int heat, diffu;
	const double *Vde, *va, *tca, *Dkm, *hi, *lim;
};
static inline XfViscF xf_visc_face_args(const XfVisc &v)
{
	XfViscF f;
	f.heat = v.heat, f.diffu = v.diffu, f.Vde = v.Vde, f.va = v.va, f.tca = v.tca, f.Dkm = v.Dkm, f.hi = v.hi, f.lim = v.lim;
	return f;
}

// Cells no ghost fill / halo pack reads: at least one more ghost width away from every face of the inner block.  The stage update can go
// straight on to their primitive recovery (k_rk_prim); the shell around them waits for the ghost fill (k_prim_shell).  GhostSpecies
// renormalisation rewrites U, so a cell must be recovered exactly once per stage and only after every reader of its raw update is done.
struct XfDeep
{
	int xlo, xhi, ylo, yhi, zlo, zhi; // deep cells: [xlo, xhi) x [ylo, yhi) x [zlo, zhi) in array indices (empty: all lo = hi = max)
};
static inline XfDeep xf_deep_cells(const XfDev &d)
{
	XfDeep p;
	p.xlo = d.DimX ? 2 * d.Bx : 0, p.xhi = d.DimX ? d.Xmax - 2 * d.Bx : d.Xmax;
	p.ylo = d.DimY ? 2 * d.By : 0, p.yhi = d.DimY ? d.Ymax - 2 * d.By : d.Ymax;
	p.zlo = d.DimZ ? 2 * d.Bz : 0, p.zhi = d.DimZ ? d.Zmax - 2 * d.Bz : d.Zmax;
	if (p.xhi <= p.xlo || p.yhi <= p.ylo || p.zhi <= p.zlo)
		p.xlo = p.xhi = d.Xmax, p.ylo = p.yhi = d.Ymax, p.zlo = p.zhi = d.Zmax;
	return p;
}

// arguments of one sweep launch beyond the block description (k_sweep x direction, k_march y / z)
struct XfMarchArgs
{
	int mode;             // XF_MODE_FW / ACC / RK
	int first;            // this direction is the first active one: LU starts from 0.0
	int flag, guard;      // RK stage 1..3; NaN guard on
	int t0, t1;           // transverse range (x, y sweeps: z-planes [t0, t1); z sweep: rows j, always [By, By + Yi))
	int ca, cb;           // y / z: cells [ca, cb) along the sweep to form the divergence for (absolute indices); faces ca - 1 .. cb - 1 are computed
	int nseg, seglen;     // y / z: the march is cut in nseg segments (gridDim.z) of seglen cells
	double *U, *U1, *LU;  // RK: fields of the update; ACC: LU
	double *Fw;           // FW: this direction's wall-flux field
	const double *dt_dev;
};

// NASA-9 tables, re-laid-out per temperature range so that a warp-uniform branch on the range gives
// uniform constant-bank addresses.  hcoef: {-a1, a2, a3, a4/2, a5/3, a6/4, a7/5, b1}; ccoef: {a1..a7}
// (Thermo_device.h:10-23,62-80).  The 1/2,1/3,1/4,1/5 products are formed on the host in the same
// FP64 operations the reference performs per call, so values are bit-identical.
struct XfThermo
{
	double hcoef[3][XF_MAXS][8];
	double ccoef[3][XF_MAXS][7];
	double Ri[XF_MAXS], _Wi[XF_MAXS];
	double Ru;
};
