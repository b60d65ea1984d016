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
#define XF_RED_PPL 16        // [16..18] uvw_c_max of the last GetDt, kept for the positivity-preserving limiter
#define XF_RED_COUNT 20

struct XfDev
{
	int Xmax, Ymax, Zmax, Xp;          // Xp: padded x pitch
	int Xi, Yi, Zi, Bx, By, Bz;        // inner sizes, ghost widths
	int DimX, DimY, DimZ;
	int weno, alpha, ghost;            // scheme + GhostSpecies
	int positivity;                    // equations.PositivityPreserving (read_json.cpp:68)
	long long N;                       // field stride = Xp*Ymax*Zmax (doubles between components)
	long long sY, sZ;                  // cell strides along y and z
	double _dx, _dy, _dz, CFL, gamma0;
	double dx, dy, dz;                 // mesh widths (WENO-CU6 epsilon = 1e-8 dl dl)
	// scalar work arrays [N] (reference FlowData, global_setup.h:218-250) + per-cell pieces of the
	// Roe-averaged pressure derivatives (Utils_device.hpp:14-35), which are pure functions of one cell
	double *u, *v, *w, *p, *H, *c, *T, *g3, *dpdrho, *e, *prho;
	double *y;                         // [NS][N]
	double *dpdrhoi;                   // [NC][N]
	double *Fw[3];                     // wall fluxes [E][N] per direction
	double *red;                       // XF_RED_* slots
	int *err;                          // [4]
	unsigned *hard_ids, *hard_count;   // cells whose Newton iteration needs more than XF_NEWTON_FAST steps (k_prim -> k_prim_hard)
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
