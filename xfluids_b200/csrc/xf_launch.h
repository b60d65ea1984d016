// xf_launch.h -- host-callable launchers exported by each compiled flavour of xf_kernels.cu
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include "xf_types.h"

// TMA tensor maps of one sweep input for one direction: conserved variables [E][N], (u v w p c) [5][N], Y [NC][N], each a 4-D
// tensor (x, y, z, component) with a box of XF_MW x TF cells along the sweep
struct XfTma
{
	CUtensorMap U, P, Y;
};

#define XF_DECLARE_LAUNCHERS(NS)                                                                                              \
	namespace NS                                                                                                              \
	{                                                                                                                         \
		int launch_prim(const XfDev &d, const XfThermo &th, int ns, int cop, double *U, int flags, cudaStream_t s, long long *launches, int k0, int k1); \
		int launch_prim_shell(const XfDev &d, const XfThermo &th, int ns, int cop, double *U, int flags, cudaStream_t s, long long *launches); \
		int launch_rk_prim(const XfDev &d, const XfThermo &th, int ns, int cop, double *U, double *U1, const double *dt_dev, int flag, int guard, int nflags, \
						   cudaStream_t s, long long *launches);                                                            \
		int launch_sweeps(const XfDev &d, int ns, int cop, const double *U, cudaStream_t s, long long *launches, int dirmask, int kp0, int kp1, int tz0, int tz1, \
						  const CUtensorMap *tmy, const CUtensorMap *tmz, const XfViscF *vf);                                   \
		int xf_z_tiles(const XfDev &d);                                                                                       \
		int launch_visc(const XfDev &d, const XfThermo &th, const XfVisc &vs, int ns, int cop, const double *U, const int bc[6], cudaStream_t s, long long *launches, int parts); \
		int launch_sweep_x(const XfDev &d, int ns, int cop, const double *U, const XfMarchArgs &a, cudaStream_t s);           \
		int launch_march(const XfDev &d, int ns, int cop, const XfTma &tm, const double *UI, const XfMarchArgs &a, int dir, cudaStream_t s); \
		int z_tile_faces();                                                                                                   \
		int launch_lu(const XfDev &d, int E, double *LU, cudaStream_t s);                                                     \
		int launch_rk(const XfDev &d, int E, double *U, double *U1, const double *LU, double dt, const double *dt_dev,       \
					  int flag, int guard, int fused, cudaStream_t s, int ka, int kb);                                        \
		int launch_nan(const XfDev &d, int E, const double *UI, const double *LU, cudaStream_t s);                           \
		int launch_bc(const XfDev &d, int E, int cop, double *U, const int bc[6], cudaStream_t s, long long *launches, int dirmask, int k0, int k1); \
		int launch_dt(const XfDev &d, const double *rho, cudaStream_t s);                                                     \
		int launch_dt_final(const XfDev &d, double t_end, cudaStream_t s);                                                    \
		int launch_layout(const XfDev &d, int E, double *soa, double *aos, int to_soa, cudaStream_t s, long long row0, long long nrows); \
		int launch_scalar_pad(const XfDev &d, double *padded, double *flat, int to_padded, cudaStream_t s);                  \
		int launch_halo(const XfDev &d, int E, double *U, double *buf, int k0, int pack, cudaStream_t s);                    \
	}

XF_DECLARE_LAUNCHERS(xf_strict)
XF_DECLARE_LAUNCHERS(xf_fast)
