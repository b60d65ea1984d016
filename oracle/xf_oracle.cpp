// oracle/xf_oracle.cpp -- TEST INFRASTRUCTURE: CPU restatement of the XFluids inviscid hot path.
//
// This file is the parity ORACLE for the CUDA kernels in xfluids_b200/csrc.  It is a plain,
// serial C++ restatement of the reference's per-RK-stage inviscid right-hand side, written from
// the reference's algorithm (file:line cited on every function; paths relative to
// /root/reference/src).  It keeps the reference's data layout (AoS conserved arrays
// A[Emax*id+n], id = Xmax*Ymax*k + Xmax*j + i, species y[NUM_SPECIES*id+n]), its evaluation
// order inside every expression and its quirks, so that, compiled with
// -O2 -ffp-contract=off, it reproduces the reference built through oracle/build_ref.sh
// bit for bit (pinned by tests/test_oracle_vs_ref.py against oracle/_ref and tests/golden).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
// It is never linked into, or called by, the product (xfluids_b200/), which has no CPU path.
//
// Compile-time macros of the reference become run-time fields of xo_cfg:
//   COP -> cop, GhostSpecies -> ghost_species, NUM_SPECIES -> NS, Emax, NUM_COP -> NCOP,
//   SCHEME_ORDER -> weno (5|6|7; 6 = WENO-CU6), Artificial_type -> alpha (1 ROE, 2 LLF, 3 GLF), NCOP_Gamma.
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <vector>

#define XO_MAXS 16 // upper bound for NUM_SPECIES / Emax-4 in fixed-size scratch arrays
#define XO_MAXE (XO_MAXS + 4)

extern "C"
{
	typedef struct
	{
		int Xmax, Ymax, Zmax, X_inner, Y_inner, Z_inner, Bw_X, Bw_Y, Bw_Z;
		int DimX, DimY, DimZ;
		int NS, Emax, NCOP;
		int cop, ghost_species, weno, alpha;
		int positivity; // equations.PositivityPreserving (read_json.cpp:68)
		double dx, dy, dz, _dx, _dy, _dz, CFL, ncop_gamma;
		int bc[6];
		const double *Hia, *Hib, *Ri, *_Wi; // NASA-9 tables Hia[n*21+m*3+range], Hib[n*6+m*3+range]; Ri=Ru/Wi; _Wi=1/Wi
		// viscous / heat-conduction / species-diffusion terms (Visc, Visc_Heat, Visc_Diffu; cmake/init_options.cmake:81-92)
		int visc, visc_heat, visc_diffu;
		double Yil_limiter, Dim_limiter, dim_max0; // Block limiters (iniset.cpp:358-359); Dim_max before scaling (0: single-process build)
		const double *fit_visc, *fit_therm, *fit_Dkj, *Wi; // Setup::GetFitCoefficient results [NS*4], [NS*4], [NS*NS*4]; Wi kg/mol
	} xo_cfg;

	// every array of one fluid, reference names (Fluids.cpp:299-339, global_setup.h FlowData)
	typedef struct
	{
		double *U, *U1, *LU, *FluxF, *FluxG, *FluxH, *FluxFw, *FluxGw, *FluxHw;
		double *eig_x, *eig_y, *eig_z;		  // eigen_local_{x,y,z}[Emax*N]
		double *eigen_block_x, *eigen_block_y, *eigen_block_z; // [Emax], running max (never reset)
		double *rho, *p, *u, *v, *w, *c, *gamma, *e, *H, *T, *y, *Ri, *Cp;
		double uvw_c_max[6];
		int error_flags[4]; // [0] rho/yi NaN guard, [1] primitive guard, [2] U/LU NaN guard, [3] unused
		// viscous work arrays (FlowData: Vde[9], viscosity_aver, thermal_conduct_aver, Dkm_aver[N*NS], hi[N*NS])
		double *Vde[9], *va, *tca, *Dkm, *hi;
	} xo_state;
}

// ---- constants: include/global_setup.h:38-54 ---------------------------------------------------
// Conditioning experiment (tests/test_conditioning.py): -DXO_PERTURB_LOG builds a variant whose log() result is moved by one
// ulp for about half of the arguments -- the size of difference any other libm (e.g. CUDA libdevice) is allowed to have.
#ifdef XO_PERTURB_LOG
#include <cstdint>
#include <cstring>
static inline double xo_perturbed_log(double x)
{
	double r = std::log(x);
	uint64_t b;
	std::memcpy(&b, &x, 8);
	return (b & 1) ? std::nextafter(r, 1e300) : r;
}
#define XO_LOG(x) xo_perturbed_log(x)
#else
#define XO_LOG(x) std::log(x)
#endif
static const double _OT = (1.0 / 3.0);
static const double _six = 1.0 / 6.0; // schemes/Utils_schemes.hpp:5
static const double Avogadro = 6.02214076e26;
static const double Boltzmann = 1.380649e-23;
static const double universal_gas_const = Avogadro * Boltzmann;
static const double Ru = universal_gas_const * 1.0E-3;

// ---- solver_Ini/Thermo_device.h:10-23 ----------------------------------------------------------
static inline double HeatCapacity_NASA(const double *Hia, const double T0, const double Ri, const int n)
{
	double T = std::fmax(T0, 200.0); // sycl::max on non-NaN doubles
	double Cpi = 0.0, _T = 1.0 / T;
	const double *a = Hia + n * 7 * 3;
	if (T >= 1000.0 && T < 6000.0)
		Cpi = Ri * ((a[0 * 3 + 1] * _T + a[1 * 3 + 1]) * _T + a[2 * 3 + 1] + (a[3 * 3 + 1] + (a[4 * 3 + 1] + (a[5 * 3 + 1] + a[6 * 3 + 1] * T) * T) * T) * T);
	else if (T < 1000.0)
		Cpi = Ri * ((a[0 * 3 + 0] * _T + a[1 * 3 + 0]) * _T + a[2 * 3 + 0] + (a[3 * 3 + 0] + (a[4 * 3 + 0] + (a[5 * 3 + 0] + a[6 * 3 + 0] * T) * T) * T) * T);
	else if (T >= 6000.0)
		Cpi = Ri * ((a[0 * 3 + 2] * _T + a[1 * 3 + 2]) * _T + a[2 * 3 + 2] + (a[3 * 3 + 2] + (a[4 * 3 + 2] + (a[5 * 3 + 2] + a[6 * 3 + 2] * T) * T) * T) * T);
	return Cpi;
}

// ---- solver_Ini/Thermo_device.h:62-80 ----------------------------------------------------------
static inline double get_Enthalpy_NASA(const double *Hia, const double *Hib, const double T0, const double Ri, const int n)
{
	double hi = 0.0, TT = T0, T = std::fmax(T0, 200.0);
	const double *a = Hia + n * 7 * 3, *b = Hib + n * 2 * 3;
	if (T >= 1000.0 && T < 6000.0)
		hi = Ri * (-a[0 * 3 + 1] / T + a[1 * 3 + 1] * XO_LOG(T) + (a[2 * 3 + 1] + (0.5 * a[3 * 3 + 1] + (a[4 * 3 + 1] * _OT + (0.25 * a[5 * 3 + 1] + 0.2 * a[6 * 3 + 1] * T) * T) * T) * T) * T + b[0 * 3 + 1]);
	else if (T < 1000.0)
		hi = Ri * (-a[0 * 3 + 0] / T + a[1 * 3 + 0] * XO_LOG(T) + (a[2 * 3 + 0] + (0.5 * a[3 * 3 + 0] + (a[4 * 3 + 0] * _OT + (0.25 * a[5 * 3 + 0] + 0.2 * a[6 * 3 + 0] * T) * T) * T) * T) * T + b[0 * 3 + 0]);
	else if (T >= 6000.0)
		hi = Ri * (-a[0 * 3 + 2] / T + a[1 * 3 + 2] * XO_LOG(T) + (a[2 * 3 + 2] + (0.5 * a[3 * 3 + 2] + (a[4 * 3 + 2] * _OT + (0.25 * a[5 * 3 + 2] + 0.2 * a[6 * 3 + 2] * T) * T) * T) * T) * T + b[0 * 3 + 2]);
	if (TT < 200.0)
	{ // linear extension below 200 K
		double Cpi = HeatCapacity_NASA(Hia, 200.0, Ri, n);
		hi += Cpi * (TT - 200.0);
	}
	return hi;
}

// ---- solver_Ini/Mixing_device.h:49-140 ---------------------------------------------------------
static inline double get_CopR(const xo_cfg &t, const double *yi)
{
	double R = 0.0;
	for (int n = 0; n < t.NS; n++)
		R += yi[n] * t._Wi[n];
	return R * Ru;
}
static inline double get_CopCp(const xo_cfg &t, const double *yi, const double T)
{
	double _CopCp = 0.0;
	for (int ii = 0; ii < t.NS; ii++)
		_CopCp += yi[ii] * HeatCapacity_NASA(t.Hia, T, t.Ri[ii], ii);
	return _CopCp;
}
static inline double get_CopCv(const xo_cfg &t, const double *yi, const double T)
{
	double _CopCv = 0.0;
	for (int ii = 0; ii < t.NS; ii++)
		_CopCv += yi[ii] * HeatCapacity_NASA(t.Hia, T, t.Ri[ii], ii);
	double _W = 0.0;
	for (int ii = 0; ii < t.NS; ii++)
		_W += yi[ii] * t._Wi[ii];
	_CopCv -= Ru * _W;
	return _CopCv;
}
static inline double get_CopW(const xo_cfg &t, const double *yi)
{
	double _W = 0.0;
	for (int ii = 0; ii < t.NS; ii++)
		_W += yi[ii] * t._Wi[ii];
	return 1.0 / _W;
}
// 3-argument overload, Mixing_device.h:102-109
static inline double get_CopGamma3(const xo_cfg &t, const double *yi, const double T)
{
	double Cp = get_CopCp(t, yi, T);
	double Cv = get_CopCv(t, yi, T);
	return Cp / Cv;
}
// 4-argument overload, Mixing_device.h:114-126 (returns -1 when gamma <= 1)
static inline double get_CopGamma4(const xo_cfg &t, const double *yi, const double Cp, const double T)
{
	double CopW = get_CopW(t, yi);
	double _CopGamma = Cp / (Cp - Ru / CopW);
	if (_CopGamma > 1.0)
		return _CopGamma;
	else
		return -1;
}
static inline double get_Coph(const xo_cfg &t, const double *yi, const double T)
{
	double h = 0.0;
	for (int i = 0; i < t.NS; i++)
	{
		double hi = get_Enthalpy_NASA(t.Hia, t.Hib, T, t.Ri[i], i);
		h += hi * yi[i];
	}
	return h;
}
// Mixing_device.h:159-193 : limited Newton iteration for T(e, Y)
static int g_newton_iters = 0, g_newton_calls = 0; // instrumentation for k_it
static inline double get_T(const xo_cfg &t, const double *yi, const double e, const double T0)
{
	double T = T0;
	double tol = 1.0e-6;
	double func_T = 0.0, dfunc_T = 0.0;
	g_newton_calls++;
	for (int i = 1; i < 101; i++)
	{
		g_newton_iters++;
		double h = get_Coph(t, yi, T);
		double R = get_CopR(t, yi);
		double Cp = get_CopCp(t, yi, T);
		func_T = h - R * T - e;
		dfunc_T = Cp - R;
		double df = std::fmin(func_T / (dfunc_T + 1.0e-30), 1e-3 * T);
		df = std::fmax(df, -1e-2 * T);
		T = T - df;
		if (std::fabs(df) <= tol)
			break;
	}
	return T;
}

#define XO_ID(i, j, k) (size_t(c.Xmax) * c.Ymax * (k) + size_t(c.Xmax) * (j) + (i))

// ---- solver_UpdateStates/Update_device.hpp:7-29 + Update_kernels.hpp:5-18 (K1) ----------------
static void Getrhoyi(const xo_cfg &c, double *UI, double &rho, double *yi)
{
	rho = UI[0];
	double rho1 = 1.0 / rho;
	if (c.cop)
	{
		const int NUM_COP = c.NCOP;
		if (c.ghost_species)
		{
			yi[NUM_COP] = 0.0;
			double sum_yi = 0.0;
			for (int ii = 0; ii < NUM_COP; ii++)
				yi[ii] = UI[ii + 5] * rho1, sum_yi += yi[ii];
			sum_yi = 1.0 / sum_yi;
			for (int ii = 0; ii < NUM_COP; ii++)
				yi[ii] *= sum_yi, UI[ii + 5] = rho * yi[ii];
		}
		else
		{
			yi[NUM_COP] = 1.0;
			for (int ii = 5; ii < c.Emax; ii++)
				yi[ii - 5] = UI[ii] * rho1, yi[NUM_COP] += -yi[ii - 5];
		}
	}
}

// ---- Update_device.hpp:33-54,81-110 + Update_kernels.hpp:20-48 (K3) --------------------------
static void UpdateFluidStates(const xo_cfg &c, xo_state &s, double *UI)
{
	const int E = c.Emax, NS = c.NS;
	const size_t N = size_t(c.Xmax) * c.Ymax * c.Zmax;
	for (size_t id = 0; id < N; id++)
		Getrhoyi(c, UI + E * id, s.rho[id], s.y + NS * id);
	for (size_t id = 0; id < N; id++)
	{
		double *U = UI + E * id, *yi = s.y + NS * id;
		double rho = s.rho[id];
		double rho1 = 1.0 / rho;
		double u = U[1] * rho1, v = U[2] * rho1, w = U[3] * rho1;
		double tme = U[4] * rho1 - 0.5 * (u * u + v * v + w * w);
		double gamma, p, T = s.T[id], Cp = s.Cp[id], R = s.Ri[id];
		if (c.cop)
		{
			double R_ = get_CopR(c, yi);
			T = get_T(c, yi, tme, T);
			p = rho * R_ * T, R = R_;
			Cp = get_CopCp(c, yi, T);
			gamma = get_CopGamma4(c, yi, Cp, T);
		}
		else
		{
			gamma = c.ncop_gamma;
			p = (c.ncop_gamma - 1.0) * rho * tme;
		}
		double H = (U[4] + p) * rho1;
		double cc = std::sqrt(gamma * p * rho1);
		s.u[id] = u, s.v[id] = v, s.w[id] = w, s.p[id] = p, s.H[id] = H, s.c[id] = cc;
		s.gamma[id] = gamma, s.T[id] = T, s.e[id] = tme, s.Cp[id] = Cp, s.Ri[id] = R;

		double *Fx = s.FluxF + E * id, *Fy = s.FluxG + E * id, *Fz = s.FluxH + E * id;
		Fx[0] = U[1], Fx[1] = U[1] * u + p, Fx[2] = U[1] * v, Fx[3] = U[1] * w, Fx[4] = (U[4] + p) * u;
		Fy[0] = U[2], Fy[1] = U[2] * u, Fy[2] = U[2] * v + p, Fy[3] = U[2] * w, Fy[4] = (U[4] + p) * v;
		Fz[0] = U[3], Fz[1] = U[3] * u, Fz[2] = U[3] * v, Fz[3] = U[3] * w + p, Fz[4] = (U[4] + p) * w;
		if (c.cop)
			for (int ii = 5; ii < E; ii++)
				Fx[ii] = U[1] * yi[ii - 5], Fy[ii] = U[2] * yi[ii - 5], Fz[ii] = U[3] * yi[ii - 5];
	}
}

// ---- guards: Estimate_kernels.hpp:5-162 (flags only, default build patches nothing) -----------
static inline bool bad(double x) { return (x < 0) || std::isnan(x) || std::isinf(x); }
static void EstimateGuards(const xo_cfg &c, xo_state &s, bool prim)
{
	for (int k = c.Bw_Z; k < c.Zmax - c.Bw_Z; k++)
		for (int j = c.Bw_Y; j < c.Ymax - c.Bw_Y; j++)
			for (int i = c.Bw_X; i < c.Xmax - c.Bw_X; i++)
			{
				size_t id = XO_ID(i, j, k);
				if (!prim)
				{ // EstimateYiKernel: rho <0/NaN/Inf, or any yi NaN/Inf
					bool e = bad(s.rho[id]);
					if (c.cop)
						for (int n = 0; n < c.NS; n++)
							e = e || std::isnan(s.y[c.NS * id + n]) || std::isinf(s.y[c.NS * id + n]);
					if (e)
						s.error_flags[0] = 1;
				}
				else if (bad(s.rho[id]) || bad(s.p[id]) || bad(s.T[id])) // EstimatePrimitiveVarKernel
					s.error_flags[1] = 1;
			}
}
// Fluids.cpp:47-87 EstimateFluidNANKernel
static void EstimateFluidNAN(const xo_cfg &c, xo_state &s, const double *UI)
{
	for (int k = c.Bw_Z; k < c.Zmax - c.Bw_Z; k++)
		for (int j = c.Bw_Y; j < c.Ymax - c.Bw_Y; j++)
			for (int i = c.Bw_X; i < c.Xmax - c.Bw_X; i++)
			{
				size_t id = XO_ID(i, j, k) * c.Emax;
				bool e = UI[id] < 0;
				for (int n = 0; n < c.Emax; n++)
					e = e || std::isnan(UI[id + n]) || std::isinf(UI[id + n]) || std::isnan(s.LU[id + n]) || std::isinf(s.LU[id + n]);
				if (e)
					s.error_flags[2] = 1;
			}
}

// ---- FDM_Method/positive-definite_eigen/Eigen_value.hpp:5-37 (K5) -----------------------------
static void GetLocalEigen(const xo_cfg &c, const xo_state &s, double AA, double BB, double CC, double *eigen_local)
{
	const int E = c.Emax;
	const size_t N = size_t(c.Xmax) * c.Ymax * c.Zmax;
	for (size_t id = 0; id < N; id++)
	{
		if (c.weno <= 6)
		{
			double uu = AA * s.u[id] + BB * s.v[id] + CC * s.w[id];
			double uuPc = uu + s.c[id];
			double uuMc = uu - s.c[id];
			eigen_local[E * id + 0] = uuMc;
			for (int ii = 1; ii < E - 1; ii++)
				eigen_local[E * id + ii] = uu;
			eigen_local[E * id + E - 1] = uuPc;
		}
		else
			for (int ii = 0; ii < E; ii++)
				eigen_local[E * id + ii] = 0.0;
	}
}
// ConVenction_block.hpp:115-170 (K6): running max, combined into the CURRENT value
static void GlobalEigenMax(const xo_cfg &c, const double *eigen_local, double *eigen_block)
{
	const int E = c.Emax;
	const size_t N = size_t(c.Xmax) * c.Ymax * c.Zmax;
	for (int nn = 0; nn < E; nn++)
		for (size_t id = 0; id < N; id++)
			eigen_block[nn] = std::fmax(eigen_block[nn], std::fabs(eigen_local[E * id + nn]));
}

// ---- Recon_device.hpp get_RoeAverage -----------------------------------------------------------
static inline double get_RoeAverage(const double left, const double right, const double D, const double D1)
{
	return (left + D * right) * D1;
}
// ---- positive-definite_eigen/Utils_device.hpp:14-35 -------------------------------------------
static inline double get_DpDrho(const double hN, const double RN, const double q2, const double Cp, const double R, const double T, const double e, const double gamma)
{
	double RNT = RN * T;
	return (gamma - 1.0) * (0.5 * q2 - hN + Cp * RNT / R);
}
static inline double get_DpDrhoi(const double hin, const double Rin, const double hiN, const double RiN, const double T, const double Cp, const double R, const double gamma)
{
	double hN_minus_hi = -hin + hiN;
	double Ri_minus_RN = (Rin - RiN);
	double temp = (gamma - 1.0) * (hN_minus_hi + Cp * Ri_minus_RN * T / R);
	return temp;
}
// ---- Utils_device.hpp:42-79 ---------------------------------------------------------------------
static inline double SoundSpeedMultiSpecies(const int NUM_COP, double *zi, double &b1, double &b3, double *_Yi, double *_dpdrhoi, double *drhoi, const double _dpdrho, const double _dpde,
											const double _dpdE, const double _prho, const double dp, const double drho, const double de, const double _rho)
{
	double _dpdrhoi_new[XO_MAXS], Sum_dpdrhoi = 0.0, Sum_drhoi = 0.0, Sum_dpdrhoi2 = 0.0, Sum_Yidpdrhoi = 0.0;
	for (int n = 0; n < NUM_COP; n++)
	{
		Sum_dpdrhoi += _dpdrhoi[n] * drhoi[n];
		Sum_dpdrhoi2 += _dpdrhoi[n] * drhoi[n] * _dpdrhoi[n] * drhoi[n];
	}
	double temp1 = dp - (_dpdrho * drho + _dpde * de + Sum_dpdrhoi);
	double temp = temp1 / (_dpdrho * _dpdrho * drho * drho + _dpde * de * _dpde * de + Sum_dpdrhoi2 + 1e-19);
	for (int n = 0; n < NUM_COP; n++)
	{
		Sum_drhoi += drhoi[n] * drhoi[n];
		Sum_Yidpdrhoi += _Yi[n] * _dpdrhoi[n];
	}
	double _dpdE_new = _dpdE + _dpdE * _dpdE * de * _rho * temp;
	double _dpdrho_new = _dpdrho + _dpdrho * _dpdrho * drho * temp;
	for (int n = 0; n < NUM_COP; n++)
		_dpdrhoi_new[n] = _dpdrhoi[n] + _dpdrhoi[n] * _dpdrhoi[n] * drhoi[n] * temp;
	double csqr = _dpdrho_new + _dpdE_new * _prho + Sum_Yidpdrhoi;
	b1 = _dpdE_new / csqr;
	for (int n = 0; n < NUM_COP; n++)
	{
		zi[n] = -_dpdrhoi_new[n] / _dpdE_new;
		b3 += _Yi[n] * zi[n];
	}
	b3 *= b1;
	return csqr;
}
// ---- Utils_device.hpp:102-140 -------------------------------------------------------------------
static inline double ReconstructSoundSpeed(const xo_cfg &c, const xo_state &s, size_t id_l, size_t id_r, const double D, const double D1, const double _rho, const double _P,
										   double *_yi, double *z, double &b1, double &b3, double &Gamma)
{
	const int NS = c.NS, NUM_COP = c.NCOP;
	const double *rho = s.rho, *u = s.u, *v = s.v, *w = s.w, *p = s.p, *T = s.T, *H = s.H;
	const double *yi_l = s.y + NS * id_l, *yi_r = s.y + NS * id_r;
	double hi_l[XO_MAXS], hi_r[XO_MAXS], _dpdrhoi[XO_MAXS], drhoi[XO_MAXS];
	for (int n = 0; n < NS; n++)
	{
		hi_l[n] = get_Enthalpy_NASA(c.Hia, c.Hib, T[id_l], c.Ri[n], n);
		hi_r[n] = get_Enthalpy_NASA(c.Hia, c.Hib, T[id_r], c.Ri[n], n);
		_yi[n] = (yi_l[n] + D * yi_r[n]) * D1;
	}
	double gamma_l = get_CopGamma3(c, yi_l, T[id_l]);
	double gamma_r = get_CopGamma3(c, yi_r, T[id_r]);
	Gamma = get_RoeAverage(gamma_l, gamma_r, D, D1);
	double q2_l = u[id_l] * u[id_l] + v[id_l] * v[id_l] + w[id_l] * w[id_l];
	double q2_r = u[id_r] * u[id_r] + v[id_r] * v[id_r] + w[id_r] * w[id_r];
	double R_l = get_CopR(c, yi_l), R_r = get_CopR(c, yi_r);
	double Cp_l = get_CopCp(c, yi_l, T[id_l]), Cp_r = get_CopCp(c, yi_r, T[id_r]);
	double e_l = H[id_l] - 0.5 * q2_l - p[id_l] / rho[id_l], e_r = H[id_r] - 0.5 * q2_r - p[id_r] / rho[id_r];
	double _dpdrho = get_RoeAverage(get_DpDrho(hi_l[NUM_COP], c.Ri[NUM_COP], q2_l, Cp_l, R_l, T[id_l], e_l, gamma_l),
									get_DpDrho(hi_r[NUM_COP], c.Ri[NUM_COP], q2_r, Cp_r, R_r, T[id_r], e_r, gamma_r), D, D1);
	for (int nn = 0; nn < NUM_COP; nn++)
	{
		_dpdrhoi[nn] = get_RoeAverage(get_DpDrhoi(hi_l[nn], c.Ri[nn], hi_l[NUM_COP], c.Ri[NUM_COP], T[id_l], Cp_l, R_l, gamma_l),
									  get_DpDrhoi(hi_r[nn], c.Ri[nn], hi_r[NUM_COP], c.Ri[NUM_COP], T[id_r], Cp_r, R_r, gamma_r), D, D1);
		drhoi[nn] = rho[id_r] * yi_r[nn] - rho[id_l] * yi_l[nn];
	}
	double _prho = get_RoeAverage(p[id_l] / rho[id_l], p[id_r] / rho[id_r], D, D1) + 0.5 * D * D1 * D1 * ((u[id_r] - u[id_l]) * (u[id_r] - u[id_l]) + (v[id_r] - v[id_l]) * (v[id_r] - v[id_l]) + (w[id_r] - w[id_l]) * (w[id_r] - w[id_l]));
	double _dpdE = get_RoeAverage(gamma_l - 1.0, gamma_r - 1.0, D, D1);
	double _dpde = get_RoeAverage((gamma_l - 1.0) * rho[id_l], (gamma_r - 1.0) * rho[id_r], D, D1);
	double c2 = SoundSpeedMultiSpecies(NUM_COP, z, b1, b3, _yi, _dpdrhoi, drhoi, _dpdrho, _dpde, _dpdE, _prho, p[id_r] - p[id_l], rho[id_r] - rho[id_l], e_r - e_l, _rho);
	double c2w = (0.0 < c2) ? 0.0 : 1.0; // sycl::step(c2, 0.0): 0 while c2 > 0, 1 while c2 <= 0
	c2 = Gamma * _P * _rho * c2w + (1.0 - c2w) * c2;
	return c2;
}

// ---- positive-definite_eigen/Eigen_matrix.hpp:7-455: one row of L / one column of R -----------
// dir: 0 x, 1 y, 2 z.  Written out per direction because the row order and signs differ.
struct RoeState
{
	double _u, _v, _w, _H, c2, b1, b3;
	const double *z, *yi;
};
static void RoeAverageLeft(const xo_cfg &c, int dir, int n, double *eigen_l, double &eigen_value, const RoeState &r)
{
	const int E = c.Emax, NUM_COP = c.NCOP, NS = c.NS;
	const double _u = r._u, _v = r._v, _w = r._w, _H = r._H, b1 = r.b1, b3 = r.b3;
	const double *z = r.z, *yi = r.yi;
	// MARCO_PREEIGEN, Eigen_callback.h:102-106
	double q2 = _u * _u + _v * _v + _w * _w;
	double _c = std::sqrt(r.c2);
	double b2 = 1.0 + b1 * q2 - b1 * _H;
	double _c1 = 1.0 / _c;
	const double un = dir == 0 ? _u : (dir == 1 ? _v : _w); // normal velocity
	double _un_c = un * _c1;
	const int nent = dir + 1; // index of the "entropy" row: x->1, y->2, z->3
	if (0 == n)
	{
		eigen_l[0] = 0.5 * (b2 + _un_c + b3);
		eigen_l[1] = dir == 0 ? -0.5 * (b1 * _u + _c1) : -0.5 * (b1 * _u);
		eigen_l[2] = dir == 1 ? -0.5 * (b1 * _v + _c1) : -0.5 * (b1 * _v);
		eigen_l[3] = dir == 2 ? -0.5 * (b1 * _w + _c1) : -0.5 * (b1 * _w);
		eigen_l[4] = 0.5 * b1;
		eigen_value = std::fabs(un - _c);
		for (int m = 0; m < NUM_COP; m++)
			eigen_l[m + E - NUM_COP] = -0.5 * b1 * z[m];
	}
	else if (nent == n)
	{
		eigen_l[0] = (1.0 - b2 - b3) / b1;
		eigen_l[1] = _u;
		eigen_l[2] = _v;
		eigen_l[3] = _w;
		eigen_l[4] = -1.0;
		eigen_value = std::fabs(un);
		for (int m = 0; m < NUM_COP; m++)
			eigen_l[m + E - NUM_COP] = z[m];
	}
	else if (n >= 1 && n <= 3)
	{ // the two shear rows (positions/signs per direction, Eigen_matrix.hpp:38-59,177-209,327-348)
		double e0 = 0, e1 = 0, e2 = 0, e3 = 0;
		if (dir == 0)
		{
			if (n == 2)
				e0 = _v, e2 = -1.0;
			else
				e0 = -_w, e3 = 1.0; // n == 3
		}
		else if (dir == 1)
		{
			if (n == 1)
				e0 = -_u, e1 = 1.0;
			else
				e0 = _w, e3 = -1.0; // n == 3
		}
		else
		{
			if (n == 1)
				e0 = _u, e1 = -1.0;
			else
				e0 = -_v, e2 = 1.0; // n == 2
		}
		eigen_l[0] = e0, eigen_l[1] = e1, eigen_l[2] = e2, eigen_l[3] = e3, eigen_l[4] = 0.0;
		eigen_value = std::fabs(un);
		for (int m = 0; m < NUM_COP; m++)
			eigen_l[m + E - NUM_COP] = 0.0;
	}
	else if (E - 1 == n)
	{
		eigen_l[0] = 0.5 * (b2 - _un_c + b3);
		eigen_l[1] = dir == 0 ? 0.5 * (-b1 * _u + _c1) : 0.5 * (-b1 * _u);
		eigen_l[2] = dir == 1 ? 0.5 * (-b1 * _v + _c1) : 0.5 * (-b1 * _v);
		eigen_l[3] = dir == 2 ? 0.5 * (-b1 * _w + _c1) : 0.5 * (-b1 * _w);
		eigen_l[4] = 0.5 * b1;
		eigen_value = std::fabs(un + _c);
		for (int m = 0; m < NUM_COP; m++)
			eigen_l[m + E - NUM_COP] = -0.5 * b1 * z[m];
	}
	else
	{ // species rows (COP only)
		eigen_l[0] = -yi[n + NS - E];
		eigen_l[1] = 0.0, eigen_l[2] = 0.0, eigen_l[3] = 0.0, eigen_l[4] = 0.0;
		eigen_value = std::fabs(un);
		for (int m = 0; m < NUM_COP; m++)
			eigen_l[m + E - NUM_COP] = (n + NS - E == m) ? 1.0 : 0.0;
	}
}
static void RoeAverageRight(const xo_cfg &c, int dir, int n, double *eigen_r, const RoeState &r)
{
	const int E = c.Emax, NUM_COP = c.NCOP, NS = c.NS;
	const double _u = r._u, _v = r._v, _w = r._w, _H = r._H, b1 = r.b1;
	const double *z = r.z, *yi = r.yi;
	double _c = std::sqrt(r.c2);
	const double un = dir == 0 ? _u : (dir == 1 ? _v : _w);
	const int nent = dir + 1;
	if (0 == n)
	{
		eigen_r[0] = 1.0;
		eigen_r[1] = dir == 0 ? _u - _c : _u;
		eigen_r[2] = dir == 1 ? _v - _c : _v;
		eigen_r[3] = dir == 2 ? _w - _c : _w;
		eigen_r[4] = _H - un * _c;
		for (int m = 0; m < NUM_COP; m++)
			eigen_r[m + E - NUM_COP] = yi[m];
	}
	else if (nent == n)
	{
		eigen_r[0] = b1;
		eigen_r[1] = dir == 2 ? b1 * _u : _u * b1; // commutative; kept as written
		eigen_r[2] = dir == 2 ? b1 * _v : _v * b1;
		eigen_r[3] = dir == 2 ? b1 * _w : _w * b1;
		eigen_r[4] = _H * b1 - 1.0;
		for (int m = 0; m < NUM_COP; m++)
			eigen_r[m + E - NUM_COP] = b1 * yi[m];
	}
	else if (n >= 1 && n <= 3)
	{
		double e1 = 0, e2 = 0, e3 = 0, e4 = 0;
		if (dir == 0)
		{
			if (n == 2)
				e2 = -1.0, e4 = -_v;
			else
				e3 = 1.0, e4 = _w;
		}
		else if (dir == 1)
		{
			if (n == 1)
				e1 = 1.0, e4 = _u;
			else
				e3 = -1.0, e4 = -_w;
		}
		else
		{
			if (n == 1)
				e1 = -1.0, e4 = -_u;
			else
				e2 = 1.0, e4 = _v;
		}
		eigen_r[0] = 0.0, eigen_r[1] = e1, eigen_r[2] = e2, eigen_r[3] = e3, eigen_r[4] = e4;
		for (int m = 0; m < NUM_COP; m++)
			eigen_r[m + E - NUM_COP] = 0.0;
	}
	else if (E - 1 == n)
	{
		eigen_r[0] = 1.0;
		eigen_r[1] = dir == 0 ? _u + _c : _u;
		eigen_r[2] = dir == 1 ? _v + _c : _v;
		eigen_r[3] = dir == 2 ? _w + _c : _w;
		eigen_r[4] = _H + un * _c;
		for (int m = 0; m < NUM_COP; m++)
			eigen_r[m + E - NUM_COP] = yi[m];
	}
	else
	{
		eigen_r[0] = 0.0, eigen_r[1] = 0.0, eigen_r[2] = 0.0, eigen_r[3] = 0.0;
		eigen_r[4] = z[n + NS - E];
		for (int m = 0; m < NUM_COP; m++)
			eigen_r[m + E - NUM_COP] = (m == n + NS - E) ? 1.0 : 0.0;
	}
}

// ---- schemes/WENO5s_schemes.hpp:12-95 ----------------------------------------------------------
static inline double weno5old_BODY(const double v1, const double v2, const double v3, const double v4, const double v5)
{
	double a1, a2, a3;
	double dtwo = 2.0, dtre = 3.0;
	a1 = v1 - dtwo * v2 + v3;
	double s1 = 13.0 * a1 * a1;
	a1 = v1 - 4.0 * v2 + dtre * v3;
	s1 += dtre * a1 * a1;
	a1 = v2 - dtwo * v3 + v4;
	double s2 = 13.0 * a1 * a1;
	a1 = v2 - v4;
	s2 += dtre * a1 * a1;
	a1 = v3 - dtwo * v4 + v5;
	double s3 = 13.0 * a1 * a1;
	a1 = dtre * v3 - 4.0 * v4 + v5;
	s3 += dtre * a1 * a1;
	double tol = 1.0E-6;
	s1 += tol, s2 += tol, s3 += tol;
	a1 = 0.1 * s2 * s2 * s3 * s3;
	a2 = 0.6 * s1 * s1 * s3 * s3;
	a3 = 0.3 * s1 * s1 * s2 * s2;
	double tw1 = 1.0 / (a1 + a2 + a3);
	a1 = a1 * tw1, a2 = a2 * tw1, a3 = a3 * tw1;
	s1 = a1 * (dtwo * v1 - 7.0 * v2 + 11.0 * v3);
	s2 = a2 * (-v2 + 5.0 * v3 + dtwo * v4);
	s3 = a3 * (dtwo * v3 + 5.0 * v4 - v5);
	return (s1 + s2 + s3);
}
static inline double weno5old_GPU(const double *f, const double *m)
{
	double temf = weno5old_BODY(f[-2], f[-1], f[0], f[1], f[2]);
	double temm = weno5old_BODY(m[3], m[2], m[1], m[0], m[-1]);
	return (temf + temm) * _six;
}
// ---- schemes/WENO6s_schemes.hpp:5-78 (WENO-CU6, SCHEME_ORDER == 6); constants Utils_schemes.hpp:5-15, global_setup.h:36-37
static const double _sxtn = 1.0 / 16.0, _twfr = 1.0 / 24.0, _ohtz = 1.0 / 120.0, _ohff = 1.0 / 144.0, _ftss = 1.0 / 5760.0;
static const double wu6a2a2 = 13.0 / 3.0, wu6a3a3 = 3129.0 / 80.0, wu6a4a4 = 87617.0 / 140.0, wu6a3a5 = 14127.0 / 224.0, wu6a5a5 = 252337135.0 / 16128.0;
static inline double WENOCU6_BODY(const double v1, const double v2, const double v3, const double v4, const double v5, const double v6, const double epsilon)
{
	double s11 = v1 - 2.0 * v2 + v3;
	double s12 = v1 - 4.0 * v2 + 3.0 * v3;
	double s1 = 13.0 * s11 * s11 + 3.0 * s12 * s12;
	double s21 = v2 - 2.0 * v3 + v4;
	double s22 = v2 - v4;
	double s2 = 13.0 * s21 * s21 + 3.0 * s22 * s22;
	double s31 = v3 - 2.0 * v4 + v5;
	double s32 = 3.0 * v3 - 4.0 * v4 + v5;
	double s3 = 13.0 * s31 * s31 + 3.0 * s32 * s32;
	double tau61 = (259.0 * v6 - 1895.0 * v5 + 6670.0 * v4 - 2590.0 * v3 - 2785.0 * v2 + 341.0 * v1) * _ftss;
	double tau62 = -(v5 - 12.0 * v4 + 22.0 * v3 - 12.0 * v2 + v1) * _sxtn;
	double tau63 = -(7.0 * v6 - 47.0 * v5 + 94.0 * v4 - 70.0 * v3 + 11.0 * v2 + 5.0 * v1) * _ohff;
	double tau64 = (v5 - 4.0 * v4 + 6.0 * v3 - 4.0 * v2 + v1) * _twfr;
	double tau65 = -(-v6 + 5.0 * v5 - 10.0 * v4 + 10.0 * v3 - 5.0 * v2 + v1) * _ohtz;
	double a1a1 = 1.0, a2a2 = wu6a2a2, a1a3 = 0.5, a3a3 = wu6a3a3, a2a4 = 4.2;
	double a1a5 = 0.125, a4a4 = wu6a4a4, a3a5 = wu6a3a5, a5a5 = wu6a5a5;
	double s6 = (tau61 * tau61 * a1a1 + tau62 * tau62 * a2a2 + tau61 * tau63 * a1a3 + tau63 * tau63 * a3a3 + tau62 * tau64 * a2a4 + tau61 * tau65 * a1a5 + tau64 * tau64 * a4a4 + tau63 * tau65 * a3a5 + tau65 * tau65 * a5a5) * 12.0;
	double s55 = (s1 + s3 + 4.0 * s2) * _six;
	double s5 = std::fabs(s6 - s55);
	double r0 = 20.0;
	double r1 = r0 + s5 / (s1 + epsilon);
	double r2 = r0 + s5 / (s2 + epsilon);
	double r3 = r0 + s5 / (s3 + epsilon);
	double r4 = r0 + s5 / (s6 + epsilon);
	double a1 = 0.05 * r1;
	double a2 = 0.45 * r2;
	double a3 = 0.45 * r3;
	double a4 = 0.05 * r4;
	double tw1 = 1.0 / (a1 + a2 + a3 + a4);
	double w1 = a1 * tw1;
	double w2 = a2 * tw1;
	double w3 = a3 * tw1;
	double w4 = a4 * tw1;
	double temp = 0.0;
	temp += w1 * (2.0 * v1 - 7.0 * v2 + 11.0 * v3);
	temp += w2 * (-v2 + 5.0 * v3 + 2.0 * v4);
	temp += w3 * (2.0 * v3 + 5.0 * v4 - v5);
	temp += w4 * (11.0 * v4 - 7.0 * v5 + 2.0 * v6);
	return temp;
}
static inline double WENOCU6_GPU(const double *f, const double *m, double delta)
{
	const double epsilon = 1.e-8 * delta * delta;
	double temf = WENOCU6_BODY(f[-2], f[-1], f[0], f[1], f[2], f[3], epsilon);
	double temm = WENOCU6_BODY(m[3], m[2], m[1], m[0], m[-1], m[-2], epsilon);
	return (temf + temm) * _six;
}
// ---- schemes/WENO7s_schemes.hpp:8-128 (the _P and _M bodies are identical after the v1..v7 pick)
static inline double weno7_BODY(const double v1, const double v2, const double v3, const double v4, const double v5, const double v6, const double v7)
{
	double ep = 1.0e-7;
	double C0 = 1.0 / 35.0, C1 = 12.0 / 35.0, C2 = 18.0 / 35.0, C3 = 4.0 / 35.0;
	double S10 = -2.0 / 6.0 * v1 + 9.0 / 6.0 * v2 - 18.0 / 6.0 * v3 + 11.0 / 6.0 * v4;
	double S11 = 1.0 / 6.0 * v2 - 6.0 / 6.0 * v3 + 3.0 / 6.0 * v4 + 2.0 / 6.0 * v5;
	double S12 = -2.0 / 6.0 * v3 - 3.0 / 6.0 * v4 + 6.0 / 6.0 * v5 - 1.0 / 6.0 * v6;
	double S13 = -11.0 / 6.0 * v4 + 18.0 / 6.0 * v5 - 9.0 / 6.0 * v6 + 2.0 / 6.0 * v7;
	double S20 = -v1 + 4.0 * v2 - 5.0 * v3 + 2.0 * v4;
	double S21 = v3 - 2.0 * v4 + v5;
	double S22 = v4 - 2.0 * v5 + v6;
	double S23 = 2.0 * v4 - 5.0 * v5 + 4.0 * v6 - 1.0 * v7;
	double S30 = -v1 + 3.0 * v2 - 3.0 * v3 + v4;
	double S31 = -v2 + 3.0 * v3 - 3.0 * v4 + v5;
	double S32 = -v3 + 3.0 * v4 - 3.0 * v5 + v6;
	double S33 = -v4 + 3.0 * v5 - 3.0 * v6 + v7;
	double S0 = S10 * S10 + 13.0 / 12.0 * S20 * S20 + 1043.0 / 960.0 * S30 * S30 + 1.0 / 12.0 * S10 * S30;
	double S1 = S11 * S11 + 13.0 / 12.0 * S21 * S21 + 1043.0 / 960.0 * S31 * S31 + 1.0 / 12.0 * S11 * S31;
	double S2 = S12 * S12 + 13.0 / 12.0 * S22 * S22 + 1043.0 / 960.0 * S32 * S32 + 1.0 / 12.0 * S12 * S32;
	double S3 = S13 * S13 + 13.0 / 12.0 * S23 * S23 + 1043.0 / 960.0 * S33 * S33 + 1.0 / 12.0 * S13 * S33;
	double a0 = C0 / ((ep + S0) * (ep + S0));
	double a1 = C1 / ((ep + S1) * (ep + S1));
	double a2 = C2 / ((ep + S2) * (ep + S2));
	double a3 = C3 / ((ep + S3) * (ep + S3));
	double W0 = a0 / (a0 + a1 + a2 + a3);
	double W1 = a1 / (a0 + a1 + a2 + a3);
	double W2 = a2 / (a0 + a1 + a2 + a3);
	double W3 = a3 / (a0 + a1 + a2 + a3);
	double q0 = -3.0 / 12.0 * v1 + 13.0 / 12.0 * v2 - 23.0 / 12.0 * v3 + 25.0 / 12.0 * v4;
	double q1 = 1.0 / 12.0 * v2 - 5.0 / 12.0 * v3 + 13.0 / 12.0 * v4 + 3.0 / 12.0 * v5;
	double q2 = -1.0 / 12.0 * v3 + 7.0 / 12.0 * v4 + 7.0 / 12.0 * v5 - 1.0 / 12.0 * v6;
	double q3 = 3.0 / 12.0 * v4 + 13.0 / 12.0 * v5 - 5.0 / 12.0 * v6 + 1.0 / 12.0 * v7;
	return W0 * q0 + W1 * q1 + W2 * q2 + W3 * q3;
}
static inline double weno7_P(const double *f) { return weno7_BODY(f[-3], f[-2], f[-1], f[0], f[1], f[2], f[3]); }
static inline double weno7_M(const double *f) { return weno7_BODY(f[4], f[3], f[2], f[1], f[0], f[-1], f[-2]); }

// ---- Reconstruction_kernels.hpp:8-199 + Eigen_callback.h:127-230 + global_marco.h:26-34 (K7) --
static void ReconstructFlux(const xo_cfg &c, const xo_state &s, int dir, const double *UI, const double *Fl, double *Fwall, const double *eigen_local, const double *eigen_block)
{
	const int E = c.Emax;
	const double *rho = s.rho, *u = s.u, *v = s.v, *w = s.w, *H = s.H, *p = s.p;
	const int i0 = c.Bw_X - (dir == 0), j0 = c.Bw_Y - (dir == 1), k0 = c.Bw_Z - (dir == 2);
	const double Roe_type = c.alpha == 1 ? 1.0 : 0.0, LLF_type = c.alpha == 2 ? 1.0 : 0.0, GLF_type = c.alpha == 3 ? 1.0 : 0.0;
	const ptrdiff_t st = dir == 0 ? 1 : (dir == 1 ? c.Xmax : ptrdiff_t(c.Xmax) * c.Ymax); // cell stride along dir
	for (int k = k0; k < c.Z_inner + c.Bw_Z; k++)
		for (int j = j0; j < c.Y_inner + c.Bw_Y; j++)
			for (int i = i0; i < c.X_inner + c.Bw_X; i++)
			{
				size_t id_l = XO_ID(i, j, k);
				size_t id_r = id_l + st;
				// MARCO_ROE
				double D = std::sqrt(rho[id_r] / rho[id_l]);
				double D1 = 1.0 / (D + 1.0);
				double _u = (u[id_l] + D * u[id_r]) * D1;
				double _v = (v[id_l] + D * v[id_r]) * D1;
				double _w = (w[id_l] + D * w[id_r]) * D1;
				double _H = (H[id_l] + D * H[id_r]) * D1;
				double _P = (p[id_l] + D * p[id_r]) * D1;
				double _rho = std::sqrt(rho[id_r] * rho[id_l]);
				// MARCO_GETC2
				double _yi[XO_MAXS], z[XO_MAXS] = {0.0}, b1 = 0.0, b3 = 0.0, Gamma0 = 1.4, c2;
				if (c.cop)
					c2 = ReconstructSoundSpeed(c, s, id_l, id_r, D, D1, _rho, _P, _yi, z, b1, b3, Gamma0);
				else
				{ // MARCO_NOCOPC2, Eigen_callback.h:87-91
					_yi[0] = 1.0, b3 = 0.0, z[0] = 0.0;
					Gamma0 = c.ncop_gamma;
					c2 = Gamma0 * _P / _rho;
					b1 = (Gamma0 - 1.0) / c2;
				}
				RoeState rs{_u, _v, _w, _H, c2, b1, b3, z, _yi};

				double uf[10], ff[10], pp[10], mm[10], f_flux, _p[XO_MAXE][XO_MAXE], eigen_lr[XO_MAXE], eigen_value, artificial_viscosity;
				for (int n = 0; n < E; n++)
				{
					double eigen_local_max = 0.0;
					RoeAverageLeft(c, dir, n, eigen_lr, eigen_value, rs);
					if (c.weno == 7)
					{ // MARCO_FLUXWALL_WENO7, Eigen_callback.h:127-176
						eigen_local_max = eigen_value;
						double lambda_l = eigen_local[E * id_l + n];
						double lambda_r = eigen_local[E * id_r + n];
						if (lambda_l * lambda_r < 0.0)
							for (int m = -3; m < 8 - 3; m++)
								eigen_local_max = std::fmax(eigen_local_max, std::fabs(eigen_local[E * (id_l + m * st) + n]));
						artificial_viscosity = Roe_type * eigen_value + LLF_type * eigen_local_max + GLF_type * eigen_block[n];
						for (int m = 0; m < 8; m++)
						{
							size_t id_local_2 = id_l + (m - 3) * st;
							uf[m] = 0.0, ff[m] = 0.0;
							for (int n1 = 0; n1 < E; n1++)
							{
								uf[m] = uf[m] + UI[E * id_local_2 + n1] * eigen_lr[n1];
								ff[m] = ff[m] + Fl[E * id_local_2 + n1] * eigen_lr[n1];
							}
							pp[m] = 0.5 * (ff[m] + artificial_viscosity * uf[m]);
							mm[m] = 0.5 * (ff[m] - artificial_viscosity * uf[m]);
						}
						f_flux = weno7_P(&pp[3]) + weno7_M(&mm[3]);
					}
					else
					{ // MARCO_FLUXWALL_WENO5, Eigen_callback.h:179-230
						for (int m = -2; m < 6 - 2; m++)
							eigen_local_max = std::fmax(eigen_local_max, std::fabs(eigen_local[E * (id_l + m * st) + n]));
						artificial_viscosity = Roe_type * eigen_value + LLF_type * eigen_local_max + GLF_type * eigen_block[n];
						for (int m = -3; m <= 4; m++)
						{
							size_t id_local = id_l + m * st;
							uf[m + 3] = 0.0, ff[m + 3] = 0.0;
							for (int n1 = 0; n1 < E; n1++)
							{
								uf[m + 3] = uf[m + 3] + UI[E * id_local + n1] * eigen_lr[n1];
								ff[m + 3] = ff[m + 3] + Fl[E * id_local + n1] * eigen_lr[n1];
							}
							pp[m + 3] = 0.5 * (ff[m + 3] + artificial_viscosity * uf[m + 3]);
							mm[m + 3] = 0.5 * (ff[m + 3] - artificial_viscosity * uf[m + 3]);
						}
						// WENO_GPU (schemes_device.hpp:13-20): weno5old for SCHEME_ORDER 5, WENO-CU6 for 6 (dl = the sweep's mesh width,
						// ConVenction_block.hpp:233,267,301)
						f_flux = c.weno == 6 ? WENOCU6_GPU(&pp[3], &mm[3], dir == 0 ? c.dx : (dir == 1 ? c.dy : c.dz)) : weno5old_GPU(&pp[3], &mm[3]);
					}
					RoeAverageRight(c, dir, n, eigen_lr, rs);
					for (int n1 = 0; n1 < E; n1++)
						_p[n][n1] = f_flux * eigen_lr[n1];
				}
				for (int n = 0; n < E; n++)
				{
					double fluxl = 0.0;
					for (int n1 = 0; n1 < E; n1++)
						fluxl += _p[n1][n];
					Fwall[E * id_l + n] = fluxl;
				}
			}
}

// ---- FDM_Method/positive-definite_eigen/PositivityPreserving_kernels.hpp:5-76 (K8), launched over the inner cells
// (ConVenction_block.hpp:330-410): id_l = the inner cell, id_r = its +dir neighbour, so the face below the first inner cell
// is never limited.  Reproduced as written, including `FF[n]` (not FF[nn]) inside the species loop.
static void PositivityPreserving(const xo_cfg &c, int dir, const double *UI, const double *Fl, double *Fwall, const double lambda_0, const double lambda)
{
	const int E = c.Emax, NUM_COP = c.NCOP;
	const ptrdiff_t st = dir == 0 ? 1 : (dir == 1 ? c.Xmax : ptrdiff_t(c.Xmax) * c.Ymax);
	double epsilon[XO_MAXS + 2];
	epsilon[0] = 1.0e-13, epsilon[1] = 1.0e-13;
	for (int ii = 2; ii < c.NS + 2; ii++)
		epsilon[ii] = 0.0;
	for (int k = c.Bw_Z; k < c.Zmax - c.Bw_Z; k++)
		for (int j = c.Bw_Y; j < c.Ymax - c.Bw_Y; j++)
			for (int i = c.Bw_X; i < c.Xmax - c.Bw_X; i++)
			{
				size_t id_l = XO_ID(i, j, k) * E, id_r = (XO_ID(i, j, k) + st) * E;
				double rho_min, theta, theta_u, theta_p, F_LF[XO_MAXE], FF_LF[XO_MAXE], FF[XO_MAXE];
				const double *UU = &(UI[id_l]), *UP = &(UI[id_r]);
				for (int n = 0; n < E; n++)
				{
					F_LF[n] = 0.5 * (Fl[n + id_l] + Fl[n + id_r] + lambda_0 * (UI[n + id_l] - UI[n + id_r]));
					FF_LF[n] = 2.0 * lambda * F_LF[n];
					FF[n] = 2.0 * lambda * Fwall[n + id_l];
				}
				theta_u = 1.0, theta_p = 1.0;
				rho_min = std::fmin(UU[0], epsilon[0]);
				if (UU[0] - FF[0] < rho_min)
					theta_u = (UU[0] - FF_LF[0] - rho_min + 1.0e-40) / (FF[0] - FF_LF[0] + 1.0e-40);
				rho_min = std::fmin(UP[0], epsilon[0]);
				if (UP[0] + FF[0] < rho_min)
					theta_p = (UP[0] + FF_LF[0] - rho_min + 1.0e-40) / (FF_LF[0] - FF[0] + 1.0e-40);
				theta = std::fmin(std::fmax(std::fmin(theta_u, theta_p), 0.0), 1.0);
				for (int n = 0; n < E; n++)
				{
					FF[n] = (1.0 - theta) * FF_LF[n] + theta * FF[n];
					Fwall[n + id_l] = (1.0 - theta) * F_LF[n] + theta * Fwall[n + id_l];
				}
				double yi_q[XO_MAXS], yi_u[XO_MAXS], yi_qp[XO_MAXS], yi_up[XO_MAXS], _rhoq, _rhou, _rhoqp, _rhoup;
				_rhoq = 1.0 / (UU[0] - FF[0]), _rhou = 1.0 / (UU[0] - FF_LF[0]);
				_rhoqp = 1.0 / (UP[0] + FF[0]), _rhoup = 1.0 / (UP[0] + FF_LF[0]);
				yi_q[NUM_COP] = 1.0, yi_u[NUM_COP] = 1.0, yi_qp[NUM_COP] = 1.0, yi_up[NUM_COP] = 1.0;
				for (int n = 0; n < NUM_COP; n++)
				{
					int tid = n + 5;
					yi_q[n] = (UU[tid] - FF[tid]) * _rhoq, yi_q[NUM_COP] -= yi_q[n];
					yi_u[n] = (UU[tid] - FF_LF[tid]) * _rhou, yi_u[NUM_COP] -= yi_u[n];
					yi_qp[n] = (UP[tid] + FF[tid]) * _rhoqp, yi_qp[NUM_COP] -= yi_qp[n];
					yi_up[n] = (UP[tid] + FF_LF[tid]) * _rhoup, yi_up[NUM_COP] -= yi_up[n];
					theta_u = 1.0, theta_p = 1.0;
					double temp = epsilon[n + 2];
					if (yi_q[n] < temp)
					{
						double yi_min = std::fmin(yi_u[n], temp);
						theta_u = (yi_u[n] - yi_min + 1.0e-40) / (yi_u[n] - yi_q[n] + 1.0e-40);
					}
					if (yi_qp[n] < temp)
					{
						double yi_min = std::fmin(yi_up[n], temp);
						theta_p = (yi_up[n] - yi_min + 1.0e-40) / (yi_up[n] - yi_qp[n] + 1.0e-40);
					}
					theta = std::fmin(std::fmax(std::fmin(theta_u, theta_p), 0.0), 1.0);
					for (int nn = 0; nn < E; nn++)
					{
						FF[n] = (1.0 - theta) * FF_LF[n] + theta * FF[n];
						Fwall[nn + id_l] = (1.0 - theta) * F_LF[nn] + theta * Fwall[nn + id_l];
					}
				}
			}
}

// ---- Reconstruction_kernels.hpp:201-234 (K10) ---------------------------------------------------
static void UpdateFluidLU(const xo_cfg &c, xo_state &s)
{
	const int E = c.Emax;
	for (int k = c.Bw_Z; k < c.Zmax - c.Bw_Z; k++)
		for (int j = c.Bw_Y; j < c.Ymax - c.Bw_Y; j++)
			for (int i = c.Bw_X; i < c.Xmax - c.Bw_X; i++)
			{
				size_t id = XO_ID(i, j, k);
				size_t id_im = id - 1, id_jm = id - c.Xmax, id_km = id - size_t(c.Xmax) * c.Ymax;
				for (int n = 0; n < E; n++)
				{
					double LU0 = 0.0;
					if (c.DimX)
						LU0 += (s.FluxFw[E * id_im + n] - s.FluxFw[E * id + n]) * c._dx;
					if (c.DimY)
						LU0 += (s.FluxGw[E * id_jm + n] - s.FluxGw[E * id + n]) * c._dy;
					if (c.DimZ)
						LU0 += (s.FluxHw[E * id_km + n] - s.FluxHw[E * id + n]) * c._dz;
					s.LU[E * id + n] = LU0;
				}
			}
}

// ---- ConVenction_block.hpp:10-617 GetLU, inviscid branch ----------------------------------------
// ---- viscous block of GetLU (FDM_Method/ConVenction_block.hpp:424-575) ------------------------------------------------------------
// viscosity/Fourth_Order/Visc_Order_kernels.hpp:16-74: velocity derivatives at the cells [B-2, B+inner+2) of every active direction
static void GetCellCenterDerivative(const xo_cfg &c, xo_state &s)
{
	const double _twle = 1.0 / 12.0;
	const double tX = c.DimX, tY = c.DimY, tZ = c.DimZ;
	const size_t sx = 1, sy = c.Xmax, sz = size_t(c.Xmax) * c.Ymax;
	const int i0 = c.DimX ? c.Bw_X - 2 : 0, i1 = c.DimX ? c.Bw_X + c.X_inner + 2 : 1;
	const int j0 = c.DimY ? c.Bw_Y - 2 : 0, j1 = c.DimY ? c.Bw_Y + c.Y_inner + 2 : 1;
	const int k0 = c.DimZ ? c.Bw_Z - 2 : 0, k1 = c.DimZ ? c.Bw_Z + c.Z_inner + 2 : 1;
	for (int k = k0; k < k1; k++)
		for (int j = j0; j < j1; j++)
			for (int i = i0; i < i1; i++)
			{
				const size_t id = sz * k + sy * j + i;
				auto D = [&](const double *q, size_t st, double _dl) { return (8.0 * (q[id + st] - q[id - st]) - (q[id + 2 * st] - q[id - 2 * st])) * _dl * _twle; };
				double d[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}; // inactive directions: the reference multiplies out-of-range reads by DimX_t = 0
				if (c.DimX)
					d[0] = D(s.u, sx, c._dx) * tX, d[1] = D(s.v, sx, c._dx) * tX * tY, d[2] = D(s.w, sx, c._dx) * tX * tZ;
				if (c.DimY)
					d[3] = D(s.u, sy, c._dy) * tY * tX, d[4] = D(s.v, sy, c._dy) * tY, d[5] = D(s.w, sy, c._dy) * tY * tZ;
				if (c.DimZ)
					d[6] = D(s.u, sz, c._dz) * tZ * tX, d[7] = D(s.v, sz, c._dz) * tZ * tY, d[8] = D(s.w, sz, c._dz) * tZ;
				for (int m = 0; m < 9; m++)
					s.Vde[m][id] = d[m];
			}
}
// viscosity/Visc_kernels.hpp:7-217: ghost fill of the four derivatives a direction's wall flux reads
static void CenterDerivativeBC(const xo_cfg &c, xo_state &s, int dir)
{
	static const int sel[3][4] = {{3, 6, 4, 8}, {1, 7, 0, 8}, {2, 5, 0, 4}};
	const int Bw = dir == 0 ? c.Bw_X : (dir == 1 ? c.Bw_Y : c.Bw_Z), inner = dir == 0 ? c.X_inner : (dir == 1 ? c.Y_inner : c.Z_inner);
	const int nmax = dir == 0 ? c.Xmax : (dir == 1 ? c.Ymax : c.Zmax);
	const size_t sy = c.Xmax, sz = size_t(c.Xmax) * c.Ymax;
	auto idx = [&](int i, int j, int k) { return sz * k + sy * j + i; };
	for (int side = 0; side < 2; side++)
	{
		const int BC = c.bc[2 * dir + side];
		const int mirror_offset = side ? inner : 0, index_inner = side ? nmax - Bw - 1 : Bw, sign = side ? -1 : 1;
		for (int k = 0; k < (dir == 2 ? Bw : c.Zmax); k++)
			for (int j = 0; j < (dir == 1 ? Bw : c.Ymax); j++)
				for (int i = 0; i < (dir == 0 ? Bw : c.Xmax); i++)
				{
					int ii = i, jj = j, kk = k;
					int &g = dir == 0 ? ii : (dir == 1 ? jj : kk);
					if (side)
						g += nmax - Bw;
					const size_t id = idx(ii, jj, kk);
					auto at = [&](int t) { return dir == 0 ? idx(t, jj, kk) : (dir == 1 ? idx(ii, t, kk) : idx(ii, jj, t)); };
					size_t src;
					double sg = 1.0;
					switch (BC)
					{
					case 2: src = at(2 * (Bw + mirror_offset) - 1 - g); break;
					case 3: src = at(g + sign * inner); break;
					case 1: src = at(index_inner); break;
					case 4:
						if (dir == 2)
							continue; // the reference's mirror index (k - offset) is out of range (Visc_kernels.hpp:197): not reproduced
						src = at(2 * (Bw + mirror_offset) - 1 - g), sg = -1.0;
						break;
					case 0:
						for (int n = 0; n < 4; n++)
							s.Vde[sel[dir][n]][id] = 0.0;
						continue;
					default: continue;
					}
					for (int n = 0; n < 4; n++)
						s.Vde[sel[dir][n]][id] = sg < 0 ? -s.Vde[sel[dir][n]][src] : s.Vde[sel[dir][n]][src];
				}
	}
}
// Visc_device.h:10-101
static inline double FitValue(const double *f, const double T)
{
	double v = f[3];
	for (int i = 2; i >= 0; i--)
		v = v * XO_LOG(T) + f[i];
	return std::exp(v);
}
// Gettransport_coeff_aver (Visc_kernels.hpp:219-241) + Get_transport_coeff_aver (Visc_device.h:107-177), all cells
static void GetTransportCoeff(const xo_cfg &c, xo_state &s)
{
	const int NS = c.NS;
	const size_t N = size_t(c.Xmax) * c.Ymax * c.Zmax;
	std::vector<double> X(NS);
	for (size_t id = 0; id < N; id++)
	{
		const double T = s.T[id], p = s.p[id], rho = s.rho[id];
		if (c.visc_diffu)
			for (int ii = 0; ii < NS; ii++)
				s.hi[ii + NS * id] = c.cop ? get_Enthalpy_NASA(c.Hia, c.Hib, T, c.Ri[ii], ii) : 0.0;
		const double *yi = &s.y[NS * id];
		double C_total = 0.0;
		for (int i = 0; i < NS; i++)
		{
			X[i] = yi[i] * c._Wi[i] * 1e-3 * rho;
			C_total = C_total + X[i];
		}
		const double _C_total = 1.0 / C_total;
		for (int i = 0; i < NS; i++)
			X[i] = X[i] * _C_total;
		double va = 0.0, tca = 0.0;
		for (int k = 0; k < NS; k++)
		{
			double denominator = 0.0;
			for (int i = 0; i < NS; i++)
			{ // PHI(specie[k], specie[i])
				double phi = std::pow(c.Wi[i] / c.Wi[k], 0.25) * std::pow(FitValue(c.fit_visc + 4 * k, T) / FitValue(c.fit_visc + 4 * i, T), 0.5);
				phi = (phi + 1.0) * (phi + 1.0) * 0.5 / std::sqrt(2.0);
				phi = phi * std::pow(1.0 + c.Wi[k] / c.Wi[i], -0.5);
				denominator = denominator + X[i] * phi;
			}
			const double _denominator = 1.0 / denominator;
			va = va + X[k] * FitValue(c.fit_visc + 4 * k, T) * _denominator;
			if (c.visc_heat)
				tca = tca + X[k] * FitValue(c.fit_therm + 4 * k, T) * _denominator;
		}
		s.va[id] = va;
		if (c.visc_heat)
			s.tca[id] = tca;
		if (c.visc_diffu)
		{
			double *Dk = &s.Dkm[NS * id];
			if (NS > 1)
				for (int k = 0; k < NS; k++)
				{
					double temp1 = 0.0, temp2 = 1.0e-20;
					for (int i = 0; i < NS; i++)
						if (i != k)
						{
							temp1 += (X[i] + 1.0e-40) * c.Wi[i];
							temp2 += (X[i] + 1.0e-40) / (FitValue(c.fit_Dkj + 4 * (i * NS + k), T) / p + 1.0e-40);
						}
					if (!(0.0 < std::ceil(temp1))) // sycl::step(ceil(temp1), 0.0)
						Dk[k] = FitValue(c.fit_Dkj + 4 * (k * NS + k), T) / p;
					else
						Dk[k] = temp1 / temp2 / rho * C_total;
					Dk[k] *= 1.0e-1;
				}
			else
			{
				Dk[0] = FitValue(c.fit_Dkj, T) / p;
				Dk[0] *= 1.0e-1;
			}
			for (int k = 0; k < NS; k++)
				Dk[k] = (Dk[k] < 1.0e-10) ? 1.0e-10 : Dk[k];
		}
	}
}
// GetWallViscousFlux{X,Y,Z} (Visc_Order_kernels.hpp:76-340) with the macros of Fourth_Order/Flux_discrete.h
static void GetWallViscousFlux(const xo_cfg &c, xo_state &s, int dir, double *Flux_wall, const double *Yil_limiter, const double *Diffu_limiter)
{
	const int NS = c.NS, E = c.Emax;
	const double _sxtn = 1.0 / 16.0, _twfr = 1.0 / 24.0;
	const double tX = c.DimX, tY = c.DimY, tZ = c.DimZ;
	const size_t sy = c.Xmax, sz = size_t(c.Xmax) * c.Ymax, st = dir == 0 ? 1 : (dir == 1 ? sy : sz);
	const double _dl = dir == 0 ? c._dx : (dir == 1 ? c._dy : c._dz);
	const int i0 = c.Bw_X - (dir == 0), j0 = c.Bw_Y - (dir == 1), k0 = c.Bw_Z - (dir == 2);
	for (int k = k0; k < c.Z_inner + c.Bw_Z; k++)
		for (int j = j0; j < c.Y_inner + c.Bw_Y; j++)
			for (int i = i0; i < c.X_inner + c.Bw_X; i++)
			{
				const size_t id = sz * k + sy * j + i, id_m1 = id - st, id_p1 = id + st, id_p2 = id + 2 * st;
				auto avg = [&](const double *q) { return (9.0 * (q[id_p1] + q[id]) - (q[id_p2] + q[id_m1])) * _sxtn; };
				const double *u = s.u, *v = s.v, *w = s.w;
				std::vector<double> F_wall_v(E);
				double f_x, f_y, f_z, u_hlf, v_hlf, w_hlf;
				const double mue = avg(s.va);
				const double lamada = -2.0 * _OT * mue;
				if (dir == 0)
				{
					const double *Ducy = s.Vde[3], *Ducz = s.Vde[6], *Dvcy = s.Vde[4], *Dwcz = s.Vde[8];
					f_x = (2.0 * mue + lamada) * (27.0 * (u[id_p1] - u[id]) - (u[id_p2] - u[id_m1])) * _dl * _twfr;
					f_x += lamada * (9.0 * (Dvcy[id_p1] + Dvcy[id]) - (Dvcy[id_p2] + Dvcy[id_m1]) + 9.0 * (Dwcz[id_p1] + Dwcz[id]) - (Dwcz[id_p2] + Dwcz[id_m1])) * _sxtn;
					f_y = mue * (27.0 * (v[id_p1] - v[id]) - (v[id_p2] - v[id_m1])) * _dl * _twfr * tY;
					f_y += mue * (9.0 * (Ducy[id_p1] + Ducy[id]) - (Ducy[id_p2] + Ducy[id_m1])) * _sxtn * tY;
					f_z = mue * (27.0 * (w[id_p1] - w[id]) - (w[id_p2] - w[id_m1])) * _dl * _twfr * tZ;
					f_z += mue * (9.0 * (Ducz[id_p1] + Ducz[id]) - (Ducz[id_p2] + Ducz[id_m1])) * _sxtn * tZ;
					u_hlf = avg(u), v_hlf = avg(v) * tY, w_hlf = avg(w) * tZ;
				}
				else if (dir == 1)
				{
					const double *Dvcx = s.Vde[1], *Dvcz = s.Vde[7], *Ducx = s.Vde[0], *Dwcz = s.Vde[8];
					f_x = mue * (27.0 * (u[id_p1] - u[id]) - (u[id_p2] - u[id_m1])) * _dl * _twfr * tX;
					f_x += mue * (9.0 * (Dvcx[id_p1] + Dvcx[id]) - (Dvcx[id_p2] + Dvcx[id_m1])) * _sxtn * tX;
					f_y = (2.0 * mue + lamada) * (27.0 * (v[id_p1] - v[id]) - (v[id_p2] - v[id_m1])) * _dl * _twfr;
					f_y += lamada * (9.0 * (Ducx[id_p1] + Ducx[id]) - (Ducx[id_p2] + Ducx[id_m1]) + 9.0 * (Dwcz[id_p1] + Dwcz[id]) - (Dwcz[id_p2] + Dwcz[id_m1])) * _sxtn;
					f_z = mue * (27.0 * (w[id_p1] - w[id]) - (w[id_p2] - w[id_m1])) * _dl * _twfr * tZ;
					f_z += mue * (9.0 * (Dvcz[id_p1] + Dvcz[id]) - (Dvcz[id_p2] + Dvcz[id_m1])) * _sxtn * tZ;
					u_hlf = avg(u) * tX, v_hlf = avg(v), w_hlf = avg(w) * tZ;
				}
				else
				{
					const double *Dwcx = s.Vde[2], *Dwcy = s.Vde[5], *Ducx = s.Vde[0], *Dvcy = s.Vde[4];
					f_x = mue * (27.0 * (u[id_p1] - u[id]) - (u[id_p2] - u[id_m1])) * _dl * _twfr * tX;
					f_x += mue * (9.0 * (Dwcx[id_p1] + Dwcx[id]) - (Dwcx[id_p2] + Dwcx[id_m1])) * _sxtn * tX;
					f_y = mue * (27.0 * (v[id_p1] - v[id]) - (v[id_p2] - v[id_m1])) * _dl * _twfr * tY;
					f_y += mue * (9.0 * (Dwcy[id_p1] + Dwcy[id]) - (Dwcy[id_p2] + Dwcy[id_m1])) * _sxtn * tY;
					f_z = (2.0 * mue + lamada) * (27.0 * (w[id_p1] - w[id]) - (w[id_p2] - w[id_m1])) * _dl * _twfr;
					f_z += lamada * (9.0 * (Ducx[id_p1] + Ducx[id]) - (Ducx[id_p2] + Ducx[id_m1]) + 9.0 * (Dvcy[id_p1] + Dvcy[id]) - (Dvcy[id_p2] + Dvcy[id_m1])) * _sxtn;
					u_hlf = avg(u) * tX, v_hlf = avg(v) * tY, w_hlf = avg(w);
				}
				F_wall_v[0] = 0.0, F_wall_v[1] = f_x, F_wall_v[2] = f_y, F_wall_v[3] = f_z;
				F_wall_v[4] = f_x * u_hlf + f_y * v_hlf + f_z * w_hlf;
				if (c.visc_heat)
				{
					double kk = avg(s.tca);
					kk *= (27.0 * (s.T[id_p1] - s.T[id]) - (s.T[id_p2] - s.T[id_m1])) * _dl * _twfr;
					F_wall_v[4] += kk;
				}
				if (c.visc_diffu)
				{
					const double rho_wall = avg(s.rho);
					double CorrectTerm = 0.0, Dim_Yil = 1.0E-20;
					std::vector<double> Yi_wall(NS, 0.0);
					for (int l = 0; l < NS; l++)
					{
						const size_t g_p1 = l + NS * id_p1, g = l + NS * id, g_p2 = l + NS * id_p2, g_m1 = l + NS * id_m1;
						const double hi_wall = (9.0 * (s.hi[g_p1] + s.hi[g]) - (s.hi[g_p2] + s.hi[g_m1])) * _sxtn;
						const double Dim_wall = (9.0 * (s.Dkm[g_p1] + s.Dkm[g]) - (s.Dkm[g_p2] + s.Dkm[g_m1])) * _sxtn;
						if (c.cop)
						{
							const double *Yi = s.y;
							auto smin = [](double a, double b) { return (b < a) ? b : a; };
							auto smax = [](double a, double b) { return (a < b) ? b : a; };
							const double Yil_wall = smin(smax((27.0 * (Yi[g_p1] - Yi[g]) - (Yi[g_p2] - Yi[g_m1])) * _dl * _twfr, -Yil_limiter[l]), Yil_limiter[l]);
							Yi_wall[l] = smin(smax((9.0 * (Yi[g_p1] + Yi[g]) - (Yi[g_p2] + Yi[g_m1])) * _sxtn, 1.0E-20), 1.0);
							Dim_Yil = smin(smax(Dim_wall * Yil_wall, -Diffu_limiter[l]), Diffu_limiter[l]);
							CorrectTerm += Dim_Yil;
						}
						F_wall_v[4] += rho_wall * hi_wall * Dim_Yil;
					}
					CorrectTerm *= rho_wall;
					for (int p = 5; p < E; p++)
						F_wall_v[p] = rho_wall * Dim_Yil - Yi_wall[p - 5] * CorrectTerm;
				}
				else
					for (int p = 5; p < E; p++)
						F_wall_v[p] = 0.0;
				for (int n = 0; n < E; n++)
					Flux_wall[n + E * id] -= F_wall_v[n];
			}
}
static void ViscousBlock(const xo_cfg &c, xo_state &s)
{
	GetCellCenterDerivative(c, s);
	for (int dir = 0; dir < 3; dir++)
		if (dir == 0 ? c.DimX : (dir == 1 ? c.DimY : c.DimZ))
			CenterDerivativeBC(c, s, dir);
	GetTransportCoeff(c, s);
	std::vector<double> yi_max(c.NS, 0.0), Dim_max(c.NS, 0.0);
	if (c.visc_diffu)
		for (int nn = 0; nn < c.NS; nn++)
		{ // ConVenction_block.hpp:458-504: both reductions start from 0.0
			double ymin = 0.0, ymax = 0.0;
			for (int k = c.Bw_Z; k < c.Zmax - c.Bw_Z; k++)
				for (int j = c.Bw_Y; j < c.Ymax - c.Bw_Y; j++)
					for (int i = c.Bw_X; i < c.Xmax - c.Bw_X; i++)
					{
						const double y = s.y[c.NS * (size_t(c.Xmax) * c.Ymax * k + size_t(c.Xmax) * j + i) + nn];
						ymin = (y < ymin) ? y : ymin, ymax = (ymax < y) ? y : ymax;
					}
			double dmax = c.dim_max0;
			if (c.dim_max0 != 0.0)
				ymin = (ymin < 0.0) ? 0.0 : ymin, ymax = (1.0 < ymax) ? 1.0 : ymax;
			ymax -= ymin;
			ymax *= c.Yil_limiter;
			dmax *= c.Dim_limiter * ymax;
			yi_max[nn] = ymax, Dim_max[nn] = dmax;
		}
	if (c.DimX)
		GetWallViscousFlux(c, s, 0, s.FluxFw, yi_max.data(), Dim_max.data());
	if (c.DimY)
		GetWallViscousFlux(c, s, 1, s.FluxGw, yi_max.data(), Dim_max.data());
	if (c.DimZ)
		GetWallViscousFlux(c, s, 2, s.FluxHw, yi_max.data(), Dim_max.data());
}

static void GetLU(const xo_cfg &c, xo_state &s, const double *UI)
{
	if (c.DimX)
		GetLocalEigen(c, s, 1.0, 0.0, 0.0, s.eig_x);
	if (c.DimY)
		GetLocalEigen(c, s, 0.0, 1.0, 0.0, s.eig_y);
	if (c.DimZ)
		GetLocalEigen(c, s, 0.0, 0.0, 1.0, s.eig_z);
	if (c.DimX)
		GlobalEigenMax(c, s.eig_x, s.eigen_block_x);
	if (c.DimY)
		GlobalEigenMax(c, s.eig_y, s.eigen_block_y);
	if (c.DimZ)
		GlobalEigenMax(c, s.eig_z, s.eigen_block_z);
	if (c.DimX)
		ReconstructFlux(c, s, 0, UI, s.FluxF, s.FluxFw, s.eig_x, s.eigen_block_x);
	if (c.DimY)
		ReconstructFlux(c, s, 1, UI, s.FluxG, s.FluxGw, s.eig_y, s.eigen_block_y);
	if (c.DimZ)
		ReconstructFlux(c, s, 2, UI, s.FluxH, s.FluxHw, s.eig_z, s.eigen_block_z);
	if (c.positivity)
	{ // ConVenction_block.hpp:330-410: lambda_d0 = uvw_c_max[d] of the last GetDt, lambda_d = CFL / lambda_d0
		const double lx0 = s.uvw_c_max[0], ly0 = s.uvw_c_max[1], lz0 = s.uvw_c_max[2];
		if (c.DimX)
			PositivityPreserving(c, 0, UI, s.FluxF, s.FluxFw, lx0, c.CFL / lx0);
		if (c.DimY)
			PositivityPreserving(c, 1, UI, s.FluxG, s.FluxGw, ly0, c.CFL / ly0);
		if (c.DimZ)
			PositivityPreserving(c, 2, UI, s.FluxH, s.FluxHw, lz0, c.CFL / lz0);
	}
	if (c.visc)
		ViscousBlock(c, s);
	UpdateFluidLU(c, s);
}

// ---- solver_UpdateStates/Update_kernels.hpp:64-94 (K12) -----------------------------------------
static void UpdateURK3rd(const xo_cfg &c, xo_state &s, const double dt, int flag)
{
	const int E = c.Emax;
	double *U = s.U, *U1 = s.U1, *LU = s.LU;
	for (int k = c.Bw_Z; k < c.Zmax - c.Bw_Z; k++)
		for (int j = c.Bw_Y; j < c.Ymax - c.Bw_Y; j++)
			for (int i = c.Bw_X; i < c.Xmax - c.Bw_X; i++)
			{
				size_t id = XO_ID(i, j, k);
				switch (flag)
				{
				case 1:
					for (int n = 0; n < E; n++)
						U1[E * id + n] = U[E * id + n] + dt * LU[E * id + n];
					break;
				case 2:
					for (int n = 0; n < E; n++)
						U1[E * id + n] = 0.75 * U[E * id + n] + 0.25 * U1[E * id + n] + 0.25 * dt * LU[E * id + n];
					break;
				case 3:
					for (int n = 0; n < E; n++)
						U[E * id + n] = (U[E * id + n] + 2.0 * U1[E * id + n] + 2.0 * dt * LU[E * id + n]) * _OT;
					break;
				}
			}
}

// ---- solver_BCs/BCs_kernels.hpp:9-258 + BCs_block.cpp:50-217 (K14) ------------------------------
// dir 0/1/2; g = ghost index along dir; mirror_offset/index_inner/sign as passed by BCs_block.cpp
static void FluidBCKernel(const xo_cfg &c, int dir, int i, int j, int k, int BC, double *d_UI, int mirror_offset, int index_inner, int sign)
{
	const int E = c.Emax, NUM_COP = c.NCOP;
	const int Bw = dir == 0 ? c.Bw_X : (dir == 1 ? c.Bw_Y : c.Bw_Z);
	const int inner = dir == 0 ? c.X_inner : (dir == 1 ? c.Y_inner : c.Z_inner);
	const int g = dir == 0 ? i : (dir == 1 ? j : k);
	size_t id = XO_ID(i, j, k);
	auto tid = [&](int t) -> size_t
	{ return dir == 0 ? XO_ID(t, j, k) : (dir == 1 ? XO_ID(i, t, k) : XO_ID(i, j, t)); };
	switch (BC)
	{
	case 2: // Symmetry
	{
		int offset = 2 * (Bw + mirror_offset) - 1;
		size_t target_id = tid(offset - g);
		for (int n = 0; n < E; n++)
			d_UI[E * id + n] = d_UI[E * target_id + n];
		d_UI[E * id + 1 + dir] = -d_UI[E * target_id + 1 + dir];
	}
	break;
	case 3: // Periodic
	{
		size_t target_id = tid(g + sign * inner);
		for (int n = 0; n < E; n++)
			d_UI[E * id + n] = d_UI[E * target_id + n];
	}
	break;
	case 0: // Inflow
		break;
	case 1: // Outflow
	{
		size_t target_id = tid(index_inner);
		for (int n = 0; n < E; n++)
			d_UI[E * id + n] = d_UI[E * target_id + n];
	}
	break;
	case 4: // nslipWall
	case 5: // viscWall  (X only; no-op in Y and Z, BCs_kernels.hpp:167-171,242-246)
	case 6: // slipWall  (X only)
	{
		if (BC != 4 && dir != 0)
			break;
		int offset = 2 * (Bw + mirror_offset) - 1;
		size_t target_id = tid(offset - g);
		d_UI[E * id + 0] = d_UI[E * target_id + 0];
		d_UI[E * id + 1] = -d_UI[E * target_id + 1];
		d_UI[E * id + 2] = -d_UI[E * target_id + 2];
		d_UI[E * id + 3] = -d_UI[E * target_id + 3];
		d_UI[E * id + 4] = d_UI[E * target_id + 4];
		if (c.cop)
			for (int n = E - NUM_COP; n < E; n++)
				d_UI[E * id + n] = d_UI[E * target_id + n];
	}
	break;
	default: // innerBlock
		break;
	}
}
static void FluidBoundaryCondition(const xo_cfg &c, double *d_UI)
{
	if (c.DimX)
		for (int k = 0; k < c.Zmax; k++)
			for (int j = 0; j < c.Ymax; j++)
				for (int g = 0; g < c.Bw_X; g++)
				{
					FluidBCKernel(c, 0, g, j, k, c.bc[0], d_UI, 0, c.Bw_X, 1);
					FluidBCKernel(c, 0, g + c.Xmax - c.Bw_X, j, k, c.bc[1], d_UI, c.X_inner, c.Xmax - c.Bw_X - 1, -1);
				}
	if (c.DimY)
		for (int k = 0; k < c.Zmax; k++)
			for (int g = 0; g < c.Bw_Y; g++)
				for (int i = 0; i < c.Xmax; i++)
				{
					FluidBCKernel(c, 1, i, g, k, c.bc[2], d_UI, 0, c.Bw_Y, 1);
					FluidBCKernel(c, 1, i, g + c.Ymax - c.Bw_Y, k, c.bc[3], d_UI, c.Y_inner, c.Ymax - c.Bw_Y - 1, -1);
				}
	if (c.DimZ)
		for (int g = 0; g < c.Bw_Z; g++)
			for (int j = 0; j < c.Ymax; j++)
				for (int i = 0; i < c.Xmax; i++)
				{
					FluidBCKernel(c, 2, i, j, g, c.bc[4], d_UI, 0, c.Bw_Z, 1);
					FluidBCKernel(c, 2, i, j, g + c.Zmax - c.Bw_Z, c.bc[5], d_UI, c.Z_inner, c.Zmax - c.Bw_Z - 1, -1);
				}
}

// ---- solver_GetDt/GlobalDt_block.hpp:6-114 (K13) ------------------------------------------------
static double GetDt(const xo_cfg &c, xo_state &s)
{
	const size_t N = size_t(c.Xmax) * c.Ymax * c.Zmax;
	for (int n = 0; n < 6; n++)
		s.uvw_c_max[n] = 0.0;
	for (size_t id = 0; id < N; id++)
	{
		double c_local = std::sqrt(1.4 * s.p[id] / s.rho[id]); // hard-coded 1.4, scans ghosts too
		if (c.DimX)
			s.uvw_c_max[0] = std::fmax(s.uvw_c_max[0], std::fabs(s.u[id]) + c_local);
		if (c.DimY)
			s.uvw_c_max[1] = std::fmax(s.uvw_c_max[1], std::fabs(s.v[id]) + c_local);
		if (c.DimZ)
			s.uvw_c_max[2] = std::fmax(s.uvw_c_max[2], std::fabs(s.w[id]) + c_local);
	}
	double dtref = s.uvw_c_max[0] * c._dx + s.uvw_c_max[1] * c._dy + s.uvw_c_max[2] * c._dz;
	return c.CFL / dtref;
}

// =================================================================================================
//  C interface used by the tests (ctypes)
// =================================================================================================
extern "C"
{
	size_t xo_ncells(const xo_cfg *c) { return size_t(c->Xmax) * c->Ymax * c->Zmax; }
	// y[i] = this machine's libm log(x[i]) -- the reference CPU path's logarithm (get_Enthalpy_NASA, Thermo_device.h:62-80), against which
	// tests/test_xf_log.py checks the device logarithm bit for bit.  Called through a volatile pointer: no builtin expansion.
	void xo_libm_log(const double *x, double *y, size_t n)
	{
		double (*volatile f)(double) = std::log;
#pragma omp parallel for
		for (long long i = 0; i < (long long)n; i++)
			y[i] = f(x[i]);
	}

	// the same for exp (which = 1) and pow (which = 2, exponents in y2): the transport fits' functions (Visc_device.h:10-41)
	void xo_libm_eval(int which, const double *x, const double *y2, double *y, size_t n)
	{
		double (*volatile fe)(double) = std::exp;
		double (*volatile fl)(double) = std::log;
		double (*volatile fp)(double, double) = std::pow;
#pragma omp parallel for
		for (long long i = 0; i < (long long)n; i++)
			y[i] = which == 0 ? fl(x[i]) : (which == 1 ? fe(x[i]) : fp(x[i], y2[i]));
	}

	xo_state *xo_state_create(const xo_cfg *c)
	{
		xo_state *s = (xo_state *)std::calloc(1, sizeof(xo_state));
		const size_t N = xo_ncells(c), E = c->Emax, NS = c->NS;
		auto A = [](size_t n)
		{ return (double *)std::calloc(n ? n : 1, sizeof(double)); };
		s->U = A(N * E), s->U1 = A(N * E), s->LU = A(N * E);
		s->FluxF = A(N * E), s->FluxG = A(N * E), s->FluxH = A(N * E);
		s->FluxFw = A(N * E), s->FluxGw = A(N * E), s->FluxHw = A(N * E);
		s->eig_x = A(N * E), s->eig_y = A(N * E), s->eig_z = A(N * E);
		s->eigen_block_x = A(E), s->eigen_block_y = A(E), s->eigen_block_z = A(E);
		s->rho = A(N), s->p = A(N), s->u = A(N), s->v = A(N), s->w = A(N), s->c = A(N), s->gamma = A(N);
		s->e = A(N), s->H = A(N), s->T = A(N), s->Ri = A(N), s->Cp = A(N), s->y = A(N * NS);
		if (c->visc)
		{
			for (int m = 0; m < 9; m++)
				s->Vde[m] = A(N);
			s->va = A(N), s->tca = A(N), s->Dkm = A(N * NS), s->hi = A(N * NS);
		}
		return s;
	}
	void xo_state_destroy(xo_state *s)
	{
		double *ps[] = {s->U, s->U1, s->LU, s->FluxF, s->FluxG, s->FluxH, s->FluxFw, s->FluxGw, s->FluxHw, s->eig_x, s->eig_y, s->eig_z,
						s->eigen_block_x, s->eigen_block_y, s->eigen_block_z, s->rho, s->p, s->u, s->v, s->w, s->c, s->gamma, s->e, s->H, s->T, s->Ri, s->Cp, s->y};
		for (double *p : ps)
			std::free(p);
		for (int m = 0; m < 9; m++)
			std::free(s->Vde[m]);
		std::free(s->va), std::free(s->tca), std::free(s->Dkm), std::free(s->hi);
		std::free(s);
	}
	// named array access: returns pointer, *len = number of doubles
	double *xo_array(const xo_cfg *c, xo_state *s, const char *name, size_t *len)
	{
		const size_t N = xo_ncells(c), E = c->Emax, NS = c->NS;
		struct
		{
			const char *n;
			double *p;
			size_t l;
		} t[] = {{"U", s->U, N * E}, {"U1", s->U1, N * E}, {"LU", s->LU, N * E}, {"FluxF", s->FluxF, N * E}, {"FluxG", s->FluxG, N * E}, {"FluxH", s->FluxH, N * E}, {"FluxFw", s->FluxFw, N * E}, {"FluxGw", s->FluxGw, N * E}, {"FluxHw", s->FluxHw, N * E}, {"rho", s->rho, N}, {"p", s->p, N}, {"u", s->u, N}, {"v", s->v, N}, {"w", s->w, N}, {"c", s->c, N}, {"gamma", s->gamma, N}, {"e", s->e, N}, {"H", s->H, N}, {"T", s->T, N}, {"R", s->Ri, N}, {"Cp", s->Cp, N}, {"y", s->y, N * NS}, {"eigen_block_x", s->eigen_block_x, E}, {"eigen_block_y", s->eigen_block_y, E}, {"eigen_block_z", s->eigen_block_z, E}, {"uvw_c_max", s->uvw_c_max, 6},
				 {"visc", s->va, N}, {"therm", s->tca, N}, {"Dkm", s->Dkm, N * NS}, {"hi", s->hi, N * NS}, {"Vde0", s->Vde[0], N}, {"Vde1", s->Vde[1], N}, {"Vde2", s->Vde[2], N},
				 {"Vde3", s->Vde[3], N}, {"Vde4", s->Vde[4], N}, {"Vde5", s->Vde[5], N}, {"Vde6", s->Vde[6], N}, {"Vde7", s->Vde[7], N}, {"Vde8", s->Vde[8], N}};
		for (auto &e : t)
			if (!std::strcmp(e.n, name))
			{
				if (len)
					*len = e.l;
				return e.p;
			}
		return nullptr;
	}
	int *xo_error_flags(xo_state *s) { return s->error_flags; }

	// the six block-level entry points of the reference, on which = 0 (U) or 1 (U1)
	void xo_boundary(const xo_cfg *c, xo_state *s, int which) { FluidBoundaryCondition(*c, which ? s->U1 : s->U); }
	int xo_update_states(const xo_cfg *c, xo_state *s, int which)
	{ // UpdateStates_block.cpp:7-191: K1, K2, K3, K4
		double *UI = which ? s->U1 : s->U;
		UpdateFluidStates(*c, *s, UI);
		EstimateGuards(*c, *s, false);
		EstimateGuards(*c, *s, true);
		return s->error_flags[0] || s->error_flags[1];
	}
	void xo_get_lu(const xo_cfg *c, xo_state *s, int which) { GetLU(*c, *s, which ? s->U1 : s->U); }
	void xo_update_u(const xo_cfg *c, xo_state *s, double dt, int flag) { UpdateURK3rd(*c, *s, dt, flag); }
	double xo_get_dt(const xo_cfg *c, xo_state *s) { return GetDt(*c, *s); }

	// one SSP-RK3 stage in the reference's order (XFLUIDS.cpp:441-525)
	int xo_rk_stage(const xo_cfg *c, xo_state *s, double dt, int flag)
	{
		const int which = flag == 1 ? 0 : 1;
		double *UI = which ? s->U1 : s->U;
		FluidBoundaryCondition(*c, UI);
		if (xo_update_states(c, s, which))
			return 1;
		GetLU(*c, *s, UI);
		EstimateFluidNAN(*c, *s, flag == 3 ? s->U : s->U1); // Fluids.cpp:963-985: checks U1,U1,U for stages 1,2,3
		if (s->error_flags[2])
			return 1;
		UpdateURK3rd(*c, *s, dt, flag);
		return 0;
	}
	// startup sequence of main.cpp:44-48 after the initial condition: BC(U), UpdateStates(U)
	int xo_startup(const xo_cfg *c, xo_state *s)
	{
		FluidBoundaryCondition(*c, s->U);
		return xo_update_states(c, s, 0);
	}
	// nsteps full time steps (XFLUIDS.cpp:172-294): dt, optional clip to t_end, 3 stages.
	// dts[step] receives the dt used; returns the number of completed steps (negative on error)
	int xo_run(const xo_cfg *c, xo_state *s, int nsteps, double t_start, double t_end, double *dts, double *t_out)
	{
		double t = t_start;
		int it = 0;
		for (; it < nsteps && t < t_end; it++)
		{
			double dt = GetDt(*c, *s);
			if (t + dt > t_end)
				dt = t_end - t;
			t += dt;
			if (dts)
				dts[it] = dt;
			for (int flag = 1; flag <= 3; flag++)
				if (xo_rk_stage(c, s, dt, flag))
				{
					if (t_out)
						*t_out = t;
					return -(it + 1);
				}
		}
		if (t_out)
			*t_out = t;
		return it;
	}
	void xo_newton_stats(int *calls, int *iters, int reset)
	{
		*calls = g_newton_calls, *iters = g_newton_iters;
		if (reset)
			g_newton_calls = g_newton_iters = 0;
	}
	double xo_Ru(void) { return Ru; }
}
