// xfh_transport.cpp -- host side of the viscous terms: Lennard-Jones transport parameters (runtime.dat/transport_data.dat), the
// Monchick-Mason collision-integral table (runtime.dat/collision_integral.dat) and the least-squares fits of ln(mu_k), ln(lambda_k),
// ln(p D_kj) as cubic polynomials in ln T that the device kernels evaluate.
//
// Follows Setup::GetFitCoefficient / Fitting / ReadOmega_table / Omega_interpolated / viscosity / thermal_conductivities / Dkj
// (reference src/read_ini/src/viscfit.cpp:148-555), GetQuadraticInterCoeff / ZrotFunc / Solve_Overdeter_equations
// (src/read_ini/mixture.hpp:43-150) and the transport part of Setup::ReadThermal (src/read_ini/src/thermal.cpp:131-165), operation by
// operation: tests/test_host_setup.py compares the fits with the arrays the unmodified reference computes (oracle/_ref, bit for bit).
#include <cmath>
#include <fstream>
#include <stdexcept>
#include "xfh_setup.hpp"

namespace xfh
{
	namespace
	{
		// global_setup.h:41-52
		const double pi = 3.1415926535897932384626433832795;
		const double kB = 1.3806549 * 1.0e-16; // erg/K
		const double NA = 6.02214129 * 1.0e23;
		const double universal_gas_const = 6.02214076e26 * 1.380649e-23;
		enum { geo = 0, epsilon_kB = 1, d_ = 2, mue = 3, alpha = 4, Zrot_298 = 5, WI = 6, SID = 8 };

		// parabola q[0] + q[1] x + q[2] x^2 through three nodes (mixture.hpp:70-77), divided differences in the reference's order
		void GetQuadraticInterCoeff(double xa, double xb, double xc, double fa, double fb, double fc, double *q)
		{
			q[2] = ((fa - fb) / (xa - xb) - (fb - fc) / (xb - xc)) / (xa - xc);
			q[1] = (fa - fb) / (xa - xb) - q[2] * (xa + xb);
			q[0] = fa - q[1] * xa - q[2] * xa * xa;
		}
		double ZrotFunc(double x) { return 1.0 + std::sqrt(pi * x) * pi / 2.0 + x * (0.25 * pi * pi + 2.0) + std::pow(pi * x, 1.5); }

		// Least squares through the normal equations, solved by Cholesky: G = A^T A = L L^T, then L y = A^T b and L^T c = y.  Every sum and
		// every quotient runs in the order of the reference's solver (mixture.hpp:89-150): the fits have to come out bit for bit.
		void Solve_Overdeter_equations(const std::vector<std::array<double, 4>> &A, const std::vector<double> &b, double *coef)
		{
			constexpr int N = 4;
			const size_t rows = b.size();
			double G[N][N] = {}, rhs[N] = {}, y[N];
			for (int c = 0; c < N; c++)
			{
				for (size_t r = 0; r < rows; r++)
					rhs[c] += A[r][c] * b[r];
				for (int c2 = 0; c2 < N; c2++)
					for (size_t r = 0; r < rows; r++)
						G[c][c2] += A[r][c] * A[r][c2];
			}
			// L overwrites the lower triangle of G, column by column
			G[0][0] = std::sqrt(G[0][0]);
			for (int r = 1; r < N; r++)
				G[r][0] /= G[0][0];
			for (int d = 1; d < N; d++)
			{
				for (int m = 0; m < d; m++)
					G[d][d] -= G[d][m] * G[d][m];
				G[d][d] = std::sqrt(G[d][d]);
				for (int r = d + 1; r < N; r++)
				{
					for (int m = 0; m < d; m++)
						G[r][d] -= G[r][m] * G[d][m];
					G[r][d] /= G[d][d];
				}
			}
			y[0] = rhs[0] / G[0][0]; // forward substitution
			for (int r = 1; r < N; r++)
			{
				for (int m = 0; m < r; m++)
					rhs[r] -= G[r][m] * y[m];
				y[r] = rhs[r] / G[r][r];
			}
			coef[N - 1] = y[N - 1] / G[N - 1][N - 1]; // back substitution, highest coefficient first
			for (int r = N - 2; r >= 0; r--)
			{
				for (int m = N - 1; m > r; m--)
					y[r] -= G[m][r] * coef[m];
				coef[r] = y[r] / G[r][r];
			}
		}
	} // namespace

	// Quirk kept (it decides the bits of the fits): Omega_interpolated picks three consecutive table rows ti1 < ti2 < ti3 around T*; for
	// T* between the last two rows (75 < T* < 100: H2 at the 3000 K fitting node, the H2-N2 / H2-O2 pairs at 5000 K) its search leaves
	// ti3 = 37, ONE PAST the tables.  The reference then reads what lies behind them in struct Setup (setupini.h:66-67: Omega_table[2][37][8],
	// delta_star[8], T_star[37], next member): Omega_table[0][37][j] is Omega_table[1][0][j], Omega_table[1][37][j] is delta_star[j], and
	// T_star[37] is the first word of the following member (a pointer / zero: a denormal as a double, indistinguishable from 0.0 in
	// x1 - x3 and x2 - x3).  The same flat memory image is kept here so that the same numbers come out.
	struct TransportFit
	{
		const Setup &S;
		double mem[2 * 37 * 8 + 8 + 37 + 8] = {0};
		struct Row
		{
			double *p;
			double &operator[](int j) { return p[j]; }
			double operator[](int j) const { return p[j]; }
		};
		struct Tab
		{
			double *p;
			Row operator[](int i) const { return Row{p + i * 8}; }
		};
		struct Tabs
		{
			double *p;
			Tab operator[](int index) const { return Tab{p + index * 37 * 8}; }
		};
		Tabs Omega_table{mem};
		double *delta_star = mem + 2 * 37 * 8, *T_star = mem + 2 * 37 * 8 + 8;
		explicit TransportFit(const Setup &s) : S(s)
		{ // ReadOmega_table (viscfit.cpp:348-364)
			const std::string fpath = S.WorkDir + "/runtime.dat/collision_integral.dat";
			std::ifstream fin(fpath);
			if (!fin)
				throw std::runtime_error("cannot open " + fpath);
			// file order: the 8 reduced dipole moments; 37 rows of {T*, Omega(2,2)* at the 8 moments}; 37 x 8 values of Omega(1,1)*
			auto take = [&](double *dst, int count)
			{
				for (int c = 0; c < count; c++)
					fin >> dst[c];
			};
			take(delta_star, 8);
			for (int row = 0; row < 37; row++)
				take(T_star + row, 1), take(mem + (37 + row) * 8, 8); // table index 1 = Omega(2,2)*
			take(mem, 37 * 8);                                        // table index 0 = Omega(1,1)*
			if (!fin)
				throw std::runtime_error("collision_integral.dat: short table");
		}
		double Omega_interpolated(double Tstar, double deltastar, int index) const
		{
			// first of the three consecutive nodes the reference's branches pick (viscfit.cpp:366-420): clamped at both ends, otherwise the
			// node below x and the two after it -- for x between the last two nodes that is one node PAST the grid (see the note above)
			auto first_node = [](const double *grid, int n, double x)
			{
				if (x <= grid[0])
					return 0;
				if (x >= grid[n - 1])
					return n - 3;
				int at = 1;
				while (x > grid[at])
					++at;
				return at - 1;
			};
			const int ti = first_node(T_star, 37, Tstar), tj = first_node(delta_star, 8, deltastar);
			double aa[3], col[3];
			for (int c = 0; c < 3; c++)
			{ // quadratic through the three T* nodes at each of the three delta* nodes, evaluated at T*
				GetQuadraticInterCoeff(T_star[ti], T_star[ti + 1], T_star[ti + 2], Omega_table[index][ti][tj + c], Omega_table[index][ti + 1][tj + c],
									   Omega_table[index][ti + 2][tj + c], aa);
				col[c] = aa[0] + aa[1] * Tstar + aa[2] * Tstar * Tstar;
			}
			GetQuadraticInterCoeff(delta_star[tj], delta_star[tj + 1], delta_star[tj + 2], col[0], col[1], col[2], aa);
			return aa[0] + aa[1] * deltastar + aa[2] * deltastar * deltastar;
		}
		double viscosity(const double *specie, double T) const
		{
			const double Tstar = T / specie[epsilon_kB];
			const double deltastar = 0.5 * specie[mue] * specie[mue] / specie[epsilon_kB] / kB / (std::pow(specie[d_], 3)) * 1.0e-12;
			const double Omega2 = Omega_interpolated(Tstar, deltastar, 0);
			double visc = 5 * 1.0e16 * std::sqrt(pi * (specie[WI] * 1e3) / NA * kB * T) / (16 * pi * specie[d_] * specie[d_] * Omega2);
			return visc = 0.1 * visc;
		}
		double Dkj(const double *specie_k, const double *specie_j, double T, double PP) const
		{
			double epsilon_jk_kB, d_jk, mue_jk_sqr;
			if ((specie_j[mue] > 0 && specie_k[mue] > 0) || (specie_j[mue] == 0 && specie_k[mue] == 0))
			{
				epsilon_jk_kB = std::sqrt(specie_j[epsilon_kB] * specie_k[epsilon_kB]);
				d_jk = (specie_j[d_] + specie_k[d_]) / 2.0;
				mue_jk_sqr = specie_j[mue] * specie_k[mue];
			}
			else
			{
				double epsilon_n_kB = 0, epsilon_p_kB = 0, alpha_n = 0, mue_p = 0, d_n = 0, d_p = 0;
				if (specie_k[mue] > 0 && specie_j[mue] == 0)
					epsilon_n_kB = specie_j[epsilon_kB], epsilon_p_kB = specie_k[epsilon_kB], alpha_n = specie_j[alpha], d_n = specie_j[d_], d_p = specie_k[d_], mue_p = specie_k[mue];
				if (specie_j[mue] > 0 && specie_k[mue] == 0)
					epsilon_n_kB = specie_k[epsilon_kB], epsilon_p_kB = specie_j[epsilon_kB], alpha_n = specie_k[alpha], d_n = specie_k[d_], d_p = specie_j[d_], mue_p = specie_j[mue];
				const double alpha_n_star = alpha_n / std::pow(d_n, 3.0);
				const double mue_p_star = mue_p / std::pow(epsilon_p_kB * kB, 0.5) / std::pow(d_p, 1.5) * 1.0e-6;
				const double ksi = 1.0 + 0.25 * alpha_n_star * mue_p_star * std::sqrt(epsilon_p_kB / epsilon_n_kB);
				epsilon_jk_kB = ksi * ksi * std::sqrt(epsilon_n_kB * epsilon_p_kB);
				d_jk = std::pow(ksi, -1.0 / 6.0) * (specie_j[d_] + specie_k[d_]) / 2.0;
				mue_jk_sqr = 0.0;
			}
			const double T_jk_star = T / epsilon_jk_kB;
			const double delta_jk_star = 0.5 * mue_jk_sqr / d_jk / d_jk / d_jk / epsilon_jk_kB / kB * 1.0e-12;
			const double W_jk = specie_k[WI] * specie_j[WI] / (specie_k[WI] + specie_j[WI]) / NA * 1.0e3;
			const double Omega1 = Omega_interpolated(T_jk_star, delta_jk_star, 1);
			const double PPP = PP * 10.0;
			return 3.0 * std::sqrt(2.0 * pi * std::pow(T * kB, 3.0) / W_jk) / (16.0 * PPP * pi * d_jk * d_jk * Omega1) * 1.0e16;
		}
		double thermal_conductivities(const double *specie, double T, double PP) const
		{
			const double Cv_trans = 1.5 * universal_gas_const;
			double Cv_rot = 0, Cv_vib = 0;
			const int id = int(specie[SID]);
			const double Cpi = HeatCapacity_NASA(S.Hia.data(), T, S.Ri[id], id);
			const double Cv = Cpi * specie[WI] * 1.0e3 - universal_gas_const;
			switch (int(specie[geo]))
			{
			case 0: Cv_rot = 0.0, Cv_vib = 0.0; break;
			case 1: Cv_rot = 1.0 * universal_gas_const, Cv_vib = Cv - 2.5 * universal_gas_const; break;
			case 2: Cv_rot = 1.5 * universal_gas_const, Cv_vib = Cv - 3.0 * universal_gas_const; break;
			}
			const double rho = PP * specie[WI] / T / universal_gas_const;
			const double Dkk = Dkj(specie, specie, T, PP);
			const double visc = viscosity(specie, T);
			const double f_vib = rho * Dkk / (visc * 10.0);
			const double Zrot = specie[Zrot_298] * ZrotFunc(specie[epsilon_kB] / 298.0) / ZrotFunc(specie[epsilon_kB] / T);
			const double Aa = 2.5 - f_vib, Bb = Zrot + 2.0 * (5.0 * Cv_rot / 3.0 / (universal_gas_const) + f_vib) / pi;
			const double f_trans = 2.5 * (1.0 - 2.0 * Cv_rot * Aa / pi / Cv_trans / Bb);
			const double f_rot = f_vib * (1.0 + 2.0 * Aa / pi / Bb);
			return visc * (f_trans * Cv_trans + f_rot * Cv_rot + f_vib * Cv_vib) / specie[WI] * 1.0e-3;
		}
		// Fitting (viscfit.cpp:317-341): indicator 0 viscosity, 1 thermal conductivity, 2 binary diffusion
		void Fitting(const std::vector<double> &TT, const double *specie_k, const double *specie_j, double *aa, int indicator) const
		{
			const int mm = int(TT.size());
			std::vector<double> b(mm);
			std::vector<std::array<double, 4>> AA(mm);
			for (int ii = 0; ii < mm; ii++)
			{
				switch (indicator)
				{
				case 0: b[ii] = std::log(viscosity(specie_k, TT[ii])); break;
				case 1: b[ii] = std::log(thermal_conductivities(specie_k, TT[ii], 1.0)); break;
				case 2: b[ii] = std::log(Dkj(specie_k, specie_j, TT[ii], 1.0)); break;
				}
				for (int jj = 0; jj < 4; jj++)
					AA[ii][jj] = std::pow(std::log(TT[ii]), jj);
			}
			Solve_Overdeter_equations(AA, b, aa);
		}
	};

	// thermal.cpp:131-165 (transport part) + Setup::GetFitCoefficient (viscfit.cpp:148-190)
	void Setup::GetFitCoefficient()
	{
		const int NS = num_species;
		species_chara.assign(NS * 9, 0.0);
		const std::string spath = WorkDir + "/runtime.dat/transport_data.dat";
		std::ifstream fint(spath);
		if (!fint)
			throw std::runtime_error("cannot open " + spath);
		std::vector<std::string> toks;
		std::string t;
		while (fint >> t)
			toks.push_back(t);
		for (int i = 0; i < NS; i++)
		{
			bool found = false;
			for (size_t p = 0; p < toks.size(); p++)
			{
				if (toks[p] == "*END")
					break;
				if (toks[p] == species_name[i] && p + 6 < toks.size())
				{
					for (int m = 0; m < 6; m++)
						species_chara[i * 9 + m] = std::stod(toks[p + 1 + m]);
					found = true;
					break;
				}
			}
			if (!found)
				throw std::runtime_error("species " + species_name[i] + " not found in transport_data.dat");
			species_chara[i * 9 + 6] = Wi[i]; // kg/mol (thermal.cpp:158-159)
			species_chara[i * 9 + 7] = 0, species_chara[i * 9 + 8] = i;
		}
		TransportFit F(*this);
		fit_visc.assign(NS * 4, 0.0), fit_therm.assign(NS * 4, 0.0), fit_Dkj.assign(NS * NS * 4, 0.0);
		for (int k = 0; k < NS; k++)
		{
			const double *specie_k = &species_chara[k * 9];
			F.Fitting(Tnode, specie_k, specie_k, &fit_visc[k * 4], 0);
			F.Fitting(Tnode, specie_k, specie_k, &fit_therm[k * 4], 1);
			for (int j = 0; j < NS; j++)
			{
				const double *specie_j = &species_chara[j * 9];
				if (k <= j)
					F.Fitting(Tnode, specie_k, specie_j, &fit_Dkj[(k * NS + j) * 4], 2);
				else
					for (int n = 0; n < 4; n++)
						fit_Dkj[(k * NS + j) * 4 + n] = fit_Dkj[(j * NS + k) * 4 + n];
			}
		}
	}

	xf_transport Setup::transport() const
	{
		xf_transport tr{};
		tr.visc = Visc ? 1 : 0, tr.visc_heat = Visc_Heat ? 1 : 0, tr.visc_diffu = Visc_Diffu ? 1 : 0;
		tr.fit_visc = fit_visc.data(), tr.fit_therm = fit_therm.data(), tr.fit_Dkj = fit_Dkj.data(), tr.Wi = Wi.data();
		tr.Yil_limiter = Yil_limiter, tr.Dim_limiter = Dim_limiter, tr.dim_max0 = diffu_dim_max0;
		return tr;
	}
} // namespace xfh
