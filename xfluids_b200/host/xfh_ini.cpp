// xfh_ini.cpp -- initial-condition hooks: the five inert samples of the BASELINE configs, restated for the host from
// src/solver_Ini/sample/<case>/ini_sample.hpp (InitialStatesKernel + InitialUFKernel).  Selected at run time by
// Setup::sample (INIT_SAMPLE in the reference).  All cells incl. ghosts are written: ghost values matter because the
// Inflow BC never touches them (the jet enters through x-min ghosts that keep their initial state forever).
// U is the reference's AoS [cell][Emax]; T is the Newton warm start of every cell (0 for single-component cases).
#include <cmath>
#include "xfh_setup.hpp"

namespace xfh
{
	namespace
	{
		const double pi = 3.1415926535897932384626433832795;

		struct Cell
		{
			double rho = 0, p = 0, T = 0, u = 0, v = 0, w = 0;
			double yi[8];
		};

		// the common tail of every multi-component sample: rho from the ideal-gas law, total energy from NASA enthalpy
		inline void finish_cop(const Setup &s, Cell &c, double *U)
		{
			const double R = get_CopR(s, c.yi);
			c.rho = c.p / R / c.T;
			const double h = get_Coph(s, c.yi, c.T);
			U[4] = c.rho * (h + 0.5 * (c.u * c.u + c.v * c.v + c.w * c.w)) - c.p;
			U[0] = c.rho, U[1] = c.rho * c.u, U[2] = c.rho * c.v, U[3] = c.rho * c.w;
			for (int ii = 5; ii < s.Emax; ii++)
				U[ii] = c.rho * c.yi[ii - 5];
		}
		inline void finish_nocop(const Setup &s, const Cell &c, double *U)
		{
			U[0] = c.rho, U[1] = c.rho * c.u, U[2] = c.rho * c.v, U[3] = c.rho * c.w;
			U[4] = c.p / (s.ncop_gamma - 1.0) + 0.5 * c.rho * (c.u * c.u + c.v * c.v + c.w * c.w);
		}
	}

	int InitialCondition(const Setup &s, double *U, double *T)
	{
		const xf_block &bl = s.bl;
		const int Xmax = bl.Xmax, Ymax = bl.Ymax, Zmax = bl.Zmax, E = s.Emax, NS = s.num_species;
		const int Bx = bl.Bwidth_X, By = bl.Bwidth_Y, Bz = bl.Bwidth_Z;
		const double dx = bl.dx, dy = bl.dy, dz = bl.dz;
		enum { ST, VORTEX, RIEMANN, SBI, JET } kind;
		if (s.sample == "1d-insert-st") kind = ST;
		else if (s.sample == "2d-euler-vortex") kind = VORTEX;
		else if (s.sample == "2d-riemann-shocks") kind = RIEMANN;
		else if (s.sample == "shock-bubble") kind = SBI;
		else if (s.sample == "2d-under-expanded-jet" || s.sample == "3d-under-expanded-jet") kind = JET;
		else return -1;
		if ((kind == VORTEX || kind == RIEMANN) == s.cop)
			return -2; // sample / mixture mismatch
		const IniShape &ini = s.ini;

#pragma omp parallel for collapse(2) schedule(static)
		for (int k = 0; k < Zmax; k++)
			for (int j = 0; j < Ymax; j++)
				for (int i = 0; i < Xmax; i++)
				{
					const size_t id = size_t(Xmax) * Ymax * k + size_t(Xmax) * j + i;
					// cell centre (every ini_sample.hpp, e.g. insert-st/ini_sample.hpp:30-33)
					const double x = bl.DimX ? (i - Bx + s.myMpiPos_x * (Xmax - Bx - Bx)) * dx + 0.5 * dx + s.Domain_xmin : 0.0;
					const double y = bl.DimY ? (j - By + s.myMpiPos_y * (Ymax - By - By)) * dy + 0.5 * dy + s.Domain_ymin : 0.0;
					const double z = bl.DimZ ? (k - Bz + s.myMpiPos_z * (Zmax - Bz - Bz)) * dz + 0.5 * dz + s.Domain_zmin : 0.0;
					Cell c;
					double *Uc = U + size_t(E) * id;
					switch (kind)
					{
					case ST: // 1D-X-Y-Z/insert-st/ini_sample.hpp:35-70
					{
						for (int n = 0; n < NS; n++)
							c.yi[n] = s.species_ratio_out[n];
						if (bl.DimX) c.T = x < 0.05 ? 400 : 1200, c.p = x < 0.05 ? 8000 : 80000;
						if (bl.DimY) c.T = y < 0.05 ? 400 : 1200, c.p = y < 0.05 ? 8000 : 80000;
						if (bl.DimZ) c.T = z < 0.05 ? 400 : 1200, c.p = z < 0.05 ? 8000 : 80000;
						finish_cop(s, c, Uc);
					}
					break;
					case VORTEX: // 2D-EulerVortex/ini_sample.hpp:32-60
					{
						const double beta = 5.0, beta2 = beta * beta;
						const double _x_vortex = (x - 0.5 * s.Domain_length), _y_vortex = (y - 0.5 * s.Domain_width);
						const double r2 = (_x_vortex * _x_vortex + _y_vortex * _y_vortex);
						const double inter1 = beta2 * std::exp(1.0 - r2) / 28.0 / pi / pi;
						const double inter2 = 0.5 * beta * std::exp(0.5 * (1.0 - r2)) / pi;
						c.rho = std::pow(1.0 - inter1, 2.5);
						c.p = std::pow(c.rho, 1.4);
						c.u = 1.0 - inter2 * _y_vortex;
						c.v = 1.0 + inter2 * _x_vortex;
						Uc[0] = c.rho, Uc[1] = c.rho * c.u, Uc[2] = c.rho * c.v, Uc[3] = c.rho * c.w;
						const double Gamma_tmp = 1.4;
						Uc[4] = c.p / (Gamma_tmp - 1.0) + 0.5 * c.rho * (c.u * c.u + c.v * c.v + c.w * c.w);
					}
					break;
					case RIEMANN: // 2D-Riemann/shocks-interaction/ini_sample.hpp:40-67 (p = rho_k and Domain_height, as written)
					{
						const double rho1 = 1.1, u1 = 0.0, v1 = 0.0, rho2 = 0.5065, u2 = 0.8939, v2 = 0.0;
						const double rho3 = 1.1, u3 = 0.8939, v3 = 0.8939, rho4 = 0.5065, u4 = 0.0, v4 = 0.8939;
						if (y > 0.5 * s.Domain_height)
						{
							if (x > 0.5 * s.Domain_length) c.rho = rho1, c.p = rho1, c.u = u1, c.v = v1;
							else c.rho = rho2, c.p = rho2, c.u = u2, c.v = v2;
						}
						else
						{
							if (x < 0.5 * s.Domain_length) c.rho = rho3, c.p = rho3, c.u = u3, c.v = v3;
							else c.rho = rho4, c.p = rho4, c.u = u4, c.v = v4;
						}
						finish_nocop(s, c, Uc);
					}
					break;
					case SBI: // shock-bubble-intera/ini_sample.hpp:26-93 (states) + :115-160 (conservatives)
					{
						if (x < ini.blast_center_x)
							c.T = ini.blast_T_in, c.p = ini.blast_pressure_in, c.u = ini.blast_u_in, c.v = ini.blast_v_in, c.w = ini.blast_w_in;
						else
							c.T = ini.blast_T_out, c.p = ini.blast_pressure_out, c.u = ini.blast_u_out, c.v = ini.blast_v_out, c.w = ini.blast_w_out;
						double *xi = c.yi;
						for (int nn = 0; nn < NS; nn++)
							xi[nn] = 0.0;
						double dy_ = 0.0, tmp = 0.0;
						if (bl.DimX) tmp = (x - ini.cop_center_x) * (x - ini.cop_center_x), dy_ += tmp * ini._xa2;
						if (bl.DimY) tmp = (y - ini.cop_center_y) * (y - ini.cop_center_y), dy_ += tmp * ini._yb2;
						if (bl.DimZ) tmp = (z - ini.cop_center_z) * (z - ini.cop_center_z), dy_ += tmp * ini._zc2;
						dy_ = std::sqrt(dy_) - 1.0;
						const double xrest = 1.0, ff = 1.0e-4, dd = 0.5 * (xrest - 2.0 * ff);
						xi[NS - 1] = 0.0;
						xi[NS - 2] = dd * (std::tanh(dy_ * ini.C)) + 0.5;
						xi[0] = 0.29 * (xrest - xi[NS - 2]);      // H2
						xi[1] = 0.15 * (xrest - xi[NS - 2]);      // O2
						xi[NS - 3] = 0.56 * (xrest - xi[NS - 2]); // Xe
						get_yi(xi, s.Wi.data(), NS);
						if (s.RSources)
						{
							const int NUM_COP = NS - 1;
							const double xre = 1.0e-15, ratios = xre * double(NUM_COP - 3) * 0.25;
							for (int n1 = 0; n1 < NUM_COP; n1++) xi[n1] -= ratios;
							for (int nn = 2; nn < NUM_COP - 2; nn++) xi[nn] = xre;
						}
						finish_cop(s, c, Uc);
					}
					break;
					case JET: // under-expanded-jet/ini_sample.hpp:42-160
					{
						for (int n = 0; n < NS; n++)
							c.yi[n] = s.species_ratio_out[n];
						const bool core = (i <= 3) && (-0.015 < y && y < 0.015);
						if (core)
							c.p = 10.0 * 101325.0, c.T = 1000.0, c.yi[0] = 0.0087, c.yi[1] = 0.2329, c.yi[2] = 0.7584;
						else
							c.p = 1.0 * 101325.0, c.T = 300.0, c.yi[0] = 0.0, c.yi[1] = 0.233, c.yi[2] = 0.767;
						const double R = get_CopR(s, c.yi);
						const double rho = c.p / R / c.T;
						const double Gamma_m = get_CopGamma(s, c.yi, c.T);
						const double cs = std::sqrt(c.p / rho * Gamma_m);
						if (i <= 3)
						{
							if (-0.015 < y && y < 0.015) c.u = cs;
							else if (-0.015 * 25 < y && y < 0.015 * 25) c.u = 0.0575 * cs;
							else c.u = 0.0 * cs;
						}
						else
							c.u = 0.0 * cs;
						finish_cop(s, c, Uc);
					}
					break;
					}
					T[id] = c.T;
				}
		return 0;
	}
} // namespace xfh
