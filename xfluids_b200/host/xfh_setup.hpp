// xfh_setup.hpp -- host-side mirror of the reference's struct Setup (src/read_ini/setupini.h:19-83): run-time inputs in
// XFluids' own formats (settings/*.json keys, runtime.dat/<mixture>/species_list.dat, runtime.dat/thermal_dynamics.dat,
// the -run/-mpi/-mpi-s/... command-line overrides) and the derived Block metrics, with the same formulas and the
// same evaluation order as the reference so that every derived number is bit-identical.
//
// What is compile-time in the reference (INIT_SAMPLE, MIXTURE_MODEL, WENO_ORDER, ARTIFICIAL_VISC_TYPE; cmake/*.cmake)
// is run-time here: optional JSON section "b200" {sample, mixture, weno, artificial, fp_mode} or -sample= -mixture=
// -weno= -alpha= -fp= -pp= -cfl= on the command line (-pp overrides equations.PositivityPreserving, -cfl run.CFLnumber).  The reference ignores unknown JSON keys, so files stay interchangeable.
#pragma once
#include <array>
#include <string>
#include <vector>
#include "../../include/xfluids_b200.h"
#include "xfh_json.hpp"

namespace xfh
{
	// reference IniShape (global_setup.h:283-316), the fields the five inert samples read
	struct IniShape
	{
		int cop_type = 0, blast_type = 1;
		double blast_center_x = 0.5, blast_center_y = 0.5, blast_center_z = 0.5;
		double Ma = 0, tau_H = 0;
		double blast_density_in = 0, blast_pressure_in = 0, blast_T_in = 298.15, blast_u_in = 0, blast_v_in = 0, blast_w_in = 0;
		double blast_density_out = 0, blast_pressure_out = 0, blast_T_out = 298.15, blast_u_out = 0, blast_v_out = 0, blast_w_out = 0;
		double blast_c_out = 0, blast_gamma_out = 0;
		double cop_center_x = 0, cop_center_y = 0, cop_center_z = 0;
		double cop_density_in = 0, cop_pressure_in = 0, cop_T_in = 0;
		double xa = 0, yb = 0, zc = 0, C = 0, _xa2 = 0, _yb2 = 0, _zc2 = 0;
	};

	struct OutStamp
	{
		double time;
		std::string spec;
	};

	struct Setup
	{
		// ---- selections that are compile-time macros in the reference
		std::string sample = "1d-insert-st"; // INIT_SAMPLE
		std::string mixture = "NO-COP";      // MIXTURE_MODEL (directory under runtime.dat/)
		bool cop = false, ghost_species = false;
		int num_species = 1, Emax = 5;
		int weno = 5, artificial = 2, fp_mode = 0;
		double ncop_gamma = 1.4;

		// ---- run (read_json.cpp:33-52)
		std::string OutputDir = "output";
		std::string select_dv = "b200"; // SelectDv of the reference's build (part of every output file name, XFLUIDS.cpp:19); -dv= overrides
		int nStepmax = 10, nStepmax_json = 10;
		std::vector<OutStamp> OutTimeStamps;
		bool write_checkpoint = false;
		int RcalInterval = 100;

		// ---- mesh / block (Block, global_setup.h:162-190; Setup::init, iniset.cpp:290-369)
		xf_block bl{};
		double Domain_xmin = 0, Domain_ymin = 0, Domain_zmin = 0, Domain_length = 1, Domain_width = 1, Domain_height = 1;
		int mx = 1, my = 1, mz = 1, myMpiPos_x = 0, myMpiPos_y = 0, myMpiPos_z = 0;
		double dl = 0;
		int Boundarys[6] = {2, 2, 2, 2, 2, 2};
		bool RSources = false, PositivityPreserving = false;
		// viscous terms: Visc / Visc_Heat / Visc_Diffu are compile-time in the reference (cmake/init_options.cmake:81-92; shock-bubble preset:
		// all three ON, fourth-order discretisation), run-time here: JSON "b200": {"visc": 1, "visc_heat": 1, "visc_diffu": 1} or -visc=0|1
		// (all three), -visc-heat=, -visc-diffu= on the command line.  -diffu-mpi=1 selects the reference MPI build's diffusion limiter.
		bool Visc = false, Visc_Heat = false, Visc_Diffu = false;
		double Yil_limiter = 0, Dim_limiter = 0, Yil_limiter_json = 1.0E10, Dim_limiter_json = 2.0E-3, diffu_dim_max0 = 0.0;
		std::vector<double> Tnode{273.15, 500.0, 750.0, 1000.0, 1250.0, 1500.0, 1750.0, 2000.0, 2250.0, 2500.0, 2750.0, 3000.0, 5000.0};
		std::vector<double> species_chara, fit_visc, fit_therm, fit_Dkj; // [NS*9], [NS*4], [NS*4], [NS*NS*4]
		size_t bytes = 0, cellbytes = 0;

		// ---- fluids
		IniShape ini;
		bool mach_shock = false;
		std::vector<std::string> species_name;
		std::vector<double> Hia, Hib, Wi, _Wi, Ri, species_ratio_in, species_ratio_out, xi_in, xi_out;

		std::string WorkDir;    // directory that contains runtime.dat/
		int myRank = 0, nRanks = 1;
		Json j_conf;
		std::vector<std::string> args;

		// Setup::Setup (constructor.cpp:12-84): ReWrite -> ReadSpecies/ReadThermal -> init -> (MpiTrans) rank position
		Setup(const std::string &json_path, const std::vector<std::string> &cli, const std::string &workdir, int rank = 0, int nranks = 1);

		void ReadIni();      // iniset.cpp:8-67
		void ReWrite();      // iniset.cpp:72-286
		void ReadSpecies();  // thermal.cpp:6-32
		void ReadThermal();  // thermal.cpp:36-175
		void init();         // iniset.cpp:290-369
		bool Mach_Shock();   // viscfit.cpp:8-140
		void GetFitCoefficient(); // viscfit.cpp:148-190 + the transport part of ReadThermal (thermal.cpp:131-165)
		xf_transport transport() const;
		void print() const;

		xf_thermal thermal() const;
		xf_scheme scheme() const { return xf_scheme{weno, artificial, fp_mode, PositivityPreserving ? 1 : 0}; }
		size_t ncells() const { return size_t(bl.Xmax) * bl.Ymax * bl.Zmax; }
		// neighbour-aware boundary list for this rank's slab: BC_COPY on interior z faces (mpiPacks.cpp:44-72)
		void rank_boundarys(int out[6]) const;

	private:
		double C_json_ = 0; // bubble_shape_x * bubble_boundary_width (read_json.cpp:148)
		std::vector<std::string> match(const std::string &opt) const;
	};

	// reference constants (global_setup.h:38-54)
	constexpr double Ru = (6.02214076e26 * 1.380649e-23) * 1.0E-3;

	// host thermo used by the initial conditions and the shock jump (Thermo_device.h / Mixing_device.h restated for the host)
	double HeatCapacity_NASA(const double *Hia, double T0, double Ri, int n);
	double get_Enthalpy_NASA(const double *Hia, const double *Hib, double T0, double Ri, int n);
	double get_CopR(const Setup &s, const double *yi);
	double get_CopCp(const Setup &s, const double *yi, double T);
	double get_CopGamma(const Setup &s, const double *yi, double T);
	double get_Coph(const Setup &s, const double *yi, double T);
	void get_yi(double *xi, const double *Wi, int ns);

	// InitializeFluidStates (solver_Ini/Ini_block.cpp:4-40) for the five inert samples: fills U (AoS [N][Emax], the
	// reference layout) and the Newton warm start T for ALL cells incl. ghosts.  Returns 0, or -1 for an unknown sample.
	int InitialCondition(const Setup &s, double *U, double *T);
} // namespace xfh
