// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE, not product code.
//
// A replacement for the reference's src/main.cpp (main.cpp:30-54) that links against the
// UNMODIFIED reference objects (Setup, XFLUIDS, Fluid) built by oracle/build_ref.sh and
// drives them through the reference's own public methods in the reference's own order
// (XFLUIDS::Evolution inner loop, XFLUIDS.cpp:167-294, minus file output / checkpoints),
// dumping raw state so the CUDA path and the C restatement (oracle/xf_oracle.cpp) can be
// compared with the reference's CPU results.
//
// Controls (environment, so that the reference's own CLI parser is left alone):
//   XF_NSTEPS=<n>          number of time steps to run (default: nStepmax from -run/JSON)
//   XF_DUMP_DIR=<dir>      where to write dumps (default: no dumps)
//   XF_DUMP_STEPS=a,b,c    dump U (AoS incl. ghosts, reference checkpoint payload layout
//                          XFLUIDS.cpp:658-687) after these step counts; 0 = initial state
//   XF_DUMP_STAGE=1        additionally dump the intermediates of step 1 / RK stage 1:
//                          primitives after UpdateStates, LU + wall fluxes after ComputeLU
//   XF_DUMP_IC=1           dump ic_U / ic_T: the raw initial condition before BC + UpdateStates
//   XF_DUMP_T=1            dump T (the Newton warm-start state) with every U dump
//   XF_OUTPUT=<n>          after the loop, write the reference's field output (XFLUIDS::Output) of the final state in the format
//                          of output stamp number n into ./output
// Prints one line "ORACLE_TIMING ..." with the wall time of the time loop.
#include "global_class.h"
#include <chrono>
#include <set>
#include <sstream>

static void dump_raw(const std::string &path, const void *p, size_t bytes)
{
	std::ofstream f(path, std::ios::binary);
	f.write(reinterpret_cast<const char *>(p), bytes);
}

static std::string envs(const char *k, const char *d = "")
{
	const char *v = std::getenv(k);
	return v ? std::string(v) : std::string(d);
}

int main(int argc, char *argv[])
{
	Setup setup(argc, argv);
	XFLUIDS solver(setup);
	sycl::queue &q = setup.q;
	solver.AllocateMemory(q);
	// InitialCondition() also tries to read a checkpoint from OutputDir; none is ever written by this driver
	solver.InitialCondition(q);
	const Block bl = setup.BlSz;
	const size_t N = size_t(bl.Xmax) * bl.Ymax * bl.Zmax;
	Fluid *fl = solver.fluids[0];
	if (!envs("XF_DUMP_DIR").empty() && (envs("XF_DUMP_STAGE") == "1" || envs("XF_DUMP_IC") == "1"))
	{ // raw initial condition: U and the Newton warm-start T exactly as the sample kernel left them
		dump_raw(envs("XF_DUMP_DIR") + "/ic_U.bin", fl->d_U, N * Emax * sizeof(real_t));
		dump_raw(envs("XF_DUMP_DIR") + "/ic_T.bin", fl->d_fstate.T, N * sizeof(real_t));
	}
	if (Visc && !envs("XF_DUMP_DIR").empty())
	{ // transport fits of Setup::GetFitCoefficient (viscfit.cpp:148-190): [NS][4] viscosity, [NS][4] conductivity, [NS*NS][4] binary diffusion,
	  // and the species characteristics they were made from
		Thermal &th = setup.h_thermal;
		dump_raw(envs("XF_DUMP_DIR") + "/fit_visc.bin", th.fitted_coefficients_visc[0], NUM_SPECIES * order_polynominal_fitted * sizeof(real_t));
		dump_raw(envs("XF_DUMP_DIR") + "/fit_therm.bin", th.fitted_coefficients_therm[0], NUM_SPECIES * order_polynominal_fitted * sizeof(real_t));
		dump_raw(envs("XF_DUMP_DIR") + "/fit_Dkj.bin", th.Dkj_matrix[0], NUM_SPECIES * NUM_SPECIES * order_polynominal_fitted * sizeof(real_t));
		dump_raw(envs("XF_DUMP_DIR") + "/species_chara.bin", th.species_chara, NUM_SPECIES * SPCH_Sz * sizeof(real_t));
	}
	solver.BoundaryCondition(q);
	solver.UpdateStates(q);

	int nsteps = setup.nStepmax;
	if (!envs("XF_NSTEPS").empty())
		nsteps = std::atoi(envs("XF_NSTEPS").c_str());
	const std::string ddir = envs("XF_DUMP_DIR");
	const bool dump_stage = envs("XF_DUMP_STAGE") == "1";
	const bool dump_T = envs("XF_DUMP_T") == "1";
	std::set<int> dump_steps;
	{
		std::stringstream ss(envs("XF_DUMP_STEPS"));
		std::string tok;
		while (std::getline(ss, tok, ','))
			if (!tok.empty())
				dump_steps.insert(std::atoi(tok.c_str()));
	}

	std::ofstream meta;
	if (!ddir.empty())
	{
		meta.open(ddir + "/meta.txt");
		meta.precision(17);
		meta << "Xmax " << bl.Xmax << "\nYmax " << bl.Ymax << "\nZmax " << bl.Zmax << "\nX_inner " << bl.X_inner
			 << "\nY_inner " << bl.Y_inner << "\nZ_inner " << bl.Z_inner << "\nBwidth_X " << bl.Bwidth_X << "\nBwidth_Y "
			 << bl.Bwidth_Y << "\nBwidth_Z " << bl.Bwidth_Z << "\nEmax " << Emax << "\nNUM_SPECIES " << NUM_SPECIES
			 << "\ndx " << bl.dx << "\ndy " << bl.dy << "\ndz " << bl.dz << "\nCFL " << bl.CFLnumber << "\n";
	}
	auto dumpU = [&](int step)
	{
		if (ddir.empty() || !dump_steps.count(step))
			return;
		dump_raw(ddir + "/U_step" + std::to_string(step) + ".bin", fl->d_U, N * Emax * sizeof(real_t));
		if (dump_T)
			dump_raw(ddir + "/T_step" + std::to_string(step) + ".bin", fl->d_fstate.T, N * sizeof(real_t));
	};
	auto dumpPrims = [&](const std::string &tag)
	{
		FlowData &f = fl->d_fstate;
		const std::pair<const char *, real_t *> sc[] = {{"rho", f.rho}, {"p", f.p}, {"u", f.u}, {"v", f.v}, {"w", f.w}, {"T", f.T}, {"H", f.H}, {"c", f.c}, {"gamma", f.gamma}, {"e", f.e}, {"Cp", f.Cp}, {"R", f.Ri}};
		for (auto &s : sc)
			dump_raw(ddir + "/" + tag + "_" + s.first + ".bin", s.second, N * sizeof(real_t));
		dump_raw(ddir + "/" + tag + "_y.bin", f.y, N * NUM_SPECIES * sizeof(real_t));
	};

	dumpU(0);
	if (!ddir.empty() && dump_stage)
		dumpPrims("ini");

	// ---- time loop: XFLUIDS::Evolution (XFLUIDS.cpp:167-294) without output/checkpoint/ARA files
	Setup::adv_nd[0].resize(1), Setup::sbm_id = 0;
	int TimeLoop = 0;
	bool stop = false, error = false;
	auto t0 = std::chrono::high_resolution_clock::now();
	while (TimeLoop < (int)setup.OutTimeStamps.size() && !stop)
	{
		real_t target_t = (solver.physicalTime < setup.OutTimeStamps[TimeLoop].time) ? setup.OutTimeStamps[TimeLoop].time : setup.OutTimeStamps[TimeLoop++].time;
		while (solver.physicalTime < target_t)
		{
			solver.Iteration++;
			solver.dt = solver.ComputeTimeStep(q);
			if (solver.physicalTime + solver.dt > target_t)
				solver.dt = target_t - solver.physicalTime;
			solver.physicalTime += solver.dt;
			if (meta.is_open())
				meta << "dt " << solver.Iteration << " " << solver.dt << " " << solver.physicalTime << "\n";

			if (dump_stage && solver.Iteration == 1 && !ddir.empty())
			{ // RungeKuttaSP3rd(flag=1) taken apart (XFLUIDS.cpp:448-470) so the intermediates can be written
				solver.BoundaryCondition(q, 0);
				dump_raw(ddir + "/s1_Ubc.bin", fl->d_U, N * Emax * sizeof(real_t));
				error = solver.UpdateStates(q, 0, solver.physicalTime, solver.Iteration, "_RK1");
				dump_raw(ddir + "/s1_Uprim.bin", fl->d_U, N * Emax * sizeof(real_t));
				dumpPrims("s1");
				solver.ComputeLU(q, 0);
				dump_raw(ddir + "/s1_LU.bin", fl->d_LU, N * Emax * sizeof(real_t));
				dump_raw(ddir + "/s1_Fwx.bin", fl->d_wallFluxF, N * Emax * sizeof(real_t));
				dump_raw(ddir + "/s1_Fwy.bin", fl->d_wallFluxG, N * Emax * sizeof(real_t));
				dump_raw(ddir + "/s1_Fwz.bin", fl->d_wallFluxH, N * Emax * sizeof(real_t));
				if (Visc)
				{ // intermediates of the viscous block of GetLU (ConVenction_block.hpp:424-575)
					FlowData &f = fl->d_fstate;
					dump_raw(ddir + "/s1_visc.bin", f.viscosity_aver, N * sizeof(real_t));
					dump_raw(ddir + "/s1_therm.bin", f.thermal_conduct_aver, N * sizeof(real_t));
					dump_raw(ddir + "/s1_Dkm.bin", f.Dkm_aver, N * NUM_SPECIES * sizeof(real_t));
					dump_raw(ddir + "/s1_hi.bin", f.hi, N * NUM_SPECIES * sizeof(real_t));
					for (int m = 0; m < 9; m++)
						dump_raw(ddir + "/s1_Vde" + std::to_string(m) + ".bin", f.Vde[m], N * sizeof(real_t));
				}
				error = error || solver.EstimateNAN(q, solver.physicalTime, solver.Iteration, 0, 1);
				solver.UpdateU(q, 1);
				dump_raw(ddir + "/s1_U1.bin", fl->d_U1, N * Emax * sizeof(real_t));
				error = error || solver.RungeKuttaSP3rd(q, 0, solver.Iteration, solver.physicalTime, 2);
				error = error || solver.RungeKuttaSP3rd(q, 0, solver.Iteration, solver.physicalTime, 3);
			}
			else
				error = solver.SinglePhaseSolverRK3rd(q, 0, solver.Iteration, solver.physicalTime);

			// ARA bookkeeping reduced to what keeps Setup::adv_nd indexable (XFLUIDS.cpp:262-283)
			Setup::sbm_id = 0;
			Setup::adv_id = (solver.Iteration < (int)Setup::adv_nd.size() && Setup::adv_push) ? solver.Iteration : 0;
			if ((int)Setup::adv_nd.size() == solver.Iteration && Setup::adv_push)
				Setup::adv_nd[0].erase(Setup::adv_nd[0].begin());
			Setup::adv_push = Setup::adv_id;

			dumpU(solver.Iteration);
			if (error || solver.Iteration >= nsteps)
			{
				stop = true;
				break;
			}
		}
	}
	if (!envs("XF_OUTPUT").empty())
	{ // one field output of the final state in the format of output stamp number XF_OUTPUT (XFLUIDS.cpp:309 does this with the last stamp)
		// (XFLUIDS keeps a private COPY of Setup, whose stamps AllocateMemory initialised, XFLUIDS.cpp:611-612; the same public call here)
		const size_t which = std::min<size_t>(std::atoi(envs("XF_OUTPUT").c_str()), setup.OutTimeStamps.size() - 1);
		setup.OutTimeStamps[which].Initialize(setup.BlSz, setup.species_name, fl->h_fstate);
		solver.Output(q, setup.OutTimeStamps[which].Reinitialize(solver.physicalTime, std::to_string(solver.Iteration)));
	}
	double secs = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
	double cells = double(bl.X_inner) * bl.Y_inner * bl.Z_inner;
	std::cout.precision(10);
	std::cout << "\nORACLE_TIMING steps=" << solver.Iteration << " seconds=" << secs << " mcell_stage_per_s=" << cells * 3.0 * solver.Iteration / secs / 1e6
			  << " time=" << solver.physicalTime << " error=" << int(error) << std::endl;
	if (meta.is_open())
		meta << "final_time " << solver.physicalTime << "\nerror " << int(error) << "\n";
	return error ? 1 : 0;
}
