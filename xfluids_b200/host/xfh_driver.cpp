#include "xfh_driver.hpp"
#include <chrono>
#include <cstdio>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <stdexcept>

namespace xfh
{
#define XFCK(call)                                                                                 \
	do                                                                                             \
	{                                                                                              \
		int rc__ = (call);                                                                         \
		if (rc__ != XF_OK)                                                                         \
			throw std::runtime_error(std::string(#call) + " failed: " + xf_last_error());          \
	} while (0)

	Fluid::Fluid(Setup &setup, int device) : Fs(setup)
	{
		xf_thermal th = Fs.thermal();
		xf_scheme sc = Fs.scheme();
		XFCK(xf_create(&Fs.bl, &th, &sc, device, &ctx));
		if (Fs.Visc)
		{
			xf_transport tr = Fs.transport();
			XFCK(xf_set_transport(ctx, &tr));
		}
		Fs.rank_boundarys(BCs);
	}
	Fluid::~Fluid()
	{
		if (ctx)
		{
			if (d_U) xf_field_free(ctx, d_U);
			if (d_U1) xf_field_free(ctx, d_U1);
			if (d_LU) xf_field_free(ctx, d_LU);
			xf_destroy(ctx);
		}
	}
	void Fluid::AllocateFluidMemory()
	{
		XFCK(xf_field_alloc(ctx, &d_U));
		XFCK(xf_field_alloc(ctx, &d_U1));
		XFCK(xf_field_alloc(ctx, &d_LU));
	}
	void Fluid::InitialU()
	{
		std::vector<double> U(Fs.ncells() * Fs.Emax), T(Fs.ncells());
		if (InitialCondition(Fs, U.data(), T.data()))
			throw std::runtime_error("unknown or inconsistent sample/mixture: " + Fs.sample + " / " + Fs.mixture);
		XFCK(xf_upload_aos(ctx, d_U, U.data()));
		XFCK(xf_upload_aos(ctx, d_U1, U.data())); // every sample sets U1 = U (e.g. insert-st/ini_sample.hpp:112-119)
		XFCK(xf_set_scalar(ctx, "T", T.data()));
	}
	void Fluid::BoundaryCondition(int flag)
	{
		double *UI = flag == 0 ? d_U : d_U1;
		XFCK(xf_boundary(ctx, UI, BCs));
		if (halo_exchange && halo_exchange(UI))
			throw std::runtime_error("halo exchange failed");
	}
	bool Fluid::UpdateFluidStates(int flag)
	{
		int err = 0;
		XFCK(xf_update_states(ctx, flag == 0 ? d_U : d_U1, &err));
		return err != 0;
	}
	void Fluid::ComputeFluidLU(int flag) { XFCK(xf_get_lu(ctx, flag == 0 ? d_U : d_U1, d_LU)); }
	void Fluid::UpdateFluidURK3(int flag, double dt) { XFCK(xf_update_u_rk3(ctx, d_U, d_U1, d_LU, dt, flag)); }
	double Fluid::GetFluidDt()
	{
		double dt = 0, m[3];
		XFCK(xf_get_dt(ctx, &dt, m));
		if (allreduce_max3)
		{ // Fluids.cpp:902-913: MAX over ranks, then the same formula
			if (allreduce_max3(m))
				throw std::runtime_error("dt all-reduce failed");
			dt = Fs.bl.CFLnumber / (m[0] * Fs.bl._dx + m[1] * Fs.bl._dy + m[2] * Fs.bl._dz);
		}
		return dt;
	}
	bool Fluid::EstimateFluidNAN(int flag)
	{
		int err = 0;
		XFCK(xf_estimate_nan(ctx, flag == 3 ? d_U : d_U1, d_LU, &err)); // Fluids.cpp:967-985: U1, U1, U for flag 1, 2, 3
		return err != 0;
	}

	XFLUIDS::XFLUIDS(Setup &setup, int device) : Ss(setup), rank(setup.myRank), nranks(setup.nRanks)
	{
		fluids.emplace_back(new Fluid(setup, device));
	}
	XFLUIDS::~XFLUIDS()
	{
		if (slab)
			xf_slab_destroy(slab);
	}
	void XFLUIDS::AttachSlab(xf_comm *comm)
	{
		Fluid &f = *fluids[0];
		if (xf_slab_create(f.ctx, comm, f.BCs, Ss.artificial, Ss.weno, nullptr, &slab) != XF_OK)
			throw std::runtime_error(std::string("xf_slab_create failed: ") + xf_slab_last_error());
		comm_ = comm;
		xf_slab *sl = slab;
		f.halo_exchange = [sl](double *UI)
		{ return xf_slab_halo(sl, UI); };
		f.allreduce_max3 = [sl](double *m3)
		{ return xf_slab_allreduce_max_host(sl, m3, 3); };
	}
	// the guards of one rank decide for all (the reference all-reduces its error counters the same way, XFLUIDS.cpp:567-595)
	static bool any_rank(xf_slab *slab, bool err)
	{
		if (!slab)
			return err;
		double v = err ? 1.0 : 0.0;
		if (xf_slab_allreduce_max_host(slab, &v, 1) != XF_OK)
			throw std::runtime_error(std::string("error all-reduce failed: ") + xf_slab_last_error());
		return v != 0.0;
	}
	void XFLUIDS::AllocateMemory() { fluids[0]->AllocateFluidMemory(); }
	void XFLUIDS::InitialCondition() { fluids[0]->InitialU(); }
	void XFLUIDS::BoundaryCondition(int flag) { fluids[0]->BoundaryCondition(flag); }
	bool XFLUIDS::UpdateStates(int flag)
	{
		const bool err = any_rank(slab, fluids[0]->UpdateFluidStates(flag));
		// global Lax-Friedrichs on several ranks: the running maxima of |lambda| are MAX-reduced before any sweep reads them (the
		// reference's MPI build reduces eigen_block the same way, ConVenction_block.hpp:115-215)
		if (slab && Ss.artificial == 3 && Ss.weno != 7 && xf_comm_allreduce_max(comm_, xf_device_glfmax(fluids[0]->ctx), 9, nullptr) != XF_OK)
			throw std::runtime_error(std::string("GLF all-reduce failed: ") + xf_slab_last_error());
		return err;
	}
	double XFLUIDS::ComputeTimeStep() { return fluids[0]->GetFluidDt(); }
	void XFLUIDS::ComputeLU(int flag) { fluids[0]->ComputeFluidLU(flag); }
	void XFLUIDS::UpdateU(int flag) { fluids[0]->UpdateFluidURK3(flag, dt); }
	bool XFLUIDS::EstimateNAN(int flag) { return any_rank(slab, fluids[0]->EstimateFluidNAN(flag)); }

	bool XFLUIDS::RungeKuttaSP3rd(int flag)
	{
		const int which = flag == 1 ? 0 : 1;
		BoundaryCondition(which);
		if (UpdateStates(which))
			return true;
		ComputeLU(which);
		if (EstimateNAN(flag))
			return true;
		UpdateU(flag);
		return false;
	}
	bool XFLUIDS::SinglePhaseSolverRK3rd()
	{
		if (RungeKuttaSP3rd(1)) return true;
		if (RungeKuttaSP3rd(2)) return true;
		return RungeKuttaSP3rd(3);
	}

	void XFLUIDS::EnableOutput() { out.reset(new FieldOutput(*this)); }

	bool XFLUIDS::Evolution(bool fused)
	{
		Fluid &f = *fluids[0];
		bool error_out = false;
		size_t TimeLoop = 0;
		// output cadence of XFLUIDS::Evolution (XFLUIDS.cpp:105-183): at Iteration % OutInterval == 0 and after every output stamp, at most
		// nOutMax + 1 times; then once more at the end (XFLUIDS.cpp:309)
		const Json &run = Ss.j_conf.at("run");
		const int OutInterval = std::max(1, int(run.value("OutInterval", double(Ss.nStepmax_json)))), nOutput = int(run.value("nOutMax", 0.0));
		int OutNum = 0;
		bool TimeLoopOut = false;
		XFCK(xf_synchronize(f.ctx));
		auto t0 = std::chrono::high_resolution_clock::now();
		while (TimeLoop < Ss.OutTimeStamps.size())
		{
			const double target_t = (physicalTime < Ss.OutTimeStamps[TimeLoop].time) ? Ss.OutTimeStamps[TimeLoop].time : Ss.OutTimeStamps[TimeLoop++].time;
			const OutStamp &OutAtThis = Ss.OutTimeStamps[TimeLoop > 0 ? TimeLoop - 1 : 0];
			if (fused)
			{
				while (physicalTime < target_t && Iteration < Ss.nStepmax)
				{
					if (out && ((Iteration % OutInterval == 0) || TimeLoopOut) && OutNum <= nOutput)
					{
						OutNum++, TimeLoopOut = false;
						out->output(OutAtThis.spec, physicalTime, std::to_string(Iteration));
					}
					int done = 0, err = 0;
					// (with field output on, a batch ends where the next Iteration % OutInterval == 0 output is due)
					const int until_out = out ? OutInterval - Iteration % OutInterval : Ss.nStepmax;
					const int nrun = std::min(Ss.nStepmax - Iteration, until_out);
					XFCK(xf_set_time(f.ctx, physicalTime));
					// one GPU: CUDA-graph replay of whole steps; z-slabs: the slab stepper (halo exchange overlapped with interior work, dt
					// MAX-reduced on the device) -- both keep dt and the physical time on the device
					int rc = slab ? xf_slab_run(slab, f.d_U, f.d_U1, f.d_LU, nrun, target_t, &done, &physicalTime, &err)
								  : xf_run(f.ctx, f.d_U, f.d_U1, f.d_LU, f.BCs, nrun, target_t, &done, &physicalTime, &err);
					if (rc != XF_OK && rc != XF_ERR_NUMERIC)
						throw std::runtime_error(std::string("fused time loop failed: ") + (slab ? xf_slab_last_error() : xf_last_error()));
					Iteration += done;
					XFCK(xf_get_time(f.ctx, nullptr, &dt));
					error_out = err != 0;
					if (verbose && rank == 0)
						std::cout << "N=" << std::setw(7) << Iteration << "  last dt: " << std::setw(14) << std::setprecision(8) << dt
								  << "  End physicalTime: " << std::setw(14) << physicalTime << "\n";
					if (error_out || done == 0)
						break;
				}
			}
			else
				while (physicalTime < target_t)
				{
					if (out && ((Iteration % OutInterval == 0) || TimeLoopOut) && OutNum <= nOutput)
					{
						OutNum++, TimeLoopOut = false;
						out->output(OutAtThis.spec, physicalTime, std::to_string(Iteration));
					}
					Iteration++;
					dt = ComputeTimeStep();
					if (physicalTime + dt > target_t)
						dt = target_t - physicalTime;
					const double tbak = physicalTime;
					physicalTime += dt;
					if (verbose && rank == 0)
						std::cout << "N=" << std::setw(7) << Iteration << "  beginning physicalTime: " << std::setw(14) << std::setprecision(8) << tbak
								  << " dt: " << std::setw(14) << dt << "End physicalTime: " << std::setw(14) << std::setprecision(8) << physicalTime << "\n";
					error_out = error_out || SinglePhaseSolverRK3rd();
					if (error_out || Ss.nStepmax <= Iteration)
						break;
				}
			if (error_out || Ss.nStepmax <= Iteration)
				break;
			TimeLoopOut = true;
		}
		XFCK(xf_synchronize(f.ctx));
		loop_seconds = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
		if (out) // the last step's output (XFLUIDS.cpp:309)
			out->output(Ss.OutTimeStamps.back().spec, physicalTime, std::to_string(Iteration));
		return error_out;
	}

	void XFLUIDS::DownloadU(double *h) const { XFCK(xf_download_aos(fluids[0]->ctx, fluids[0]->d_U, h)); }

	// int Step; real_t Time; float elapsed; real_t U[Xmax*Ymax*Zmax*Emax]   (XFLUIDS.cpp:658-687)
	void XFLUIDS::Output_Ubak(const std::string &path) const
	{
		std::vector<double> U(Ss.ncells() * Ss.Emax);
		DownloadU(U.data());
		std::ofstream fout(path, std::ios::binary);
		if (!fout)
			throw std::runtime_error("cannot write " + path);
		const int Step = Iteration;
		const double Time = physicalTime;
		const float el = float(loop_seconds);
		fout.write((const char *)&Step, sizeof(Step));
		fout.write((const char *)&Time, sizeof(Time));
		fout.write((const char *)&el, sizeof(el));
		fout.write((const char *)U.data(), U.size() * sizeof(double));
	}

	// XFLUIDS::Read_Ubak (XFLUIDS.cpp:689-724), called by the reference from InitialCondition (XFLUIDS.cpp:616-623): Step, Time and
	// U replace the freshly initialised state; everything else -- notably the Newton warm-start T -- stays as the initial
	// condition left it, exactly as in the reference.
	bool XFLUIDS::Read_Ubak(const std::string &path)
	{
		std::ifstream fin(path, std::ios::binary);
		if (!fin.is_open())
		{
			if (rank == 0 && verbose)
				std::cout << "CheckingPoint-file not exist or open failed, CheckingPoint closed." << std::endl;
			return false;
		}
		int Step = 0;
		double Time = 0;
		float el = 0;
		std::vector<double> U(Ss.ncells() * Ss.Emax);
		fin.read((char *)&Step, sizeof(Step));
		fin.read((char *)&Time, sizeof(Time));
		fin.read((char *)&el, sizeof(el));
		fin.read((char *)U.data(), U.size() * sizeof(double));
		if (!fin || size_t(fin.gcount()) != U.size() * sizeof(double))
			throw std::runtime_error("CheckingPoint file " + path + " does not match this block (Xmax*Ymax*Zmax*Emax doubles expected)");
		Iteration = Step, physicalTime = Time;
		XFCK(xf_upload_aos(fluids[0]->ctx, fluids[0]->d_U, U.data()));
		return true;
	}
} // namespace xfh
