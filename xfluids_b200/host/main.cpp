// xfluids_b200 executable: the reference's src/main.cpp flow (main.cpp:30-54) on the CUDA engine.
//   xfluids <settings.json> [-run=nx,ny,nz[,nsteps]] [-sample=..] [-mixture=..] [-weno=5|6|7] [-alpha=LLF|GLF|ROE]
//           [-fp=0|1] [-pp=0|1] [-cfl=x] [-bc=a,b,c,d,e,f] [-dev=n] [-blocks] [-ckpt=path] [-restart=path] [-quiet]
//           [-mpi=1,1,N -mpi-s=weak|strong] [-out [-outdir=dir] [-dv=name]] [-visc=0|1]
// -out writes the reference's field files (VTI / CVTI .vti + .pvti, CPLT .dat) as run.OutDAT / OutVTI and the output stamps ask.
// -ckpt writes the reference's CheckingPoint format at the end, -restart continues from such a file (XFLUIDS.cpp:616-623,689-724).
// -mpi=1,1,N: N z-slabs on the GPUs dev .. dev+N-1 of this box.  Where the reference starts N MPI processes (mpiPacks.cpp:3-75), this
// executable runs one host thread per GPU in ONE process; the ranks talk through the NCCL slab stepper of the CUDA library (halo
// exchange over NVLink overlapped with interior work, MAX all-reduce of dt); with -ckpt every rank writes <path>.rank<r>.
// runtime.dat/ is searched upwards from the executable like the reference does (external/fworkdir.hpp:11-27).
#include <cstring>
#include <filesystem>
#include <iostream>
#include <mutex>
#include <thread>
#include "xfh_driver.hpp"

static std::string find_workdir(const std::string &exe)
{
	namespace fs = std::filesystem;
	fs::path p = fs::absolute(exe).parent_path();
	for (;;)
	{
		if (fs::exists(p / "runtime.dat"))
			return p.string();
		if (!p.has_parent_path() || p == p.parent_path())
			break;
		p = p.parent_path();
	}
	throw std::runtime_error("Error: cannot find WorkDir, run executable file under Program Directory.");
}

struct Options
{
	int device = 0;
	bool fused = true, quiet = false;
	std::string ckpt, restart;
	int nranks = 1;
	bool output = false; // -out: write the field files of run.OutDAT / OutVTI at the output stamps (off by default: benchmarks)
};

struct RankResult
{
	int iteration = 0, restored = 0;
	double seconds = 0, time = 0, cells = 0;
	bool error = false;
	std::string failure;
};

// main.cpp:30-54 for one rank
static void run_rank(const char *json, const std::vector<std::string> &cli, const std::string &workdir, const Options &o, int rank, const char *nccl_id, RankResult *res)
{
	try
	{
		xfh::Setup setup(json, cli, workdir, rank, o.nranks);
		if (!o.quiet && rank == 0)
			setup.print();
		xfh::XFLUIDS solver(setup, o.device + rank);
		solver.verbose = !o.quiet;
		xf_comm *comm = nullptr;
		if (o.nranks > 1)
		{
			if (xf_comm_create(nccl_id, rank, o.nranks, o.device + rank, &comm) != XF_OK)
				throw std::runtime_error(std::string("xf_comm_create failed: ") + xf_slab_last_error());
			solver.AttachSlab(comm);
		}
		solver.AllocateMemory();
		if (o.output)
			solver.EnableOutput();
		solver.InitialCondition();
		const std::string suffix = o.nranks > 1 ? ".rank" + std::to_string(rank) : "";
		if (!o.restart.empty() && !solver.Read_Ubak(o.restart + suffix))
			throw std::runtime_error("cannot read checkpoint " + o.restart + suffix);
		res->restored = solver.Iteration;
		solver.BoundaryCondition();
		if (solver.UpdateStates())
			throw std::runtime_error("errors of primitive variables captured in the initial state");
		res->error = solver.Evolution(o.fused);
		res->iteration = solver.Iteration, res->seconds = solver.loop_seconds, res->time = solver.physicalTime;
		res->cells = double(setup.bl.X_inner) * setup.bl.Y_inner * setup.bl.Z_inner;
		if (!o.ckpt.empty())
			solver.Output_Ubak(o.ckpt + suffix);
		if (solver.slab)
			xf_slab_destroy(solver.slab), solver.slab = nullptr;
		if (comm)
			xf_comm_destroy(comm);
	}
	catch (const std::exception &e)
	{
		res->failure = e.what();
	}
}

int main(int argc, char *argv[])
{
	try
	{
		if (argc < 2)
		{
			std::cerr << "usage: " << argv[0] << " settings.json [-run=nx,ny,nz[,nsteps]] [-sample=..] [-mixture=..] [-weno=5|6|7] [-alpha=LLF] [-pp=0|1] [-dev=n] [-blocks] [-ckpt=file] [-restart=file] [-mpi=1,1,N -mpi-s=weak|strong]\n";
			return 2;
		}
		std::vector<std::string> cli(argv + 2, argv + argc);
		Options o;
		for (auto &a : cli)
		{
			if (!a.compare(0, 5, "-dev=")) o.device = std::atoi(a.c_str() + 5);
			if (a == "-blocks") o.fused = false;
			if (a == "-quiet") o.quiet = true;
			if (a == "-out") o.output = true;
			if (!a.compare(0, 6, "-ckpt=")) o.ckpt = a.substr(6);
			if (!a.compare(0, 9, "-restart=")) o.restart = a.substr(9);
			if (!a.compare(0, 5, "-mpi="))
			{
				int mx = 1, my = 1, mz = 1;
				if (std::sscanf(a.c_str() + 5, "%d,%d,%d", &mx, &my, &mz) == 3)
					o.nranks = mz;
			}
		}
		const std::string workdir = find_workdir(argv[0]);
		std::vector<RankResult> res(o.nranks);
		char nccl_id[128] = {0};
		if (o.nranks > 1 && xf_comm_unique_id(nccl_id) != XF_OK)
			throw std::runtime_error(std::string("NCCL is not available: ") + xf_slab_last_error());
		if (o.nranks == 1)
			run_rank(argv[1], cli, workdir, o, 0, nccl_id, &res[0]);
		else
		{
			std::vector<std::thread> th;
			for (int r = 0; r < o.nranks; r++)
				th.emplace_back(run_rank, argv[1], std::cref(cli), std::cref(workdir), std::cref(o), r, nccl_id, &res[r]);
			for (auto &t : th)
				t.join();
		}
		bool err = false;
		double cells = 0, seconds = 0;
		for (int r = 0; r < o.nranks; r++)
		{
			if (!res[r].failure.empty())
				throw std::runtime_error("rank " + std::to_string(r) + ": " + res[r].failure);
			err = err || res[r].error;
			cells += res[r].cells;
			seconds = std::max(seconds, res[r].seconds);
		}
		// throughput of THIS run: the steps a restart restored were not computed here
		const int steps_here = res[0].iteration - res[0].restored;
		std::cout.precision(10);
		std::cout << "XFLUIDS_B200 steps=" << res[0].iteration << " seconds=" << seconds << " mcell_stage_per_s=" << cells * 3.0 * steps_here / seconds / 1e6
				  << " time=" << res[0].time << " ranks=" << o.nranks << " error=" << int(err) << std::endl;
		return err ? 1 : 0;
	}
	catch (const std::exception &e)
	{
		std::cerr << "xfluids_b200: " << e.what() << std::endl;
		return 3;
	}
}
