// xfluids_b200 executable: the reference's src/main.cpp flow (main.cpp:30-54) on the CUDA engine.
//   xfluids <settings.json> [-run=nx,ny,nz[,nsteps]] [-sample=..] [-mixture=..] [-weno=5|7] [-alpha=LLF|GLF|ROE]
//           [-fp=0|1] [-pp=0|1] [-cfl=x] [-dev=n] [-blocks] [-ckpt=path] [-restart=path] [-quiet]
// -ckpt writes the reference's CheckingPoint format at the end, -restart continues from such a file (XFLUIDS.cpp:616-623,689-724).
// runtime.dat/ is searched upwards from the executable like the reference does (external/fworkdir.hpp:11-27).
#include <cstring>
#include <filesystem>
#include <iostream>
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

int main(int argc, char *argv[])
{
	try
	{
		if (argc < 2)
		{
			std::cerr << "usage: " << argv[0] << " settings.json [-run=nx,ny,nz[,nsteps]] [-sample=..] [-mixture=..] [-weno=5|6|7] [-alpha=LLF] [-pp=0|1] [-dev=n] [-blocks] [-ckpt=file] [-restart=file]\n";
			return 2;
		}
		std::vector<std::string> cli(argv + 2, argv + argc);
		int device = 0;
		bool fused = true, quiet = false;
		std::string ckpt, restart;
		for (auto &a : cli)
		{
			if (!a.compare(0, 5, "-dev=")) device = std::atoi(a.c_str() + 5);
			if (a == "-blocks") fused = false;
			if (a == "-quiet") quiet = true;
			if (!a.compare(0, 6, "-ckpt=")) ckpt = a.substr(6);
			if (!a.compare(0, 9, "-restart=")) restart = a.substr(9);
		}
		xfh::Setup setup(argv[1], cli, find_workdir(argv[0]));
		if (!quiet)
			setup.print();
		xfh::XFLUIDS solver(setup, device);
		solver.verbose = !quiet;
		solver.AllocateMemory();
		solver.InitialCondition();
		if (!restart.empty() && !solver.Read_Ubak(restart))
			throw std::runtime_error("cannot read checkpoint " + restart);
		solver.BoundaryCondition();
		if (solver.UpdateStates())
			throw std::runtime_error("errors of primitive variables captured in the initial state");
		const bool err = solver.Evolution(fused);
		const double cells = double(setup.bl.X_inner) * setup.bl.Y_inner * setup.bl.Z_inner;
		std::cout.precision(10);
		std::cout << "XFLUIDS_B200 steps=" << solver.Iteration << " seconds=" << solver.loop_seconds
				  << " mcell_stage_per_s=" << cells * 3.0 * solver.Iteration / solver.loop_seconds / 1e6 << " time=" << solver.physicalTime
				  << " error=" << int(err) << std::endl;
		if (!ckpt.empty())
			solver.Output_Ubak(ckpt);
		return err ? 1 : 0;
	}
	catch (const std::exception &e)
	{
		std::cerr << "xfluids_b200: " << e.what() << std::endl;
		return 3;
	}
}
