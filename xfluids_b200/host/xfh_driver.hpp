// xfh_driver.hpp -- host mirror of the reference's driver classes: Fluid (src/Fluids.cpp, global_class.h:14-55) owns the
// device arrays of one fluid and forwards to the block-level entry points -- here the C ABI of libxfluids_b200.so --
// and XFLUIDS (src/XFLUIDS.cpp, global_class.h:57-121) sequences main -> Allocate -> IC -> BC -> UpdateStates -> Evolution.
#pragma once
#include <functional>
#include <memory>
#include <string>
#include <vector>
#include "xfh_setup.hpp"

namespace xfh
{
	class Fluid
	{
		Setup &Fs;

	public:
		xf_ctx *ctx = nullptr;
		double *d_U = nullptr, *d_U1 = nullptr, *d_LU = nullptr; // reference Fluid::d_U, d_U1, d_LU (SoA on the device here)
		int BCs[6];
		int error_patched_times = 0;
		// hook called between the local ghost fill and the primitive recovery of every stage: the z-slab halo exchange of a
		// multi-GPU run (reference: MpiTrans::MpiTransBuf inside FluidBoundaryCondition).  Empty on one GPU.
		std::function<int(double *d_UI)> halo_exchange;
		// hook for the cross-rank MAX of the three dt maxima / error word (reference: MPI_Allreduce in Fluid::GetFluidDt)
		std::function<int(double *m3)> allreduce_max3;

		Fluid(Setup &setup, int device);
		~Fluid();
		void AllocateFluidMemory();                 // Fluids.cpp:270-583
		void InitialU();                            // Fluids.cpp:585-588 -> InitializeFluidStates
		void BoundaryCondition(int flag);           // Fluids.cpp:922-933
		bool UpdateFluidStates(int flag);           // Fluids.cpp:935-944 ; true = error captured
		void ComputeFluidLU(int flag);              // Fluids.cpp:951-961
		void UpdateFluidURK3(int flag, double dt);  // Fluids.cpp:946-949
		double GetFluidDt();                        // Fluids.cpp:897-920
		bool EstimateFluidNAN(int flag);            // Fluids.cpp:963-1040
	};

	class XFLUIDS;
	struct OutFmt;
	struct OutVar;
	// field output in the reference's file formats (xfh_output.cpp; XFLUIDS::Output & co, src/XFLUIDS.cpp:726-1793)
	class FieldOutput
	{
	public:
		struct Host;
		explicit FieldOutput(XFLUIDS &x);
		~FieldOutput();
		// XFLUIDS::Output: `spec` = the "{-C=..;-P=..;-V=..}" part of the output stamp in charge, `step` = the iteration label
		void output(const std::string &spec, double time, const std::string &step);
		int OutDAT = 1, OutVTI = 0;
		std::string dir, prefix;

	private:
		XFLUIDS &X;
		Setup &S;
		Host *h;
		int nb[3], mn[3], mxi[3]; // OutSize VTI (XFLUIDS.cpp:35-62)
		void copy_from_device();
		std::vector<OutVar> variables() const;
		void write_vti(const OutFmt &f, const std::vector<OutVar> &vars, const std::string &step, double time, bool compressed);
		void write_cplt(const std::string &step, double time);
	};

	class XFLUIDS
	{
	public:
		Setup &Ss;
		double dt = 0, physicalTime = 0;
		int Iteration = 0, rank = 0, nranks = 1;
		std::vector<std::unique_ptr<Fluid>> fluids;
		bool verbose = true;
		double loop_seconds = 0;
		std::unique_ptr<FieldOutput> out; // set by EnableOutput(): field files at the output stamps and at the end, like XFLUIDS::Evolution
		void EnableOutput();

		XFLUIDS(Setup &setup, int device);
		~XFLUIDS();
		// multi-GPU (z-slabs, one XFLUIDS per rank -- a thread of `xfluids -mpi=1,1,N` or a process): binds this rank's fluid to the
		// C++ slab stepper of the CUDA library (xf_slab_*, NCCL over NVLink).  Sets Fluid::halo_exchange / allreduce_max3, the two
		// places where the reference calls MPI (BCs_block.cpp:50-217, Fluids.cpp:902-913); Evolution(fused) then runs xf_slab_run.
		void AttachSlab(xf_comm *comm);
		xf_slab *slab = nullptr;
		xf_comm *comm_ = nullptr;
		void AllocateMemory();
		void InitialCondition();
		void BoundaryCondition(int flag = 0);
		bool UpdateStates(int flag = 0);
		double ComputeTimeStep();                                  // XFLUIDS.cpp:527-544
		bool SinglePhaseSolverRK3rd();                             // XFLUIDS.cpp:377-411
		bool RungeKuttaSP3rd(int flag);                            // XFLUIDS.cpp:441-525
		void ComputeLU(int flag);
		void UpdateU(int flag);
		bool EstimateNAN(int flag);
		// XFLUIDS::Evolution (XFLUIDS.cpp:105-311).  fused = false: the reference's call-by-call loop (host reads dt and the
		// error flags every stage); fused = true: xf_run -- device-resident dt, one CUDA-graph replay per step.
		bool Evolution(bool fused);
		void Output_Ubak(const std::string &path) const;           // XFLUIDS.cpp:658-687 checkpoint format
		bool Read_Ubak(const std::string &path);                   // XFLUIDS.cpp:689-724 restart from a checkpoint (reference or ours)
		void DownloadU(double *h_aos) const;
	};
} // namespace xfh
