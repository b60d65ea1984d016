// xfh_capi.cpp -- C entry points of libxfluids_host.so: lets the Python tests / bench drive the C++ host layer (Setup,
// initial conditions, the XFLUIDS driver) that sits above the CUDA C ABI.
#include <cstring>
#include "xfh_driver.hpp"

using namespace xfh;
static thread_local std::string g_herr;

extern "C"
{
	const char *xfh_last_error(void) { return g_herr.c_str(); }

	void *xfh_setup_create(const char *json, const char *workdir, int argc, const char **argv, int rank, int nranks)
	{
		try
		{
			std::vector<std::string> cli;
			for (int i = 0; i < argc; i++)
				cli.push_back(argv[i]);
			return new Setup(json, cli, workdir, rank, nranks);
		}
		catch (const std::exception &e)
		{
			g_herr = e.what();
			return nullptr;
		}
	}
	void xfh_setup_destroy(void *s) { delete (Setup *)s; }
	void xfh_setup_block(void *s, xf_block *b) { *b = ((Setup *)s)->bl; }
	void xfh_setup_thermal(void *s, xf_thermal *t) { *t = ((Setup *)s)->thermal(); }
	void xfh_setup_scheme(void *s, xf_scheme *sc) { *sc = ((Setup *)s)->scheme(); }
	void xfh_setup_transport(void *s, xf_transport *t) { *t = ((Setup *)s)->transport(); }
	void xfh_setup_bc(void *s, int bc[6]) { ((Setup *)s)->rank_boundarys(bc); }
	// info[0..7] = Emax, num_species, cop, ghost_species, nStepmax, mz, myMpiPos_z, n_stamps
	void xfh_setup_info(void *s_, int info[8])
	{
		Setup *s = (Setup *)s_;
		info[0] = s->Emax, info[1] = s->num_species, info[2] = s->cop, info[3] = s->ghost_species, info[4] = s->nStepmax;
		info[5] = s->mz, info[6] = s->myMpiPos_z, info[7] = (int)s->OutTimeStamps.size();
	}
	void xfh_setup_stamps(void *s_, double *t)
	{
		Setup *s = (Setup *)s_;
		for (size_t i = 0; i < s->OutTimeStamps.size(); i++)
			t[i] = s->OutTimeStamps[i].time;
	}
	// ini[0..9]: post-shock rho,p,T,u ; pre-shock rho,p,T ; bubble C, _xa2, tau_H  (Mach_Shock results, for tests)
	void xfh_setup_ini(void *s_, double ini[10])
	{
		const IniShape &i = ((Setup *)s_)->ini;
		ini[0] = i.blast_density_in, ini[1] = i.blast_pressure_in, ini[2] = i.blast_T_in, ini[3] = i.blast_u_in;
		ini[4] = i.blast_density_out, ini[5] = i.blast_pressure_out, ini[6] = i.blast_T_out, ini[7] = i.C, ini[8] = i._xa2, ini[9] = i.tau_H;
	}
	int xfh_initial_condition(void *s, double *U, double *T) { return InitialCondition(*(Setup *)s, U, T); }

	void *xfh_solver_create(void *s, int device)
	{
		try
		{
			XFLUIDS *x = new XFLUIDS(*(Setup *)s, device);
			x->verbose = false;
			x->AllocateMemory();
			return x;
		}
		catch (const std::exception &e)
		{
			g_herr = e.what();
			return nullptr;
		}
	}
	void xfh_solver_destroy(void *x) { delete (XFLUIDS *)x; }
	// main.cpp:40-48: IC -> BC -> UpdateStates
	int xfh_solver_init(void *x_)
	{
		try
		{
			XFLUIDS *x = (XFLUIDS *)x_;
			x->InitialCondition();
			x->BoundaryCondition();
			return x->UpdateStates() ? 1 : 0;
		}
		catch (const std::exception &e)
		{
			g_herr = e.what();
			return -1;
		}
	}
	int xfh_solver_evolve(void *x_, int fused, int *iteration, double *time, double *seconds)
	{
		try
		{
			XFLUIDS *x = (XFLUIDS *)x_;
			const bool err = x->Evolution(fused != 0);
			*iteration = x->Iteration, *time = x->physicalTime, *seconds = x->loop_seconds;
			return err ? 1 : 0;
		}
		catch (const std::exception &e)
		{
			g_herr = e.what();
			return -1;
		}
	}
	int xfh_solver_download(void *x_, double *U)
	{
		try
		{
			((XFLUIDS *)x_)->DownloadU(U);
			return 0;
		}
		catch (const std::exception &e)
		{
			g_herr = e.what();
			return -1;
		}
	}
	int xfh_solver_checkpoint(void *x_, const char *path)
	{
		try
		{
			((XFLUIDS *)x_)->Output_Ubak(path);
			return 0;
		}
		catch (const std::exception &e)
		{
			g_herr = e.what();
			return -1;
		}
	}
	void *xfh_solver_ctx(void *x_) { return ((XFLUIDS *)x_)->fluids[0]->ctx; }
	void xfh_solver_fields(void *x_, double **U, double **U1, double **LU)
	{
		Fluid &f = *((XFLUIDS *)x_)->fluids[0];
		*U = f.d_U, *U1 = f.d_U1, *LU = f.d_LU;
	}
}
