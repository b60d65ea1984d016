// xfh_setup.cpp -- see xfh_setup.hpp.  Formulas and their evaluation order follow the reference files cited per function.
#include "xfh_setup.hpp"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <iostream>

namespace xfh
{
	static std::vector<std::string> split(const std::string &s, char c)
	{
		std::vector<std::string> r;
		std::string cur;
		for (char ch : s)
		{
			if (ch == c)
				r.push_back(cur), cur.clear();
			else
				cur += ch;
		}
		r.push_back(cur);
		return r;
	}

	// external/options.hpp:19-47  AppendParas::match
	std::vector<std::string> Setup::match(const std::string &opt) const
	{
		const std::string key = opt + "=";
		for (const std::string &a : args)
			if (a.find(key) != std::string::npos)
				return split(std::string(a).erase(0, key.length()), ',');
		return {};
	}

	Setup::Setup(const std::string &json_path, const std::vector<std::string> &cli, const std::string &workdir, int rank, int nranks)
		: WorkDir(workdir), myRank(rank), nRanks(nranks), args(cli)
	{
		j_conf = ReadJson(json_path);
		ReWrite();
		ReadSpecies();
		if (Visc)
			GetFitCoefficient(); // constructor.cpp:36-38
		init();
		// MpiTrans::MpiTrans (mpiPacks.cpp:3-75): position of this rank in the process grid; only z-slabs here
		if (mx != 1 || my != 1)
			throw std::runtime_error("xfluids_b200 decomposes along z only: mx and my must be 1");
		if (mz != nRanks && mz != 1)
			throw std::runtime_error("mz must equal the number of ranks");
		myMpiPos_z = (mz > 1) ? myRank : 0;
	}

	// ---- iniset.cpp:8-67 + read_json.cpp:30-148 ---------------------------------------------------------
	void Setup::ReadIni()
	{
		const Json &run = j_conf.at("run"), &mpi = j_conf.at("mpi"), &eq = j_conf.at("equations"), &mesh = j_conf.at("mesh"), &init_ = j_conf.at("init");
		const Json &b2 = j_conf.at("b200");
		sample = b2.value("sample", sample.c_str());
		mixture = b2.value("mixture", mixture.c_str());
		weno = int(b2.value("weno", double(weno)));
		const std::string art = b2.value("artificial", "LLF");
		artificial = art == "ROE" ? 1 : (art == "GLF" ? 3 : 2);
		fp_mode = int(b2.value("fp_mode", 0.0));

		OutputDir = run.value("OutputDir", "output");
		nStepmax = int(run.value("nStepMax", 10.0));
		nStepmax_json = nStepmax;
		RcalInterval = int(run.value("RcalInterval", 100.0));
		RSources = eq.value("Sources_React", false);
		PositivityPreserving = eq.value("PositivityPreserving", false);
		Tnode = eq.value("ViscosityFittingTnode", Tnode);                    // read_json.cpp:70
		Yil_limiter_json = mesh.value("Yil_limiter", 1.0E10), Dim_limiter_json = mesh.value("Dim_limiter", 2.0E-3); // read_json.cpp:104-105
		Visc = b2.value("visc", 0.0) != 0.0, Visc_Heat = b2.value("visc_heat", Visc ? 1.0 : 0.0) != 0.0, Visc_Diffu = b2.value("visc_diffu", Visc ? 1.0 : 0.0) != 0.0;
		mx = int(mpi.value("mx", 1.0)), my = int(mpi.value("my", 1.0)), mz = int(mpi.value("mz", 1.0));

		bl.CFLnumber = run.value("CFLnumber", 0.4);
		const std::vector<double> Inner = mesh.value("Resolution", std::vector<double>{1, 0, 0});
		const std::vector<double> Bwidth = mesh.value("Ghost_width", std::vector<double>{4, 4, 4});
		const std::vector<double> medg = mesh.value("DOMAIN_Medg", std::vector<double>{0.0, 0.0, 0.0});
		const std::vector<double> size = mesh.value("DOMAIN_Size", std::vector<double>{1.0, 1.0, 1.0});
		bl.DimX = Inner[0] != 0, bl.DimY = Inner[1] != 0, bl.DimZ = Inner[2] != 0;
		bl.X_inner = int(Inner[0]), bl.Y_inner = int(Inner[1]), bl.Z_inner = int(Inner[2]);
		bl.Bwidth_X = int(Bwidth[0]), bl.Bwidth_Y = int(Bwidth[1]), bl.Bwidth_Z = int(Bwidth[2]);
		Domain_xmin = medg[0], Domain_ymin = medg[1], Domain_zmin = medg[2];
		Domain_length = size[0], Domain_width = size[1], Domain_height = size[2];
		const std::vector<double> bcs = mesh.value("Boundarys", std::vector<double>{2, 2, 2, 2, 2, 2});
		for (int i = 0; i < 6; i++)
			Boundarys[i] = int(bcs[i]);

		// init section (read_json.cpp:131-148)
		ini.Ma = init_.value("blast_mach", 0.0);
		ini.cop_type = int(init_.value("cop_type", 0.0));
		ini.blast_type = int(init_.value("blast_type", 1.0));
		const std::vector<double> cop_pos = init_.value("cop_center", std::vector<double>{0.0, 0.0, 0.0});
		const std::vector<double> blast_pos = init_.value("blast_center", std::vector<double>{0.5, 0.5, 0.5});
		const std::vector<double> up = init_.value("blast_upstream", std::vector<double>{0.0, 0.0, 298.15, 0.0, 0.0, 0.0});
		const std::vector<double> down = init_.value("blast_downstream", std::vector<double>{0.0, 0.0, 298.15, 0.0, 0.0, 0.0});
		const std::vector<double> cop_in = init_.value("cop_inside", std::vector<double>{down[0], down[1], down[2], 0.0, 0.0, 0.0});
		// ComputeDminJson (read_json.cpp:180-192): uses the JSON's own Dimensions
		double Dmin = size[0] + size[1] + size[2];
		if (Inner[0] != 0) Dmin = std::min(size[0], Dmin);
		if (Inner[1] != 0) Dmin = std::min(size[1], Dmin);
		if (Inner[2] != 0) Dmin = std::min(size[2], Dmin);
		const double xa_json = init_.value("bubble_shape_x", 0.4 * Dmin);
		const double yb_first = xa_json / init_.value("bubble_shape_ratioy", 1.0);
		const double zc_first = xa_json / init_.value("bubble_shape_ratioz", 1.0);
		const double yb_json = init_.value("bubble_shape_y", yb_first);
		const double zc_json = init_.value("bubble_shape_z", zc_first);
		const double bubble_boundary = init_.value("bubble_boundary_cells", 2.0);
		C_json_ = xa_json * init_.value("bubble_boundary_width", bubble_boundary);
		ini.blast_center_x = blast_pos[0], ini.blast_center_y = blast_pos[1], ini.blast_center_z = blast_pos[2];
		ini.xa = xa_json, ini.yb = yb_json, ini.zc = zc_json;
		ini.blast_density_in = up[0], ini.blast_pressure_in = up[1], ini.blast_T_in = up[2];
		ini.blast_u_in = up[3], ini.blast_v_in = up[4], ini.blast_w_in = up[5];
		ini.blast_density_out = down[0], ini.blast_pressure_out = down[1], ini.blast_T_out = down[2];
		ini.blast_u_out = down[3], ini.blast_v_out = down[4], ini.blast_w_out = down[5];
		ini.cop_center_x = cop_pos[0], ini.cop_center_y = cop_pos[1], ini.cop_center_z = cop_pos[2];
		ini.cop_density_in = cop_in[0], ini.cop_pressure_in = cop_in[1], ini.cop_T_in = cop_in[2];
	}

	// ---- iniset.cpp:72-286 ------------------------------------------------------------------------------------
	void Setup::ReWrite()
	{
		ReadIni();
		auto ints = [&](const std::string &o)
		{ std::vector<int> r; for (auto &t : match(o)) r.push_back(std::atoi(t.c_str())); return r; };
		auto dbls = [&](const std::string &o)
		{ std::vector<double> r; for (auto &t : match(o)) r.push_back(std::atof(t.c_str())); return r; };

		std::vector<int> Inner_size = ints("-run");
		if (Inner_size.size() >= 3)
		{
			bl.X_inner = Inner_size[0], bl.Y_inner = Inner_size[1], bl.Z_inner = Inner_size[2];
			bl.DimX = bool(Inner_size[0]), bl.DimY = bool(Inner_size[1]), bl.DimZ = bool(Inner_size[2]);
			if (4 == Inner_size.size())
				nStepmax = Inner_size[3];
		}
		std::vector<int> gcs = ints("-gcs");
		if (gcs.size() >= 3)
			bl.Bwidth_X = gcs[0], bl.Bwidth_Y = gcs[1], bl.Bwidth_Z = gcs[2];
		std::vector<double> dom = dbls("-domain");
		if (dom.size() >= 3)
		{
			if (dom[0] > 0) Domain_length = dom[0];
			if (dom[1] > 0) Domain_width = dom[1];
			if (dom[2] > 0) Domain_height = dom[2];
		}
		std::vector<int> mpiapa = ints("-mpi");
		if (mpiapa.size() >= 3)
			mx = mpiapa[0], my = mpiapa[1], mz = mpiapa[2];
		std::vector<std::string> mpis = match("-mpi-s");
		if (!mpis.empty())
		{
			if (mpis[0] == "strong")
			{
				if (!(bl.X_inner % mx + bl.Y_inner % my + bl.Z_inner % mz))
					bl.X_inner /= mx, bl.Y_inner /= my, bl.Z_inner /= mz;
				else
					throw std::runtime_error("Error: the number of blocks in each direction is not divisible by the number of MPI processes in that direction!");
			}
			else if (mpis[0] == "weak")
				Domain_length *= mx, Domain_width *= my, Domain_height *= mz;
		}
		// run-time versions of the reference's compile-time selections
		if (!match("-sample").empty()) sample = match("-sample")[0];
		if (!match("-mixture").empty()) mixture = match("-mixture")[0];
		if (!match("-weno").empty()) weno = std::atoi(match("-weno")[0].c_str());
		if (!match("-fp").empty()) fp_mode = std::atoi(match("-fp")[0].c_str());
		if (!match("-pp").empty()) PositivityPreserving = std::atoi(match("-pp")[0].c_str()) != 0;
		if (!match("-cfl").empty()) bl.CFLnumber = std::atof(match("-cfl")[0].c_str());
		if (!match("-dv").empty()) select_dv = match("-dv")[0];
		if (!match("-outdir").empty()) OutputDir = match("-outdir")[0];
		if (!match("-visc").empty()) Visc = Visc_Heat = Visc_Diffu = std::atoi(match("-visc")[0].c_str()) != 0;
		if (!match("-visc-heat").empty()) Visc_Heat = std::atoi(match("-visc-heat")[0].c_str()) != 0;
		if (!match("-visc-diffu").empty()) Visc_Diffu = std::atoi(match("-visc-diffu")[0].c_str()) != 0;
		if (!match("-diffu-mpi").empty()) diffu_dim_max0 = std::atoi(match("-diffu-mpi")[0].c_str()) != 0 ? 1.0 : 0.0;
		std::vector<int> bcv = ints("-bc"); // run-time override of mesh.Boundarys (xmin, xmax, ymin, ymax, zmin, zmax; BConditions codes)
		if (bcv.size() == 6)
			for (int i = 0; i < 6; i++)
				Boundarys[i] = bcv[i];
		if (!match("-alpha").empty())
		{
			const std::string a = match("-alpha")[0];
			artificial = a == "ROE" ? 1 : (a == "GLF" ? 3 : 2);
		}

		// output time stamps (iniset.cpp:201-285): they clip dt (XFLUIDS.cpp:197-198), so they are part of the numerics
		const Json &run = j_conf.at("run");
		const std::vector<std::string> arrays = run.strings("OutTimeArrays"), stamps = run.strings("OutTimeStamps");
		for (const std::string &a : arrays)
		{
			std::vector<std::string> temp = split(a, ':');
			std::vector<std::string> tempt = split(temp[0], ';');
			std::vector<std::string> temp0 = split(tempt[1], '*');
			const double tn_b = std::stod(tempt[0]);
			const double cnt = std::stod(temp0[0]), itv = std::stod(temp0[1]);
			for (size_t tn = 1; tn <= size_t(cnt); tn++)
				OutTimeStamps.push_back({tn * itv + tn_b, temp.size() > 1 ? temp[1] : ""});
		}
		for (const std::string &st : stamps)
		{
			std::vector<std::string> temp = split(st, ':');
			const double the_time = std::stod(temp[0]);
			const OutStamp s{the_time, temp.size() > 1 ? temp[1] : ""};
			if (!arrays.empty())
			{
				for (size_t tn = 0; tn + 1 < OutTimeStamps.size(); tn++)
					if (OutTimeStamps[tn + 1].time > the_time)
					{
						OutTimeStamps.insert(OutTimeStamps.begin() + (tn + 1), s);
						break;
					}
			}
			else
				OutTimeStamps.push_back(s);
		}
		if (OutTimeStamps.empty())
			OutTimeStamps.push_back({1.0e300, ""});
	}

	// ---- thermal.cpp:6-32 ---------------------------------------------------------------------------------------
	void Setup::ReadSpecies()
	{
		cop = mixture != "NO-COP";
		const std::string path = WorkDir + "/runtime.dat/" + mixture + "/species_list.dat";
		std::ifstream fins(path);
		if (!fins)
			throw std::runtime_error("cannot open " + path);
		std::vector<std::string> toks;
		std::string t;
		while (fins >> t)
			toks.push_back(t);
		if (cop)
		{
			if (toks.size() % 3)
				throw std::runtime_error("species_list.dat: expected names / outside ratios / inside ratios");
			num_species = int(toks.size() / 3);
			species_name.assign(toks.begin(), toks.begin() + num_species);
			species_ratio_out.resize(num_species), species_ratio_in.resize(num_species);
			for (int n = 0; n < num_species; n++)
				species_ratio_out[n] = std::stod(toks[num_species + n]), species_ratio_in[n] = std::stod(toks[2 * num_species + n]);
			// GhostSpecies (runtime.dat/<mixture>/case_setup.h) is a compile-time macro in the reference; its data files mark
			// it by repeating the filler species' name as the last entry (N2 N2) or by a zero mole ratio on both sides
			ghost_species = (num_species >= 2 && species_name[num_species - 1] == species_name[num_species - 2]) ||
							(species_ratio_out[num_species - 1] == 0.0 && species_ratio_in[num_species - 1] == 0.0);
			Emax = num_species + 4;
		}
		else
		{
			num_species = 1, Emax = 5, ghost_species = false;
			species_name = {toks.empty() ? std::string("NCOP") : toks[0]};
			species_ratio_out = {1.0}, species_ratio_in = {1.0};
		}
		ReadThermal();
	}

	// ---- thermal.cpp:36-175 --------------------------------------------------------------------------------------
	void Setup::ReadThermal()
	{
		const int NS = num_species;
		Hia.assign(NS * 7 * 3, 0.0), Hib.assign(NS * 2 * 3, 0.0), Wi.assign(NS, 0.0), _Wi.assign(NS, 0.0), Ri.assign(NS, 0.0);
		const std::string path = WorkDir + "/runtime.dat/thermal_dynamics.dat";
		std::ifstream fincn(path);
		if (!fincn)
			throw std::runtime_error("cannot open " + path);
		std::vector<std::string> toks;
		std::string t;
		while (fincn >> t)
			toks.push_back(t);
		for (int n = 0; n < NS; n++)
		{
			const std::string key = "*" + species_name[n];
			size_t p = 0;
			for (; p < toks.size(); p++)
				if (toks[p] == "*END" || toks[p] == key)
					break;
			if (p >= toks.size() || toks[p] == "*END")
				throw std::runtime_error("species " + species_name[n] + " not found in thermal_dynamics.dat");
			p++;
			for (int r = 0; r < 3; r++)
			{ // 200-1000 K, 1000-6000 K, 6000-20000 K: a1-a7 then b1,b2
				for (int m = 0; m < 7; m++)
					Hia[n * 7 * 3 + m * 3 + r] = std::stod(toks[p++]);
				for (int m = 0; m < 2; m++)
					Hib[n * 2 * 3 + m * 3 + r] = std::stod(toks[p++]);
			}
			double W = std::stod(toks[p++]); // g/mol
			W *= 1e-3;                       // kg/mol (thermal.cpp:158)
			Wi[n] = W;
			_Wi[n] = 1.0 / Wi[n];
			Ri[n] = Ru / Wi[n];
		}
		xi_in = species_ratio_in, xi_out = species_ratio_out;
		if (cop)
		{ // mole -> mass fractions
			get_yi(species_ratio_in.data(), Wi.data(), NS);
			get_yi(species_ratio_out.data(), Wi.data(), NS);
		}
	}

	// ---- iniset.cpp:290-369 -----------------------------------------------------------------------------------------
	void Setup::init()
	{
		bl.X_inner = bl.DimX ? bl.X_inner : 1;
		bl.Y_inner = bl.DimY ? bl.Y_inner : 1;
		bl.Z_inner = bl.DimZ ? bl.Z_inner : 1;
		bl.Bwidth_X = bl.DimX ? bl.Bwidth_X : 0;
		bl.Bwidth_Y = bl.DimY ? bl.Bwidth_Y : 0;
		bl.Bwidth_Z = bl.DimZ ? bl.Bwidth_Z : 0;
		Domain_length = bl.DimX ? Domain_length : 1.0;
		Domain_width = bl.DimY ? Domain_width : 1.0;
		Domain_height = bl.DimZ ? Domain_height : 1.0;
		bl.dx = bl.DimX ? Domain_length / double(mx * bl.X_inner) : 1.0;
		bl.dy = bl.DimY ? Domain_width / double(my * bl.Y_inner) : 1.0;
		bl.dz = bl.DimZ ? Domain_height / double(mz * bl.Z_inner) : 1.0;
		bl.Xmax = bl.DimX ? (bl.X_inner + 2 * bl.Bwidth_X) : 1;
		bl.Ymax = bl.DimY ? (bl.Y_inner + 2 * bl.Bwidth_Y) : 1;
		bl.Zmax = bl.DimZ ? (bl.Z_inner + 2 * bl.Bwidth_Z) : 1;
		dl = bl.dx + bl.dy + bl.dz;
		if (bl.DimX) dl = std::min(dl, bl.dx);
		if (bl.DimY) dl = std::min(dl, bl.dy);
		if (bl.DimZ) dl = std::min(dl, bl.dz);
		bl._dx = 1.0 / bl.dx, bl._dy = 1.0 / bl.dy, bl._dz = 1.0 / bl.dz;
		// robust limiters of the species-diffusion flux (iniset.cpp:358-359)
		Dim_limiter = std::min(std::max(Dim_limiter_json, 0.0), 1.0);
		Yil_limiter = std::max(std::max(bl._dx, bl._dy), bl._dz) * std::min(std::max(Yil_limiter_json, 0.0), 1.0);
		ini._xa2 = 1.0 / (ini.xa * ini.xa);
		ini._yb2 = 1.0 / (ini.yb * ini.yb);
		ini._zc2 = 1.0 / (ini.zc * ini.zc);
		ini.C = C_json_ * mx * bl.X_inner;
		bytes = ncells() * sizeof(double), cellbytes = size_t(Emax) * bytes; // size_t: the reference's int overflows beyond ~29.8M cells
		mach_shock = Mach_Shock();
	}

	// ---- viscfit.cpp:8-140 ---------------------------------------------------------------------------------------------
	static double get_MixtureR(const Setup &s, const double *yi)
	{ // mixture.hpp:12-20 : sum of yi * Ru / Wi (a division per species, unlike get_CopR)
		double R = 0.0;
		for (int n = 0; n < s.num_species; n++)
			R += yi[n] * Ru / s.Wi[n];
		return R;
	}
	bool Setup::Mach_Shock()
	{
		double Ma_1 = ini.Ma * ini.Ma;
		if (Ma_1 < 1.0)
			return false;
		double p2 = ini.blast_pressure_out, T2 = ini.blast_T_out;
		double R = get_MixtureR(*this, species_ratio_out.data());
		double Gamma_m2 = get_CopGamma(*this, species_ratio_out.data(), T2);
		double c2 = std::sqrt(Gamma_m2 * R * T2);
		ini.blast_c_out = c2, ini.blast_gamma_out = Gamma_m2, ini.tau_H = 2.0 * ini.xa / (ini.Ma * ini.blast_c_out);
		ini.blast_density_out = p2 / R / T2;
		double rho2 = ini.blast_density_out;
		ini.blast_v_in = ini.blast_v_out;
		ini.blast_w_in = ini.blast_w_out;
		ini.cop_pressure_in = ini.blast_pressure_out;
		ini.cop_T_in = ini.blast_T_out;
		double R_cop = get_MixtureR(*this, species_ratio_in.data());
		ini.cop_density_in = ini.cop_pressure_in / R_cop / ini.cop_T_in;

		double rho1, p1, u1;
		// read_json.cpp:133: Mach_Modified = (blast_mach > 1) selects the closed-form jump of Ref0 (viscfit.cpp:124-133);
		// the iterative branch (:70-121) only runs for Ma == 1 exactly.
		if (!(ini.Ma > 1))
		{
			double Ma = ini.Ma, e2, u2, E2, Si, T1, e1, E1;
			e2 = get_Coph(*this, species_ratio_out.data(), T2) - R * T2;
			u2 = ini.blast_u_out;
			E2 = e2 + 0.5 * (u2 * u2 + ini.blast_v_out * ini.blast_v_out + ini.blast_w_out * ini.blast_w_out);
			Si = Ma * c2;
			p1 = Ma * p2;
			rho1 = Ma * ini.blast_density_out;
			T1 = T2;
			double residual = 0, threshold = 1.0e-6;
			int iter = 0;
			do
			{
				if (iter != 0)
				{
					double delta_rho = 1.0e-6 * rho1;
					rho1 += delta_rho;
					u1 = rho2 * (u2 - Si) / rho1 + Si;
					p1 = rho2 * (u2 - Si) * u2 + p2 - rho1 * (u1 - Si) * u1;
					T1 = p1 / rho1 / R;
					e1 = get_Coph(*this, species_ratio_out.data(), T1) - p1 / rho1;
					E1 = e1 + 0.5 * (u1 * u1 + ini.blast_v_in * ini.blast_v_in + ini.blast_w_in * ini.blast_w_in);
					double residual_new = rho2 * (u2 - Si) * E2 - rho1 * (u1 - Si) * E1 + p2 * u2 - p1 * u1;
					double dfdrho = (residual_new - residual) / delta_rho;
					rho1 -= delta_rho;
					rho1 = rho1 - residual / dfdrho;
				}
				if (iter > 1000)
					throw std::runtime_error("Mach number Iteration failed: Over 1000 steps has been done.");
				u1 = rho2 * (u2 - Si) / rho1 + Si;
				p1 = rho2 * (u2 - Si) * u2 + p2 - rho1 * (u1 - Si) * u1;
				T1 = p1 / rho1 / R;
				e1 = get_Coph(*this, species_ratio_out.data(), T1) - p1 / rho1;
				E1 = e1 + 0.5 * (u1 * u1 + ini.blast_v_in * ini.blast_v_in + ini.blast_w_in * ini.blast_w_in);
				residual = rho2 * (u2 - Si) * E2 - rho1 * (u1 - Si) * E1 + p2 * u2 - p1 * u1;
				iter++;
			} while (std::fabs(residual) > threshold);
		}
		else
		{
			rho1 = rho2 * (Gamma_m2 + 1.0) * Ma_1 / (2.0 + (Gamma_m2 - 1.0) * Ma_1);
			p1 = p2 * (1.0 + 2.0 * Gamma_m2 * (Ma_1 - 1.0) / (Gamma_m2 + 1.0));
			u1 = ini.Ma * c2 * (1.0 - rho2 / rho1);
		}
		ini.blast_density_in = rho1;
		ini.blast_pressure_in = p1;
		ini.blast_T_in = ini.blast_pressure_in / R / ini.blast_density_in;
		ini.blast_u_in = u1;
		return true;
	}

	void Setup::print() const
	{
		std::cout << "<--------------------------------------------------->\n"
				  << "xfluids_b200  sample: " << sample << "  mixture: " << mixture << " (species " << num_species << (ghost_species ? ", ghost" : "")
				  << ")  WENO" << weno << "  alpha: " << (artificial == 1 ? "ROE" : artificial == 2 ? "LLF" : "GLF") << "  fp_mode: " << (fp_mode ? "fast" : "strict") << "\n"
				  << "Resolution of Domain:                 " << bl.X_inner << " x " << bl.Y_inner << " x " << bl.Z_inner << "  (rank " << myRank << "/" << nRanks << ", mz " << mz << ")\n"
				  << "GhostWidth Cells: Bx, By, Bz:         " << bl.Bwidth_X << ",  " << bl.Bwidth_Y << ",  " << bl.Bwidth_Z << "\n"
				  << "XYZ dir Domain size:                  " << Domain_length << " x " << Domain_width << " x " << Domain_height << "\n"
				  << "Difference steps: dx, dy, dz:         " << bl.dx << ", " << bl.dy << ", " << bl.dz << "\n"
				  << "<--------------------------------------------------->" << std::endl;
	}

	xf_thermal Setup::thermal() const
	{
		xf_thermal t{};
		t.num_species = num_species, t.cop = cop, t.ghost_species = ghost_species, t.ncop_gamma = ncop_gamma;
		t.Hia = Hia.data(), t.Hib = Hib.data(), t.Ri = Ri.data(), t._Wi = _Wi.data();
		return t;
	}
	void Setup::rank_boundarys(int out[6]) const
	{
		for (int i = 0; i < 6; i++)
			out[i] = Boundarys[i];
		if (mz > 1)
		{ // periodic Cartesian communicator: interior faces exchange, outer faces keep the physical BC -- except that a
			// periodic z boundary is itself an exchange with the rank at the other end (mpiPacks.cpp:44-72)
			if (myMpiPos_z > 0 || Boundarys[4] == XF_BC_PERIODIC) out[4] = XF_BC_COPY;
			if (myMpiPos_z < mz - 1 || Boundarys[5] == XF_BC_PERIODIC) out[5] = XF_BC_COPY;
		}
	}

	// ---- host thermo (Thermo_device.h:10-23,62-80; Mixing_device.h:17-140) ---------------------------------------
	static const double _OT = (1.0 / 3.0);
	double HeatCapacity_NASA(const double *Hia, double T0, double Ri, int n)
	{
		double T = std::max(T0, 200.0);
		double Cpi = 0.0, _T = 1.0 / T;
		const double *a = Hia + n * 21;
		const int r = (T >= 1000.0 && T < 6000.0) ? 1 : (T < 1000.0 ? 0 : 2);
		Cpi = Ri * ((a[0 * 3 + r] * _T + a[1 * 3 + r]) * _T + a[2 * 3 + r] + (a[3 * 3 + r] + (a[4 * 3 + r] + (a[5 * 3 + r] + a[6 * 3 + r] * T) * T) * T) * T);
		return Cpi;
	}
	double get_Enthalpy_NASA(const double *Hia, const double *Hib, double T0, double Ri, int n)
	{
		double hi = 0.0, TT = T0, T = std::max(T0, 200.0);
		const double *a = Hia + n * 21, *b = Hib + n * 6;
		const int r = (T >= 1000.0 && T < 6000.0) ? 1 : (T < 1000.0 ? 0 : 2);
		hi = Ri * (-a[0 * 3 + r] / T + a[1 * 3 + r] * std::log(T) + (a[2 * 3 + r] + (0.5 * a[3 * 3 + r] + (a[4 * 3 + r] * _OT + (0.25 * a[5 * 3 + r] + 0.2 * a[6 * 3 + r] * T) * T) * T) * T) * T + b[0 * 3 + r]);
		if (TT < 200.0)
			hi += HeatCapacity_NASA(Hia, 200.0, Ri, n) * (TT - 200.0);
		return hi;
	}
	double get_CopR(const Setup &s, const double *yi)
	{
		double R = 0.0;
		for (int n = 0; n < s.num_species; n++)
			R += yi[n] * s._Wi[n];
		return R * Ru;
	}
	double get_CopCp(const Setup &s, const double *yi, double T)
	{
		double cp = 0.0;
		for (int n = 0; n < s.num_species; n++)
			cp += yi[n] * HeatCapacity_NASA(s.Hia.data(), T, s.Ri[n], n);
		return cp;
	}
	double get_CopGamma(const Setup &s, const double *yi, double T)
	{
		double Cp = get_CopCp(s, yi, T);
		double Cv = get_CopCp(s, yi, T);
		double _W = 0.0;
		for (int n = 0; n < s.num_species; n++)
			_W += yi[n] * s._Wi[n];
		Cv -= Ru * _W;
		return Cp / Cv;
	}
	double get_Coph(const Setup &s, const double *yi, double T)
	{
		double h = 0.0;
		for (int i = 0; i < s.num_species; i++)
			h += get_Enthalpy_NASA(s.Hia.data(), s.Hib.data(), T, s.Ri[i], i) * yi[i];
		return h;
	}
	void get_yi(double *xi, const double *Wi, int ns)
	{
		double W_mix = 0.0;
		for (int i = 0; i < ns; i++)
			W_mix += xi[i] * Wi[i];
		double _W_mix = 1.0 / W_mix;
		for (int n = 0; n < ns; n++)
			xi[n] = xi[n] * Wi[n] * _W_mix;
	}
} // namespace xfh
