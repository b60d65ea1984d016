// xfh_output.cpp -- field output of the XFLUIDS driver mirror in the reference's own file formats:
//   VTI_*.vti / .pvti    XFLUIDS::Output_svti  (src/XFLUIDS.cpp:1210-1387)  VTK ImageData, raw appended Float32 blocks + the parallel header
//   CVTI_*.vti / .pvti   XFLUIDS::Output_cvti  (src/XFLUIDS.cpp:1390-1555)  the same for a "compressed dimensions" stamp (-C=X,Y,0.0)
//   CPLT_*.dat           XFLUIDS::Output_cplt  (src/XFLUIDS.cpp:1676-1793)  Tecplot ASCII, point format
// selected per output stamp by the mini-language of run.OutTimeStamps / OutTimeArrays ("time: {-C=..;-P=..;-V=..}", parsed as
// src/read_ini/src/iniset.cpp:201-285 + src/read_ini/outformat/outformat.cpp do) and run.OutDAT / OutVTI / OutBoundary
// (read_json.cpp:36-40); dispatch as XFLUIDS::Output (XFLUIDS.cpp:1024-1073).  tests/test_gpu_output.py compares the files byte for
// byte with the ones the unmodified reference writes (tests/golden/out/, made by tests/golden/make_golden_out.py).
//
// Reference behaviour kept because the bytes depend on it: XFLUIDS::AllocateMemory initialises every stamp with an EMPTY variable list
// (OutFmt::Initialize's default arguments, XFLUIDS.cpp:611-612 / outformat.cpp:55-62), so a stamp's -V list never takes effect and every
// VTI file carries all variables; the single-process build never evaluates the slice position of a -C stamp (GetCPT_OutRanks sits
// behind USE_MPI), so a "compressed" file holds the whole block under a degenerate extent; -P criteria only decide whether a rank
// writes at all.  Variables: axis_x/y/z, velocity_u/v/w, rho, p, T, e (= E/rho - q^2/2, GetStates' tme), c, g (gamma), vorticity
// (viscous runs), y<k>[name].
#include <cmath>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include "xfh_driver.hpp"

namespace xfh
{
	namespace
	{
		std::vector<std::string> split(const std::string &s, char c)
		{ // Stringsplit (external/strsplit): empty fields are dropped
			std::vector<std::string> out;
			std::string cur;
			for (char ch : s)
			{
				if (ch == c)
				{
					if (!cur.empty())
						out.push_back(cur);
					cur.clear();
				}
				else
					cur += ch;
			}
			if (!cur.empty())
				out.push_back(cur);
			return out;
		}
		// AppendParas::match(str, option, split) (external/options.hpp:50-63)
		std::vector<std::string> match(const std::vector<std::string> &str, std::string option, char sp = ',')
		{
			option += "=";
			for (const std::string &t : str)
				if (t.find(option) != std::string::npos)
					return split(std::string(t).erase(0, option.length()), sp);
			return {};
		}
		struct OutString
		{ // outvars.hpp:24-47
			std::string time, step, rank;
			OutString(double t, int r, const std::string &st)
			{
				std::ostringstream a, b, c;
				a.width(11), a.fill('0'), a << t * 1E9;
				b.width(7), b.fill('0'), b << st;
				c.width(5), c.fill('0'), c << r;
				time = a.str(), step = b.str(), rank = c.str();
			}
		};
	} // namespace

	struct OutVar
	{ // outvars.hpp:49-125
		std::string name;
		const double *var = nullptr;
		size_t num = 1, num_id = 0;
		int axis = -1;
	};
	struct OutFmt
	{
		bool OutDirX = true, OutDirY = true, OutDirZ = true, CPOut = false, SPOut = false;
		double outpos[3] = {0, 0, 0};
		struct Cri
		{
			const double *var = nullptr;
			size_t num = 1, num_id = 0;
			int opera = 0;
			double line = 0;
			bool in_range(size_t id) const
			{
				const double r = var[id * num + num_id];
				switch (opera)
				{
				case 0: return r == line;
				case 1: return r < line;
				case 2: return r > line;
				case -1: return r <= line;
				default: return r >= line;
				}
			}
		};
		std::vector<Cri> cri;
	};

	struct FieldOutput::Host
	{
		std::vector<double> rho, p, T, u, v, w, c, e, gamma, y, vort; // y: [N][NS] (the reference's AoS), vort: [N]
	};

	FieldOutput::FieldOutput(XFLUIDS &x) : X(x), S(x.Ss), h(new Host)
	{
		const xf_block &b = S.bl;
		const Json &run = S.j_conf.at("run");
		OutDAT = int(run.value("OutDAT", 1.0)), OutVTI = int(run.value("OutVTI", 0.0));
		const bool OutBoundary = run.value("OutBoundary", 0.0) != 0.0;
		if (OutBoundary)
			nb[0] = b.Xmax, mn[0] = 0, mxi[0] = b.Xmax, nb[1] = b.Ymax, mn[1] = 0, mxi[1] = b.Ymax, nb[2] = b.Zmax, mn[2] = 0, mxi[2] = b.Zmax;
		else
			nb[0] = b.X_inner, mn[0] = b.Bwidth_X, mxi[0] = b.Xmax - b.Bwidth_X, nb[1] = b.Y_inner, mn[1] = b.Bwidth_Y, mxi[1] = b.Ymax - b.Bwidth_Y,
			nb[2] = b.Z_inner, mn[2] = b.Bwidth_Z, mxi[2] = b.Zmax - b.Bwidth_Z;
		prefix = "-" + S.select_dv + "-" + S.sample; // XFLUIDS.cpp:19-25
		if (b.DimZ) prefix = "Z" + prefix;
		if (b.DimY) prefix = "Y" + prefix;
		if (b.DimX) prefix = "X" + prefix;
		dir = S.OutputDir;
	}
	FieldOutput::~FieldOutput() { delete h; }

	// XFLUIDS::CopyDataFromDevice (XFLUIDS.cpp:726-828): the primitive fields in the reference's cell order
	void FieldOutput::copy_from_device()
	{
		const size_t N = S.ncells();
		const int NS = S.num_species;
		xf_ctx *ctx = X.fluids[0]->ctx;
		auto get = [&](const char *name, std::vector<double> &a)
		{
			a.resize(N);
			if (xf_get_scalar(ctx, name, a.data()) != XF_OK)
				throw std::runtime_error(std::string("xf_get_scalar ") + name + ": " + xf_last_error());
		};
		get("rho", h->rho), get("p", h->p), get("T", h->T), get("u", h->u), get("v", h->v), get("w", h->w), get("c", h->c);
		h->y.assign(N * NS, 1.0);
		if (S.cop)
		{
			std::vector<double> yk;
			for (int k = 0; k < NS; k++)
			{
				const std::string nm = "y" + std::to_string(k);
				get(nm.c_str(), yk);
				for (size_t id = 0; id < N; id++)
					h->y[id * NS + k] = yk[id];
			}
		}
		// e = tme of GetStates (Update_device.hpp:37-58) and gamma (4-argument get_CopGamma, Mixing_device.h:114-126 / NCOP_Gamma): both
		// are output-only in the reference's FlowData; formed here from the same inputs by the same operations
		// (the reference's FlowData holds the state of the LAST UpdateStates call -- the input of stage 3 of the last step, not the final U;
		// "UI4" is the energy component of that field)
		std::vector<double> E4;
		get("UI4", E4);
		h->e.resize(N), h->gamma.resize(N);
		for (size_t id = 0; id < N; id++)
		{
			const double rho1 = 1.0 / h->rho[id];
			const double u = h->u[id], v = h->v[id], w = h->w[id];
			h->e[id] = E4[id] * rho1 - 0.5 * (u * u + v * v + w * w);
			if (S.cop)
			{
				const double *yi = &h->y[id * NS];
				const double Cp = get_CopCp(S, yi, h->T[id]);
				double Wm = 0.0;
				for (int n = 0; n < NS; n++)
					Wm += yi[n] * S._Wi[n];
				const double CopW = 1.0 / Wm;
				const double g = Cp / (Cp - Ru / CopW);
				h->gamma[id] = g > 1.0 ? g : -1.0;
			}
			else
				h->gamma[id] = S.ncop_gamma;
		}
		h->vort.clear();
		if (S.Visc)
		{ // |curl u| from the velocity derivatives of the last viscous block (GetInnerCellCenterDerivativeKernel, Visc_Order_kernels.hpp:70-73)
			std::vector<double> D[9];
			for (int m = 0; m < 9; m++)
			{
				const std::string nm = "Vde" + std::to_string(m);
				get(nm.c_str(), D[m]);
			}
			h->vort.resize(N);
			for (size_t id = 0; id < N; id++)
			{
				const double wx = D[5][id] - D[7][id], wy = D[6][id] - D[2][id], wz = D[1][id] - D[3][id];
				h->vort[id] = std::sqrt(wx * wx + wy * wy + wz * wz);
			}
		}
	}

	// OutFmt::Initialize_V with an empty list = every variable (outformat.cpp:55-133)
	std::vector<OutVar> FieldOutput::variables() const
	{
		const xf_block &b = S.bl;
		std::vector<OutVar> v;
		auto add = [&](const std::string &n, const std::vector<double> &a, size_t num = 1, size_t idn = 0)
		{ OutVar o; o.name = n, o.var = a.data(), o.num = num, o.num_id = idn; v.push_back(o); };
		auto axis = [&](const std::string &n, int d)
		{ OutVar o; o.name = n, o.axis = d; v.push_back(o); };
		if (b.DimX) axis("axis_x", 0), add("velocity_u", h->u);
		if (b.DimY) axis("axis_y", 1), add("velocity_v", h->v);
		if (b.DimZ) axis("axis_z", 2), add("velocity_w", h->w);
		add("rho", h->rho), add("p", h->p), add("T", h->T), add("e", h->e), add("c", h->c), add("g", h->gamma);
		if (S.Visc && !h->vort.empty())
			add("vorticity", h->vort);
		if (S.cop)
			for (int n = 0; n < S.num_species; n++)
				add("y" + std::to_string(n + 1) + "[" + S.species_name[n] + "]", h->y, S.num_species, n);
		return v;
	}

	static OutFmt parse_fmt(const Setup &S, const std::string &spec_in, const FieldOutput::Host &h)
	{
		// iniset.cpp:201-231: a stamp without a spec gets " {-C=X,Y,Z}" with 0.0 in the inactive directions
		std::string spec = spec_in;
		if (spec.empty())
			spec = std::string(" {-C=") + (S.bl.DimX ? "X," : "0.0,") + (S.bl.DimY ? "Y," : "0.0,") + (S.bl.DimZ ? "Z}" : "0.0}");
		spec.erase(0, 2), spec.erase(spec.size() - 1, 1);
		const std::vector<std::string> items = split(spec, ';');
		OutFmt f;
		const std::vector<std::string> C = match(items, "-C"), P = match(items, "-P");
		if (C.size() >= 3)
		{ // OutFmt::Initialize_C (outformat.cpp:8-31)
			if (C[0] != "X") f.OutDirX = false, f.outpos[0] = std::stod(C[0]);
			if (C[1] != "Y") f.OutDirY = false, f.outpos[1] = std::stod(C[1]);
			if (C[2] != "Z") f.OutDirZ = false, f.outpos[2] = std::stod(C[2]);
		}
		f.CPOut = !(f.OutDirX && f.OutDirY && f.OutDirZ);
		for (const std::string &p : P)
		{ // Criterion (criterion.hpp:17-72): "<var> <op> <value>"
			const std::vector<std::string> t = split(p, ' ');
			if (t.size() < 3)
				continue;
			OutFmt::Cri c;
			const std::string &sv = t[0];
			if (sv == "rho") c.var = h.rho.data();
			else if (sv == "p" || sv == "P") c.var = h.p.data();
			else if (sv == "T") c.var = h.T.data();
			else if (sv == "u") c.var = h.u.data();
			else if (sv == "v") c.var = h.v.data();
			else if (sv == "w") c.var = h.w.data();
			else if (sv == "Gamma") c.var = h.gamma.data();
			else if (sv == "vorticity" && S.Visc && !h.vort.empty()) c.var = h.vort.data();
			else if (sv.find("yi[") != std::string::npos)
			{
				c.var = h.y.data(), c.num = S.num_species;
				for (int n = 0; n < S.num_species; n++)
					if (sv == "yi[" + S.species_name[n] + "]")
						c.num_id = n;
			}
			else
				continue;
			c.line = std::stod(t[2]);
			c.opera = t[1] == "=" ? 0 : (t[1] == "<" ? 1 : (t[1] == ">" ? 2 : (t[1] == "<=" ? -1 : -2)));
			f.cri.push_back(c);
		}
		if (!P.empty())
			f.SPOut = f.OutDirX && f.OutDirY && f.OutDirZ; // OutFmt::Initialize_P (outformat.cpp:44-46)
		return f;
	}

	// OutVar::vti_binary<float> (outvars.hpp:77-109)
	static void vti_binary(const Setup &S, const OutVar &o, std::ostream &out, const int mn[3], const int mxi[3], const int nb[3])
	{
		const xf_block &b = S.bl;
		const unsigned int nbOfWords = (unsigned int)(nb[0] * nb[1] * nb[2] * sizeof(float));
		out.write((const char *)&nbOfWords, sizeof(unsigned int));
		for (size_t k = mn[2]; k < (size_t)mxi[2]; k++)
			for (size_t j = mn[1]; j < (size_t)mxi[1]; j++)
				for (size_t i = mn[0]; i < (size_t)mxi[0]; i++)
				{
					float temp;
					if (o.axis < 0)
					{
						const size_t id = (size_t)b.Xmax * b.Ymax * k + (size_t)b.Xmax * j + i;
						temp = static_cast<float>(o.var[o.num * id + o.num_id]);
					}
					else
					{ // size_t arithmetic as written: (i - Bwidth + pos * inner + 0.5) with i unsigned
						const double ax = o.axis == 0, ay = o.axis == 1, az = o.axis == 2;
						double tmp = 0.0;
						tmp += ax * ((i - b.Bwidth_X + S.myMpiPos_x * (b.X_inner) + 0.5) * b.dx + S.Domain_xmin);
						tmp += ay * ((j - b.Bwidth_Y + S.myMpiPos_y * (b.Y_inner) + 0.5) * b.dy + S.Domain_ymin);
						tmp += az * ((k - b.Bwidth_Z + S.myMpiPos_z * (b.Z_inner) + 0.5) * b.dz + S.Domain_zmin);
						temp = static_cast<float>(tmp);
					}
					out.write((const char *)&temp, sizeof(float));
				}
	}

	// Output_svti / Output_cvti share everything but names, the degenerate extents of a -C stamp and who writes the header
	void FieldOutput::write_vti(const OutFmt &f, const std::vector<OutVar> &vars, const std::string &step, double time, bool compressed)
	{
		const xf_block &b = S.bl;
		const OutString osr(time, X.rank, step);
		const bool multi = X.nranks > 1;
		double dx = 0.0, dy = 0.0, dz = 0.0;
		int e[6] = {0, 0, 0, 0, 0, 0};
		const bool ox = !compressed || f.OutDirX, oy = !compressed || f.OutDirY, oz = !compressed || f.OutDirZ;
		if (ox && b.DimX) e[0] = S.myMpiPos_x * nb[0], e[1] = S.myMpiPos_x * nb[0] + nb[0], dx = b.dx;
		if (oy && b.DimY) e[2] = S.myMpiPos_y * nb[1], e[3] = S.myMpiPos_y * nb[1] + nb[1], dy = b.dy;
		if (oz && b.DimZ) e[4] = S.myMpiPos_z * nb[2], e[5] = S.myMpiPos_z * nb[2] + nb[2], dz = b.dz;
		const std::string kind = compressed ? "CVTI_" : "VTI_";
		const std::string temp_name = "./" + kind + prefix + "_Step_Time_" + osr.step + "." + osr.time;
		std::string file_name = dir + "/" + temp_name + (multi ? "_rank_" + osr.rank : "") + ".vti";
		// partial output: a rank without a cell that meets a criterion writes nothing (GetSPT_OutRanks, XFLUIDS.cpp:884-921)
		bool mine = true;
		if (!compressed && !f.cri.empty())
		{
			mine = false;
			for (int k = mn[2]; k < mxi[2] && !mine; k++)
				for (int j = mn[1]; j < mxi[1] && !mine; j++)
					for (int i = mn[0]; i < mxi[0] && !mine; i++)
					{
						const size_t id = (size_t)b.Xmax * b.Ymax * k + (size_t)b.Xmax * j + i;
						for (const OutFmt::Cri &c : f.cri)
							if (c.in_range(id))
							{
								mine = true;
								break;
							}
					}
		}
		const int dims = b.DimX + b.DimY + b.DimZ;
		if (X.rank == 0 && (compressed || mine))
		{ // .pvti header (rank 0; pieces: every rank of the z-slab decomposition)
			const std::string header = dir + "/" + kind + prefix + "_Time_" + osr.time + ".pvti";
			std::ofstream o(header);
			o << "<?xml version=\"1.0\"?>" << std::endl;
			o << "<VTKFile type=\"PImageData\" version=\"0.1\" byte_order=\"LittleEndian\">" << std::endl;
			o << "  <PImageData WholeExtent=\"";
			if (compressed)
				o << 0 << " " << (f.OutDirX ? S.mx : 0) * nb[0] << " " << 0 << " " << (f.OutDirY ? S.my : 0) * nb[1] << " " << 0 << " " << (f.OutDirZ ? S.mz : 0) * nb[2];
			else
				o << 0 << " " << S.mx * nb[0] * int(b.DimX) << " " << 0 << " " << S.my * nb[1] * int(b.DimY) << " " << 0 << " " << S.mz * nb[2] * int(b.DimZ);
			o << "\" GhostLevel=\"0\" Origin=\"" << S.Domain_xmin << " " << S.Domain_ymin << " " << S.Domain_zmin << "\" Spacing=\"" << dx << " " << dy << " " << dz << "\">" << std::endl;
			o << "    <PCellData Scalars=\"Scalars_\">" << std::endl;
			for (const OutVar &v : vars)
				o << "      <PDataArray type=\"Float32\" Name=\"" << v.name << "\"/>" << std::endl;
			o << "    </PCellData>" << std::endl;
			for (int r = 0; r < X.nranks; r++)
			{
				std::ostringstream pf;
				pf.width(5), pf.fill('0'), pf << r;
				const std::string piece = temp_name + (multi ? "_rank_" + pf.str() : "") + ".vti";
				const int coords[3] = {0, 0, multi ? r : 0};
				const int On[3] = {(!compressed || f.OutDirX) ? nb[0] : 0, (!compressed || f.OutDirY) ? nb[1] : 0, (!compressed || f.OutDirZ) ? nb[2] : 0};
				o << " <Piece Extent=\"";
				auto ext = [&](int d_) { if (coords[d_] == 0) o << 0 << " " << On[d_] << " "; else o << coords[d_] * On[d_] << " " << coords[d_] * On[d_] + On[d_] << " "; };
				if (compressed)
				{
					if (b.DimX) ext(0); else o << 0 << " " << 0;
					if (b.DimY) ext(1); else o << 0 << " " << 0;
					if (b.DimZ) ext(2); else o << 0 << " " << 0;
				}
				else
				{
					ext(0), ext(1);
					if (dims == 3) ext(2); else o << 0 << " " << 0;
				}
				o << "\" Source=\"" << piece << "\"/>" << std::endl;
			}
			o << "</PImageData>" << std::endl;
			o << "</VTKFile>" << std::endl;
		}
		if (!mine)
			return;
		std::ofstream o(file_name, std::ios::binary);
		if (!o)
			throw std::runtime_error("cannot write " + file_name);
		o << "<VTKFile type=\"ImageData\" version=\"0.1\" byte_order=\"LittleEndian\">\n";
		o << "  <ImageData WholeExtent=\"" << e[0] << " " << e[1] << " " << e[2] << " " << e[3] << " " << e[4] << " " << e[5] << "\" Origin=\"" << S.Domain_xmin << " " << S.Domain_ymin
		  << " " << S.Domain_zmin << "\" Spacing=\"" << dx << " " << dy << " " << dz << "\">" << std::endl;
		o << "  <Piece Extent=\"" << e[0] << " " << e[1] << " " << e[2] << " " << e[3] << " " << e[4] << " " << e[5] << "\">" << std::endl;
		o << "    <PointData>\n    </PointData>\n";
		o << "    <CellData>" << std::endl;
		for (size_t iv = 0; iv < vars.size(); iv++)
			o << "     <DataArray type=\"Float32\" Name=\"" << vars[iv].name << "\" format=\"appended\" offset=\"" << iv * nb[0] * nb[1] * nb[2] * sizeof(float) + iv * sizeof(unsigned int)
			  << "\" />" << std::endl;
		o << "    </CellData>" << std::endl;
		o << "  </Piece>" << std::endl;
		o << "  </ImageData>" << std::endl;
		o << "  <AppendedData encoding=\"raw\">" << std::endl;
		o << "_";
		for (const OutVar &v : vars)
			vti_binary(S, v, o, mn, mxi, nb);
		o << "  </AppendedData>" << std::endl;
		o << "</VTKFile>" << std::endl;
	}

	// Output_cplt (XFLUIDS.cpp:1676-1793)
	void FieldOutput::write_cplt(const std::string &step, double time)
	{
		const xf_block &b = S.bl;
		const OutString osr(time, X.rank, step);
		const int NS = S.num_species;
		const int NSout = S.num_species - (S.ghost_species ? 1 : 0); // Block::num_species: the ghost species is not listed (iniset.cpp:348-352)
		const double temx = 0.5 * b.dx + S.Domain_xmin, temy = 0.5 * b.dy + S.Domain_ymin, temz = 0.5 * b.dz + S.Domain_zmin;
		const double posx = -b.Bwidth_X + S.myMpiPos_x * (b.X_inner), posy = -b.Bwidth_Y + S.myMpiPos_y * (b.Y_inner), posz = -b.Bwidth_Z + S.myMpiPos_z * (b.Z_inner);
		const std::string file_name = dir + "/CPLT_" + prefix + "_Step_Time_" + osr.step + "." + osr.time + "_" + osr.rank + ".dat";
		std::ofstream out(file_name);
		if (!out)
			throw std::runtime_error("cannot write " + file_name);
		const bool vort = S.Visc && (b.DimX + b.DimY + b.DimZ > 1) && !h->vort.empty();
		out << "title='" << prefix << "'\nvariables=";
		if (b.DimX) out << "x[m], ";
		if (b.DimY) out << "y[m], ";
		if (b.DimZ) out << "z[m], ";
		out << "<i><greek>r</greek></i>[kg/m<sup>3</sup>], <i>p</i>[Pa], <i>c</i>[m/s]";
		if (b.DimX) out << ", <i>u</i>[m/s]";
		if (b.DimY) out << ", <i>v</i>[m/s]";
		if (b.DimZ) out << ", <i>w</i>[m/s]";
		out << ", <i><greek>g</greek></i>[-], <i>T</i>[K], <i>e</i>[J]";
		if (vort)
			out << ", <i><greek>w</greek></i>|[s<sup>-1</sup>], <i><greek>w</greek></i><sub>x</sub>[s<sup>-1</sup>], <i><greek>w</greek></i><sub>y</sub>[s<sup>-1</sup>], <i><greek>w</greek></i><sub>z</sub>[s<sup>-1</sup>]";
		if (S.cop)
			for (int n = 0; n < NSout; n++)
				out << ", <i>Y(" << S.species_name[n] << ")</i>[-]";
		out << "\n";
		out << "zone t='" << prefix << "_" << osr.time;
		if (X.nranks > 1)
			out << "_rank_" << std::to_string(X.rank);
		out << "', i= " << nb[0] << ", j= " << nb[1] << ", k= " << nb[2] << ", SOLUTIONTIME= " << osr.time << "\n";
		for (int k = mn[2]; k < mxi[2]; k++)
			for (int j = mn[1]; j < mxi[1]; j++)
				for (int i = mn[0]; i < mxi[0]; i++)
				{
					const size_t id = (size_t)b.Xmax * b.Ymax * k + (size_t)b.Xmax * j + i;
					const int pos_x = int(i + posx), pos_y = int(j + posy), pos_z = int(k + posz);
					if (b.DimX) out << (pos_x)*b.dx + temx << " ";
					if (b.DimY) out << (pos_y)*b.dy + temy << " ";
					if (b.DimZ) out << (pos_z)*b.dz + temz << " ";
					out << h->rho[id] << " " << h->p[id] << " " << h->c[id] << " ";
					if (b.DimX) out << h->u[id] << " ";
					if (b.DimY) out << h->v[id] << " ";
					if (b.DimZ) out << h->w[id] << " ";
					out << h->gamma[id] << " " << h->T[id] << " " << h->e[id] << " ";
					if (S.cop)
						for (int n = 0; n < NSout; n++)
							out << h->y[id * NS + n] << " ";
					out << "\n";
				}
	}

	// XFLUIDS::Output (XFLUIDS.cpp:1024-1073)
	void FieldOutput::output(const std::string &spec, double time, const std::string &step)
	{
		if (!OutDAT && !OutVTI)
			return;
		copy_from_device();
		const OutFmt f = parse_fmt(S, spec, *h);
		const std::vector<OutVar> vars = variables();
		const char *what;
		if (f.CPOut)
		{
			if (OutDAT) write_cplt(step, time);
			if (OutVTI) write_vti(f, vars, step, time, true);
			what = "Compress Dimensions solution";
		}
		else if (f.SPOut)
		{
			if (OutVTI) write_vti(f, vars, step, time, false);
			what = "Partial Domain solution";
		}
		else
		{
			if (OutDAT) write_cplt(step, time);
			if (OutVTI) write_vti(f, vars, step, time, false);
			what = "Common Domain solution";
		}
		if (X.rank == 0 && X.verbose)
			std::cout << what << " has been done at Step = " << step << ", Time = " << time << std::endl;
	}
} // namespace xfh
