// xfh_json.hpp -- a small JSON reader for XFluids' settings/*.json (reference parser: src/read_ini/settings/read_json.cpp:150-178).
// Keeps the reference's comment-stripping quirk: every line is cut at its FIRST '/' character (find_first_of("//")),
// which is why no value in those files may contain a slash.
#pragma once
#include <cctype>
#include <cstdlib>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace xfh
{
	struct Json
	{
		enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
		bool b = false;
		double num = 0;
		std::string str;
		std::vector<Json> arr;
		std::map<std::string, Json> obj;

		bool has(const std::string &k) const { return kind == Obj && obj.count(k); }
		const Json &at(const std::string &k) const
		{
			static const Json empty_obj = [] { Json j; j.kind = Obj; return j; }();
			auto it = obj.find(k);
			return (kind == Obj && it != obj.end()) ? it->second : empty_obj; // a missing section behaves like {}
		}
		double value(const std::string &k, double d) const { return has(k) && (obj.at(k).kind == Num || obj.at(k).kind == Bool) ? (obj.at(k).kind == Num ? obj.at(k).num : double(obj.at(k).b)) : d; }
		bool value(const std::string &k, bool d) const { return has(k) ? (obj.at(k).kind == Bool ? obj.at(k).b : obj.at(k).num != 0) : d; }
		std::string value(const std::string &k, const char *d) const { return has(k) && obj.at(k).kind == Str ? obj.at(k).str : std::string(d); }
		std::vector<double> value(const std::string &k, std::vector<double> d) const
		{
			if (!has(k) || obj.at(k).kind != Arr)
				return d;
			std::vector<double> r;
			for (auto &e : obj.at(k).arr)
				r.push_back(e.num);
			return r;
		}
		std::vector<std::string> strings(const std::string &k) const
		{
			std::vector<std::string> r;
			if (has(k) && obj.at(k).kind == Arr)
				for (auto &e : obj.at(k).arr)
					r.push_back(e.str);
			return r;
		}
	};

	class JsonParser
	{
		const std::string &s;
		size_t p = 0;
		void ws()
		{
			while (p < s.size() && std::isspace((unsigned char)s[p]))
				p++;
		}
		[[noreturn]] void bad(const char *m) { throw std::runtime_error(std::string("json: ") + m + " at offset " + std::to_string(p)); }
		Json val()
		{
			ws();
			if (p >= s.size())
				bad("unexpected end");
			Json j;
			char c = s[p];
			if (c == '{')
			{
				j.kind = Json::Obj, p++, ws();
				if (s[p] == '}') { p++; return j; }
				for (;;)
				{
					ws();
					if (s[p] != '"') bad("key expected");
					std::string k = str();
					ws();
					if (s[p++] != ':') bad("':' expected");
					j.obj[k] = val();
					ws();
					if (s[p] == ',') { p++; continue; }
					if (s[p] == '}') { p++; break; }
					bad("',' or '}' expected");
				}
			}
			else if (c == '[')
			{
				j.kind = Json::Arr, p++, ws();
				if (s[p] == ']') { p++; return j; }
				for (;;)
				{
					j.arr.push_back(val());
					ws();
					if (s[p] == ',') { p++; continue; }
					if (s[p] == ']') { p++; break; }
					bad("',' or ']' expected");
				}
			}
			else if (c == '"')
				j.kind = Json::Str, j.str = str();
			else if (!s.compare(p, 4, "true"))
				j.kind = Json::Bool, j.b = true, p += 4;
			else if (!s.compare(p, 5, "false"))
				j.kind = Json::Bool, j.b = false, p += 5;
			else if (!s.compare(p, 4, "null"))
				p += 4;
			else
			{
				char *e = nullptr;
				j.kind = Json::Num, j.num = std::strtod(s.c_str() + p, &e);
				if (e == s.c_str() + p) bad("value expected");
				p = e - s.c_str();
			}
			return j;
		}
		std::string str()
		{
			std::string r;
			p++;
			while (p < s.size() && s[p] != '"')
			{
				if (s[p] == '\\' && p + 1 < s.size())
					p++;
				r += s[p++];
			}
			p++;
			return r;
		}

	public:
		explicit JsonParser(const std::string &text) : s(text) {}
		Json parse() { return val(); }
	};

	inline Json ReadJson(const std::string &filename)
	{
		std::ifstream in(filename.c_str());
		if (!in)
			throw std::runtime_error("error while reading json configure file " + filename);
		std::stringstream out;
		std::string line;
		while (std::getline(in, line))
			out << line.substr(0, line.find_first_of("//")) << "\n"; // the reference's comment rule (first '/')
		const std::string text = out.str();
		return JsonParser(text).parse();
	}
} // namespace xfh
