/* Control-file front end with the reference's names: FlowParserOptions, parse_flow_controlfile,
 * extract_spatial_physics_config / extract_spatial_numerics_config / firstorder_spatial_numerics_config
 * (reference: src/utilities/controlparser.{hpp,cpp}:60-290). The reference reads the file with
 * Boost.PropertyTree's INFO parser; Boost is not available here, so InfoTree below parses the subset of the INFO
 * grammar the reference's control files use - `key value`, `key { ... }`, quoted strings, `;` comments and the
 * `#include "file"` directive (the control files under tests/inv-2dcyl) - with Boost's state machine: after a key's value, further
 * words on the same line start NEW keys (so `boundary_values 0.0 290.0` yields the value "0.0", as in the
 * reference; quote the list to get both numbers).
 *
 * Differences from the reference, on purpose:
 *  - `spatial_discretization.limiter_parameter` IS read (the reference never assigns
 *    FlowParserOptions::limiter_param and uses an indeterminate value); default 1.0 when absent;
 *  - command-line overrides come as a string map (no Boost.ProgramOptions): "mesh_file", "source_dir"
 *    (substituted for @CMAKE_SOURCE_DIR@ in #include paths and mesh paths, which CMake does for the reference's tests);
 *  - implicit pseudo-time options are parsed but the case driver rejects them (out of scope: needs PETSc).
 */
#ifndef FVENS_B200_CONTROLPARSER_HPP
#define FVENS_B200_CONTROLPARSER_HPP

#include "fvens_b200.hpp"
#include <fstream>
#include <sstream>
#include <algorithm>
#include <cctype>

namespace fvens {

#ifndef FVENS_B200_PI
#define FVENS_B200_PI
constexpr double PI = 3.14159265358979323846;
#endif

class InputNotGivenError : public std::runtime_error {
public:
	InputNotGivenError(const std::string& msg) : std::runtime_error(msg) {}
};

/// A node of the INFO tree: data string + ordered children (duplicate keys allowed; lookups return the first)
class InfoTree {
public:
	std::string data;
	std::vector<std::pair<std::string, InfoTree>> children;

	const InfoTree* find(const std::string& path) const {
		const InfoTree *cur = this;
		size_t pos = 0;
		while(pos <= path.size()) {
			const size_t dot = path.find('.', pos);
			const std::string key = path.substr(pos, dot == std::string::npos ? std::string::npos : dot - pos);
			const InfoTree *next = nullptr;
			for(const auto& c : cur->children) if(c.first == key) { next = &c.second; break; }
			if(!next) return nullptr;
			cur = next;
			if(dot == std::string::npos) break;
			pos = dot + 1;
		}
		return cur;
	}
	bool has(const std::string& path) const { return find(path) != nullptr; }

	/// ptree::get<T>(path): throws if the key is missing or the value does not convert
	template <typename T> T get(const std::string& path) const {
		const InfoTree *n = find(path);
		if(!n) throw InputNotGivenError("No such node (" + path + ")");
		return convert<T>(n->data, path);
	}
	/// ptree::get(path, default)
	template <typename T> T get(const std::string& path, const T& dflt) const {
		const InfoTree *n = find(path);
		return n ? convert<T>(n->data, path) : dflt;
	}

	static InfoTree read_info(const std::string& file, const std::string& source_dir = "") {
		InfoTree root;
		std::vector<InfoTree*> stack(1, &root);
		InfoTree *last = nullptr;
		bool expect_data = false;
		parse_file(file, source_dir, stack, last, expect_data, 0);
		if(stack.size() != 1) throw std::runtime_error(file + ": unmatched '{'");
		return root;
	}

private:
	template <typename T> static T convert(const std::string& s, const std::string& path);

	static std::string substitute(std::string s, const std::string& source_dir) {
		const std::string key = "@CMAKE_SOURCE_DIR@";
		for(size_t p = s.find(key); p != std::string::npos; p = s.find(key, p)) s.replace(p, key.size(), source_dir);
		return s;
	}

	/// One line -> tokens; a quoted string is one token (escapes \" \\ \n \t), ';' outside quotes ends the line
	static std::vector<std::pair<std::string,bool>> tokenize(const std::string& line, const std::string& where) {
		std::vector<std::pair<std::string,bool>> toks;      // (text, was quoted)
		size_t i = 0;
		while(i < line.size()) {
			if(std::isspace((unsigned char)line[i])) { i++; continue; }
			if(line[i] == ';') break;
			if(line[i] == '"') {
				std::string s; i++;
				bool closed = false;
				while(i < line.size()) {
					if(line[i] == '\\' && i + 1 < line.size()) {
						const char c = line[i+1];
						s += c == 'n' ? '\n' : (c == 't' ? '\t' : c);
						i += 2;
					}
					else if(line[i] == '"') { closed = true; i++; break; }
					else s += line[i++];
				}
				if(!closed) throw std::runtime_error(where + ": unterminated string");
				toks.push_back(std::make_pair(s, true));
			}
			else if(line[i] == '{' || line[i] == '}') { toks.push_back(std::make_pair(std::string(1, line[i]), false)); i++; }
			else {
				const size_t st = i;
				while(i < line.size() && !std::isspace((unsigned char)line[i]) && line[i] != ';' && line[i] != '{' && line[i] != '}') i++;
				toks.push_back(std::make_pair(line.substr(st, i - st), false));
			}
		}
		return toks;
	}

	static void parse_file(const std::string& file, const std::string& source_dir, std::vector<InfoTree*>& stack,
	                       InfoTree*& last, bool& expect_data, const int depth) {
		if(depth > 16) throw std::runtime_error(file + ": #include nesting too deep");
		std::ifstream in(file);
		if(!in) throw std::runtime_error("cannot open control file " + file);
		std::string line;
		int lineno = 0;
		while(std::getline(in, line)) {
			lineno++;
			const std::string where = file + "(" + std::to_string(lineno) + ")";
			size_t f = line.find_first_not_of(" \t\r");
			if(f != std::string::npos && line.compare(f, 8, "#include") == 0) {
				if(expect_data) throw std::runtime_error(where + ": #include where a value was expected");
				const auto t = tokenize(line.substr(f + 8), where);
				if(t.size() != 1 || !t[0].second) throw std::runtime_error(where + ": #include needs one quoted file name");
				parse_file(substitute(t[0].first, source_dir), source_dir, stack, last, expect_data, depth + 1);
				continue;
			}
			for(const auto& tk : tokenize(line, where)) {
				const bool brace_open = !tk.second && tk.first == "{", brace_close = !tk.second && tk.first == "}";
				if(brace_open) {
					if(!last) throw std::runtime_error(where + ": unexpected {");
					stack.push_back(last); last = nullptr; expect_data = false;
				}
				else if(brace_close) {
					if(stack.size() <= 1) throw std::runtime_error(where + ": unmatched }");
					stack.pop_back(); last = nullptr; expect_data = false;
				}
				else if(expect_data) { last->data = tk.first; expect_data = false; }
				else {
					stack.back()->children.push_back(std::make_pair(tk.first, InfoTree()));
					last = &stack.back()->children.back().second;
					expect_data = true;
				}
			}
		}
	}
};

template <> inline std::string InfoTree::convert<std::string>(const std::string& s, const std::string&) { return s; }
template <> inline double InfoTree::convert<double>(const std::string& s, const std::string& path) {
	size_t used = 0; double v = 0;
	try { v = std::stod(s, &used); } catch(std::exception&) { used = 0; }
	if(used == 0 || used != s.size()) throw std::runtime_error("conversion of data to a number failed at " + path + ": '" + s + "'");
	return v;
}
template <> inline int InfoTree::convert<int>(const std::string& s, const std::string& path) {
	size_t used = 0; long v = 0;
	try { v = std::stol(s, &used); } catch(std::exception&) { used = 0; }
	if(used == 0 || used != s.size()) throw std::runtime_error("conversion of data to an integer failed at " + path + ": '" + s + "'");
	return (int)v;
}
template <> inline bool InfoTree::convert<bool>(const std::string& s, const std::string& path) {
	if(s == "true" || s == "1") return true;
	if(s == "false" || s == "0") return false;
	throw std::runtime_error("conversion of data to bool failed at " + path + ": '" + s + "'");
}

inline std::string to_upper_copy(std::string s) { for(char& c : s) c = (char)std::toupper((unsigned char)c); return s; }
inline std::string to_lower_copy(std::string s) { for(char& c : s) c = (char)std::tolower((unsigned char)c); return s; }

/// Reference: parseStringToVector (controlparser.cpp:42-53): space-separated list
template <typename T> inline std::vector<T> parseStringToVector(const std::string& str) {
	T elem; std::vector<T> vec; std::stringstream ss(str);
	while(ss >> elem) vec.push_back(elem);
	return vec;
}

/// Reference: bcTypeMap (spatial/abctypemap.cpp:16-33)
inline BCType bcTypeFromString(const std::string& name) {
	static const std::pair<const char*, BCType> tab[] = {
		{"slipwall", SLIP_WALL_BC}, {"isothermalwall", ISOTHERMAL_WALL_BC}, {"adiabaticwall", ADIABATIC_WALL_BC},
		{"farfield", FARFIELD_BC}, {"inflowoutflow", INFLOW_OUTFLOW_BC}, {"subsonic_inflow", SUBSONIC_INFLOW_BC},
		{"extrapolation", EXTRAPOLATION_BC}, {"periodic", PERIODIC_BC}};
	for(const auto& e : tab) if(name == e.first) return e.second;
	throw std::out_of_range("bcTypeMap: unknown boundary condition type '" + name + "'");     // bimap::at throws out_of_range
}

/// Reference: FlowParserOptions (utilities/controlparser.hpp:19-76), same field names
struct FlowParserOptions {
	std::string meshfile, vtu_output_file, logfile, flowtype, init_soln_file, invflux, invfluxjac, gradientmethod, limiter,
		pseudotimetype, constvisc, surfnameprefix, volnameprefix, vol_output_reqd, sim_type, time_integrator, nl_update_scheme;
	freal initcfl = 0, endcfl = 0, tolerance = 0, firstinitcfl = 0, firstendcfl = 0, firsttolerance = 0,
		Minf = 0, alpha = 0, Reinf = 0, Tinf = 0, Pr = 0, gamma = 0, limiter_param = 1.0, final_time = 0, phy_timestep = 0, phy_cfl = 0;
	freal min_nl_update = 0.2;
	int maxiter = 0, rampstart = 0, rampend = 0, firstmaxiter = 0, firstrampstart = 0, firstrampend = 0,
		num_out_walls = 0, num_out_others = 0, time_order = 1;
	std::vector<FlowBCConfig> bcconf;
	short soln_init_type = 0, usestarter = 0;
	bool lognres = false, write_final_lin_sys = false, useconstvisc = false, viscsim = false, order2 = true;
	std::vector<int> lwalls, lothers;
};

namespace detail {
inline std::string get_upperCaseString(const InfoTree& tree, const std::string& path) {
	return to_upper_copy(tree.get<std::string>(path));
}

/// Reference: parse_BC_options (controlparser.cpp:242-290): sections bc0, bc1, ... numbered consecutively
inline std::vector<FlowBCConfig> parse_BC_options(const InfoTree& infopts, const std::string& c_bcs) {
	std::vector<FlowBCConfig> bcvec;
	for(int ibc = 0; ; ibc++) {
		const std::string base = c_bcs + ".bc" + std::to_string(ibc);
		if(!infopts.has(base)) break;
		FlowBCConfig bconf;
		bconf.bc_type = bcTypeFromString(to_lower_copy(infopts.get<std::string>(base + ".type")));
		bconf.bc_tag = infopts.get<int>(base + ".marker");
		if(bconf.bc_type == ADIABATIC_WALL_BC || bconf.bc_type == ISOTHERMAL_WALL_BC || bconf.bc_type == SUBSONIC_INFLOW_BC)
			bconf.bc_vals = parseStringToVector<freal>(infopts.get<std::string>(base + ".boundary_values"));
		if(bconf.bc_type == PERIODIC_BC)
			bconf.bc_opts = parseStringToVector<int>(infopts.get<std::string>(base + ".options"));
		bcvec.push_back(bconf);
	}
	return bcvec;
}
}

/// Reference: parse_flow_controlfile (controlparser.cpp:60-216). cmdvars: "mesh_file" overrides io.mesh_file,
/// "write_final_linear_system" as in the reference, "source_dir" replaces @CMAKE_SOURCE_DIR@, "log_file_prefix"
/// stands for the PETSc option -fvens_log_file_prefix.
inline FlowParserOptions parse_flow_controlfile(const std::string& controlfile,
                                                const std::map<std::string,std::string>& cmdvars = std::map<std::string,std::string>())
{
	using detail::get_upperCaseString;
	FlowParserOptions opts;
	opts.time_integrator = "NONE";
	const std::string c_io = "io", c_flowconds = "flow_conditions", c_bcs = "bc", c_phy_time = "time",
		c_spatial = "spatial_discretization", c_pseudotime = "pseudotime";
	const std::string pt_main = "main", pt_init = "initialization";
	const auto srcdir = cmdvars.find("source_dir");
	const std::string source_dir = srcdir != cmdvars.end() ? srcdir->second : std::string(".");

	const InfoTree infopts = InfoTree::read_info(controlfile, source_dir);

	opts.meshfile = infopts.get<std::string>("io.mesh_file");
	if(cmdvars.count("mesh_file")) {
		std::cout << "Read mesh file from the command line rather than the control file.\n";
		opts.meshfile = cmdvars.at("mesh_file");
	}
	{
		const std::string key = "@CMAKE_SOURCE_DIR@";
		for(size_t p = opts.meshfile.find(key); p != std::string::npos; p = opts.meshfile.find(key, p)) opts.meshfile.replace(p, key.size(), source_dir);
	}
	opts.write_final_lin_sys = cmdvars.count("write_final_linear_system") && cmdvars.at("write_final_linear_system") == "true";

	opts.vtu_output_file = infopts.get<std::string>(c_io + ".solution_output_file");
	opts.logfile = infopts.get<std::string>(c_io + ".log_file_prefix");
	opts.lognres = infopts.get<bool>(c_io + ".convergence_history_required");

	opts.flowtype = get_upperCaseString(infopts, c_flowconds + ".flow_type");
	opts.gamma = infopts.get<freal>(c_flowconds + ".adiabatic_index");
	opts.alpha = PI/180.0*infopts.get<freal>(c_flowconds + ".angle_of_attack");
	opts.Minf = infopts.get<freal>(c_flowconds + ".freestream_Mach_number");
	if(opts.flowtype == "NAVIERSTOKES" || opts.flowtype == "RANS") {
		opts.viscsim = true;
		opts.Tinf = infopts.get<freal>(c_flowconds + ".freestream_temperature");
		opts.Reinf = infopts.get<freal>(c_flowconds + ".freestream_Reynolds_number");
		opts.Pr = infopts.get<freal>(c_flowconds + ".Prandtl_number");
		opts.useconstvisc = infopts.get<bool>(c_flowconds + ".use_constant_viscosity", false);
		if(opts.flowtype == "RANS") throw UnsupportedOptionError("flowtype RANS");
	}
	else {
		// the reference stores 1/0 and 0/0 here; they are never used by an Euler run
		opts.viscsim = false; opts.Tinf = 298.0; opts.Reinf = HUGE_VAL; opts.Pr = std::nan("");
	}

	opts.bcconf = detail::parse_BC_options(infopts, c_bcs);

	if(infopts.has(c_bcs + ".listof_output_wall_boundaries"))
		opts.lwalls = parseStringToVector<int>(infopts.get<std::string>(c_bcs + ".listof_output_wall_boundaries"));
	opts.num_out_walls = (int)opts.lwalls.size();
	if(infopts.has(c_bcs + ".listof_output_other_boundaries"))
		opts.lothers = parseStringToVector<int>(infopts.get<std::string>(c_bcs + ".listof_output_other_boundaries"));
	opts.num_out_others = (int)opts.lothers.size();
	if(opts.num_out_others > 0 || opts.num_out_walls > 0)
		opts.surfnameprefix = infopts.get<std::string>(c_bcs + ".surface_output_file_prefix");
	if(infopts.has(c_bcs + ".volume_output_file_prefix")) {
		opts.volnameprefix = infopts.get<std::string>(c_bcs + ".volume_output_file_prefix");
		opts.vol_output_reqd = "YES";
	} else opts.vol_output_reqd = "NO";

	opts.sim_type = get_upperCaseString(infopts, c_phy_time + ".simulation_type");
	if(opts.sim_type == "UNSTEADY") {
		opts.final_time = infopts.get<freal>(c_phy_time + ".final_time");
		opts.time_integrator = get_upperCaseString(infopts, c_phy_time + ".time_integrator");
		opts.time_order = infopts.get<int>(c_phy_time + ".temporal_order");
		if(opts.time_integrator == "TVDRK") opts.phy_cfl = infopts.get<freal>(c_phy_time + ".physical_cfl");
		else opts.phy_timestep = infopts.get<freal>(c_phy_time + ".physical_time_step");
	}

	opts.invflux = get_upperCaseString(infopts, c_spatial + ".inviscid_flux");
	opts.gradientmethod = get_upperCaseString(infopts, c_spatial + ".gradient_method");
	if(opts.gradientmethod == "NONE") opts.order2 = false;
	opts.limiter = get_upperCaseString(infopts, c_spatial + ".limiter");
	opts.limiter_param = infopts.get<freal>(c_spatial + ".limiter_parameter", 1.0);

	opts.pseudotimetype = get_upperCaseString(infopts, c_pseudotime + ".pseudotime_stepping_type");
	opts.initcfl = infopts.get<freal>(c_pseudotime + "." + pt_main + ".cfl_min");
	opts.endcfl = infopts.get<freal>(c_pseudotime + "." + pt_main + ".cfl_max");
	opts.tolerance = infopts.get<freal>(c_pseudotime + "." + pt_main + ".tolerance");
	opts.maxiter = infopts.get<int>(c_pseudotime + "." + pt_main + ".max_timesteps");
	if(infopts.has(c_pseudotime + "." + pt_init)) {
		opts.usestarter = 1;
		opts.firstinitcfl = infopts.get<freal>(c_pseudotime + "." + pt_init + ".cfl_min");
		opts.firstendcfl = infopts.get<freal>(c_pseudotime + "." + pt_init + ".cfl_max");
		opts.firsttolerance = infopts.get<freal>(c_pseudotime + "." + pt_init + ".tolerance");
		opts.firstmaxiter = infopts.get<int>(c_pseudotime + "." + pt_init + ".max_timesteps");
	}
	if(opts.pseudotimetype == "IMPLICIT") {
		opts.invfluxjac = get_upperCaseString(infopts, "Jacobian_inviscid_flux");
		if(opts.invfluxjac == "CONSISTENT") opts.invfluxjac = opts.invflux;
		opts.nl_update_scheme = get_upperCaseString(infopts, c_pseudotime + ".nonlinear_update_scheme");
		opts.min_nl_update = infopts.get<freal>(c_pseudotime + ".min_nonlinear_relaxation_factor", 0.2);
	}
	if(cmdvars.count("log_file_prefix")) opts.logfile = cmdvars.at("log_file_prefix");
	return opts;
}

/// Reference: controlparser.cpp:218-240
inline FlowPhysicsConfig extract_spatial_physics_config(const FlowParserOptions& opts) {
	const FlowPhysicsConfig pconf { opts.gamma, opts.Minf, opts.Tinf, opts.Reinf, opts.Pr, opts.alpha,
		opts.viscsim, opts.useconstvisc, opts.bcconf };
	return pconf;
}
inline FlowNumericsConfig extract_spatial_numerics_config(const FlowParserOptions& opts) {
	const FlowNumericsConfig nconf { opts.invflux, opts.invfluxjac, opts.gradientmethod, opts.limiter, opts.limiter_param, opts.order2 };
	return nconf;
}
inline FlowNumericsConfig firstorder_spatial_numerics_config(const FlowParserOptions& opts) {
	const FlowNumericsConfig nconf { opts.invflux, opts.invfluxjac, "NONE", "NONE", 1.0, false };
	return nconf;
}

}
#endif
