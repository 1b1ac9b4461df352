/* C++ test of the reference class surface in fvens_b200/host/fvens_b200.hpp. Needs a B200 (every call below
 * launches kernels through libfvens_b200.so). Run by tests/test_gpu_cpp_surface.py:
 *     test_host_surface <mesh dir>
 * The cases read like the reference's own unit tests:
 *   flux KAT              a Roe flux produced by the reference's object code (tier A probe)
 *   wall BCs              tests/flow-general/testwallbcs.cpp:14-79 (zero mass/energy flux through walls)
 *   1-exact gradients     tests/finite-volume/testgradientschemes.cpp:36-90 (WLS + linear reconstruction)
 *   FlowFV                compute_residual contract (adds into the residual; host Vec == device Vec)
 *   forward Euler         ode/aodesolver.cpp:136-282 (Tolerance_error at maxiter, convergence history)
 *   factories             utilities/afactory.cpp (unknown keys), spatial/abc.cpp:493 (unknown BC throws)
 */
#include "../../fvens_b200/host/fvens_b200.hpp"
#include <cstdio>
#include <limits>

using namespace fvens;

static int nfail = 0;
#define CHECK(cond, msg) do { if(!(cond)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, msg); nfail++; } } while(0)

static FlowPhysicsConfig inviscid_physics() {
	FlowPhysicsConfig p;
	p.gamma = 1.4; p.Minf = 0.8; p.Tinf = 288.15; p.Reinf = 5000.0; p.Pr = 0.72; p.aoa = 1.25*M_PI/180.0;
	p.viscous_sim = false; p.const_visc = false;
	p.bcconf = { {2, SLIP_WALL_BC, {}, {}}, {4, FARFIELD_BC, {}, {}} };
	return p;
}

static void test_flux_kat()
{
	const IdealGasPhysics<freal> phy(1.4, 0.5, 288.15, 5000.0, 0.72);
	const InviscidFlux<freal> *roe = create_const_inviscidflux<freal>("ROE", &phy);
	const double ul[4] = {1, .9, .1, 2.6}, ur[4] = {1.1, .8, .05, 2.9}, n[2] = {.6, .8};
	const double expect[4] = {0.541628860055636, 1.07319342226683, 0.82808158309487, 1.92662723278903};
	double f[4];
	roe->get_flux(ul, ur, n, f);
	for(int k = 0; k < 4; k++) CHECK(std::fabs(f[k] - expect[k]) < 1e-12*std::fabs(expect[k]), "Roe flux differs from the reference's value");
	delete roe;
	CHECK(create_const_inviscidflux<freal>("NOSUCHFLUX", &phy) == nullptr, "unknown flux key must return nullptr");
}

static void test_wall_bcs(const std::string& dir)
{
	const UMesh<freal,NDIM> m = constructMesh(dir + "/testperiodic.msh");
	const IdealGasPhysics<freal> phy(1.4, 0.5, 288.15, 5000.0, 0.72);
	const std::array<freal,NVARS> uinf = phy.compute_freestream_state(0.0);
	// tests/flow-general/test.ctrl: 4 far field, 2 adiabatic wall, 3 isothermal wall; plus a slip wall on 2 in a second pass
	const double u[4] = {1.0, 0.5, 0.5, 10.0/0.4 + 0.25};
	// the reference's FLUX_TOL is 10*ZERO_TOL = 2.2e-15 for its own non-FMA arithmetic; the device functions contract
	// to FMA and refine SFU reciprocals, which moves an O(20) energy flux by a few more ulps: one extra decade
	const double tol = 10*2.2e-16*10;
	const char *fluxes[] = {"HLLC", "ROE", "AUSM", "AUSMPLUS", "HLL", "LLF"};
	for(int pass = 0; pass < 2; pass++) {
		std::vector<FlowBCConfig> conf = { {4, FARFIELD_BC, {}, {}}, {3, ISOTHERMAL_WALL_BC, {0.0, 290.0/288.15}, {}} };
		if(pass == 0) conf.push_back({2, ADIABATIC_WALL_BC, {0.0}, {}}); else conf.push_back({2, SLIP_WALL_BC, {}, {}});
		auto bcs = create_const_flowBCs<freal>(conf, phy, uinf);
		for(const char *fk : fluxes) {
			const InviscidFlux<freal> *fl = create_const_inviscidflux<freal>(fk, &phy);
			for(fint f = m.gPhyBFaceStart(); f < m.gPhyBFaceEnd(); f++) {
				if(m.gbtags(f,0) != 2) continue;
				const double n[2] = {m.gfacemetric(f,0), m.gfacemetric(f,1)};
				double ug[4], flux[4];
				bcs.at(m.gbtags(f,0))->computeGhostState(u, n, ug);
				fl->get_flux(u, ug, n, flux);
				CHECK(std::fabs(flux[0]) <= tol, "wall mass flux not zero");
				CHECK(std::fabs(flux[3]) <= 10*tol, "wall energy flux not zero");
			}
			delete fl;
		}
		for(auto& kv : bcs) delete kv.second;
	}
	// abc.cpp:493-494: a type without a FlowBC class throws std::runtime_error. PERIODIC_BC is such a type in the reference;
	// here it has a class (its faces become interior faces of the device mesh) whose ghost state must never be asked for
	bool threw = false;
	try { create_const_flowBCs<freal>({{7, (BCType)99, {}, {}}}, phy, uinf); } catch(std::runtime_error&) { threw = true; }
	CHECK(threw, "a BC type without a class must throw std::runtime_error");
	{
		auto pb = create_const_flowBCs<freal>({{7, PERIODIC_BC, {}, {}}}, phy, uinf);
		bool refused = false;
		double g[4]; const double nn[2] = {1.0, 0.0};
		try { pb.at(7)->computeGhostState(u, nn, g); } catch(UnsupportedOptionError&) { refused = true; }
		CHECK(refused, "the periodic marker has no ghost state");
		for(auto& kv : pb) delete kv.second;
	}
}

struct GeomProbe : public Spatial<freal,NVARS> {
	GeomProbe(const UMesh<freal,NDIM> *mesh) : Spatial<freal,NVARS>(mesh) {}
	StatusCode compute_residual(const Vec, Vec, const bool, Vec) const { return 0; }
	void getGradients(const Vec, GradBlock_t<freal,NDIM,NVARS> *const) const {}
	using Spatial<freal,NVARS>::rch; using Spatial<freal,NVARS>::rcbp; using Spatial<freal,NVARS>::gr;
};

static double linearfunc(const double *x, int k) { return (k+1)*(2.0*x[0] + 0.5*x[1] + 2.5); }

static void test_one_exact(const std::string& path)
{
	const UMesh<freal,NDIM> m = constructMesh(path);
	const GeomProbe sp(&m);
	const fint ne = m.gnelem(), nb = m.gnbface(), nf = m.gnaface();
	const GradientScheme<freal,NVARS> *wls = create_const_gradientscheme<freal,NVARS>("LEASTSQUARES", &m, sp.rch.data(), sp.rcbp.data());
	std::vector<GradBlock_t<freal,NDIM,NVARS>> grads(ne);
	MVector<freal> u(ne, NVARS);
	amat::Array2d<freal> ug(nb, NVARS), ul(nf, NVARS), ur(nf, NVARS);
	for(fint i = 0; i < ne; i++) for(int k = 0; k < NVARS; k++) u(i,k) = linearfunc(&sp.rch[(size_t)i*NDIM], k);
	for(fint i = 0; i < nb; i++) for(int k = 0; k < NVARS; k++) ug(i,k) = linearfunc(&sp.rcbp(i,0), k);
	wls->compute_gradients(amat::Array2dView<freal>(u.data(), ne, NVARS), amat::Array2dView<freal>(ug.data(), nb, NVARS), &grads[0](0,0));
	const LinearUnlimitedReconstruction<freal,NVARS> lur(&m, sp.rch.data(), sp.rcbp.data(), sp.gr);
	lur.compute_face_values(u, amat::Array2dView<freal>(ug.data(), nb, NVARS), &grads[0](0,0),
	                        amat::Array2dMutableView<freal>(ul.data(), nf, NVARS), amat::Array2dMutableView<freal>(ur.data(), nf, NVARS));
	double err = 0, lrerr = 0;
	for(fint f = 0; f < nf; f++)
		for(int k = 0; k < NVARS; k++) {
			const double e = (ul(f,k) - linearfunc(&sp.gr(f,0), k))/(k+1);
			err += e*e;
			if(f >= m.gSubDomFaceStart() && f < m.gSubDomFaceEnd()) { const double d = (ul(f,k) - ur(f,k))/(k+1); lrerr += d*d; }
		}
	err = std::sqrt(err/(nf*NVARS)); lrerr = std::sqrt(lrerr/(nf*NVARS));
	std::printf("  1-exact on %s: error norm %.3e, LR error norm %.3e\n", path.c_str(), err, lrerr);
	const double eps = std::numeric_limits<double>::epsilon();
	CHECK(err < 10*eps*8, "WLS + linear reconstruction is not 1-exact");       // values are O(8), the reference's are O(4)
	CHECK(lrerr < 10*eps*8, "left and right face values differ");
	delete wls;
}

static void test_flowfv_and_solver(const std::string& dir)
{
	const UMesh<freal,NDIM> m = constructMesh(dir + "/naca0012luo.msh");
	const FlowPhysicsConfig pc = inviscid_physics();
	FlowNumericsConfig nc; nc.conv_numflux = "ROE"; nc.conv_numflux_jac = "ROE"; nc.gradientscheme = "LEASTSQUARES";
	nc.reconstruction = "VENKATAKRISHNAN"; nc.limiter_param = 2.0; nc.order2 = true;
	const FlowFV_base<freal> *const prob = create_const_flowSpatialDiscretization<freal>(&m, pc, nc);
	const fint ne = m.gnelem();

	Vec u, r, dt, ud, rd, dtd;
	createGhostedSystemVector(&m, NVARS, &u); createSystemVector(&m, NVARS, &r); createSystemVector(&m, 1, &dt);
	createGhostedSystemVector(&m, NVARS, &ud, VEC_DEVICE); createSystemVector(&m, NVARS, &rd, VEC_DEVICE); createSystemVector(&m, 1, &dtd, VEC_DEVICE);
	initializeSystemVector(pc, m, u);
	// perturb so that the residual is not trivially the boundary's
	for(fint i = 0; i < ne; i++) { u->host[4*(size_t)i] *= 1.0 + 0.01*std::sin(0.37*i); u->host[4*(size_t)i+3] *= 1.0 + 0.01*std::cos(0.11*i); }
	VecCopyFromHost(ud, u->host.data());

	CHECK(prob->compute_residual(u, r, true, dt) == 0, "compute_residual (host Vecs) failed");
	CHECK(prob->compute_residual(ud, rd, true, dtd) == 0, "compute_residual (device Vecs) failed");
	std::vector<double> r1(r->host), rdev(r->host.size()), dtdev(ne);
	VecCopyToHost(rd, rdev.data()); VecCopyToHost(dtd, dtdev.data());
	double rmax = 0;
	bool same = true;
	for(size_t k = 0; k < r1.size(); k++) { rmax = std::max(rmax, std::fabs(r1[k])); same = same && r1[k] == rdev[k]; }
	for(fint i = 0; i < ne; i++) same = same && dt->host[i] == dtdev[i] && dt->host[i] > 0;
	CHECK(rmax > 1e-8, "residual is trivially zero");
	CHECK(same, "host-Vec and device-Vec residuals differ");
	// the reference's contract: the residual is ADDED to the vector
	CHECK(prob->compute_residual(u, r, false, nullptr) == 0, "compute_residual without time steps failed");
	bool doubled = true;
	for(size_t k = 0; k < r1.size(); k++) doubled = doubled && std::fabs(r->host[k] - 2.0*r1[k]) <= 4e-16*std::fabs(r1[k]) + 1e-300;
	CHECK(doubled, "compute_residual must add into the residual vector");

	std::vector<GradBlock_t<freal,NDIM,NVARS>> g(ne);
	prob->getGradients(u, g.data());
	MVector<freal> out;
	const auto cd = prob->computeSurfaceData(amat::Array2dView<freal>(u->host.data(), ne, NVARS), g.data(), 2, out);
	CHECK(std::isfinite(std::get<0>(cd)) && std::isfinite(std::get<1>(cd)), "surface data not finite");

	// forward Euler: maxiter reached -> Tolerance_error, history recorded, state changed
	SteadySolverConfig sc; sc.lognres = false; sc.write_final_lin_sys = false; sc.cflinit = 0.5; sc.cflfin = 0.5; sc.rampstart = 0; sc.rampend = 0;
	sc.tol = 1e-12; sc.maxiter = 30; sc.linmaxiterstart = 0; sc.linmaxiterend = 0;
	initializeSystemVector(pc, m, ud);
	initializeSystemVector(pc, m, u);
	SteadyForwardEulerSolver<NVARS> solver(prob, ud, sc);
	bool tol_thrown = false;
	try { solver.solve(ud); } catch(Tolerance_error& e) { tol_thrown = true; }
	CHECK(tol_thrown, "Tolerance_error expected at maxiter");
	const TimingData td = solver.getTimingData();
	CHECK(td.num_timesteps == 30 && td.convhis.size() == 30 && !td.converged, "timing data / history wrong");
	// the same 30 steps through host Vecs end in the same state
	SteadyForwardEulerSolver<NVARS> solver2(prob, u, sc);
	try { solver2.solve(u); } catch(Tolerance_error&) {}
	std::vector<double> uh(u->size());
	VecCopyToHost(ud, uh.data());
	bool same_state = true;
	for(size_t k = 0; k < uh.size(); k++) same_state = same_state && uh[k] == u->host[k];
	CHECK(same_state, "host-Vec and device-Vec solves differ");
	// loose tolerance converges and returns 0
	sc.tol = 2.0; sc.maxiter = 500;      // met after the first step (ratio 1): the loop ends and solve returns 0
	SteadyForwardEulerSolver<NVARS> solver3(prob, ud, sc);
	int rc = -1;
	try { rc = solver3.solve(ud); } catch(...) {}
	CHECK(rc == 0 && solver3.getTimingData().converged, "solver should converge to a loose tolerance");

	VecDestroy(&u); VecDestroy(&r); VecDestroy(&dt); VecDestroy(&ud); VecDestroy(&rd); VecDestroy(&dtd);
	delete prob;

	bool threw = false;
	nc.conv_numflux = "NOSUCHFLUX";
	try { delete create_const_flowSpatialDiscretization<freal>(&m, pc, nc); } catch(UnsupportedOptionError&) { threw = true; }
	CHECK(threw, "unknown flux key must be rejected");
}

int main(int argc, char **argv)
{
	const std::string dir = argc > 1 ? argv[1] : "tests/golden/meshes";
	int ndev = 0;
	if(fvg_device_count(&ndev) != 0 || ndev < 1) { std::printf("no CUDA device: %s\n", fvg_last_error()); return 77; }
	try {
		test_flux_kat();
		test_wall_bcs(dir);
		test_one_exact(dir + "/testperiodic.msh");
		test_one_exact(dir + "/2dcylinderhybrid.msh");
		test_one_exact(dir + "/squareunsquad0.msh");
		test_flowfv_and_solver(dir);
	} catch(std::exception& e) { std::printf("FAIL exception: %s\n", e.what()); return 1; }
	std::printf(nfail ? "HOST_SURFACE FAIL (%d)\n" : "HOST_SURFACE OK\n", nfail);
	return nfail ? 1 : 0;
}
