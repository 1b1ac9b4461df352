/* CPU-only test of the unsteady driver of the host class surface (TVDRKSolver, fvens_b200/host/fvens_b200.hpp) on a
 * Spatial that is NOT the engine's: the generic loop drives compute_residual on host Vecs, so no GPU is touched.
 *     test_ode_host <mesh file> <scratch log file>
 * The model problem is du/dt = -lambda_i u per cell (residual = area * lambda_i * u, local time steps given), whose
 * exact solution is known: the observed order of accuracy of orders 1, 2, 3 must be 1, 2, 3 - which it is only if the
 * stages are evaluated at the stage state and the update has the right sign (the reference's loop, ode/aodesolver.cpp:
 * 708-742, does neither). Also: dt = cfl * min(dtm) from the first stage, the loop condition and step count, the
 * divergence check, the log-file line, the coefficient table of initialize_TVDRK_Coeffs (:45-67).
 * Run by tests/test_ode_host.py.
 */
#include "../../fvens_b200/host/fvens_b200.hpp"
#include "../../fvens_b200/host/controlparser.hpp"
#include "../../fvens_b200/host/casesolvers.hpp"
#include <cstdio>

using namespace fvens;

static int nfail = 0;
#define CHECK(cond, msg) do { if(!(cond)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, msg); nfail++; } } while(0)

class DecaySpatial : public Spatial<freal,NVARS> {
public:
	DecaySpatial(const UMesh<freal,NDIM> *const mesh, const double dt_small, const bool poison = false)
		: Spatial<freal,NVARS>(mesh), dts(dt_small), bad(poison), nevals(0) {}
	double lambda(const fint i) const { return 1.0 + 0.5*std::sin(0.37*i); }
	StatusCode compute_residual(const Vec u, Vec residual, const bool gettimesteps, Vec dtm) const {
		nevals++;
		for(fint i = 0; i < m->gnelem(); i++) {
			for(int k = 0; k < NVARS; k++)
				residual->host[(size_t)i*NVARS+k] += -lambda(i)*(k+1)*u->host[(size_t)i*NVARS+k]*m->garea(i);
			if(gettimesteps) dtm->host[i] = (i == 7 ? dts : 3.0*dts + 1e-3*i);
		}
		if(bad && gettimesteps && nevals > 4) dtm->host[3] = std::numeric_limits<double>::quiet_NaN();
		return 0;
	}
	void getGradients(const Vec, GradBlock_t<freal,NDIM,NVARS> *const) const {}
	const double dts; const bool bad;
	mutable int nevals;
};

static double run(const UMesh<freal,NDIM>& m, const int order, const double dts, int *steps_out = nullptr, double *time_out = nullptr)
{
	DecaySpatial sp(&m, dts);
	Vec u = nullptr;
	createSystemVector(&m, NVARS, &u);
	for(size_t j = 0; j < u->host.size(); j++) u->host[j] = 1.0 + 0.1*(j % 5);
	const std::vector<double> u0 = u->host;
	TVDRKSolver<NVARS> time(&sp, u, order, "", 0.5);
	time.solve(1.0);
	const double T = time.physicalTime();
	double err = 0;
	for(fint i = 0; i < m.gnelem(); i++)
		for(int k = 0; k < NVARS; k++) {
			const size_t j = (size_t)i*NVARS + k;
			err = std::max(err, std::fabs(u->host[j] - u0[j]*std::exp(-sp.lambda(i)*(k+1)*T)));
		}
	if(steps_out) *steps_out = time.numSteps();
	if(time_out) *time_out = T;
	CHECK(sp.nevals == order*time.numSteps(), "one residual evaluation per stage");
	VecDestroy(&u);
	return err;
}

int main(int argc, char **argv)
{
	if(argc < 3) { std::printf("usage: test_ode_host <mesh file> <log file>\n"); return 2; }
	UMesh<freal,NDIM> m = constructMesh(argv[1]);
	CHECK(m.gnelem() > 8, "mesh too small for the test");

	// coefficient table
	double c[9];
	CHECK(fvg_tvdrk_coefficients(3, c) == 0 && c[0] == 1 && c[1] == 0 && c[2] == 1 && c[3] == 0.75 && c[4] == 0.25 && c[5] == 0.25
	      && c[6] == 0.3333333333333333 && c[7] == 0.6666666666666667 && c[8] == 0.6666666666666667, "order-3 table");
	CHECK(fvg_tvdrk_coefficients(2, c) == 0 && c[3] == 0.5 && c[4] == 0.5 && c[5] == 0.5, "order-2 table");
	CHECK(fvg_tvdrk_coefficients(4, c) != 0, "order 4 is not available");

	// time step = cfl * smallest local step, from the first stage; loop runs while time <= T - 1e-12, last step not clipped
	int steps = 0; double T = 0;
	run(m, 2, 0.02, &steps, &T);
	CHECK(steps == 100 && std::fabs(T - 1.0) < 1e-9, "100 steps of 0.5*0.02 expected");
	run(m, 1, 0.03, &steps, &T);
	CHECK(steps == 67 && T > 1.0 && T < 1.0 + 0.015 + 1e-9, "67 steps of 0.015 expected (the last one overshoots, as the reference)");

	// observed order of accuracy
	for(int order = 1; order <= 3; order++) {
		const double e1 = run(m, order, 0.04), e2 = run(m, order, 0.02), e3 = run(m, order, 0.01);
		const double p12 = std::log2(e1/e2), p23 = std::log2(e2/e3);
		std::printf("order %d: errors %.3e %.3e %.3e, observed %.3f %.3f\n", order, e1, e2, e3, p12, p23);
		CHECK(std::fabs(p23 - order) < 0.08 && std::fabs(p12 - order) < 0.15, "observed order of accuracy");
	}

	// divergence: a NaN local time step is seen by the min and raised as Numerical_error; u keeps the last good step
	{
		DecaySpatial sp(&m, 0.02, true);
		Vec u = nullptr; createSystemVector(&m, NVARS, &u);
		VecSet(u, 1.0);
		TVDRKSolver<NVARS> time(&sp, u, 2, "", 0.5);
		bool threw = false;
		try { time.solve(1.0); } catch(Numerical_error& e) { threw = std::string(e.what()).find("dtmin is Nan or inf") != std::string::npos; }
		CHECK(threw, "Numerical_error expected");
		CHECK(time.numSteps() == 2 && std::isfinite(u->host[0]) && u->host[0] < 1.0, "state of the last completed step");
		VecDestroy(&u);
	}
	// unsupported order, log-file line
	{
		DecaySpatial sp(&m, 0.05);
		Vec u = nullptr; createSystemVector(&m, NVARS, &u); VecSet(u, 1.0);
		TVDRKSolver<NVARS> bad(&sp, u, 4, "", 0.5);
		bool threw = false;
		try { bad.solve(1.0); } catch(UnsupportedOptionError&) { threw = true; }
		CHECK(threw, "order 4 must be rejected");
		{ std::ofstream f(argv[2]); f << "case"; }
		TVDRKSolver<NVARS> time(&sp, u, 3, argv[2], 0.5);
		CHECK(time.solve(0.2) == 0, "solve");
		std::ifstream f(argv[2]); std::string line; std::getline(f, line);
		int tabs = 0; for(char ch : line) tabs += ch == '\t';
		CHECK(line.compare(0, 5, "case\t") == 0 && tabs == 3, "log line: <tab>threads<tab>wall<tab>cpu appended");
		const std::tuple<double,double> rt = time.getRunTimes();
		CHECK(std::get<0>(rt) >= 0 && std::get<1>(rt) >= 0, "run times");
		VecDestroy(&u);
	}
	// UnsteadyFlowCase::execute (utilities/casesolvers.cpp:429-445): TVDRK with the control file's order, CFL and final time
	{
		DecaySpatial sp(&m, 0.02);
		Vec u = nullptr; createSystemVector(&m, NVARS, &u); VecSet(u, 1.0);
		FlowParserOptions o;
		o.time_integrator = "TVDRK"; o.time_order = 2; o.phy_cfl = 0.5; o.final_time = 0.1; o.logfile = "";
		CHECK(UnsteadyFlowCase(o).execute(&sp, u) == 0 && sp.nevals == 20 && u->host[0] < 1.0, "UnsteadyFlowCase: 10 steps of order 2");
		o.time_integrator = "BDF";
		bool threw = false;
		try { UnsteadyFlowCase(o).execute(&sp, u); } catch(UnsupportedOptionError&) { threw = true; }
		CHECK(threw, "only TVDRK exists");
		VecDestroy(&u);
	}
	std::printf(nfail ? "FAILED (%d)\n" : "ALL PASSED\n", nfail);
	return nfail ? 1 : 0;
}
