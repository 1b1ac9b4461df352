/* C++ test of the multi-GPU class surface (fvens_b200/host/fvens_b200.hpp: DistributedFlowFV + SteadyForwardEulerSolver).
 * One process per rank; every rank runs
 *     test_dist_surface <mesh file> <rank> <nranks> <scratch dir> [device]
 * builds its subdomain of the mesh (Hilbert-curve partition), steps 60 forward-Euler iterations with the reference's
 * driver class on the distributed FlowFV, and writes its own rows of the final state plus the residual history to the
 * scratch directory; tests/test_gpu_cpp_surface.py merges them and compares with the single-GPU run (bitwise state).
 * The set-up all-gather is done through files in the scratch directory - the transport is the caller's business
 * (an MPI program passes MPI_Allgather), nothing of the data path goes through it.
 */
#include "../../fvens_b200/host/fvens_b200.hpp"
#include <chrono>
#include <cstdio>
#include <thread>

using namespace fvens;

int main(int argc, char **argv)
{
	if(argc < 5) { std::printf("usage: test_dist_surface <mesh> <rank> <nranks> <scratch dir> [device]\n"); return 2; }
	const std::string meshfile = argv[1], dir = argv[4];
	const int rank = std::atoi(argv[2]), nranks = std::atoi(argv[3]), device = argc > 5 ? std::atoi(argv[5]) : 0;
	try {
		const UMesh<freal,NDIM> m = constructMesh(meshfile);
		FlowPhysicsConfig p;
		p.gamma = 1.4; p.Minf = 0.8; p.Tinf = 288.15; p.Reinf = 5000.0; p.Pr = 0.72; p.aoa = 1.25*M_PI/180.0;
		p.viscous_sim = false; p.const_visc = false;
		p.bcconf = { {2, SLIP_WALL_BC, {}, {}}, {4, FARFIELD_BC, {}, {}} };
		FlowNumericsConfig n;
		n.conv_numflux = "ROE"; n.conv_numflux_jac = "ROE"; n.gradientscheme = "LEASTSQUARES"; n.reconstruction = "VENKATAKRISHNAN";
		n.limiter_param = 2.0; n.order2 = true;

		// cell -> rank map: the Hilbert order cut into equal chunks (a Scotch map drops in the same way)
		std::vector<int> part((size_t)m.gnelem());
		{
			fvg_umesh *um = nullptr;
			fvg_throw(fvg_umesh_read(meshfile.c_str(), &um), "fvg_umesh_read");
			fvg_throw(fvg_partition_sfc(um, nranks, part.data()), "fvg_partition_sfc");
			fvg_umesh_destroy(um);
		}
		const AllGather allgather = [&](const void *mine, void *all, size_t bytes) {
			{
				const std::string tmp = dir + "/gather_" + std::to_string(rank) + ".tmp", fin = dir + "/gather_" + std::to_string(rank) + ".bin";
				std::FILE *f = std::fopen(tmp.c_str(), "wb");
				if(!f || std::fwrite(mine, 1, bytes, f) != bytes) throw std::runtime_error("cannot write " + tmp);
				std::fclose(f);
				std::rename(tmp.c_str(), fin.c_str());
			}
			for(int r = 0; r < nranks; r++) {
				const std::string fin = dir + "/gather_" + std::to_string(r) + ".bin";
				for(int tries = 0; ; tries++) {
					std::FILE *f = std::fopen(fin.c_str(), "rb");
					if(f) {
						const size_t got = std::fread(static_cast<unsigned char*>(all) + bytes*(size_t)r, 1, bytes, f);
						std::fclose(f);
						if(got == bytes) break;
					}
					if(tries > 3000) throw std::runtime_error("rank " + std::to_string(r) + " never arrived");
					std::this_thread::sleep_for(std::chrono::milliseconds(10));
				}
			}
		};
		DistributedFlowFV<freal,true,false> flow(&m, p, n, part, rank, nranks, device, allgather);
		std::printf("rank %d: %d own cells, %d ghost cells\n", rank, (int)flow.ncell(), (int)flow.nghostcell());

		// free stream everywhere, as SteadyFlowCase::execute_starter does (casesolvers.cpp:100-140)
		const std::array<freal,NVARS> uinf = flow.freestream();
		std::vector<double> u0((size_t)flow.ncell()*NVARS);
		for(fint i = 0; i < flow.ncell(); i++) for(int k = 0; k < NVARS; k++) u0[(size_t)i*NVARS+k] = uinf[k];
		Vec u = nullptr;
		VecCreateBlocked(flow.ncell(), flow.nghostcell(), NVARS, VEC_DEVICE, &u);
		VecCopyFromHost(u, u0.data());

		SteadySolverConfig sc;
		sc.lognres = false; sc.logfile = ""; sc.write_final_lin_sys = false;
		sc.cflinit = 0.4; sc.cflfin = 0.4; sc.rampstart = 0; sc.rampend = 0; sc.tol = 1e-30; sc.maxiter = 60;
		sc.linmaxiterstart = 0; sc.linmaxiterend = 0;
		SteadyForwardEulerSolver<NVARS> solver(&flow, u, sc);
		bool threw = false;
		try { solver.solve(u); } catch(Tolerance_error&) { threw = true; }      // maxiter reached, as the reference reports it
		flow.check_neighbours();
		const TimingData td = solver.getTimingData();

		// residual of the final state through the Spatial interface (adds into a zeroed Vec)
		Vec r = nullptr, dt = nullptr;
		VecCreateBlocked(flow.ncell(), 0, NVARS, VEC_DEVICE, &r);
		VecCreateBlocked(flow.ncell(), 0, 1, VEC_DEVICE, &dt);
		VecSet(r, 0.0);
		if(flow.compute_residual(u, r, true, dt) != 0) throw std::runtime_error(fvg_last_error());

		std::vector<double> uh((size_t)u->size()), rh((size_t)flow.ncell()*NVARS);
		VecCopyToHost(u, uh.data()); VecCopyToHost(r, rh.data());
		std::FILE *f = std::fopen((dir + "/state_" + std::to_string(rank) + ".bin").c_str(), "wb");
		const int nown = (int)flow.ncell(), nst = (int)td.convhis.size(), flag = threw ? 1 : 0;
		std::fwrite(&nown, sizeof(int), 1, f); std::fwrite(&nst, sizeof(int), 1, f); std::fwrite(&flag, sizeof(int), 1, f);
		for(int i = 0; i < nown; i++) { const int g = (int)flow.global_cell(i); std::fwrite(&g, sizeof(int), 1, f); }
		std::fwrite(uh.data(), sizeof(double), (size_t)nown*NVARS, f);
		std::fwrite(rh.data(), sizeof(double), (size_t)nown*NVARS, f);
		for(int s = 0; s < nst; s++) { const double h = td.convhis[(size_t)s].absrmsres; std::fwrite(&h, sizeof(double), 1, f); }
		std::fclose(f);
		VecDestroy(&u); VecDestroy(&r); VecDestroy(&dt);
		std::printf("DIST_SURFACE OK rank %d steps %d\n", rank, nst);
	}
	catch(std::exception& e) { std::printf("DIST_SURFACE FAIL rank %d: %s\n", rank, e.what()); return 1; }
	return 0;
}
