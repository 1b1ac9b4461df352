/* ORACLE — TEST INFRASTRUCTURE ONLY (see orc_physics.hpp header).
 *
 * CPU restatement of FVENS's finite-volume residual path: Spatial geometry, gradient schemes,
 * reconstructions, FlowFV::compute_residual / compute_fluxes / compute_max_timestep, the steady
 * forward-Euler driver and the surface functionals. The loop structure of the reference is KEPT
 * (separate passes P1..P11, face loops with `omp atomic`, per-call scratch allocation) because
 * this code is also the CPU baseline timed by bench.py.
 *
 * PINNED: tests/test_oracle_ref_b.py and tests/test_oracle_ref_c.py hold this restatement against the reference's
 * own object code for the same loops (oracle/ref_tier_b.cpp, ref_tier_c.cpp: the reference's sources compiled in
 * place) to 1e-12; tests/test_oracle_kat.py against the reference's known-answer tests.
 *
 * Deliberate deviations from the reference, documented in SURVEY.md section 0 and DESIGN.md section 2:
 *  H1: Barth-Jespersen / Venkatakrishnan index the cell-state array with boundary ghost ids
 *      (out of bounds in the reference). Here bnd_policy 0 uses the boundary ghost state `ug`,
 *      bnd_policy 1 skips boundary neighbours (as the reference's WENO does).
 *  H2: limiter_param is an explicit input.
 *  The Green-Gauss boundary-face update is atomic (the reference's races under OpenMP).
 */
#ifndef ORC_SPATIAL_HPP
#define ORC_SPATIAL_HPP

#include "orc_physics.hpp"
#include "orc_mesh.hpp"
#include <map>
#include <memory>
#include <cstring>

namespace orc {

enum GradId { GRAD_ZERO = 0, GRAD_GG = 1, GRAD_WLS = 2 };
enum ReconId { RECON_NONE = 0, RECON_WENO = 1, RECON_VANALBADA = 2, RECON_BJ = 3, RECON_VENKAT = 4 };

struct Numerics {
	int flux = FLUX_ROE;
	int gradient = GRAD_WLS;
	int recon = RECON_NONE;
	double limiter_param = 1.0;
	int order2 = 1;
	int bnd_policy = 0;
};

struct PhysConf {
	double gamma = 1.4, Minf = 0.5, Tinf = 298.0, Reinf = 1e300, Pr = 0.72, aoa = 0.0;
	int viscous = 0, const_visc = 0;
};

/// Spatial + FlowFV_base + FlowFV rolled together (spatial/aspatial.cpp, spatial/flow_spatial.cpp)
struct Flow {
	const Mesh *m;
	PhysConf pc;
	Numerics nc;
	Gas physics;
	double uinf[4];
	std::map<int,BC> bcs;

	std::vector<double> rc;    ///< cell centres [nelem][2]
	std::vector<double> gr;    ///< face midpoints [naface][2]
	std::vector<double> rcbp;  ///< ghost cell centres [nbface][2]
	std::vector<double> V;     ///< WLS inverse matrices [nelem][4] = (0,0),(0,1),(1,0),(1,1)
	std::vector<double> clength; ///< Venkatakrishnan length scale

	Flow(const Mesh *mesh, const PhysConf &p, const Numerics &n, const std::vector<BC> &bclist)
		: m(mesh), pc(p), nc(n), physics(p.gamma, p.Minf, p.Tinf, p.Reinf, p.Pr)
	{
		physics.freestreamState(pc.aoa, uinf);
		for(const BC &b : bclist) bcs[b.tag] = b;
		setup_geometry();
		if(nc.gradient == GRAD_WLS) setup_wls();
		if(nc.recon == RECON_VENKAT) setup_venkat();
	}

	/// aspatial.cpp:37-76 (face midpoints), mesh.cpp:317-328 (centres), aspatial.cpp:98-119 (ghosts)
	void setup_geometry()
	{
		rc.resize(2*(size_t)m->nelem);
		for(int i = 0; i < m->nelem; i++)
			for(int d = 0; d < 2; d++) {
				double c = 0;
				for(int j = 0; j < m->nnode[i]; j++) c += m->x(m->inpoel[4*i+j], d);
				c /= (double)m->nnode[i];
				rc[2*i+d] = c;
			}
		gr.assign(2*(size_t)m->naface, 0.0);
		for(int f = 0; f < m->naface; f++)
			for(int d = 0; d < 2; d++) {
				double g = 0;
				for(int iv = 0; iv < 2; iv++) g += m->x(m->intfac[4*f+2+iv], d);
				g /= 2;
				gr[2*f+d] = g;
			}
		rcbp.resize(2*(size_t)m->nbface);
		for(int f = 0; f < m->nbface; f++) {
			const int ielem = m->intfac[4*f];
			for(int d = 0; d < 2; d++) {
				double mid = 0;
				for(int k = 0; k < 2; k++) mid += m->x(m->intfac[4*f+2+k], d);
				mid /= 2;
				rcbp[2*f+d] = 2.0*mid - rc[2*ielem+d];
			}
		}
	}

	/// agradientschemes.cpp:219-317; 2x2 inverse as Eigen's fixed-size path: adj * (1/det)
	void setup_wls()
	{
		std::vector<double> A(4*(size_t)m->nelem, 0.0);
		for(int f = 0; f < m->nbface; f++) {
			const int ie = m->intfac[4*f];
			double w2 = 0, dr[2];
			for(int d = 0; d < 2; d++) {
				w2 += (rc[2*ie+d]-rcbp[2*f+d])*(rc[2*ie+d]-rcbp[2*f+d]);
				dr[d] = rc[2*ie+d]-rcbp[2*f+d];
			}
			w2 = 1.0/(w2);
			for(int i = 0; i < 2; i++) for(int j = 0; j < 2; j++) A[4*ie+2*i+j] += w2*dr[i]*dr[j];
		}
		for(int f = m->nbface; f < m->naface; f++) {
			const int ie = m->intfac[4*f], je = m->intfac[4*f+1];
			double w2 = 0, dr[2];
			for(int d = 0; d < 2; d++) {
				w2 += (rc[2*ie+d]-rc[2*je+d])*(rc[2*ie+d]-rc[2*je+d]);
				dr[d] = rc[2*ie+d]-rc[2*je+d];
			}
			w2 = 1.0/(w2);
			for(int i = 0; i < 2; i++) for(int j = 0; j < 2; j++) {
				A[4*ie+2*i+j] += w2*dr[i]*dr[j];
				A[4*je+2*i+j] += w2*dr[i]*dr[j];
			}
		}
		V.resize(4*(size_t)m->nelem);
		for(int i = 0; i < m->nelem; i++) {
			const double a = A[4*i], b = A[4*i+1], c = A[4*i+2], d = A[4*i+3];
			const double invdet = 1.0/(a*d - c*b);
			V[4*i+0] =  d*invdet;
			V[4*i+2] = -c*invdet;
			V[4*i+1] = -b*invdet;
			V[4*i+3] =  a*invdet;
		}
	}

	/// limitedlinearreconstruction.cpp:189-205
	void setup_venkat()
	{
		clength.assign(m->nelem, 0.0);
		for(int iel = 0; iel < m->nelem; iel++) {
			const int nn = m->nnode[iel];
			for(int ifa = 0; ifa < nn; ifa++) {
				double llen = 0;
				const int a = m->inpoel[4*iel+ifa], b = m->inpoel[4*iel+(ifa+1)%nn];
				for(int d = 0; d < 2; d++)
					llen += std::pow(m->x(a,d) - m->x(b,d), 2);
				if(clength[iel] < llen) clength[iel] = llen;
			}
			clength[iel] = std::sqrt(clength[iel]);
		}
	}

	/// flow_spatial.cpp:85-93 (bcs.at throws std::out_of_range for an unregistered tag)
	void boundary_state(int ied, const double *ins, double *gs) const
	{
		ghost_state(bcs.at(m->btags[ied]), physics, uinf, ins, &m->facemetric[3*ied], gs);
	}
	/// flow_spatial.cpp:74-83
	void boundary_states(const double *ins, double *gs) const
	{
#pragma omp parallel for default(shared)
		for(int ied = 0; ied < m->nbface; ied++)
			boundary_state(ied, ins + ied*4, gs + ied*4);
	}

	// ---------------------------------------------------------------- gradients
	/// grad memory layout = GradBlock_t col-major 2x4: grad[8*cell + idim + 2*ivar]

	/// agradientschemes.cpp:62-214 (no connectivity faces)
	void gradients_gg(const double *u, const double *ug, double *grad) const
	{
#pragma omp parallel default(shared)
		{
#pragma omp for
			for(int i = 0; i < m->nelem*8; i++) grad[i] = 0;

#pragma omp for
			for(int f = 0; f < m->nbface; f++) {
				const int ie = m->intfac[4*f];
				double mid[2] = {0,0};
				for(int k = 2; k < 4; k++)
					for(int d = 0; d < 2; d++) mid[d] += m->x(m->intfac[4*f+k], d);
				for(int d = 0; d < 2; d++) mid[d] /= 2;
				double dL = 0, dR = 0;
				for(int d = 0; d < 2; d++) {
					dL += (mid[d]-rc[2*ie+d])*(mid[d]-rc[2*ie+d]);
					dR += (mid[d]-rcbp[2*f+d])*(mid[d]-rcbp[2*f+d]);
				}
				dL = 1.0/std::sqrt(dL);
				dR = 1.0/std::sqrt(dR);
				const double areainv1 = 1.0/m->area[ie];
				for(int iv = 0; iv < 4; iv++) {
					const double ut = (u[4*ie+iv]*dL + ug[4*f+iv]*dR)/(dL+dR) * m->facemetric[3*f+2];
					// The reference updates grad[ielem] WITHOUT `omp atomic` in this loop (agradientschemes.cpp:109-110):
					// a cell with two boundary faces (domain corners) races when the faces fall to different threads,
					// and its gradient - then the residual of the cells around it - comes out wrong now and then. The
					// oracle states the intended (sequential) arithmetic, so the update is atomic here.
					for(int d = 0; d < 2; d++) {
#pragma omp atomic update
						grad[8*ie+d+2*iv] += (ut * m->facemetric[3*f+d])*areainv1;
					}
				}
			}

#pragma omp for
			for(int f = m->nbface; f < m->naface; f++) {
				const int ie = m->intfac[4*f], je = m->intfac[4*f+1];
				double mid[2] = {0,0};
				for(int k = 2; k < 4; k++)
					for(int d = 0; d < 2; d++) mid[d] += m->x(m->intfac[4*f+k], d);
				for(int d = 0; d < 2; d++) mid[d] /= 2;
				double dL = 0, dR = 0;
				for(int d = 0; d < 2; d++) {
					dL += (mid[d]-rc[2*ie+d])*(mid[d]-rc[2*ie+d]);
					dR += (mid[d]-rc[2*je+d])*(mid[d]-rc[2*je+d]);
				}
				dL = 1.0/std::sqrt(dL);
				dR = 1.0/std::sqrt(dR);
				const double areainv1 = 1.0/m->area[ie];
				const double areainv2 = 1.0/m->area[je];
				for(int iv = 0; iv < 4; iv++) {
					const double ut = (u[4*ie+iv]*dL + u[4*je+iv]*dR)/(dL+dR) * m->facemetric[3*f+2];
					for(int d = 0; d < 2; d++) {
#pragma omp atomic update
						grad[8*ie+d+2*iv] += (ut * m->facemetric[3*f+d])*areainv1;
#pragma omp atomic update
						grad[8*je+d+2*iv] -= (ut * m->facemetric[3*f+d])*areainv2;
					}
				}
			}
		}
	}

	/// agradientschemes.cpp:323-440
	void gradients_wls(const double *u, const double *ug, double *grad) const
	{
		std::vector<double> fv(8*(size_t)m->nelem);    // per-call allocation, as the reference (H10)
		double *const f = fv.data();
#pragma omp parallel for default(shared)
		for(int i = 0; i < m->nelem*8; i++) f[i] = 0;

#pragma omp parallel for default(shared)
		for(int fc = 0; fc < m->nbface; fc++) {
			const int ie = m->intfac[4*fc];
			double w2 = 0, dr[2], du[4];
			for(int d = 0; d < 2; d++) {
				w2 += (rc[2*ie+d]-rcbp[2*fc+d])*(rc[2*ie+d]-rcbp[2*fc+d]);
				dr[d] = rc[2*ie+d]-rcbp[2*fc+d];
			}
			w2 = 1.0/(w2);
			for(int iv = 0; iv < 4; iv++) du[iv] = u[4*ie+iv] - ug[4*fc+iv];
			for(int iv = 0; iv < 4; iv++)
				for(int d = 0; d < 2; d++) {
#pragma omp atomic update
					f[8*ie+d+2*iv] += w2*dr[d]*du[iv];
				}
		}

#pragma omp parallel for default(shared)
		for(int fc = m->nbface; fc < m->naface; fc++) {
			const int ie = m->intfac[4*fc], je = m->intfac[4*fc+1];
			double w2 = 0, dr[2], du[4];
			for(int d = 0; d < 2; d++) {
				w2 += (rc[2*ie+d]-rc[2*je+d])*(rc[2*ie+d]-rc[2*je+d]);
				dr[d] = rc[2*ie+d]-rc[2*je+d];
			}
			w2 = 1.0/(w2);
			for(int iv = 0; iv < 4; iv++) du[iv] = u[4*ie+iv] - u[4*je+iv];
			for(int iv = 0; iv < 4; iv++)
				for(int d = 0; d < 2; d++) {
#pragma omp atomic update
					f[8*ie+d+2*iv] += w2*dr[d]*du[iv];
#pragma omp atomic update
					f[8*je+d+2*iv] += w2*dr[d]*du[iv];
				}
		}

#pragma omp parallel for default(shared)
		for(int ie = 0; ie < m->nelem; ie++)
			for(int iv = 0; iv < 4; iv++)
				for(int d = 0; d < 2; d++)
					grad[8*ie+d+2*iv] = V[4*ie+2*d]*f[8*ie+0+2*iv] + V[4*ie+2*d+1]*f[8*ie+1+2*iv];
	}

	/// agradientschemes.cpp:36-51, and the dispatch of afactory.cpp:111-127
	void gradients(const double *u, const double *ug, double *grad) const
	{
		if(nc.gradient == GRAD_WLS) gradients_wls(u, ug, grad);
		else if(nc.gradient == GRAD_GG) gradients_gg(u, ug, grad);
		else {
#pragma omp parallel for default(shared)
			for(int i = 0; i < m->nelem*8; i++) grad[i] = 0;
		}
	}

	// ---------------------------------------------------------------- reconstruction

	/// reconstruction_utils.hpp:17-32
	static double extrapolate(double ucell, const double *g /*cell block*/, int ivar, double lim,
	                          const double *gp, const double *rcc)
	{
		double uface = ucell;
		for(int d = 0; d < 2; d++)
			uface += lim*g[d+2*ivar]*(gp[d] - rcc[d]);
		return uface;
	}

	/// areconstruction.cpp:52-103
	void recon_linear(const double *u, const double *grad, double *ufl, double *ufr) const
	{
#pragma omp parallel default(shared)
		{
#pragma omp for nowait
			for(int f = m->nbface; f < m->naface; f++) {
				const int ie = m->intfac[4*f], je = m->intfac[4*f+1];
				for(int i = 0; i < 4; i++) {
					ufl[4*f+i] = extrapolate(u[4*ie+i], grad+8*ie, i, 1.0, &gr[2*f], &rc[2*ie]);
					ufr[4*f+i] = extrapolate(u[4*je+i], grad+8*je, i, 1.0, &gr[2*f], &rc[2*je]);
				}
			}
#pragma omp for
			for(int f = 0; f < m->nbface; f++) {
				const int ie = m->intfac[4*f];
				for(int i = 0; i < 4; i++)
					ufl[4*f+i] = extrapolate(u[4*ie+i], grad+8*ie, i, 1.0, &gr[2*f], &rc[2*ie]);
			}
		}
	}

	/// musclreconstruction.cpp:35-59 helpers, 70-130 loops (eps = 1e-8, k = 1/3)
	void recon_muscl(const double *u, const double *ug, const double *grad, double *ufl, double *ufr) const
	{
		const double eps = 1e-8, k = 1.0/3.0;
		auto biased = [](const double *ri, const double *rj, double ui, double uj, const double *g) {
			double del = 0;
			for(int d = 0; d < 2; d++) del += g[d]*(rj[d]-ri[d]);
			return 2.0*del - (uj-ui);
		};
#pragma omp parallel for default(shared)
		for(int f = 0; f < m->nbface; f++) {
			const int ie = m->intfac[4*f];
			for(int i = 0; i < 4; i++) {
				const double ui = u[4*ie+i], uj = ug[4*f+i];
				const double deltam = biased(&rc[2*ie], &rcbp[2*f], ui, uj, grad+8*ie+2*i);
				double phi_l = (2.0*deltam * (uj - ui) + eps) / (deltam*deltam + (uj - ui)*(uj - ui) + eps);
				if(phi_l < 0.0) phi_l = 0.0;
				ufl[4*f+i] = ui + phi_l/4.0*( (1.0-k*phi_l)*deltam + (1.0+k*phi_l)*(uj - ui) );
			}
		}
#pragma omp parallel for default(shared)
		for(int f = m->nbface; f < m->naface; f++) {
			const int ie = m->intfac[4*f], je = m->intfac[4*f+1];
			for(int i = 0; i < 4; i++) {
				const double ui = u[4*ie+i], uj = u[4*je+i];
				const double deltam = biased(&rc[2*ie], &rc[2*je], ui, uj, grad+8*ie+2*i);
				const double deltap = biased(&rc[2*ie], &rc[2*je], ui, uj, grad+8*je+2*i);
				double phi_l = (2.0*deltam * (uj - ui) + eps) / (deltam*deltam + (uj - ui)*(uj - ui) + eps);
				if(phi_l < 0.0) phi_l = 0.0;
				double phi_r = (2*deltap * (uj - ui) + eps) / (deltap*deltap + (uj - ui)*(uj - ui) + eps);
				if(phi_r < 0.0) phi_r = 0.0;
				ufl[4*f+i] = ui + phi_l/4.0*( (1.0-k*phi_l)*deltam + (1.0+k*phi_l)*(uj - ui) );
				ufr[4*f+i] = uj - phi_r/4.0*( (1.0-k*phi_r)*deltap + (1.0+k*phi_r)*(uj - ui) );
			}
		}
	}

	/// limitedlinearreconstruction.cpp:39-105 (gamma = 4, epsilon = 1e-5, lambda = limiter_param)
	void recon_weno(const double *u, const double *grad, double *ufl, double *ufr) const
	{
		const double gamma = 4.0, lambda = nc.limiter_param, epsilon = 1.0e-5;
		auto mag2 = [](const double *g, int iv) {
			double r = 0;
			for(int d = 0; d < 2; d++) r += g[d+2*iv]*g[d+2*iv];
			return r;
		};
#pragma omp parallel for default(shared)
		for(int ie = 0; ie < m->nelem; ie++)
			for(int iv = 0; iv < 4; iv++) {
				double wsum = 0, lgrad[2] = {0,0};
				{
					const double denom = std::pow( mag2(grad+8*ie,iv) + epsilon, gamma );
					const double w = lambda / denom;
					wsum += w;
					for(int d = 0; d < 2; d++) lgrad[d] += w*grad[8*ie+d+2*iv];
				}
				for(int j = 0; j < m->nnode[ie]; j++) {
					const int je = m->esuel[4*ie+j];
					if(je >= m->nelem) continue;
					const double denom = std::pow( mag2(grad+8*je,iv) + epsilon, gamma );
					const double w = 1.0 / denom;
					wsum += w;
					for(int d = 0; d < 2; d++) lgrad[d] += w*grad[8*je+d+2*iv];
				}
				for(int d = 0; d < 2; d++) lgrad[d] /= wsum;

				for(int j = 0; j < m->nnode[ie]; j++) {
					const int face = m->elemface[4*ie+j];
					const int je = m->esuel[4*ie+j];
					double val = u[4*ie+iv];
					for(int d = 0; d < 2; d++)
						val += lgrad[d]*(gr[2*face+d] - rc[2*ie+d]);
					if(ie < je) ufl[4*face+iv] = val;
					else        ufr[4*face+iv] = val;
				}
			}
	}

	/// limitedlinearreconstruction.cpp:117-176 (BJ) and 209-268 (Venkatakrishnan), with the H1 patch
	void recon_limited(bool venkat, const double *u, const double *ug, const double *grad,
	                   double *ufl, double *ufr) const
	{
#pragma omp parallel for default(shared)
		for(int iel = 0; iel < m->nelem; iel++) {
			const double eps2 = venkat ? std::pow(nc.limiter_param*clength[iel], 3) : 0.0;
			for(int iv = 0; iv < 4; iv++) {
				double duimin = 0, duimax = 0;
				for(int j = 0; j < m->nnode[iel]; j++) {
					const int jel = m->esuel[4*iel+j];
					double dui;
					if(jel < m->nelem) dui = u[4*jel+iv]-u[4*iel+iv];
					else if(nc.bnd_policy == 0) dui = ug[4*(jel-m->nelem)+iv]-u[4*iel+iv];   // H1
					else continue;                                                          // H1 alt
					if(dui > duimax) duimax = dui;
					if(dui < duimin) duimin = dui;
				}
				double lim = 1.0;
				for(int j = 0; j < m->nnode[iel]; j++) {
					const int face = m->elemface[4*iel+j];
					const double uface = extrapolate(u[4*iel+iv], grad+8*iel, iv, 1.0, &gr[2*face], &rc[2*iel]);
					double phiik;
					if(venkat) {
						const double dm = uface - u[4*iel+iv];
						const double dp = dm < 0 ? duimin : duimax;
						phiik = (dp*dp + 2*dp*dm + eps2)/(dp*dp + dp*dm + 2*dm*dm + eps2);
					} else {
						const double diff = uface - u[4*iel+iv];
						if(diff > 0)      phiik = 1 < duimax/diff ? 1 : duimax/diff;
						else if(diff < 0) phiik = 1 < duimin/diff ? 1 : duimin/diff;
						else              phiik = 1;
					}
					if(phiik < lim) lim = phiik;
				}
				for(int j = 0; j < m->nnode[iel]; j++) {
					const int face = m->elemface[4*iel+j];
					const int jel = m->esuel[4*iel+j];
					const double val = extrapolate(u[4*iel+iv], grad+8*iel, iv, lim, &gr[2*face], &rc[2*iel]);
					if(iel < jel) ufl[4*face+iv] = val;
					else          ufr[4*face+iv] = val;
				}
			}
		}
	}

	/// Dispatch of afactory.cpp:178-211
	void face_values(const double *u, const double *ug, const double *grad, double *ufl, double *ufr) const
	{
		switch(nc.recon) {
		case RECON_NONE:      recon_linear(u, grad, ufl, ufr); break;
		case RECON_WENO:      recon_weno(u, grad, ufl, ufr); break;
		case RECON_VANALBADA: recon_muscl(u, ug, grad, ufl, ufr); break;
		case RECON_BJ:        recon_limited(false, u, ug, grad, ufl, ufr); break;
		default:              recon_limited(true, u, ug, grad, ufl, ufr); break;
		}
	}

	// ---------------------------------------------------------------- residual

	/// flow_spatial.cpp:489-563
	void compute_fluxes(const double *u, const double *gradients, const double *uleft, const double *uright,
	                    const double *ug, double *res) const
	{
		const bool o2 = nc.order2;
#pragma omp parallel for default(shared)
		for(int ied = 0; ied < m->naface; ied++) {
			const double *n = &m->facemetric[3*ied];
			const double len = m->facemetric[3*ied+2];
			const int lelem = m->intfac[4*ied], relem = m->intfac[4*ied+1];
			double fluxes[4];
			inviscid_flux(nc.flux, physics, &uleft[ied*4], &uright[ied*4], n, fluxes);
			for(int iv = 0; iv < 4; iv++) fluxes[iv] *= len;

			if(pc.viscous) {
				const bool isb = ied < m->nbface;
				const double *rcr = isb ? &rcbp[2*ied] : &rc[2*relem];
				const double *ucr = isb ? &ug[ied*4] : &u[relem*4];
				const double *gl = o2 ? gradients+8*lelem : nullptr;
				const double *grr = o2 ? (isb ? gradients+8*lelem : gradients+8*relem) : nullptr;
				double vflux[4];
				cell_viscous_flux(physics, o2, pc.const_visc, n, &rc[2*lelem], rcr, &u[lelem*4], ucr,
				                  gl, grr, &uleft[ied*4], &uright[ied*4], vflux);
				for(int iv = 0; iv < 4; iv++) fluxes[iv] += vflux[iv]*len;
			}

			for(int iv = 0; iv < 4; iv++) {
#pragma omp atomic update
				res[4*lelem+iv] -= fluxes[iv];
			}
			if(relem < m->nelem)
				for(int iv = 0; iv < 4; iv++) {
#pragma omp atomic update
					res[4*relem+iv] += fluxes[iv];
				}
		}
	}

	/// flow_spatial.cpp:567-634
	void compute_max_timestep(const double *uleft, const double *uright, double *timesteps) const
	{
		std::vector<double> integv(m->nelem);     // per-call allocation as the reference
		double *const integ = integv.data();
#pragma omp parallel for simd default(shared)
		for(int iel = 0; iel < m->nelem; iel++) integ[iel] = 0.0;

#pragma omp parallel for default(shared)
		for(int ied = 0; ied < m->naface; ied++) {
			const double *n = &m->facemetric[3*ied];
			const double len = m->facemetric[3*ied+2];
			const int lelem = m->intfac[4*ied], relem = m->intfac[4*ied+1];
			const double *ul = &uleft[4*ied], *ur = &uright[4*ied];
			const double ci = physics.soundSpeedFromConserved(ul);
			const double cj = physics.soundSpeedFromConserved(ur);
			const double vni = dot2(&ul[1],n)/ul[0];
			const double vnj = dot2(&ur[1],n)/ur[0];
			double specradi = (std::fabs(vni)+ci)*len;
			double specradj = (std::fabs(vnj)+cj)*len;
			if(pc.viscous) {
				double mui, muj;
				if(pc.const_visc) { mui = physics.constantViscosity(); muj = physics.constantViscosity(); }
				else { mui = physics.viscosityFromConserved(ul); muj = physics.viscosityFromConserved(ur); }
				const double coi = std::max(4.0/(3*ul[0]), physics.g/ul[0]);
				const double coj = std::max(4.0/(3*ur[0]), physics.g/ur[0]);
				specradi += coi*mui/physics.Pr * len*len/m->area[lelem];
				if(relem < m->nelem)
					specradj += coj*muj/physics.Pr * len*len/m->area[relem];
			}
#pragma omp atomic update
			integ[lelem] += specradi;
			if(relem < m->nelem) {
#pragma omp atomic update
				integ[relem] += specradj;
			}
		}
#pragma omp parallel for simd default(shared)
		for(int iel = 0; iel < m->nelem; iel++)
			timesteps[iel] = m->area[iel]/integ[iel];
	}

	/// flow_spatial.cpp:637-816. Adds -r(u) into res (caller zeroes). If face_out != null the final
	/// conserved left/right face states are copied there ([2][naface][4]) for unit parity.
	void compute_residual(const double *u, double *res, bool gettimesteps, double *dtm,
	                      double *grad_out = nullptr, double *face_out = nullptr) const
	{
		const int nb = m->nbface, nf = m->naface, ne = m->nelem;
		std::vector<double> ulv(4*(size_t)nf), urv(4*(size_t)nf);
		double *const uleft = ulv.data(), *const uright = urv.data();

		// P1
#pragma omp parallel for default(shared)
		for(int ied = 0; ied < nb; ied++) {
			const int ie = m->intfac[4*ied];
			for(int iv = 0; iv < 4; iv++) uleft[4*ied+iv] = u[4*ie+iv];
		}
		std::vector<double> ubcell(4*(size_t)nb);
		std::vector<double> gradv;

		if(nc.order2) {
			boundary_states(uleft, uright);                          // P2
			std::vector<double> up(4*(size_t)ne);
#pragma omp parallel default(shared)
			{
#pragma omp for
				for(int f = 0; f < nb; f++) {                         // P3
					for(int j = 0; j < 4; j++) ubcell[4*f+j] = uright[4*f+j];
					physics.primitiveFromConserved(&uright[4*f], &uright[4*f]);
				}
#pragma omp for
				for(int iel = 0; iel < ne; iel++)                     // P4
					physics.primitiveFromConserved(&u[4*iel], &up[4*iel]);
			}
			gradv.resize(8*(size_t)ne);
			gradients(up.data(), uright, gradv.data());              // P5
			face_values(up.data(), uright, gradv.data(), uleft, uright);   // P6
#pragma omp parallel default(shared)
			{
#pragma omp for
				for(int f = nb; f < nf; f++) {                        // P8
					physics.conservedFromPrimitive(&uleft[4*f], &uleft[4*f]);
					physics.conservedFromPrimitive(&uright[4*f], &uright[4*f]);
				}
#pragma omp for
				for(int f = 0; f < nb; f++)
					physics.conservedFromPrimitive(&uleft[4*f], &uleft[4*f]);
			}
		}
		else {
#pragma omp parallel for default(shared)
			for(int ied = nb; ied < nf; ied++) {
				const int ie = m->intfac[4*ied], je = m->intfac[4*ied+1];
				for(int iv = 0; iv < 4; iv++) {
					uleft[4*ied+iv] = u[4*ie+iv];
					uright[4*ied+iv] = u[4*je+iv];
				}
			}
		}

		boundary_states(uleft, uright);                              // P9

		const double *ug_pb = nc.order2 ? ubcell.data() : uright;
		compute_fluxes(u, nc.order2 ? gradv.data() : nullptr, uleft, uright, ug_pb, res);   // P10
		if(gettimesteps)
			compute_max_timestep(uleft, uright, dtm);                // P11

		if(grad_out && nc.order2) std::memcpy(grad_out, gradv.data(), sizeof(double)*8*(size_t)ne);
		if(face_out) {
			std::memcpy(face_out, uleft, sizeof(double)*4*(size_t)nf);
			std::memcpy(face_out+4*(size_t)nf, uright, sizeof(double)*4*(size_t)nf);
		}
	}

	/// flow_spatial.cpp:96-112 : gradients of CONSERVED variables with conserved ghost states
	void get_gradients(const double *u, double *grads) const
	{
		std::vector<double> ug(4*(size_t)m->nbface);
		for(int f = 0; f < m->nbface; f++)
			boundary_state(f, &u[4*m->intfac[4*f]], &ug[4*f]);
		gradients(u, ug.data(), grads);
	}

	/// flow_spatial.cpp:131-310: returns (Cl, Cdp, Cdf) over faces carrying marker iwbcm
	void surface_data(const double *u, const double *grad, int iwbcm, double *out3) const
	{
		const double wind[2] = { std::cos(pc.aoa)*std::cos(0.0), std::sin(pc.aoa)*std::cos(0.0) };
		double totalarea = 0, Cdf = 0, Cdp = 0, Cl = 0;
		const double pinf = physics.freestreamPressure();
		const double flownormal[2] = { -wind[1], wind[0] };
		for(int f = 0; f < m->nbface; f++) {
			if(m->btags[f] != iwbcm) continue;
			const int lelem = m->intfac[4*f];
			const double *n = &m->facemetric[3*f];
			const double len = m->facemetric[3*f+2];
			const double tangf[2] = { n[1], -n[0] };
			const double *urec = &u[4*lelem];
			const double cp = (physics.pressureFromConserved(urec) - pinf)*2.0;
			const double muhat = physics.viscosityFromConserved(urec);
			double gradu[2][2];
			for(int i = 0; i < 2; i++)
				for(int j = 0; j < 2; j++)
					gradu[i][j] = (grad[8*lelem+j+2*(i+1)]*urec[0] - urec[i+1]*grad[8*lelem+j+2*0]) /
						(urec[0]*urec[0]);
			double force[2];
			for(int i = 0; i < 2; i++) {
				force[i] = 0;
				for(int j = 0; j < 2; j++)
					force[i] += (gradu[i][j] + gradu[j][i])*n[j];
			}
			const double tauw = muhat*dot2(force,tangf);
			const double cf = 2*tauw;
			const double ndotw = dot2(n,wind);
			const double ndotnw = dot2(n,flownormal);
			const double tdotw = dot2(tangf,wind);
			totalarea += len;
			Cl += cp*ndotnw*len;
			Cdp += cp*ndotw*len;
			Cdf += cf*tdotw*len;
		}
		out3[0] = Cl/totalarea; out3[1] = Cdp/totalarea; out3[2] = Cdf/totalarea;
	}

	/// aoutput.cpp:28-63
	double entropy_error(const double *u) const
	{
		const double sinf = physics.entropyFromConserved(uinf);
		double error = 0;
		for(int iel = 0; iel < m->nelem; iel++) {
			const double s_err = (physics.entropyFromConserved(&u[4*iel]) - sinf) / sinf;
			error += s_err*s_err*m->area[iel];
		}
		return std::sqrt(error);
	}
};

/// ode/aodesolver.cpp:136-282. Returns 0 converged, 1 = hit maxiter (Tolerance_error in the
/// reference), 2 = non-finite residual (Numerical_error). hist[step] = resi (absolute) if non-null.
static inline int forward_euler(const Flow &fl, double *u, double cfl, double tol, int maxiter,
                                int *steps_out, double *hist)
{
	const Mesh *m = fl.m;
	const int ne = m->nelem;
	if(maxiter <= 0) { *steps_out = 0; return 0; }
	std::vector<double> r(4*(size_t)ne), dtm(ne);
	int step = 0;
	double resi = 1.0, initres = 1.0;
	while(resi/initres > tol && step < maxiter) {
#pragma omp parallel for simd default(shared)
		for(int i = 0; i < ne*4; i++) r[i] = 0;

		fl.compute_residual(u, r.data(), true, dtm.data());

#pragma omp parallel for default(shared)
		for(int iel = 0; iel < ne; iel++)
			for(int i = 0; i < 4; i++)
				u[4*iel+i] += cfl*dtm[iel] * 1.0/m->area[iel]*r[4*iel+i];

		double locresenergy = 0;
#pragma omp parallel for simd reduction(+:locresenergy) default(shared)
		for(int iel = 0; iel < ne; iel++)
			locresenergy += r[4*iel+3]*r[4*iel+3]*m->area[iel];

		resi = std::sqrt(locresenergy);
		if(step == 0) initres = resi;
		if(hist) hist[step] = resi;
		step++;
		if(!std::isfinite(resi)) { *steps_out = step; return 2; }
	}
	*steps_out = step;
	return (step == maxiter) ? 1 : 0;
}

} // namespace orc
#endif
