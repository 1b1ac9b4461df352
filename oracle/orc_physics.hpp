/* ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the pointwise gas dynamics used by FVENS's residual path. Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use anything
 * under oracle/. The product (fvens_b200/) never includes, links or calls this.
 *
 * Every function cites the reference lines it follows (paths relative to /root/reference/src).
 * Arithmetic is written in the same operation order as the reference so that, compiled with
 * -ffp-contract=off, results agree with the reference's own object code (oracle/_ref, tier A)
 * to the last bit for everything except libm calls.
 */
#ifndef ORC_PHYSICS_HPP
#define ORC_PHYSICS_HPP

#include <cmath>
#include <algorithm>

namespace orc {

constexpr int NDIM = 2;
constexpr int NVARS = 4;

/// BC type ids, same order as the reference enum (spatial/abctypes.hpp:13-22)
enum BCType {
	SLIP_WALL_BC = 0, FARFIELD_BC = 1, INFLOW_OUTFLOW_BC = 2, SUBSONIC_INFLOW_BC = 3,
	EXTRAPOLATION_BC = 4, PERIODIC_BC = 5, ISOTHERMAL_WALL_BC = 6, ADIABATIC_WALL_BC = 7
};

/// Flux ids; string keys as in utilities/afactory.cpp:38-81
enum FluxId { FLUX_LLF = 0, FLUX_VANLEER = 1, FLUX_AUSM = 2, FLUX_AUSMPLUS = 3, FLUX_ROE = 4,
              FLUX_HLL = 5, FLUX_HLLC = 6 };

static inline double dot2(const double *a, const double *b) {
	// mathutils.hpp:24-30 : dot starts at 0 and accumulates
	double d = 0;
	d += a[0]*b[0];
	d += a[1]*b[1];
	return d;
}

/// physics/aphysics.hpp:48-336, ctor physics/aphysics.cpp:17-19
struct Gas {
	double g, Minf, Tinf, Reinf, Pr, sC;

	Gas(double g_, double M, double T, double Re, double Pr_)
		: g(g_), Minf(M), Tinf(T), Reinf(Re), Pr(Pr_), sC(110.5) { }

	// aphysics_defs.hpp:53-63
	double pressure(double ie) const { return (g-1.0)*ie; }
	double pressureFromConserved(const double *uc) const {
		return pressure(uc[3] - 0.5*dot2(&uc[1],&uc[1])/uc[0]);
	}
	// aphysics_defs.hpp:119-122
	double temperature(double rho, double p) const { return p/rho * g*Minf*Minf; }
	// aphysics_defs.hpp:140-143
	double soundSpeed(double rho, double p) const { return std::sqrt(g * p/rho); }
	// aphysics_defs.hpp:160-163
	double soundSpeedFromConserved(const double *uc) const {
		return soundSpeed(uc[0], pressureFromConserved(uc));
	}
	// aphysics_defs.hpp:204-207
	double entropyFromConserved(const double *uc) const {
		return pressureFromConserved(uc)/std::pow(uc[0],g);
	}
	// aphysics_defs.hpp:211-223
	double energyFromPressure(double p, double d, double vmag2) const {
		return p/(g-1.0) + 0.5*d*vmag2;
	}
	double energyFromTemperature(double T, double d, double vmag2) const {
		return d * (T/(g*(g-1.0)*Minf*Minf) + 0.5*vmag2);
	}
	// aphysics_defs.hpp:241-244
	double energyFromPrimitive(const double *up) const {
		return energyFromPressure(up[3], up[0], dot2(&up[1],&up[1]));
	}
	// aphysics_defs.hpp:259-267 (in-place safe)
	void primitiveFromConserved(const double *uc, double *up) const {
		const double rho = uc[0];
		const double p = pressureFromConserved(uc);
		up[0] = rho;
		up[1] = uc[1]/rho;
		up[2] = uc[2]/rho;
		up[3] = p;
	}
	// aphysics_defs.hpp:273-281
	void primitive2FromConserved(const double *uc, double *up) const {
		const double rho = uc[0];
		const double p = pressureFromConserved(uc);
		up[0] = rho;
		up[1] = uc[1]/rho;
		up[2] = uc[2]/rho;
		up[3] = temperature(rho,p);
	}
	// aphysics_defs.hpp:287-295 (in-place safe)
	void conservedFromPrimitive(const double *up, double *uc) const {
		const double rho = up[0];
		const double rhoE = energyFromPrimitive(up);
		uc[0] = rho;
		uc[1] = rho*up[1];
		uc[2] = rho*up[2];
		uc[3] = rhoE;
	}
	// aphysics_defs.hpp:299-303
	double densityFromPressureTemperature(double p, double T) const { return g*Minf*Minf*p/T; }
	// aphysics_defs.hpp:319-322
	double temperatureFromConserved(const double *uc) const {
		return temperature(uc[0], pressureFromConserved(uc));
	}
	// aphysics_defs.hpp:349-353
	double gradTemperature(double rho, double gradrho, double p, double gradp) const {
		return (gradp*rho - p*gradrho) / (rho*rho) * g*Minf*Minf;
	}
	// aphysics_defs.hpp:410-421
	double viscosityFromTemperature(double T) const {
		return (1.0+sC/Tinf)/(T+sC/Tinf) * std::pow(T,1.5) / Reinf;
	}
	double viscosityFromConserved(const double *uc) const {
		return viscosityFromTemperature(temperatureFromConserved(uc));
	}
	// aphysics_defs.hpp:443-451
	double constantViscosity() const { return 1.0/Reinf; }
	double thermalConductivity(double muhat) const { return muhat / (Minf*Minf*(g-1.0)*Pr); }
	// aphysics_defs.hpp:465-467
	double freestreamPressure() const { return 1.0/(g*Minf*Minf); }

	// aphysics_defs.hpp:15-23
	void directionalFlux(const double *uc, const double *n, double vn, double p, double *flux) const {
		flux[0] = vn*uc[0];
		flux[1] = vn*uc[1] + p*n[0];
		flux[2] = vn*uc[2] + p*n[1];
		flux[3] = vn*(uc[3] + p);
	}
	// aphysics.cpp:28-35
	void directionalFluxFromConserved(const double *u, const double *n, double *flux) const {
		const double vn = dot2(&u[1],n)/u[0];
		const double p = pressure(u[3] - 0.5*dot2(&u[1],&u[1])/u[0]);
		directionalFlux(u, n, vn, p, flux);
	}
	// aphysics_defs.hpp:27-38
	void varsFromConserved(const double *uc, const double *n, double *v, double &vn, double &p,
	                       double &H) const {
		v[0] = uc[1]/uc[0];
		v[1] = uc[2]/uc[0];
		vn = dot2(v,n);
		const double vmag2 = dot2(v,v);
		p = (g-1.0)*(uc[3] - 0.5*uc[0]*vmag2);
		H = (uc[3]+p)/uc[0];
	}
	// aphysics.cpp:44-58 with mathutils.hpp:61-70 (beta = 0)
	void freestreamState(double aoa, double *uinf) const {
		uinf[0] = 1.0;
		uinf[1] = std::cos(aoa)*std::cos(0.0);
		uinf[2] = std::sin(aoa)*std::cos(0.0);
		uinf[3] = energyFromPressure(freestreamPressure(), 1.0, 1.0);
	}
	// aphysics_defs.hpp:471-487
	void stressTensor(double mu, const double grad[NDIM][NVARS], double stress[NDIM][NDIM]) const {
		double ldiv = 0;
		for(int j = 0; j < NDIM; j++)
			ldiv += grad[j][j+1];
		ldiv *= 2.0/3.0*mu;
		for(int i = 0; i < NDIM; i++) {
			for(int j = 0; j < NDIM; j++)
				stress[i][j] = mu*(grad[i][j+1] + grad[j][i+1]);
			stress[i][i] -= ldiv;
		}
	}
};

// ------------------------------------------------------------------------------------------------
// Inviscid numerical fluxes: spatial/anumericalflux.cpp
// ------------------------------------------------------------------------------------------------

/// anumericalflux.cpp:41-61
static inline void flux_llf(const Gas &ph, const double *ul, const double *ur, const double *n,
                            double *flux)
{
	double vi[2], vj[2], vni, vnj, pi, pj, Hi, Hj;
	ph.varsFromConserved(ul, n, vi, vni, pi, Hi);
	ph.varsFromConserved(ur, n, vj, vnj, pj, Hj);
	const double ci = ph.soundSpeed(ul[0],pi);
	const double cj = ph.soundSpeed(ur[0],pj);
	const double eig = std::fabs(vni)+ci > std::fabs(vnj)+cj ? std::fabs(vni)+ci : std::fabs(vnj)+cj;
	double fl[4], fr[4];
	ph.directionalFluxFromConserved(ul,n,fl);
	ph.directionalFluxFromConserved(ur,n,fr);
	for(int i = 0; i < 4; i++)
		flux[i] = 0.5*( fl[i] + fr[i] - eig*(ur[i]-ul[i]) );
}

/// anumericalflux.cpp:203-250
static inline void flux_vanleer(const Gas &ph, const double *ul, const double *ur, const double *n,
                                double *flux)
{
	const double g = ph.g;
	double fiplus[4], fjminus[4];
	double vi[2], vj[2], vni, vnj, pi, pj, Hi, Hj;
	ph.varsFromConserved(ul, n, vi, vni, pi, Hi);
	ph.varsFromConserved(ur, n, vj, vnj, pj, Hj);
	const double ci = ph.soundSpeed(ul[0],pi);
	const double cj = ph.soundSpeed(ur[0],pj);
	const double Mni = vni/ci;
	const double Mnj = vnj/cj;

	if(Mni < -1.0)
		for(int i = 0; i < 4; i++) fiplus[i] = 0;
	else if(Mni > 1.0)
		ph.directionalFlux(ul,n,vni,pi,fiplus);
	else {
		const double vmags = std::pow(ul[1]/ul[0], 2) + std::pow(ul[2]/ul[0], 2);
		fiplus[0] = ul[0]*ci*std::pow(Mni+1, 2)/4.0;
		fiplus[1] = fiplus[0] * (ul[1]/ul[0] + n[0]*(2.0*ci - vni)/g);
		fiplus[2] = fiplus[0] * (ul[2]/ul[0] + n[1]*(2.0*ci - vni)/g);
		fiplus[3] = fiplus[0] * ( (vmags - vni*vni)/2.0 + std::pow((g-1)*vni+2*ci, 2)/(2*(g*g-1)) );
	}

	if(Mnj > 1.0)
		for(int i = 0; i < 4; i++) fjminus[i] = 0;
	else if(Mnj < -1.0)
		ph.directionalFlux(ur,n,vnj,pj,fjminus);
	else {
		const double vmags = std::pow(ur[1]/ur[0], 2) + std::pow(ur[2]/ur[0], 2);
		fjminus[0] = -ur[0]*cj*std::pow(Mnj-1, 2)/4.0;
		fjminus[1] = fjminus[0] * (ur[1]/ur[0] + n[0]*(-2.0*cj - vnj)/g);
		fjminus[2] = fjminus[0] * (ur[2]/ur[0] + n[1]*(-2.0*cj - vnj)/g);
		fjminus[3] = fjminus[0] * ( (vmags - vnj*vnj)/2.0 + std::pow((g-1)*vnj-2*cj, 2)/(2*(g*g-1)) );
	}

	for(int i = 0; i < 4; i++)
		flux[i] = fiplus[i] + fjminus[i];
}

/// anumericalflux.cpp:265-315
static inline void flux_ausm(const Gas &ph, const double *ul, const double *ur, const double *n,
                             double *flux)
{
	double ML, MR, pL, pR;
	double vi[2], vj[2], vni, vnj, pi, pj, Hi, Hj;
	ph.varsFromConserved(ul, n, vi, vni, pi, Hi);
	ph.varsFromConserved(ur, n, vj, vnj, pj, Hj);
	const double ci = ph.soundSpeed(ul[0],pi);
	const double cj = ph.soundSpeed(ur[0],pj);
	const double Mni = vni/ci, Mnj = vnj/cj;

	if(std::fabs(Mni) <= 1.0) {
		ML = 0.25*(Mni+1)*(Mni+1);
		pL = ML*pi*(2.0-Mni);
	}
	else if(Mni < -1.0) { ML = 0; pL = 0; }
	else { ML = Mni; pL = pi; }

	if(std::fabs(Mnj) <= 1.0) {
		MR = -0.25*(Mnj-1)*(Mnj-1);
		pR = -MR*pj*(2.0+Mnj);
	}
	else if(Mnj < -1.0) { MR = Mnj; pR = pj; }
	else { MR = 0; pR = 0; }

	const double Mhalf = ML+MR;
	const double phalf = pL+pR;

	flux[0] = Mhalf/2.0*(ul[0]*ci+ur[0]*cj) -std::fabs(Mhalf)/2.0*(ur[0]*cj-ul[0]*ci);
	for(int j = 1; j < 3; j++)
		flux[j] = Mhalf/2.0*(ul[j]*ci+ur[j]*cj) -std::fabs(Mhalf)/2.0*(ur[j]*cj-ul[j]*ci) + phalf*n[j-1];
	flux[3] = Mhalf/2.0*(ci*(ul[3]+pi)+cj*(ur[3]+pj))
		-std::fabs(Mhalf)/2.0*(cj*(ur[3]+pj)-ci*(ul[3]+pi));
}

/// anumericalflux.cpp:480-553
static inline void flux_ausmplus(const Gas &ph, const double *ul, const double *ur, const double *n,
                                 double *flux)
{
	const double g = ph.g;
	double ML, MR, pL, pR;
	double vi[2], vj[2], vni, vnj, pi, pj, Hi, Hj;
	ph.varsFromConserved(ul, n, vi, vni, pi, Hi);
	ph.varsFromConserved(ur, n, vj, vnj, pj, Hj);
	const double ci = ph.soundSpeed(ul[0],pi);
	const double cj = ph.soundSpeed(ur[0],pj);
	const double vmag2i = dot2(vi,vi);
	const double vmag2j = dot2(vj,vj);

	double csi = std::sqrt((ci*ci/(g-1.0)+0.5*vmag2i)*2.0*(g-1.0)/(g+1.0));
	double csj = std::sqrt((cj*cj/(g-1.0)+0.5*vmag2j)*2.0*(g-1.0)/(g+1.0));
	const double corri = (csi > vni) ? csi : vni;
	const double corrj = (csj > -vnj) ? csj : -vnj;
	csi = csi*csi/corri;
	csj = csj*csj/corrj;
	const double chalf = (csi < csj) ? csi : csj;

	const double Mni = vni/chalf, Mnj = vnj/chalf;

	if(std::fabs(Mni) <= 1.0) {
		ML = 0.25*(Mni+1)*(Mni+1) + 1.0/8.0*(Mni*Mni-1.0)*(Mni*Mni-1.0);
		pL = pi*(0.25*(Mni+1)*(Mni+1)*(2.0-Mni) + 3.0/16*Mni*(Mni*Mni-1.0)*(Mni*Mni-1.0));
	}
	else if(Mni < -1.0) { ML = 0; pL = 0; }
	else { ML = Mni; pL = pi; }

	if(std::fabs(Mnj) <= 1.0) {
		MR = -0.25*(Mnj-1)*(Mnj-1) - 1.0/8.0*(Mnj*Mnj-1.0)*(Mnj*Mnj-1.0);
		pR = pj*(0.25*(Mnj-1)*(Mnj-1)*(2.0+Mnj) - 3.0/16*Mnj*(Mnj*Mnj-1.0)*(Mnj*Mnj-1.0));
	}
	else if(Mnj < -1.0) { MR = Mnj; pR = pj; }
	else { MR = 0; pR = 0; }

	const double Mhalf = ML+MR;
	const double phalf = pL+pR;

	flux[0] = chalf* (Mhalf/2.0*(ul[0]+ur[0]) -std::fabs(Mhalf)/2.0*(ur[0]-ul[0]));
	for(int j = 1; j < 3; j++)
		flux[j] = chalf* (Mhalf/2.0*(ul[j]+ur[j]) -std::fabs(Mhalf)/2.0*(ur[j]-ul[j])) + phalf*n[j-1];
	flux[3] = chalf* (Mhalf/2.0*(ul[3]+pi+ur[3]+pj) -std::fabs(Mhalf)/2.0*((ur[3]+pj)-(ul[3]+pi)));
}

/// Roe averages, anumericalflux.hpp:175-189
struct RoeAvg { double Rij, rhoij, vij[2], vm2ij, vnij, Hij, cij; };
static inline RoeAvg roe_averages(const Gas &ph, const double *ul, const double *ur, const double *n,
                                  const double *vi, double Hi, const double *vj, double Hj)
{
	RoeAvg a;
	a.Rij = std::sqrt(ur[0]/ul[0]);
	a.rhoij = a.Rij*ul[0];
	for(int i = 0; i < 2; i++)
		a.vij[i] = (a.Rij*vj[i] + vi[i])/(a.Rij + 1.0);
	a.Hij = (a.Rij*Hj + Hi)/(a.Rij + 1.0);
	a.vm2ij = dot2(a.vij,a.vij);
	a.vnij = dot2(a.vij,n);
	a.cij = std::sqrt( (ph.g-1.0)*(a.Hij - a.vm2ij*0.5) );
	return a;
}

/// anumericalflux.cpp:664 (fixeps), 668-732
static inline void flux_roe(const Gas &ph, const double *ul, const double *ur, const double *n,
                            double *flux)
{
	const double fixeps = 1.0e-4;
	double vi[2], vj[2], vni, vnj, pi, pj, Hi, Hj;
	ph.varsFromConserved(ul, n, vi, vni, pi, Hi);
	ph.varsFromConserved(ur, n, vj, vnj, pj, Hj);
	const RoeAvg a = roe_averages(ph, ul,ur,n,vi,Hi,vj,Hj);
	const double rhoij=a.rhoij, vm2ij=a.vm2ij, vnij=a.vnij, Hij=a.Hij, cij=a.cij;
	const double *vij = a.vij;

	double l[4];
	l[0] = std::fabs(vnij-cij);
	l[1] = std::fabs(vnij);
	l[2] = std::fabs(vnij);
	l[3] = std::fabs(vnij+cij);

	const double delta = fixeps*cij;
	for(int ivar = 0; ivar < 4; ivar++)
		if(l[ivar] < delta)
			l[ivar] = (l[ivar]*l[ivar] + delta*delta)/(2.0*delta);

	const double devn = vnj-vni, dep = pj-pi, derho = ur[0]-ul[0];
	double adu[4], lalpha[4];
	lalpha[0] = l[0]*(dep-rhoij*cij*devn)/(2.0*cij*cij);
	lalpha[1] = l[1]*(derho - dep/(cij*cij));
	lalpha[2] = l[1]*rhoij;
	lalpha[3] = l[3]*(dep+rhoij*cij*devn)/(2.0*cij*cij);

	adu[0] = lalpha[0];
	adu[1] = lalpha[0]*(vij[0]-cij*n[0]);
	adu[2] = lalpha[0]*(vij[1]-cij*n[1]);
	adu[3] = lalpha[0]*(Hij-cij*vnij);

	adu[0] += lalpha[1];
	adu[1] += lalpha[1]*vij[0] +      lalpha[2]*(vj[0]-vi[0] - devn*n[0]);
	adu[2] += lalpha[1]*vij[1] +      lalpha[2]*(vj[1]-vi[1] - devn*n[1]);
	adu[3] += lalpha[1]*vm2ij/2.0 + lalpha[2] *(vij[0]*(vj[0]-vi[0]) +vij[1]*(vj[1]-vi[1]) -vnij*devn);

	adu[0] += lalpha[3];
	adu[1] += lalpha[3]*(vij[0]+cij*n[0]);
	adu[2] += lalpha[3]*(vij[1]+cij*n[1]);
	adu[3] += lalpha[3]*(Hij+cij*vnij);

	double fi[4], fj[4];
	ph.directionalFlux(ul,n,vni,pi,fi);
	ph.directionalFlux(ur,n,vnj,pj,fj);
	for(int ivar = 0; ivar < 4; ivar++)
		flux[ivar] = 0.5*(fi[ivar]+fj[ivar] - adu[ivar]);
}

/// anumericalflux.cpp:974-1007
static inline void flux_hll(const Gas &ph, const double *ul, const double *ur, const double *n,
                            double *flux)
{
	double vi[2], vj[2], vni, vnj, pi, pj, Hi, Hj;
	ph.varsFromConserved(ul, n, vi, vni, pi, Hi);
	ph.varsFromConserved(ur, n, vj, vnj, pj, Hj);
	const double ci = ph.soundSpeed(ul[0], pi);
	const double cj = ph.soundSpeed(ur[0], pj);
	const RoeAvg a = roe_averages(ph, ul,ur,n,vi,Hi,vj,Hj);

	double sl = vni - ci;
	if(sl > a.vnij-a.cij) sl = a.vnij-a.cij;
	double sr = vnj+cj;
	if(sr < a.vnij+a.cij) sr = a.vnij+a.cij;
	const double sr0 = sr > 0 ? 0 : sr;
	const double sl0 = sl > 0 ? 0 : sl;

	const double t1 = (sr0 - sl0)/(sr-sl); const double t2 = 1.0 - t1;
	const double t3 = 0.5*(sr*std::fabs(sl)-sl*std::fabs(sr))/(sr-sl);
	flux[0] = t1*vnj*ur[0] + t2*vni*ul[0]                     - t3*(ur[0]-ul[0]);
	flux[1] = t1*(vnj*ur[1]+pj*n[0]) + t2*(vni*ul[1]+pi*n[0]) - t3*(ur[1]-ul[1]);
	flux[2] = t1*(vnj*ur[2]+pj*n[1]) + t2*(vni*ul[2]+pi*n[1]) - t3*(ur[2]-ul[2]);
	flux[3] = t1*(vnj*ur[0]*Hj) + t2*(vni*ul[0]*Hi)           - t3*(ur[3]-ul[3]);
}

/// anumericalflux.cpp:1071-1081
static inline void hllc_star(const double *u, const double *n, double vn, double p, double ss,
                             double sm, double *ustr)
{
	const double pstar = u[0]*(vn-ss)*(vn-sm) + p;
	ustr[0] = u[0] * (ss - vn)/(ss-sm);
	ustr[1] = ( (ss-vn)*u[1] + (pstar-p)*n[0] )/(ss-sm);
	ustr[2] = ( (ss-vn)*u[2] + (pstar-p)*n[1] )/(ss-sm);
	ustr[3] = ( (ss-vn)*u[3] - p*vn + pstar*sm )/(ss-sm);
}

/// anumericalflux.cpp:1176-1228
static inline void flux_hllc(const Gas &ph, const double *ul, const double *ur, const double *n,
                             double *flux)
{
	double vi[2], vj[2], vni, vnj, pi, pj, Hi, Hj;
	ph.varsFromConserved(ul, n, vi, vni, pi, Hi);
	ph.varsFromConserved(ur, n, vj, vnj, pj, Hj);
	const double ci = ph.soundSpeed(ul[0], pi);
	const double cj = ph.soundSpeed(ur[0], pj);
	const RoeAvg a = roe_averages(ph, ul,ur,n,vi,Hi,vj,Hj);

	double sl = vni - ci;
	if(sl > a.vnij-a.cij) sl = a.vnij-a.cij;
	double sr = vnj+cj;
	if(sr < a.vnij+a.cij) sr = a.vnij+a.cij;
	const double sm = ( ur[0]*vnj*(sr-vnj) - ul[0]*vni*(sl-vni) + pi-pj )
		/ ( ur[0]*(sr-vnj) - ul[0]*(sl-vni) );

	if(sl > 0)
		ph.directionalFlux(ul,n,vni,pi,flux);
	else if(sl <= 0 && sm > 0) {
		ph.directionalFlux(ul,n,vni,pi,flux);
		double ustr[4];
		hllc_star(ul,n,vni,pi,sl,sm,ustr);
		for(int k = 0; k < 4; k++)
			flux[k] += sl * ( ustr[k] - ul[k]);
	}
	else if(sm <= 0 && sr >= 0) {
		ph.directionalFlux(ur,n,vnj,pj,flux);
		double ustr[4];
		hllc_star(ur,n,vnj,pj,sr,sm,ustr);
		for(int k = 0; k < 4; k++)
			flux[k] += sr * ( ustr[k] - ur[k]);
	}
	else
		ph.directionalFlux(ur,n,vnj,pj,flux);
}

static inline void inviscid_flux(int id, const Gas &ph, const double *ul, const double *ur,
                                 const double *n, double *flux)
{
	switch(id) {
	case FLUX_LLF:      flux_llf(ph,ul,ur,n,flux); break;
	case FLUX_VANLEER:  flux_vanleer(ph,ul,ur,n,flux); break;
	case FLUX_AUSM:     flux_ausm(ph,ul,ur,n,flux); break;
	case FLUX_AUSMPLUS: flux_ausmplus(ph,ul,ur,n,flux); break;
	case FLUX_ROE:      flux_roe(ph,ul,ur,n,flux); break;
	case FLUX_HLL:      flux_hll(ph,ul,ur,n,flux); break;
	default:            flux_hllc(ph,ul,ur,n,flux); break;
	}
}

// ------------------------------------------------------------------------------------------------
// Boundary conditions: spatial/abc.cpp
// ------------------------------------------------------------------------------------------------

/// One boundary condition: tag -> (type, values). abc.hpp:34-40, factory abc.cpp:461-500
struct BC {
	int tag;
	int type;
	double vals[2];   ///< adiabatic wall: tangential velocity; isothermal: (vt, T); inflow: (p0, T0)
};

/// abc.cpp:49-84 (InOutFlow), 152-176 (InFlow), 194-199 (Farfield), 218-226 (Slipwall),
/// 272-280 (Adiabaticwall2D), 354-369 (Isothermalwall2D), 417-423 (Extrapolation)
static inline void ghost_state(const BC &bc, const Gas &phy, const double *uinf, const double *ins,
                               const double *n, double *gs)
{
	switch(bc.type) {
	case INFLOW_OUTFLOW_BC: {
		const double vni = dot2(&ins[1],&n[0])/ins[0];
		const double ci = phy.soundSpeedFromConserved(ins);
		const double Mni = vni/ci;
		const double pinf = phy.freestreamPressure();
		if(Mni <= 0) {
			for(int i = 0; i < 4; i++) gs[i] = uinf[i];
		}
		else if(Mni < 1) {
			gs[0] = ins[0];
			gs[1] = ins[1];
			gs[2] = ins[2];
			gs[3] = phy.energyFromPressure(pinf, ins[0], dot2(&ins[1],&ins[1])/(ins[0]*ins[0]) );
		}
		else {
			for(int i = 0; i < 4; i++) gs[i] = ins[i];
		}
		break;
	}
	case SUBSONIC_INFLOW_BC: {
		const double ptotal = bc.vals[0], ttotal = bc.vals[1];
		const double ci = phy.soundSpeedFromConserved(ins);
		const double Rminus = dot2(&ins[1],&n[0])/ins[0] - ci/(2*phy.g - 1.0);
		const double co2 = ci*ci + (phy.g-1.0)/2.0 * dot2(&ins[1],&ins[1])/(ins[0]*ins[0]);
		const double q = std::sqrt((phy.g+1)*co2/((phy.g-1)*Rminus*Rminus) - (phy.g-1)/2.0);
		const double cg = -Rminus*(phy.g-1)/(phy.g+1) * (1.0 + q);
		const double tg = ttotal*cg*cg/co2;
		const double pg = ptotal * std::pow(tg/ttotal, phy.g/(phy.g-1.0));
		gs[0] = phy.densityFromPressureTemperature(pg,tg);
		const double vgmag = std::sqrt(2.0/(phy.g-1.0)*(co2 - cg*cg));
		// mathutils.hpp:34-57 in 2D: cosphi = 1
		const double vg0 = vgmag*1.0*n[0];
		const double vg1 = vgmag*1.0*n[1];
		gs[1] = gs[0]*vg0;
		gs[2] = gs[0]*vg1;
		gs[3] = phy.energyFromPressure(pg,gs[0],vgmag*vgmag);
		break;
	}
	case FARFIELD_BC:
		for(int i = 0; i < 4; i++) gs[i] = uinf[i];
		break;
	case SLIP_WALL_BC: {
		const double vni = dot2(&ins[1],&n[0])/ins[0];
		gs[0] = ins[0];
		gs[1] = ins[1] - 2.0*vni*n[0]*ins[0];
		gs[2] = ins[2] - 2.0*vni*n[1]*ins[0];
		gs[3] = ins[3];
		break;
	}
	case ADIABATIC_WALL_BC: {
		const double tangMomentum = bc.vals[0] * ins[0];
		gs[0] = ins[0];
		gs[1] =  2.0*tangMomentum*n[1] - ins[1];
		gs[2] = -2.0*tangMomentum*n[0] - ins[2];
		gs[3] = ins[3];
		break;
	}
	case ISOTHERMAL_WALL_BC: {
		const double tangvel = bc.vals[0], walltemperature = bc.vals[1];
		const double p = phy.pressureFromConserved(ins);
		const double gtemp = 2.0*walltemperature - phy.temperature(ins[0],p);
		gs[0] = ins[0];
		gs[1] = gs[0]*( 2.0*tangvel*n[1] - ins[1]/ins[0]);
		gs[2] = gs[0]*(-2.0*tangvel*n[0] - ins[2]/ins[0]);
		const double vmag2 = dot2(&gs[1],&gs[1])/(gs[0]*gs[0]);
		gs[3] = phy.energyFromTemperature(gtemp, gs[0], vmag2);
		break;
	}
	default: // EXTRAPOLATION_BC
		for(int k = 0; k < 4; k++) gs[k] = ins[k];
	}
}

// ------------------------------------------------------------------------------------------------
// Viscous flux pieces: physics/viscousphysics.cpp, spatial/aspatial.cpp
// ------------------------------------------------------------------------------------------------

/// physics/viscousphysics.cpp:15-68. gradl/gradr are [dim][var] row-major and are converted in place.
static inline void primitive2_states_and_gradients(const Gas &ph, bool order2,
                                                   const double *ucl, const double *ucr,
                                                   double *gradl, double *gradr,
                                                   double *uctl, double *uctr)
{
	if(order2) {
		ph.primitiveFromConserved(ucl, uctl);
		ph.primitiveFromConserved(ucr, uctr);
		for(int j = 0; j < NDIM; j++) {
			gradl[j*NVARS+3] = ph.gradTemperature(uctl[0], gradl[j*NVARS], uctl[3], gradl[j*NVARS+3]);
			gradr[j*NVARS+3] = ph.gradTemperature(uctr[0], gradr[j*NVARS], uctr[3], gradr[j*NVARS+3]);
		}
		uctl[3] = ph.temperature(uctl[0], uctl[3]);
		uctr[3] = ph.temperature(uctr[0], uctr[3]);
	}
	else {
		ph.primitive2FromConserved(ucl, uctl);
		ph.primitive2FromConserved(ucr, uctr);
	}
}

/// spatial/aspatial.cpp:173-205
static inline void face_gradient_modified_average(const double *rcl, const double *rcr,
                                                  const double *ucl, const double *ucr,
                                                  const double *gradl, const double *gradr,
                                                  double grad[NDIM][NVARS])
{
	double dr[NDIM], dist=0;
	for(int i = 0; i < NDIM; i++) {
		dr[i] = rcr[i]-rcl[i];
		dist += dr[i]*dr[i];
	}
	dist = std::sqrt(dist);
	for(int i = 0; i < NDIM; i++)
		dr[i] /= dist;

	for(int i = 0; i < NVARS; i++) {
		double davg[NDIM];
		for(int j = 0; j < NDIM; j++)
			davg[j] = 0.5*(gradl[j*NVARS+i] + gradr[j*NVARS+i]);
		const double corr = (ucr[i]-ucl[i])/dist;
		const double ddr = dot2(davg,dr);
		for(int j = 0; j < NDIM; j++)
			grad[j][i] = davg[j] - ddr*dr[j] + corr*dr[j];
	}
}

/// physics/viscousphysics.cpp:71-122
static inline void viscous_flux(const Gas &ph, bool constVisc, const double *n,
                                const double grad[NDIM][NVARS], const double *ul, const double *ur,
                                double *vflux)
{
	const double muRe = constVisc ? ph.constantViscosity()
		: 0.5*( ph.viscosityFromConserved(ul) + ph.viscosityFromConserved(ur) );
	const double kdiff = ph.thermalConductivity(muRe);

	double stress[NDIM][NDIM];
	ph.stressTensor(muRe, grad, stress);

	vflux[0] = 0;
	for(int i = 0; i < NDIM; i++) {
		vflux[i+1] = 0;
		for(int j = 0; j < NDIM; j++)
			vflux[i+1] -= stress[i][j] * n[j];
	}

	double vavg[NDIM];
	for(int j = 0; j < NDIM; j++)
		vavg[j] = 0.5*( ul[j+1]/ul[0] + ur[j+1]/ur[0] );

	vflux[3] = 0;
	for(int i = 0; i < NDIM; i++) {
		double comp = 0;
		for(int j = 0; j < NDIM; j++)
			comp += stress[i][j]*vavg[j];
		comp += kdiff*grad[i][3];
		vflux[3] -= comp * n[i];
	}
}

/// spatial/flow_spatial.cpp:349-395: glue. gradsl/gradsr are GradBlock_t memory order
/// (col-major 2x4: index idim + 2*ivar); may be null when !order2.
static inline void cell_viscous_flux(const Gas &ph, bool order2, bool constVisc, const double *normal,
                                     const double *rcl, const double *rcr,
                                     const double *ucell_l, const double *ucell_r,
                                     const double *gradsl, const double *gradsr,
                                     const double *ul, const double *ur, double *vflux)
{
	double uctl[NVARS], uctr[NVARS];
	double gradl[NDIM*NVARS], gradr[NDIM*NVARS];
	if(order2) {
		for(int i = 0; i < NDIM; i++)
			for(int j = 0; j < NVARS; j++) {
				gradl[i*NVARS+j] = gradsl[i+NDIM*j];
				gradr[i*NVARS+j] = gradsr[i+NDIM*j];
			}
	}
	primitive2_states_and_gradients(ph, order2, ucell_l, ucell_r, gradl, gradr, uctl, uctr);
	if(!order2) {
		for(int k = 0; k < NDIM*NVARS; k++) { gradl[k] = 0; gradr[k] = 0; }
	}
	double grad[NDIM][NVARS];
	face_gradient_modified_average(rcl, rcr, uctl, uctr, gradl, gradr, grad);
	viscous_flux(ph, constVisc, normal, grad, ul, ur, vflux);
}

} // namespace orc
#endif
