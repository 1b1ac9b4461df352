/* Device-side ideal-gas physics, numerical fluxes, boundary ghost states and the viscous flux
 * for the FVENS residual path. Device code only: there is deliberately no host fallback.
 *
 * What each piece restates (paths relative to the reference's src/):
 *   Side / load_side            physics/aphysics_defs.hpp:27-38, 140-163 (getVarsFromConserved, sound speed)
 *   flux<...>                   spatial/anumericalflux.cpp:41-61 (LLF), 203-250 (Van Leer), 265-315 (AUSM),
 *                               480-553 (AUSM+), 668-732 (Roe-Pike + Harten fix), 974-1007 (HLL),
 *                               1071-1081 + 1176-1228 (HLLC); Roe averages spatial/anumericalflux.hpp:175-189
 *   ghost_state                 spatial/abc.cpp:49-84, 152-176, 194-199, 218-226, 272-280, 354-369, 417-423
 *   viscous_face_flux           physics/viscousphysics.cpp:15-122, spatial/aspatial.cpp:173-205,
 *                               spatial/flow_spatial.cpp:349-395
 * The arithmetic is arranged for the GPU (shared reciprocals, no virtual dispatch, compile-time
 * flux selection); parity with the reference is to round-off (tests/ gate it at 1e-12 relative).
 */
#pragma once
#include <cuda_runtime.h>

namespace fvg {

enum FluxId { FLUX_LLF = 0, FLUX_VANLEER = 1, FLUX_AUSM = 2, FLUX_AUSMPLUS = 3, FLUX_ROE = 4,
              FLUX_HLL = 5, FLUX_HLLC = 6, FLUX_COUNT = 7 };

/// Same numbering as the reference enum (spatial/abctypes.hpp:13-22)
enum BCType { SLIP_WALL_BC = 0, FARFIELD_BC = 1, INFLOW_OUTFLOW_BC = 2, SUBSONIC_INFLOW_BC = 3,
              EXTRAPOLATION_BC = 4, PERIODIC_BC = 5, ISOTHERMAL_WALL_BC = 6, ADIABATIC_WALL_BC = 7 };

constexpr int MAX_BC = 16;

struct BCEntry {
	int tag;
	int type;
	double v0, v1;
};

/// Everything the kernels need to know about the gas and the boundary conditions; passed by value.
struct GasParams {
	double g;        ///< adiabatic index
	double Minf, Tinf, Reinf, Pr;
	double sCT;      ///< Sutherland constant over Tinf (110.5/Tinf)
	double gm1;      ///< g-1
	double igm1;     ///< 1/(g-1)
	double gM2;      ///< g*Minf^2
	double pinf;     ///< 1/(g Minf^2)
	double uinf[4];
	double limiter_param;
	int nbc;
	BCEntry bc[MAX_BC];
};

// ---------------------------------------------------------------------------------------------
// FP64 reciprocal / square root without the IEEE slow paths: the SFU seed (rcp.approx / rsqrt.approx,
// relative error < 2^-22) refined by two Newton steps in FMA arithmetic. Results are within ~1 ulp of
// the correctly rounded value; every use is on strictly positive, normal-range arguments (densities,
// pressures, distances), so no special-case branches are needed. This is what keeps the hot kernels
// free of the BSSY/BSYNC-wrapped subroutine calls that `/` and sqrt() compile to.

__device__ __forceinline__ double frcp(double x) {
	double y;
	asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
	double e = fma(-x, y, 1.0);
	y = fma(y, e, y);
	e = fma(-x, y, 1.0);
	return fma(y, e, y);
}
__device__ __forceinline__ double frsqrt(double x) {
	double y;
	asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
	const double h = 0.5*x;
	double e = fma(-h*y, y, 0.5);
	y = fma(y, e, y);
	e = fma(-h*y, y, 0.5);
	return fma(y, e, y);
}
/// sqrt(x) for x > 0: x * rsqrt(x) with one Heron correction. The correction squares the error of its input,
/// so ONE Newton step on the SFU seed (2^-22 -> 2^-43) is enough to reach the last bit afterwards.
__device__ __forceinline__ double fsqrt(double x) {
	double y;
	asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
	const double h = 0.5*x;
	const double e = fma(-h*y, y, 0.5);
	y = fma(y, e, y);
	const double s = x*y;
	return fma(0.5*y, fma(-s, s, x), s);
}
/// sqrt(x) and 1/x for x > 0 from one SFU seed: r = rsqrt(x) to full precision, then sqrt = x*r (corrected), 1/x = r*r
__device__ __forceinline__ void fsqrt_rcp(double x, double &sq, double &rc) {
	const double r = frsqrt(x);
	const double s = x*r;
	sq = fma(0.5*r, fma(-s, s, x), s);
	rc = r*r;
}

// ---------------------------------------------------------------------------------------------
// basic gas relations on conserved variables u = (rho, rho vx, rho vy, rho E)

__device__ __forceinline__ double pressure_cons(const GasParams &G, const double u[4]) {
	return G.gm1*(u[3] - 0.5*(u[1]*u[1] + u[2]*u[2])/u[0]);
}
/// Written with explicit fma / non-contractable operations: the same cell is converted in different
/// places (as a tile cell, as a halo cell of the neighbouring tile, as a ghost on another GPU) and all of
/// them must produce the same bits, whatever the compiler would contract in each context.
__device__ __forceinline__ void cons2prim(const GasParams &G, const double u[4], double p[4]) {
	const double ir = frcp(u[0]);
	const double m2 = fma(u[2], u[2], __dmul_rn(u[1], u[1]));
	const double ie = fma(__dmul_rn(-0.5, m2), ir, u[3]);
	p[0] = u[0]; p[1] = __dmul_rn(u[1], ir); p[2] = __dmul_rn(u[2], ir); p[3] = __dmul_rn(G.gm1, ie);
}
/// u_face = u_cell + g . (mid - centre) for the four primitive variables (reconstruction_utils.hpp:17-32),
/// gradients in GradBlock order split in two rows of four; explicit fma order for the same reason as above
__device__ __forceinline__ void extrapolate_prim(const double pc[4], const double ga[4], const double gb[4],
                                                 double midx, double midy, double rcx, double rcy, double pf[4]) {
	const double dx = __dsub_rn(midx, rcx), dy = __dsub_rn(midy, rcy);
	pf[0] = fma(ga[1], dy, fma(ga[0], dx, pc[0]));
	pf[1] = fma(ga[3], dy, fma(ga[2], dx, pc[1]));
	pf[2] = fma(gb[1], dy, fma(gb[0], dx, pc[2]));
	pf[3] = fma(gb[3], dy, fma(gb[2], dx, pc[3]));
}
__device__ __forceinline__ void prim2cons(const GasParams &G, const double p[4], double u[4]) {
	const double e = p[3]*G.igm1 + 0.5*p[0]*(p[1]*p[1] + p[2]*p[2]);
	u[0] = p[0]; u[1] = p[0]*p[1]; u[2] = p[0]*p[2]; u[3] = e;
}
__device__ __forceinline__ double temperature(const GasParams &G, double rho, double p) {
	return p/rho*G.gM2;
}
__device__ __forceinline__ double sutherland(const GasParams &G, double T) {
	// (1 + C/Tinf)/(T + C/Tinf) * T^1.5 / Re
	return (1.0 + G.sCT)/(T + G.sCT)*(T*sqrt(T))/G.Reinf;
}
__device__ __forceinline__ double viscosity_cons(const GasParams &G, const double u[4]) {
	return sutherland(G, temperature(G, u[0], pressure_cons(G, u)));
}

/// One side of a face: everything the fluxes and the spectral radius need.
struct Side {
	double r, mx, my, E;  ///< conserved
	double vx, vy, vn;    ///< velocity and its normal component
	double p, H, c;       ///< pressure, total specific enthalpy, sound speed
	double ir;            ///< 1/rho
};

template <bool WITH_C>
__device__ __forceinline__ Side load_side(const GasParams &G, const double u[4], double nx, double ny) {
	Side s;
	s.r = u[0]; s.mx = u[1]; s.my = u[2]; s.E = u[3];
	const double ir = frcp(u[0]);
	s.ir = ir;
	s.vx = u[1]*ir; s.vy = u[2]*ir;
	s.vn = s.vx*nx + s.vy*ny;
	s.p = G.gm1*(u[3] - 0.5*u[0]*(s.vx*s.vx + s.vy*s.vy));
	s.H = (u[3] + s.p)*ir;
	s.c = WITH_C ? fsqrt(G.g*s.p*ir) : 0.0;
	return s;
}

/// Side straight from a primitive state (rho, vx, vy, p): skips the prim->cons->prim round trip the
/// reference makes between reconstruction and flux (flow_spatial.cpp:740-754); differs at round-off only.
template <bool WITH_C>
__device__ __forceinline__ Side side_from_prim(const GasParams &G, const double p[4], double nx, double ny) {
	Side s;
	s.r = p[0]; s.vx = p[1]; s.vy = p[2]; s.p = p[3];
	s.mx = p[0]*p[1]; s.my = p[0]*p[2];
	s.E = p[3]*G.igm1 + 0.5*p[0]*(p[1]*p[1] + p[2]*p[2]);
	s.vn = p[1]*nx + p[2]*ny;
	const double ir = frcp(p[0]);
	s.ir = ir;
	s.H = (s.E + s.p)*ir;
	s.c = WITH_C ? fsqrt(G.g*s.p*ir) : 0.0;
	return s;
}

__device__ __forceinline__ void normal_flux(const Side &s, double nx, double ny, double f[4]) {
	f[0] = s.vn*s.r;
	f[1] = s.vn*s.mx + s.p*nx;
	f[2] = s.vn*s.my + s.p*ny;
	f[3] = s.vn*(s.E + s.p);
}

struct RoeAvg { double R, rho, vx, vy, vm2, vn, H, c, ic2; };

__device__ __forceinline__ RoeAvg roe_average(const GasParams &G, const Side &a, const Side &b,
                                              double nx, double ny) {
	RoeAvg q;
	q.R = fsqrt(b.r*a.ir);
	q.rho = q.R*a.r;
	const double iw = frcp(q.R + 1.0);
	q.vx = (q.R*b.vx + a.vx)*iw;
	q.vy = (q.R*b.vy + a.vy)*iw;
	q.H = (q.R*b.H + a.H)*iw;
	q.vm2 = q.vx*q.vx + q.vy*q.vy;
	q.vn = q.vx*nx + q.vy*ny;
	fsqrt_rcp(G.gm1*(q.H - 0.5*q.vm2), q.c, q.ic2);
	return q;
}

template <int FLUX> struct FluxTraits { static constexpr bool needs_c = true; };
template <> struct FluxTraits<FLUX_ROE> { static constexpr bool needs_c = false; };

/// Numerical flux through a face with unit normal (nx,ny), from the two sides' data.
template <int FLUX>
__device__ __forceinline__ void flux_from_sides(const GasParams &G, const Side &a, const Side &b,
                                                double nx, double ny, double f[4])
{
	if(FLUX == FLUX_LLF) {
		const double ea = fabs(a.vn) + a.c, eb = fabs(b.vn) + b.c;
		const double eig = ea > eb ? ea : eb;
		double fa[4], fb[4];
		normal_flux(a, nx, ny, fa);
		normal_flux(b, nx, ny, fb);
		f[0] = 0.5*(fa[0] + fb[0] - eig*(b.r - a.r));
		f[1] = 0.5*(fa[1] + fb[1] - eig*(b.mx - a.mx));
		f[2] = 0.5*(fa[2] + fb[2] - eig*(b.my - a.my));
		f[3] = 0.5*(fa[3] + fb[3] - eig*(b.E - a.E));
	}
	else if(FLUX == FLUX_VANLEER) {
		const double g = G.g;
		const double Ma = a.vn*frcp(a.c), Mb = b.vn*frcp(b.c);
		const double ie = 1.0/(2.0*(g*g - 1.0));
		double fp[4], fm[4];
		if(Ma < -1.0) { fp[0] = fp[1] = fp[2] = fp[3] = 0.0; }
		else if(Ma > 1.0) normal_flux(a, nx, ny, fp);
		else {
			const double vm2 = a.vx*a.vx + a.vy*a.vy;
			const double t = (2.0*a.c - a.vn)*(1.0/g);
			const double w = G.gm1*a.vn + 2.0*a.c;
			fp[0] = a.r*a.c*(Ma + 1.0)*(Ma + 1.0)*0.25;
			fp[1] = fp[0]*(a.vx + nx*t);
			fp[2] = fp[0]*(a.vy + ny*t);
			fp[3] = fp[0]*((vm2 - a.vn*a.vn)*0.5 + w*w*ie);
		}
		if(Mb > 1.0) { fm[0] = fm[1] = fm[2] = fm[3] = 0.0; }
		else if(Mb < -1.0) normal_flux(b, nx, ny, fm);
		else {
			const double vm2 = b.vx*b.vx + b.vy*b.vy;
			const double t = (-2.0*b.c - b.vn)*(1.0/g);
			const double w = G.gm1*b.vn - 2.0*b.c;
			fm[0] = -b.r*b.c*(Mb - 1.0)*(Mb - 1.0)*0.25;
			fm[1] = fm[0]*(b.vx + nx*t);
			fm[2] = fm[0]*(b.vy + ny*t);
			fm[3] = fm[0]*((vm2 - b.vn*b.vn)*0.5 + w*w*ie);
		}
		for(int k = 0; k < 4; k++) f[k] = fp[k] + fm[k];
	}
	else if(FLUX == FLUX_AUSM) {
		const double Ma = a.vn*frcp(a.c), Mb = b.vn*frcp(b.c);
		double ML, MR, pL, pR;
		if(fabs(Ma) <= 1.0) { ML = 0.25*(Ma + 1.0)*(Ma + 1.0); pL = ML*a.p*(2.0 - Ma); }
		else if(Ma < -1.0) { ML = 0.0; pL = 0.0; }
		else { ML = Ma; pL = a.p; }
		if(fabs(Mb) <= 1.0) { MR = -0.25*(Mb - 1.0)*(Mb - 1.0); pR = -MR*b.p*(2.0 + Mb); }
		else if(Mb < -1.0) { MR = Mb; pR = b.p; }
		else { MR = 0.0; pR = 0.0; }
		const double Mh = 0.5*(ML + MR), aM = fabs(Mh), ph = pL + pR;
		f[0] = Mh*(a.r*a.c + b.r*b.c) - aM*(b.r*b.c - a.r*a.c);
		f[1] = Mh*(a.mx*a.c + b.mx*b.c) - aM*(b.mx*b.c - a.mx*a.c) + ph*nx;
		f[2] = Mh*(a.my*a.c + b.my*b.c) - aM*(b.my*b.c - a.my*a.c) + ph*ny;
		const double ha = a.c*(a.E + a.p), hb = b.c*(b.E + b.p);
		f[3] = Mh*(ha + hb) - aM*(hb - ha);
	}
	else if(FLUX == FLUX_AUSMPLUS) {
		const double g = G.g;
		const double k = 2.0*G.gm1/(g + 1.0);
		const double vm2a = a.vx*a.vx + a.vy*a.vy, vm2b = b.vx*b.vx + b.vy*b.vy;
		const double igm1 = 1.0/G.gm1;
		const double cs2a = (a.c*a.c*igm1 + 0.5*vm2a)*k;
		const double cs2b = (b.c*b.c*igm1 + 0.5*vm2b)*k;
		const double csa = fsqrt(cs2a), csb = fsqrt(cs2b);
		const double corra = csa > a.vn ? csa : a.vn;
		const double corrb = csb > -b.vn ? csb : -b.vn;
		const double cta = csa*csa*frcp(corra), ctb = csb*csb*frcp(corrb);
		const double ch = cta < ctb ? cta : ctb;
		const double ich = frcp(ch);
		const double Ma = a.vn*ich, Mb = b.vn*ich;
		double ML, MR, pL, pR;
		if(fabs(Ma) <= 1.0) {
			const double q = (Ma*Ma - 1.0)*(Ma*Ma - 1.0);
			ML = 0.25*(Ma + 1.0)*(Ma + 1.0) + 0.125*q;
			pL = a.p*(0.25*(Ma + 1.0)*(Ma + 1.0)*(2.0 - Ma) + 0.1875*Ma*q);
		}
		else if(Ma < -1.0) { ML = 0.0; pL = 0.0; }
		else { ML = Ma; pL = a.p; }
		if(fabs(Mb) <= 1.0) {
			const double q = (Mb*Mb - 1.0)*(Mb*Mb - 1.0);
			MR = -0.25*(Mb - 1.0)*(Mb - 1.0) - 0.125*q;
			pR = b.p*(0.25*(Mb - 1.0)*(Mb - 1.0)*(2.0 + Mb) - 0.1875*Mb*q);
		}
		else if(Mb < -1.0) { MR = Mb; pR = b.p; }
		else { MR = 0.0; pR = 0.0; }
		const double Mh = 0.5*(ML + MR), aM = fabs(Mh), ph = pL + pR;
		f[0] = ch*(Mh*(a.r + b.r) - aM*(b.r - a.r));
		f[1] = ch*(Mh*(a.mx + b.mx) - aM*(b.mx - a.mx)) + ph*nx;
		f[2] = ch*(Mh*(a.my + b.my) - aM*(b.my - a.my)) + ph*ny;
		f[3] = ch*(Mh*(a.E + a.p + b.E + b.p) - aM*((b.E + b.p) - (a.E + a.p)));
	}
	else if(FLUX == FLUX_ROE) {
		const RoeAvg q = roe_average(G, a, b, nx, ny);
		double l0 = fabs(q.vn - q.c), l1 = fabs(q.vn), l3 = fabs(q.vn + q.c);
		const double delta = 1.0e-4*q.c;
		if(l0 < delta || l1 < delta || l3 < delta) {
			const double i2d = frcp(2.0*delta), d2 = delta*delta;
			if(l0 < delta) l0 = (l0*l0 + d2)*i2d;
			if(l1 < delta) l1 = (l1*l1 + d2)*i2d;
			if(l3 < delta) l3 = (l3*l3 + d2)*i2d;
		}
		const double dvn = b.vn - a.vn, dp = b.p - a.p, dr = b.r - a.r;
		const double dvx = b.vx - a.vx, dvy = b.vy - a.vy;
		const double ic2 = q.ic2;
		const double rc = q.rho*q.c;
		const double a0 = l0*(dp - rc*dvn)*(0.5*ic2);
		const double a1 = l1*(dr - dp*ic2);
		const double a2 = l1*q.rho;
		const double a3 = l3*(dp + rc*dvn)*(0.5*ic2);
		double d0 = a0 + a1 + a3;
		double d1 = a0*(q.vx - q.c*nx) + a1*q.vx + a2*(dvx - dvn*nx) + a3*(q.vx + q.c*nx);
		double d2_ = a0*(q.vy - q.c*ny) + a1*q.vy + a2*(dvy - dvn*ny) + a3*(q.vy + q.c*ny);
		double d3 = a0*(q.H - q.c*q.vn) + a1*(0.5*q.vm2) + a2*(q.vx*dvx + q.vy*dvy - q.vn*dvn)
		            + a3*(q.H + q.c*q.vn);
		double fa[4], fb[4];
		normal_flux(a, nx, ny, fa);
		normal_flux(b, nx, ny, fb);
		f[0] = 0.5*(fa[0] + fb[0] - d0);
		f[1] = 0.5*(fa[1] + fb[1] - d1);
		f[2] = 0.5*(fa[2] + fb[2] - d2_);
		f[3] = 0.5*(fa[3] + fb[3] - d3);
	}
	else if(FLUX == FLUX_HLL) {
		const RoeAvg q = roe_average(G, a, b, nx, ny);
		double sl = a.vn - a.c; if(sl > q.vn - q.c) sl = q.vn - q.c;
		double sr = b.vn + b.c; if(sr < q.vn + q.c) sr = q.vn + q.c;
		const double sr0 = sr > 0.0 ? 0.0 : sr;
		const double sl0 = sl > 0.0 ? 0.0 : sl;
		const double is = frcp(sr - sl);
		const double t1 = (sr0 - sl0)*is, t2 = 1.0 - t1;
		const double t3 = 0.5*(sr*fabs(sl) - sl*fabs(sr))*is;
		f[0] = t1*b.vn*b.r + t2*a.vn*a.r - t3*(b.r - a.r);
		f[1] = t1*(b.vn*b.mx + b.p*nx) + t2*(a.vn*a.mx + a.p*nx) - t3*(b.mx - a.mx);
		f[2] = t1*(b.vn*b.my + b.p*ny) + t2*(a.vn*a.my + a.p*ny) - t3*(b.my - a.my);
		f[3] = t1*(b.vn*b.r*b.H) + t2*(a.vn*a.r*a.H) - t3*(b.E - a.E);
	}
	else { // HLLC
		const RoeAvg q = roe_average(G, a, b, nx, ny);
		double sl = a.vn - a.c; if(sl > q.vn - q.c) sl = q.vn - q.c;
		double sr = b.vn + b.c; if(sr < q.vn + q.c) sr = q.vn + q.c;
		const double mb = b.r*(sr - b.vn), ma = a.r*(sl - a.vn);
		const double sm = (mb*b.vn - ma*a.vn + a.p - b.p)*frcp(mb - ma);
		if(sl > 0.0) normal_flux(a, nx, ny, f);
		else if(sm > 0.0) {
			normal_flux(a, nx, ny, f);
			const double pst = a.r*(a.vn - sl)*(a.vn - sm) + a.p;
			const double k = frcp(sl - sm);
			const double w = sl - a.vn;
			f[0] += sl*(a.r*w*k - a.r);
			f[1] += sl*((w*a.mx + (pst - a.p)*nx)*k - a.mx);
			f[2] += sl*((w*a.my + (pst - a.p)*ny)*k - a.my);
			f[3] += sl*((w*a.E - a.p*a.vn + pst*sm)*k - a.E);
		}
		else if(sr >= 0.0) {
			normal_flux(b, nx, ny, f);
			const double pst = b.r*(b.vn - sr)*(b.vn - sm) + b.p;
			const double k = frcp(sr - sm);
			const double w = sr - b.vn;
			f[0] += sr*(b.r*w*k - b.r);
			f[1] += sr*((w*b.mx + (pst - b.p)*nx)*k - b.mx);
			f[2] += sr*((w*b.my + (pst - b.p)*ny)*k - b.my);
			f[3] += sr*((w*b.E - b.p*b.vn + pst*sm)*k - b.E);
		}
		else normal_flux(b, nx, ny, f);
	}
}

/// Convenience: flux straight from conserved states (pointwise test hook and generic callers).
template <int FLUX>
__device__ __forceinline__ void inviscid_flux(const GasParams &G, const double ul[4], const double ur[4],
                                              double nx, double ny, double f[4])
{
	const Side a = load_side<FluxTraits<FLUX>::needs_c>(G, ul, nx, ny);
	const Side b = load_side<FluxTraits<FLUX>::needs_c>(G, ur, nx, ny);
	flux_from_sides<FLUX>(G, a, b, nx, ny, f);
}

// ---------------------------------------------------------------------------------------------
// boundary ghost states

/// The few gas constants the boundary conditions need, passed BY VALUE to the out-of-line routine below
/// (passing the kernel-parameter struct by reference would make the compiler copy it to the stack).
struct BCGas { double g, gm1, gM2, pinf, Minf; double u0, u1, u2, u3; };

/// Out of line on purpose: boundary faces are O(sqrt(N)), and inlining the eight-way switch (with its pow,
/// sqrt and divisions) into the face and cell kernels costs every interior face registers and code size.
static __device__ __noinline__ double4 ghost_state_ool(const BCGas B, const int type, const double v0, const double v1,
                                                       const double4 in4, const double nx, const double ny)
{
	const double ins[4] = {in4.x, in4.y, in4.z, in4.w};
	double gs[4];
	switch(type) {
	case INFLOW_OUTFLOW_BC: {
		const double ir = 1.0/ins[0];
		const double vn = (ins[1]*nx + ins[2]*ny)*ir;
		const double m2 = ins[1]*ins[1] + ins[2]*ins[2];
		const double p = B.gm1*(ins[3] - 0.5*m2*ir);
		const double c = sqrt(B.g*p*ir);
		const double Mn = vn/c;
		if(Mn <= 0.0) { gs[0] = B.u0; gs[1] = B.u1; gs[2] = B.u2; gs[3] = B.u3; }
		else if(Mn < 1.0) {
			gs[0] = ins[0]; gs[1] = ins[1]; gs[2] = ins[2];
			gs[3] = B.pinf/B.gm1 + 0.5*ins[0]*(m2/(ins[0]*ins[0]));
		}
		else { gs[0] = ins[0]; gs[1] = ins[1]; gs[2] = ins[2]; gs[3] = ins[3]; }
		break;
	}
	case SUBSONIC_INFLOW_BC: {
		const double g = B.g, ptot = v0, ttot = v1;
		const double ir = 1.0/ins[0];
		const double m2 = ins[1]*ins[1] + ins[2]*ins[2];
		const double p = B.gm1*(ins[3] - 0.5*m2*ir);
		const double ci = sqrt(g*p*ir);
		const double Rm = (ins[1]*nx + ins[2]*ny)*ir - ci/(2.0*g - 1.0);
		const double co2 = ci*ci + 0.5*B.gm1*m2/(ins[0]*ins[0]);
		const double q = sqrt((g + 1.0)*co2/(B.gm1*Rm*Rm) - 0.5*B.gm1);
		const double cg = -Rm*B.gm1/(g + 1.0)*(1.0 + q);
		const double tg = ttot*cg*cg/co2;
		const double pg = ptot*pow(tg/ttot, g/B.gm1);
		gs[0] = B.gM2*pg/tg;
		const double vg = sqrt(2.0/B.gm1*(co2 - cg*cg));
		gs[1] = gs[0]*(vg*nx);
		gs[2] = gs[0]*(vg*ny);
		gs[3] = pg/B.gm1 + 0.5*gs[0]*(vg*vg);
		break;
	}
	case FARFIELD_BC:
		gs[0] = B.u0; gs[1] = B.u1; gs[2] = B.u2; gs[3] = B.u3;
		break;
	case SLIP_WALL_BC: {
		const double vn = (ins[1]*nx + ins[2]*ny)/ins[0];
		gs[0] = ins[0];
		gs[1] = ins[1] - 2.0*vn*nx*ins[0];
		gs[2] = ins[2] - 2.0*vn*ny*ins[0];
		gs[3] = ins[3];
		break;
	}
	case ADIABATIC_WALL_BC: {
		const double tm = v0*ins[0];
		gs[0] = ins[0];
		gs[1] = 2.0*tm*ny - ins[1];
		gs[2] = -2.0*tm*nx - ins[2];
		gs[3] = ins[3];
		break;
	}
	case ISOTHERMAL_WALL_BC: {
		const double vt = v0, Tw = v1;
		const double ir = 1.0/ins[0];
		const double p = B.gm1*(ins[3] - 0.5*(ins[1]*ins[1] + ins[2]*ins[2])*ir);
		const double Tg = 2.0*Tw - p*ir*B.gM2;
		gs[0] = ins[0];
		gs[1] = gs[0]*(2.0*vt*ny - ins[1]*ir);
		gs[2] = gs[0]*(-2.0*vt*nx - ins[2]*ir);
		const double vm2 = (gs[1]*gs[1] + gs[2]*gs[2])/(gs[0]*gs[0]);
		gs[3] = gs[0]*(Tg/(B.g*B.gm1*B.Minf*B.Minf) + 0.5*vm2);
		break;
	}
	default: // EXTRAPOLATION_BC (and anything unknown is rejected at flow creation)
		gs[0] = ins[0]; gs[1] = ins[1]; gs[2] = ins[2]; gs[3] = ins[3];
	}
	return make_double4(gs[0], gs[1], gs[2], gs[3]);
}

__device__ __forceinline__ void ghost_state(const GasParams &G, const BCEntry &bc, const double ins[4],
                                            double nx, double ny, double gs[4])
{
	const BCGas B = {G.g, G.gm1, G.gM2, G.pinf, G.Minf, G.uinf[0], G.uinf[1], G.uinf[2], G.uinf[3]};
	const double4 r = ghost_state_ool(B, bc.type, bc.v0, bc.v1, make_double4(ins[0], ins[1], ins[2], ins[3]), nx, ny);
	gs[0] = r.x; gs[1] = r.y; gs[2] = r.z; gs[3] = r.w;
}

// ---------------------------------------------------------------------------------------------
// viscous flux (modified-average face gradient)

enum ViscMode { VISC_NONE = 0, VISC_CONST = 1, VISC_SUTHERLAND = 2 };

/** Viscous flux through a face.
 * ucl,ucr : conserved cell-centre states either side (right = boundary ghost state on a boundary face)
 * gl, gr  : cell gradients of the primitive variables (rho, vx, vy, p) in GradBlock order
 *           g[idim + 2*ivar]; ignored when !ORDER2 (treated as zero, flow_spatial.cpp:383-389)
 * ul, ur  : conserved face states (used for the viscosity and the average velocity)
 */
template <bool ORDER2, bool CONSTVISC>
__device__ __forceinline__ void viscous_face_flux(const GasParams &G, double nx, double ny,
                                                  double rclx, double rcly, double rcrx, double rcry,
                                                  const double ucl[4], const double ucr[4],
                                                  const double gl[8], const double gr[8],
                                                  const double ul[4], const double ur[4], double vf[4])
{
	// cell states in (rho, vx, vy, T)
	double tl[4], tr[4];
	cons2prim(G, ucl, tl);
	cons2prim(G, ucr, tr);
	double dl[2][4], dr_[2][4];   // [dim][var] gradients of (rho, vx, vy, T)
	if(ORDER2) {
		for(int d = 0; d < 2; d++) {
			for(int v = 0; v < 3; v++) { dl[d][v] = gl[d + 2*v]; dr_[d][v] = gr[d + 2*v]; }
			// grad T from grad p and grad rho
			dl[d][3] = (gl[d + 6]*tl[0] - tl[3]*gl[d])/(tl[0]*tl[0])*G.gM2;
			dr_[d][3] = (gr[d + 6]*tr[0] - tr[3]*gr[d])/(tr[0]*tr[0])*G.gM2;
		}
	} else {
		for(int d = 0; d < 2; d++) for(int v = 0; v < 4; v++) { dl[d][v] = 0.0; dr_[d][v] = 0.0; }
	}
	tl[3] = temperature(G, tl[0], tl[3]);
	tr[3] = temperature(G, tr[0], tr[3]);

	// modified average
	double ex = rcrx - rclx, ey = rcry - rcly;
	const double dist = sqrt(ex*ex + ey*ey);
	ex /= dist; ey /= dist;
	double gf[2][4];
	for(int v = 0; v < 4; v++) {
		const double ax = 0.5*(dl[0][v] + dr_[0][v]), ay = 0.5*(dl[1][v] + dr_[1][v]);
		const double corr = (tr[v] - tl[v])/dist;
		const double ddr = ax*ex + ay*ey;
		gf[0][v] = ax - ddr*ex + corr*ex;
		gf[1][v] = ay - ddr*ey + corr*ey;
	}

	const double mu = CONSTVISC ? 1.0/G.Reinf : 0.5*(viscosity_cons(G, ul) + viscosity_cons(G, ur));
	const double kd = mu/(G.Minf*G.Minf*G.gm1*G.Pr);

	const double ldiv = (gf[0][1] + gf[1][2])*(2.0/3.0*mu);
	const double sxx = mu*(gf[0][1] + gf[0][1]) - ldiv;
	const double sxy = mu*(gf[0][2] + gf[1][1]);
	const double syy = mu*(gf[1][2] + gf[1][2]) - ldiv;

	const double vax = 0.5*(ul[1]/ul[0] + ur[1]/ur[0]);
	const double vay = 0.5*(ul[2]/ul[0] + ur[2]/ur[0]);

	vf[0] = 0.0;
	vf[1] = -(sxx*nx + sxy*ny);
	vf[2] = -(sxy*nx + syy*ny);
	vf[3] = -((sxx*vax + sxy*vay + kd*gf[0][3])*nx + (sxy*vax + syy*vay + kd*gf[1][3])*ny);
}

/** The same viscous flux for the fused face kernel: every division and square root of the formula above is replaced by
 * the branch-free reciprocals of this file (about 35 IEEE divisions and 5 square roots per face otherwise - they made the
 * viscous face pass four times as long as the inviscid one), shared sub-expressions are computed once (1/rho of each
 * cell, 1/|d|, the two face viscosities, which the spectral radius needs too). Results agree with viscous_face_flux to a
 * few ulps. mu_l, mu_r: viscosities of the two FACE states (constant-viscosity flows pass 1/Re for both). */
template <bool ORDER2>
__device__ __forceinline__ void viscous_face_flux_fast(const GasParams &G, double nx, double ny,
                                                       double rclx, double rcly, double rcrx, double rcry,
                                                       const double ucl[4], const double ucr[4],
                                                       const double gl[8], const double gr[8],
                                                       double vlx, double vly, double vrx, double vry,
                                                       double mu_l, double mu_r, double vf[4])
{
	// cell states in (rho, vx, vy, T)
	const double irl = frcp(ucl[0]), irr = frcp(ucr[0]);
	double tl[4], tr[4];
	tl[0] = ucl[0]; tl[1] = ucl[1]*irl; tl[2] = ucl[2]*irl;
	tr[0] = ucr[0]; tr[1] = ucr[1]*irr; tr[2] = ucr[2]*irr;
	const double pl = G.gm1*(ucl[3] - 0.5*(ucl[1]*ucl[1] + ucl[2]*ucl[2])*irl);
	const double pr = G.gm1*(ucr[3] - 0.5*(ucr[1]*ucr[1] + ucr[2]*ucr[2])*irr);
	double dl[2][4], dr_[2][4];   // [dim][var] gradients of (rho, vx, vy, T)
	if(ORDER2) {
		const double sl = irl*irl*G.gM2, sr = irr*irr*G.gM2;
		#pragma unroll
		for(int d = 0; d < 2; d++) {
			#pragma unroll
			for(int v = 0; v < 3; v++) { dl[d][v] = gl[d + 2*v]; dr_[d][v] = gr[d + 2*v]; }
			// grad T from grad p and grad rho
			dl[d][3] = (gl[d + 6]*tl[0] - pl*gl[d])*sl;
			dr_[d][3] = (gr[d + 6]*tr[0] - pr*gr[d])*sr;
		}
	} else {
		for(int d = 0; d < 2; d++) for(int v = 0; v < 4; v++) { dl[d][v] = 0.0; dr_[d][v] = 0.0; }
	}
	tl[3] = pl*irl*G.gM2;
	tr[3] = pr*irr*G.gM2;

	// modified average
	double ex = rcrx - rclx, ey = rcry - rcly;
	const double idist = frsqrt(ex*ex + ey*ey);
	ex *= idist; ey *= idist;
	double gf[2][4];
	#pragma unroll
	for(int v = 0; v < 4; v++) {
		const double ax = 0.5*(dl[0][v] + dr_[0][v]), ay = 0.5*(dl[1][v] + dr_[1][v]);
		const double corr = (tr[v] - tl[v])*idist;
		const double ddr = ax*ex + ay*ey;
		gf[0][v] = ax - ddr*ex + corr*ex;
		gf[1][v] = ay - ddr*ey + corr*ey;
	}

	const double mu = 0.5*(mu_l + mu_r);
	const double kd = mu*frcp(G.Minf*G.Minf*G.gm1*G.Pr);

	const double ldiv = (gf[0][1] + gf[1][2])*(2.0/3.0*mu);
	const double sxx = mu*(gf[0][1] + gf[0][1]) - ldiv;
	const double sxy = mu*(gf[0][2] + gf[1][1]);
	const double syy = mu*(gf[1][2] + gf[1][2]) - ldiv;

	const double vax = 0.5*(vlx + vrx);
	const double vay = 0.5*(vly + vry);

	vf[0] = 0.0;
	vf[1] = -(sxx*nx + sxy*ny);
	vf[2] = -(sxy*nx + syy*ny);
	vf[3] = -((sxx*vax + sxy*vay + kd*gf[0][3])*nx + (sxy*vax + syy*vay + kd*gf[1][3])*ny);
}
/// Sutherland viscosity of a state given its pressure and 1/rho, with the reciprocals of this file
__device__ __forceinline__ double sutherland_fast(const GasParams &G, double p, double ir) {
	const double T = p*ir*G.gM2;
	return (1.0 + G.sCT)*(T*fsqrt(T))*frcp((T + G.sCT)*G.Reinf);
}

} // namespace fvg
