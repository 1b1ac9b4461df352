"""Deterministic synthetic hybrid meshes and states (Gmsh is not available offline).

Generators return (coords[npoin,2], nnode[nelem], inpoel[nelem,4] (-1 padded), bface[nbface,3]) in
the layout of the reference's MeshData (src/mesh/meshreaders.hpp:29-56). Geometry follows the
reference's .geo files: tests/inv-gaussianbump/gaussian_channel.geo (bump channel) and
testcases/2dcylinder/grids/2dcylstruct.geo (O-grid cylinder).
"""
import numpy as np


def _hash01(i, j, seed):
    """Deterministic per-(i,j) pseudo-random number in [0,1) (splitmix-style integer mixing)."""
    with np.errstate(over="ignore"):
        h = (i.astype(np.uint64)*np.uint64(0x9E3779B97F4A7C15)) ^ (j.astype(np.uint64)*np.uint64(0xC2B2AE3D27D4EB4F)) \
            ^ np.uint64((seed*0x165667B19E3779F9) & 0xFFFFFFFFFFFFFFFF)
        h ^= h >> np.uint64(30)
        h *= np.uint64(0xBF58476D1CE4E5B9)
        h ^= h >> np.uint64(27)
        h *= np.uint64(0x94D049BB133111EB)
        h ^= h >> np.uint64(31)
    return (h >> np.uint64(40)).astype(np.float64)/float(1 << 24)


def _lattice_cells(nx, ny, split, node_id):
    """Cells of an nx x ny lattice of quads, row-major; quads flagged in `split` become two triangles
    whose diagonal alternates with (i+j)&1. node_id(i,j) maps lattice indices to node numbers."""
    I, J = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")     # J rows, I columns
    I = I.ravel(); J = J.ravel(); sp = split.ravel()
    n00, n10, n11, n01 = node_id(I, J), node_id(I+1, J), node_id(I+1, J+1), node_id(I, J+1)
    ncell_of = np.where(sp, 2, 1)
    off = np.concatenate(([0], np.cumsum(ncell_of)))
    nelem = int(off[-1])
    inpoel = np.full((nelem, 4), -1, dtype=np.int32)
    nnode = np.empty(nelem, dtype=np.int32)
    q = ~sp
    oq = off[:-1][q]
    inpoel[oq, 0], inpoel[oq, 1], inpoel[oq, 2], inpoel[oq, 3] = n00[q], n10[q], n11[q], n01[q]
    nnode[oq] = 4
    odd = ((I + J) & 1).astype(bool)
    a = sp & ~odd            # diagonal n00-n11
    oa = off[:-1][a]
    inpoel[oa, 0], inpoel[oa, 1], inpoel[oa, 2] = n00[a], n10[a], n11[a]
    inpoel[oa+1, 0], inpoel[oa+1, 1], inpoel[oa+1, 2] = n00[a], n11[a], n01[a]
    b = sp & odd             # diagonal n10-n01
    ob = off[:-1][b]
    inpoel[ob, 0], inpoel[ob, 1], inpoel[ob, 2] = n00[b], n10[b], n01[b]
    inpoel[ob+1, 0], inpoel[ob+1, 1], inpoel[ob+1, 2] = n10[b], n11[b], n01[b]
    for o in (oa, ob):
        nnode[o] = 3
        nnode[o+1] = 3
    return nnode, inpoel


def bump_channel(nx, ny, tri_fraction=1.0/3.0, seed=12345, jitter=0.15):
    """Gaussian-bump channel x in [-1,1], y from 0.02*exp(-100 x^2) to 0.75, hybrid tri/quad.
    Markers: 2 = bottom and top walls, 3 = inlet (left), 4 = outlet (right) as tests/inv-gaussianbump/base.ctrl.
    tri_fraction = probability that a base quad is split; 1/3 gives a 50:50 tri/quad cell count."""
    xi = np.linspace(-1.0, 1.0, nx+1)
    eta = np.linspace(0.0, 1.0, ny+1)
    X = np.repeat(xi[None, :], ny+1, axis=0)
    E = np.repeat(eta[:, None], nx+1, axis=1)
    hx, he = 2.0/nx, 1.0/ny
    rng = np.random.default_rng(seed)
    dX = rng.uniform(-jitter, jitter, size=X.shape)*hx
    dE = rng.uniform(-jitter, jitter, size=X.shape)*he
    dX[:, 0] = dX[:, -1] = 0.0; dX[0, :] = dX[-1, :] = 0.0
    dE[:, 0] = dE[:, -1] = 0.0; dE[0, :] = dE[-1, :] = 0.0
    X = X + dX; E = E + dE
    yb = 0.02*np.exp(-100.0*X*X)
    Y = yb + E*(0.75 - yb)
    coords = np.stack([X.ravel(), Y.ravel()], axis=1)

    def nid(i, j):
        return (j*(nx+1) + i).astype(np.int32)

    I, J = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    split = _hash01(I, J, seed) < tri_fraction
    nnode, inpoel = _lattice_cells(nx, ny, split, nid)
    ii = np.arange(nx); jj = np.arange(ny)
    z = np.zeros_like
    bottom = np.stack([nid(ii, z(ii)), nid(ii+1, z(ii)), np.full(nx, 2)], axis=1)
    right = np.stack([nid(z(jj)+nx, jj), nid(z(jj)+nx, jj+1), np.full(ny, 4)], axis=1)
    top = np.stack([nid(ii+1, z(ii)+ny), nid(ii, z(ii)+ny), np.full(nx, 2)], axis=1)
    left = np.stack([nid(z(jj), jj+1), nid(z(jj), jj), np.full(ny, 3)], axis=1)
    bface = np.concatenate([bottom, right, top, left]).astype(np.int32)
    return coords, nnode, inpoel, bface


def ogrid_cylinder(ntheta, nr, r0=0.5, rfar=20.0, tri_fraction=0.0, seed=12345):
    """O-grid around a cylinder of radius r0 out to rfar, geometric radial stretching.
    Markers: 2 = cylinder wall, 4 = far field (testcases/2dcylinder)."""
    ratio = (rfar/r0)**(1.0/nr)
    r = r0*ratio**np.arange(nr+1)
    th = 2.0*np.pi*np.arange(ntheta)/ntheta
    # node (i=theta index, j=radial index); theta runs clockwise so that cells are counter-clockwise
    X = (r[:, None]*np.cos(-th)[None, :]).ravel()
    Y = (r[:, None]*np.sin(-th)[None, :]).ravel()
    coords = np.stack([X, Y], axis=1)

    def nid(i, j):
        return (j*ntheta + (i % ntheta)).astype(np.int32)

    I, J = np.meshgrid(np.arange(ntheta), np.arange(nr), indexing="xy")
    split = _hash01(I, J, seed) < tri_fraction
    # lattice orientation: i runs clockwise in theta and j outward in r, so (n00,n10,n11,n01) is counter-clockwise
    nnode, inpoel = _lattice_cells(ntheta, nr, split, nid)
    ii = np.arange(ntheta)
    z = np.zeros_like
    wall = np.stack([nid(ii, z(ii)), nid(ii+1, z(ii)), np.full(ntheta, 2)], axis=1)
    far = np.stack([nid(ii+1, z(ii)+nr), nid(ii, z(ii)+nr), np.full(ntheta, 4)], axis=1)
    bface = np.concatenate([wall, far]).astype(np.int32)
    return coords, nnode, inpoel, bface


def square(n, lo=-5.0, hi=5.0, tri_fraction=0.0, seed=12345, jitter=0.0):
    """[lo,hi]^2 with n x n base quads. Markers: 1 bottom, 2 right, 3 top, 4 left."""
    xi = np.linspace(lo, hi, n+1)
    X = np.repeat(xi[None, :], n+1, axis=0)
    Y = np.repeat(xi[:, None], n+1, axis=1)
    if jitter > 0:
        rng = np.random.default_rng(seed)
        h = (hi-lo)/n
        dX = rng.uniform(-jitter, jitter, size=X.shape)*h
        dY = rng.uniform(-jitter, jitter, size=X.shape)*h
        dX[:, 0] = dX[:, -1] = 0.0; dX[0, :] = dX[-1, :] = 0.0
        dY[:, 0] = dY[:, -1] = 0.0; dY[0, :] = dY[-1, :] = 0.0
        X = X + dX; Y = Y + dY
    coords = np.stack([X.ravel(), Y.ravel()], axis=1)

    def nid(i, j):
        return (j*(n+1) + i).astype(np.int32)

    I, J = np.meshgrid(np.arange(n), np.arange(n), indexing="xy")
    split = _hash01(I, J, seed) < tri_fraction
    nnode, inpoel = _lattice_cells(n, n, split, nid)
    ii = np.arange(n)
    z = np.zeros_like
    bottom = np.stack([nid(ii, z(ii)), nid(ii+1, z(ii)), np.full(n, 1)], axis=1)
    right = np.stack([nid(z(ii)+n, ii), nid(z(ii)+n, ii+1), np.full(n, 2)], axis=1)
    top = np.stack([nid(ii+1, z(ii)+n), nid(ii, z(ii)+n), np.full(n, 3)], axis=1)
    left = np.stack([nid(z(ii), ii+1), nid(z(ii), ii), np.full(n, 4)], axis=1)
    bface = np.concatenate([bottom, right, top, left]).astype(np.int32)
    return coords, nnode, inpoel, bface


def periodic_square(n, lo=-5.0, hi=5.0, tri_fraction=0.0, seed=12345, jitter=0.0):
    """The doubly periodic box of the isentropic-vortex series (tests/isentropic-vortex/grids/dom.geo: [-5,5]^2): as
    square(), with marker 3 on the left and right sides (periodic in x) and 4 on the bottom and top (periodic in y).
    Jitter moves interior nodes only, so that partner faces keep matching midpoints."""
    coords, nnode, inpoel, bface = square(n, lo, hi, tri_fraction, seed, jitter)
    bface = bface.copy()
    m = bface[:, 2]
    bface[:, 2] = np.where((m == 2) | (m == 4), 3, 4)
    return coords, nnode, inpoel, bface


def boundary_edges(nnode, inpoel):
    """(node0, node1) of every edge that belongs to exactly one cell, oriented as in that cell (counter-clockwise)."""
    n = len(nnode)
    j = np.arange(4)[None, :]
    a = inpoel
    b = np.take_along_axis(inpoel, (j + 1) % nnode[:, None], axis=1)
    valid = j < nnode[:, None]
    a = a[valid]; b = b[valid]
    lo = np.minimum(a, b).astype(np.int64); hi = np.maximum(a, b).astype(np.int64)
    key = lo*(int(inpoel.max()) + 1) + hi
    _, inv, cnt = np.unique(key, return_inverse=True, return_counts=True)
    once = cnt[inv] == 1
    return np.stack([a[once], b[once]], axis=1).astype(np.int32)


def unfold_periodic(coords, nnode, inpoel, period, marker=9):
    """3 x 3 copies of a doubly periodic mesh laid side by side (the centre copy first, so that its cells keep their
    indices), nodes on the seams merged, every outer edge a boundary face with `marker`. On this mesh an ordinary
    (non-periodic) evaluation gives, in the centre copy, exactly what a periodic evaluation gives on the original mesh:
    the oracle for the periodic pairing, which the reference cannot run."""
    shifts = [(0, 0)] + [(i, j) for j in (-1, 0, 1) for i in (-1, 0, 1) if (i, j) != (0, 0)]
    npo = len(coords)
    allc = np.concatenate([coords + np.array([i*period, j*period]) for (i, j) in shifts])
    h = np.sqrt(period*period/len(nnode))
    keyxy = np.round(allc/(1e-6*h)).astype(np.int64)
    _, first, inv = np.unique(keyxy, axis=0, return_index=True, return_inverse=True)
    inv = inv.reshape(-1)
    newcoords = allc[first]
    cells = []
    for k in range(len(shifts)):
        ip = np.where(inpoel >= 0, inv[np.where(inpoel >= 0, inpoel, 0) + k*npo], -1)
        cells.append(ip)
    ip = np.concatenate(cells).astype(np.int32)
    nn = np.tile(nnode, len(shifts)).astype(np.int32)
    be = boundary_edges(nn, ip)
    bface = np.concatenate([be, np.full((len(be), 1), marker, dtype=np.int32)], axis=1).astype(np.int32)
    return newcoords, nn, ip, bface


def cell_centres(coords, nnode, inpoel):
    idx = np.where(inpoel < 0, 0, inpoel)
    w = (inpoel >= 0).astype(np.float64)
    cx = (coords[idx, 0]*w).sum(axis=1)/nnode
    cy = (coords[idx, 1]*w).sum(axis=1)/nnode
    return np.stack([cx, cy], axis=1)


def hilbert_order(rc, order=20):
    """new2old permutation of the points rc along a Hilbert curve on a 2^order grid over their bounding box: the numpy
    twin of fvg_umesh_hilbert_ordering (csrc/device_mesh.cu hilbert_order / hilbert_d), ties broken by index. Lets a
    process build the Hilbert-ordered benchmark mesh without loading libfvens_b200.so (bench.py --impl reference)."""
    rc = np.asarray(rc, dtype=np.float64)
    lo = rc.min(axis=0); hi = rc.max(axis=0)
    span = max(hi[0] - lo[0], hi[1] - lo[1], 1e-300)
    scale = float((1 << order) - 1)/span
    x = ((rc[:, 0] - lo[0])*scale).astype(np.uint32)
    y = ((rc[:, 1] - lo[1])*scale).astype(np.uint32)
    n = np.uint32(1 << order)
    d = np.zeros(len(rc), dtype=np.uint64)
    s = 1 << (order - 1)
    while s > 0:
        s32 = np.uint32(s)
        rx = (x & s32) != 0
        ry = (y & s32) != 0
        d += np.uint64(s*s)*((3*rx.astype(np.uint64)) ^ ry.astype(np.uint64))
        flip = ~ry & rx
        x = np.where(flip, n - np.uint32(1) - x, x)
        y = np.where(flip, n - np.uint32(1) - y, y)
        swap = ~ry
        x, y = np.where(swap, y, x), np.where(swap, x, y)
        s >>= 1
    return np.argsort(d, kind="stable").astype(np.int32)


def freestream_state(gamma, Minf, aoa):
    """IdealGasPhysics::compute_freestream_state (src/physics/aphysics.cpp:44-58)"""
    pinf = 1.0/(gamma*Minf*Minf)
    return np.array([1.0, np.cos(aoa), np.sin(aoa), pinf/(gamma-1.0) + 0.5])


def perturbed_state(rc, gamma, Minf, aoa=0.0, amp=0.05, shock=False):
    """Smooth admissible test state: freestream with delta = amp*sin(2 pi x)*cos(3 pi y) on
    rho, v, p; optionally a normal-shock jump (Rankine-Hugoniot, M = 1.3) at x = 0."""
    x, y = rc[:, 0], rc[:, 1]
    d = amp*np.sin(2*np.pi*x)*np.cos(3*np.pi*y)
    pinf = 1.0/(gamma*Minf*Minf)
    rho = 1.0 + d
    vx = np.cos(aoa)*(1.0 + d)
    vy = np.sin(aoa) + 0.5*d
    p = pinf*(1.0 + d)
    if shock:
        Ms = 1.3
        pr = 1.0 + 2.0*gamma/(gamma+1.0)*(Ms*Ms-1.0)
        rr = (gamma+1.0)*Ms*Ms/((gamma-1.0)*Ms*Ms+2.0)
        right = x > 0
        rho = np.where(right, rho*rr, rho)
        p = np.where(right, p*pr, p)
        vx = np.where(right, vx/rr, vx)
    E = p/(gamma-1.0) + 0.5*rho*(vx*vx + vy*vy)
    return np.ascontiguousarray(np.stack([rho, rho*vx, rho*vy, E], axis=1))


def isentropic_vortex(rc, gamma, Minf, t=0.0, strength=0.6, sigma=0.8, clength=1.0, centre=(0.0, 0.0)):
    """Isentropic vortex convected by the unit free stream along x, evaluated at the points rc at time t (conserved
    variables, the reference's non-dimensionalisation: rho_inf = |v_inf| = 1, p_inf = 1/(gamma M^2)).
    omega = strength * exp(-r^2 / (2 sigma^2 clength^2)), v = v_inf + (-y, x)/clength * omega, T/T_inf = 1 -
    (gamma-1)/2 (M sigma)^2 omega^2, rho = T^(1/(gamma-1)), p = p_inf T^(gamma/(gamma-1)) - an exact solution of the
    Euler equations (radial balance dp/dr = rho v_theta^2 / r). The reference's own vortex test
    (tests/isentropic-vortex/isentropicvortex.cpp:18-43, disabled in tests/CMakeLists.txt:46, written for units with
    c_inf = 1) has the same form without the (M sigma)^2 factor; BASELINE configs[4] / SURVEY 8d config 5."""
    x = rc[:, 0] - centre[0] - t
    y = rc[:, 1] - centre[1]
    om = strength*np.exp(-(x*x + y*y)/(2.0*sigma*sigma*clength*clength))
    th = 1.0 - 0.5*(gamma - 1.0)*(Minf*sigma)**2*om*om
    rho = th**(1.0/(gamma - 1.0))
    p = th**(gamma/(gamma - 1.0))/(gamma*Minf*Minf)
    vx = 1.0 - y/clength*om
    vy = x/clength*om
    return np.stack([rho, rho*vx, rho*vy, p/(gamma - 1.0) + 0.5*rho*(vx*vx + vy*vy)], axis=1)
