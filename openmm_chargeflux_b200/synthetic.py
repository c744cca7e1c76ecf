"""Synthetic inputs for the benchmark configs of BASELINE.json.

All generators are deterministic functions of ``seed`` (numpy ``Generator(PCG64(seed))``), return
unwrapped double coordinates in nm and a filled :class:`CoulForce`.
"""
import numpy as np

from .force import CoulForce

WATER_DENSITY = 33.43  # molecules / nm^3

# parameter set of SURVEY.md section 8d
Q_O, Q_H = -0.834, 0.417
SIG_O, EPS_O = 0.315075, 0.635968
SIG_H, EPS_H = 0.1, 0.0
R_OH, THETA_HOH = 0.09572, np.deg2rad(104.52)
FLUX_BOND_K, FLUX_BOND_B = 2.0, 0.09572
FLUX_ANGLE_K, FLUX_ANGLE_THETA0 = 0.08, 1.82421813


def _random_rotations(rng, n):
    """Uniform random rotation matrices [n,3,3] from unit quaternions."""
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    return np.stack([
        np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], axis=1),
        np.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], axis=1),
        np.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], axis=1),
    ], axis=1)


def _lattice_sites(n, box_len):
    m = int(np.ceil(n ** (1.0 / 3.0) - 1e-9))
    while m ** 3 < n:
        m += 1
    idx = np.arange(n)
    ijk = np.stack([idx // (m * m), (idx // m) % m, idx % m], axis=1)
    return (ijk + 0.5) * (box_len / m)


def _water_geometry(rng, centres):
    n = len(centres)
    r1 = rng.normal(R_OH, 0.002, size=n)
    r2 = rng.normal(R_OH, 0.002, size=n)
    th = rng.normal(THETA_HOH, np.deg2rad(2.0), size=n)
    h1 = np.stack([r1 * np.sin(th / 2), np.zeros(n), r1 * np.cos(th / 2)], axis=1)
    h2 = np.stack([-r2 * np.sin(th / 2), np.zeros(n), r2 * np.cos(th / 2)], axis=1)
    rot = _random_rotations(rng, n)
    pos = np.empty((n, 3, 3))
    pos[:, 0] = centres
    pos[:, 1] = centres + np.einsum("nij,nj->ni", rot, h1)
    pos[:, 2] = centres + np.einsum("nij,nj->ni", rot, h2)
    return pos.reshape(-1, 3)


def water_box(n_waters, seed, periodic=True, cutoff=1.0, ewald_tol=1e-4, flux="bond+angle"):
    """Flexible-water box: returns (positions [3*n_w,3], box [3,3], CoulForce).

    Atom order O,H1,H2 per molecule; exclusions (O,H1),(O,H2),(H1,H2); flux bonds (O,H1),(O,H2) and
    flux angle (H1,O,H2), or one flux-water term per molecule when ``flux == "water"``.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    box_len = (n_waters / WATER_DENSITY) ** (1.0 / 3.0)
    centres = _lattice_sites(n_waters, box_len) + rng.uniform(-0.02, 0.02, size=(n_waters, 3))
    pos = _water_geometry(rng, centres)
    o = 3 * np.arange(n_waters)
    charge = np.tile([Q_O, Q_H, Q_H], n_waters)
    sigma = np.tile([SIG_O, SIG_H, SIG_H], n_waters)
    eps = np.tile([EPS_O, EPS_H, EPS_H], n_waters)
    excl = np.stack([np.stack([o, o + 1], 1), np.stack([o, o + 2], 1), np.stack([o + 1, o + 2], 1)], 1).reshape(-1, 2)
    f = CoulForce()
    bonds = angles = waters = None
    if flux == "bond+angle":
        bidx = np.stack([np.stack([o, o + 1], 1), np.stack([o, o + 2], 1)], 1).reshape(-1, 2)
        bonds = (bidx, np.tile([FLUX_BOND_K, FLUX_BOND_B], (2 * n_waters, 1)))
        angles = (np.stack([o + 1, o, o + 2], 1), np.tile([FLUX_ANGLE_K, FLUX_ANGLE_THETA0], (n_waters, 1)))
    elif flux == "water":
        ub0 = 2 * R_OH * np.sin(THETA_HOH / 2)
        waters = (np.stack([o, o + 1, o + 2], 1), np.tile([1.6, 0.4, -0.3, R_OH, ub0], (n_waters, 1)))
    elif flux != "none":
        raise ValueError(flux)
    f._bulk(charge, sigma, eps, excl, bonds, angles, waters)
    f.setUsesPeriodicBoundaryConditions(periodic)
    f.setCutoffDistance(cutoff)
    f.setEwaldErrorTolerance(ewald_tol)
    return pos, np.diag([box_len] * 3), f


def methanol_water(n_methanol=100, n_water=300, seed=5, cutoff=1.0, ewald_tol=1e-5):
    """Config C5: methanol/water mixture with 1-2/1-3 exclusions and mixed flux parameters.

    Methanol atom order C,H,H,H,O,HO; OPLS-like charges C .145 / HC .04 / O -.683 / HO .418;
    12 exclusions per methanol (5 bonds + 7 angles); every bond is a flux bond and every angle a flux
    angle with per-type k. Waters as in :func:`water_box` but described by flux-water terms.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    n_mol = n_methanol + n_water
    n_atoms = 6 * n_methanol + 3 * n_water
    box_len = (n_atoms / 96.0) ** (1.0 / 3.0)   # ~96 atoms/nm^3 -> L ~ 2.5 nm at 1500 atoms
    sites = _lattice_sites(n_mol, box_len) + rng.uniform(-0.02, 0.02, size=(n_mol, 3))
    perm = rng.permutation(n_mol)
    meth_sites, water_sites = sites[perm[:n_methanol]], sites[perm[n_methanol:]]
    # methanol template (nm): C at origin, O along +z, methyl H's tetrahedral, hydroxyl H bent
    rco, rch, roh = 0.1410, 0.1090, 0.0945
    tet = np.deg2rad(109.5)
    tmpl = np.zeros((6, 3))
    for k in range(3):
        phi = 2 * np.pi * k / 3
        tmpl[1 + k] = rch * np.array([np.sin(tet) * np.cos(phi), np.sin(tet) * np.sin(phi), np.cos(tet)])
    tmpl[4] = [0, 0, rco]
    coh = np.deg2rad(108.5)
    tmpl[5] = tmpl[4] + roh * np.array([np.sin(np.pi - coh), 0, np.cos(np.pi - coh)])
    rot = _random_rotations(rng, n_methanol)
    mpos = meth_sites[:, None, :] + np.einsum("nij,aj->nai", rot, tmpl) + rng.normal(0, 0.002, size=(n_methanol, 6, 3))
    wpos = _water_geometry(rng, water_sites)
    pos = np.concatenate([mpos.reshape(-1, 3), wpos], axis=0)

    f = CoulForce()
    m_q = [0.145, 0.04, 0.04, 0.04, -0.683, 0.418]
    m_sig = [0.35, 0.25, 0.25, 0.25, 0.312, 0.1]
    m_eps = [0.276144, 0.0, 0.0, 0.0, 0.711280, 0.0]
    m_bonds = [(0, 1), (0, 2), (0, 3), (0, 4), (4, 5)]
    m_angles = [(1, 0, 2), (1, 0, 3), (2, 0, 3), (1, 0, 4), (2, 0, 4), (3, 0, 4), (0, 4, 5)]
    bond_k = {(0, 1): 0.35, (0, 2): 0.35, (0, 3): 0.35, (0, 4): -1.1, (4, 5): 1.7}
    angle_k = {0: 0.02, 4: 0.06}
    for m in range(n_methanol):
        base = 6 * m
        for a in range(6):
            f.addParticle(m_q[a], m_sig[a], m_eps[a])
    for w in range(n_water):
        f.addParticle(Q_O, SIG_O, EPS_O)
        f.addParticle(Q_H, SIG_H, EPS_H)
        f.addParticle(Q_H, SIG_H, EPS_H)
    for m in range(n_methanol):
        base = 6 * m
        for (a, b) in m_bonds:
            f.addException(base + a, base + b)
            ra = np.linalg.norm(tmpl[a] - tmpl[b])
            f.addFluxBond(base + a, base + b, bond_k[(a, b)], ra)
        for (a, b, c) in m_angles:
            f.addException(base + a, base + c)
            v1, v2 = tmpl[a] - tmpl[b], tmpl[c] - tmpl[b]
            th0 = np.arccos(v1 @ v2 / np.linalg.norm(v1) / np.linalg.norm(v2))
            f.addFluxAngle(base + a, base + b, base + c, angle_k[b], th0)
    ub0 = 2 * R_OH * np.sin(THETA_HOH / 2)
    for w in range(n_water):
        o = 6 * n_methanol + 3 * w
        f.addException(o, o + 1)
        f.addException(o, o + 2)
        f.addException(o + 1, o + 2)
        f.addFluxWater(o, o + 1, o + 2, 1.6, 0.4, -0.3, R_OH, ub0)
    f.setUsesPeriodicBoundaryConditions(True)
    f.setCutoffDistance(cutoff)
    f.setEwaldErrorTolerance(ewald_tol)
    return pos, np.diag([box_len] * 3), f


def rock_salt(cells=2, a=0.564, charge=1.0, cutoff=None, ewald_tol=1e-6):
    """NaCl rock-salt supercell (no flux, no LJ): the Madelung known-answer case."""
    pts, q = [], []
    for i in range(2 * cells):
        for j in range(2 * cells):
            for k in range(2 * cells):
                pts.append([i, j, k])
                q.append(charge if (i + j + k) % 2 == 0 else -charge)
    pos = np.asarray(pts, dtype=np.float64) * (a / 2)
    box_len = cells * a
    f = CoulForce()
    f._bulk(q, np.zeros(len(q)), np.zeros(len(q)))
    f.setUsesPeriodicBoundaryConditions(True)
    f.setCutoffDistance(cutoff if cutoff is not None else 0.49 * box_len)
    f.setEwaldErrorTolerance(ewald_tol)
    return pos, np.diag([box_len] * 3), f


# the benchmark configurations of BASELINE.json
CONFIGS = {
    "c1": dict(n_waters=64, seed=64, periodic=False),
    "c2": dict(n_waters=1365, seed=4096, periodic=True, cutoff=1.0, ewald_tol=1e-4),
    "c3": dict(n_waters=10922, seed=32768, periodic=True, cutoff=1.0, ewald_tol=1e-5),
    "c4": dict(n_waters=87381, seed=262144, periodic=True, cutoff=1.0, ewald_tol=1e-5),
}


def config(name):
    if name == "c5":
        return methanol_water()
    return water_box(**CONFIGS[name])


def ballistic_frames(pos, n_frames, dt_ps=0.0005, temperature=300.0, seed=17):
    """Positions for a benchmark whose atoms move every step: frame k = pos + k dt v with per-atom Maxwell-Boltzmann
    velocities (O: 15.999, H: 1.008 g/mol; kT in kJ/mol), i.e. straight-line motion at thermal speed. Harder on a
    neighbour list than real MD (bonded atoms vibrate instead of flying apart): the fastest hydrogens cover 0.05 nm in
    about ten 0.5 fs steps. Callers walk the frames forwards and backwards so that the geometry stays bounded.
    Returns float64 [n_frames, N, 3]."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n = len(pos)
    mass = np.tile([15.999, 1.008, 1.008], (n + 2) // 3)[:n]
    sigma = np.sqrt(0.0083144626 * temperature / mass)          # nm/ps
    vel = rng.normal(size=(n, 3)) * sigma[:, None]
    k = np.arange(n_frames, dtype=np.float64)[:, None, None]
    return pos[None, :, :] + k * dt_ps * vel[None, :, :]


def ping_pong(step, n_frames):
    """Frame index of step `step` when walking 0, 1, ..., n-1, n-2, ..., 1, 0, 1, ..."""
    period = 2 * (n_frames - 1)
    r = step % period
    return r if r < n_frames else period - r
