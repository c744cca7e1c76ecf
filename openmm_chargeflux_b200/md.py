"""MD harness around the CalcCoulForce kernel: harmonic bonds/angles +
velocity Verlet on the GPU, the pieces OpenMM itself contributes around the plugin in a flexible-water
simulation. Used for the "NVE MD" benchmark configuration and the energy-conservation tests."""
import ctypes as C

import numpy as np

from . import _abi, runtime

KB = 0.0083144626            # kJ/mol/K
# flexible TIP3P-like intramolecular terms (kJ/mol/nm^2, nm, kJ/mol/rad^2, rad)
WATER_BOND_K, WATER_BOND_R0 = 462750.4, 0.09572
WATER_ANGLE_K, WATER_ANGLE_THETA0 = 836.8, 1.82421813
MASS_O, MASS_H = 15.999, 1.008


class NVESimulation:
    def __init__(self, force, box, masses, bonds=None, angles=None, device=-1):
        self.kernel = runtime.CalcCoulForceKernel(device=device)
        self.kernel.initialize(box, force)
        self.box = runtime._box9(box)
        self.n = force.getNumParticles()
        lib = self.kernel._lib
        lib.cfx_md_create.argtypes = [C.c_void_p, _abi.c_double_p, C.c_int32, _abi.c_int32_p, _abi.c_double_p,
                                      C.c_int32, _abi.c_int32_p, _abi.c_double_p, C.POINTER(C.c_void_p)]
        lib.cfx_md_destroy.argtypes = [C.c_void_p]
        lib.cfx_md_destroy.restype = None
        lib.cfx_md_set_state.argtypes = [C.c_void_p, _abi.c_double_p, _abi.c_double_p, _abi.c_double_p]
        lib.cfx_md_get_state.argtypes = [C.c_void_p, _abi.c_double_p, _abi.c_double_p]
        lib.cfx_md_minimize.argtypes = [C.c_void_p, C.c_int32, C.c_double]
        lib.cfx_md_step.argtypes = [C.c_void_p, C.c_int32, C.c_double, C.POINTER(C.c_float)]
        lib.cfx_md_energies.argtypes = [C.c_void_p, _abi.c_double_p]
        self._lib = lib
        m = np.ascontiguousarray(masses, dtype=np.float64)
        bi, bp = (np.zeros(0, np.int32), np.zeros(0)) if bonds is None else \
            (np.ascontiguousarray(bonds[0], np.int32).reshape(-1), np.ascontiguousarray(bonds[1], np.float64).reshape(-1))
        ai, ap = (np.zeros(0, np.int32), np.zeros(0)) if angles is None else \
            (np.ascontiguousarray(angles[0], np.int32).reshape(-1), np.ascontiguousarray(angles[1], np.float64).reshape(-1))
        self.masses = m
        h = C.c_void_p()
        self._check(lib.cfx_md_create(self.kernel._h, m.ctypes.data_as(_abi.c_double_p), len(bi) // 2,
                                      bi.ctypes.data_as(_abi.c_int32_p), bp.ctypes.data_as(_abi.c_double_p), len(ai) // 3,
                                      ai.ctypes.data_as(_abi.c_int32_p), ap.ctypes.data_as(_abi.c_double_p), C.byref(h)))
        self._md = h

    def _check(self, code):
        if code != 0:
            raise runtime.CfxError(self._lib.cfx_last_error().decode())

    def set_state(self, positions, velocities=None):
        p = np.ascontiguousarray(positions, dtype=np.float64).reshape(-1)
        v = None if velocities is None else np.ascontiguousarray(velocities, dtype=np.float64).reshape(-1)
        self._check(self._lib.cfx_md_set_state(self._md, p.ctypes.data_as(_abi.c_double_p),
                                               None if v is None else v.ctypes.data_as(_abi.c_double_p),
                                               self.box.ctypes.data_as(_abi.c_double_p)))

    def get_state(self):
        p, v = np.zeros(3 * self.n), np.zeros(3 * self.n)
        self._check(self._lib.cfx_md_get_state(self._md, p.ctypes.data_as(_abi.c_double_p), v.ctypes.data_as(_abi.c_double_p)))
        return p.reshape(-1, 3), v.reshape(-1, 3)

    def minimize(self, steps, max_displacement=0.002):
        self._check(self._lib.cfx_md_minimize(self._md, steps, max_displacement))

    def step(self, nsteps, dt):
        ms = C.c_float(0)
        self._check(self._lib.cfx_md_step(self._md, nsteps, dt, C.byref(ms)))
        return ms.value

    def energies(self):
        e = np.zeros(4)
        self._check(self._lib.cfx_md_energies(self._md, e.ctypes.data_as(_abi.c_double_p)))
        return dict(kinetic=e[0], bonded=e[1], coulomb=e[2], total=e[3])

    def maxwell_boltzmann(self, temperature, seed=0):
        rng = np.random.Generator(np.random.PCG64(seed))
        v = rng.normal(size=(self.n, 3)) * np.sqrt(KB * temperature / self.masses)[:, None]
        v -= (self.masses[:, None] * v).sum(0) / self.masses.sum()
        return v

    def close(self):
        if getattr(self, "_md", None):
            self._lib.cfx_md_destroy(self._md)
            self._md = None
            self.kernel.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def flexible_water_simulation(n_waters, seed, cutoff=1.0, ewald_tol=1e-5, device=-1):
    """The benchmark water box with intramolecular harmonic terms: returns (sim, positions)."""
    from . import synthetic
    pos, box, force = synthetic.water_box(n_waters, seed, cutoff=cutoff, ewald_tol=ewald_tol)
    o = 3 * np.arange(n_waters)
    bonds = (np.stack([np.stack([o, o + 1], 1), np.stack([o, o + 2], 1)], 1).reshape(-1, 2),
             np.tile([WATER_BOND_K, WATER_BOND_R0], (2 * n_waters, 1)))
    angles = (np.stack([o + 1, o, o + 2], 1), np.tile([WATER_ANGLE_K, WATER_ANGLE_THETA0], (n_waters, 1)))
    masses = np.tile([MASS_O, MASS_H, MASS_H], n_waters)
    sim = NVESimulation(force, box, masses, bonds, angles, device=device)
    sim.set_state(pos)
    return sim, pos
