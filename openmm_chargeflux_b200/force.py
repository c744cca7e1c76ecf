"""Host-side mirror of the plugin's API class ``CoulPlugin::CoulForce``.

Same method names, argument order, defaults and storage layout as the reference
(/root/reference/openmmapi/include/CoulForce.h:16-150, src/CoulForce.cpp:12-140) and therefore as the
``openmmcoul.CoulForce`` SWIG class (python/openmmcoul.i:50-75): a script that fills a
``openmmcoul.CoulForce`` fills this class with the same calls. Getters that return through C++
reference arguments return tuples here, as SWIG's typemaps do.
"""
import ctypes as C

import numpy as np

from . import _abi


class CoulForce:
    def __init__(self):
        # CoulForce.cpp:12-16
        self._cutoff = 1.0
        self._ewald_tol = 0.0001
        self._pbc = False
        self._charges = []
        self._ljparams = []          # sigma, epsilon interleaved (CoulForce.h:145)
        self._exclusions = []
        self._fbond_idx, self._fbond_params = [], []
        self._fangle_idx, self._fangle_params = [], []
        self._fwater_idx, self._fwater_params = [], []
        self._force_group = 0

    # -- particles ------------------------------------------------------------------------------
    def addParticle(self, charge, sigma, epsilon):
        self._charges.append(float(charge))
        self._ljparams += [float(sigma), float(epsilon)]

    def getNumParticles(self):
        return len(self._charges)

    def getParticleParameters(self, index):
        return self._charges[index], self._ljparams[2 * index], self._ljparams[2 * index + 1]

    def setParticleParameters(self, index, charge, sigma, epsilon):
        self._charges[index] = float(charge)
        self._ljparams[2 * index] = float(sigma)
        self._ljparams[2 * index + 1] = float(epsilon)

    # -- global settings ------------------------------------------------------------------------
    def getCutoffDistance(self):
        return self._cutoff

    def setCutoffDistance(self, cutoff):
        self._cutoff = float(cutoff)

    def usesPeriodicBoundaryConditions(self):
        return self._pbc

    def setUsesPeriodicBoundaryConditions(self, ifPeriod):
        self._pbc = bool(ifPeriod)

    def setEwaldErrorTolerance(self, tol):
        self._ewald_tol = float(tol)

    def getEwaldErrorTolerance(self):
        return self._ewald_tol

    def getForceGroup(self):
        return self._force_group

    def setForceGroup(self, group):
        self._force_group = int(group)

    # -- exclusions -----------------------------------------------------------------------------
    def addException(self, p1, p2):
        self._exclusions.append((int(p1), int(p2)))

    def getNumExceptions(self):
        return len(self._exclusions)

    def getExceptionParameters(self, index):
        return self._exclusions[index]

    # -- charge-flux terms ----------------------------------------------------------------------
    def addFluxBond(self, p1, p2, k, b):
        self._fbond_idx += [int(p1), int(p2)]
        self._fbond_params += [float(k), float(b)]

    def getFluxBondParameters(self, index):
        return (self._fbond_idx[2 * index], self._fbond_idx[2 * index + 1],
                self._fbond_params[2 * index], self._fbond_params[2 * index + 1])

    def getNumFluxBonds(self):
        return len(self._fbond_idx) // 2

    def addFluxAngle(self, p1, p2, p3, k, theta):
        self._fangle_idx += [int(p1), int(p2), int(p3)]
        self._fangle_params += [float(k), float(theta)]

    def getFluxAngleParameters(self, index):
        return (*self._fangle_idx[3 * index:3 * index + 3], *self._fangle_params[2 * index:2 * index + 2])

    def getNumFluxAngles(self):
        return len(self._fangle_idx) // 3

    def addFluxWater(self, po, ph1, ph2, k1, k2, kub, b0, ub0):
        self._fwater_idx += [int(po), int(ph1), int(ph2)]
        self._fwater_params += [float(k1), float(k2), float(kub), float(b0), float(ub0)]

    def getFluxWaterParameters(self, index):
        return (*self._fwater_idx[3 * index:3 * index + 3], *self._fwater_params[5 * index:5 * index + 5])

    def getNumFluxWaters(self):
        return len(self._fwater_idx) // 3

    # -- bulk fill (not in the reference API; avoids 1e5 Python calls for the synthetic boxes) ----
    def _bulk(self, charge, sigma, epsilon, exceptions=(), bonds=None, angles=None, waters=None):
        charge, sigma, epsilon = (np.asarray(a, dtype=np.float64) for a in (charge, sigma, epsilon))
        self._charges += charge.tolist()
        self._ljparams += np.stack([sigma, epsilon], axis=1).ravel().tolist()
        self._exclusions += [tuple(int(v) for v in p) for p in np.asarray(exceptions, dtype=np.int64).reshape(-1, 2)]
        if bonds is not None:
            idx, par = bonds
            self._fbond_idx += np.asarray(idx, dtype=np.int64).ravel().tolist()
            self._fbond_params += np.asarray(par, dtype=np.float64).ravel().tolist()
        if angles is not None:
            idx, par = angles
            self._fangle_idx += np.asarray(idx, dtype=np.int64).ravel().tolist()
            self._fangle_params += np.asarray(par, dtype=np.float64).ravel().tolist()
        if waters is not None:
            idx, par = waters
            self._fwater_idx += np.asarray(idx, dtype=np.int64).ravel().tolist()
            self._fwater_params += np.asarray(par, dtype=np.float64).ravel().tolist()
        return self

    # -- C ABI parameter block --------------------------------------------------------------------
    def to_desc(self, default_box):
        """Build the ``cfx_system_desc`` block (include/cfx_b200.h). Returns (desc, keepalive)."""
        n = self.getNumParticles()
        lj = np.asarray(self._ljparams, dtype=np.float64).reshape(n, 2)
        arrays = {
            "charge": np.ascontiguousarray(self._charges, dtype=np.float64),
            "sigma": np.ascontiguousarray(lj[:, 0]),
            "epsilon": np.ascontiguousarray(lj[:, 1]),
            "exception_pairs": np.ascontiguousarray(self._exclusions, dtype=np.int32).reshape(-1),
            "flux_bond_idx": np.ascontiguousarray(self._fbond_idx, dtype=np.int32),
            "flux_bond_params": np.ascontiguousarray(self._fbond_params, dtype=np.float64),
            "flux_angle_idx": np.ascontiguousarray(self._fangle_idx, dtype=np.int32),
            "flux_angle_params": np.ascontiguousarray(self._fangle_params, dtype=np.float64),
            "flux_water_idx": np.ascontiguousarray(self._fwater_idx, dtype=np.int32),
            "flux_water_params": np.ascontiguousarray(self._fwater_params, dtype=np.float64),
        }
        d = _abi.SystemDesc()
        d.num_particles = n
        d.num_exceptions = self.getNumExceptions()
        d.num_flux_bonds = self.getNumFluxBonds()
        d.num_flux_angles = self.getNumFluxAngles()
        d.num_flux_waters = self.getNumFluxWaters()
        for name, arr in arrays.items():
            ptr_t = _abi.c_int32_p if arr.dtype == np.int32 else _abi.c_double_p
            setattr(d, name, arr.ctypes.data_as(ptr_t))
        d.cutoff = self._cutoff
        d.ewald_tol = self._ewald_tol
        d.use_pbc = 1 if self._pbc else 0
        box = np.asarray(default_box, dtype=np.float64)
        if box.shape == (3,):
            box = np.diag(box)
        d.default_box = (C.c_double * 9)(*box.reshape(9).tolist())
        return d, arrays
