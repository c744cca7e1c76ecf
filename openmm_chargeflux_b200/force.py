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

    # -- serialization (the reference registers no proxy) ------------------
    # Same XML as plugin/CoulForceProxy.cpp writes through OpenMM's XmlSerializer, byte for byte: attributes in
    # alphabetical order (the node's property map), doubles as %.17g, one tab per level.
    def to_xml(self):
        def el(depth, name, attrs, children=None):
            head = "\t" * depth + "<" + name + "".join(' %s="%s"' % (k, attrs[k]) for k in sorted(attrs))
            if not children:
                return [head + "/>"]
            return [head + ">"] + children + ["\t" * depth + "</" + name + ">"]

        def d(v):
            return "%.17g" % v

        n = self.getNumParticles()
        kids = []
        kids += el(1, "Particles", {}, [x for i in range(n) for x in el(2, "Particle", {
            "q": d(self._charges[i]), "sig": d(self._ljparams[2 * i]), "eps": d(self._ljparams[2 * i + 1])})])
        kids += el(1, "Exceptions", {}, [x for p1, p2 in self._exclusions for x in el(2, "Exception", {"p1": p1, "p2": p2})])
        kids += el(1, "FluxBonds", {}, [x for i in range(self.getNumFluxBonds()) for x in el(2, "Bond", dict(
            zip(("p1", "p2"), self._fbond_idx[2 * i:2 * i + 2]), **dict(zip(("k", "b"), map(d, self._fbond_params[2 * i:2 * i + 2])))))])
        kids += el(1, "FluxAngles", {}, [x for i in range(self.getNumFluxAngles()) for x in el(2, "Angle", dict(
            zip(("p1", "p2", "p3"), self._fangle_idx[3 * i:3 * i + 3]),
            **dict(zip(("k", "theta"), map(d, self._fangle_params[2 * i:2 * i + 2])))))])
        kids += el(1, "FluxWaters", {}, [x for i in range(self.getNumFluxWaters()) for x in el(2, "Water", dict(
            zip(("po", "ph1", "ph2"), self._fwater_idx[3 * i:3 * i + 3]),
            **dict(zip(("k1", "k2", "kub", "b0", "ub0"), map(d, self._fwater_params[5 * i:5 * i + 5])))))])
        root = el(0, "Force", {"type": "CoulForce", "version": 1, "forceGroup": self._force_group, "cutoff": d(self._cutoff),
                               "ewaldTolerance": d(self._ewald_tol), "usesPeriodic": 1 if self._pbc else 0}, kids)
        return '<?xml version="1.0" ?>\n' + "\n".join(root) + "\n"

    @classmethod
    def from_xml(cls, text):
        import xml.etree.ElementTree as ET
        root = ET.fromstring(text)
        if root.get("type") != "CoulForce":
            raise ValueError("the XML does not hold a CoulForce (type=%r)" % root.get("type"))
        if int(root.get("version")) != 1:
            raise ValueError("Unsupported version number")
        f = cls()
        f.setForceGroup(int(root.get("forceGroup", 0)))
        f.setCutoffDistance(float(root.get("cutoff")))
        f.setEwaldErrorTolerance(float(root.get("ewaldTolerance")))
        f.setUsesPeriodicBoundaryConditions(int(root.get("usesPeriodic")) != 0)
        for p in root.find("Particles"):
            f.addParticle(float(p.get("q")), float(p.get("sig")), float(p.get("eps")))
        for e in root.find("Exceptions"):
            f.addException(int(e.get("p1")), int(e.get("p2")))
        for b in root.find("FluxBonds"):
            f.addFluxBond(int(b.get("p1")), int(b.get("p2")), float(b.get("k")), float(b.get("b")))
        for a in root.find("FluxAngles"):
            f.addFluxAngle(int(a.get("p1")), int(a.get("p2")), int(a.get("p3")), float(a.get("k")), float(a.get("theta")))
        for w in root.find("FluxWaters"):
            f.addFluxWater(int(w.get("po")), int(w.get("ph1")), int(w.get("ph2")), float(w.get("k1")), float(w.get("k2")),
                           float(w.get("kub")), float(w.get("b0")), float(w.get("ub0")))
        return f

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
