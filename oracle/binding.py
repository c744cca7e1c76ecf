"""ctypes bindings for the two CPU checkers. TEST INFRASTRUCTURE, NOT PRODUCT.

* :class:`Oracle`         -- oracle/libcfx_oracle.so, the restatement (oracle/cfx_oracle.cpp).
* :class:`ReferenceBuild` -- oracle/_ref/libcfx_ref.so, the plugin's own sources compiled against the
  OpenMM stand-in (oracle/ref_harness.cpp). Exists only where it was built (this container).

Both take a filled ``CoulForce`` (openmm_chargeflux_b200.force) and evaluate
``execute(positions, box, includeForces, includeEnergy) -> (energy[5], forces)``.
"""
import ctypes as C
import os

import numpy as np

from openmm_chargeflux_b200 import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_LIB = os.path.join(_HERE, "libcfx_oracle.so")
REF_LIB = os.path.join(_HERE, "_ref", "libcfx_ref.so")
REF_PLUGIN = os.path.join(_HERE, "_ref", "libOpenMMCoulReference.so")


def oracle_available():
    return os.path.exists(ORACLE_LIB)


def reference_available():
    return os.path.exists(REF_LIB) and os.path.exists(REF_PLUGIN)


def _dp(a):
    return a.ctypes.data_as(_abi.c_double_p)


def _ip(a):
    return a.ctypes.data_as(_abi.c_int32_p)


class _Base:
    prefix = None
    lib = None

    def _fn(self, name):
        return getattr(self.lib, self.prefix + name)

    def _check(self, code):
        if code != 0:
            err = self._fn("last_error")
            err.restype = C.c_char_p
            raise RuntimeError("%s: %s" % (self.prefix, err().decode()))

    def execute(self, positions, box, include_forces=True, include_energy=True, forces_in=None):
        pos = np.ascontiguousarray(positions, dtype=np.float64).reshape(-1)
        box = np.ascontiguousarray(box, dtype=np.float64).reshape(9)
        energy = np.zeros(_abi.E_COUNT)
        forces = np.zeros(3 * self.n) if forces_in is None else np.array(forces_in, dtype=np.float64).reshape(-1)
        self._check(self._fn("execute")(self.h, _dp(pos), _dp(box), int(include_forces), int(include_energy),
                                        _dp(energy), _dp(forces)))
        return energy, forces.reshape(-1, 3)

    def ewald_params(self):
        p = _abi.EwaldParams()
        self._check(self._fn("get_ewald_params")(self.h, C.byref(p)))
        return p.alpha, tuple(p.kmax), p.num_kvectors

    def charges(self):
        q = np.zeros(self.n)
        self._check(self._fn("get_charges")(self.h, _dp(q)))
        return q

    def jacobian(self):
        f = self._fn("num_jacobian_rows")
        p = f(self.h)
        dq, dx, val = np.zeros(p, np.int32), np.zeros(p, np.int32), np.zeros(3 * p)
        self._check(self._fn("get_jacobian")(self.h, _ip(dq), _ip(dx), _dp(val)))
        return dq, dx, val.reshape(-1, 3)

    def neighbor_pairs(self):
        cnt = C.c_int64(0)
        self._check(self._fn("get_neighbor_pairs")(self.h, None, 0, C.byref(cnt)))
        pairs = np.zeros(2 * cnt.value, np.int32)
        self._check(self._fn("get_neighbor_pairs")(self.h, _ip(pairs), cnt.value, C.byref(cnt)))
        return pairs.reshape(-1, 2)

    def close(self):
        if getattr(self, "h", None):
            self._fn("destroy")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Oracle(_Base):
    prefix = "cfxo_"

    def __init__(self, force, default_box):
        if Oracle.lib is None:
            Oracle.lib = C.CDLL(ORACLE_LIB)
            Oracle.lib.cfxo_destroy.argtypes = [C.c_void_p]
        self.n = force.getNumParticles()
        desc, self._keep = force.to_desc(default_box)
        self.h = C.c_void_p()
        self._check(self.lib.cfxo_create(C.byref(desc), C.byref(self.h)))

    def dedq(self):
        v = np.zeros(self.n)
        self._check(self.lib.cfxo_get_dedq(self.h, _dp(v)))
        return v

    def set_kx_range(self, lo, hi):
        self.lib.cfxo_set_kx_range(self.h, int(lo), int(hi))

    def stats(self):
        s = _abi.Stats()
        self.lib.cfxo_get_stats(self.h, C.byref(s))
        return s


class ReferenceBuild(_Base):
    """The plugin's own Reference-platform kernel (or, with platform="B200" and the B200 plugin
    loaded, this repository's kernel behind the same unmodified CoulForce/CoulForceImpl)."""
    prefix = "cfxref_"

    def __init__(self, force, default_box, platform="Reference", plugin=None):
        if ReferenceBuild.lib is None:
            ReferenceBuild.lib = C.CDLL(REF_LIB)
            ReferenceBuild.lib.cfxref_destroy.argtypes = [C.c_void_p]
        self._check(self.lib.cfxref_load_plugin((plugin or REF_PLUGIN).encode()))
        self.n = force.getNumParticles()
        desc, self._keep = force.to_desc(default_box)
        self.h = C.c_void_p()
        self._check(self.lib.cfxref_create(C.byref(desc), platform.encode(), C.byref(self.h)))

    def to_xml(self):
        """XmlSerializer::serialize of the context's CoulForce (a loaded library must have registered a proxy for it)."""
        need = C.c_int64(0)
        self.lib.cfxref_serialize_xml.argtypes = [C.c_void_p, C.c_char_p, C.c_int64, C.POINTER(C.c_int64)]
        self._check(self.lib.cfxref_serialize_xml(self.h, None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value)
        self._check(self.lib.cfxref_serialize_xml(self.h, buf, need.value, C.byref(need)))
        return buf.value.decode()

    @classmethod
    def from_xml(cls, xml, default_box, platform="Reference", plugin=None):
        """A context around XmlSerializer::deserialize(xml)."""
        self = cls.__new__(cls)
        if ReferenceBuild.lib is None:
            ReferenceBuild.lib = C.CDLL(REF_LIB)
            ReferenceBuild.lib.cfxref_destroy.argtypes = [C.c_void_p]
        self._check(self.lib.cfxref_load_plugin((plugin or REF_PLUGIN).encode()))
        box = np.ascontiguousarray(default_box, dtype=np.float64).reshape(9)
        self.h = C.c_void_p()
        self._check(self.lib.cfxref_create_from_xml(xml.encode(), _dp(box), platform.encode(), C.byref(self.h)))
        self.lib.cfxref_num_particles.argtypes = [C.c_void_p]
        self.n = self.lib.cfxref_num_particles(self.h)
        return self
