"""Host-side binding of the ``CalcCoulForce`` kernel to libcfx_b200.so (include/cfx_b200.h).

Mirrors the plugin's kernel interface (openmmapi/include/CoulKernels.h:15-38):

    kernel = CalcCoulForceKernel()
    kernel.initialize(default_box, force)                    # initialize(system, force)
    energy = kernel.execute(positions, box, forces_out, includeForces, includeEnergy)

There is no CPU fallback: importing works anywhere, but constructing a kernel without the built CUDA
library or without a B200 raises.
"""
import ctypes as C
import os

import numpy as np

from . import _abi

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libcfx_b200.so")
_lib = None


class CfxError(RuntimeError):
    """Raised for any non-zero status of the C ABI (the C++ adapter throws OpenMMException instead)."""


def load_library():
    """dlopen libcfx_b200.so and declare argument types. Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CfxError("libcfx_b200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                       "there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    lib.cfx_last_error.restype = C.c_char_p
    lib.cfx_create.argtypes = [C.POINTER(_abi.SystemDesc), C.POINTER(_abi.Options), C.POINTER(C.c_void_p)]
    lib.cfx_destroy.argtypes = [C.c_void_p]
    lib.cfx_destroy.restype = None
    # plain addresses: building a typed ctypes pointer from a numpy array costs ~5 us each, the per-step call passes four
    lib.cfx_execute.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.cfx_execute_device.argtypes = [C.c_void_p, C.c_void_p, _abi.c_double_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p]
    lib.cfx_execute_shard.argtypes = [C.c_void_p, C.c_void_p, _abi.c_double_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.cfx_execute_sharded.argtypes = [C.c_void_p, C.c_void_p, _abi.c_double_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.cfx_comm_get_unique_id.argtypes = [C.c_char_p]
    lib.cfx_comm_init.argtypes = [C.c_void_p, C.c_char_p]
    lib.cfx_comm_size.argtypes = [C.c_void_p]
    lib.cfx_multi_create.argtypes = [C.POINTER(_abi.SystemDesc), _abi.c_int32_p, C.c_int32, C.POINTER(C.c_void_p)]
    lib.cfx_multi_destroy.argtypes = [C.c_void_p]
    lib.cfx_multi_destroy.restype = None
    lib.cfx_multi_num_devices.argtypes = [C.c_void_p]
    lib.cfx_multi_handle.argtypes = [C.c_void_p, C.c_int32]
    lib.cfx_multi_handle.restype = C.c_void_p
    lib.cfx_multi_execute.argtypes = [C.c_void_p, _abi.c_double_p, _abi.c_double_p, C.c_int, C.c_int, _abi.c_double_p, _abi.c_double_p]
    lib.cfx_padded_num_particles.argtypes = [C.c_void_p]
    lib.cfx_get_ewald_params.argtypes = [C.c_void_p, C.POINTER(_abi.EwaldParams)]
    lib.cfx_get_stats.argtypes = [C.c_void_p, C.POINTER(_abi.Stats)]
    lib.cfx_get_charges.argtypes = [C.c_void_p, _abi.c_double_p]
    lib.cfx_get_dedq.argtypes = [C.c_void_p, _abi.c_double_p]
    lib.cfx_num_jacobian_rows.argtypes = [C.c_void_p]
    lib.cfx_get_jacobian.argtypes = [C.c_void_p, _abi.c_int32_p, _abi.c_int32_p, _abi.c_double_p]
    lib.cfx_get_neighbor_pairs.argtypes = [C.c_void_p, _abi.c_int32_p, C.c_int64, C.POINTER(C.c_int64)]
    lib.cfx_get_exclusions.argtypes = [C.c_void_p, _abi.c_int32_p, _abi.c_int32_p, C.c_int64, C.POINTER(C.c_int64)]
    lib.cfx_time_device.argtypes = [C.c_void_p, C.c_void_p, _abi.c_double_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]
    lib.cfx_time_kernels.argtypes = [C.c_void_p, C.c_void_p, _abi.c_double_p, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int,
                                     C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_int)]
    lib.cfx_measure_fp32_peak.argtypes = [C.c_int, C.c_int, _abi.c_double_p, _abi.c_double_p]
    _lib = lib
    return lib


def _dp(a):
    return a.ctypes.data_as(_abi.c_double_p)


def _addr(a):
    """Address of a numpy array's buffer without going through ndarray.ctypes (the per-step call path)."""
    try:
        return C.addressof(C.c_char.from_buffer(a))
    except (TypeError, ValueError):        # read-only or otherwise unexportable buffer
        return a.ctypes.data


def _box9(box):
    b = np.asarray(box, dtype=np.float64)
    if b.shape == (3,):
        b = np.diag(b)
    return np.ascontiguousarray(b.reshape(9))


class CalcCoulForceKernel:
    """The B200 implementation of the plugin's ``CalcCoulForce`` kernel."""

    @staticmethod
    def Name():
        return "CalcCoulForce"      # CoulKernels.h:17-19

    def __init__(self, device=-1, shard_rank=0, shard_count=1, use_graph=True, pin_caller_buffers=False,
                 skip_discarded_energy=False, list_skin=None, kmax_follows_box=False):
        """list_skin: skin (nm) of the direct-space candidate lists (None = library default 0.1 nm, 0 = rebuild at every
        evaluation).
        pin_caller_buffers: the caller keeps the positions / forces arrays it passes to execute() alive until it passes
        different ones or closes the kernel; they are then page-locked in place (CFX_OPT_PIN_CALLER_BUFFERS).
        kmax_follows_box: re-derive kmax from the box of the call when it changes (CFX_OPT_KMAX_FOLLOWS_BOX); the default
        keeps the kmax of the default box for the life of the kernel, as the reference does."""
        self._lib = load_library()
        self._opts = _abi.Options(device=device, shard_rank=shard_rank, shard_count=shard_count,
                                  use_graph=1 if use_graph else 0,
                                  flags=(_abi.OPT_PIN_CALLER_BUFFERS if pin_caller_buffers else 0)
                                  | (_abi.OPT_SKIP_DISCARDED_ENERGY if skip_discarded_energy else 0)
                                  | (_abi.OPT_KMAX_FOLLOWS_BOX if kmax_follows_box else 0),
                                  list_skin_pm=0 if list_skin is None else (-1 if list_skin <= 0 else max(1, int(round(list_skin * 1e3)))))
        self._h = None
        self.num_particles = 0
        self._e5 = np.zeros(_abi.E_COUNT)       # per-call scratch of execute(), addresses taken once
        self._b9 = np.zeros(9)
        self._e5_addr, self._b9_addr = self._e5.ctypes.data, self._b9.ctypes.data

    def _check(self, code):
        if code != _abi.CFX_OK:
            raise CfxError(self._lib.cfx_last_error().decode())

    # CalcCoulForceKernel::initialize(const System&, const CoulForce&): the System contributes the
    # particle count and the default periodic box (ReferenceCoulKernels.cpp:231,400).
    def initialize(self, default_box, force):
        self.close()
        desc, keep = force.to_desc(_box9(default_box).reshape(3, 3))
        h = C.c_void_p()
        self._check(self._lib.cfx_create(C.byref(desc), C.byref(self._opts), C.byref(h)))
        self._h = h
        self.num_particles = force.getNumParticles()
        del keep

    def copyParametersToContext(self, default_box, force):
        """New parameter values for the same topology (what CoulForce.updateParametersInContext would reach; the
        reference has no such call). Index lists, exceptions, cutoff and tolerance are not re-read."""
        desc, keep = force.to_desc(_box9(default_box).reshape(3, 3))
        self._lib.cfx_update_parameters.argtypes = [C.c_void_p, C.c_void_p]
        self._check(self._lib.cfx_update_parameters(self._h, C.byref(desc)))
        del keep

    # CalcCoulForceKernel::execute(ContextImpl&, includeForces, includeEnergy): adds to `forces`
    # (a [N,3] float64 array, the platform's force vector) and returns the energy in kJ/mol.
    def execute(self, positions, box, forces=None, includeForces=True, includeEnergy=True, components=None):
        pos = np.ascontiguousarray(positions, dtype=np.float64)
        if pos.size != 3 * self.num_particles:
            raise CfxError("positions has %d values, expected %d" % (pos.size, 3 * self.num_particles))
        self._b9[:] = _box9(box)
        fptr = None
        if forces is not None:
            if forces.dtype != np.float64 or not forces.flags.c_contiguous or forces.size != pos.size:
                raise CfxError("forces must be a C-contiguous float64 array of shape [N,3]")
            fptr = _addr(forces)
        e5 = self._e5
        if self._lib.cfx_execute(self._h, _addr(pos), self._b9_addr, 1 if includeForces else 0, 1 if includeEnergy else 0,
                                 self._e5_addr, fptr) != _abi.CFX_OK:
            raise CfxError(self._lib.cfx_last_error().decode())
        if components is not None:
            components[:] = e5
        return float(e5[_abi.E_TOTAL])

    def execute_device(self, d_positions, box, d_force_fixed, d_dedq_fixed=0, d_energy=0, stream=0,
                       includeForces=True, includeEnergy=True):
        """Device-pointer entry (ints, e.g. ``tensor.data_ptr()``); asynchronous on ``stream``."""
        b = _box9(box)
        self._check(self._lib.cfx_execute_device(self._h, d_positions, _dp(b), int(includeForces), int(includeEnergy),
                                                 d_force_fixed, d_dedq_fixed or None, d_energy or None, stream or None))

    def execute_shard(self, d_positions, box, d_reduce, stream=0, includeForces=True, includeEnergy=True):
        """Sharded step: zero + fill the int64 reduction buffer [3*Npad + 8] (forces 2^32, energies 2^24); asynchronous."""
        self._check(self._lib.cfx_execute_shard(self._h, d_positions, _dp(_box9(box)), int(includeForces), int(includeEnergy),
                                                d_reduce, stream or None))

    # -- multi-GPU: one process per GPU, NCCL communicator owned by the handle (include/cfx_b200.h) ------------
    def comm_init(self, unique_id):
        """Collective over the shard_count ranks: `unique_id` is the 128-byte id rank 0 drew with comm_unique_id()."""
        self._check(self._lib.cfx_comm_init(self._h, bytes(unique_id)))

    def comm_size(self):
        return self._lib.cfx_comm_size(self._h)

    def execute_sharded(self, d_positions, box, d_reduce, stream=0, includeForces=True, includeEnergy=True):
        """execute_shard + in-place sum all-reduce of d_reduce over the communicator, one CUDA graph; asynchronous."""
        self._check(self._lib.cfx_execute_sharded(self._h, d_positions, _dp(_box9(box)), int(includeForces), int(includeEnergy),
                                                  d_reduce, stream or None))

    # -- derived parameters, counters, parity getters --------------------------------------------
    def padded_num_particles(self):
        return self._lib.cfx_padded_num_particles(self._h)

    def ewald_params(self):
        p = _abi.EwaldParams()
        self._check(self._lib.cfx_get_ewald_params(self._h, C.byref(p)))
        return p.alpha, tuple(p.kmax), p.num_kvectors

    def stats(self):
        s = _abi.Stats()
        self._check(self._lib.cfx_get_stats(self._h, C.byref(s)))
        return s

    def charges(self):
        q = np.zeros(self.num_particles)
        self._check(self._lib.cfx_get_charges(self._h, _dp(q)))
        return q

    def dedq(self):
        v = np.zeros(self.num_particles)
        self._check(self._lib.cfx_get_dedq(self._h, _dp(v)))
        return v

    def jacobian(self):
        p = self._lib.cfx_num_jacobian_rows(self._h)
        dq, dx, val = np.zeros(p, np.int32), np.zeros(p, np.int32), np.zeros(3 * p)
        self._check(self._lib.cfx_get_jacobian(self._h, dq.ctypes.data_as(_abi.c_int32_p), dx.ctypes.data_as(_abi.c_int32_p), _dp(val)))
        return dq, dx, val.reshape(-1, 3)

    def neighbor_pairs(self):
        cnt = C.c_int64(0)
        self._check(self._lib.cfx_get_neighbor_pairs(self._h, None, 0, C.byref(cnt)))
        pairs = np.zeros(2 * cnt.value, np.int32)
        self._check(self._lib.cfx_get_neighbor_pairs(self._h, pairs.ctypes.data_as(_abi.c_int32_p), cnt.value, C.byref(cnt)))
        return pairs.reshape(-1, 2)

    def exclusions(self):
        cnt = C.c_int64(0)
        ptr = np.zeros(self.num_particles + 1, np.int32)
        self._check(self._lib.cfx_get_exclusions(self._h, ptr.ctypes.data_as(_abi.c_int32_p), None, 0, C.byref(cnt)))
        cols = np.zeros(cnt.value, np.int32)
        self._check(self._lib.cfx_get_exclusions(self._h, ptr.ctypes.data_as(_abi.c_int32_p), cols.ctypes.data_as(_abi.c_int32_p),
                                                 cnt.value, C.byref(cnt)))
        return ptr, cols

    # -- timing helpers (bench.py) ------------------------------------------------------------------
    def time_device(self, d_positions, box, iters, includeForces=True, includeEnergy=True):
        ms = C.c_float(0)
        self._check(self._lib.cfx_time_device(self._h, d_positions, _dp(_box9(box)), int(includeForces), int(includeEnergy),
                                              int(iters), C.byref(ms)))
        return ms.value

    def time_kernels(self, d_positions, box, iters, includeForces=True, includeEnergy=True):
        names = C.create_string_buffer(4096)
        ms = (C.c_float * 64)()
        cnt = C.c_int(0)
        self._check(self._lib.cfx_time_kernels(self._h, d_positions, _dp(_box9(box)), int(includeForces), int(includeEnergy), int(iters), names, 4096, ms, 64, C.byref(cnt)))
        return dict(zip(names.value.decode().split(";"), [ms[i] for i in range(cnt.value)]))

    def close(self):
        if self._h is not None:
            self._lib.cfx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def comm_unique_id():
    """128-byte NCCL id for comm_init (rank 0 draws it, the launcher broadcasts it)."""
    lib = load_library()
    buf = C.create_string_buffer(_abi.COMM_ID_BYTES)
    if lib.cfx_comm_get_unique_id(buf) != 0:
        raise CfxError(lib.cfx_last_error().decode())
    return buf.raw


class MultiGpuCoulKernel:
    """One process, several GPUs (cfx_multi_*): the kernel a plugin inside a single OpenMM process would hold.
    ``execute`` has the contract of CalcCoulForceKernel.execute."""

    def __init__(self, default_box, force, devices):
        self._lib = load_library()
        desc, keep = force.to_desc(_box9(default_box).reshape(3, 3))
        devs = np.asarray(devices, dtype=np.int32)
        h = C.c_void_p()
        if self._lib.cfx_multi_create(C.byref(desc), devs.ctypes.data_as(_abi.c_int32_p), len(devs), C.byref(h)) != 0:
            raise CfxError(self._lib.cfx_last_error().decode())
        self._m = h
        self.num_particles = force.getNumParticles()
        del keep

    def execute(self, positions, box, forces=None, includeForces=True, includeEnergy=True, components=None):
        pos = np.ascontiguousarray(positions, dtype=np.float64).reshape(-1)
        e5 = np.zeros(_abi.E_COUNT)
        fptr = _dp(forces) if forces is not None else None
        if self._lib.cfx_multi_execute(self._m, _dp(pos), _dp(_box9(box)), int(includeForces), int(includeEnergy), _dp(e5), fptr) != 0:
            raise CfxError(self._lib.cfx_last_error().decode())
        if components is not None:
            components[:] = e5
        return float(e5[_abi.E_TOTAL])

    def close(self):
        if self._m is not None:
            self._lib.cfx_multi_destroy(self._m)
            self._m = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def measure_tf32_peak(device=-1, iters=20000):
    """Dense TF32 tcgen05 throughput in TFLOP/s (roofline denominator of the tensor-core k-space kernels)."""
    lib = load_library()
    lib.cfx_measure_tf32_peak.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double)]
    tf = C.c_double(0.0)
    if lib.cfx_measure_tf32_peak(int(device), int(iters), C.byref(tf)) != 0:
        raise CfxError(lib.cfx_last_error().decode())
    return tf.value


def measure_i8_peak(device=-1, iters=20000):
    """Dense INT8 tcgen05 throughput in TOP/s (roofline denominator of the integer structure-factor kernel)."""
    lib = load_library()
    lib.cfx_measure_i8_peak.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double)]
    tops = C.c_double(0.0)
    if lib.cfx_measure_i8_peak(int(device), int(iters), C.byref(tops)) != 0:
        raise CfxError(lib.cfx_last_error().decode())
    return tops.value


def measure_fp32_peak(device=-1, iters=5):
    lib = load_library()
    tf, mhz = C.c_double(0), C.c_double(0)
    if lib.cfx_measure_fp32_peak(device, iters, C.byref(tf), C.byref(mhz)) != 0:
        raise CfxError(lib.cfx_last_error().decode())
    return tf.value, mhz.value


class CoulContext:
    """Minimal stand-in for ``openmm.Context`` around one CoulForce: owns the kernel, evaluates
    energy and forces for given positions (what ``context.getState(getEnergy=True, getForces=True)``
    returns for a System whose only force is the CoulForce)."""

    def __init__(self, force, box, device=-1, **kernel_options):
        self.force = force
        self.box = np.asarray(box, dtype=np.float64)
        self.kernel = CalcCoulForceKernel(device=device, **kernel_options)
        self.kernel.initialize(self.box, force)

    def evaluate(self, positions, includeForces=True, includeEnergy=True):
        forces = np.zeros((self.force.getNumParticles(), 3))
        comps = np.zeros(_abi.E_COUNT)
        e = self.kernel.execute(positions, self.box, forces, includeForces, includeEnergy, components=comps)
        return e, forces, comps
