"""ctypes view of include/cfx_b200.h (struct layouts and argument types).

The same ``cfx_system_desc`` block is consumed by the product library (libcfx_b200.so), by the CPU
oracle (oracle/libcfx_oracle.so) and by the reference harness (oracle/_ref/libcfx_ref.so), so the
three can be driven with identical inputs.
"""
import ctypes as C

CFX_OK, CFX_ERR_ARGUMENT, CFX_ERR_CUDA, CFX_ERR_STATE = 0, 1, 2, 3
OPT_PIN_CALLER_BUFFERS = 1
OPT_SKIP_DISCARDED_ENERGY = 2
OPT_KMAX_FOLLOWS_BOX = 4
COMM_ID_BYTES = 128
E_SELF, E_RECIP, E_DIRECT, E_EXCL, E_TOTAL, E_COUNT = 0, 1, 2, 3, 4, 5
ONE_4PI_EPS0 = 138.935456

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)


class SystemDesc(C.Structure):
    _fields_ = [
        ("num_particles", C.c_int32),
        ("charge", c_double_p), ("sigma", c_double_p), ("epsilon", c_double_p),
        ("num_exceptions", C.c_int32), ("exception_pairs", c_int32_p),
        ("num_flux_bonds", C.c_int32), ("flux_bond_idx", c_int32_p), ("flux_bond_params", c_double_p),
        ("num_flux_angles", C.c_int32), ("flux_angle_idx", c_int32_p), ("flux_angle_params", c_double_p),
        ("num_flux_waters", C.c_int32), ("flux_water_idx", c_int32_p), ("flux_water_params", c_double_p),
        ("cutoff", C.c_double), ("ewald_tol", C.c_double), ("use_pbc", C.c_int32),
        ("default_box", C.c_double * 9),
    ]


class Options(C.Structure):
    _fields_ = [("device", C.c_int32), ("shard_rank", C.c_int32), ("shard_count", C.c_int32),
                ("use_graph", C.c_int32), ("flags", C.c_int32), ("list_skin_pm", C.c_int32), ("reserved", C.c_int32 * 2)]


class EwaldParams(C.Structure):
    _fields_ = [("alpha", C.c_double), ("kmax", C.c_int32 * 3), ("num_kvectors", C.c_int64)]


class Stats(C.Structure):
    _fields_ = [("pairs_in_cutoff", C.c_int64), ("pair_candidates", C.c_int64), ("kernel_launches", C.c_int64),
                ("cells", C.c_int32 * 3), ("longest_pair_list", C.c_int32), ("pair_list_builds", C.c_int64)]


# every symbol include/cfx_b200.h declares (tests check the built library exports all of them)
EXPORTED_SYMBOLS = [
    "cfx_last_error", "cfx_device_count", "cfx_create", "cfx_destroy", "cfx_update_parameters", "cfx_execute", "cfx_execute_device", "cfx_execute_shard", "cfx_execute_platform",
    "cfx_comm_get_unique_id", "cfx_comm_init", "cfx_comm_size", "cfx_execute_sharded",
    "cfx_multi_create", "cfx_multi_destroy", "cfx_multi_num_devices", "cfx_multi_handle", "cfx_multi_execute",
    "cfx_padded_num_particles", "cfx_get_ewald_params", "cfx_get_stats", "cfx_get_charges", "cfx_get_dedq",
    "cfx_num_jacobian_rows", "cfx_get_jacobian", "cfx_get_neighbor_pairs", "cfx_get_exclusions",
    "cfx_time_device", "cfx_time_kernels", "cfx_measure_fp32_peak", "cfx_measure_tf32_peak", "cfx_measure_i8_peak",
    "cfx_md_create", "cfx_md_destroy", "cfx_md_set_state", "cfx_md_get_state", "cfx_md_minimize", "cfx_md_step", "cfx_md_energies",
]
