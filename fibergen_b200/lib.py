"""ctypes binding of libfgb200.so (include/fgb200.h + include/fgb200_lssolver.h).

The shared library is built in tree by ``make`` / ``__graft_entry__.build()``.  There is no CPU
fallback: if the library is missing or no sm_100 device is present every compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfgb200.so")

# error codes (fgb200.h)
FGB_OK, FGB_EINVAL, FGB_ENODEV, FGB_ENOMEM, FGB_ECUDA, FGB_EUNSUPPORTED, FGB_ENUMERIC, FGB_ECOMM = 0, -1, -2, -3, -4, -5, -6, -7
MODES = {"elasticity": 0, "hyperelasticity": 1, "viscosity": 2, "heat": 3, "porous": 4}
SCHEMES = {"collocated": 0, "staggered": 1, "willot": 2}
LAWS = {"iso": 0, "general": 1, "tiso": 2, "scalar": 3, "aniso3": 4, "svk": 5, "nh": 6, "nh2": 7}
MIXING = {"voigt": 0, "reuss": 1, "laminate": 2}

c_dp = C.POINTER(C.c_double)
c_dpp = C.POINTER(c_dp)
CALLBACK = C.CFUNCTYPE(C.c_int, C.c_void_p)

# name -> (restype, argtypes); this table is also what the "every declared symbol is exported" test walks
PROTOTYPES = {
    # ---- fgb200.h
    "fgb_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                             C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "fgb_destroy": (None, [C.c_void_p]),
    "fgb_last_error": (C.c_char_p, [C.c_void_p]),
    "fgb_version": (C.c_char_p, []),
    "fgb_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "fgb_synchronize": (C.c_int, [C.c_void_p]),
    "fgb_comm_unique_id": (C.c_int, [C.c_void_p]),
    "fgb_comm_init": (C.c_int, [C.c_void_p, C.c_void_p]),
    "fgb_local_nx": (C.c_int, [C.c_void_p]),
    "fgb_local_x0": (C.c_int, [C.c_void_p]),
    "fgb_plane_elems": (C.c_size_t, [C.c_void_p]),
    "fgb_dim": (C.c_int, [C.c_void_p]),
    "fgb_field_alloc": (C.c_int, [C.c_void_p]),
    "fgb_field_free": (C.c_int, [C.c_void_p, C.c_int]),
    "fgb_field_upload": (C.c_int, [C.c_void_p, C.c_int, c_dpp]),
    "fgb_field_download": (C.c_int, [C.c_void_p, C.c_int, c_dpp]),
    "fgb_field_device_ptr": (C.c_void_p, [C.c_void_p, C.c_int, C.c_int]),
    "fgb_set_num_phases": (C.c_int, [C.c_void_p, C.c_int]),
    "fgb_set_phase": (C.c_int, [C.c_void_p, C.c_int, c_dp]),
    "fgb_set_dfg": (C.c_int, [C.c_void_p, C.c_int]),
    "fgb_dfg_prolongate": (C.c_int, [C.c_void_p, C.c_int]),
    "fgb_dfg_restrict": (C.c_int, [C.c_void_p, C.c_int]),
    "fgb_set_law": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_dp, C.c_int]),
    "fgb_set_normals": (C.c_int, [C.c_void_p, c_dpp]),
    "fgb_set_orientation": (C.c_int, [C.c_void_p, c_dpp]),
    "fgb_set_mixing": (C.c_int, [C.c_void_p, C.c_int, c_dp, C.c_int]),
    "fgb_init_phase_capsules": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_double, c_dp, C.c_int, C.c_int]),
    "fgb_get_phase": (C.c_int, [C.c_void_p, C.c_int, c_dp]),
    "fgb_set_freq_hack": (C.c_int, [C.c_void_p, C.c_int]),
    "fgb_set_bc": (C.c_int, [C.c_void_p, c_dp, c_dp, C.c_double]),
    "fgb_set_constant": (C.c_int, [C.c_void_p, C.c_int, c_dp]),
    "fgb_add_constant": (C.c_int, [C.c_void_p, C.c_int, c_dp]),
    "fgb_copy": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "fgb_xpay": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int]),
    "fgb_xpaymz": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int]),
    "fgb_adjust_residual": (C.c_int, [C.c_void_p, C.c_int, c_dp, C.c_int]),
    "fgb_inner": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_dp]),
    "fgb_average": (C.c_int, [C.c_void_p, C.c_int, c_dp]),
    "fgb_component_dot": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_dp]),
    "fgb_mean_pk1": (C.c_int, [C.c_void_p, C.c_int, C.c_double, c_dp]),
    "fgb_mean_energy": (C.c_int, [C.c_void_p, C.c_int, c_dp]),
    "fgb_min_detF": (C.c_int, [C.c_void_p, C.c_int, c_dp]),
    "fgb_mean_cauchy": (C.c_int, [C.c_void_p, C.c_int, C.c_double, c_dp]),
    "fgb_calc_displacement": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double]),
    "fgb_g0div_hyper": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double]),
    "fgb_grad_hyper": (C.c_int, [C.c_void_p, C.c_int]),
    "fgb_grad_g0div_hyper": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double]),
    "fgb_calc_pressure": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double]),
    "fgb_extrapolate_polynomial": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), c_dp, c_dp, C.c_int]),
    "fgb_ref_material": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_dp, c_dp]),
    "fgb_calc_stress": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]),
    "fgb_calc_stress_deriv": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]),
    "fgb_calc_stress_const": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double]),
    "fgb_calc_polarization": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int]),
    "fgb_gamma": (C.c_int, [C.c_void_p, C.c_int, c_dp, C.c_double, C.c_double, C.c_double, C.c_double]),
    "fgb_div_staggered": (C.c_int, [C.c_void_p, C.c_int]),
    "fgb_g0_staggered": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_double]),
    "fgb_eps_staggered": (C.c_int, [C.c_void_p, C.c_int, c_dp]),
    "fgb_u_upload": (C.c_int, [C.c_void_p, c_dpp, C.c_int]),
    "fgb_u_download": (C.c_int, [C.c_void_p, c_dpp, C.c_int]),
    "fgb_fft_forward": (C.c_int, [C.c_void_p, C.c_int]),
    "fgb_fft_backward": (C.c_int, [C.c_void_p, C.c_int]),
    "fgb_basic_step": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_dp, C.c_double, C.c_double]),
    "fgb_polarization_step": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_dp, C.c_double, C.c_double]),
    "fgb_cg_apply": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, c_dp]),
    "fgb_cg_step": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, c_dp]),
    "fgb_cg_update": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, c_dp]),
    "fgb_cg_implicit_w_supported": (C.c_int, [C.c_void_p]),
    "fgb_cg_direction": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double]),
    "fgb_cg_tangent_prepare": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_double]),
    "fgb_cgdev_begin": (C.c_int, [C.c_void_p, C.c_double]),
    "fgb_cgdev_step": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double]),
    "fgb_cgdev_update": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "fgb_cgdev_wait": (C.c_int, [C.c_void_p, C.c_int, c_dp]),
    "fgb_check_numeric": (C.c_int, [C.c_void_p]),
    "fgb_launch_count": (C.c_uint64, [C.c_void_p]),
    "fgb_launch_count_reset": (None, [C.c_void_p]),
    "fgb_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "fgb_profile_get": (C.c_int, [C.c_void_p, C.c_char_p, c_dp, C.POINTER(C.c_uint64)]),
    "fgb_profile_names": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    # ---- fgb200_lssolver.h
    "fgls_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                              C.c_int, C.c_int, C.c_int]),
    "fgls_destroy": (None, [C.c_void_p]),
    "fgls_last_error": (C.c_char_p, [C.c_void_p]),
    "fgls_set": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p]),
    "fgls_add_material": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p, c_dp, C.c_int]),
    "fgls_set_reference": (C.c_int, [C.c_void_p, C.c_double, C.c_double]),
    "fgls_init": (C.c_int, [C.c_void_p]),
    "fgls_init_comm": (C.c_int, [C.c_void_p, C.c_void_p]),
    "fgls_set_phase": (C.c_int, [C.c_void_p, C.c_int, c_dp]),
    "fgls_init_phase_capsules": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "fgls_get_phase": (C.c_int, [C.c_void_p, C.c_int, c_dp]),
    "fgls_set_normals": (C.c_int, [C.c_void_p, c_dpp]),
    "fgls_set_orientation": (C.c_int, [C.c_void_p, c_dpp]),
    "fgls_set_strain": (C.c_int, [C.c_void_p, c_dp]),
    "fgls_set_stress": (C.c_int, [C.c_void_p, c_dp]),
    "fgls_set_bc_projector": (C.c_int, [C.c_void_p, c_dp]),
    "fgls_set_callback": (C.c_int, [C.c_void_p, CALLBACK, C.c_void_p]),
    "fgls_run": (C.c_int, [C.c_void_p]),
    "fgls_cancel": (C.c_int, [C.c_void_p]),
    "fgls_num_residuals": (C.c_int, [C.c_void_p]),
    "fgls_get_residuals": (C.c_int, [C.c_void_p, c_dp, C.c_int]),
    "fgls_mean_stress": (C.c_int, [C.c_void_p, c_dp]),
    "fgls_mean_strain": (C.c_int, [C.c_void_p, c_dp]),
    "fgls_mean_energy": (C.c_int, [C.c_void_p, c_dp]),
    "fgls_mean_cauchy_stress": (C.c_int, [C.c_void_p, c_dp]),
    "fgls_field_components": (C.c_int, [C.c_void_p, C.c_char_p]),
    "fgls_effective_properties": (C.c_int, [C.c_void_p, c_dp]),
    "fgls_get_field": (C.c_int, [C.c_void_p, C.c_char_p, c_dpp]),
    "fgls_ref_material": (C.c_int, [C.c_void_p, c_dp, c_dp]),
    "fgls_calc_ref_material": (C.c_int, [C.c_void_p]),
    "fgls_bc_matrices": (C.c_int, [C.c_void_p, c_dp, c_dp]),
    "fgls_dim": (C.c_int, [C.c_void_p]),
    "fgls_local_nx": (C.c_int, [C.c_void_p]),
    "fgls_solve_time": (C.c_double, [C.c_void_p]),
    "fgls_launches": (C.c_uint64, [C.c_void_p]),
    "fgls_ctx": (C.c_void_p, [C.c_void_p]),
}



class Capsule(C.Structure):
    """fgb_capsule (fgb200.h): a <place_fiber> capsule"""
    _fields_ = [("c", C.c_double * 3), ("a", C.c_double * 3), ("L0", C.c_double), ("R", C.c_double), ("material", C.c_int)]


_lib = None


class FgbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("fgb200 error %d: %s" % (code, msg))
        self.code = code
        self.message = msg


def load():
    """dlopen libfgb200.so and attach the prototypes; raises if the extension was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libfgb200.so is missing (%s): run `make` or __graft_entry__.build(); "
                          "there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
