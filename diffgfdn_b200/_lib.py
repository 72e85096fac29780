"""ctypes binding of libdiffgfdn_b200.so (the C ABI declared in include/diffgfdn_b200.h).

There is no CPU fallback: if the library cannot be loaded, or a kernel reports an error, a RuntimeError is raised."""
import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_void_p, POINTER

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libdiffgfdn_b200.so")

_lib = None

# name -> (restype, argtypes); mirrors include/diffgfdn_b200.h one to one
SIGNATURES = {
    "dgfdn_last_error": (c_char_p, []),
    "dgfdn_version": (c_int, []),
    "dgfdn_sm_count": (c_int, []),
    "dgfdn_copy_rows_h2d": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_void_p]),
    "dgfdn_skew_expm_fwd": (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "dgfdn_skew_expm_bwd": (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dgfdn_coupled_feedback_fwd": (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dgfdn_coupled_feedback_bwd": (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dgfdn_solve_factors_bytes": (c_int64, [c_int, c_int64]),
    "dgfdn_solve_fwd": (c_int, [c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dgfdn_solve_bwd_ws_bytes": (c_int64, [c_int]),
    "dgfdn_solve_bwd": (c_int, [c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p]),
    "dgfdn_solve_groups_factors_bytes": (c_int64, [c_int, c_int, c_int64]),
    "dgfdn_solve_groups_fwd": (c_int, [c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_void_p]),
    "dgfdn_solve_groups_bwd_ws_bytes": (c_int64, [c_int]),
    "dgfdn_solve_groups_bwd": (c_int, [c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_void_p]),
    "dgfdn_solve_fir_fwd": (c_int, [c_int, c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dgfdn_solve_fir_bwd": (c_int, [c_int, c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_void_p]),
    "dgfdn_peer_state_bytes": (c_int64, []),
    "dgfdn_peer_flags_bytes": (c_int64, []),
    "dgfdn_peer_push": (c_int, [c_int, c_int, c_int, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_int64,
                                c_void_p]),
    "dgfdn_peer_gather": (c_int, [c_int, c_int, c_int, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_void_p,
                                  c_int64, c_void_p]),
    "dgfdn_peer_reduce": (c_int, [c_int, c_int, c_int, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_void_p,
                                  c_int64, c_void_p]),
    "dgfdn_solve_colorless_ws_bytes": (c_int64, [c_int]),
    "dgfdn_solve_colorless": (c_int, [c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dgfdn_project_fwd": (c_int, [c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64,
                                  c_void_p]),
    "dgfdn_project_bwd": (c_int, [c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int,
                                  c_void_p, c_void_p]),
    "dgfdn_project_svf_fwd": (c_int, [c_int, c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                      c_void_p, c_int64, c_void_p]),
    "dgfdn_project_svf_bwd_ws_bytes": (c_int64, [c_int, c_int, c_int64, c_int64]),
    "dgfdn_project_svf_bwd": (c_int, [c_int, c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                      c_void_p, c_void_p, c_void_p, c_void_p]),
    "dgfdn_project_sh_fwd": (c_int, [c_int, c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dgfdn_project_sh_bwd": (c_int, [c_int, c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                     c_void_p, c_void_p]),
    "dgfdn_mix_channels": (c_int, [c_int, c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dgfdn_czt_plan_create": (c_int, [c_int64, c_int64, c_int64, POINTER(c_void_p)]),
    "dgfdn_czt_plan_destroy": (c_int, [c_void_p]),
    "dgfdn_czt_plan_mc": (c_int64, [c_void_p]),
    "dgfdn_irfft_window_fwd": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dgfdn_irfft_window_bwd": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_int64,
                                       c_void_p]),
    "dgfdn_edc_db": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    "dgfdn_edc_loss_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    "dgfdn_edc_loss_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_double, c_void_p, c_void_p]),
    "dgfdn_edr_db": (c_int, [c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p]),
    "dgfdn_edr_ws_bytes": (c_int64, [c_int64, c_int64]),
    "dgfdn_edr_loss_fwd": (c_int, [c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dgfdn_edr_loss_bwd": (c_int, [c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dgfdn_td_edc_step": (c_int, [c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64,
                                  c_void_p, c_double, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "dgfdn_td_mix": (c_int, [c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64,
                             c_void_p]),
    "dgfdn_td_contract_ws_bytes": (c_int64, [c_int, c_int64, c_int64]),
    "dgfdn_td_contract": (c_int, [c_int, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_int, c_void_p,
                                  c_void_p]),
    "dgfdn_td_edc_fused_supported": (c_int, [c_int, c_int64]),
    "dgfdn_td_edc_fused_ws_bytes": (c_int64, [c_int, c_int64, c_int64]),
    "dgfdn_td_edc_fused_ws_init": (c_int, [c_void_p, c_int, c_int64, c_int64, c_void_p]),
    "dgfdn_td_edc_fused": (c_int, [c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64,
                                   c_void_p, c_double, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "dgfdn_td_edc_fused_info": (c_int, [c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dgfdn_mlp_supported": (c_int, [c_int, c_int, c_int, c_int, c_int]),
    "dgfdn_mlp_num_params": (c_int64, [c_int, c_int, c_int, c_int]),
    "dgfdn_mlp_bwd_ws_bytes": (c_int64, [c_int64, c_int, c_int, c_int, c_int]),
    "dgfdn_mlp_fwd": (c_int, [c_int64, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_int, c_void_p,
                              c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dgfdn_mlp_bwd": (c_int, [c_int64, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_int, c_void_p,
                              c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_void_p]),
    "dgfdn_colorless_fwd": (c_int, [c_int, c_int64, c_void_p, c_int, c_void_p, c_void_p]),
    "dgfdn_colorless_bwd": (c_int, [c_int, c_int64, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "dgfdn_render_groups": (c_int, [c_int, c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_void_p]),
    "dgfdn_render_mix": (c_int, [c_int, c_int, c_int64, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_void_p]),
}


def load():
    """Load the shared library (once). Raises RuntimeError with build instructions if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not found: build it with `python -m diffgfdn_b200.build` "
                           "(nvcc, sm_100a). diffgfdn_b200 has no CPU or eager fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI drifted from the header
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def call(name, *args):
    """Call an int-returning entry point and raise on a non-zero status."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed: {lib.dgfdn_last_error().decode()}")
