"""ctypes binding of libmpcb200.so (include/mpcb200.h).  No torch types cross this boundary: device pointers are
passed as integers (`tensor.data_ptr()`), the stream as `torch.cuda.current_stream().cuda_stream`.

There is NO CPU fallback: if the library is missing or CUDA is unavailable, loading/creating raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MPCB200_LIB: development aid -- load another BUILD of the same library (e.g. a variant compiled with different flags for an
# A/B measurement).  There is still no fallback: whatever path is named must exist and export the ABI.
LIB_PATH = os.environ.get("MPCB200_LIB") or os.path.join(_HERE, "csrc", "libmpcb200.so")

F32, F64 = 0, 1
HESS_GAUSS_NEWTON, HESS_EXACT = 0, 1
ABI_VERSION = 4

# status codes (mirror test/FORCESNLPsolver/include/FORCESNLPsolver.h:70-106)
ST_OPTIMAL, ST_MAXIT, ST_STALLED, ST_NAN, ST_NOPROGRESS, ST_INFEASIBLE_X0 = 1, 0, 3, -6, -7, -8


class Config(C.Structure):
    """mpcb200_config"""
    _fields_ = ([(n, C.c_int32) for n in ("abi_version", "device", "N", "max_batch", "precision", "hessian",
                                          "max_iter", "ls_max")] +
                [("dt", C.c_double), ("l_wb", C.c_double), ("l_fric", C.c_double), ("Q", C.c_double * 5),
                 ("R", C.c_double * 2), ("deltav_min", C.c_double), ("deltav_max", C.c_double), ("a_max", C.c_double),
                 ("delta_min", C.c_double), ("delta_max", C.c_double), ("v_min", C.c_double), ("v_max", C.c_double),
                 ("r_sum", C.c_double), ("ego_offset", C.c_double), ("obstacle", C.c_double * 6),
                 ("mu0", C.c_double), ("mu_min", C.c_double), ("mu_factor", C.c_double), ("tol_step", C.c_double),
                 ("tol_feas", C.c_double), ("tau_min", C.c_double), ("bound_push", C.c_double), ("mu_min_alpha", C.c_double), ("mu_up_alpha", C.c_double), ("mu_up_factor", C.c_double), ("mu_max", C.c_double), ("mu_factor_full", C.c_double),
                 ("kappa_sigma", C.c_double), ("screen_inv_curv", C.c_double), ("trust_step", C.c_double), ("acc_factor", C.c_double),
                 ("acc_iters", C.c_int32), ("stall_iters", C.c_int32), ("refine_f64", C.c_int32), ("init_rollout", C.c_int32),
                 ("mu_warm", C.c_double), ("warm_push", C.c_double), ("kappa_warm", C.c_double), ("stiff_slack", C.c_double),
                 ("warm_duals", C.c_int32), ("warps_per_cta", C.c_int32), ("host_route", C.c_int32), ("host_chunks", C.c_int32)])


class Scenario(C.Structure):
    """mpcb200_scenario"""
    _fields_ = [("dt", C.c_double), ("Q", C.c_double * 5), ("R", C.c_double * 2), ("r_sum", C.c_double), ("obstacle", C.c_double * 6)]


EXPORTS = {
    "mpcb200_default_config": (None, [C.POINTER(Config), C.c_int32, C.c_int32]),
    "mpcb200_create": (C.c_int, [C.POINTER(Config), C.POINTER(C.c_void_p)]),
    "mpcb200_destroy": (None, [C.c_void_p]),
    "mpcb200_last_error": (C.c_char_p, [C.c_void_p]),
    "mpcb200_solve": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "mpcb200_solve_dual": (C.c_int, [C.c_void_p] * 7 + [C.c_int32, C.c_void_p]),
    "mpcb200_lam_words": (C.c_int32, [C.c_void_p]),
    "mpcb200_sqp_begin": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "mpcb200_sqp_iter": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "mpcb200_sqp_end": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mpcb200_plant_step_shift": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "mpcb200_build_ref_window": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_double,
                                           C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "mpcb200_closed_loop": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "mpcb200_solve_cold": (C.c_int, [C.c_void_p] * 6 + [C.c_int32, C.c_void_p]),
    "mpcb200_solve_host": (C.c_int, [C.c_void_p] * 8 + [C.c_int32]),
    "mpcb200_set_scenarios": (C.c_int, [C.c_void_p, C.POINTER(Scenario), C.c_int32]),
    "mpcb200_solve_scenarios": (C.c_int, [C.c_void_p] * 7 + [C.c_int32, C.c_void_p]),
    "mpcb200_forces_stage_eval": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "mpcb200_forces_solve": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)] + [C.c_void_p] * 6 + [C.c_int32, C.c_void_p]),
    "mpcb200_forces_set_road_boundaries": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_double]),
    "mpcb200_forces_closed_loop": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.c_int32] + [C.c_void_p] * 8 + [C.c_int32, C.c_void_p]),
    "mpcb200_launch_count": (C.c_int64, [C.c_void_p]),
    "mpcb200_workspace_words": (C.c_int32, [C.c_void_p]),
    "mpcb200_slab_in_smem": (C.c_int32, [C.c_void_p]),
    "mpcb200_abi_version": (C.c_int32, []),
}

_lib = None


class Mpcb200Error(RuntimeError):
    pass


def load():
    """dlopen libmpcb200.so and type every export declared in include/mpcb200.h.  Raises if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Mpcb200Error(f"{LIB_PATH} not built -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.mpcb200_abi_version() != ABI_VERSION:
            raise Mpcb200Error("libmpcb200.so ABI version mismatch")
        _lib = lib
    return _lib


def default_config(N, precision=F32):
    cfg = Config()
    load().mpcb200_default_config(C.byref(cfg), N, precision)
    return cfg


_INT_FIELDS = {n for n, t in Config._fields_ if t is C.c_int32}
_SCALAR_FIELDS = {n for n, t in Config._fields_ if t in (C.c_int32, C.c_double)}


def set_options(cfg, opts):
    """Apply solver options (fields of mpcb200_config) by name.  A name that is not a scalar field of the struct is an error:
    `setattr` on a ctypes.Structure would otherwise create a plain Python attribute and the option would be silently ignored."""
    for k, v in opts.items():
        if k not in _SCALAR_FIELDS:
            raise Mpcb200Error(f"unknown solver option {k!r}; mpcb200_config has: {sorted(_SCALAR_FIELDS)}")
        setattr(cfg, k, int(v) if k in _INT_FIELDS else float(v))
    return cfg


class Handle:
    """RAII wrapper of mpcb200_handle."""

    def __init__(self, cfg):
        self.lib = load()
        self.cfg = cfg
        h = C.c_void_p()
        rc = self.lib.mpcb200_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            msg = self.lib.mpcb200_last_error(None)
            raise Mpcb200Error(f"mpcb200_create failed ({rc}): {msg.decode() if msg else ''}")
        self.h = h

    def check(self, rc):
        if rc != 0:
            msg = self.lib.mpcb200_last_error(self.h)
            raise Mpcb200Error(f"libmpcb200 call failed ({rc}): {msg.decode() if msg else ''}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.mpcb200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launch_count(self):
        return int(self.lib.mpcb200_launch_count(self.h))
