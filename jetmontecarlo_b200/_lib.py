"""ctypes binding of libjmc_b200.so (declared in include/jmc_b200.h).

There is no CPU fallback: if the shared library is missing or cannot be
loaded, importing this module raises.  Build it with
``python jetmontecarlo_b200/csrc/build.py`` (or ``__graft_entry__.build()``).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libjmc_b200.so")

ABI_VERSION = 2

# ---- enums (include/jmc_b200.h) ---------------------------------------------
LIN, LOG = 0, 1
QUARK, GLUON = 0, 1
ACC_LL, ACC_LL_NONSING, ACC_MLL = 0, 1, 2
BC_FIRST, BC_LAST_PLUS, BC_LAST_MINUS = 0, 1, 2
F64, F32 = 0, 1
RSS_PRECRIT, RSS_CRIT, RSS_SUB, RSS_FEW_PRES = 1, 2, 4, 8
SAMPLER_SIMPLE, SAMPLER_CRITICAL, SAMPLER_PRECRITICAL, SAMPLER_UNGROOMED = 0, 1, 2, 3
(FN_ALPHA_S, FN_SPLITTING, FN_RADIATOR_WEIGHT, FN_CRIT_PDF_FC, FN_CRIT_RAD_FC, FN_SUB_RAD_FC,
 FN_SUB_RADPRIME_FC, FN_SUB_PDF_FC, FN_PRE_RAD_FC, FN_CRIT_RAD_RC, FN_CRIT_RADPRIME_RC,
 FN_CRIT_PDF_RC, FN_SUB_RAD_RC, FN_SUB_RADPRIME_RC, FN_SUB_PDF_RC, FN_PRE_WEIGHT, FN_C_GROOMED,
 FN_C_UNGROOMED, FN_TWO_EMISSION, FN_PRE_WEIGHT_ZEROBIN, FN_ALPHA_SPLITTING, FN_MUL, FN_MASK_FINITE,
 FN_EWOC_WEIGHT, FN_SUB_SC1, FN_SUB_SC2, FN_SUB_SC3, FN_SUB_HC1, FN_SUB_HC2, FN_CRIT_SC1, FN_CRIT_SC2, FN_CRIT_SC3,
 FN_CRIT_HC1, FN_CRIT_HC2) = range(34)
KIND_RADIATOR, KIND_CRIT, KIND_CRIT_SUB, KIND_PRE_CRIT_SUB, KIND_SIMPLE, KIND_SHOWER = range(6)

SUDAKOV_PRECRITICAL, SUDAKOV_SUBSEQUENT = 1, 2
METHODS = {'lin': LIN, 'log': LOG}
JETS = {'quark': QUARK, 'gluon': GLUON}
ACCS = {'LL': ACC_LL, 'LL_nonsing': ACC_LL_NONSING, 'MLL': ACC_MLL}

EXPORTS = (
    "jmc_abi_version", "jmc_last_error", "jmc_device_count", "jmc_set_device", "jmc_sm_count",
    "jmc_sampler_area", "jmc_integrand_uniforms", "jmc_integrand_area", "jmc_sample", "jmc_uniforms", "jmc_inverse_transform",
    "jmc_eval_fn", "jmc_eval_integrand", "jmc_bin_index", "jmc_hist", "jmc_hist2d", "jmc_finalize2d", "jmc_minmax", "jmc_run", "jmc_run_batch", "jmc_finalize_batch", "jmc_dump", "jmc_shower_dump", "jmc_shower_chains",
    "jmc_finalize", "jmc_cumulative", "jmc_set_density_host", "jmc_integrate_host", "jmc_sample_host",
    "jmc_eval_fn_host", "jmc_launch_count", "jmc_create", "jmc_destroy", "jmc_handle_run", "jmc_handle_integrate_host",
    "jmc_chain_ecf", "jmc_process_sample", "jmc_process_pairs", "jmc_process_ewoc",
    "jmc_conditional_inverse_transform", "jmc_sudakov_generate", "jmc_fill", "jmc_memset", "jmc_tile", "jmc_count_finite", "jmc_pack_histograms", "jmc_pack_minmax",
)


class Sampler(C.Structure):
    _fields_ = [("kind", C.c_int32), ("method", C.c_int32), ("epsilon", C.c_double),
                ("zc", C.c_double), ("radius", C.c_double), ("lo", C.c_double), ("hi", C.c_double)]


class FnParams(C.Structure):
    _fields_ = [("jet", C.c_int32), ("acc", C.c_int32), ("fixed_coupling", C.c_int32),
                ("reserved", C.c_int32), ("z_cut", C.c_double), ("beta", C.c_double),
                ("max_radius", C.c_double), ("z_pre", C.c_double), ("f_soft", C.c_double),
                ("epsilon", C.c_double), ("s0", C.c_double), ("s1", C.c_double),
                ("s2", C.c_double), ("s3", C.c_double)]


class Integrand(C.Structure):
    _fields_ = [("kind", C.c_int32), ("method", C.c_int32), ("jet", C.c_int32),
                ("fixed_coupling", C.c_int32), ("obs_acc", C.c_int32), ("split_acc", C.c_int32),
                ("nan_to_num", C.c_int32), ("shower_groomer", C.c_int32),
                ("shower_n_emissions", C.c_int32), ("counts", C.c_int32),
                ("epsilon", C.c_double), ("z_cut", C.c_double), ("radius", C.c_double),
                ("beta", C.c_double), ("f_soft", C.c_double), ("z_pre", C.c_double),
                ("theta_scale", C.c_double), ("lo", C.c_double), ("hi", C.c_double),
                ("beta_sd", C.c_double), ("shower_cutoff", C.c_double),
                ("shower_rss_flags", C.c_int32), ("reserved", C.c_int32)]


class Process(C.Structure):
    _fields_ = [("kind", C.c_int32), ("method", C.c_int32), ("epsilon", C.c_double),
                ("lo", C.c_double * 4), ("hi", C.c_double * 4), ("energy", C.c_double), ("prefactor", C.c_double)]


PROCESS_EE2HADRONS_LO = 0
PAIR_COSTHETA, PAIR_THETA, PAIR_M2, PAIR_GECF, PAIR_Z_I, PAIR_Z_J = range(6)


class JmcError(RuntimeError):
    """A libjmc_b200 call failed (CUDA error, size limit, missing device)."""


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "jetmontecarlo_b200: %s is missing. This package has no CPU fallback; build the "
            "CUDA library with `python jetmontecarlo_b200/csrc/build.py`." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, u64, i64, i32, dbl = C.c_void_p, C.c_uint64, C.c_int64, C.c_int, C.c_double
    P = C.POINTER
    sig = {
        "jmc_abi_version": (i32, []),
        "jmc_last_error": (C.c_char_p, []),
        "jmc_device_count": (i32, []),
        "jmc_set_device": (i32, [i32]),
        "jmc_sm_count": (i32, []),
        "jmc_sampler_area": (i32, [P(Sampler), P(dbl)]),
        "jmc_integrand_uniforms": (i32, [P(Integrand)]),
        "jmc_integrand_area": (i32, [P(Integrand), P(dbl)]),
        "jmc_sample": (i32, [P(Sampler), u64, u64, u64, vp, vp, vp]),
        "jmc_uniforms": (i32, [u64, u64, u64, i32, vp, vp]),
        "jmc_inverse_transform": (i32, [vp, vp, i32, u64, u64, u64, vp, vp]),
        "jmc_eval_fn": (i32, [i32, P(FnParams), u64, vp, i64, vp, i64, vp, i64, vp, i64, vp, vp]),
        "jmc_eval_integrand": (i32, [P(Integrand), u64, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp]),
        "jmc_bin_index": (i32, [vp, u64, vp, i32, vp, vp]),
        "jmc_hist": (i32, [vp, vp, u64, vp, i32, vp, vp, vp, vp]),
        "jmc_hist2d": (i32, [vp, vp, vp, u64, vp, i32, vp, i32, vp, vp, vp, vp]),
        "jmc_finalize2d": (i32, [vp, vp, vp, vp, i32, vp, i32, dbl, u64, i32, dbl, vp, vp, vp, vp, vp]),
        "jmc_minmax": (i32, [vp, u64, vp, vp]),
        "jmc_run": (i32, [P(Integrand), u64, u64, u64, i32, vp, i32, vp, vp, vp, vp, vp]),
        "jmc_run_batch": (i32, [vp, i32, u64, u64, u64, i32, vp, i32, vp, vp, vp, vp, vp]),
        "jmc_finalize_batch": (i32, [i32, vp, vp, vp, vp, i32, vp, u64, i32, dbl, vp, vp, vp, vp, vp]),
        "jmc_dump": (i32, [P(Integrand), u64, u64, u64, i32, vp, vp, vp, vp]),
        "jmc_shower_dump": (i32, [P(Integrand), u64, u64, u64, i32, vp, vp, vp, vp]),
        "jmc_shower_chains": (i32, [P(Integrand), u64, u64, u64, i32, i32, vp, vp, vp]),
        "jmc_finalize": (i32, [vp, vp, vp, vp, i32, dbl, u64, i32, dbl, vp, vp, vp, vp, vp]),
        "jmc_cumulative": (i32, [vp, vp, vp, i32, i32, dbl, vp, vp, vp]),
        "jmc_set_density_host": (i32, [vp, vp, u64, vp, i32, dbl, i32, dbl, vp, vp, vp, vp, vp, vp, vp]),
        "jmc_integrate_host": (i32, [P(Integrand), u64, u64, u64, i32, vp, i32, i32, dbl,
                                     vp, vp, vp, vp, vp, vp, vp]),
        "jmc_sample_host": (i32, [P(Sampler), u64, u64, u64, vp, vp]),
        "jmc_eval_fn_host": (i32, [i32, P(FnParams), u64, vp, i64, vp, i64, vp, i64, vp, i64, vp]),
        "jmc_launch_count": (u64, []),
        "jmc_chain_ecf": (i32, [P(Integrand), vp, vp, u64, i32, vp, vp]),
        "jmc_process_sample": (i32, [P(Process), u64, u64, u64, vp, vp, vp, vp]),
        "jmc_process_pairs": (i32, [P(Process), vp, u64, i32, dbl, vp, vp]),
        "jmc_process_ewoc": (i32, [P(Process), u64, u64, u64, i32, dbl, dbl, dbl, vp, i32, vp, vp, vp, vp, vp, vp]),
        "jmc_conditional_inverse_transform": (i32, [vp, i32, vp, i32, vp, i32, vp, vp, vp, vp, u64, i32, i32, vp, vp]),
        "jmc_sudakov_generate": (i32, [i32, vp, i32, vp, i32, vp, i32, vp, dbl, vp, u64, u64, vp, vp, vp]),
        "jmc_fill": (i32, [vp, u64, dbl, vp]),
        "jmc_memset": (i32, [vp, i32, u64, vp]),
        "jmc_tile": (i32, [vp, u64, i32, vp, vp]),
        "jmc_count_finite": (i32, [vp, u64, P(u64), vp]),
        "jmc_pack_histograms": (i32, [vp, vp, vp, u64, vp, i32, vp]),
        "jmc_pack_minmax": (i32, [vp, u64, vp, i32, vp]),
        "jmc_create": (i32, [P(vp)]),
        "jmc_destroy": (i32, [vp]),
        "jmc_handle_run": (i32, [vp, P(Integrand), u64, u64, u64, i32, vp, i32, vp, vp, vp, vp, vp]),
        "jmc_handle_integrate_host": (i32, [vp, P(Integrand), u64, u64, u64, i32, vp, i32, i32, dbl,
                                            vp, vp, vp, vp, vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)       # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.jmc_abi_version() != ABI_VERSION:
        raise ImportError("libjmc_b200.so ABI version %d, expected %d" % (lib.jmc_abi_version(), ABI_VERSION))
    return lib


lib = _load()


def check(rc):
    """Raise on a negative return code with the library's message."""
    if rc < 0:
        msg = lib.jmc_last_error().decode("utf-8", "replace")
        raise JmcError("libjmc_b200 error %d: %s" % (rc, msg))
    return rc


def launch_count():
    return int(lib.jmc_launch_count())


class Handle:
    """``jmc_handle``: cached job description + device workspace + pinned staging (include/jmc_b200.h).  One per
    (thread, device); ``default_handle()`` keeps them."""

    def __init__(self):
        self._h = C.c_void_p()
        check(lib.jmc_create(C.byref(self._h)))

    @property
    def ptr(self):
        return self._h

    def close(self):
        if self._h:
            lib.jmc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_handles = {}


def default_handle(device_index):
    """the calling thread's handle for a device"""
    import threading
    key = (threading.get_ident(), int(device_index))
    h = _handles.get(key)
    if h is None:
        h = _handles[key] = Handle()
    return h
