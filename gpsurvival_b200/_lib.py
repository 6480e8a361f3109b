"""ctypes binding of include/gpsurv.h.  There is no fallback: if libgpsurv.so is missing or
cannot be loaded this module raises, and without an sm_100 GPU ``gps_create`` fails."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgpsurv.so")

GPS_OK, GPS_ERR_ARG, GPS_ERR_CUDA, GPS_ERR_NOT_PD, GPS_ERR_UNSUPPORTED, GPS_ERR_STATE = range(6)
COV = {"SqExp": 0, "ARD": 1, "Matern1": 2, "Matern3": 3, "Matern5": 4}


def cov_code(cov_form, extra_param=None):
    """covFuncForm (+ extraParam = maternParam for 'Matern', CovFunc.R:31) -> gps_cov_form."""
    if cov_form == "Matern":
        key = "Matern" + str(int(extra_param)) if extra_param is not None else "Matern?"
        if key not in COV:
            raise GpsError(GPS_ERR_ARG, f"Matern needs extraParam (maternParam) 1, 3 or 5, got {extra_param!r}")
        return COV[key]
    if cov_form not in COV:
        raise GpsError(GPS_ERR_UNSUPPORTED, f"covFuncForm {cov_form!r} is not built (SqExp, ARD, Matern)")
    return COV[cov_form]
MEAN = {"Zero": 0, "Linear": 1}
NOISE_CORR = {False: 0, None: 0, "none": 0, "noiseCorrLearned": 1, "noiseCorrVec": 2}
ARD_MODE = {"intended": 0, "as_written": 1}

_dp = C.POINTER(C.c_double)
_h = C.c_void_p

# name -> (restype, argtypes); every symbol include/gpsurv.h declares
PROTOTYPES = {
    "gps_version": (C.c_char_p, []),
    "gps_last_error": (C.c_char_p, [_h]),
    "gps_last_pivot": (C.c_int, [_h]),
    "gps_batch_pivots": (C.c_int, [_h, C.POINTER(C.c_int)]),
    "gps_set_near_pd": (C.c_int, [_h, C.c_int]),
    "gps_near_pd_used": (C.c_int, [_h, C.POINTER(C.c_int)]),
    "gps_create": (C.c_int, [C.c_int, C.POINTER(_h)]),
    "gps_create_batch": (C.c_int, [C.c_int, C.c_int, C.POINTER(_h)]),
    "gps_destroy": (C.c_int, [_h]),
    "gps_set_options": (C.c_int, [_h, C.c_int, C.c_int, C.c_int, C.c_int]),
    "gps_set_data": (C.c_int, [_h, _dp, C.c_int, C.c_int, _dp, _dp]),
    "gps_set_targets": (C.c_int, [_h, _dp]),
    "gps_set_noise_corr": (C.c_int, [_h, _dp, C.c_int]),
    "gps_cov": (C.c_int, [_h, _dp, C.c_int, _dp, C.c_int, C.c_int, C.c_double, _dp, C.c_int, C.c_int, C.c_int, _dp]),
    "gps_nlml": (C.c_int, [_h, _dp, C.c_int, _dp]),
    "gps_nlml_grad": (C.c_int, [_h, _dp, C.c_int, _dp, _dp]),
    "gps_regress_finalize": (C.c_int, [_h, _dp, C.c_int, _dp, _dp, _dp, _dp, _dp]),
    "gps_predict": (C.c_int, [_h, _dp, C.c_int, C.c_int, _dp, C.c_int, _dp, _dp, _dp]),
    "gps_adjust": (C.c_int, [_h, _dp, _dp, _dp, C.c_int, _dp, _dp]),
    "gps_metrics": (C.c_int, [_h, _dp, _dp, _dp, C.c_int, _dp, _dp, C.POINTER(C.c_longlong)]),
    "gps_synth_generate": (C.c_int, [_h, C.c_ulonglong, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int,
                                     C.c_double, C.c_double, _dp, _dp, _dp, _dp, C.POINTER(C.c_int)]),
    "gps_chol_matvec": (C.c_int, [_h, _dp, C.c_int, _dp, _dp]),
    "gps_fit_batch": (C.c_int, [_h, _dp, C.c_int, C.c_int, _dp, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int,
                                C.c_int, _dp, C.c_int, _dp, _dp, _dp, C.POINTER(C.c_int), _dp, _dp, _dp, _dp, _dp,
                                C.POINTER(C.c_longlong)]),
    "gps_fit_batch_status": (C.c_int, [_h, C.POINTER(C.c_int)]),
    "gps_set_data_indexed": (C.c_int, [_h, _dp, C.c_int, C.c_int, _dp, _dp, C.POINTER(C.c_int), C.c_int]),
    "gps_predict_rows": (C.c_int, [_h, _dp, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int, _dp, _dp, _dp]),
    "gps_partition_lpt": (C.c_int, [_dp, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "gps_fit_batch_multi": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                      _dp, C.c_int, C.c_int, _dp, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int,
                                      _dp, C.c_int, _dp, _dp, _dp, C.POINTER(C.c_int), _dp, _dp, _dp, _dp, _dp, C.POINTER(C.c_int),
                                      C.POINTER(C.c_longlong)]),
    "gps_fit_folds_multi": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                      _dp, C.c_int, C.c_int, _dp, _dp, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int), C.c_int,
                                      _dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int,
                                      _dp, _dp, _dp, C.POINTER(C.c_int), _dp, _dp, _dp, _dp, _dp, C.POINTER(C.c_int),
                                      C.POINTER(C.c_longlong)]),
    "gps_model_token": (C.c_ulonglong, [_h]),
    "gps_get_shape": (C.c_int, [_h, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "gps_launch_count": (C.c_longlong, [_h]),
    "gps_last_device_ms": (C.c_double, [_h]),
    "gps_set_graphs": (C.c_int, [_h, C.c_int]),
    "gps_time_eval": (C.c_int, [_h, C.c_int, C.c_int, _dp]),
    "gps_time_stages": (C.c_int, [_h, _dp]),
    "gps_bench_gemm": (C.c_int, [_h, C.c_int, C.c_int, C.c_int, C.c_int, _dp]),
    "gps_debug_panel_clocks": (C.c_int, [_h, C.POINTER(C.c_longlong), C.c_int]),
    "gps_multi_release": (C.c_int, []),
    "gps_debug_dag_trace": (C.c_int, [_h, C.c_int, C.POINTER(C.c_longlong), C.c_longlong]),
}

_lib = None


class GpsError(RuntimeError):
    def __init__(self, code, msg, pivot=0):
        super().__init__(f"libgpsurv error {code}: {msg}")
        self.code = code
        self.pivot = pivot


class NotPositiveDefinite(GpsError):
    """GPS_ERR_NOT_PD — lets a caller keep the reference's nearPD fallback host-side
    (ForOptimisationLog.R:43-51)."""


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -m gpsurvival_b200.build` "
            "(gpsurvival_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)     # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def dptr(a):
    return None if a is None else a.ctypes.data_as(_dp)


def f64(a, order="F"):
    """Contiguous FP64 copy in R (column-major) layout."""
    return np.require(np.asarray(a, dtype=np.float64), requirements=["F" if order == "F" else "C", "A"])


def check(rc, handle=None):
    if rc == GPS_OK:
        return
    lib = load()
    msg = lib.gps_last_error(handle)
    msg = msg.decode() if msg else ""
    if rc == GPS_ERR_NOT_PD:
        raise NotPositiveDefinite(rc, msg, lib.gps_last_pivot(handle) if handle else 0)
    raise GpsError(rc, msg)
