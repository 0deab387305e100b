"""ctypes binding of libmgp_b200.so (include/mgp.h).  Thin: argument marshalling + error mapping."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmgp_b200.so")

MGP_OK, MGP_EINVAL, MGP_ECUDA, MGP_ESTATE, MGP_ENOMEM, MGP_ENOTPD, MGP_ENCCL = 0, -1, -2, -3, -4, -5, -6
CORR = {"squared_exponential": 0, "matern": 1, "absolute_exponential": 2}
ACQ = {"EI": 0, "PI": 1, "EpsilonPI": 1, "UCB": 2, "MGFI": 3}
MODE = {"noiseless": 0, "noisy": 1, "noise_estim": 2}
LIK_RESTRICTED = 16
TREND = {("constant", False): 0, ("constant", True): 1, ("linear", False): 2, ("linear", True): 3}
MAX_SETS = 16
ABI_VERSION = 4

c_dp = ctypes.POINTER(ctypes.c_double)
c_ip = ctypes.POINTER(ctypes.c_int)
c_lp = ctypes.POINTER(ctypes.c_int64)
_vp, _i, _l, _d = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_double

# name -> (restype, argtypes): every symbol include/mgp.h declares
SIGNATURES = {
    "mgp_abi_version": (_i, []),
    "mgp_create": (_i, [_i, ctypes.POINTER(_vp)]),
    "mgp_destroy": (_i, [_vp]),
    "mgp_last_error": (ctypes.c_char_p, [_vp]),
    "mgp_set_train": (_i, [_vp, _i, _i, c_dp, c_dp, _i, _i, c_dp]),
    "mgp_nll_batch": (_i, [_vp, _i, c_dp, _i, _i, _d, _i, c_dp, c_dp, c_ip]),
    "mgp_finalize_fit": (_i, [_vp, c_dp, _i, _i, _d, c_dp]),
    "mgp_append_train": (_i, [_vp, _i, c_dp, c_dp, c_dp]),
    "mgp_nll_env": (_i, [_vp, c_dp, _i, _i, _d, c_dp, c_dp, c_dp, c_dp, c_dp]),
    "mgp_get_state": (_i, [_vp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp]),
    "mgp_set_state": (_i, [_vp, c_dp, c_dp, c_dp, _d, c_dp]),
    "mgp_get_trend": (_i, [_vp, c_dp]),
    "mgp_predict": (_i, [_vp, _l, c_dp, c_dp, c_dp]),
    "mgp_gradient": (_i, [_vp, _l, c_dp, c_dp, c_dp]),
    "mgp_acq": (_i, [_vp, _l, c_dp, _i, _i, c_dp, _d, _i, c_dp, c_dp]),
    "mgp_acq_topk": (_i, [_vp, _l, c_dp, _i, _i, c_dp, _d, _i, _i, c_dp, c_lp]),
    "mgp_predict_dev": (_i, [_vp, _l, _vp, _vp, _vp, _vp]),
    "mgp_acq_dev": (_i, [_vp, _l, _vp, _i, _i, c_dp, _d, _i, _vp, _vp, _vp]),
    "mgp_acq_topk_dev": (_i, [_vp, _l, _vp, _i, _i, c_dp, _d, _i, _i, _vp, _vp, _vp]),
    "mgp_nll_batch_dev": (_i, [_vp, _i, _vp, _i, _i, _d, _i, _vp, _vp, _vp, _vp]),
    "mgp_comm_unique_id": (_i, [ctypes.c_char_p]),
    "mgp_comm_init": (_i, [_vp, _i, _i, ctypes.c_char_p]),
    "mgp_comm_destroy": (_i, [_vp]),
    "mgp_comm_info": (ctypes.c_char_p, [_vp]),
    "mgp_bcast_model": (_i, [_vp, _i]),
    "mgp_allgather_topk": (_i, [_vp, _i, _i, c_dp, c_lp, _l, c_dp, c_lp]),
    "mgp_allgather_topk_dev": (_i, [_vp, _i, _i, _vp, _vp, _l, _vp, _vp, _vp]),
    "mgp_allgather_f64_dev": (_i, [_vp, _vp, _l, _vp, _vp]),
    "mgp_set_option": (_i, [_vp, ctypes.c_char_p, _l]),
    "mgp_get_counter": (_l, [_vp, ctypes.c_char_p]),
}

_lib = None


class MgpError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libmgp_b200 error %d: %s" % (code, msg))
        self.code = code


def load():
    """Load the shared library (once).  Raises if it has not been built: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libmgp_b200.so not found at %s -- build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (needs nvcc); mipego_b200 has no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.mgp_abi_version() != ABI_VERSION:
        raise ImportError("libmgp_b200.so ABI version mismatch")
    _lib = lib
    return lib


def dptr(a):
    return a.ctypes.data_as(c_dp) if a is not None else None


def as_f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


class Handle:
    """Owns one mgp_t*.  One per model object; not thread-safe."""

    def __init__(self, device=None):
        lib = load()
        if device is None:
            device = int(os.environ.get("MGP_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        h = _vp()
        rc = lib.mgp_create(int(device), ctypes.byref(h))
        if rc != 0:
            raise MgpError(rc, "mgp_create failed on device %d: no usable sm_100 CUDA device "
                               "(mipego_b200 has no CPU fallback)" % device)
        self.lib, self.h, self.device = lib, h, int(device)

    def check(self, rc):
        if rc != 0:
            raise MgpError(rc, self.lib.mgp_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.mgp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def counter(self, name):
        return int(self.lib.mgp_get_counter(self.h, name.encode()))

    def set_option(self, name, value):
        self.check(self.lib.mgp_set_option(self.h, name.encode(), int(value)))
