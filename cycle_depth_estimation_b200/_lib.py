"""ctypes binding of libcdb200.so (include/cdb200.h). Fails loudly when the library is missing."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libcdb200.so")

BF16, F32 = 0, 1
ACT_NONE, ACT_RELU, ACT_LEAKY, ACT_TANH, ACT_SIGMOID = 0, 1, 2, 3, 4
NORM_NONE, NORM_INSTANCE, NORM_BATCH = 0, 1, 2
PAD_ZERO, PAD_REFLECT = 0, 1


class CdbAct(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("c", C.c_int32),
                ("sn", C.c_int64), ("sh", C.c_int64), ("sw", C.c_int64), ("dtype", C.c_int32),
                ("reserved", C.c_int32)]


class CdbOut(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("c", C.c_int32),
                ("cstore", C.c_int32), ("dtype", C.c_int32),
                ("sn", C.c_int64), ("sh", C.c_int64), ("sw", C.c_int64), ("sc", C.c_int64)]


class CdbConvGeom(C.Structure):
    _fields_ = [("r", C.c_int32), ("s", C.c_int32), ("stride", C.c_int32), ("pad_h", C.c_int32),
                ("pad_w", C.c_int32), ("dil", C.c_int32), ("transposed", C.c_int32), ("rowpack", C.c_int32),
                ("flip", C.c_int32), ("reserved", C.c_int32)]


class CdbEpilogue(C.Structure):
    _fields_ = [("bias", C.c_void_p), ("act", C.c_int32), ("slope", C.c_float), ("stats", C.c_void_p),
                ("flags", C.c_int64)]


class CdbNormDesc(C.Structure):
    _fields_ = [("norm", C.c_int32), ("act", C.c_int32), ("slope", C.c_float), ("eps", C.c_float),
                ("channels", C.c_int32), ("pad", C.c_int32), ("use_running", C.c_int32),
                ("update_running", C.c_int32), ("momentum", C.c_float), ("flags", C.c_int32),
                ("stats", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p),
                ("running_mean", C.c_void_p), ("running_var", C.c_void_p), ("conv_bias", C.c_void_p),
                ("count_scale", C.c_float), ("reserved_", C.c_int32)]


class CdbAdamEntry(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("numel", C.c_int64)]


class CdbAdamPackEntry(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("numel", C.c_int64), ("pack", C.c_void_p * 2), ("d0", C.c_int32), ("d1", C.c_int32), ("r", C.c_int32),
                ("s", C.c_int32), ("rows_are_dim0", C.c_int32 * 2), ("rowpack", C.c_int32 * 2)]


class CdbPackEntry(C.Structure):
    _fields_ = [("w4", C.c_void_p), ("out", C.c_void_p), ("d0", C.c_int32), ("d1", C.c_int32), ("r", C.c_int32),
                ("s", C.c_int32), ("rows_are_dim0", C.c_int32), ("rowpack", C.c_int32)]


_lib = None


def lib():
    """Returns the loaded library; raises if it has not been built (no silent fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libcdb200.so is not built (%s). Run `python -m cycle_depth_estimation_b200.build`; "
            "this package has no CPU or cuDNN fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    L.cdb_version.restype = C.c_int
    L.cdb_last_error.restype = C.c_char_p
    L.cdb_device_abort_flag.restype = C.c_int
    L.cdb_launch_count.restype = C.c_longlong
    L.cdb_conv2d_wgrad_workspace.restype = C.c_size_t
    L.cdb_conv2d_toeplitz_wgrad_workspace.restype = C.c_size_t
    L.cdb_depth_metrics_workspace.restype = C.c_size_t
    L.cdb_validation_workspace.restype = C.c_size_t
    L.cdb_depth_labels_workspace.restype = C.c_size_t
    L.cdb_pil_resample_workspace.restype = C.c_size_t
    _lib = L
    return L


def check(rc):
    if rc != 0:
        msg = lib().cdb_last_error().decode("utf-8", "replace")
        if rc == -2:
            raise NotImplementedError("cdb200: " + msg)
        raise RuntimeError("cdb200 error %d: %s" % (rc, msg))


def exported_symbols_from_header():
    """Names of the functions include/cdb200.h declares (used by the CPU-side export test)."""
    import re
    hdr = os.path.join(_HERE, "..", "include", "cdb200.h")
    txt = open(hdr).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(cdb_[a-z0-9_]+)\s*\(", txt)))
