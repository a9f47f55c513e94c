"""ctypes binding of libb200mvs.so (include/b200mvs.h).  There is no fallback:
if the library is missing or no B200 is visible, calls raise."""
import ctypes
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libb200mvs.so")
NUM_LEVELS = 5

c_float_p = ctypes.POINTER(ctypes.c_float)
c_u8_p = ctypes.POINTER(ctypes.c_uint8)


class Shape(ctypes.Structure):
    _fields_ = [("batch", ctypes.c_int32), ("views", ctypes.c_int32), ("rows", ctypes.c_int32),
                ("cols", ctypes.c_int32), ("num_idepth_samples", ctypes.c_int32),
                ("do_cost_volume_filter", ctypes.c_int32), ("do_refiners", ctypes.c_int32 * NUM_LEVELS)]


# name -> (restype, argtypes); every symbol include/b200mvs.h declares.
SYMBOLS = {
    "b200mvs_last_error": (ctypes.c_char_p, []),
    "b200mvs_version": (ctypes.c_char_p, []),
    "b200mvs_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_char_p),
                                      ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int64),
                                      ctypes.POINTER(ctypes.c_void_p)]),
    "b200mvs_destroy": (None, [ctypes.c_void_p]),
    "b200mvs_forward": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(Shape)] + [ctypes.POINTER(ctypes.c_void_p)] * 8
                        + [ctypes.c_void_p]),
    "b200mvs_forward_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(Shape)]
                             + [ctypes.POINTER(ctypes.c_void_p)] * 8
                             + [ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64)]),
    "b200mvs_set_option": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int]),
    "b200mvs_conv3x3_c32": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32,
                                           ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                           ctypes.c_void_p, ctypes.c_void_p]),
    "b200mvs_probe_select": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p]),
    "b200mvs_probe_read": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_double),
                                          ctypes.POINTER(ctypes.c_int64)]),
    "b200mvs_last_launch_count": (ctypes.c_int64, [ctypes.c_void_p]),
    "b200mvs_last_stage_profile": (ctypes.c_char_p, [ctypes.c_void_p]),
    "b200mvs_get_stage": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int64,
                                         ctypes.POINTER(ctypes.c_int64), ctypes.c_void_p]),
    "b200mvs_set_debug": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "b200mvs_homography_warp": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32,
                                               ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p,
                                               ctypes.c_void_p, ctypes.c_void_p]),
    "b200mvs_upsample_mask": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32,
                                             ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p,
                                             ctypes.c_void_p]),
    "b200mvs_reproject": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32,
                                         ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                         ctypes.c_int32] + [ctypes.c_void_p] * 7),
    "b200mvs_area_downsample": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                               ctypes.c_void_p, ctypes.c_void_p]),
    "b200mvs_prepare_cameras": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.c_int32,
                                               ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(ctypes.c_int32),
                                               ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                               ctypes.c_void_p]),
    "b200mvs_depth_metrics": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32,
                                             ctypes.c_float, ctypes.c_float, ctypes.c_int32, ctypes.c_int64,
                                             ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
}

_lib = None


def load():
    """Loads the CUDA library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m multi_view_stereonet_b200.build` "
                "(there is no CPU or PyTorch fallback for this path)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def check(rc, what):
    if rc != 0:
        msg = load().b200mvs_last_error().decode("utf-8", "replace")
        # -1 mirrors the reference's bare `assert`s (multi_view_stereonet.py:548-549)
        exc = AssertionError if rc == -1 else RuntimeError
        raise exc(f"{what} failed ({rc}): {msg}")


def ptr_array(ptrs):
    arr = (ctypes.c_void_p * len(ptrs))()
    for i, p in enumerate(ptrs):
        arr[i] = p
    return arr
