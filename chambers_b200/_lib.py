"""ctypes binding of ``libchambers_aug.so`` (``include/chambers_aug.h``).

There is NO fallback: if the shared library is missing or no B200 is visible the
calls raise.  Nothing in this package imports ``oracle/``.
"""

import ctypes
import os
import threading

_PKG = os.path.dirname(os.path.abspath(__file__))
# CHB_LIB selects another build of the same library (tools/variants.py: timeline / experiment builds).
LIB_PATH = os.environ.get("CHB_LIB") or os.path.join(_PKG, "libchambers_aug.so")

CHB_MAX_SUBOPS = 4
CHB_MAX_CHAIN = 8
CHB_MAX_TABLE_OPS = 64
CHB_SCHED_FIELDS = 5

OP_KINDS = [
    "AutoContrast", "Equalize", "Invert", "Brightness", "Contrast", "Color", "Sharpness",
    "ShearX", "ShearY", "TranslateX", "TranslateY", "Posterize", "Solarize", "SolarizeAdd",
    "CutOut", "Rotate",
]
OP_KIND = {n: i for i, n in enumerate(OP_KINDS)}
INTERPOLATIONS = {"nearest": 0, "bilinear": 1}
FILL_MODES = {"constant": 0, "reflect": 1, "wrap": 2, "nearest": 3}


class ChbOp(ctypes.Structure):
    _fields_ = [
        ("kind", ctypes.c_int32),
        ("interpolation", ctypes.c_int32),
        ("fill_mode", ctypes.c_int32),
        ("ivalue", ctypes.c_int32 * 2),
        ("fill_value", ctypes.c_float),
        ("probability", ctypes.c_double),
        ("value", ctypes.c_double),
    ]


class ChbTransform(ctypes.Structure):
    _fields_ = [
        ("n_ops", ctypes.c_int32),
        ("_pad", ctypes.c_int32),
        ("ops", ChbOp * CHB_MAX_SUBOPS),
    ]


class ChbPolicy(ctypes.Structure):
    _fields_ = [
        ("n_table", ctypes.c_int32),
        ("n_draws", ctypes.c_int32),
        ("elementwise", ctypes.c_int32),
        ("_pad", ctypes.c_int32),
        ("table", ctypes.POINTER(ChbTransform)),
    ]


class ChambersAugError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("libchambers_aug error %d: %s" % (code, message))
        self.code = code


_lib = None
_lib_lock = threading.Lock()

_vp = ctypes.c_void_p
_i = ctypes.c_int
_i64 = ctypes.c_int64
_u64 = ctypes.c_uint64
_u32 = ctypes.c_uint32

# name -> (restype, argtypes); also the list the CPU test checks against include/chambers_aug.h.
SIGNATURES = {
    "chb_version": (_i, []),
    "chb_init": (_i, [_i, ctypes.POINTER(_vp)]),
    "chb_destroy": (None, [_vp]),
    "chb_last_error": (ctypes.c_char_p, [_vp]),
    "chb_kernel_launches": (_i64, [_vp]),
    "chb_randaugment_table": (_i, [ctypes.c_double, ctypes.POINTER(ChbTransform)]),
    "chb_autoaugment_table": (_i, [ctypes.POINTER(ChbTransform)]),
    "chb_policy_apply": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, ctypes.POINTER(ChbPolicy), _i64, _i64,
                              _u64, _u32, _vp, _vp, _vp]),
    "chb_policy_apply_normalized": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, ctypes.POINTER(ChbPolicy), _i, _i64, _i64,
                                         _u64, _u32, _vp, _vp, _vp]),
    "chb_randaugment": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, ctypes.c_double, _i, _i64, _i64, _u64,
                             _u32, _vp, _vp, _vp]),
    "chb_autoaugment": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i64, _i64, _u64, _u32, _vp, _vp, _vp]),
    "chb_apply_op": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, ctypes.POINTER(ChbOp), _i64, _i64, _u64, _u32,
                          _vp, _vp, _vp]),
    "chb_policy_apply_host": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, ctypes.POINTER(ChbPolicy), _i64, _i64,
                                   _u64, _u32, _vp, _vp]),
    "chb_set_debug": (_i, [_vp, _i]),
    "chb_imagenet_normalize": (_i, [_vp, _vp, _i, _vp, _i64, _i, _i, _vp]),
    "chb_resize_min_max_shape": (_i, [_i, _i, _i, _i, ctypes.POINTER(_i), ctypes.POINTER(_i)]),
    "chb_resize": (_i, [_vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "chb_set_engine": (_i, [_vp, _i]),
    "chb_last_engine": (_i, [_vp]),
    "chb_debug_timeline": (_i, [_vp, _vp, _i]),
    "chb_tile_plan": (_i, [_i, _i] + [ctypes.POINTER(_i)] * 4),
}


def load():
    """dlopen the library (once) and declare every prototype.  Raises if it is not built."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "chambers_b200: %s is missing. Build it with `python -m chambers_b200.build` "
                "(needs nvcc); there is no CPU fallback." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


_tls = threading.local()


def context(device):
    """The chb_ctx of (this thread, device); created on first use."""
    lib = load()
    table = getattr(_tls, "ctx", None)
    if table is None:
        table = _tls.ctx = {}
    ctx = table.get(device)
    if ctx is None:
        handle = _vp()
        rc = lib.chb_init(int(device), ctypes.byref(handle))
        if rc != 0:
            raise ChambersAugError(rc, (lib.chb_last_error(None) or b"").decode())
        ctx = table[device] = handle
    return ctx


def check(ctx, rc):
    if rc != 0:
        raise ChambersAugError(rc, (load().chb_last_error(ctx) or b"").decode())


def kernel_launches(device):
    return int(load().chb_kernel_launches(context(device)))


def tile_plan(H, W):
    """(tiles_x, tiles_y, tile_width, tile_height) of an H x W image (pure function of the shape)."""
    out = [_i() for _ in range(4)]
    load().chb_tile_plan(int(H), int(W), *[ctypes.byref(o) for o in out])
    return tuple(o.value for o in out)


def set_debug(device, force_generic):
    """Route every tile through the scalar executor (tests cross-check it against the fast ones)."""
    check(context(device), load().chb_set_debug(context(device), 1 if force_generic else 0))


NORM_MODES = {"caffe": 0, "tf": 1, "torch": 2}
ENGINES = {"auto": 0, "tiles": 1, "resident": 2}


def resize_min_max_shape(H, W, min_side, max_side):
    """(new_height, new_width) of ResizingMinMax (chb_resize_min_max_shape; host arithmetic, no GPU)."""
    oh, ow = _i(), _i()
    rc = load().chb_resize_min_max_shape(int(H), int(W), int(min_side or 0), int(max_side or 0),
                                         ctypes.byref(oh), ctypes.byref(ow))
    if rc != 0:
        raise ValueError("Must specify either 'min_side' or 'max_side'.")
    return oh.value, ow.value


def set_engine(device, engine):
    """Select the engine of (this thread, device): "auto" | "tiles" | "resident" (chb_set_engine)."""
    check(context(device), load().chb_set_engine(context(device), ENGINES[engine]))


def last_engine(device):
    """Engine the last device call of (this thread, device) ran on: "tiles" | "resident"."""
    v = int(load().chb_last_engine(context(device)))
    return {1: "tiles", 2: "resident"}.get(v, "none")
