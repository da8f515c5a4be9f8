"""TEST INFRASTRUCTURE ONLY -- builds and loads the plain-C part of the oracle
(oracle/c/*.c -> oracle/_build/liboracle.so) with gcc.  ``__graft_entry__.build()``
calls :func:`build` so the prebuilt library travels to the GPU box."""
import ctypes
import glob
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_LIB = None


def _stale():
    if not os.path.exists(_SO):
        return True
    t = os.path.getmtime(_SO)
    return any(os.path.getmtime(f) > t for f in glob.glob(os.path.join(_HERE, "c", "*.c")))


def build(force=False):
    if not force and not _stale():
        return _SO
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(_HERE, "c", "*.c")))
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", _SO] + srcs + ["-lm"]
    subprocess.check_call(cmd)
    return _SO


def load():
    global _LIB
    if _LIB is None:
        build()
        _LIB = ctypes.CDLL(_SO)
        vp, i, l, f = ctypes.c_void_p, ctypes.c_int, ctypes.c_long, ctypes.c_float
        _LIB.oracle_single_iou_rotated.restype = f
        _LIB.oracle_single_iou_rotated.argtypes = [vp, vp, i]
        _LIB.oracle_box_iou_rotated.argtypes = [vp, vp, vp, l, l, i, i]
        _LIB.oracle_nms_rotated.argtypes = [vp, vp, vp, l, f]
        _LIB.oracle_roi_align.argtypes = [vp, vp, vp, i, i, i, i, l, i, f, i, i]
        _LIB.oracle_roi_align_rotated.argtypes = [vp, vp, vp, i, i, i, i, l, i, f, i, i, i]
    return _LIB
