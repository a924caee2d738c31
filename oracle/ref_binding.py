"""ctypes binding of oracle/_ref/libvolrend_ref.so -- the UNMODIFIED reference CUDA renderer
(built by oracle/Makefile.ref from /root/reference).  TEST INFRASTRUCTURE ONLY; needs a GPU."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libvolrend_ref.so")
HEADLESS_PATH = os.path.join(_HERE, "_ref", "volrend_headless_ref")


class ref_options(C.Structure):
    _fields_ = [("step_size", C.c_float), ("sigma_thresh", C.c_float), ("stop_thresh", C.c_float),
                ("background_brightness", C.c_float), ("render_bbox", C.c_float * 6),
                ("basis_minmax", C.c_int32 * 2), ("rot_dirs", C.c_float * 3), ("render_depth", C.c_int32)]


def available() -> bool:
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        h = C.CDLL(LIB_PATH)
        h.ref_tree_open.restype = C.c_void_p
        h.ref_tree_open.argtypes = [C.c_char_p]
        h.ref_tree_close.restype = None
        h.ref_tree_close.argtypes = [C.c_void_p]
        h.ref_tree_info.restype = C.c_int
        h.ref_tree_info.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 4 + [C.POINTER(C.c_longlong), C.POINTER(C.c_int)]
        h.ref_render_u8.restype = C.c_int
        h.ref_render_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p,
                                    C.POINTER(ref_options), C.c_void_p, C.c_void_p, C.c_void_p]
        h.ref_render_f32.restype = C.c_int
        h.ref_render_f32.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p,
                                     C.POINTER(ref_options), C.c_void_p]
        h.ref_render_f32_composite.restype = C.c_int
        h.ref_render_f32_composite.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p,
                                               C.POINTER(ref_options), C.c_void_p, C.c_void_p, C.c_void_p]
        h.ref_time_frames.restype = C.c_float
        h.ref_time_frames.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_int,
                                      C.POINTER(ref_options), C.c_int, C.c_void_p]
        _lib = h
    return _lib


def make_options(**kw) -> ref_options:
    o = ref_options()
    o.step_size, o.sigma_thresh, o.stop_thresh, o.background_brightness = 1e-4, 1e-2, 1e-2, 1.0
    for i, v in enumerate((0, 0, 0, 1, 1, 1)):
        o.render_bbox[i] = v
    o.basis_minmax[0], o.basis_minmax[1] = 0, 24
    for k, v in kw.items():
        if k in ("render_bbox", "rot_dirs", "basis_minmax"):
            arr = getattr(o, k)
            for i, x in enumerate(v):
                arr[i] = x
        else:
            setattr(o, k, v)
    return o


class RefTree:
    def __init__(self, npz_path: str):
        self.h = lib().ref_tree_open(npz_path.encode())
        if not self.h:
            raise RuntimeError(f"reference loader could not open {npz_path}")

    def info(self) -> dict:
        N, dd, bd, fmt, ndc = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
        cap = C.c_longlong()
        lib().ref_tree_info(self.h, N, dd, bd, fmt, cap, ndc)
        return dict(N=N.value, data_dim=dd.value, basis_dim=bd.value, format=fmt.value, capacity=cap.value,
                    use_ndc=ndc.value)

    def close(self):
        if self.h:
            lib().ref_tree_close(self.h)
            self.h = None

    def render_u8(self, w, h, fx, fy, c2w12, opt, rgba_in=None, depth_in=None) -> np.ndarray:
        out = np.zeros((h, w, 4), np.uint8)
        c = np.ascontiguousarray(c2w12, np.float32)
        rin = None if rgba_in is None else np.ascontiguousarray(rgba_in, np.uint8)
        din = None if depth_in is None else np.ascontiguousarray(depth_in, np.float32)
        rc = lib().ref_render_u8(self.h, w, h, fx, fy, c.ctypes.data, C.byref(opt),
                                 rin.ctypes.data if rin is not None else None,
                                 din.ctypes.data if din is not None else None, out.ctypes.data)
        if rc:
            raise RuntimeError(f"ref_render_u8 failed {rc}")
        return out

    def render_f32(self, w, h, fx, fy, c2w12, opt, rgba_in=None, depth_in=None) -> np.ndarray:
        out = np.zeros((h, w, 4), np.float32)
        c = np.ascontiguousarray(c2w12, np.float32)
        if rgba_in is not None:
            rin, din = np.ascontiguousarray(rgba_in, np.uint8), np.ascontiguousarray(depth_in, np.float32)
            rc = lib().ref_render_f32_composite(self.h, w, h, fx, fy, c.ctypes.data, C.byref(opt), rin.ctypes.data,
                                                din.ctypes.data, out.ctypes.data)
        else:
            rc = lib().ref_render_f32(self.h, w, h, fx, fy, c.ctypes.data, C.byref(opt), out.ctypes.data)
        if rc:
            raise RuntimeError(f"ref_render_f32 failed {rc}")
        return out

    def time_frames(self, w, h, fx, fy, c2w12s, opt, with_d2h=False, host_out=None) -> float:
        c = np.ascontiguousarray(c2w12s, np.float32).reshape(-1, 12)
        ptr = None
        if with_d2h:
            ptr = host_out.data_ptr() if hasattr(host_out, "data_ptr") else host_out.ctypes.data
        ms = lib().ref_time_frames(self.h, w, h, fx, fy, c.ctypes.data, c.shape[0], C.byref(opt), int(with_d2h), ptr)
        if ms < 0:
            raise RuntimeError("ref_time_frames failed")
        return float(ms)
