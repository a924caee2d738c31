"""ctypes binding of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs; never from the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
ORC_BASIS_MAX = 25


class orc_tree(C.Structure):
    _fields_ = [("child", C.c_void_p), ("data", C.c_void_p), ("extra", C.c_void_p), ("capacity", C.c_int64),
                ("N", C.c_int32), ("data_dim", C.c_int32), ("format", C.c_int32), ("basis_dim", C.c_int32),
                ("offset", C.c_float * 3), ("scale", C.c_float * 3),
                ("ndc_width", C.c_float), ("ndc_height", C.c_float), ("ndc_focal", C.c_float)]


class orc_camera(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("fx", C.c_float), ("fy", C.c_float),
                ("c2w", C.c_float * 12)]


class orc_options(C.Structure):
    _fields_ = [("step_size", C.c_float), ("sigma_thresh", C.c_float), ("stop_thresh", C.c_float),
                ("background_brightness", C.c_float), ("render_bbox", C.c_float * 6),
                ("basis_minmax", C.c_int32 * 2), ("rot_dirs", C.c_float * 3), ("render_depth", C.c_int32)]


class orc_counters(C.Structure):
    _fields_ = [("samples", C.c_uint64), ("child_loads", C.c_uint64), ("shaded", C.c_uint64),
                ("rays_hit", C.c_uint64)]


_lib = None


def build() -> None:
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-pthread",
                           "-o", LIB_PATH, os.path.join(_HERE, "march_oracle.c"), "-lm"])


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        h = C.CDLL(LIB_PATH)
        h.orc_render.restype = C.c_int
        h.orc_render.argtypes = [C.POINTER(orc_tree), C.POINTER(orc_camera), C.POINTER(orc_options),
                                 C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.POINTER(orc_counters), C.c_int]
        h.orc_default_options.restype = None
        h.orc_default_options.argtypes = [C.POINTER(orc_options)]
        _lib = h
    return _lib


_FMT = {"RGBA": 0, "SH": 1, "SG": 2, "ASG": 3}


def parse_format(s: str):
    """DataFormat::parse, src/n3tree.cpp:55-78."""
    idx = next((i for i, ch in enumerate(s) if not ch.isalpha()), -1)
    if idx < 0:
        return 0, -1
    return _FMT.get(s[:idx], 0), int(s[idx:])


class OracleTree:
    """Keeps the numpy arrays alive and exposes the C struct."""

    def __init__(self, child, data, offset, scale, data_dim, data_format: str, extra=None,
                 ndc=None):
        self.child = np.ascontiguousarray(child, np.int32)
        self.data = np.ascontiguousarray(np.asarray(data).view(np.uint16))
        self.extra = None if extra is None else np.ascontiguousarray(extra, np.float32)
        fmt, bd = parse_format(data_format)
        t = orc_tree()
        t.child = self.child.ctypes.data
        t.data = self.data.ctypes.data
        t.extra = self.extra.ctypes.data if self.extra is not None else None
        t.capacity = self.child.shape[0]
        t.N = self.child.shape[1] if self.child.ndim > 1 else 0
        t.data_dim, t.format, t.basis_dim = int(data_dim), fmt, bd
        for i in range(3):
            t.offset[i] = float(offset[i])
            t.scale[i] = float(scale[i])
        if ndc is None:
            t.ndc_width = -1.0
        else:
            t.ndc_width, t.ndc_height, t.ndc_focal = ndc
        self.c = t

    @classmethod
    def from_synth(cls, s, ndc=None):
        return cls(s.child, s.data, s.offset, s.invradius3, s.data_dim, s.data_format, s.extra, ndc)


def make_camera(width, height, fx, fy, c2w12) -> orc_camera:
    c = orc_camera()
    c.width, c.height, c.fx, c.fy = int(width), int(height), float(fx), float(fy)
    for i in range(12):
        c.c2w[i] = float(c2w12[i])
    return c


def make_options(**kw) -> orc_options:
    o = orc_options()
    lib().orc_default_options(C.byref(o))
    for k, v in kw.items():
        if k in ("render_bbox", "rot_dirs", "basis_minmax"):
            arr = getattr(o, k)
            for i, x in enumerate(v):
                arr[i] = x
        else:
            setattr(o, k, v)
    return o


def render(tree: OracleTree, cam: orc_camera, opt: orc_options, tile=None, rgba_in=None, depth_in=None,
           want_float=True, want_u8=True, want_counters=True, nthreads=None):
    """Returns (rgba_f32 [h,w,4] | None, rgba8 [h,w,4] | None, counters dict | None)."""
    if tile is None:
        tile = (0, 0, cam.width, cam.height)
    x0, y0, w, h = tile
    f32 = np.zeros((h, w, 4), np.float32) if want_float else None
    u8 = np.zeros((h, w, 4), np.uint8) if want_u8 else None
    cnt = orc_counters() if want_counters else None
    if nthreads is None:
        nthreads = os.cpu_count() or 1
    rin = None if rgba_in is None else np.ascontiguousarray(rgba_in, np.uint8)
    din = None if depth_in is None else np.ascontiguousarray(depth_in, np.float32)
    rc = lib().orc_render(C.byref(tree.c), C.byref(cam), C.byref(opt), x0, y0, w, h,
                          rin.ctypes.data if rin is not None else None,
                          din.ctypes.data if din is not None else None,
                          f32.ctypes.data if f32 is not None else None,
                          u8.ctypes.data if u8 is not None else None,
                          C.byref(cnt) if cnt is not None else None, int(nthreads))
    if rc != 0:
        raise RuntimeError(f"orc_render failed: {rc}")
    cd = None if cnt is None else dict(samples=cnt.samples, child_loads=cnt.child_loads, shaded=cnt.shaded,
                                       rays_hit=cnt.rays_hit)
    return f32, u8, cd


def algorithmic_bytes(cnt: dict, basis_dim: int, n_pixels: int) -> int:
    """A_frame = 4*sum(d_s) + 2*S + 6*basis_dim*S_shaded + 4*W*H  (BASELINE.md 3)."""
    per_shaded = 6 * basis_dim if basis_dim > 0 else 6
    return 4 * cnt["child_loads"] + 2 * cnt["samples"] + per_shaded * cnt["shaded"] + 4 * n_pixels
