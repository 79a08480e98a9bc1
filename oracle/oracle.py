"""ORACLE — TEST INFRASTRUCTURE ONLY. ctypes wrapper around oracle/_build/liboracle.so (the CPU restatement of the
reference's tessellate -> stencil-then-cover path). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this module. PARITY UNPINNED: the reference has no tests or golden vectors and
cannot be compiled here (Rust); see DESIGN.md.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence

import numpy as np

from contrast_renderer_b200 import _abi
from contrast_renderer_b200.path import PathSoA, DynamicStrokeOptions, dynamic_stroke_options_array

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")


def build(force: bool = False) -> str:
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE] + (["-B"] if force else []), stdout=subprocess.DEVNULL)
    return _LIB_PATH


class RenderCommandC(C.Structure):
    _fields_ = [("shape", C.c_uint32), ("instance_begin", C.c_uint32), ("instance_end", C.c_uint32), ("operation", C.c_uint32),
                ("clip_depth", C.c_uint32), ("save_layer", C.c_uint32), ("restore_layer", C.c_uint32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.oracle_shape_from_paths.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(_abi.PathSoAC), C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p)]
        _lib.oracle_shape_destroy.argtypes = [C.c_void_p]
        _lib.oracle_shape_destroy.restype = None
        _lib.oracle_shape_get_layout.argtypes = [C.c_void_p, C.POINTER(_abi.ShapeLayoutC)]
        for f in ("oracle_shape_vertex_buffer", "oracle_shape_index_buffer", "oracle_shape_stroke_buffer"):
            getattr(_lib, f).argtypes = [C.c_void_p]
            getattr(_lib, f).restype = C.c_void_p
        _lib.oracle_shape_set_dynamic_stroke_options.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        _lib.oracle_tessellate_batch.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(_abi.PathSoAC), C.c_void_p, C.c_uint32, C.c_int,
                                                 C.POINTER(C.c_uint64), C.POINTER(C.c_double)]
        _lib.oracle_render.argtypes = [C.POINTER(_abi.ConfigC), C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_size_t,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_uint64), C.c_void_p]
        _lib.oracle_atan2.argtypes = [C.c_float, C.c_float]
        _lib.oracle_atan2.restype = C.c_float
        _lib.oracle_acos.argtypes = [C.c_float]
        _lib.oracle_acos.restype = C.c_float
        _lib.oracle_pow.argtypes = [C.c_float, C.c_float]
        _lib.oracle_pow.restype = C.c_float
        _lib.oracle_wgsl_mod.argtypes = [C.c_float, C.c_float]
        _lib.oracle_wgsl_mod.restype = C.c_float
        _lib.oracle_sincos.argtypes = [C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        _lib.oracle_sincos.restype = None
        _lib.oracle_solve.argtypes = [C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int)]
        _lib.oracle_uniform_tangent_angle.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_int]
        _lib.oracle_curve_eval.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]
        _lib.oracle_curve_eval.restype = None
        _lib.oracle_andrew.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        _lib.oracle_ga_triple.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.oracle_ga_triple.restype = C.c_float
        _lib.oracle_ga_join.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.oracle_ga_join.restype = None
    return _lib


class OracleError(Exception):
    def __init__(self, status: int):
        super().__init__(f"oracle status {status}")
        self.status = status


VERTEX_DTYPES = [
    np.dtype([("pos", "<f4", 2), ("tex", "<f4", 2), ("flags", "<u4")]),   # line    Vertex2f1i 20 B
    np.dtype([("pos", "<f4", 2), ("tex", "<f4", 3), ("flags", "<u4")]),   # joint   Vertex3f1i 24 B
    np.dtype([("pos", "<f4", 2)]),                                        # solid   Vertex0     8 B
    np.dtype([("pos", "<f4", 2), ("w", "<f4", 2)]),                       # iq      Vertex2f   16 B
    np.dtype([("pos", "<f4", 2), ("w", "<f4", 3)]),                       # ic      Vertex3f   20 B
    np.dtype([("pos", "<f4", 2), ("w", "<f4", 3)]),                       # rq      Vertex3f   20 B
    np.dtype([("pos", "<f4", 2), ("w", "<f4", 4)]),                       # rc      Vertex4f   24 B
    np.dtype([("pos", "<f4", 2)]),                                        # hull    Vertex0     8 B
]
CATEGORY_NAMES = ["line", "joint", "solid", "integral_quadratic", "integral_cubic", "rational_quadratic", "rational_cubic", "hull"]


def split_vertex_buffer(buf: np.ndarray, vertex_offsets) -> List[np.ndarray]:
    out, begin = [], 0
    for k in range(8):
        end = int(vertex_offsets[k])
        out.append(np.frombuffer(buf[begin:end].tobytes(), dtype=VERTEX_DTYPES[k]))
        begin = end
    return out


def split_index_buffer(buf: np.ndarray, index_offsets) -> List[np.ndarray]:
    out, begin = [], 0
    for k in range(3):
        end = int(index_offsets[k])
        out.append(np.frombuffer(buf[begin:end].tobytes(), dtype=np.uint16))
        begin = end
    return out


class OracleShape:
    def __init__(self, handle):
        self._h = handle
        layout = _abi.ShapeLayoutC()
        lib().oracle_shape_get_layout(self._h, C.byref(layout))
        self.vertex_offsets = [int(v) for v in layout.vertex_offsets]
        self.index_offsets = [int(v) for v in layout.index_offsets]
        self.n_groups = int(layout.dynamic_stroke_options_count)
        self.proto_hull_points = int(layout.proto_hull_points)

    def _bytes(self, fn, size) -> np.ndarray:
        if size == 0:
            return np.zeros(0, np.uint8)
        ptr = getattr(lib(), fn)(self._h)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(size,)).copy()

    @property
    def vertex_buffer(self) -> np.ndarray:
        return self._bytes("oracle_shape_vertex_buffer", self.vertex_offsets[7])

    @property
    def index_buffer(self) -> np.ndarray:
        return self._bytes("oracle_shape_index_buffer", self.index_offsets[2])

    @property
    def stroke_buffer(self) -> np.ndarray:
        return self._bytes("oracle_shape_stroke_buffer", 48 * self.n_groups)

    def vertices(self) -> List[np.ndarray]:
        return split_vertex_buffer(self.vertex_buffer, self.vertex_offsets)

    def indices(self) -> List[np.ndarray]:
        return split_index_buffer(self.index_buffer, self.index_offsets)

    def set_dynamic_stroke_options(self, index: int, options: DynamicStrokeOptions) -> None:
        c = options.to_c()
        st = lib().oracle_shape_set_dynamic_stroke_options(self._h, index, C.byref(c))
        if st:
            raise OracleError(st)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_shape_destroy(self._h)
            self._h = None


def shape_from_paths(dynamic_stroke_options: Sequence[DynamicStrokeOptions], soa: PathSoA, path_begin: int = 0,
                     path_end: Optional[int] = None) -> OracleShape:
    groups = dynamic_stroke_options_array(dynamic_stroke_options)
    c = soa.as_c()
    h = C.c_void_p()
    st = lib().oracle_shape_from_paths(groups, len(dynamic_stroke_options), C.byref(c), path_begin,
                                       soa.n_paths if path_end is None else path_end, C.byref(h))
    if st:
        raise OracleError(st)
    return OracleShape(h)


def tessellate_batch(dynamic_stroke_options, soa: PathSoA, shape_path_begin: np.ndarray, threads: int = 1):
    """Returns (output bytes, seconds) of tessellating every shape on `threads` host threads."""
    groups = dynamic_stroke_options_array(dynamic_stroke_options)
    c = soa.as_c()
    begin = np.ascontiguousarray(shape_path_begin, dtype=np.uint32)
    nbytes, secs = C.c_uint64(), C.c_double()
    st = lib().oracle_tessellate_batch(groups, len(dynamic_stroke_options), C.byref(c), begin.ctypes.data, len(begin) - 1, threads,
                                       C.byref(nbytes), C.byref(secs))
    if st:
        raise OracleError(st)
    return int(nbytes.value), float(secs.value)


def render(config: _abi.ConfigC, width: int, height: int, shapes: Sequence[OracleShape], commands: np.ndarray, transforms: np.ndarray,
           colors: Optional[np.ndarray], color: Optional[np.ndarray] = None, stencil: Optional[np.ndarray] = None,
           alpha_layers: Optional[np.ndarray] = None, threads: int = 1, depth: Optional[np.ndarray] = None):
    """commands: structured array / list of (shape, inst_begin, inst_end, op, clip_depth, save_layer, restore_layer).
    Returns (color[h,w,s,4] f32, stencil[h,w,s] u8, alpha_layers[l,h,w,s] f32, covered_samples). `depth`: the f32 depth
    attachment [h,w,s], updated in place (pass np.ones(...) for a pass that clears depth to 1.0); None = no depth attachment."""
    s = int(config.msaa_sample_count)
    if color is None:
        color = np.zeros((height, width, s, 4), np.float32)
    if stencil is None:
        stencil = np.zeros((height, width, s), np.uint8)
    if alpha_layers is None:
        alpha_layers = np.zeros((max(1, int(config.alpha_layer_count)), height, width, s), np.float32)
    cmds = (RenderCommandC * max(1, len(commands)))()
    for i, cmd in enumerate(commands):
        cmds[i] = RenderCommandC(*[int(v) for v in cmd])
    handles = (C.c_void_p * max(1, len(shapes)))(*[sh._h for sh in shapes])
    transforms = np.ascontiguousarray(transforms, dtype=np.float32)
    colors_arr = np.ascontiguousarray(colors, dtype=np.float32) if colors is not None else None
    covered = C.c_uint64()
    st = lib().oracle_render(C.byref(config), width, height, handles, len(shapes), cmds, len(commands), transforms.ctypes.data,
                             colors_arr.ctypes.data if colors_arr is not None else None, color.ctypes.data, stencil.ctypes.data,
                             alpha_layers.ctypes.data, threads, C.byref(covered), depth.ctypes.data if depth is not None else None)
    if st:
        raise OracleError(st)
    return color, stencil, alpha_layers, int(covered.value)


def max_threads() -> int:
    return int(lib().oracle_max_threads())


# ---- unit-level probes ------------------------------------------------------------------------------------------
def solve(coefficients: Sequence[float], margin: float = 1e-4):
    c = np.asarray(coefficients, dtype=np.float32)
    roots = np.zeros(12, np.float32)
    disc, real_root = C.c_float(), C.c_int()
    n = lib().oracle_solve(len(c) - 1, c.ctypes.data, margin, roots.ctypes.data, C.byref(disc), C.byref(real_root))
    return float(disc.value), roots[: 3 * n].reshape(n, 3), int(real_root.value)


def uniform_tangent_angle(kind: int, control_points, weights, angle_step: float) -> np.ndarray:
    cp = np.ascontiguousarray(control_points, dtype=np.float32).reshape(-1)
    w = np.ascontiguousarray(weights, dtype=np.float32) if weights is not None else None
    out = np.zeros(4096, np.float32)
    n = lib().oracle_uniform_tangent_angle(kind, cp.ctypes.data, w.ctypes.data if w is not None else None, angle_step, out.ctypes.data, len(out))
    return out[:n].copy()


def curve_eval(kind: int, control_points, weights, t: float):
    cp = np.ascontiguousarray(control_points, dtype=np.float32).reshape(-1)
    w = np.ascontiguousarray(weights, dtype=np.float32) if weights is not None else None
    xy, normal = np.zeros(2, np.float32), np.zeros(2, np.float32)
    lib().oracle_curve_eval(kind, cp.ctypes.data, w.ctypes.data if w is not None else None, t, xy.ctypes.data, normal.ctypes.data)
    return xy, normal


def ga_triple(a, b, c) -> float:
    """(A v B) v C of three unweighted ppga2d points, as the oracle evaluates it."""
    a, b, c = (np.ascontiguousarray(v, dtype=np.float32) for v in (a, b, c))
    return float(lib().oracle_ga_triple(a.ctypes.data, b.ctypes.data, c.ctypes.data))


def ga_join(p, q) -> np.ndarray:
    """P v Q of two unweighted ppga2d points: the Plane (g0, g1, g2)."""
    p, q = (np.ascontiguousarray(v, dtype=np.float32) for v in (p, q))
    out = np.zeros(3, np.float32)
    lib().oracle_ga_join(p.ctypes.data, q.ctypes.data, out.ctypes.data)
    return out


def andrew(points) -> np.ndarray:
    pts = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 2)
    out = np.zeros((2 * len(pts) + 2, 2), np.float32)
    n = lib().oracle_andrew(pts.ctypes.data, len(pts), out.ctypes.data)
    return out[:n].copy()
