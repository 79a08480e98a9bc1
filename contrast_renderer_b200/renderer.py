"""Host-side mirror of the reference's renderer API (src/renderer.rs) on top of the C-ABI of libcontrast_b200.so.

`Renderer`, `Shape`, `RenderOperation`, `Configuration` and the `Error` variants keep the reference's names,
argument meaning and error behaviour (src/renderer.rs:145-160,177,267,360,380-405,432,892,932,941,979;
src/error.rs:5-16), so a test written against the Rust crate reads the same here. wgpu objects have no counterpart:
`RenderPass` stands for the `wgpu::RenderPass` the reference records into, and `Renderer` additionally owns the
colour / stencil attachments.

Everything computes on the GPU through the shared library; there is no CPU path. Importing this module without the
built library, or creating a `Renderer` without a CUDA device, raises.
"""
from __future__ import annotations

import ctypes as C
import enum
import os
import weakref
from typing import Optional, Sequence

import numpy as np

from . import _abi
from .path import DynamicStrokeOptions, Path, PathSoA, dynamic_stroke_options_array

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libcontrast_b200.so")
if os.environ.get("CONTRAST_B200_LIB"):   # experiments only: an alternative build of the same library
    _LIB_PATH = os.environ["CONTRAST_B200_LIB"]
_lib = None


def library_path() -> str:
    return _LIB_PATH


def lib():
    """The loaded libcontrast_b200.so. Raises if it has not been built (`python -m contrast_renderer_b200.build`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise ImportError(f"{_LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(nvcc, sm_100a). There is no CPU fallback.")
        l = C.CDLL(_LIB_PATH)
        vp, u32, sz = C.c_void_p, C.c_uint32, C.c_size_t
        sig = {
            "cr_renderer_create": [C.POINTER(_abi.ConfigC), C.POINTER(vp)],
            "cr_renderer_get_config": [vp, C.POINTER(_abi.ConfigC)],
            "cr_renderer_resize": [vp, u32, u32],
            "cr_renderer_set_stream": [vp, vp],
            "cr_renderer_synchronize": [vp],
            "cr_renderer_set_pipelining": [vp, u32],
            "cr_shape_from_paths": [vp, vp, sz, C.POINTER(_abi.PathSoAC), vp, C.POINTER(vp)],
            "cr_shape_batch_from_paths": [vp, vp, sz, C.POINTER(_abi.PathSoAC), vp, u32, vp, C.POINTER(vp)],
            "cr_shape_set_dynamic_stroke_options": [vp, sz, vp],
            "cr_shape_batch_set_dynamic_stroke_options": [vp, sz, vp],
            "cr_shape_get_layout": [vp, C.POINTER(_abi.ShapeLayoutC)],
            "cr_shape_read_vertex_buffer": [vp, vp, sz],
            "cr_shape_read_index_buffer": [vp, vp, sz],
            "cr_shape_read_stroke_buffer": [vp, vp, sz],
            "cr_pass_begin": [vp, u32, u32, C.POINTER(vp)],
            "cr_pass_begin_depth": [vp, u32, u32, u32, C.c_float, C.POINTER(vp)],
            "cr_renderer_read_depth": [vp, vp, sz],
            "cr_renderer_read_color_texels": [vp, vp, sz],
            "cr_pass_set_instances": [vp, vp, vp, u32, u32],
            "cr_pass_set_clip_depth": [vp, u32],
            "cr_pass_save_alpha_context": [vp, u32],
            "cr_pass_restore_alpha_context": [vp, u32],
            "cr_shape_render": [vp, vp, u32, u32, u32],
            "cr_pass_render_batch": [vp, vp, vp, sz],
            "cr_pass_render_script": [vp, vp, vp, sz],
            "cr_pass_submit": [vp],
            "cr_renderer_read_color": [vp, vp, sz],
            "cr_renderer_read_stencil": [vp, vp, sz],
            "cr_renderer_read_alpha_layer": [vp, u32, vp, sz],
            "cr_renderer_get_attachments": [vp, C.POINTER(vp), C.POINTER(vp)],
            "cr_renderer_get_stats": [vp, C.POINTER(_abi.StatsC)],
            "cr_renderer_get_settled_pass_stats": [vp, C.POINTER(_abi.StatsC)],
            "cr_renderer_read_color_texels_async": [vp, vp, C.c_size_t, C.POINTER(C.c_uint64)],
            "cr_renderer_wait_readback": [vp, C.c_uint64],
            "cr_renderer_enable_timing": [vp, u32],
            "cr_renderer_set_tile_sharding": [vp, u32, u32],
            "cr_renderer_set_order_sharding": [vp, u32, u32],
            "cr_renderer_export_exchange": [vp, vp],
            "cr_renderer_import_peer_exchange": [vp, u32, vp],
            "cr_renderer_export_attachments": [vp, vp, vp],
            "cr_renderer_import_peer_attachments": [vp, u32, vp, vp],
        }
        for name, args in sig.items():
            fn = getattr(l, name)
            fn.argtypes = args
            fn.restype = C.c_int
        for name in ("cr_renderer_destroy", "cr_shape_destroy", "cr_shape_batch_destroy", "cr_pass_abort"):
            getattr(l, name).argtypes = [vp]
            getattr(l, name).restype = None
        l.cr_shape_batch_size.argtypes = [vp]
        l.cr_shape_batch_size.restype = u32
        l.cr_shape_batch_get.argtypes = [vp, u32]
        l.cr_shape_batch_get.restype = vp
        l.cr_status_string.argtypes = [C.c_int]
        l.cr_status_string.restype = C.c_char_p
        l.cr_last_error_message.restype = C.c_char_p
        l.cr_abi_version.restype = u32
        _lib = l
    return _lib


class Error(Exception):
    """`enum Error` (src/error.rs:5-16) plus the library's own status codes (>= 100)."""

    def __init__(self, status: int):
        self.status = int(status)
        self.name = lib().cr_status_string(self.status).decode()
        super().__init__(f"{self.name}: {lib().cr_last_error_message().decode()}")


class NumberOfStencilBitsIsUnsupported(Error):
    pass


class ClipStackOverflow(Error):
    pass


class TooManyNestedOpacityGroups(Error):
    pass


class TooManyDashIntervals(Error):
    pass


class DynamicStrokeOptionsIndexOutOfBounds(Error):
    pass


_ERROR_CLASSES = {1: NumberOfStencilBitsIsUnsupported, 2: ClipStackOverflow, 3: TooManyNestedOpacityGroups, 4: TooManyDashIntervals,
                  5: DynamicStrokeOptionsIndexOutOfBounds}


def _check(status: int) -> None:
    if status != _abi.CR_OK:
        raise _ERROR_CLASSES.get(status, Error)(status)


class RenderOperation(enum.IntEnum):  # src/renderer.rs:145-160
    Stencil = 0
    Clip = 1
    UnClip = 2
    Color = 3
    SaveAlphaContext = 4
    ScaleAlphaContext = 5
    RestoreAlphaContext = 6


class Blending(enum.IntEnum):
    PremultipliedOver = 0
    Replace = 1


class CullMode(enum.IntEnum):
    Off = 0
    Front = 1
    Back = 2


class CompareFunction(enum.IntEnum):   # wgpu::CompareFunction (Configuration::depth_compare, src/renderer.rs:388)
    Never = 1
    Less = 2
    Equal = 3
    LessEqual = 4
    Greater = 5
    NotEqual = 6
    GreaterEqual = 7
    Always = 8


class ColorFormat(enum.IntEnum):       # Configuration::blending.format (src/renderer.rs:382)
    Rgba32Float = 0                    # parity mode: no quantisation between blend operations
    Rgba8Unorm = 1
    Bgra8Unorm = 2                     # the demo's surface format (examples/application_framework.rs:175)


class Configuration:
    """`struct Configuration` (src/renderer.rs:380-405) minus the wgpu-only fields."""

    def __init__(self, msaa_sample_count: int = 1, clip_nesting_counter_bits: int = 4, winding_counter_bits: int = 4, alpha_layer_count: int = 0,
                 blending: Blending = Blending.PremultipliedOver, cull_mode: CullMode = CullMode.Off, device: int = -1,
                 depth_compare: CompareFunction = CompareFunction.Always, depth_write_enabled: bool = False,
                 color_format: ColorFormat = ColorFormat.Rgba32Float):
        self.depth_compare = depth_compare
        self.depth_write_enabled = depth_write_enabled
        self.color_format = color_format
        self.msaa_sample_count = msaa_sample_count
        self.clip_nesting_counter_bits = clip_nesting_counter_bits
        self.winding_counter_bits = winding_counter_bits
        self.alpha_layer_count = alpha_layer_count
        self.blending = blending
        self.cull_mode = cull_mode
        self.device = device

    def to_c(self) -> _abi.ConfigC:
        return _abi.ConfigC(self.msaa_sample_count, self.clip_nesting_counter_bits, self.winding_counter_bits, self.alpha_layer_count,
                            int(self.blending), int(self.cull_mode), self.device, int(self.depth_compare), 1 if self.depth_write_enabled else 0,
                            int(self.color_format))

    @property
    def has_depth(self) -> bool:
        return self.depth_compare != CompareFunction.Always or bool(self.depth_write_enabled)


def _as_soa(paths) -> PathSoA:
    return paths if isinstance(paths, PathSoA) else PathSoA.from_paths(paths)


class Renderer:
    """`struct Renderer` (src/renderer.rs:408-985)."""

    def __init__(self, config: Optional[Configuration] = None):
        self.config = config or Configuration()
        self._h = C.c_void_p()
        c = self.config.to_c()
        self._children = weakref.WeakSet()  # shapes / batches / passes: they must be released before the renderer
        _check(lib().cr_renderer_create(C.byref(c), C.byref(self._h)))
        self.width = self.height = 0

    # Renderer::get_config
    def get_config(self) -> Configuration:
        c = _abi.ConfigC()
        _check(lib().cr_renderer_get_config(self._h, C.byref(c)))
        return Configuration(c.msaa_sample_count, c.clip_nesting_counter_bits, c.winding_counter_bits, c.alpha_layer_count, Blending(c.blending),
                             CullMode(c.cull_mode), c.device, CompareFunction(c.depth_compare or 8), bool(c.depth_write_enabled), ColorFormat(c.color_format))

    # Renderer::resize_internal_buffers
    def resize_internal_buffers(self, width: int, height: int) -> None:
        _check(lib().cr_renderer_resize(self._h, width, height))
        self.width, self.height = width, height

    def set_stream(self, cuda_stream: int) -> None:
        _check(lib().cr_renderer_set_stream(self._h, cuda_stream))

    def synchronize(self) -> None:
        _check(lib().cr_renderer_synchronize(self._h))

    def set_pipelining(self, enabled: bool = True) -> None:
        """Frame pipelining (`cr_renderer_set_pipelining`): rebuilding a batch for frame N + 1 overlaps rasterising frame N.
        Device-memory inputs must then be complete when `ShapeBatch(...)` / `Shape.from_paths` is called."""
        _check(lib().cr_renderer_set_pipelining(self._h, 1 if enabled else 0))

    def begin_render_pass(self, clear_color: bool = True, clear_stencil: bool = True, clear_depth: Optional[bool] = None,
                          depth_clear_value: float = 1.0) -> "RenderPass":
        """`clear_depth=None`: the depth aspect follows the stencil aspect (one depth-stencil attachment, cleared to 1.0)."""
        return RenderPass(self, clear_color, clear_stencil, clear_stencil if clear_depth is None else clear_depth, depth_clear_value)

    def enable_timing(self, enabled: bool = True) -> None:
        _check(lib().cr_renderer_enable_timing(self._h, 1 if enabled else 0))

    def stats(self) -> _abi.StatsC:
        s = _abi.StatsC()
        _check(lib().cr_renderer_get_stats(self._h, C.byref(s)))
        return s

    def settled_pass_stats(self) -> _abi.StatsC:
        """Counters of the last pass the renderer has settled (submit settles the pass before the one it submits); never waits."""
        s = _abi.StatsC()
        _check(lib().cr_renderer_get_settled_pass_stats(self._h, C.byref(s)))
        return s

    def attachments(self):
        color, stencil = C.c_void_p(), C.c_void_p()
        _check(lib().cr_renderer_get_attachments(self._h, C.byref(color), C.byref(stencil)))
        return color.value, stencil.value

    # ---- one target spanning several GPUs (include/contrast_b200.h, "tile sharding"); orchestration in sharding.py
    def set_tile_sharding(self, world: int, rank: int) -> None:
        _check(lib().cr_renderer_set_tile_sharding(self._h, world, rank))

    # ---- one target composed from draw-order slices ("order sharding"); orchestration in sharding.py
    def set_order_sharding(self, world: int, rank: int) -> None:
        _check(lib().cr_renderer_set_order_sharding(self._h, world, rank))

    def export_exchange(self) -> bytes:
        handle = (C.c_uint8 * _abi.CR_IPC_HANDLE_BYTES)()
        _check(lib().cr_renderer_export_exchange(self._h, handle))
        return bytes(handle)

    def import_peer_exchange(self, peer_rank: int, handle: bytes) -> None:
        assert len(handle) == _abi.CR_IPC_HANDLE_BYTES
        _check(lib().cr_renderer_import_peer_exchange(self._h, peer_rank, (C.c_uint8 * _abi.CR_IPC_HANDLE_BYTES).from_buffer_copy(handle)))

    def export_attachments(self) -> bytes:
        """The two 64-byte CUDA IPC handles (colour, stencil) of this renderer's attachments, concatenated."""
        color, stencil = (C.c_uint8 * _abi.CR_IPC_HANDLE_BYTES)(), (C.c_uint8 * _abi.CR_IPC_HANDLE_BYTES)()
        _check(lib().cr_renderer_export_attachments(self._h, color, stencil))
        return bytes(color) + bytes(stencil)

    def import_peer_attachments(self, peer_rank: int, handles: bytes) -> None:
        n = _abi.CR_IPC_HANDLE_BYTES
        assert len(handles) == 2 * n
        color, stencil = (C.c_uint8 * n).from_buffer_copy(handles[:n]), (C.c_uint8 * n).from_buffer_copy(handles[n:])
        _check(lib().cr_renderer_import_peer_attachments(self._h, peer_rank, color, stencil))

    def read_color(self) -> np.ndarray:
        out = np.empty((self.height, self.width, self.config.msaa_sample_count, 4), np.float32)
        _check(lib().cr_renderer_read_color(self._h, out.ctypes.data, out.nbytes))
        return out

    def read_stencil(self) -> np.ndarray:
        out = np.empty((self.height, self.width, self.config.msaa_sample_count), np.uint8)
        _check(lib().cr_renderer_read_stencil(self._h, out.ctypes.data, out.nbytes))
        return out

    def read_color_texels(self, dst_address: int, capacity_bytes: int) -> None:
        """The colour attachment as stored (16 B per sample, or one unorm8 texel), into caller memory (pinned for PCIe speed)."""
        _check(lib().cr_renderer_read_color_texels(self._h, dst_address, capacity_bytes))

    def read_color_texels_async(self, dst_address: int, capacity_bytes: int) -> int:
        """Starts the read-back of the frame of the pass submitted last (snapshot behind the pass, host copy on its own stream) and
        returns a ticket for `wait_readback`; the pipeline keeps running."""
        ticket = C.c_uint64(0)
        _check(lib().cr_renderer_read_color_texels_async(self._h, dst_address, capacity_bytes, C.byref(ticket)))
        return int(ticket.value)

    def wait_readback(self, ticket: int) -> None:
        _check(lib().cr_renderer_wait_readback(self._h, ticket))

    def read_depth(self) -> np.ndarray:
        out = np.empty((self.height, self.width, self.config.msaa_sample_count), np.float32)
        _check(lib().cr_renderer_read_depth(self._h, out.ctypes.data, out.nbytes))
        return out

    def read_alpha_layer(self, layer: int) -> np.ndarray:
        out = np.empty((self.height, self.width, self.config.msaa_sample_count), np.float32)
        _check(lib().cr_renderer_read_alpha_layer(self._h, layer, out.ctypes.data, out.nbytes))
        return out

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            for child in list(getattr(self, "_children", ())):
                child.close()
            lib().cr_renderer_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Shape:
    """`struct Shape` (src/renderer.rs:163-377)."""

    def __init__(self, handle, renderer: Renderer, owner=None):
        self._h = handle
        self._renderer = renderer
        self._owner = owner  # ShapeBatch that owns a borrowed view
        if owner is None:
            renderer._children.add(self)

    @staticmethod
    def from_paths(renderer: Renderer, dynamic_stroke_options: Sequence[DynamicStrokeOptions], paths, existing: Optional["Shape"] = None,
                   memory_space: int = _abi.CR_MEM_HOST, pointers=None) -> "Shape":
        """Shape::from_paths (src/renderer.rs:177). `paths`: a sequence of `Path` or a `PathSoA`. `existing` is consumed."""
        soa = _as_soa(paths)
        groups = dynamic_stroke_options_array(dynamic_stroke_options)
        c = soa.as_c(memory_space, pointers)
        h = C.c_void_p()
        old = None
        if existing is not None:
            old, existing._h = existing._h, None
        _check(lib().cr_shape_from_paths(renderer._h, groups, len(dynamic_stroke_options), C.byref(c), old, C.byref(h)))
        return Shape(h, renderer)

    # Shape::set_dynamic_stroke_options
    def set_dynamic_stroke_options(self, index: int, options: DynamicStrokeOptions) -> None:
        c = options.to_c()
        _check(lib().cr_shape_set_dynamic_stroke_options(self._h, index, C.byref(c)))

    # Shape::render
    def render(self, render_pass: "RenderPass", instance_indices: range, render_operation: RenderOperation) -> None:
        _check(lib().cr_shape_render(render_pass._h, self._h, instance_indices.start, instance_indices.stop, int(render_operation)))

    def layout(self) -> _abi.ShapeLayoutC:
        l = _abi.ShapeLayoutC()
        _check(lib().cr_shape_get_layout(self._h, C.byref(l)))
        return l

    def vertex_buffer(self) -> np.ndarray:
        n = int(self.layout().vertex_offsets[7])
        out = np.zeros(max(n, 1), np.uint8)
        _check(lib().cr_shape_read_vertex_buffer(self._h, out.ctypes.data, out.nbytes))
        return out[:n]

    def index_buffer(self) -> np.ndarray:
        n = int(self.layout().index_offsets[2])
        out = np.zeros(max(n, 2), np.uint8)
        _check(lib().cr_shape_read_index_buffer(self._h, out.ctypes.data, out.nbytes))
        return out[:n]

    def stroke_buffer(self) -> np.ndarray:
        n = 48 * int(self.layout().dynamic_stroke_options_count)
        out = np.zeros(max(n, 1), np.uint8)
        _check(lib().cr_shape_read_stroke_buffer(self._h, out.ctypes.data, out.nbytes))
        return out[:n]

    def close(self) -> None:
        if self._h is not None and self._owner is None:
            lib().cr_shape_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ShapeBatch:
    """Many Shapes tessellated by one launch sequence (`cr_shape_batch_from_paths`): shape i = paths
    [shape_path_begin[i], shape_path_begin[i+1]). Equivalent to calling `Shape::from_paths` once per slice."""

    def __init__(self, renderer: Renderer, dynamic_stroke_options: Sequence[DynamicStrokeOptions], paths, shape_path_begin,
                 existing: Optional["ShapeBatch"] = None, memory_space: int = _abi.CR_MEM_HOST, pointers=None):
        soa = _as_soa(paths)
        groups = dynamic_stroke_options_array(dynamic_stroke_options)
        c = soa.as_c(memory_space, pointers)
        begin = np.ascontiguousarray(shape_path_begin, dtype=np.uint32)
        self._renderer = renderer
        self._h = C.c_void_p()
        old = None
        if existing is not None:
            old, existing._h = existing._h, None
        _check(lib().cr_shape_batch_from_paths(renderer._h, groups, len(dynamic_stroke_options), C.byref(c), begin.ctypes.data, len(begin) - 1, old,
                                               C.byref(self._h)))
        renderer._children.add(self)

    def __len__(self) -> int:
        return int(lib().cr_shape_batch_size(self._h))

    def __getitem__(self, index: int) -> Shape:
        h = lib().cr_shape_batch_get(self._h, index)
        if not h:
            raise IndexError(index)
        return Shape(C.c_void_p(h), self._renderer, owner=self)

    def set_dynamic_stroke_options(self, index: int, options: DynamicStrokeOptions) -> None:
        c = options.to_c()
        _check(lib().cr_shape_batch_set_dynamic_stroke_options(self._h, index, C.byref(c)))

    def close(self) -> None:
        if self._h is not None and self._h.value:
            lib().cr_shape_batch_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RenderPass:
    """The `wgpu::RenderPass` of src/renderer.rs:267-355 plus the per-pass renderer state (`set_clip_depth`,
    `save_alpha_context`, `restore_alpha_context`). Records; `submit()` runs everything on the renderer's stream."""

    def __init__(self, renderer: Renderer, clear_color: bool = True, clear_stencil: bool = True, clear_depth: Optional[bool] = None,
                 depth_clear_value: float = 1.0):
        self._renderer = renderer
        self._h = C.c_void_p()
        self._keep = []  # host arrays referenced by the pass until submit
        clear_depth = clear_stencil if clear_depth is None else clear_depth
        _check(lib().cr_pass_begin_depth(renderer._h, 1 if clear_color else 0, 1 if clear_stencil else 0, 1 if clear_depth else 0,
                                         float(depth_clear_value), C.byref(self._h)))
        renderer._children.add(self)

    def set_instances(self, transforms, colors=None, count: Optional[int] = None, memory_space: int = _abi.CR_MEM_HOST) -> None:
        """Vertex buffer slot 0 (instance mat4, 16 floats each = four column vectors) and the colour slot (rgba).
        Arrays (host memory) or raw addresses with `count` (host or device memory, per `memory_space`)."""
        if isinstance(transforms, (int, np.integer)):
            _check(lib().cr_pass_set_instances(self._h, int(transforms), int(colors) if colors else None, int(count), memory_space))
            return
        t = np.ascontiguousarray(transforms, dtype=np.float32).reshape(-1, 16)
        col = np.ascontiguousarray(colors, dtype=np.float32).reshape(-1, 4) if colors is not None else None
        self._keep += [t, col]
        _check(lib().cr_pass_set_instances(self._h, t.ctypes.data, col.ctypes.data if col is not None else None, len(t), _abi.CR_MEM_HOST))

    def set_clip_depth(self, clip_depth: int) -> None:  # Renderer::set_clip_depth
        _check(lib().cr_pass_set_clip_depth(self._h, clip_depth))

    def save_alpha_context(self, alpha_layer: int) -> None:  # Renderer::save_alpha_context
        _check(lib().cr_pass_save_alpha_context(self._h, alpha_layer))

    def restore_alpha_context(self, alpha_layer: int) -> None:  # Renderer::restore_alpha_context
        _check(lib().cr_pass_restore_alpha_context(self._h, alpha_layer))

    def render_batch(self, batch: ShapeBatch, commands: np.ndarray) -> None:
        """commands: [n, 4] u32 rows of (shape_index, instance_begin, instance_end, render_operation)."""
        cmds = np.ascontiguousarray(commands, dtype=np.uint32).reshape(-1, 4)
        _check(lib().cr_pass_render_batch(self._h, batch._h, cmds.ctypes.data, len(cmds)))

    def render_script(self, batch: ShapeBatch, script: np.ndarray) -> None:
        """script: [n, 7] u32 rows of (shape_index, instance_begin, instance_end, render_operation, clip_depth, save_alpha_layer,
        restore_alpha_layer): `cr_pass_render_script`, i.e. the state calls (where the state changes) and the draw, per row."""
        rows = np.ascontiguousarray(script, dtype=np.uint32).reshape(-1, 7)
        _check(lib().cr_pass_render_script(self._h, batch._h, rows.ctypes.data, len(rows)))

    def submit(self) -> None:
        h, self._h = self._h, None
        try:
            _check(lib().cr_pass_submit(h))
        finally:
            self._keep = []

    def close(self) -> None:
        """Dropping the pass without submitting it: nothing runs."""
        if self._h is not None and self._h.value:
            lib().cr_pass_abort(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def orthographic_transform(width: float, height: float) -> np.ndarray:
    """Instance mat4 (column vectors, src/shaders.wgsl:13-27) mapping pixel coordinates [0,W]x[0,H] (y down) to NDC."""
    m = np.zeros(16, np.float32)
    m[0] = 2.0 / width
    m[5] = -2.0 / height
    m[10] = 1.0
    m[12] = -1.0
    m[13] = 1.0
    m[15] = 1.0
    return m


__all__ = ["Renderer", "Shape", "ShapeBatch", "RenderPass", "RenderOperation", "Configuration", "Blending", "CullMode", "Error",
           "NumberOfStencilBitsIsUnsupported", "ClipStackOverflow", "TooManyNestedOpacityGroups", "TooManyDashIntervals",
           "DynamicStrokeOptionsIndexOutOfBounds", "orthographic_transform", "lib", "library_path", "Path"]
