"""ctypes mirror of include/contrast_b200.h (plain C structs; no torch types cross the boundary)."""
import ctypes as C

CR_MAX_DASH_INTERVALS = 4
CR_DASH_PATTERN_CAPACITY = 8
CR_MAX_STEPS_PER_INTERVAL = 4194304

# enum cr_status — 1..5 are `enum Error` of the reference in declaration order (src/error.rs:5-16)
CR_OK = 0
CR_ERR_NUMBER_OF_STENCIL_BITS_IS_UNSUPPORTED = 1
CR_ERR_CLIP_STACK_OVERFLOW = 2
CR_ERR_TOO_MANY_NESTED_OPACITY_GROUPS = 3
CR_ERR_TOO_MANY_DASH_INTERVALS = 4
CR_ERR_DYNAMIC_STROKE_OPTIONS_INDEX_OUT_OF_BOUNDS = 5
CR_ERR_INVALID_ARGUMENT = 100
CR_ERR_CUDA = 101
CR_ERR_NON_FINITE = 102
CR_ERR_CURVE_STEPS_CAPACITY = 103
CR_ERR_CUBIC_TRIANGULATION = 104
CR_ERR_NO_DEVICE = 105
CR_ERR_NOT_RESIZED = 106

CR_SEG_LINE, CR_SEG_INTEGRAL_QUADRATIC, CR_SEG_INTEGRAL_CUBIC, CR_SEG_RATIONAL_QUADRATIC, CR_SEG_RATIONAL_CUBIC = range(5)
CR_MEM_HOST, CR_MEM_DEVICE = 0, 1
CR_FORMAT_RGBA32F, CR_FORMAT_RGBA8_UNORM, CR_FORMAT_BGRA8_UNORM = 0, 1, 2
CR_STROKE_FLAG_STROKED, CR_STROKE_FLAG_CLOSED, CR_STROKE_FLAG_UNIFORM_TANGENT_ANGLE = 1, 2, 4
SEGMENT_FLOATS = (2, 4, 6, 5, 10)  # floats per segment of each type


class _Approx(C.Union):
    _fields_ = [("angle_step", C.c_float), ("steps", C.c_uint32)]


class StrokeOptionsC(C.Structure):
    _fields_ = [("width", C.c_float), ("offset", C.c_float), ("miter_clip", C.c_float), ("flags", C.c_uint32),
                ("dynamic_stroke_options_group", C.c_uint32), ("approximation", _Approx)]


class DashIntervalC(C.Structure):
    _fields_ = [("gap_start", C.c_float), ("gap_end", C.c_float), ("dash_start", C.c_uint32), ("dash_end", C.c_uint32)]


class DynamicStrokeOptionsC(C.Structure):
    _fields_ = [("dashed", C.c_uint32), ("join", C.c_uint32), ("start", C.c_uint32), ("end", C.c_uint32),
                ("pattern_len", C.c_uint32), ("phase", C.c_float), ("pattern", DashIntervalC * CR_DASH_PATTERN_CAPACITY)]


class PathSoAC(C.Structure):
    _fields_ = [("n_paths", C.c_uint32), ("n_segments", C.c_uint32), ("memory_space", C.c_uint32), ("_reserved", C.c_uint32),
                ("start", C.c_void_p), ("segment_begin", C.c_void_p), ("segment_types", C.c_void_p), ("type_begin", C.c_void_p),
                ("line_segments", C.c_void_p), ("integral_quadratic", C.c_void_p), ("integral_cubic", C.c_void_p),
                ("rational_quadratic", C.c_void_p), ("rational_cubic", C.c_void_p), ("stroke_options", C.c_void_p)]


class ConfigC(C.Structure):
    _fields_ = [("msaa_sample_count", C.c_uint32), ("clip_nesting_counter_bits", C.c_uint32), ("winding_counter_bits", C.c_uint32),
                ("alpha_layer_count", C.c_uint32), ("blending", C.c_uint32), ("cull_mode", C.c_uint32), ("device", C.c_int32),
                ("depth_compare", C.c_uint32), ("depth_write_enabled", C.c_uint32), ("color_format", C.c_uint32), ("_reserved", C.c_uint32 * 2)]


class ShapeLayoutC(C.Structure):
    _fields_ = [("vertex_offsets", C.c_uint64 * 8), ("index_offsets", C.c_uint64 * 3), ("dynamic_stroke_options_count", C.c_uint64),
                ("proto_hull_points", C.c_uint64)]


class DrawCommandC(C.Structure):
    _fields_ = [("shape_index", C.c_uint32), ("instance_begin", C.c_uint32), ("instance_end", C.c_uint32),
                ("render_operation", C.c_uint32)]


class StatsC(C.Structure):
    _fields_ = [("covered_samples", C.c_uint64), ("primitives", C.c_uint64), ("tile_pairs", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("tessellated_paths", C.c_uint64), ("vertex_bytes", C.c_uint64),
                ("input_bytes", C.c_uint64), ("last_tess_ms", C.c_float), ("last_raster_ms", C.c_float),
                ("last_bin_ms", C.c_float), ("last_hull_sort_ms", C.c_float), ("last_hull_chain_ms", C.c_float), ("_reserved", C.c_float),
                ("proto_hull_points", C.c_uint64), ("hull_vertices", C.c_uint64)]


CR_IPC_HANDLE_BYTES = 64
CR_TILE = 16

# Every symbol include/contrast_b200.h declares (checked against the built library by tests/test_abi.py).
EXPORTED_SYMBOLS = (
    "cr_renderer_create", "cr_renderer_destroy", "cr_renderer_get_config", "cr_renderer_resize", "cr_renderer_set_stream",
    "cr_renderer_synchronize", "cr_renderer_set_pipelining", "cr_shape_from_paths", "cr_shape_destroy", "cr_shape_batch_from_paths", "cr_shape_batch_destroy",
    "cr_shape_batch_size", "cr_shape_batch_get", "cr_shape_set_dynamic_stroke_options", "cr_shape_batch_set_dynamic_stroke_options",
    "cr_shape_get_layout", "cr_shape_read_vertex_buffer", "cr_shape_read_index_buffer", "cr_shape_read_stroke_buffer",
    "cr_pass_begin", "cr_pass_begin_depth", "cr_renderer_read_depth", "cr_renderer_read_color_texels", "cr_renderer_read_color_texels_async", "cr_renderer_wait_readback", "cr_pass_set_instances", "cr_pass_set_clip_depth", "cr_pass_save_alpha_context",
    "cr_pass_restore_alpha_context", "cr_shape_render", "cr_pass_render_batch", "cr_pass_render_script", "cr_pass_submit", "cr_pass_abort", "cr_renderer_read_color",
    "cr_renderer_read_stencil", "cr_renderer_read_alpha_layer", "cr_renderer_get_attachments", "cr_renderer_get_stats", "cr_renderer_get_settled_pass_stats",
    "cr_renderer_enable_timing", "cr_renderer_set_tile_sharding", "cr_renderer_set_order_sharding", "cr_renderer_export_exchange", "cr_renderer_import_peer_exchange", "cr_renderer_export_attachments", "cr_renderer_import_peer_attachments",
    "cr_status_string", "cr_last_error_message", "cr_abi_version",
)
