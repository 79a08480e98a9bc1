/* contrast_b200.h — C-ABI of libcontrast_b200.so: the B200-native drop-in for the tessellate → stencil-then-cover
 * hot path of Lichtso/contrast_renderer (reference @ a189d64).
 *
 * The reference has no FFI; its boundary for this path is the Rust public API of src/renderer.rs + src/path.rs.
 * Every entry point below names the reference item it replaces (file:line relative to /root/reference).
 * Plain pointers and sizes only — no torch / CUDA types in any signature. A Rust `extern "C"` block binds these
 * 1:1.
 *
 * Threading: one caller thread per cr_renderer (the reference never spawns threads). All device work is ordered
 * on one CUDA stream owned by the renderer (or adopted via cr_renderer_set_stream).
 * Memory: every input array may live in host memory (memory_space = CR_MEM_HOST; the library stages it with
 * cudaMemcpyAsync) or already in device memory (CR_MEM_DEVICE; used as-is, zero copies).
 */
#ifndef CONTRAST_B200_H
#define CONTRAST_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ----------------------------------------------------------------------------------------------- status codes
 * 1..5 are the variants of `enum Error` in declaration order (src/error.rs:5-16). Conditions on which the
 * reference panics instead of returning (SafeFloat finite assert src/safe_float.rs:46,114; cubic quadrilateral
 * asserts src/fill.rs:174,178) are reported as codes >= 100 instead of aborting the process. */
typedef enum cr_status {
    CR_OK = 0,
    CR_ERR_NUMBER_OF_STENCIL_BITS_IS_UNSUPPORTED = 1,     /* src/renderer.rs:433 */
    CR_ERR_CLIP_STACK_OVERFLOW = 2,                       /* src/renderer.rs:933 */
    CR_ERR_TOO_MANY_NESTED_OPACITY_GROUPS = 3,            /* src/renderer.rs:947,980 */
    CR_ERR_TOO_MANY_DASH_INTERVALS = 4,                   /* src/renderer.rs:32 */
    CR_ERR_DYNAMIC_STROKE_OPTIONS_INDEX_OUT_OF_BOUNDS = 5,/* src/renderer.rs:189,366 */
    CR_ERR_INVALID_ARGUMENT = 100,
    CR_ERR_CUDA = 101,
    CR_ERR_NON_FINITE = 102,               /* reference: assert!(value.is_finite()) panic */
    CR_ERR_CURVE_STEPS_CAPACITY = 103,     /* more than CR_MAX_STEPS_PER_INTERVAL samples in one inflection-free interval */
    CR_ERR_CUBIC_TRIANGULATION = 104,      /* reference: assert_eq!/assert_ne! panic in fill.rs:174,178 */
    CR_ERR_NO_DEVICE = 105,
    CR_ERR_NOT_RESIZED = 106
} cr_status;

/* ------------------------------------------------------------------------------------------------ enumerations */
typedef enum cr_segment_type {            /* enum SegmentType, src/path.rs:56-67 */
    CR_SEG_LINE = 0,
    CR_SEG_INTEGRAL_QUADRATIC = 1,
    CR_SEG_INTEGRAL_CUBIC = 2,
    CR_SEG_RATIONAL_QUADRATIC = 3,
    CR_SEG_RATIONAL_CUBIC = 4
} cr_segment_type;

typedef enum cr_join { CR_JOIN_MITER = 0, CR_JOIN_BEVEL = 1, CR_JOIN_ROUND = 2 } cr_join;  /* src/path.rs:71-82 */
typedef enum cr_cap {                                                                       /* src/path.rs:86-101 */
    CR_CAP_SQUARE = 0, CR_CAP_ROUND = 1, CR_CAP_OUT = 2, CR_CAP_IN = 3, CR_CAP_RIGHT = 4, CR_CAP_LEFT = 5, CR_CAP_BUTT = 6
} cr_cap;

typedef enum cr_render_operation {        /* enum RenderOperation, src/renderer.rs:145-160 */
    CR_OP_STENCIL = 0,
    CR_OP_CLIP = 1,
    CR_OP_UNCLIP = 2,
    CR_OP_COLOR = 3,
    CR_OP_SAVE_ALPHA_CONTEXT = 4,
    CR_OP_SCALE_ALPHA_CONTEXT = 5,
    CR_OP_RESTORE_ALPHA_CONTEXT = 6
} cr_render_operation;

typedef enum cr_memory_space { CR_MEM_HOST = 0, CR_MEM_DEVICE = 1 } cr_memory_space;
typedef enum cr_blending { CR_BLEND_PREMULTIPLIED_OVER = 0, CR_BLEND_REPLACE = 1 } cr_blending;
typedef enum cr_cull_mode { CR_CULL_NONE = 0, CR_CULL_FRONT = 1, CR_CULL_BACK = 2 } cr_cull_mode;
/* wgpu::CompareFunction with its own numbering (Configuration::depth_compare, src/renderer.rs:388); 0 reads as Always so that
 * a zero-initialised cr_config has no depth test. */
typedef enum cr_compare_function {
    CR_COMPARE_DEFAULT_ALWAYS = 0, CR_COMPARE_NEVER = 1, CR_COMPARE_LESS = 2, CR_COMPARE_EQUAL = 3, CR_COMPARE_LESS_EQUAL = 4,
    CR_COMPARE_GREATER = 5, CR_COMPARE_NOT_EQUAL = 6, CR_COMPARE_GREATER_EQUAL = 7, CR_COMPARE_ALWAYS = 8
} cr_compare_function;
/* Format of the colour attachment (Configuration::blending.format, src/renderer.rs:382; the demo's surface is Bgra8Unorm,
 * examples/application_framework.rs:175). RGBA32F is the parity mode: no quantisation between blend operations. With an 8-bit
 * format every blend result is stored as unorm8 (clamp, x * 255 + 0.5, truncate) and the alpha layers are R8Unorm like the
 * reference's (src/renderer.rs:783,898). */
typedef enum cr_color_format { CR_FORMAT_RGBA32F = 0, CR_FORMAT_RGBA8_UNORM = 1, CR_FORMAT_BGRA8_UNORM = 2 } cr_color_format;

#define CR_MAX_DASH_INTERVALS 4           /* src/path.rs:121 */
#define CR_DASH_PATTERN_CAPACITY 8        /* struct capacity; > CR_MAX_DASH_INTERVALS yields CR_ERR_TOO_MANY_DASH_INTERVALS */
#define CR_MAX_STEPS_PER_INTERVAL 4194304 /* samples of one interpolate_normal! run (src/curve.rs:228-252): keeps the 32-bit vertex counts of a path from wrapping */

/* ------------------------------------------------------------------------------------------------ path model */

/* struct StrokeOptions + CurveApproximation, src/path.rs:153-192. 24 bytes. */
#define CR_STROKE_FLAG_STROKED 1u          /* Path::stroke_options is Some (src/path.rs:215) */
#define CR_STROKE_FLAG_CLOSED 2u           /* StrokeOptions::closed */
#define CR_STROKE_FLAG_UNIFORM_TANGENT_ANGLE 4u /* else UniformlySpacedParameters */
typedef struct cr_stroke_options {
    float width;
    float offset;
    float miter_clip;
    uint32_t flags;
    uint32_t dynamic_stroke_options_group;
    union {
        float angle_step;                  /* CurveApproximation::UniformTangentAngle */
        uint32_t steps;                    /* CurveApproximation::UniformlySpacedParameters */
    } approximation;
} cr_stroke_options;

/* struct DashInterval / enum DynamicStrokeOptions, src/path.rs:105-149 */
typedef struct cr_dash_interval {
    float gap_start;
    float gap_end;
    uint32_t dash_start;                   /* cr_cap */
    uint32_t dash_end;                     /* cr_cap */
} cr_dash_interval;

typedef struct cr_dynamic_stroke_options {
    uint32_t dashed;                       /* 0: Solid{join,start,end}; 1: Dashed{join,pattern,phase} */
    uint32_t join;                         /* cr_join */
    uint32_t start;                        /* cr_cap (Solid) */
    uint32_t end;                          /* cr_cap (Solid) */
    uint32_t pattern_len;                  /* Dashed */
    float phase;                           /* Dashed */
    cr_dash_interval pattern[CR_DASH_PATTERN_CAPACITY];
} cr_dynamic_stroke_options;

/* A set of `Path`s (src/path.rs:213-230) in structure-of-arrays form. The five per-type arrays are the
 * concatenation, in path order, of each Path's `line_segments`, `integral_quadratic_curve_segments`, ...;
 * `segment_types` is the concatenation of every Path's `segment_types`. Cursor tables say where a path's slice
 * of each array begins, so a path is walked with five cursors exactly like the five iterators in
 * src/stroke.rs:210-214 / src/fill.rs:273-277. */
typedef struct cr_path_soa {
    uint32_t n_paths;
    uint32_t n_segments;
    uint32_t memory_space;                 /* cr_memory_space of every pointer below */
    uint32_t _reserved;
    const float* start;                    /* [n_paths][2]                Path::start */
    const uint32_t* segment_begin;         /* [n_paths + 1]               into segment_types */
    const uint8_t* segment_types;          /* [n_segments]                cr_segment_type */
    const uint32_t* type_begin;            /* [5][n_paths + 1]            row t = cursor table of per-type array t */
    const float* line_segments;            /* [n_line][2]                 control_points[0] */
    const float* integral_quadratic;       /* [n_iq][4]                   control_points[0..2] */
    const float* integral_cubic;           /* [n_ic][6]                   control_points[0..3] */
    const float* rational_quadratic;       /* [n_rq][5]                   weight, control_points[0..2] */
    const float* rational_cubic;           /* [n_rc][10]                  weights[0..4], control_points[0..3] */
    const cr_stroke_options* stroke_options; /* [n_paths]; flags & STROKED == 0 means a filled Path; NULL = every Path is filled */
} cr_path_soa;

/* --------------------------------------------------------------------------------------------- renderer setup */

/* struct Configuration, src/renderer.rs:380-405, minus the wgpu-only fields. Depth: gl_Position = instance_transform *
 * vec4(position, 0, 1) (src/shaders.wgsl:72), so clip z = col0.z * x + col1.z * y + col3.z; the colour cover is the only
 * pipeline with a depth test and depth writes (src/renderer.rs:743-745; every other pipeline: Always / false), and a depth
 * failure keeps the stencil value (depth_fail_op: Keep, src/renderer.rs:442). The depth attachment is 32-bit float per
 * sample (a conforming Depth24Plus, examples/showcase/main.rs:46) and only exists when depth_compare / depth_write ask for it. */
typedef struct cr_config {
    uint32_t msaa_sample_count;            /* 1 or 4 */
    uint32_t clip_nesting_counter_bits;
    uint32_t winding_counter_bits;
    uint32_t alpha_layer_count;
    uint32_t blending;                     /* cr_blending of the colour cover */
    uint32_t cull_mode;                    /* cr_cull_mode of the colour cover */
    int32_t device;                        /* CUDA device ordinal; -1 = current */
    uint32_t depth_compare;                /* cr_compare_function of the colour cover (src/renderer.rs:744) */
    uint32_t depth_write_enabled;          /* src/renderer.rs:745 */
    uint32_t color_format;                 /* cr_color_format */
    uint32_t _reserved[2];
} cr_config;

typedef struct cr_renderer cr_renderer;   /* struct Renderer, src/renderer.rs:408 (+ the colour / stencil attachments) */
typedef struct cr_shape cr_shape;         /* struct Shape,    src/renderer.rs:163 */
typedef struct cr_shape_batch cr_shape_batch; /* many Shapes tessellated by one launch sequence */
typedef struct cr_pass cr_pass;           /* wgpu::RenderPass as used by src/renderer.rs:267-355 */

/* Renderer::new, src/renderer.rs:432. */
int cr_renderer_create(const cr_config* config, cr_renderer** out);
void cr_renderer_destroy(cr_renderer* renderer);
/* Renderer::get_config, src/renderer.rs:887. */
int cr_renderer_get_config(const cr_renderer* renderer, cr_config* out);
/* Renderer::resize_internal_buffers, src/renderer.rs:892. Also (re)allocates the colour (RGBA32F per sample)
 * and stencil (u8 per sample) attachments, which the reference's caller owns as wgpu textures. */
int cr_renderer_resize(cr_renderer* renderer, uint32_t width, uint32_t height);
/* Adopt an existing cudaStream_t (passed as void*); NULL restores the renderer's own stream. */
int cr_renderer_set_stream(cr_renderer* renderer, void* cuda_stream);
int cr_renderer_synchronize(cr_renderer* renderer);
/* Frame pipelining (no reference counterpart; wgpu pipelines frames by itself): Shape::from_paths work runs on a second stream
 * of the renderer and a rebuild of a batch (`existing`) writes the set of arrays the last pass is NOT reading, so tessellating
 * frame N + 1 overlaps rasterising frame N. Passes wait on the device for the builds they render; results observed through this
 * API are ordered as before. The one difference for the caller: input arrays in DEVICE memory must be complete when
 * cr_shape_from_paths / cr_shape_batch_from_paths is called (work merely enqueued on the renderer's stream is not waited for). */
int cr_renderer_set_pipelining(cr_renderer* renderer, uint32_t enabled);

/* ---------------------------------------------------------------------------------------------- shape building */

/* Shape::from_paths, src/renderer.rs:177. `existing` (may be NULL) is consumed: its device buffers are reused in
 * place when the byte lengths are unchanged (Buffer::update, src/renderer.rs:89-95). */
int cr_shape_from_paths(cr_renderer* renderer, const cr_dynamic_stroke_options* dynamic_stroke_options,
                        size_t dynamic_stroke_options_count, const cr_path_soa* paths, cr_shape* existing,
                        cr_shape** out);
void cr_shape_destroy(cr_shape* shape);

/* Batched Shape::from_paths: shape s is built from paths [shape_path_begin[s], shape_path_begin[s+1]) (host array
 * of n_shapes + 1 entries); all shapes share `dynamic_stroke_options`. One count / scan / emit launch sequence
 * covers every path of every shape. `existing` (may be NULL) is consumed and its allocations are reused. */
int cr_shape_batch_from_paths(cr_renderer* renderer, const cr_dynamic_stroke_options* dynamic_stroke_options,
                              size_t dynamic_stroke_options_count, const cr_path_soa* paths,
                              const uint32_t* shape_path_begin, uint32_t n_shapes, cr_shape_batch* existing,
                              cr_shape_batch** out);
void cr_shape_batch_destroy(cr_shape_batch* batch);
uint32_t cr_shape_batch_size(const cr_shape_batch* batch);
/* Borrowed view of shape `index`; valid until the batch is destroyed. Do not pass it to cr_shape_destroy. */
cr_shape* cr_shape_batch_get(cr_shape_batch* batch, uint32_t index);

/* Shape::set_dynamic_stroke_options, src/renderer.rs:360 — a 48-byte in-place write at index * 48. */
int cr_shape_set_dynamic_stroke_options(cr_shape* shape, size_t index, const cr_dynamic_stroke_options* options);
int cr_shape_batch_set_dynamic_stroke_options(cr_shape_batch* batch, size_t index,
                                              const cr_dynamic_stroke_options* options);

/* The private fields of struct Shape (src/renderer.rs:163-171), exposed for parity checks: cumulative BYTE ends
 * of [line | joint | solid | integral quadratic | integral cubic | rational quadratic | rational cubic | hull]
 * in the vertex buffer and of [line | joint | solid] in the u16 index buffer (concat_buffers!, :121-141,198-209). */
typedef struct cr_shape_layout {
    uint64_t vertex_offsets[8];
    uint64_t index_offsets[3];
    uint64_t dynamic_stroke_options_count;
    uint64_t proto_hull_points;            /* size of `proto_hull` before convex_hull::andrew (src/renderer.rs:184) */
} cr_shape_layout;
int cr_shape_get_layout(cr_shape* shape, cr_shape_layout* out);
/* Device→host copies in the reference's exact byte layout (src/vertex.rs:1-26; u16 indices with 0xFFFF restarts,
 * wrapping like `start_index as u16`, src/stroke.rs:108,128, src/fill.rs:363; 48-byte DynamicStrokeDescriptor,
 * src/renderer.rs:18-27). `capacity` must be at least the size reported by cr_shape_get_layout. */
int cr_shape_read_vertex_buffer(cr_shape* shape, void* dst, size_t capacity);
int cr_shape_read_index_buffer(cr_shape* shape, void* dst, size_t capacity);
int cr_shape_read_stroke_buffer(cr_shape* shape, void* dst, size_t capacity);

/* ------------------------------------------------------------------------------------------------- render pass */

/* wgpu begin_render_pass with LoadOp::Clear / LoadOp::Load (examples/showcase/main.rs:211-234): clear colour is
 * transparent black, clear stencil is 0. As in wgpu the clear is part of the pass and executes with it at cr_pass_submit
 * (the tile kernel starts cleared tiles from zero and writes every tile: no memset, no read of the old contents); a pass
 * that is aborted clears nothing. */
int cr_pass_begin(cr_renderer* renderer, uint32_t clear_color, uint32_t clear_stencil, cr_pass** out);
/* The same with the depth aspect's load operation spelled out (depth_ops: LoadOp::Clear(1.0) in the demo, main.rs:222-225).
 * cr_pass_begin clears depth to 1.0 whenever it clears the stencil (one depth-stencil attachment). */
int cr_pass_begin_depth(cr_renderer* renderer, uint32_t clear_color, uint32_t clear_stencil, uint32_t clear_depth,
                        float depth_clear_value, cr_pass** out);
/* Vertex buffer slot 0 (instance mat4: four vec4 that become the matrix COLUMNS, src/shaders.wgsl:13-27,
 * src/renderer.rs:462-466; 64 B each) and the instance colour slot (16 B each, src/renderer.rs:502-506).
 * `colors` may be NULL if no colour-consuming operation is recorded. */
int cr_pass_set_instances(cr_pass* pass, const float* transforms, const float* colors, uint32_t count,
                          uint32_t memory_space);
/* Renderer::set_clip_depth, src/renderer.rs:932. */
int cr_pass_set_clip_depth(cr_pass* pass, uint32_t clip_depth);
/* Renderer::save_alpha_context, src/renderer.rs:941: selects the R8 layer the next SAVE_ALPHA_CONTEXT writes. */
int cr_pass_save_alpha_context(cr_pass* pass, uint32_t alpha_layer);
/* Renderer::restore_alpha_context, src/renderer.rs:979: selects the layer the next RESTORE_ALPHA_CONTEXT reads. */
int cr_pass_restore_alpha_context(cr_pass* pass, uint32_t alpha_layer);
/* Shape::render, src/renderer.rs:267. Records; nothing runs until cr_pass_submit.
 * Rasterisation follows the WebGPU rules the reference relies on (restated in oracle/raster.hpp): clip = M (x, y, 0, 1)
 * (src/shaders.wgsl:72), 1/256-pixel snapping, top-left fill rule, perspective-correct per-sample attributes, the standard 4x
 * sample pattern. Primitives with a corner on or behind the eye plane (w <= 0), or further than 2^21 pixels away, are CLIPPED
 * in clip space (eye plane + guard band) like a GPU clips them, not dropped; depth is not clipped against the near / far planes
 * (unclipped-depth semantics), the depth test of the colour cover compares z / w against the f32 depth attachment. */
int cr_shape_render(cr_pass* pass, cr_shape* shape, uint32_t instance_begin, uint32_t instance_end,
                    uint32_t render_operation);

/* Bulk recording: command i renders shape `shape_index[i]` of `batch` for instances
 * [instance_begin[i], instance_end[i]) with operation `operation[i]`; host arrays. Equivalent to n calls of
 * cr_shape_render in order (the stencil reference / alpha layer state current at the time of the call applies). */
typedef struct cr_draw_command {
    uint32_t shape_index;
    uint32_t instance_begin;
    uint32_t instance_end;
    uint32_t render_operation;
} cr_draw_command;
int cr_pass_render_batch(cr_pass* pass, cr_shape_batch* batch, const cr_draw_command* commands, size_t count);
/* Bulk recording WITH the pass state: entry i is equivalent to cr_pass_set_clip_depth(clip_depth),
 * cr_pass_save_alpha_context(save_alpha_layer), cr_pass_restore_alpha_context(restore_alpha_layer) (each only when the
 * value differs from the pass's current one, with the same errors) followed by cr_shape_render. One call records a whole
 * clip / opacity-group script (src/renderer.rs:253-266) — 60 000 draws of BASELINE config 4 in ~1 ms of host time. */
typedef struct cr_scripted_draw {
    uint32_t shape_index;
    uint32_t instance_begin;
    uint32_t instance_end;
    uint32_t render_operation;
    uint32_t clip_depth;
    uint32_t save_alpha_layer;
    uint32_t restore_alpha_layer;
} cr_scripted_draw;
int cr_pass_render_script(cr_pass* pass, cr_shape_batch* batch, const cr_scripted_draw* draws, size_t count);

/* queue.submit(encoder.finish()) (examples/showcase/main.rs:252): bins, sorts and rasterises everything
 * recorded, asynchronously on the renderer's stream. The pass object is consumed. */
int cr_pass_submit(cr_pass* pass);
/* Dropping a wgpu::RenderPass / CommandEncoder without submitting it: frees the pass, runs nothing. */
void cr_pass_abort(cr_pass* pass);

/* Attachment read-back (tests / image dump). color: [height][width][samples][4] f32 premultiplied;
 * stencil: [height][width][samples] u8 (clip bits << winding bits | winding bits);
 * alpha layer: [height][width][samples] f32. These synchronise the stream. */
int cr_renderer_read_color(cr_renderer* renderer, float* dst, size_t capacity_bytes);
/* The colour attachment as stored: 16 bytes per sample (RGBA32F) or one packed unorm8 texel per sample (RGBA8 / BGRA8). This
 * is the frame a presenter consumes; `dst` should be pinned host memory for the copy to run at PCIe speed. */
int cr_renderer_read_color_texels(cr_renderer* renderer, void* dst, size_t capacity_bytes);
/* The same frame without stopping the pipeline (wgpu: copy_texture_to_buffer + map_async): the attachment is snapshot on the
 * renderer's stream behind the pass submitted last and copied to `dst` (pinned host memory, untouched until waited for) on a
 * stream of its own while the next frames are rendered. At most four read-backs are in flight; a fifth waits for the oldest.
 * cr_renderer_wait_readback returns when `ticket`'s bytes have arrived (and are the frame of a pass that was not skipped). */
int cr_renderer_read_color_texels_async(cr_renderer* renderer, void* dst, size_t capacity_bytes, uint64_t* ticket);
int cr_renderer_wait_readback(cr_renderer* renderer, uint64_t ticket);
int cr_renderer_read_stencil(cr_renderer* renderer, uint8_t* dst, size_t capacity_bytes);
int cr_renderer_read_alpha_layer(cr_renderer* renderer, uint32_t layer, float* dst, size_t capacity_bytes);
/* depth: [height][width][samples] f32; CR_ERR_INVALID_ARGUMENT if the configuration has no depth attachment. */
int cr_renderer_read_depth(cr_renderer* renderer, float* dst, size_t capacity_bytes);
/* Device pointers of the attachments (for zero-copy consumers and multi-GPU tile exchange). */
int cr_renderer_get_attachments(cr_renderer* renderer, void** color_dev, void** stencil_dev);

/* One render target spanning several GPUs of a node ("NCCL reduce of framebuffer tiles only when one render target spans
 * the box", BASELINE north_star; SURVEY 8e tile sharding). No reference counterpart: wgpu renders on one device.
 * Every rank creates a renderer of the same configuration and extent, records the SAME pass and submits it; tile
 * (tx, ty) of the 16 x 16 pixel grid is owned by rank (tx + ty) % world. A rank bins and rasterises only its own tiles,
 * and the tile kernel stores each finished tile into its own attachments and, with P2P stores over NVLink, into the
 * imported attachments of every other rank: after all ranks have submitted (and a barrier), each rank holds the
 * complete colour and stencil attachments. Alpha layers and cr_stats.covered_samples stay per-owner.
 * The caller orders the ranks: a barrier before cr_pass_submit (no rank may still be reading or clearing the previous
 * frame when another rank's tiles arrive), and one after
 * cr_pass_submit (contrast_renderer_b200/sharding.py does both with stream-ordered NCCL collectives).
 * cr_renderer_set_tile_sharding(r, 1, 0) returns to a single-GPU target and closes the imported handles. */
#define CR_IPC_HANDLE_BYTES 64
int cr_renderer_set_tile_sharding(cr_renderer* renderer, uint32_t world, uint32_t rank);
/* cudaIpcMemHandle_t of the colour / stencil attachments (valid until the next cr_renderer_resize). */
int cr_renderer_export_attachments(cr_renderer* renderer, uint8_t color_handle[CR_IPC_HANDLE_BYTES], uint8_t stencil_handle[CR_IPC_HANDLE_BYTES]);
/* Maps the attachments of rank `peer_rank` (handles from its cr_renderer_export_attachments, in another process). */
int cr_renderer_import_peer_attachments(cr_renderer* renderer, uint32_t peer_rank, const uint8_t color_handle[CR_IPC_HANDLE_BYTES],
                                        const uint8_t stencil_handle[CR_IPC_HANDLE_BYTES]);

/* One render target composed from DRAW-ORDER slices ("path instances shard across GPUs by batch" into one target, BASELINE
 * north_star; SURVEY 8e batch sharding). No reference counterpart. Rank r of `world` tessellates and submits only its contiguous
 * slice of the draw order (slices must not cut through a clip or an opacity group). Per 16 x 16 tile the ranks whose slices
 * touch it form a chain in rank order: each waits for its predecessor's tile state (colour + stencil, stored into ITS
 * attachments over NVLink), rasterises its slice on top — the same operations in the same order as a single GPU would execute,
 * so the composed frame is bit-identical to the single-GPU frame — and hands the tile on; the last rank of the chain stores
 * the finished tile into every rank's attachments. Every rank submits the same number of passes; the caller puts a barrier
 * before cr_pass_submit (nobody still reads the previous frame) and one after it (contrast_renderer_b200/sharding.py).
 * Setup: cr_renderer_set_order_sharding, then exchange the handles of cr_renderer_export_attachments and
 * cr_renderer_export_exchange and import them with cr_renderer_import_peer_attachments / cr_renderer_import_peer_exchange.
 * cr_stats.covered_samples stays per rank (the sum over ranks is the frame's). (world, rank) = (1, 0) switches it off. */
int cr_renderer_set_order_sharding(cr_renderer* renderer, uint32_t world, uint32_t rank);
int cr_renderer_export_exchange(cr_renderer* renderer, uint8_t handle[CR_IPC_HANDLE_BYTES]);
int cr_renderer_import_peer_exchange(cr_renderer* renderer, uint32_t peer_rank, const uint8_t handle[CR_IPC_HANDLE_BYTES]);

/* Counters of the last submitted pass / last from_paths call (synchronises the stream). */
typedef struct cr_stats {
    uint64_t covered_samples;              /* samples that passed the stencil test of a COLOR cover */
    uint64_t primitives;                   /* triangles considered by the binner */
    uint64_t tile_pairs;                   /* (tile, primitive) pairs rasterised */
    uint64_t kernel_launches;              /* kernels of this library launched since renderer creation */
    uint64_t tessellated_paths;
    uint64_t vertex_bytes;                 /* bytes of vertex + index output of the last from_paths */
    uint64_t input_bytes;                  /* bytes of path input of the last from_paths */
    float last_tess_ms;                    /* CUDA-event duration of the emit kernels (0 if timing disabled) */
    float last_raster_ms;                  /* CUDA-event duration of the tile raster kernel */
    float last_bin_ms;
    float last_hull_sort_ms;               /* CUDA-event duration of hull_sort_kernel / hull_chain_kernel of the last from_paths */
    float last_hull_chain_ms;
    float _reserved;
    uint64_t proto_hull_points;            /* points that went through convex_hull::andrew in the last from_paths, over all shapes */
    uint64_t hull_vertices;                /* hull vertices it produced */
} cr_stats;
int cr_renderer_get_stats(cr_renderer* renderer, cr_stats* out);
/* The same record as of the last pass the renderer has settled (cr_pass_submit settles the pass submitted before), without waiting:
 * with frame pipelining the result of frame N is read while frame N + 1 runs (wgpu: mapping a buffer of an earlier submission). */
int cr_renderer_get_settled_pass_stats(cr_renderer* renderer, cr_stats* out);
int cr_renderer_enable_timing(cr_renderer* renderer, uint32_t enabled);

const char* cr_status_string(int status);
const char* cr_last_error_message(void);
uint32_t cr_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* CONTRAST_B200_H */
