// api.cu — the C-ABI of libcontrast_b200.so (include/contrast_b200.h): object lifetimes, host<->device staging and
// the launch sequences that string the kernels of tess.cu / raster.cu / prims.cu together.
//
// Mirrors, entry point by entry point, the Rust public API of the reference for this path:
//   Renderer::new / resize_internal_buffers / set_clip_depth / save_alpha_context / restore_alpha_context
//   (src/renderer.rs:432,892,932,941,979), Shape::from_paths / render / set_dynamic_stroke_options
//   (src/renderer.rs:177,267,360), convert_dynamic_stroke_options (src/renderer.rs:29-60).
// There is NO CPU fallback anywhere in this file: without a CUDA device every entry point that would compute
// returns CR_ERR_NO_DEVICE / CR_ERR_CUDA.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <memory>
#include <new>
#include <vector>
#include "device_common.cuh"
#include "prims.h"
#include "raster.h"
#include "tess.h"

// ----------------------------------------------------------------------------------------------- error plumbing
static thread_local char g_error_message[512] = "";
void cr_set_error_message(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error_message, sizeof(g_error_message), fmt, ap);
    va_end(ap);
}
#define CR_TRY(expr)                     \
    do {                                 \
        const int _st = (expr);          \
        if (_st != CR_OK) return _st;    \
    } while (0)
static int fail(int status, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error_message, sizeof(g_error_message), fmt, ap);
    va_end(ap);
    return status;
}

namespace {

// A grow-only device allocation (stream-ordered allocator, so re-tessellating every frame does not synchronise).
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    bool plain = false;   // cudaMalloc instead of the stream-ordered pool: such memory can be exported with cudaIpcGetMemHandle
    int reserve(cudaStream_t stream, size_t bytes) {
        if (bytes <= cap && p) return CR_OK;
        release(stream);
        const size_t want = bytes < 256 ? 256 : bytes;
        if (plain) CR_CUDA_TRY(cudaMalloc(&p, want));
        else CR_CUDA_TRY(cudaMallocAsync(&p, want, stream));
        cap = want;
        return CR_OK;
    }
    void release(cudaStream_t stream) {
        if (p && plain) { cudaStreamSynchronize(stream); cudaFree(p); }
        else if (p) cudaFreeAsync(p, stream);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct Descriptor48 {   // DynamicStrokeDescriptor (src/renderer.rs:18-27)
    float gap_start[CR_MAX_DASH_INTERVALS];
    float gap_end[CR_MAX_DASH_INTERVALS];
    uint32_t caps;
    uint32_t count_dashed_join;
    float phase;
    uint32_t _padding;
};
static_assert(sizeof(Descriptor48) == 48, "DynamicStrokeDescriptor is 48 bytes");

// convert_dynamic_stroke_options (src/renderer.rs:29-60)
int convert_dynamic_stroke_options(const cr_dynamic_stroke_options& o, Descriptor48& out) {
    memset(&out, 0, sizeof(out));
    if (o.join > CR_JOIN_ROUND) return fail(CR_ERR_INVALID_ARGUMENT, "join %u is not a cr_join", o.join);
    if (o.dashed) {
        if (o.pattern_len > CR_MAX_DASH_INTERVALS) return fail(CR_ERR_TOO_MANY_DASH_INTERVALS, "%u dash intervals > %d", o.pattern_len, CR_MAX_DASH_INTERVALS);
        if (o.pattern_len == 0) return fail(CR_ERR_INVALID_ARGUMENT, "a dashed stroke needs at least one interval");
        out.count_dashed_join = ((o.pattern_len - 1) << 3) | 4u | o.join;
        out.phase = o.phase;
        for (uint32_t i = 0; i < o.pattern_len; ++i) {
            if (o.pattern[i].dash_start > CR_CAP_BUTT || o.pattern[i].dash_end > CR_CAP_BUTT) return fail(CR_ERR_INVALID_ARGUMENT, "bad cap");
            out.gap_start[i] = o.pattern[i].gap_start;
            out.gap_end[i] = o.pattern[i].gap_end;
            out.caps |= o.pattern[i].dash_start << (((i + o.pattern_len - 1) % o.pattern_len) * 8);
            out.caps |= o.pattern[i].dash_end << (i * 8 + 4);
        }
    } else {
        if (o.start > CR_CAP_BUTT || o.end > CR_CAP_BUTT) return fail(CR_ERR_INVALID_ARGUMENT, "bad cap");
        out.caps = o.start | (o.end << 4);
        out.count_dashed_join = o.join;
    }
    return CR_OK;
}

int decode_device_error(uint32_t flags) {
    if (flags & CR_DEVERR_BAD_TABLES) return fail(CR_ERR_INVALID_ARGUMENT, "cr_path_soa: segment_begin / type_begin / segment_types are inconsistent (not monotone, not covering [0, n_segments], a type code above 4, or per-type counts that disagree with the type stream)");
    if (flags & CR_DEVERR_GROUP_OOB) return fail(CR_ERR_DYNAMIC_STROKE_OPTIONS_INDEX_OUT_OF_BOUNDS, "a path references a dynamic stroke options group that does not exist");
    if (flags & CR_DEVERR_NON_FINITE) return fail(CR_ERR_NON_FINITE, "tessellation produced a non-finite hull point");
    if (flags & CR_DEVERR_STEPS) return fail(CR_ERR_CURVE_STEPS_CAPACITY, "more than %d tangent-angle steps in one curve interval", CR_MAX_STEPS_PER_INTERVAL);
    if (flags & CR_DEVERR_CUBIC) return fail(CR_ERR_CUBIC_TRIANGULATION, "cubic control quadrilateral is in no recognised configuration");
    return CR_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------- objects
// Words of the renderer's pinned read-back area.
enum { PIN_TESS = 0 /* 11 totals, err, max_proto, 5 type totals */, PIN_PASS = 24 /* PassCounters snapshot, 10 words */, PIN_MISC = 40 /* one 48-byte descriptor */, PIN_WORDS = 64 };

struct cr_pass;
struct cr_renderer {
    cr_config config;
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    // Frame pipelining: a frame goes through four streams — host inputs are copied on `copy`, count / scan / emit run on `tess`,
    // the convex hulls (sort + sequential chains: latency bound, few warps) on `hull`, the passes on `stream` — so that the copy
    // of frame N + 2, the tessellation of frame N + 1... and the raster of frame N - 1 overlap. Events order the stages.
    cudaStream_t tess = nullptr, copy = nullptr, hull = nullptr;
    bool pipelined = false;
    cudaStream_t tstream() const { return pipelined ? tess : stream; }
    cudaStream_t cstream() const { return pipelined ? copy : stream; }
    cudaStream_t hstream() const { return pipelined ? hull : stream; }
    int staging_cur = 0;              // which of the two staging sets the build being enqueued copies into
    cudaEvent_t staging_free[2] = {nullptr, nullptr};   // recorded on the tessellation stream behind the last kernel that reads the set
    cudaEvent_t ev_copied = nullptr, ev_emitted = nullptr, ev_hull_idle = nullptr;
    uint32_t width = 0, height = 0, tiles_x = 0, tiles_y = 0;
    DevBuf color, stencil, alpha_layers, depth;
    // scratch shared by every from_paths / submit of this renderer
    DevBuf staging[2][10], counts, scan_scratch_tess, scan_scratch, shape_begin_dev, hull_scratch_a, hull_scratch_b;
    DevBuf compact_dev, cmds_dev, batches_dev, cmd_cands, cand_tiles, records, big_list, pair_tile, pair_cand, pair_tile_alt, pair_cand_alt, tile_prims, clip_list, clip_attrs, radix_scratch, tile_begin,
        inst_transforms, inst_colors, pass_counters;
    uint32_t* pinned = nullptr;   // PIN_WORDS words of pinned read-back area
    cr_stats stats{};
    cr_shape_batch* stats_batch = nullptr;   // the batch of the last from_paths, for the hull-vertex statistic (its counts arrive later)
    bool timing = false;
    cudaEvent_t ev[9] = {};   // tess begin/end, bin begin/end, raster begin/end, hull begin / after sort / end
    bool ev_valid[3] = {false, false, false};
    cudaEvent_t ev_sizes = nullptr;   // from_paths: the sizes of the build have reached the pinned area
    cudaEvent_t ev_pass = nullptr;    // submit: the counters of the pass have reached the pinned area
    cudaEvent_t ev_update = nullptr;  // set_dynamic_stroke_options: the 48-byte staging slot has been read
    // Draw commands are recorded straight into a pinned, grow-only arena (one recording pass at a time; a second concurrent
    // pass falls back to a heap vector), so that cr_pass_submit uploads them with a truly asynchronous copy.
    CompactCommand* cmd_arena = nullptr;
    size_t cmd_arena_cap = 0;
    bool cmd_arena_busy = false;
    // Capacities the next pass is sized with (candidates, (tile, candidate) pairs): what the last pass needed plus slack. 0: unknown.
    uint32_t cand_cap = 0, pair_cap = 0, last_cands = 0, last_pairs = 0;
    // asynchronous frame read-back (cr_renderer_read_color_texels_async): the attachment is snapshot on the renderer's stream behind the
    // pass (device-to-device, a few microseconds) and copied to the host on a stream of its own while the next frame is rendered
    struct Readback { void* dst = nullptr; size_t bytes = 0; uint64_t ticket = 0, pass_serial = 0; bool pending = false; cudaEvent_t snapshot = nullptr, done = nullptr; DevBuf buf; };
    Readback readback[4];
    cudaStream_t out = nullptr;
    uint64_t readback_tickets = 0, pass_serial = 0, inflight_serial = 0;
    uint32_t radix_layout_cap = 0xFFFFFFFFu;   // pair capacity the radix scratch is laid out (and zeroed) for
    uint32_t clip_cap = 1024;         // triangles frustum clipping may produce in one pass (grows when a pass needs more)
    cr_pass* inflight = nullptr;      // the last submitted pass until its device-side sizes have been checked (settle)
    int deferred_status = CR_OK;      // an error found while settling, reported by the next entry point that can fail
    char deferred_message[256] = "";
    uint32_t shard_world = 1, shard_rank = 0;              // tile sharding of one target across GPUs
    void* peer_color[CR_MAX_PEERS] = {};                   // peer-mapped attachments of the other ranks, slot = rank - (rank > shard_rank)
    void* peer_stencil[CR_MAX_PEERS] = {};
    uint32_t order_world = 1, order_rank = 0, order_epoch = 0;   // draw-order sharding into one target (batch sharding composed)
    DevBuf exchange;                                       // touched-tile bitmaps and flags (raster.h), exported to the other ranks
    void* peer_exchange[CR_MAX_PEERS] = {};
    uint64_t live_objects = 0;        // shape batches and passes that still point at this renderer
    bool destroy_requested = false;   // cr_renderer_destroy was called while live_objects > 0: the last child frees it
};

struct cr_shape {
    cr_shape_batch* batch;
    uint32_t index;
    bool owns_batch;
};

// The device arrays of one build of a batch. A batch owns CR_BATCH_SETS of them: with frame pipelining
// (cr_renderer_set_pipelining) a rebuild writes the set used least recently, so that tessellating frame N + 1 and building the
// hulls of frame N overlap rasterising frame N - 1.
#define CR_BATCH_SETS 3
struct BatchStorage {
    DevBuf vtx[7], proto, hull, idx[3], cat_begin, hull_count, stroke, desc_dev;
    DevBuf err;                        // [0] error bits of this build, [1] largest proto-hull slice of any shape
    cudaEvent_t written = nullptr;     // recorded on the tessellation stream when the build (and its descriptor) is complete
    cudaEvent_t last_read = nullptr;   // recorded on the renderer's stream behind the last pass that reads these arrays
    bool was_read = false;
};
struct cr_shape_batch {
    cr_renderer* renderer = nullptr;
    uint32_t n_shapes = 0, n_paths = 0, n_groups = 0, n_segments = 0;
    BatchStorage set[CR_BATCH_SETS];
    int cur = 0;                      // the set holding the latest build
    BatchStorage& store() { return set[cur]; }
    const BatchStorage& store() const { return set[cur]; }
    // Host mirrors of the slice tables ([CNT_COUNT][n_shapes + 1] cat_begin, then [n_shapes] hull_count, then the final error
    // word), in pinned memory; they arrive asynchronously (ev_mirrors) and are only waited for by the layout / read-back calls.
    uint32_t* mirrors = nullptr;
    size_t mirrors_cap = 0;
    cudaEvent_t ev_mirrors = nullptr;
    bool mirrors_pending = false;
    bool built = false;               // holds a finished build whose sizes can seed an optimistic rebuild
    bool has_cubics = true;
    uint32_t max_proto = 0;
    std::vector<cr_shape> views;
    uint64_t totals[CNT_COUNT] = {};
    const uint32_t* cat_begin_host() const { return mirrors; }
    const uint32_t* hull_count_host() const { return mirrors + (size_t)CNT_COUNT * (n_shapes + 1); }
};

struct InstanceSet {
    const float* transforms;
    const float* colors;
    uint32_t count, space, base;
};
struct cr_pass {
    cr_renderer* renderer;
    std::vector<CompactCommand> commands;   // heap fallback; the usual home of the commands is the renderer's pinned arena
    bool arena = false;
    size_t n_arena = 0;
    std::vector<cr_shape_batch*> batches;
    std::vector<BatchStorage*> sets;        // (after submit) the build of each batch this pass renders
    std::vector<InstanceSet> instance_sets;
    uint32_t instance_total = 0;
    uint32_t clip_depth = 0, save_layer = 0, restore_layer = 0;
    bool clear_color = false, clear_stencil = false, clear_depth = false;   // LoadOp::Clear of the attachments, executed at submit (fused into the tile kernel)
    float depth_clear_value = 1.0f;
    bool any_color = false;           // (after submit) the instance colour slot was bound
};

namespace {

struct DeviceGuard {
    int previous = -1;
    bool ok = true;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&previous) != cudaSuccess) { ok = false; return; }
        if (previous != device && cudaSetDevice(device) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() { if (previous >= 0) cudaSetDevice(previous); }
};
#define CR_GUARD(r)                                                                    \
    DeviceGuard _guard((r)->device);                                                   \
    if (!_guard.ok) return fail(CR_ERR_CUDA, "cannot select CUDA device %d", (r)->device)

int settle(cr_renderer* r);
// An error found while settling an earlier pass is handed to the caller by the next entry point that returns a status.
int take_deferred(cr_renderer* r) {
    if (r->deferred_status == CR_OK) return CR_OK;
    const int st = r->deferred_status;
    r->deferred_status = CR_OK;
    return fail(st, "%s", r->deferred_message);
}

void batch_release(cr_shape_batch* b) {
    cudaStream_t st = b->renderer->tstream();
    for (BatchStorage& B : b->set) {
        if (B.was_read && B.last_read) cudaStreamWaitEvent(st, B.last_read, 0);   // freed behind the last pass that reads them
        for (auto& v : B.vtx) v.release(st);
        for (auto& v : B.idx) v.release(st);
        if (B.written) cudaStreamWaitEvent(st, B.written, 0);                     // ... and behind the build itself (the hull stream records it)
        B.proto.release(st); B.hull.release(st); B.cat_begin.release(st); B.hull_count.release(st); B.stroke.release(st); B.desc_dev.release(st); B.err.release(st);
        if (B.written) cudaEventDestroy(B.written);
        if (B.last_read) cudaEventDestroy(B.last_read);
        B.written = B.last_read = nullptr;
        B.was_read = false;
    }
    if (b->mirrors_pending && b->ev_mirrors) cudaEventSynchronize(b->ev_mirrors);
    if (b->mirrors) cudaFreeHost(b->mirrors);
    if (b->ev_mirrors) cudaEventDestroy(b->ev_mirrors);
    b->mirrors = nullptr; b->mirrors_cap = 0; b->ev_mirrors = nullptr; b->mirrors_pending = false; b->built = false;
}

// The scan scratch carries ticket counters that must start at zero (prims.h): zero it whenever it is (re)allocated.
int reserve_scan_scratch(cr_renderer* r, cudaStream_t st, DevBuf& buf, size_t words) {
    (void)r;
    if (buf.p && words * 4 <= buf.cap) return CR_OK;
    const size_t want = std::max<size_t>(words * 2, 4096);
    CR_TRY(buf.reserve(st, want * 4));
    return cr_scan_prepare(st, buf.as<uint32_t>(), buf.cap / 4);
}

// Bring one input array to the device: host memory is staged (cudaMemcpyAsync on the renderer's stream), device
// memory is used where it lies.
template <typename T>
int stage(cr_renderer* r, int slot, const T* src, size_t count, uint32_t space, const T** out) {
    if (count == 0) { *out = nullptr; return CR_OK; }
    if (!src) return fail(CR_ERR_INVALID_ARGUMENT, "null input array (slot %d)", slot);
    if (space == CR_MEM_DEVICE) { *out = src; return CR_OK; }
    DevBuf& buf = r->staging[r->staging_cur][slot];
    CR_TRY(buf.reserve(r->cstream(), count * sizeof(T)));
    CR_CUDA_TRY(cudaMemcpyAsync(buf.p, src, count * sizeof(T), cudaMemcpyHostToDevice, r->cstream()));
    *out = buf.as<T>();
    return CR_OK;
}

// The host mirrors of a batch's slice tables: wait for them (and for the build's final error word) if they are still in flight.
int ensure_mirrors(cr_shape_batch* b) {
    if (!b->mirrors_pending) return CR_OK;
    CR_CUDA_TRY(cudaEventSynchronize(b->ev_mirrors));
    b->mirrors_pending = false;
    const uint32_t err = b->mirrors[(size_t)CNT_COUNT * (b->n_shapes + 1) + b->n_shapes];
    return decode_device_error(err & ~CR_DEVERR_FATAL_MASK);   // an emit-pass error of an optimistic rebuild (non-finite vertex) surfaces here
}

// Shape::from_paths for many shapes. Two ways through:
//  * cold (no previous build to go by): count, scan, bounds; the host waits for the sizes, allocates, then emit + hull;
//  * optimistic (`b` holds a finished build of the same numbers of paths / shapes / segments): emit + hull are enqueued
//    behind the count pass at once, writing into the previous build's arrays; shape_bounds_kernel checks on the device that
//    they are large enough (else nothing is written and the host falls back to the cold order). The host only waits for
//    the event after the size read-back, never for the stream: the GPU is already running the emit pass by then.
int build_batch(cr_renderer* r, const cr_dynamic_stroke_options* groups, size_t n_groups, const cr_path_soa* soa, const uint32_t* shape_path_begin,
                uint32_t n_shapes, cr_shape_batch* b) {
    if (!soa) return fail(CR_ERR_INVALID_ARGUMENT, "paths is null");
    if (n_groups > 65536) return fail(CR_ERR_INVALID_ARGUMENT, "at most 65536 dynamic stroke option groups (the vertex flag word keeps the group in 16 bits, src/shaders.wgsl:273)");
    if (n_groups && !groups) return fail(CR_ERR_INVALID_ARGUMENT, "dynamic_stroke_options is null");
    if (soa->memory_space > CR_MEM_DEVICE) return fail(CR_ERR_INVALID_ARGUMENT, "bad memory_space");
    if (n_shapes == 0 || !shape_path_begin) return fail(CR_ERR_INVALID_ARGUMENT, "a batch needs at least one shape");
    if (shape_path_begin[0] != 0 || shape_path_begin[n_shapes] != soa->n_paths) return fail(CR_ERR_INVALID_ARGUMENT, "shape_path_begin must cover [0, n_paths]");
    for (uint32_t s = 0; s < n_shapes; ++s)
        if (shape_path_begin[s] > shape_path_begin[s + 1]) return fail(CR_ERR_INVALID_ARGUMENT, "shape_path_begin must be non-decreasing");
    cudaStream_t st = r->tstream(), hs = r->hstream();
    const uint32_t n_paths = soa->n_paths;
    const size_t stride = (size_t)n_paths + 1;
    if (n_paths && !soa->type_begin) return fail(CR_ERR_INVALID_ARGUMENT, "type_begin is null");

    // descriptors first: TooManyDashIntervals is reported before any tessellation work (src/renderer.rs:210-215 runs
    // after the loop in the reference, but both are pure functions of the input and an error discards the Shape)
    std::vector<Descriptor48> descs(n_groups);
    for (size_t i = 0; i < n_groups; ++i) CR_TRY(convert_dynamic_stroke_options(groups[i], descs[i]));

    // A pass in flight may still have to be re-submitted from the arrays of its batches (settle): in place they must not be
    // overwritten before that is known; with frame pipelining this build goes to the batch's OTHER set of arrays instead, and
    // the tessellation stream only waits (on the device) for the last pass that read that set.
    if (!r->pipelined) CR_TRY(settle(r));
    if (b->mirrors_pending) { CR_CUDA_TRY(cudaEventSynchronize(b->ev_mirrors)); b->mirrors_pending = false; }

    const bool optimistic = b->built && b->n_paths == n_paths && b->n_shapes == n_shapes && b->n_segments == soa->n_segments && b->n_groups == n_groups;
    const int target = (r->pipelined && b->built) ? (b->cur + 1) % CR_BATCH_SETS : b->cur;
    BatchStorage& B = b->set[target];
    if (B.was_read) CR_CUDA_TRY(cudaStreamWaitEvent(st, B.last_read, 0));
    if (r->pipelined && B.written) CR_CUDA_TRY(cudaStreamWaitEvent(st, B.written, 0));   // the set's previous build (its hulls run on another stream) is complete
    b->built = false;
    const bool host_inputs = soa->memory_space == CR_MEM_HOST;
    if (r->pipelined && host_inputs) {   // the other staging set; the copies wait for the kernels that read its previous contents
        r->staging_cur ^= 1;
        if (r->staging_free[r->staging_cur]) CR_CUDA_TRY(cudaStreamWaitEvent(r->copy, r->staging_free[r->staging_cur], 0));
    }

    if (r->timing) CR_CUDA_TRY(cudaEventRecord(r->ev[0], st));
    DevicePaths P{};
    P.n_paths = n_paths;
    P.n_segments = soa->n_segments;
    // per-type segment totals: from the host table, or (device inputs) from the previous build of the same shape until the
    // read-back below confirms them — they size the staging copies (host inputs only) and pick the kernel variant
    uint32_t type_totals[5] = {0, 0, 0, 0, 0};
    bool totals_known = n_paths == 0;
    if (n_paths && soa->memory_space == CR_MEM_HOST) {
        for (int t = 0; t < 5; ++t) type_totals[t] = soa->type_begin[t * stride + n_paths];
        totals_known = true;
        if ((uint64_t)type_totals[0] + type_totals[1] + type_totals[2] + type_totals[3] + type_totals[4] != soa->n_segments)
            return fail(CR_ERR_INVALID_ARGUMENT, "cr_path_soa: the per-type totals type_begin[t][n_paths] do not add up to n_segments");
    }
    static const int kSegFloats[5] = {2, 4, 6, 5, 10};
    CR_TRY(stage(r, 0, soa->start, 2 * (size_t)n_paths, soa->memory_space, &P.start));
    CR_TRY(stage(r, 1, soa->segment_begin, n_paths ? stride : 0, soa->memory_space, &P.segment_begin));
    CR_TRY(stage(r, 2, soa->segment_types, soa->n_segments, soa->memory_space, &P.segment_types));
    CR_TRY(stage(r, 3, soa->type_begin, n_paths ? 5 * stride : 0, soa->memory_space, &P.type_begin));
    const float* seg_src[5] = {soa->line_segments, soa->integral_quadratic, soa->integral_cubic, soa->rational_quadratic, soa->rational_cubic};
    for (int t = 0; t < 5; ++t) {
        if (soa->memory_space == CR_MEM_DEVICE) P.seg[t] = seg_src[t];
        else CR_TRY(stage(r, 4 + t, seg_src[t], (size_t)type_totals[t] * kSegFloats[t], soa->memory_space, &P.seg[t]));
    }
    // stroke_options == NULL: every Path has `stroke_options: None` (src/path.rs:215), i.e. all paths are filled
    if (soa->stroke_options) CR_TRY(stage(r, 9, soa->stroke_options, n_paths, soa->memory_space, &P.stroke_options));
    else P.stroke_options = nullptr;
    if (r->pipelined && host_inputs) {   // the tessellation kernels wait for the copies
        CR_CUDA_TRY(cudaEventRecord(r->ev_copied, r->copy));
        CR_CUDA_TRY(cudaStreamWaitEvent(st, r->ev_copied, 0));
    }

    // ---- pass A: count, scan, per-shape slice boundaries
    CR_TRY(r->counts.reserve(st, CNT_COUNT * stride * sizeof(uint32_t)));
    CR_TRY(reserve_scan_scratch(r, st, r->scan_scratch_tess, cr_scan_scratch_words((uint32_t)stride, CNT_COUNT)));
    CR_TRY(B.err.reserve(st, 8));
    CR_CUDA_TRY(cudaMemsetAsync(B.err.p, 0, 8, st));
    CR_TRY(r->shape_begin_dev.reserve(st, (size_t)(n_shapes + 1) * 4));
    CR_CUDA_TRY(cudaMemcpyAsync(r->shape_begin_dev.p, shape_path_begin, (size_t)(n_shapes + 1) * 4, cudaMemcpyHostToDevice, st));
    CR_TRY(B.cat_begin.reserve(st, (size_t)CNT_COUNT * (n_shapes + 1) * 4));
    CR_TRY(B.hull_count.reserve(st, (size_t)n_shapes * 4));
    uint32_t* counts = r->counts.as<uint32_t>();
    uint32_t* err = B.err.as<uint32_t>();

    auto run_sizes = [&](bool has_cubics, const TessCapacity& caps) -> int {
        CR_TRY(cr_tess_count(st, P, (uint32_t)n_groups, counts, err, has_cubics));
        CR_TRY(cr_scan_exclusive(st, counts, (uint32_t)stride, CNT_COUNT, r->scan_scratch_tess.as<uint32_t>()));
        CR_TRY(cr_tess_shape_bounds(st, counts, n_paths, r->shape_begin_dev.as<uint32_t>(), n_shapes, B.cat_begin.as<uint32_t>(), err + 1, caps, err));
        for (int c = 0; c < CNT_COUNT; ++c)
            CR_CUDA_TRY(cudaMemcpyAsync(&r->pinned[PIN_TESS + c], counts + c * stride + n_paths, 4, cudaMemcpyDeviceToHost, st));
        CR_CUDA_TRY(cudaMemcpyAsync(&r->pinned[PIN_TESS + CNT_COUNT], err, 8, cudaMemcpyDeviceToHost, st));
        if (n_paths && soa->memory_space == CR_MEM_DEVICE)
            for (int t = 0; t < 5; ++t)
                CR_CUDA_TRY(cudaMemcpyAsync(&r->pinned[PIN_TESS + CNT_COUNT + 2 + t], soa->type_begin + t * stride + n_paths, 4, cudaMemcpyDeviceToHost, st));
        CR_CUDA_TRY(cudaEventRecord(r->ev_sizes, st));
        return CR_OK;
    };
    auto reserve_outputs = [&](const uint64_t* totals) -> int {
        for (int c = 0; c < 7; ++c) CR_TRY(B.vtx[c].reserve(st, totals[c] * (size_t)kCategoryStride[c]));
        CR_TRY(B.proto.reserve(st, totals[CNT_PROTO] * 8));
        CR_TRY(B.hull.reserve(st, totals[CNT_PROTO] * 8));
        CR_TRY(r->hull_scratch_a.reserve(st, totals[CNT_PROTO] * 8));
        CR_TRY(r->hull_scratch_b.reserve(st, totals[CNT_PROTO] * 8));
        for (int k = 0; k < 3; ++k) CR_TRY(B.idx[k].reserve(st, totals[CNT_LINE_IDX + k] * 4));
        return CR_OK;
    };
    bool emit_has_cubics = true;   // false: the batch is known to hold no cubic segment (every filled path goes through the per-segment kernel)
    auto run_emit = [&](uint32_t max_proto) -> int {
        TessOutput out{};
        for (int c = 0; c < 7; ++c) out.vtx[c] = B.vtx[c].p;
        out.proto = B.proto.as<float2>();
        for (int k = 0; k < 3; ++k) out.idx[k] = B.idx[k].as<uint32_t>();
        CR_TRY(cr_tess_emit(st, P, counts, r->shape_begin_dev.as<uint32_t>(), n_shapes, out, err, emit_has_cubics));
        if (r->pipelined) {   // the input (staging) arrays are free again; the hulls continue on their own stream
            if (host_inputs) CR_CUDA_TRY(cudaEventRecord(r->staging_free[r->staging_cur], st));
            CR_CUDA_TRY(cudaEventRecord(r->ev_emitted, st));
            CR_CUDA_TRY(cudaStreamWaitEvent(hs, r->ev_emitted, 0));
        }
        if (r->timing) CR_CUDA_TRY(cudaEventRecord(r->ev[6], hs));
        CR_TRY(cr_tess_hull(hs, out.proto, r->hull_scratch_a.as<float2>(), r->hull_scratch_b.as<float2>(),
                            B.cat_begin.as<uint32_t>() + (size_t)CNT_PROTO * (n_shapes + 1), n_shapes, B.hull.as<float2>(), B.hull_count.as<uint32_t>(), max_proto, err,
                            r->timing ? r->ev[7] : nullptr));
        if (r->timing) { CR_CUDA_TRY(cudaEventRecord(r->ev[8], hs)); CR_CUDA_TRY(cudaEventRecord(r->ev[1], hs)); r->ev_valid[0] = true; }
        return CR_OK;
    };
    TessCapacity unlimited;
    for (int c = 0; c < CNT_COUNT; ++c) unlimited.v[c] = 0xFFFFFFFFu;

    bool emitted = false;
    if (optimistic) {
        TessCapacity caps;
        for (int c = 0; c < 7; ++c) caps.v[c] = (uint32_t)std::min<size_t>(0xFFFFFFFFu, B.vtx[c].cap / kCategoryStride[c]);
        caps.v[CNT_PROTO] = (uint32_t)std::min<size_t>(0xFFFFFFFFu, std::min(std::min(B.proto.cap, B.hull.cap), std::min(r->hull_scratch_a.cap, r->hull_scratch_b.cap)) / 8);
        for (int k = 0; k < 3; ++k) caps.v[CNT_LINE_IDX + k] = (uint32_t)std::min<size_t>(0xFFFFFFFFu, B.idx[k].cap / 4);
        emit_has_cubics = totals_known ? (type_totals[CR_SEG_INTEGRAL_CUBIC] != 0 || type_totals[CR_SEG_RATIONAL_CUBIC] != 0) : b->has_cubics;
        CR_TRY(run_sizes(emit_has_cubics, caps));
        CR_TRY(run_emit(b->max_proto));   // the sort's shared-memory capacity is a launch parameter: larger shapes take its global-memory path
        emitted = true;
    } else {
        // device inputs: the kernel variant without the cubic builder may only be used once the totals are known
        CR_TRY(run_sizes(totals_known ? (type_totals[CR_SEG_INTEGRAL_CUBIC] != 0 || type_totals[CR_SEG_RATIONAL_CUBIC] != 0) : true, unlimited));
    }
    CR_CUDA_TRY(cudaEventSynchronize(r->ev_sizes));
    uint32_t flags = r->pinned[PIN_TESS + CNT_COUNT];
    if (n_paths && soa->memory_space == CR_MEM_DEVICE) {
        for (int t = 0; t < 5; ++t) type_totals[t] = r->pinned[PIN_TESS + CNT_COUNT + 2 + t];
        if ((uint64_t)type_totals[0] + type_totals[1] + type_totals[2] + type_totals[3] + type_totals[4] != soa->n_segments)
            return fail(CR_ERR_INVALID_ARGUMENT, "cr_path_soa: the per-type totals type_begin[t][n_paths] do not add up to n_segments");
    }
    const bool has_cubics = type_totals[CR_SEG_INTEGRAL_CUBIC] != 0 || type_totals[CR_SEG_RATIONAL_CUBIC] != 0;
    if (emitted && (flags & (CR_DEVERR_CAPACITY | CR_DEVERR_MODE))) {
        // the optimistic launch did nothing (emit and hull return when they see these bits): redo in the cold order
        CR_CUDA_TRY(cudaStreamSynchronize(st));
        if (r->pipelined) CR_CUDA_TRY(cudaStreamSynchronize(hs));
        CR_CUDA_TRY(cudaMemsetAsync(B.err.p, 0, 8, st));
        CR_TRY(run_sizes(has_cubics, unlimited));
        CR_CUDA_TRY(cudaEventSynchronize(r->ev_sizes));
        flags = r->pinned[PIN_TESS + CNT_COUNT];
        emitted = false;
    }
    CR_TRY(decode_device_error(flags));
    for (int c = 0; c < CNT_COUNT; ++c) b->totals[c] = r->pinned[PIN_TESS + c];
    const uint32_t max_proto = r->pinned[PIN_TESS + CNT_COUNT + 1];

    b->renderer = r;
    b->n_shapes = n_shapes;
    b->n_paths = n_paths;
    b->n_groups = (uint32_t)n_groups;
    b->n_segments = soa->n_segments;
    b->has_cubics = has_cubics;
    b->max_proto = max_proto;
    if (!emitted) {
        emit_has_cubics = has_cubics;
        if (r->pipelined) {   // the hull scratch arrays may be re-allocated here (on the tessellation stream): the hulls of the previous build must be done with them
            CR_CUDA_TRY(cudaEventRecord(r->ev_hull_idle, hs));
            CR_CUDA_TRY(cudaStreamWaitEvent(st, r->ev_hull_idle, 0));
        }
        CR_TRY(reserve_outputs(b->totals));
        CR_TRY(run_emit(max_proto));
    }
    // everything below follows the hulls: it is enqueued on their stream, which is the build's last stage
    CR_TRY(B.stroke.reserve(st, n_groups * sizeof(Descriptor48)));
    CR_TRY(B.desc_dev.reserve(st, sizeof(DeviceBatch)));
    if (r->pipelined && !emitted) { CR_CUDA_TRY(cudaEventRecord(r->ev_emitted, st)); CR_CUDA_TRY(cudaStreamWaitEvent(hs, r->ev_emitted, 0)); }   // (re-)allocations above precede their use
    if (n_groups) CR_CUDA_TRY(cudaMemcpyAsync(B.stroke.p, descs.data(), n_groups * sizeof(Descriptor48), cudaMemcpyHostToDevice, hs));

    // ---- the rasteriser's view of this batch + host mirrors of the slice tables (asynchronous, pinned)
    DeviceBatch db{};
    for (int c = 0; c < 7; ++c) db.vtx[c] = B.vtx[c].p;
    db.hull = B.hull.as<float2>();
    for (int k = 0; k < 3; ++k) db.idx[k] = B.idx[k].as<uint32_t>();
    db.cat_begin = B.cat_begin.as<uint32_t>();
    db.hull_count = B.hull_count.as<uint32_t>();
    db.stroke = B.stroke.p;
    db.n_shapes = n_shapes;
    db.n_groups = (uint32_t)n_groups;
    CR_CUDA_TRY(cudaMemcpyAsync(B.desc_dev.p, &db, sizeof(db), cudaMemcpyHostToDevice, hs));
    const size_t table_words = (size_t)CNT_COUNT * (n_shapes + 1), mirror_words = table_words + n_shapes + 2;
    if (mirror_words > b->mirrors_cap) {
        if (b->mirrors) cudaFreeHost(b->mirrors);
        b->mirrors = nullptr;
        CR_CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&b->mirrors), mirror_words * 4, cudaHostAllocDefault));
        b->mirrors_cap = mirror_words;
    }
    if (!b->ev_mirrors) CR_CUDA_TRY(cudaEventCreateWithFlags(&b->ev_mirrors, cudaEventDisableTiming));
    CR_CUDA_TRY(cudaMemcpyAsync(b->mirrors, B.cat_begin.p, table_words * 4, cudaMemcpyDeviceToHost, hs));
    CR_CUDA_TRY(cudaMemcpyAsync(b->mirrors + table_words, B.hull_count.p, (size_t)n_shapes * 4, cudaMemcpyDeviceToHost, hs));
    CR_CUDA_TRY(cudaMemcpyAsync(b->mirrors + table_words + n_shapes, err, 4, cudaMemcpyDeviceToHost, hs));
    CR_CUDA_TRY(cudaEventRecord(b->ev_mirrors, hs));
    b->mirrors_pending = true;
    if (!B.written) CR_CUDA_TRY(cudaEventCreateWithFlags(&B.written, cudaEventDisableTiming));
    CR_CUDA_TRY(cudaEventRecord(B.written, hs));
    b->cur = target;
    b->views.resize(n_shapes);
    for (uint32_t s = 0; s < n_shapes; ++s) b->views[s] = cr_shape{b, s, false};
    b->built = true;
    if (!optimistic) CR_TRY(ensure_mirrors(b));   // a first build reports every error of the emit pass (non-finite vertices) before it returns

    uint64_t input_bytes = 8ull * n_paths;
    for (int t = 0; t < 5; ++t) input_bytes += (uint64_t)type_totals[t] * (1 + 4 * kSegFloats[t]);
    uint64_t out_bytes = 0;
    for (int c = 0; c < 7; ++c) out_bytes += b->totals[c] * (uint64_t)kCategoryStride[c];
    for (int k = 0; k < 3; ++k) out_bytes += 2ull * b->totals[CNT_LINE_IDX + k];
    // B_in of SURVEY §8d: 8 + 24 [stroked] per path; the stroked count is not known on the host for device inputs,
    // so the 24 B record is counted for every path that carries one (all of them in this ABI).
    r->stats.input_bytes = input_bytes + (soa->stroke_options ? 24ull * n_paths : 0ull);
    r->stats.vertex_bytes = out_bytes;   // + 8 B per hull vertex, added by cr_renderer_get_stats once the hull counts have arrived
    r->stats.tessellated_paths = n_paths;
    r->stats.proto_hull_points = b->totals[CNT_PROTO];
    r->stats.hull_vertices = 0;
    r->stats_batch = b;
    return CR_OK;
}

}  // namespace

// ============================================================================================= exported C-ABI
extern "C" {

uint32_t cr_abi_version(void) { return 3; }
const char* cr_last_error_message(void) { return g_error_message; }
const char* cr_status_string(int status) {
    switch (status) {
        case CR_OK: return "Ok";
        case CR_ERR_NUMBER_OF_STENCIL_BITS_IS_UNSUPPORTED: return "NumberOfStencilBitsIsUnsupported";
        case CR_ERR_CLIP_STACK_OVERFLOW: return "ClipStackOverflow";
        case CR_ERR_TOO_MANY_NESTED_OPACITY_GROUPS: return "TooManyNestedOpacityGroups";
        case CR_ERR_TOO_MANY_DASH_INTERVALS: return "TooManyDashIntervals";
        case CR_ERR_DYNAMIC_STROKE_OPTIONS_INDEX_OUT_OF_BOUNDS: return "DynamicStrokeOptionsIndexOutOfBounds";
        case CR_ERR_INVALID_ARGUMENT: return "InvalidArgument";
        case CR_ERR_CUDA: return "Cuda";
        case CR_ERR_NON_FINITE: return "NonFinite";
        case CR_ERR_CURVE_STEPS_CAPACITY: return "CurveStepsCapacity";
        case CR_ERR_CUBIC_TRIANGULATION: return "CubicTriangulation";
        case CR_ERR_NO_DEVICE: return "NoDevice";
        case CR_ERR_NOT_RESIZED: return "NotResized";
        default: return "Unknown";
    }
}

// Renderer::new (src/renderer.rs:432-437)
int cr_renderer_create(const cr_config* config, cr_renderer** out) {
    if (!config || !out) return fail(CR_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    if (config->winding_counter_bits == 0 || config->clip_nesting_counter_bits + config->winding_counter_bits > 8)
        return fail(CR_ERR_NUMBER_OF_STENCIL_BITS_IS_UNSUPPORTED, "winding_counter_bits = %u, clip_nesting_counter_bits = %u", config->winding_counter_bits,
                    config->clip_nesting_counter_bits);
    if (config->msaa_sample_count != 1 && config->msaa_sample_count != 4) return fail(CR_ERR_INVALID_ARGUMENT, "msaa_sample_count must be 1 or 4");
    if (config->blending > CR_BLEND_REPLACE || config->cull_mode > CR_CULL_BACK) return fail(CR_ERR_INVALID_ARGUMENT, "bad blending / cull_mode");
    if (config->depth_compare > CR_COMPARE_ALWAYS || config->depth_write_enabled > 1u) return fail(CR_ERR_INVALID_ARGUMENT, "bad depth_compare / depth_write_enabled");
    if (config->color_format > CR_FORMAT_BGRA8_UNORM) return fail(CR_ERR_INVALID_ARGUMENT, "bad color_format");
    int n_devices = 0;
    if (cudaGetDeviceCount(&n_devices) != cudaSuccess || n_devices == 0) {
        cudaGetLastError();
        return fail(CR_ERR_NO_DEVICE, "no CUDA device: libcontrast_b200 has no CPU path");
    }
    int device = config->device;
    if (device < 0) CR_CUDA_TRY(cudaGetDevice(&device));
    if (device >= n_devices) return fail(CR_ERR_NO_DEVICE, "device %d of %d", device, n_devices);
    std::unique_ptr<cr_renderer> r(new (std::nothrow) cr_renderer());
    if (!r) return fail(CR_ERR_INVALID_ARGUMENT, "out of host memory");
    r->config = *config;
    r->config.device = device;
    r->device = device;
    DeviceGuard guard(device);
    if (!guard.ok) return fail(CR_ERR_CUDA, "cannot select CUDA device %d", device);
    CR_CUDA_TRY(cudaStreamCreateWithFlags(&r->own_stream, cudaStreamNonBlocking));
    r->stream = r->own_stream;
    cudaMemPool_t pool;
    CR_CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t threshold = ~0ull;   // keep freed blocks cached: shape rebuilds every frame must not hit cudaMalloc
    CR_CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold));
    CR_CUDA_TRY(cudaMallocHost(reinterpret_cast<void**>(&r->pinned), PIN_WORDS * sizeof(uint32_t)));
    memset(r->pinned, 0, PIN_WORDS * sizeof(uint32_t));
    for (auto& e : r->ev) CR_CUDA_TRY(cudaEventCreate(&e));
    CR_CUDA_TRY(cudaEventCreateWithFlags(&r->ev_sizes, cudaEventDisableTiming));
    CR_CUDA_TRY(cudaEventCreateWithFlags(&r->ev_pass, cudaEventDisableTiming));
    *out = r.release();
    return CR_OK;
}

static void close_peers(cr_renderer* r) {
    for (int i = 0; i < CR_MAX_PEERS; ++i) {
        if (r->peer_color[i]) cudaIpcCloseMemHandle(r->peer_color[i]);
        if (r->peer_stencil[i]) cudaIpcCloseMemHandle(r->peer_stencil[i]);
        if (r->peer_exchange[i]) cudaIpcCloseMemHandle(r->peer_exchange[i]);
        r->peer_color[i] = r->peer_stencil[i] = r->peer_exchange[i] = nullptr;
    }
}
static void pass_free(cr_pass* p);
static void renderer_free(cr_renderer* r) {
    DeviceGuard guard(r->device);
    cudaStreamSynchronize(r->stream);
    for (cudaStream_t q : {r->tess, r->copy, r->hull}) if (q) cudaStreamSynchronize(q);
    if (r->inflight) { pass_free(r->inflight); r->inflight = nullptr; }
    close_peers(r);
    cudaStream_t st = r->stream;
    DevBuf* all[] = {&r->color, &r->stencil, &r->alpha_layers, &r->depth, &r->exchange, &r->counts, &r->scan_scratch, &r->shape_begin_dev, &r->hull_scratch_a,
                     &r->hull_scratch_b, &r->scan_scratch_tess, &r->compact_dev, &r->cmds_dev, &r->batches_dev, &r->cmd_cands, &r->cand_tiles, &r->records, &r->big_list, &r->pair_tile, &r->pair_cand,
                     &r->pair_tile_alt, &r->pair_cand_alt, &r->tile_prims, &r->clip_list, &r->clip_attrs, &r->radix_scratch, &r->tile_begin, &r->inst_transforms, &r->inst_colors, &r->pass_counters};
    for (DevBuf* d : all) d->release(st);
    for (auto& set : r->staging) for (auto& d : set) d.release(st);
    if (r->out) cudaStreamSynchronize(r->out);
    for (auto& rb : r->readback) { rb.buf.release(st); if (rb.snapshot) cudaEventDestroy(rb.snapshot); if (rb.done) cudaEventDestroy(rb.done); }
    cudaStreamSynchronize(st);
    for (auto& e : r->ev) if (e) cudaEventDestroy(e);
    if (r->pinned) cudaFreeHost(r->pinned);
    if (r->cmd_arena) cudaFreeHost(r->cmd_arena);
    if (r->ev_sizes) cudaEventDestroy(r->ev_sizes);
    if (r->ev_pass) cudaEventDestroy(r->ev_pass);
    if (r->ev_update) cudaEventDestroy(r->ev_update);
    for (cudaEvent_t e : {r->staging_free[0], r->staging_free[1], r->ev_copied, r->ev_emitted, r->ev_hull_idle}) if (e) cudaEventDestroy(e);
    for (cudaStream_t q : {r->tess, r->copy, r->hull, r->out}) if (q) cudaStreamDestroy(q);
    if (r->own_stream) cudaStreamDestroy(r->own_stream);
    delete r;
}
// Shapes and passes hold a pointer to their renderer (Rust: a borrow). Destroying the renderer first is tolerated: it is
// kept alive until its last child is gone.
static void renderer_release_child(cr_renderer* r) {
    if (r->live_objects > 0) --r->live_objects;
    if (r->destroy_requested && r->live_objects == 0) renderer_free(r);
}
void cr_renderer_destroy(cr_renderer* r) {
    if (!r) return;
    { DeviceGuard guard(r->device); settle(r); }   // the pass in flight is one of the live objects
    if (r->live_objects > 0) { r->destroy_requested = true; return; }
    renderer_free(r);
}

int cr_renderer_get_config(const cr_renderer* r, cr_config* out) {
    if (!r || !out) return fail(CR_ERR_INVALID_ARGUMENT, "null argument");
    *out = r->config;
    return CR_OK;
}

static bool has_depth(const cr_renderer* r) {
    const uint32_t f = r->config.depth_compare;
    return (f != CR_COMPARE_DEFAULT_ALWAYS && f != CR_COMPARE_ALWAYS) || r->config.depth_write_enabled != 0u;
}
static size_t color_texel_bytes(const cr_renderer* r) { return r->config.color_format == CR_FORMAT_RGBA32F ? 16 : 4; }
// Renderer::resize_internal_buffers (src/renderer.rs:892-929) + the caller-owned colour / depth-stencil textures.
int cr_renderer_resize(cr_renderer* r, uint32_t width, uint32_t height) {
    if (!r) return fail(CR_ERR_INVALID_ARGUMENT, "null renderer");
    if (width == 0 || height == 0 || width > 32768 || height > 32768) return fail(CR_ERR_INVALID_ARGUMENT, "bad extent %ux%u", width, height);
    CR_GUARD(r);
    CR_TRY(settle(r));
    if (r->out) CR_CUDA_TRY(cudaStreamSynchronize(r->out));   // read-backs of the old extent
    for (auto& rb : r->readback) rb.pending = false;
    r->cand_cap = r->pair_cap = 0;   // sized for another extent
    const size_t samples = (size_t)width * height * r->config.msaa_sample_count;
    if (r->order_world > 1) return fail(CR_ERR_INVALID_ARGUMENT, "resize of an order-sharded target: call cr_renderer_set_order_sharding(r, 1, 0) first");
    if (r->peer_color[0] || r->peer_stencil[0]) return fail(CR_ERR_INVALID_ARGUMENT, "resize while peer attachments are imported: call cr_renderer_set_tile_sharding(r, 1, 0) first");
    r->color.plain = r->stencil.plain = true;   // exportable to the other ranks of a tile-sharded target
    const size_t texel = color_texel_bytes(r), layer_texel = r->config.color_format == CR_FORMAT_RGBA32F ? 4 : 1;
    CR_TRY(r->color.reserve(r->stream, samples * texel));
    CR_TRY(r->stencil.reserve(r->stream, samples));
    CR_TRY(r->alpha_layers.reserve(r->stream, samples * layer_texel * std::max<uint32_t>(1, r->config.alpha_layer_count)));
    if (has_depth(r)) {   // wgpu clears depth to 1.0 in the pass (LoadOp::Clear); a fresh attachment starts there too
        CR_TRY(r->depth.reserve(r->stream, samples * 4));
        std::vector<float> ones(samples, 1.0f);
        CR_CUDA_TRY(cudaMemcpyAsync(r->depth.p, ones.data(), samples * 4, cudaMemcpyHostToDevice, r->stream));
        CR_CUDA_TRY(cudaStreamSynchronize(r->stream));
    }
    r->width = width;
    r->height = height;
    r->tiles_x = (width + CR_TILE - 1) / CR_TILE;
    r->tiles_y = (height + CR_TILE - 1) / CR_TILE;
    CR_CUDA_TRY(cudaMemsetAsync(r->color.p, 0, samples * texel, r->stream));
    CR_CUDA_TRY(cudaMemsetAsync(r->stencil.p, 0, samples, r->stream));
    CR_CUDA_TRY(cudaMemsetAsync(r->alpha_layers.p, 0, samples * layer_texel * std::max<uint32_t>(1, r->config.alpha_layer_count), r->stream));
    return CR_OK;
}

int cr_renderer_set_stream(cr_renderer* r, void* cuda_stream) {
    if (!r) return fail(CR_ERR_INVALID_ARGUMENT, "null renderer");
    CR_GUARD(r);
    CR_TRY(settle(r));
    CR_CUDA_TRY(cudaStreamSynchronize(r->stream));
    for (cudaStream_t q : {r->tess, r->copy, r->hull}) if (q) CR_CUDA_TRY(cudaStreamSynchronize(q));
    r->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : r->own_stream;
    return CR_OK;
}
int cr_renderer_synchronize(cr_renderer* r) {
    if (!r) return fail(CR_ERR_INVALID_ARGUMENT, "null renderer");
    CR_GUARD(r);
    CR_TRY(settle(r));
    CR_CUDA_TRY(cudaStreamSynchronize(r->stream));
    if (r->pipelined) for (cudaStream_t q : {r->copy, r->tess, r->hull}) CR_CUDA_TRY(cudaStreamSynchronize(q));
    return take_deferred(r);
}

// Frame pipelining: Shape::from_paths work moves to a second, high-priority stream of the renderer and every rebuild of a
// batch (`existing`) writes the set of arrays the last pass is not reading, so that tessellating frame N + 1 overlaps
// rasterising frame N. Passes wait (on the device) for the builds of the batches they render; everything the caller observes
// through this API is ordered as before. What changes for the caller: input arrays in DEVICE memory must be complete when
// cr_shape_from_paths / cr_shape_batch_from_paths is called — work merely enqueued on the renderer's stream is not waited for.
int cr_renderer_set_pipelining(cr_renderer* r, uint32_t enabled) {
    if (!r) return fail(CR_ERR_INVALID_ARGUMENT, "null renderer");
    CR_GUARD(r);
    CR_TRY(settle(r));
    CR_CUDA_TRY(cudaStreamSynchronize(r->stream));
    for (cudaStream_t q : {r->tess, r->copy, r->hull}) if (q) CR_CUDA_TRY(cudaStreamSynchronize(q));
    if (enabled && !r->tess) {
        int least = 0, greatest = 0;
        CR_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        CR_CUDA_TRY(cudaStreamCreateWithPriority(&r->tess, cudaStreamNonBlocking, greatest));
        CR_CUDA_TRY(cudaStreamCreateWithPriority(&r->hull, cudaStreamNonBlocking, greatest));
        CR_CUDA_TRY(cudaStreamCreateWithPriority(&r->copy, cudaStreamNonBlocking, greatest));
        for (cudaEvent_t* e : {&r->staging_free[0], &r->staging_free[1], &r->ev_copied, &r->ev_emitted, &r->ev_hull_idle}) CR_CUDA_TRY(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    }
    r->pipelined = enabled != 0;
    return CR_OK;
}

// ---------------------------------------------------------------------------------------------- shape building
int cr_shape_batch_from_paths(cr_renderer* r, const cr_dynamic_stroke_options* groups, size_t n_groups, const cr_path_soa* paths,
                              const uint32_t* shape_path_begin, uint32_t n_shapes, cr_shape_batch* existing, cr_shape_batch** out) {
    if (!r || !out) return fail(CR_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    CR_GUARD(r);
    if (!r->pipelined) {   // a pass in flight may still have to be re-submitted from the arrays this call is about to overwrite
        CR_TRY(settle(r));
        CR_TRY(take_deferred(r));
    }
    cr_shape_batch* b = existing;   // consumed: its allocations are reused in place when large enough (Buffer::update, src/renderer.rs:89-95)
    if (b && b->renderer != r) return fail(CR_ERR_INVALID_ARGUMENT, "existing batch belongs to another renderer");
    if (!b) {
        b = new (std::nothrow) cr_shape_batch();
        if (!b) return fail(CR_ERR_INVALID_ARGUMENT, "out of host memory");
        b->renderer = r;
        ++r->live_objects;
    }
    const int st = build_batch(r, groups, n_groups, paths, shape_path_begin, n_shapes, b);
    if (st != CR_OK) {
        char message[sizeof(g_error_message)];
        memcpy(message, g_error_message, sizeof(message));
        settle(r);   // the batch is about to be released: a pass in flight that renders it must be complete
        memcpy(g_error_message, message, sizeof(message));
        if (r->stats_batch == b) r->stats_batch = nullptr;
        batch_release(b);
        delete b;
        renderer_release_child(r);
        return st;
    }
    *out = b;
    return CR_OK;
}
void cr_shape_batch_destroy(cr_shape_batch* b) {
    if (!b) return;
    cr_renderer* r = b->renderer;
    {
        DeviceGuard guard(r->device);
        settle(r);   // a pass in flight may reference this batch
        if (r->stats_batch == b) r->stats_batch = nullptr;
        batch_release(b);
    }
    delete b;
    renderer_release_child(r);
}
uint32_t cr_shape_batch_size(const cr_shape_batch* b) { return b ? b->n_shapes : 0; }
cr_shape* cr_shape_batch_get(cr_shape_batch* b, uint32_t index) { return (b && index < b->n_shapes) ? &b->views[index] : nullptr; }

// Shape::from_paths (src/renderer.rs:177-249)
int cr_shape_from_paths(cr_renderer* r, const cr_dynamic_stroke_options* groups, size_t n_groups, const cr_path_soa* paths, cr_shape* existing,
                        cr_shape** out) {
    if (!r || !out || !paths) return fail(CR_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    if (existing && !existing->owns_batch) return fail(CR_ERR_INVALID_ARGUMENT, "existing is a borrowed view of a batch");
    const uint32_t begin[2] = {0, paths->n_paths};
    cr_shape_batch* old_batch = existing ? existing->batch : nullptr;
    cr_shape_batch* batch = nullptr;
    const int st = cr_shape_batch_from_paths(r, groups, n_groups, paths, begin, 1, old_batch, &batch);
    if (existing) delete existing;   // consumed either way; on failure its batch was released above
    if (st != CR_OK) return st;
    cr_shape* s = new (std::nothrow) cr_shape{batch, 0, true};
    if (!s) { cr_shape_batch_destroy(batch); return fail(CR_ERR_INVALID_ARGUMENT, "out of host memory"); }
    *out = s;
    return CR_OK;
}
void cr_shape_destroy(cr_shape* s) {
    if (!s || !s->owns_batch) return;
    cr_shape_batch_destroy(s->batch);
    delete s;
}

// Shape::set_dynamic_stroke_options (src/renderer.rs:360-376)
int cr_shape_batch_set_dynamic_stroke_options(cr_shape_batch* b, size_t index, const cr_dynamic_stroke_options* options) {
    if (!b || !options) return fail(CR_ERR_INVALID_ARGUMENT, "null argument");
    if (index >= b->n_groups) return fail(CR_ERR_DYNAMIC_STROKE_OPTIONS_INDEX_OUT_OF_BOUNDS, "group %zu of %u", index, b->n_groups);
    Descriptor48 d;
    CR_TRY(convert_dynamic_stroke_options(*options, d));
    cr_renderer* r = b->renderer;
    CR_GUARD(r);
    CR_TRY(settle(r));   // a pass in flight was recorded with the old descriptor
    // staged through a slot of the pinned area (one 48-byte asynchronous copy, Buffer::update of src/renderer.rs:89-95); the slot
    // is reused by the next update, which therefore waits for this copy — but not for the stream
    if (!r->ev_update) CR_CUDA_TRY(cudaEventCreateWithFlags(&r->ev_update, cudaEventDisableTiming));
    else CR_CUDA_TRY(cudaEventSynchronize(r->ev_update));
    memcpy(&r->pinned[PIN_MISC], &d, sizeof(d));
    BatchStorage& B = b->store();
    if (B.written) CR_CUDA_TRY(cudaStreamWaitEvent(r->stream, B.written, 0));   // ordered behind the build and between the passes on the renderer's stream
    CR_CUDA_TRY(cudaMemcpyAsync(static_cast<char*>(B.stroke.p) + index * sizeof(d), &r->pinned[PIN_MISC], sizeof(d), cudaMemcpyHostToDevice, r->stream));
    CR_CUDA_TRY(cudaEventRecord(r->ev_update, r->stream));
    return CR_OK;
}
int cr_shape_set_dynamic_stroke_options(cr_shape* s, size_t index, const cr_dynamic_stroke_options* options) {
    if (!s) return fail(CR_ERR_INVALID_ARGUMENT, "null shape");
    return cr_shape_batch_set_dynamic_stroke_options(s->batch, index, options);
}

int cr_shape_get_layout(cr_shape* s, cr_shape_layout* out) {
    if (!s || !out) return fail(CR_ERR_INVALID_ARGUMENT, "null argument");
    cr_shape_batch* b = s->batch;
    CR_TRY(ensure_mirrors(b));
    const size_t stride = (size_t)b->n_shapes + 1;
    const uint32_t* cb = b->cat_begin_host();
    uint64_t acc = 0;
    for (int c = 0; c < 7; ++c) {
        acc += (uint64_t)(cb[c * stride + s->index + 1] - cb[c * stride + s->index]) * kCategoryStride[c];
        out->vertex_offsets[c] = acc;
    }
    acc += 8ull * b->hull_count_host()[s->index];
    out->vertex_offsets[7] = acc;
    acc = 0;
    for (int k = 0; k < 3; ++k) {
        acc += 2ull * (cb[(CNT_LINE_IDX + k) * stride + s->index + 1] - cb[(CNT_LINE_IDX + k) * stride + s->index]);
        out->index_offsets[k] = acc;
    }
    out->dynamic_stroke_options_count = b->n_groups;
    out->proto_hull_points = cb[CNT_PROTO * stride + s->index + 1] - cb[CNT_PROTO * stride + s->index];
    return CR_OK;
}

int cr_shape_read_vertex_buffer(cr_shape* s, void* dst, size_t capacity) {
    if (!s || !dst) return fail(CR_ERR_INVALID_ARGUMENT, "null argument");
    cr_shape_layout layout;
    CR_TRY(cr_shape_get_layout(s, &layout));
    if (capacity < layout.vertex_offsets[7]) return fail(CR_ERR_INVALID_ARGUMENT, "capacity %zu < %llu", capacity, (unsigned long long)layout.vertex_offsets[7]);
    const cr_shape_batch* b = s->batch;
    CR_GUARD(b->renderer);
    cudaStream_t st = b->renderer->stream;
    const BatchStorage& B = b->store();
    if (B.written) CR_CUDA_TRY(cudaStreamWaitEvent(st, B.written, 0));
    const size_t stride = (size_t)b->n_shapes + 1;
    const uint32_t* cb = b->cat_begin_host();
    char* d = static_cast<char*>(dst);
    uint64_t at = 0;
    for (int c = 0; c < 8; ++c) {
        const uint64_t bytes = layout.vertex_offsets[c] - at;
        if (bytes) {
            const char* src = c < 7 ? static_cast<const char*>(B.vtx[c].p) + (size_t)cb[c * stride + s->index] * kCategoryStride[c]
                                    : static_cast<const char*>(B.hull.p) + (size_t)cb[CNT_PROTO * stride + s->index] * 8;
            CR_CUDA_TRY(cudaMemcpyAsync(d + at, src, bytes, cudaMemcpyDeviceToHost, st));
        }
        at = layout.vertex_offsets[c];
    }
    CR_CUDA_TRY(cudaStreamSynchronize(st));
    return CR_OK;
}

int cr_shape_read_index_buffer(cr_shape* s, void* dst, size_t capacity) {
    if (!s || !dst) return fail(CR_ERR_INVALID_ARGUMENT, "null argument");
    cr_shape_layout layout;
    CR_TRY(cr_shape_get_layout(s, &layout));
    if (capacity < layout.index_offsets[2]) return fail(CR_ERR_INVALID_ARGUMENT, "capacity too small");
    const cr_shape_batch* b = s->batch;
    CR_GUARD(b->renderer);
    cudaStream_t st = b->renderer->stream;
    const BatchStorage& B = b->store();
    if (B.written) CR_CUDA_TRY(cudaStreamWaitEvent(st, B.written, 0));
    const size_t stride = (size_t)b->n_shapes + 1;
    const uint32_t* cb = b->cat_begin_host();
    uint16_t* d = static_cast<uint16_t*>(dst);
    std::vector<uint32_t> wide;
    for (int k = 0; k < 3; ++k) {
        const uint32_t begin = cb[(CNT_LINE_IDX + k) * stride + s->index], n = cb[(CNT_LINE_IDX + k) * stride + s->index + 1] - begin;
        wide.resize(n);
        if (n) CR_CUDA_TRY(cudaMemcpyAsync(wide.data(), B.idx[k].as<uint32_t>() + begin, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        CR_CUDA_TRY(cudaStreamSynchronize(st));
        // `start_index as u16` (src/stroke.rs:108,128, src/fill.rs:363): the reference's u16 indices wrap modulo 65536
        for (uint32_t i = 0; i < n; ++i) *d++ = wide[i] == CR_RESTART ? (uint16_t)0xFFFF : (uint16_t)(wide[i] >> 1);
    }
    return CR_OK;
}

int cr_shape_read_stroke_buffer(cr_shape* s, void* dst, size_t capacity) {
    if (!s || !dst) return fail(CR_ERR_INVALID_ARGUMENT, "null argument");
    const cr_shape_batch* b = s->batch;
    const size_t bytes = (size_t)b->n_groups * sizeof(Descriptor48);
    if (capacity < bytes) return fail(CR_ERR_INVALID_ARGUMENT, "capacity too small");
    CR_GUARD(b->renderer);
    if (b->store().written) CR_CUDA_TRY(cudaStreamWaitEvent(b->renderer->stream, b->store().written, 0));
    if (bytes) CR_CUDA_TRY(cudaMemcpyAsync(dst, b->store().stroke.p, bytes, cudaMemcpyDeviceToHost, b->renderer->stream));
    CR_CUDA_TRY(cudaStreamSynchronize(b->renderer->stream));
    return CR_OK;
}

// ------------------------------------------------------------------------------------------------- render pass
static void pass_free(cr_pass* p) {
    cr_renderer* r = p->renderer;
    if (p->arena) r->cmd_arena_busy = false;
    delete p;
}

int cr_pass_begin(cr_renderer* r, uint32_t clear_color, uint32_t clear_stencil, cr_pass** out) {
    return cr_pass_begin_depth(r, clear_color, clear_stencil, clear_stencil, 1.0f, out);
}
int cr_pass_begin_depth(cr_renderer* r, uint32_t clear_color, uint32_t clear_stencil, uint32_t clear_depth, float depth_clear_value, cr_pass** out) {
    if (!r || !out) return fail(CR_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    if (r->width == 0) return fail(CR_ERR_NOT_RESIZED, "cr_renderer_resize has not been called");
    CR_GUARD(r);
    CR_TRY(settle(r));   // the previous pass gives the command arena back (its device-side sizes have arrived long ago)
    CR_TRY(take_deferred(r));
    cr_pass* p = new (std::nothrow) cr_pass();
    if (!p) return fail(CR_ERR_INVALID_ARGUMENT, "out of host memory");
    p->renderer = r;
    // Like a wgpu render pass, the clear belongs to the pass and happens when the pass executes (cr_pass_submit): the tile
    // kernel starts cleared tiles from zero instead of loading them and writes every tile, so the attachments are neither
    // memset nor read. A pass that is dropped without submit clears nothing.
    p->clear_color = clear_color != 0;
    p->clear_stencil = clear_stencil != 0;
    p->clear_depth = clear_depth != 0;
    p->depth_clear_value = depth_clear_value;
    if (!r->cmd_arena_busy) { p->arena = true; r->cmd_arena_busy = true; }
    ++r->live_objects;
    *out = p;
    return CR_OK;
}

int cr_pass_set_instances(cr_pass* p, const float* transforms, const float* colors, uint32_t count, uint32_t memory_space) {
    if (!p) return fail(CR_ERR_INVALID_ARGUMENT, "null pass");
    if (count && !transforms) return fail(CR_ERR_INVALID_ARGUMENT, "null transforms");
    if (memory_space > CR_MEM_DEVICE) return fail(CR_ERR_INVALID_ARGUMENT, "bad memory_space");
    p->instance_sets.push_back(InstanceSet{transforms, colors, count, memory_space, p->instance_total});
    p->instance_total += count;
    return CR_OK;
}

// Renderer::set_clip_depth (src/renderer.rs:932-938)
int cr_pass_set_clip_depth(cr_pass* p, uint32_t clip_depth) {
    if (!p) return fail(CR_ERR_INVALID_ARGUMENT, "null pass");
    if (clip_depth >= (1u << p->renderer->config.clip_nesting_counter_bits)) return fail(CR_ERR_CLIP_STACK_OVERFLOW, "clip depth %u", clip_depth);
    p->clip_depth = clip_depth;
    return CR_OK;
}
// Renderer::save_alpha_context (src/renderer.rs:941-976)
int cr_pass_save_alpha_context(cr_pass* p, uint32_t alpha_layer) {
    if (!p) return fail(CR_ERR_INVALID_ARGUMENT, "null pass");
    if (alpha_layer >= p->renderer->config.alpha_layer_count) return fail(CR_ERR_TOO_MANY_NESTED_OPACITY_GROUPS, "alpha layer %u", alpha_layer);
    p->save_layer = alpha_layer;
    return CR_OK;
}
// Renderer::restore_alpha_context (src/renderer.rs:979-985)
int cr_pass_restore_alpha_context(cr_pass* p, uint32_t alpha_layer) {
    if (!p) return fail(CR_ERR_INVALID_ARGUMENT, "null pass");
    if (alpha_layer >= p->renderer->config.alpha_layer_count) return fail(CR_ERR_TOO_MANY_NESTED_OPACITY_GROUPS, "alpha layer %u", alpha_layer);
    p->restore_layer = alpha_layer;
    return CR_OK;
}

// Records one draw: 32 bytes and no table look-up (the slice tables of the batch live on the device and may still be in
// flight; the expansion into per-category candidate ranges happens on the device at submit).
static int record(cr_pass* p, cr_shape_batch* b, uint32_t shape, uint32_t instance_begin, uint32_t instance_end, uint32_t op) {
    if (op > CR_OP_RESTORE_ALPHA_CONTEXT) return fail(CR_ERR_INVALID_ARGUMENT, "bad render operation %u", op);
    if (b->renderer != p->renderer) return fail(CR_ERR_INVALID_ARGUMENT, "shape belongs to another renderer");
    if (shape >= b->n_shapes) return fail(CR_ERR_INVALID_ARGUMENT, "shape index %u of %u", shape, b->n_shapes);
    if (p->instance_sets.empty()) return fail(CR_ERR_INVALID_ARGUMENT, "cr_pass_set_instances has not been called");
    const InstanceSet& is = p->instance_sets.back();
    if (instance_begin > instance_end || instance_end > is.count) return fail(CR_ERR_INVALID_ARGUMENT, "instances [%u, %u) of %u", instance_begin, instance_end, is.count);
    const bool needs_color = op == CR_OP_COLOR || op == CR_OP_SCALE_ALPHA_CONTEXT || op == CR_OP_RESTORE_ALPHA_CONTEXT;
    if (needs_color && !is.colors) return fail(CR_ERR_INVALID_ARGUMENT, "this operation reads the instance colour slot, which is not bound");
    if ((op == CR_OP_SAVE_ALPHA_CONTEXT || op == CR_OP_RESTORE_ALPHA_CONTEXT) && p->renderer->config.alpha_layer_count == 0)
        return fail(CR_ERR_TOO_MANY_NESTED_OPACITY_GROUPS, "alpha_layer_count is 0");
    if (instance_begin == instance_end) return CR_OK;
    uint32_t bi = 0;
    for (; bi < p->batches.size(); ++bi) if (p->batches[bi] == b) break;
    if (bi == p->batches.size()) p->batches.push_back(b);
    CompactCommand c{};
    c.batch = bi;
    c.shape = shape;
    c.instance_begin = is.base + instance_begin;
    c.instance_count = instance_end - instance_begin;
    c.operation = op;
    c.ref = p->clip_depth << p->renderer->config.winding_counter_bits;
    c.layers = p->save_layer | (p->restore_layer << 16);
    if (!p->arena) { p->commands.push_back(c); return CR_OK; }
    cr_renderer* r = p->renderer;
    if (p->n_arena == r->cmd_arena_cap) {
        const size_t cap = std::max<size_t>(4096, 2 * r->cmd_arena_cap);
        CompactCommand* grown = nullptr;
        CR_CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&grown), cap * sizeof(CompactCommand), cudaHostAllocDefault));
        if (p->n_arena) memcpy(grown, r->cmd_arena, p->n_arena * sizeof(CompactCommand));
        if (r->cmd_arena) cudaFreeHost(r->cmd_arena);
        r->cmd_arena = grown;
        r->cmd_arena_cap = cap;
    }
    r->cmd_arena[p->n_arena++] = c;
    return CR_OK;
}

// Shape::render (src/renderer.rs:267-355)
int cr_shape_render(cr_pass* p, cr_shape* s, uint32_t instance_begin, uint32_t instance_end, uint32_t op) {
    if (!p || !s) return fail(CR_ERR_INVALID_ARGUMENT, "null argument");
    return record(p, s->batch, s->index, instance_begin, instance_end, op);
}
int cr_pass_render_batch(cr_pass* p, cr_shape_batch* b, const cr_draw_command* commands, size_t count) {
    if (!p || !b || (count && !commands)) return fail(CR_ERR_INVALID_ARGUMENT, "null argument");
    if (!p->arena) p->commands.reserve(p->commands.size() + count);
    for (size_t i = 0; i < count; ++i)
        CR_TRY(record(p, b, commands[i].shape_index, commands[i].instance_begin, commands[i].instance_end, commands[i].render_operation));
    return CR_OK;
}

int cr_pass_render_script(cr_pass* p, cr_shape_batch* b, const cr_scripted_draw* draws, size_t count) {
    if (!p || !b || (count && !draws)) return fail(CR_ERR_INVALID_ARGUMENT, "null argument");
    if (!p->arena) p->commands.reserve(p->commands.size() + count);
    for (size_t i = 0; i < count; ++i) {
        const cr_scripted_draw& d = draws[i];
        if (d.clip_depth != p->clip_depth) CR_TRY(cr_pass_set_clip_depth(p, d.clip_depth));
        if (d.save_alpha_layer != p->save_layer) CR_TRY(cr_pass_save_alpha_context(p, d.save_alpha_layer));
        if (d.restore_alpha_layer != p->restore_layer) CR_TRY(cr_pass_restore_alpha_context(p, d.restore_alpha_layer));
        CR_TRY(record(p, b, d.shape_index, d.instance_begin, d.instance_end, d.render_operation));
    }
    return CR_OK;
}

// The pass's view of the attachments and of the renderer's configuration.
static RasterTarget make_target(const cr_pass* p) {
    const cr_renderer* r = p->renderer;
    RasterTarget tg{};
    tg.color = r->color.p;
    tg.stencil = r->stencil.as<uint8_t>();
    tg.alpha_layers = r->alpha_layers.p;
    tg.depth = has_depth(r) ? r->depth.as<float>() : nullptr;
    tg.depth_compare = r->config.depth_compare;
    tg.depth_write = r->config.depth_write_enabled;
    tg.clear_depth = p->clear_depth ? 1u : 0u;
    tg.depth_clear_value = p->depth_clear_value;
    tg.color_format = r->config.color_format;
    tg.width = r->width; tg.height = r->height; tg.tiles_x = r->tiles_x; tg.tiles_y = r->tiles_y;
    tg.samples = r->config.msaa_sample_count;
    tg.sample_lo = tg.samples == 4 ? 32 : 128;
    tg.sample_hi = tg.samples == 4 ? 224 : 128;
    tg.wmask = (1u << r->config.winding_counter_bits) - 1u;
    tg.cmask = ((1u << r->config.clip_nesting_counter_bits) - 1u) << r->config.winding_counter_bits;
    tg.blending = r->config.blending;
    tg.cull_mode = r->config.cull_mode;
    {
        static const int env_run = getenv("CR_PIXEL_RUN_MAX") ? atoi(getenv("CR_PIXEL_RUN_MAX")) : -1;   // tuning knob for experiments
        tg.pixel_run_max = env_run >= 0 ? (uint32_t)env_run : 12u;
    }
    tg.clear_color = p->clear_color ? 1u : 0u;
    tg.clear_stencil = p->clear_stencil ? 1u : 0u;
    tg.shard_world = r->shard_world;
    tg.shard_rank = r->shard_rank;
    for (int i = 0; i < CR_MAX_PEERS; ++i) { tg.peer_color[i] = r->peer_color[i]; tg.peer_stencil[i] = static_cast<uint8_t*>(r->peer_stencil[i]); }
    tg.order_world = r->order_world;
    tg.order_rank = r->order_rank;
    tg.order_epoch = r->order_epoch;
    tg.order_mask_words = (r->tiles_x * r->tiles_y + 31) / 32;
    tg.exchange = r->exchange.as<uint32_t>();
    for (int i = 0; i < CR_MAX_PEERS; ++i) tg.peer_exchange[i] = static_cast<uint32_t*>(r->peer_exchange[i]);
    return tg;
}

// The pass's clear when there is nothing to rasterise. Single GPU: plain memsets. Tile-sharded target: a rank may only touch
// the tiles it owns (the others arrive from their owners, possibly before this call), so the tile kernel runs on an empty
// tile table: every owned tile is cleared in all ranks' attachments. The depth aspect (cleared to a float) goes the same way.
static int clear_attachments(cr_pass* p) {
    cr_renderer* r = p->renderer;
    const bool depth_clear = p->clear_depth && has_depth(r);
    if (r->order_world <= 1 && !p->clear_color && !p->clear_stencil && !depth_clear) return CR_OK;
    if (r->shard_world <= 1 && r->order_world <= 1 && !depth_clear) {
        const size_t samples = (size_t)r->width * r->height * r->config.msaa_sample_count;
        if (p->clear_color) CR_CUDA_TRY(cudaMemsetAsync(r->color.p, 0, samples * color_texel_bytes(r), r->stream));
        if (p->clear_stencil) CR_CUDA_TRY(cudaMemsetAsync(r->stencil.p, 0, samples, r->stream));
        return CR_OK;
    }
    const uint32_t n_tiles = r->tiles_x * r->tiles_y;
    CR_TRY(r->tile_begin.reserve(r->stream, (size_t)(n_tiles + 1) * 4));
    CR_CUDA_TRY(cudaMemsetAsync(r->tile_begin.p, 0, (size_t)(n_tiles + 1) * 4, r->stream));
    RasterScene none{};
    if (r->order_world > 1) {   // a rank whose slice draws nothing still tells the others so, in this pass's epoch
        r->order_epoch += 1;
        CR_TRY(cr_raster_publish_touched_tiles(r->stream, make_target(p), r->tile_begin.as<uint32_t>()));
    }
    return cr_raster_tiles(r->stream, none, make_target(p), nullptr, r->tile_begin.as<uint32_t>(), nullptr);
}

// Enqueues the whole pass. `sized` = false: OPTIMISTIC — buffers and grids are sized from the capacities the previous pass left
// behind, nothing is read back before everything is enqueued; the kernels check the capacities on the device and the tile
// kernel does not run when one did not suffice (settle() then re-submits). `sized` = true: the host waits for the candidate
// and pair totals and sizes exactly (first pass of a renderer, tile-sharded targets, re-submission).
static int enqueue_pass(cr_pass* p, bool sized, int attempt = 0) {
    cr_renderer* r = p->renderer;
    cudaStream_t st = r->stream;
    const CompactCommand* const cmds = p->arena ? r->cmd_arena : p->commands.data();
    const uint32_t n_cmds = (uint32_t)(p->arena ? p->n_arena : p->commands.size());
    CR_TRY(r->pass_counters.reserve(st, sizeof(PassCounters)));
    PassCounters* counters = r->pass_counters.as<PassCounters>();
    if (r->order_world > 1) r->order_epoch += 1;   // every rank submits the same number of passes: the epoch names this one on all of them
    // ---- scene description: batches, commands (expanded on the device), candidate numbering
    CR_TRY(r->batches_dev.reserve(st, p->batches.size() * sizeof(DeviceBatch)));
    if (p->sets.size() != p->batches.size()) {   // first enqueue of this pass: it renders the build each batch holds NOW (a re-submission keeps them)
        p->sets.clear();
        for (cr_shape_batch* b : p->batches) p->sets.push_back(&b->store());
    }
    for (size_t i = 0; i < p->batches.size(); ++i) {
        BatchStorage* B = p->sets[i];
        if (B->written) CR_CUDA_TRY(cudaStreamWaitEvent(st, B->written, 0));   // the build may still be running on the tessellation stream
        CR_CUDA_TRY(cudaMemcpyAsync(r->batches_dev.as<DeviceBatch>() + i, B->desc_dev.p, sizeof(DeviceBatch), cudaMemcpyDeviceToDevice, st));
    }
    CR_TRY(r->compact_dev.reserve(st, (size_t)n_cmds * sizeof(CompactCommand)));
    CR_CUDA_TRY(cudaMemcpyAsync(r->compact_dev.p, cmds, (size_t)n_cmds * sizeof(CompactCommand), cudaMemcpyHostToDevice, st));
    CR_TRY(r->cmds_dev.reserve(st, (size_t)n_cmds * sizeof(DeviceCommand)));
    CR_TRY(r->cmd_cands.reserve(st, (size_t)(n_cmds + 1) * 4));
    CR_TRY(reserve_scan_scratch(r, st, r->scan_scratch, cr_scan_scratch_words(n_cmds + 1, 1)));
    if (r->timing) CR_CUDA_TRY(cudaEventRecord(r->ev[2], st));
    CR_TRY(cr_raster_expand(st, r->compact_dev.as<CompactCommand>(), n_cmds, r->batches_dev.as<DeviceBatch>(), r->cmds_dev.as<DeviceCommand>(), r->cmd_cands.as<uint32_t>(),
                            counters, r->scan_scratch.as<uint32_t>()));
    if (sized) {
        CR_CUDA_TRY(cudaMemcpyAsync(&r->pinned[PIN_PASS], counters, 8, cudaMemcpyDeviceToHost, st));
        CR_CUDA_TRY(cudaStreamSynchronize(st));
        unsigned long long cand_total = 0;
        memcpy(&cand_total, &r->pinned[PIN_PASS], 8);
        if (cand_total >= 0xFFFFFFFFull) return fail(CR_ERR_INVALID_ARGUMENT, "%llu candidate primitives in one pass exceed 2^32; submit in several passes", cand_total);
        r->last_cands = (uint32_t)cand_total;
        r->cand_cap = std::max<uint32_t>(r->cand_cap, (uint32_t)cand_total);
    }
    const uint32_t cand_cap = r->cand_cap;

    RasterScene sc{};
    sc.batches = r->batches_dev.as<DeviceBatch>();
    sc.commands = r->cmds_dev.as<DeviceCommand>();
    sc.cmd_cand_begin = r->cmd_cands.as<uint32_t>();
    sc.n_commands = n_cmds;
    sc.transforms = r->inst_transforms.as<float>();
    sc.colors = p->any_color ? r->inst_colors.as<float>() : nullptr;
    const uint32_t clip_cap = r->clip_cap;
    CR_TRY(r->clip_attrs.reserve(st, (size_t)clip_cap * sizeof(ClipAttr)));
    sc.clip_attrs = r->clip_attrs.as<ClipAttr>();
    const RasterTarget tg = make_target(p);
    const uint32_t n_tiles = r->tiles_x * r->tiles_y;

    // ---- bin: count, scan, emit, sort by tile (stable => draw order survives inside every tile)
    CR_TRY(r->cand_tiles.reserve(st, (size_t)(cand_cap + 1) * 4));
    CR_TRY(r->records.reserve(st, ((size_t)std::max<uint32_t>(cand_cap, 1u) + clip_cap) * sizeof(PrimRecord)));   // the fan triangles of frustum clipping live behind the candidates
    CR_TRY(r->clip_list.reserve(st, (size_t)(cand_cap + 1) * 4));
    CR_TRY(reserve_scan_scratch(r, st, r->scan_scratch, cr_scan_scratch_words(cand_cap + 1, 1)));
    CR_TRY(r->big_list.reserve(st, (size_t)(cand_cap + 1) * 4));
    CR_TRY(cr_raster_setup(st, sc, tg, cand_cap, r->records.as<PrimRecord>(), r->cand_tiles.as<uint32_t>(), r->big_list.as<uint32_t>(), r->clip_list.as<uint32_t>(),
                           r->clip_attrs.as<ClipAttr>(), clip_cap, counters));
    if (sized) {
        CR_CUDA_TRY(cudaMemcpyAsync(&r->pinned[PIN_PASS], counters, sizeof(PassCounters), cudaMemcpyDeviceToHost, st));
        CR_CUDA_TRY(cudaStreamSynchronize(st));
        PassCounters seen;
        memcpy(&seen, &r->pinned[PIN_PASS], sizeof(seen));
        if ((seen.flags & CR_PASS_OVERFLOW_CLIP) != 0u && attempt == 0) {   // more clipped triangles than the capacity: size it and run the vertex stage again
            r->clip_cap = std::max<uint32_t>(r->clip_cap, seen.clip_total + seen.clip_total / 8 + 64);
            if (r->order_world > 1) r->order_epoch -= 1;   // nothing of this attempt has used the epoch yet
            return enqueue_pass(p, true, 1);
        }
        const unsigned long long pair_total = seen.pair_total;
        if (pair_total >= 0xFFFFFFFFull)
            return fail(CR_ERR_INVALID_ARGUMENT, "%llu (tile, primitive) pairs in one pass exceed 2^32; submit in several passes", pair_total);
        r->last_pairs = (uint32_t)pair_total;
        r->pair_cap = std::max<uint32_t>(r->pair_cap, (uint32_t)pair_total);
    }
    const uint32_t pair_cap = r->pair_cap;
    if (cand_cap) CR_TRY(cr_scan_exclusive(st, r->cand_tiles.as<uint32_t>(), cand_cap + 1, 1, r->scan_scratch.as<uint32_t>()));
    CR_TRY(r->pair_tile.reserve(st, (size_t)std::max<uint32_t>(pair_cap, 1u) * 4));
    CR_TRY(r->pair_cand.reserve(st, (size_t)std::max<uint32_t>(pair_cap, 1u) * 4));
    CR_TRY(r->pair_tile_alt.reserve(st, (size_t)std::max<uint32_t>(pair_cap, 1u) * 4));
    CR_TRY(r->pair_cand_alt.reserve(st, (size_t)std::max<uint32_t>(pair_cap, 1u) * 4));
    CR_TRY(reserve_scan_scratch(r, st, r->radix_scratch, cr_radix_scratch_words(pair_cap)));
    if (r->radix_layout_cap != pair_cap) {
        // The sort keeps its histograms in front of the scan's tickets and status words, at an offset that depends on the pair
        // capacity: with another capacity the tickets would lie where histogram values were (a scan that finds a non-zero ticket
        // never ends). Zero the scratch whenever the layout moves.
        CR_TRY(cr_scan_prepare(st, r->radix_scratch.as<uint32_t>(), r->radix_scratch.cap / 4));
        r->radix_layout_cap = pair_cap;
    }
    CR_TRY(r->tile_begin.reserve(st, (size_t)(n_tiles + 1) * 4));
    CR_TRY(cr_raster_bin_emit(st, tg, cand_cap, pair_cap, r->records.as<PrimRecord>(), r->cand_tiles.as<uint32_t>(), r->big_list.as<uint32_t>(), r->clip_list.as<uint32_t>(),
                              r->pair_tile.as<uint32_t>(), r->pair_cand.as<uint32_t>(), counters));
    if (cand_cap == 0) {   // no candidates: bin_emit (which publishes the live pair count) did not run
        CR_CUDA_TRY(cudaMemsetAsync(&counters->n_pairs_live, 0, 4, st));
    }
    uint32_t key_bits = 1;
    while ((1u << key_bits) < n_tiles) ++key_bits;
    uint32_t *sorted_tile = r->pair_tile.as<uint32_t>(), *sorted_cand = r->pair_cand.as<uint32_t>();
    CR_TRY(cr_radix_sort_pairs(st, r->pair_tile.as<uint32_t>(), r->pair_cand.as<uint32_t>(), r->pair_tile_alt.as<uint32_t>(), r->pair_cand_alt.as<uint32_t>(), pair_cap,
                               &counters->n_pairs_live, key_bits, r->radix_scratch.as<uint32_t>(), &sorted_tile, &sorted_cand));
    CR_TRY(cr_lower_bounds(st, sorted_tile, pair_cap, &counters->n_pairs_live, r->tile_begin.as<uint32_t>(), n_tiles + 1));
    if (r->order_world > 1) CR_TRY(cr_raster_publish_touched_tiles(st, tg, r->tile_begin.as<uint32_t>()));
    if (r->timing) { CR_CUDA_TRY(cudaEventRecord(r->ev[3], st)); CR_CUDA_TRY(cudaEventRecord(r->ev[4], st)); }
    // tile-ordered primitive stream (set up per (tile, primitive) pair), then the tile kernel that bulk-loads it
    CR_TRY(r->tile_prims.reserve(st, (size_t)std::max<uint32_t>(pair_cap, 1u) * cr_tile_prim_bytes()));
    CR_TRY(cr_raster_tile_prims(st, sc, tg, r->records.as<PrimRecord>(), sorted_tile, sorted_cand, pair_cap, r->tile_prims.p, counters));
    CR_TRY(cr_raster_tiles(st, sc, tg, r->tile_prims.p, r->tile_begin.as<uint32_t>(), counters));
    if (r->timing) { CR_CUDA_TRY(cudaEventRecord(r->ev[5], st)); r->ev_valid[1] = r->ev_valid[2] = true; }
    CR_CUDA_TRY(cudaMemcpyAsync(&r->pinned[PIN_PASS], counters, sizeof(PassCounters), cudaMemcpyDeviceToHost, st));
    CR_CUDA_TRY(cudaEventRecord(r->ev_pass, st));
    for (BatchStorage* B : p->sets) {   // a rebuild into these arrays (and their release) waits for this pass
        if (!B->last_read) CR_CUDA_TRY(cudaEventCreateWithFlags(&B->last_read, cudaEventDisableTiming));
        CR_CUDA_TRY(cudaEventRecord(B->last_read, st));
        B->was_read = true;
    }
    return CR_OK;
}

static int submit(cr_pass* p) {
    cr_renderer* r = p->renderer;
    cudaStream_t st = r->stream;
    const uint32_t n_cmds = (uint32_t)(p->arena ? p->n_arena : p->commands.size());
    if (n_cmds == 0) {
        r->stats.primitives = r->stats.tile_pairs = r->stats.covered_samples = 0;
        return clear_attachments(p);
    }
    // ---- instance slots: always copied, so that a re-submission (settle) reads what this call was given
    CR_TRY(r->inst_transforms.reserve(st, (size_t)p->instance_total * 64));
    CR_TRY(r->inst_colors.reserve(st, (size_t)p->instance_total * 16));
    p->any_color = false;
    for (const InstanceSet& is : p->instance_sets) {
        if (!is.count) continue;
        const cudaMemcpyKind kind = is.space == CR_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
        CR_CUDA_TRY(cudaMemcpyAsync(r->inst_transforms.as<float>() + 16 * (size_t)is.base, is.transforms, (size_t)is.count * 64, kind, st));
        if (is.colors) {
            CR_CUDA_TRY(cudaMemcpyAsync(r->inst_colors.as<float>() + 4 * (size_t)is.base, is.colors, (size_t)is.count * 16, kind, st));
            p->any_color = true;
        }
    }
    // optimistic unless there is nothing to go by, or the target spans several GPUs (the other ranks wait for this rank's tiles)
    const bool sized = r->cand_cap == 0 || r->pair_cap == 0 || r->shard_world > 1 || r->order_world > 1;
    CR_TRY(enqueue_pass(p, sized));
    if (sized && r->shard_world <= 1 && r->order_world <= 1) {
        // leave slack for the next (optimistic) pass: scenes drift from frame to frame. Based on what THIS pass needed, so that
        // a renderer whose passes are all sized never compounds it.
        r->cand_cap = std::max<uint32_t>(r->cand_cap, (uint32_t)std::min<uint64_t>(0xFFFFFFFEull, (uint64_t)r->last_cands + r->last_cands / 8 + 4096));
        r->pair_cap = std::max<uint32_t>(r->pair_cap, (uint32_t)std::min<uint64_t>(0xFFFFFFFEull, (uint64_t)r->last_pairs + r->last_pairs / 8 + 4096));
    }
    return CR_OK;
}

extern "C++" {
namespace {
// The device-side sizes of the pass submitted last: wait for them (they have long arrived unless the caller comes straight
// back), keep them as statistics and capacities, and re-submit the pass with exact sizes if a capacity did not suffice.
int settle(cr_renderer* r) {
    cr_pass* p = r->inflight;
    if (!p) return CR_OK;
    r->inflight = nullptr;
    int status = CR_OK;
    if (cudaEventSynchronize(r->ev_pass) != cudaSuccess) status = fail(CR_ERR_CUDA, "waiting for the pass failed: %s", cudaGetErrorString(cudaGetLastError()));
    PassCounters pc;
    memcpy(&pc, &r->pinned[PIN_PASS], sizeof(pc));
    if (status == CR_OK && pc.flags != 0u) {
        if (pc.cand_total >= 0xFFFFFFFFull) status = fail(CR_ERR_INVALID_ARGUMENT, "%llu candidate primitives in one pass exceed 2^32; submit in several passes", pc.cand_total);
        else {
            r->cand_cap = std::max<uint32_t>(r->cand_cap, (uint32_t)pc.cand_total);
            if (pc.flags & CR_PASS_OVERFLOW_CLIP) r->clip_cap = std::max<uint32_t>(r->clip_cap, pc.clip_total + pc.clip_total / 8 + 64);
            status = enqueue_pass(p, true);
            if (status == CR_OK) {
                if (cudaEventSynchronize(r->ev_pass) != cudaSuccess) status = fail(CR_ERR_CUDA, "waiting for the re-submitted pass failed");
                memcpy(&pc, &r->pinned[PIN_PASS], sizeof(pc));
            }
            // an asynchronous read-back taken behind the first (skipped) attempt holds the frame as it was before the pass: take it again
            for (auto& rb : r->readback) {
                if (!rb.pending || rb.pass_serial != r->inflight_serial || status != CR_OK) continue;
                cudaStreamSynchronize(r->out);
                if (cudaMemcpyAsync(rb.dst, r->color.p, rb.bytes, cudaMemcpyDeviceToHost, r->stream) != cudaSuccess || cudaStreamSynchronize(r->stream) != cudaSuccess)
                    status = fail(CR_ERR_CUDA, "repeating the frame read-back failed");
                cudaEventRecord(rb.done, r->stream);
            }
        }
    }
    if (status == CR_OK) {
        r->stats.primitives = pc.cand_total;
        r->stats.tile_pairs = pc.pair_total;
        r->stats.covered_samples = pc.covered;
    } else if (r->deferred_status == CR_OK) {
        r->deferred_status = status;
        snprintf(r->deferred_message, sizeof(r->deferred_message), "(pass submitted earlier) %s", g_error_message);
    }
    pass_free(p);
    renderer_release_child(r);
    return CR_OK;
}
}  // namespace
}  // extern "C++"

// queue.submit(encoder.finish()) (examples/showcase/main.rs:252). Returns once everything is enqueued. A pass whose sizes were
// only estimated (see enqueue_pass) is checked by the next call on the renderer that needs its result or its resources;
// an error found then (more than 2^32 primitives or pairs) is returned by that call.
int cr_pass_submit(cr_pass* p) {
    if (!p) return fail(CR_ERR_INVALID_ARGUMENT, "null pass");
    cr_renderer* r = p->renderer;
    int st;
    bool keep = false;
    {
        DeviceGuard guard(r->device);
        if (!guard.ok) st = fail(CR_ERR_CUDA, "cannot select CUDA device");
        else {
            st = settle(r);
            if (st == CR_OK) st = take_deferred(r);
            if (st == CR_OK) st = submit(p);
            const uint32_t n_cmds = (uint32_t)(p->arena ? p->n_arena : p->commands.size());
            keep = st == CR_OK && n_cmds != 0;
        }
    }
    if (keep) { r->inflight = p; r->inflight_serial = ++r->pass_serial; return CR_OK; }   // its commands, batches and instance copies stay until settle()
    pass_free(p);
    renderer_release_child(r);
    return st;
}

void cr_pass_abort(cr_pass* p) {
    if (!p) return;
    cr_renderer* r = p->renderer;
    pass_free(p);
    renderer_release_child(r);
}

// ------------------------------------------------------------------------------------------------- read-back
static int read_back(cr_renderer* r, const void* src, size_t bytes, void* dst, size_t capacity) {
    if (!r || !dst) return fail(CR_ERR_INVALID_ARGUMENT, "null argument");
    if (r->width == 0) return fail(CR_ERR_NOT_RESIZED, "cr_renderer_resize has not been called");
    if (capacity < bytes) return fail(CR_ERR_INVALID_ARGUMENT, "capacity %zu < %zu", capacity, bytes);
    CR_GUARD(r);
    CR_TRY(settle(r));
    CR_TRY(take_deferred(r));
    CR_CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, r->stream));
    CR_CUDA_TRY(cudaStreamSynchronize(r->stream));
    return CR_OK;
}
int cr_renderer_read_color(cr_renderer* r, float* dst, size_t capacity_bytes) {
    if (!r) return fail(CR_ERR_INVALID_ARGUMENT, "null renderer");
    const size_t samples = (size_t)r->width * r->height * r->config.msaa_sample_count;
    if (r->config.color_format == CR_FORMAT_RGBA32F) return read_back(r, r->color.p, samples * 16, dst, capacity_bytes);
    if (capacity_bytes < samples * 16) return fail(CR_ERR_INVALID_ARGUMENT, "capacity %zu < %zu", capacity_bytes, samples * 16);
    std::vector<uint32_t> texels(samples);   // unorm8 texels -> the floats a shader would read: c / 255, in RGBA order
    CR_TRY(read_back(r, r->color.p, samples * 4, texels.data(), samples * 4));
    const bool bgra = r->config.color_format == CR_FORMAT_BGRA8_UNORM;
    for (size_t i = 0; i < samples; ++i) {
        const uint32_t t = texels[i];
        const float c0 = (float)(t & 255u) / 255.0f, c1 = (float)((t >> 8) & 255u) / 255.0f, c2 = (float)((t >> 16) & 255u) / 255.0f;
        dst[4 * i + 0] = bgra ? c2 : c0; dst[4 * i + 1] = c1; dst[4 * i + 2] = bgra ? c0 : c2; dst[4 * i + 3] = (float)(t >> 24) / 255.0f;
    }
    return CR_OK;
}
int cr_renderer_read_color_texels(cr_renderer* r, void* dst, size_t capacity_bytes) {
    if (!r) return fail(CR_ERR_INVALID_ARGUMENT, "null renderer");
    return read_back(r, r->color.p, (size_t)r->width * r->height * r->config.msaa_sample_count * color_texel_bytes(r), dst, capacity_bytes);
}
// The frame of the pass submitted last, without stopping the pipeline: snapshot behind the pass, host copy on its own stream.
int cr_renderer_read_color_texels_async(cr_renderer* r, void* dst, size_t capacity_bytes, uint64_t* ticket) {
    if (!r || !dst || !ticket) return fail(CR_ERR_INVALID_ARGUMENT, "null argument");
    if (r->width == 0) return fail(CR_ERR_NOT_RESIZED, "cr_renderer_resize has not been called");
    const size_t bytes = (size_t)r->width * r->height * r->config.msaa_sample_count * color_texel_bytes(r);
    if (capacity_bytes < bytes) return fail(CR_ERR_INVALID_ARGUMENT, "capacity %zu < %zu", capacity_bytes, bytes);
    CR_GUARD(r);
    if (!r->out) CR_CUDA_TRY(cudaStreamCreateWithFlags(&r->out, cudaStreamNonBlocking));
    cr_renderer::Readback& rb = r->readback[r->readback_tickets & 3u];
    if (rb.pending) {   // four read-backs are in flight at most: the older one of this slot finishes first
        CR_CUDA_TRY(cudaEventSynchronize(rb.done));
        rb.pending = false;
    }
    if (!rb.snapshot) { CR_CUDA_TRY(cudaEventCreateWithFlags(&rb.snapshot, cudaEventDisableTiming)); CR_CUDA_TRY(cudaEventCreateWithFlags(&rb.done, cudaEventDisableTiming)); }
    CR_TRY(rb.buf.reserve(r->stream, bytes));
    CR_CUDA_TRY(cudaMemcpyAsync(rb.buf.p, r->color.p, bytes, cudaMemcpyDeviceToDevice, r->stream));
    CR_CUDA_TRY(cudaEventRecord(rb.snapshot, r->stream));
    CR_CUDA_TRY(cudaStreamWaitEvent(r->out, rb.snapshot, 0));
    CR_CUDA_TRY(cudaMemcpyAsync(dst, rb.buf.p, bytes, cudaMemcpyDeviceToHost, r->out));
    CR_CUDA_TRY(cudaEventRecord(rb.done, r->out));
    rb.dst = dst; rb.bytes = bytes; rb.pending = true;
    rb.pass_serial = r->inflight ? r->inflight_serial : 0;   // 0: no pass in flight, nothing can be re-submitted under it
    rb.ticket = ++r->readback_tickets;
    *ticket = rb.ticket;
    return CR_OK;
}
// Waits until the read-back `ticket` has arrived in its destination. The pass it belongs to is settled first: if that pass had to
// be re-submitted (a capacity did not suffice), the frame is read again.
int cr_renderer_wait_readback(cr_renderer* r, uint64_t ticket) {
    if (!r) return fail(CR_ERR_INVALID_ARGUMENT, "null renderer");
    CR_GUARD(r);
    for (auto& rb : r->readback) {
        if (!rb.pending || rb.ticket != ticket) continue;
        if (r->inflight && rb.pass_serial == r->inflight_serial) { CR_TRY(settle(r)); CR_TRY(take_deferred(r)); }
        CR_CUDA_TRY(cudaEventSynchronize(rb.done));
        rb.pending = false;
        return CR_OK;
    }
    return CR_OK;   // already waited for (or replaced by a newer read-back of its slot, which waited for it)
}
int cr_renderer_read_depth(cr_renderer* r, float* dst, size_t capacity_bytes) {
    if (!r) return fail(CR_ERR_INVALID_ARGUMENT, "null renderer");
    if (!has_depth(r)) return fail(CR_ERR_INVALID_ARGUMENT, "this configuration has no depth attachment (depth_compare Always, depth_write_enabled 0)");
    return read_back(r, r->depth.p, (size_t)r->width * r->height * r->config.msaa_sample_count * 4, dst, capacity_bytes);
}
int cr_renderer_read_stencil(cr_renderer* r, uint8_t* dst, size_t capacity_bytes) {
    if (!r) return fail(CR_ERR_INVALID_ARGUMENT, "null renderer");
    return read_back(r, r->stencil.p, (size_t)r->width * r->height * r->config.msaa_sample_count, dst, capacity_bytes);
}
int cr_renderer_read_alpha_layer(cr_renderer* r, uint32_t layer, float* dst, size_t capacity_bytes) {
    if (!r) return fail(CR_ERR_INVALID_ARGUMENT, "null renderer");
    if (layer >= r->config.alpha_layer_count) return fail(CR_ERR_TOO_MANY_NESTED_OPACITY_GROUPS, "alpha layer %u", layer);
    const size_t samples = (size_t)r->width * r->height * r->config.msaa_sample_count;
    if (r->config.color_format == CR_FORMAT_RGBA32F)
        return read_back(r, static_cast<const char*>(r->alpha_layers.p) + layer * samples * 4, samples * 4, dst, capacity_bytes);
    if (capacity_bytes < samples * 4) return fail(CR_ERR_INVALID_ARGUMENT, "capacity %zu < %zu", capacity_bytes, samples * 4);
    std::vector<uint8_t> texels(samples);   // R8Unorm layer (src/renderer.rs:898)
    CR_TRY(read_back(r, static_cast<const char*>(r->alpha_layers.p) + layer * samples, samples, texels.data(), samples));
    for (size_t i = 0; i < samples; ++i) dst[i] = (float)texels[i] / 255.0f;
    return CR_OK;
}
int cr_renderer_get_attachments(cr_renderer* r, void** color_dev, void** stencil_dev) {
    if (!r) return fail(CR_ERR_INVALID_ARGUMENT, "null renderer");
    if (r->width == 0) return fail(CR_ERR_NOT_RESIZED, "cr_renderer_resize has not been called");
    if (color_dev) *color_dev = r->color.p;
    if (stencil_dev) *stencil_dev = r->stencil.p;
    return CR_OK;
}

// ---- one render target spanning several GPUs: tile ownership + peer-mapped attachments
int cr_renderer_set_tile_sharding(cr_renderer* r, uint32_t world, uint32_t rank) {
    if (!r) return fail(CR_ERR_INVALID_ARGUMENT, "null renderer");
    if (world == 0 || world > CR_MAX_PEERS + 1 || rank >= world) return fail(CR_ERR_INVALID_ARGUMENT, "bad tile sharding %u of %u (at most %d ranks)", rank, world, CR_MAX_PEERS + 1);
    CR_GUARD(r);
    CR_TRY(settle(r));
    CR_CUDA_TRY(cudaStreamSynchronize(r->stream));
    close_peers(r);
    r->order_world = 1; r->order_rank = 0;
    r->shard_world = world;
    r->shard_rank = rank;
    r->cand_cap = r->pair_cap = 0;   // the pair counts of a rank depend on the tiles it owns
    return CR_OK;
}
int cr_renderer_export_attachments(cr_renderer* r, uint8_t* color_handle, uint8_t* stencil_handle) {
    static_assert(sizeof(cudaIpcMemHandle_t) == CR_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t is 64 bytes");
    if (!r || !color_handle || !stencil_handle) return fail(CR_ERR_INVALID_ARGUMENT, "null argument");
    if (r->width == 0) return fail(CR_ERR_NOT_RESIZED, "cr_renderer_resize has not been called");
    CR_GUARD(r);
    CR_CUDA_TRY(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(color_handle), r->color.p));
    CR_CUDA_TRY(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(stencil_handle), r->stencil.p));
    return CR_OK;
}
int cr_renderer_import_peer_attachments(cr_renderer* r, uint32_t peer_rank, const uint8_t* color_handle, const uint8_t* stencil_handle) {
    if (!r || !color_handle || !stencil_handle) return fail(CR_ERR_INVALID_ARGUMENT, "null argument");
    const uint32_t world = std::max(r->shard_world, r->order_world), rank = r->order_world > 1 ? r->order_rank : r->shard_rank;
    if (peer_rank >= world || peer_rank == rank) return fail(CR_ERR_INVALID_ARGUMENT, "peer rank %u is not another rank of a %u-rank target", peer_rank, world);
    CR_GUARD(r);
    const uint32_t slot = peer_rank - (peer_rank > rank ? 1u : 0u);
    if (r->peer_color[slot] || r->peer_stencil[slot]) return fail(CR_ERR_INVALID_ARGUMENT, "peer rank %u is already imported", peer_rank);
    cudaIpcMemHandle_t hc, hs;
    memcpy(&hc, color_handle, sizeof(hc));
    memcpy(&hs, stencil_handle, sizeof(hs));
    CR_CUDA_TRY(cudaIpcOpenMemHandle(&r->peer_color[slot], hc, cudaIpcMemLazyEnablePeerAccess));
    CR_CUDA_TRY(cudaIpcOpenMemHandle(&r->peer_stencil[slot], hs, cudaIpcMemLazyEnablePeerAccess));
    return CR_OK;
}

// ---- one render target composed from draw-order slices (batch sharding into one target)
int cr_renderer_set_order_sharding(cr_renderer* r, uint32_t world, uint32_t rank) {
    if (!r) return fail(CR_ERR_INVALID_ARGUMENT, "null renderer");
    if (world == 0 || world > CR_MAX_PEERS + 1 || rank >= world) return fail(CR_ERR_INVALID_ARGUMENT, "bad order sharding %u of %u (at most %d ranks)", rank, world, CR_MAX_PEERS + 1);
    if (r->width == 0) return fail(CR_ERR_NOT_RESIZED, "cr_renderer_resize has not been called");
    if (world > 1 && (has_depth(r) || r->config.alpha_layer_count != 0))
        return fail(CR_ERR_INVALID_ARGUMENT, "draw-order sharding hands colour and stencil from rank to rank; depth and alpha layers are not exchanged");
    CR_GUARD(r);
    CR_TRY(settle(r));
    CR_CUDA_TRY(cudaStreamSynchronize(r->stream));
    close_peers(r);
    r->shard_world = 1; r->shard_rank = 0;
    r->order_world = world;
    r->order_rank = rank;
    r->order_epoch = 0;
    r->cand_cap = r->pair_cap = 0;
    if (world > 1) {
        const size_t words = cr_exchange_words(r->tiles_x * r->tiles_y);
        r->exchange.plain = true;   // exportable with cudaIpcGetMemHandle
        CR_TRY(r->exchange.reserve(r->stream, words * 4));
        CR_CUDA_TRY(cudaMemsetAsync(r->exchange.p, 0, words * 4, r->stream));   // epoch 0 is never live
        CR_CUDA_TRY(cudaStreamSynchronize(r->stream));
    }
    return CR_OK;
}
int cr_renderer_export_exchange(cr_renderer* r, uint8_t* handle) {
    if (!r || !handle) return fail(CR_ERR_INVALID_ARGUMENT, "null argument");
    if (r->order_world <= 1 || !r->exchange.p) return fail(CR_ERR_INVALID_ARGUMENT, "cr_renderer_set_order_sharding has not been called");
    CR_GUARD(r);
    CR_CUDA_TRY(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle), r->exchange.p));
    return CR_OK;
}
int cr_renderer_import_peer_exchange(cr_renderer* r, uint32_t peer_rank, const uint8_t* handle) {
    if (!r || !handle) return fail(CR_ERR_INVALID_ARGUMENT, "null argument");
    if (peer_rank >= r->order_world || peer_rank == r->order_rank) return fail(CR_ERR_INVALID_ARGUMENT, "peer rank %u is not another rank of a %u-rank target", peer_rank, r->order_world);
    CR_GUARD(r);
    const uint32_t slot = peer_rank - (peer_rank > r->order_rank ? 1u : 0u);
    if (r->peer_exchange[slot]) return fail(CR_ERR_INVALID_ARGUMENT, "peer rank %u is already imported", peer_rank);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    CR_CUDA_TRY(cudaIpcOpenMemHandle(&r->peer_exchange[slot], h, cudaIpcMemLazyEnablePeerAccess));
    return CR_OK;
}

int cr_renderer_enable_timing(cr_renderer* r, uint32_t enabled) {
    if (!r) return fail(CR_ERR_INVALID_ARGUMENT, "null renderer");
    r->timing = enabled != 0;
    r->ev_valid[0] = r->ev_valid[1] = r->ev_valid[2] = false;
    return CR_OK;
}
// The counters of the most recent pass that has been SETTLED (cr_pass_submit settles the pass before the one it submits), without
// waiting for anything: with frame pipelining the host reads the result of frame N while frame N + 1 is running.
int cr_renderer_get_settled_pass_stats(cr_renderer* r, cr_stats* out) {
    if (!r || !out) return fail(CR_ERR_INVALID_ARGUMENT, "null argument");
    CR_TRY(take_deferred(r));
    r->stats.kernel_launches = g_cr_kernel_launches;
    *out = r->stats;
    return CR_OK;
}
int cr_renderer_get_stats(cr_renderer* r, cr_stats* out) {
    if (!r || !out) return fail(CR_ERR_INVALID_ARGUMENT, "null argument");
    CR_GUARD(r);
    CR_TRY(settle(r));
    CR_TRY(take_deferred(r));
    CR_CUDA_TRY(cudaStreamSynchronize(r->stream));
    if (r->pipelined) for (cudaStream_t q : {r->copy, r->tess, r->hull}) CR_CUDA_TRY(cudaStreamSynchronize(q));
    if (r->stats_batch) {   // hull vertices of the last from_paths: its counts arrive with the batch's mirrors
        cr_shape_batch* b = r->stats_batch;
        CR_TRY(ensure_mirrors(b));
        uint64_t hv = 0;
        for (uint32_t s = 0; s < b->n_shapes; ++s) hv += b->hull_count_host()[s];
        r->stats.vertex_bytes += 8ull * (hv - r->stats.hull_vertices);
        r->stats.hull_vertices = hv;
    }
    r->stats.kernel_launches = g_cr_kernel_launches;
    float ms = 0.0f;
    r->stats.last_tess_ms = (r->ev_valid[0] && cudaEventElapsedTime(&ms, r->ev[0], r->ev[1]) == cudaSuccess) ? ms : 0.0f;
    r->stats.last_bin_ms = (r->ev_valid[1] && cudaEventElapsedTime(&ms, r->ev[2], r->ev[3]) == cudaSuccess) ? ms : 0.0f;
    r->stats.last_raster_ms = (r->ev_valid[2] && cudaEventElapsedTime(&ms, r->ev[4], r->ev[5]) == cudaSuccess) ? ms : 0.0f;
    r->stats.last_hull_sort_ms = (r->ev_valid[0] && cudaEventElapsedTime(&ms, r->ev[6], r->ev[7]) == cudaSuccess) ? ms : 0.0f;
    r->stats.last_hull_chain_ms = (r->ev_valid[0] && cudaEventElapsedTime(&ms, r->ev[7], r->ev[8]) == cudaSuccess) ? ms : 0.0f;
    cudaGetLastError();
    *out = r->stats;
    return CR_OK;
}

}  // extern "C"
