// raster.cu — primitive setup, the binner and K3, the software stencil-then-cover tile rasteriser for sm_100a.
//
// Replaces the wgpu render pass of the reference: Shape::render draw recording (src/renderer.rs:267-355), the 13
// pipeline stencil / blend states (src/renderer.rs:565-861) and every entry point of src/shaders.wgsl.
//
// Data flow (all in HBM; nothing is read back before the pass has been enqueued — the host sizes everything from capacities and
// the kernels check them on the device, api.cu "optimistic submit"):
//   compact commands --expand--> commands + candidate numbering (one candidate per index slot / list triangle / hull-strip
//   triangle, in exact draw order) --prim_setup--> 64-byte PrimRecords (vertex stage done once) + tiles touched per candidate
//   (candidates with a big tile box: bin_big, one warp each; candidates that cross the eye plane: clip_kernel, fan triangles
//   behind the candidates) --scan--> --bin_emit--> (tile, candidate) pairs --stable radix sort by tile--> per-tile ranges in
//   draw order --tile_prims--> the tile-ordered 128-byte TilePrim stream (every pair set up for its tile once) --K3--> framebuffer.
//
// K3: one CTA per 16x16 tile. The pixel's stencil byte, RGBA colour (and depth) live in the REGISTERS of "its" thread for the
// whole pass, so the tile is read once and written once (128-bit accesses) and overdraw costs no HBM traffic. The tile's
// primitives arrive as one contiguous stream, 64 at a time, by bulk asynchronous copies (cp.async.bulk + mbarrier) into two
// shared-memory buffers — the next chunk is in flight while this one is rasterised. Inside a chunk, maximal runs of primitives
// whose stencil effect commutes are executed in parallel into a shared-memory accumulator:
//   * fill stencil (src/renderer.rs:577-582): the test only looks at the clip bits, the op is +-1 mod 2^winding_bits
//     => the run's net effect on a pixel is the signed count of covering front/back faces;
//   * stroke stencil (src/renderer.rs:571-576): passes only while the winding bits are still equal to the reference
//     (zero), then sets them to one => the run's net effect is "any primitive covers".
// Runs of at most PIXEL_RUN_MAX primitives (the usual hull cover, small fans) skip the shared accumulator altogether: every
// thread walks the run's primitives for its own pixel (no barrier, all threads busy); runs up to ROW_SWEEP_MAX lay (primitive,
// row) items out on a 16 x 16 grid; longer runs compact their rows with a per-warp prefix sum. Cover operations (colour, clip,
// alpha contexts) are order dependent and run one thread per pixel. Edge functions are evaluated in 32-bit integers whenever
// every value over the tile fits (exact), else in 64-bit. The pass's LoadOp::Clear is fused: cleared tiles start from zero
// instead of being loaded and every tile is written (empty ones too), so the target is neither memset nor read.
// One target on several GPUs: with tile sharding a CTA only processes tiles its rank owns and stores the finished tile into
// every rank's attachments (peer-mapped, P2P over NVLink); with draw-order sharding it waits for its predecessor's tile state
// and hands the tile to its successor (see RasterTarget in raster.h).
//
// Rasterisation contract: see the header comment of oracle/raster.hpp (written independently, same rules).
#include "device_common.cuh"
#include "prims.h"
#include "raster.h"

namespace {

#define RCHUNK 64             // primitives of a tile in shared memory at a time (two buffers: the next chunk streams in while this one is rasterised)
static_assert(RCHUNK == 64, "K3 scans a run's row counts two primitives per lane and keeps the run starts in one 64-bit mask");
#ifndef K3_MIN_BLOCKS
#define K3_MIN_BLOCKS 6   // resident K3 CTAs per SM asked of the compiler (40 registers per thread): 6 measured best of 4..8 on configs 3, 4 and 5
#endif
#ifndef K3_MIN_BLOCKS_MSAA
#define K3_MIN_BLOCKS_MSAA 3
#endif
#ifndef K3_MIN_BLOCKS_U8
#define K3_MIN_BLOCKS_U8 5   // 1x, 8-bit colour: the unorm conversions cost a few registers more
#endif
#define PIXEL_RUN_MAX 12u     // runs of at most this many primitives execute in pixel mode (see K3)
#define ROW_SWEEP_MAX 32u     // stencil runs up to this length use the fixed 16 x 16 (primitive, row) grid, longer ones compact their rows (measured: text 8 % slower, dashed strokes 21 % faster with compaction everywhere)
#define BIG_TILE_BOX 8        // candidates touching more tiles than this are binned by the whole warp

struct Descriptor {   // DynamicStrokeDescriptor, src/renderer.rs:18-27 (48 B)
    float gap_start[4];
    float gap_end[4];
    uint32_t caps;
    uint32_t count_dashed_join;
    float phase;
    uint32_t pad;
};

enum Pipe : uint32_t {
    P_STROKE_LINE = 0, P_STROKE_JOINT = 1, P_FILL_SOLID = 2, P_FILL_IQ = 3, P_FILL_IC = 4, P_FILL_RQ = 5, P_FILL_RC = 6,
    P_CLIP = 7, P_UNCLIP = 8, P_COLOR = 9, P_SAVE_ALPHA = 10, P_SCALE_ALPHA = 11, P_RESTORE_ALPHA = 12
};
#define META_FRONT 16u
#define META_SWAPPED 32u
#define META_VALID 64u
#define META_E32 4096u    // (tile-local) the edge functions fit 32-bit integers everywhere in the tile
#define META_FULL 2048u   // (tile-local) every pixel centre of the tile is inside the primitive
#define META_RUN_START 8192u   // (tile-local) this primitive does not commute with its predecessor in the tile: it opens a run
#define META_BIAS_SHIFT 16     // (tile-local) bit 16 + e: edge e is neither a top nor a left edge (bias -1)
#define META_KIND_SHIFT 19     // (tile-local) bits 19-20: run kind of the primitive's pipeline (0 stroke stencil, 1 fill stencil, 2 cover)
#define META_CLIPPED 16384u    // a fan triangle produced by frustum clipping: its corners' 1/w and attributes are in ClipAttr[v[0]]
#define META_CLIP_PARENT 32768u   // (record of a clipped candidate, not VALID) X[0] = first fan triangle, X[1] = their number

__device__ __forceinline__ const float* vertex_ptr(const DeviceBatch& b, uint32_t cat, uint32_t v) {
    switch (cat) {
        case 0: return reinterpret_cast<const float*>(b.vtx[0]) + (size_t)v * 5;
        case 1: return reinterpret_cast<const float*>(b.vtx[1]) + (size_t)v * 6;
        case 2: return reinterpret_cast<const float*>(b.vtx[2]) + (size_t)v * 2;
        case 3: return reinterpret_cast<const float*>(b.vtx[3]) + (size_t)v * 4;
        case 4: return reinterpret_cast<const float*>(b.vtx[4]) + (size_t)v * 5;
        case 5: return reinterpret_cast<const float*>(b.vtx[5]) + (size_t)v * 5;
        case 6: return reinterpret_cast<const float*>(b.vtx[6]) + (size_t)v * 6;
        default: return reinterpret_cast<const float*>(b.hull) + (size_t)v * 2;
    }
}

struct SnapVertex { int X, Y; bool ok; };
// vertex shader + viewport transform + 1/256 px snapping (src/shaders.wgsl:66-74)
__device__ __forceinline__ float clip_w(const float* __restrict__ m, float x, float y) { return (m[3] * x + m[7] * y) + m[15]; }
__device__ __forceinline__ float clip_z(const float* __restrict__ m, float x, float y) { return (m[2] * x + m[6] * y) + m[14]; }
// wgpu::CompareFunction of the depth test: does the fragment's depth `z` pass against the stored `d`?
__device__ __forceinline__ bool depth_passes(uint32_t f, float z, float d) {
    switch (f) {
        case CR_COMPARE_NEVER: return false;
        case CR_COMPARE_LESS: return z < d;
        case CR_COMPARE_EQUAL: return z == d;
        case CR_COMPARE_LESS_EQUAL: return z <= d;
        case CR_COMPARE_GREATER: return z > d;
        case CR_COMPARE_NOT_EQUAL: return z != d;
        case CR_COMPARE_GREATER_EQUAL: return z >= d;
        default: return true;
    }
}
// Storing a blend result into an 8-bit unorm attachment and reading it back: clamp, x * 255 + 0.5, truncate; texel / 255.
__device__ __forceinline__ float unorm8(float x) {
    x = x > 0.0f ? x : 0.0f;   // NaN -> 0
    x = x < 1.0f ? x : 1.0f;
    return cr::floor_f(x * 255.0f + 0.5f) / 255.0f;
}
__device__ __forceinline__ uint32_t unorm8_bits(float q) { return (uint32_t)(q * 255.0f + 0.5f); }   // q is k / 255 exactly rounded: recovers k
__device__ __forceinline__ SnapVertex snap_vertex(const float* __restrict__ m, float x, float y, uint32_t W, uint32_t H) {
    SnapVertex v;
    const float cx = (m[0] * x + m[4] * y) + m[12];
    const float cy = (m[1] * x + m[5] * y) + m[13];
    const float cw = clip_w(m, x, y);
    v.ok = cw > 0.0f;
    const float invw = 1.0f / cw;
    const float fx = ((cx * invw) * 0.5f + 0.5f) * (float)W;
    const float fy = (0.5f - (cy * invw) * 0.5f) * (float)H;
    if (!(cr::fabs_f(fx) <= 2097152.0f) || !(cr::fabs_f(fy) <= 2097152.0f)) v.ok = false;
    v.X = v.ok ? (int)cr::floor_f(fx * 256.0f + 0.5f) : 0;
    v.Y = v.ok ? (int)cr::floor_f(fy * 256.0f + 0.5f) : 0;
    return v;
}

// Edge e runs from vertex e to vertex e+1 of a clockwise (y down) triangle: E_e(P) = A_e (P.y - Y_e) - B_e (P.x - X_e),
// inside when E_e + bias_e >= 0 (top-left rule: bias 0 on top and left edges, -1 otherwise).
struct Edges { int A[3], B[3], bias[3]; };
__device__ __forceinline__ Edges make_edges(const int* X, const int* Y) {
    Edges t;
#pragma unroll
    for (int e = 0; e < 3; ++e) {
        const int n = e == 2 ? 0 : e + 1;
        t.A[e] = X[n] - X[e];
        t.B[e] = Y[n] - Y[e];
        t.bias[e] = ((t.B[e] == 0 && t.A[e] > 0) || t.B[e] < 0) ? 0 : -1;
    }
    return t;
}

// Pixel and tile extent of a triangle: pixel p has its samples at 256 p + [sample_lo, sample_hi] (128 = the centre for 1x).
struct Extent { int px0, px1, py0, py1, tx0, tx1, ty0, ty1; bool empty; };
__device__ __forceinline__ Extent extent_of(const int* X, const int* Y, const RasterTarget& tg) {
    Extent x;
    const int minX = min(X[0], min(X[1], X[2])), maxX = max(X[0], max(X[1], X[2]));
    const int minY = min(Y[0], min(Y[1], Y[2])), maxY = max(Y[0], max(Y[1], Y[2]));
    x.px0 = max(0, (minX - tg.sample_hi + 255) >> 8); x.px1 = min((int)tg.width - 1, (maxX - tg.sample_lo) >> 8);
    x.py0 = max(0, (minY - tg.sample_hi + 255) >> 8); x.py1 = min((int)tg.height - 1, (maxY - tg.sample_lo) >> 8);
    x.empty = x.px0 > x.px1 || x.py0 > x.py1;
    x.tx0 = x.px0 / CR_TILE; x.tx1 = x.px1 / CR_TILE; x.ty0 = x.py0 / CR_TILE; x.ty1 = x.py1 / CR_TILE;
    return x;
}
// Can any sample of tile (tx, ty), restricted to the triangle's pixel extent, be inside? (per-edge trivial reject)
__device__ __forceinline__ bool tile_hit(const int* X, const int* Y, const Edges& t, const Extent& x, int tx, int ty, const RasterTarget& tg) {
    const int ylo = max(x.py0, ty * CR_TILE) * 256 + tg.sample_lo, yhi = min(x.py1, ty * CR_TILE + CR_TILE - 1) * 256 + tg.sample_hi;
    const int xlo = max(x.px0, tx * CR_TILE) * 256 + tg.sample_lo, xhi = min(x.px1, tx * CR_TILE + CR_TILE - 1) * 256 + tg.sample_hi;
    bool hit = true;
#pragma unroll
    for (int e = 0; e < 3; ++e) {
        const int PX = t.B[e] < 0 ? xhi : xlo, PY = t.A[e] > 0 ? yhi : ylo;   // corner maximising E_e
        const long long E = (long long)t.A[e] * (PY - Y[e]) - (long long)t.B[e] * (PX - X[e]);
        if (E + t.bias[e] < 0) hit = false;
    }
    return hit;
}

// ------------------------------------------------------------------------------------------ vertex stage
// Which command a candidate belongs to: last command whose first candidate is <= cand.
__device__ __forceinline__ uint32_t find_command(const uint32_t* __restrict__ begin, uint32_t n_commands, uint32_t cand) {
    uint32_t lo = 0, hi = n_commands;
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (begin[mid] <= cand) lo = mid; else hi = mid; }
    return lo;
}

// What a candidate number stands for: the command, the vertex category, the instance and the three vertices of its triangle.
struct CandInfo { uint32_t ci, cat, instance, v[3]; bool odd; };
// `hint`: a command at or before the candidate's (0xFFFFFFFF: none, binary search); walking forwards from it is one or two steps
// for the consecutive candidates of a CTA.
__device__ bool candidate_info(const RasterScene& sc, uint32_t cand, const uint32_t* __restrict__ cmd_begin, CandInfo& c, uint32_t hint = 0xFFFFFFFFu) {
    if (hint == 0xFFFFFFFFu) c.ci = find_command(cmd_begin, sc.n_commands, cand);
    else { c.ci = hint; while (c.ci + 1u < sc.n_commands && cmd_begin[c.ci + 1u] <= cand) ++c.ci; }
    const DeviceCommand& cmd = sc.commands[c.ci];
    uint32_t rem = cand - cmd_begin[c.ci];
    uint32_t cat = 0, prev_end = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { const uint32_t e = cmd.cat_end[k]; if (rem >= e) { cat = k + 1; prev_end = e; } }
    if (cat > 7) return false;
    rem -= prev_end;
    const uint32_t n = cmd.slots[cat];
    const uint32_t inst = rem / n, local = rem - inst * n;
    const DeviceBatch& b = sc.batches[cmd.batch];
    c.odd = false;
    if (cat <= 2) {   // indexed triangle strips with primitive restart (src/renderer.rs:476)
        if (local + 2 >= n) return false;
        const uint32_t* idx = b.idx[cat] + cmd.ibase[cat] + local;
        const uint32_t i0 = idx[0], i1 = idx[1], i2 = idx[2];
        if (i0 == CR_RESTART || i1 == CR_RESTART || i2 == CR_RESTART) return false;
        c.v[0] = cmd.vbase[cat] + (i0 >> 1); c.v[1] = cmd.vbase[cat] + (i1 >> 1); c.v[2] = cmd.vbase[cat] + (i2 >> 1);
        c.odd = (i0 & 1u) != 0;
    } else if (cat <= 6) {   // triangle lists
        c.v[0] = cmd.vbase[cat] + 3 * local; c.v[1] = c.v[0] + 1; c.v[2] = c.v[0] + 2;
    } else {   // non-indexed hull strip (src/renderer.rs:354)
        c.v[0] = cmd.vbase[7] + local; c.v[1] = c.v[0] + 1; c.v[2] = c.v[0] + 2;
        c.odd = (local & 1u) != 0;
    }
    c.cat = cat;
    c.instance = cmd.instance_begin + inst;
    return true;
}
// Orientation, culling and the record of one snapped triangle. false: zero area or culled.
__device__ __forceinline__ bool finish_record(const RasterTarget& tg, const DeviceCommand& cmd, const CandInfo& c, const SnapVertex* sv, PrimRecord& rec) {
    const long long area2 = (long long)(sv[1].X - sv[0].X) * (sv[2].Y - sv[0].Y) - (long long)(sv[2].X - sv[0].X) * (sv[1].Y - sv[0].Y);
    if (area2 == 0) return false;
    const bool front = (area2 < 0) != c.odd;   // counter-clockwise in NDC (y up) = negative area in y-down pixels; odd strip triangles flip
    const bool swapped = area2 < 0;
    if (cmd.operation == CR_OP_COLOR) {   // cull_mode applies to the colour cover only (src/renderer.rs:743)
        if (tg.cull_mode == CR_CULL_BACK && !front) return false;
        if (tg.cull_mode == CR_CULL_FRONT && front) return false;
    }
    const uint32_t pipe = cmd.operation == CR_OP_STENCIL ? c.cat : P_CLIP + (cmd.operation - CR_OP_CLIP);   // CLIP..RESTORE follow the enum order
    rec.X[0] = sv[0].X; rec.Y[0] = sv[0].Y;
    rec.X[1] = swapped ? sv[2].X : sv[1].X; rec.Y[1] = swapped ? sv[2].Y : sv[1].Y;
    rec.X[2] = swapped ? sv[1].X : sv[2].X; rec.Y[2] = swapped ? sv[1].Y : sv[2].Y;
    rec.meta = pipe | (front ? META_FRONT : 0u) | (swapped ? META_SWAPPED : 0u) | META_VALID | (c.cat << 8);
    rec.cmd = c.ci;
    rec.instance = c.instance;
    rec.v[0] = c.v[0]; rec.v[1] = c.v[1]; rec.v[2] = c.v[2];
    rec.ref = cmd.ref; rec.layers = cmd.layers; rec.batch = cmd.batch; rec._pad = 0;
    return true;
}
// 0: nothing to draw, 1: `rec` is the candidate's record, 2: a vertex cannot be snapped (eye plane / range): frustum clipping decides
__device__ int build_record(const RasterScene& sc, const RasterTarget& tg, uint32_t cand, const uint32_t* __restrict__ cmd_begin, PrimRecord& rec, uint32_t hint) {
    rec.meta = 0;
    CandInfo c;
    if (!candidate_info(sc, cand, cmd_begin, c, hint)) return 0;
    const DeviceCommand& cmd = sc.commands[c.ci];
    const DeviceBatch& b = sc.batches[cmd.batch];
    const float* m = sc.transforms + 16 * (size_t)c.instance;
    SnapVertex sv[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float* p = vertex_ptr(b, c.cat, c.v[i]);   // 20- and 24-byte vertices are only 4-byte aligned
        sv[i] = snap_vertex(m, p[0], p[1], tg.width, tg.height);
    }
    if (!sv[0].ok || !sv[1].ok || !sv[2].ok) return 2;
    return finish_record(tg, cmd, c, sv, rec) ? 1 : 0;
}

__device__ __forceinline__ void store_record(PrimRecord* dst, const PrimRecord& rec) {
    uint4* d = reinterpret_cast<uint4*>(dst);
    d[0] = make_uint4((uint32_t)rec.X[0], (uint32_t)rec.X[1], (uint32_t)rec.X[2], (uint32_t)rec.Y[0]);
    d[1] = make_uint4((uint32_t)rec.Y[1], (uint32_t)rec.Y[2], rec.meta, rec.cmd);
    d[2] = make_uint4(rec.instance, rec.v[0], rec.v[1], rec.v[2]);
    d[3] = make_uint4(rec.ref, rec.layers, rec.batch, 0u);
}
template <bool FULL>   // FULL == false: only the geometry half (X, Y, meta, cmd) that the binner needs
__device__ __forceinline__ PrimRecord load_record(const PrimRecord* src) {
    const uint4* s = reinterpret_cast<const uint4*>(src);
    const uint4 a = s[0], b = s[1];
    PrimRecord r;
    r.X[0] = (int)a.x; r.X[1] = (int)a.y; r.X[2] = (int)a.z; r.Y[0] = (int)a.w;
    r.Y[1] = (int)b.x; r.Y[2] = (int)b.y; r.meta = b.z; r.cmd = b.w;
    if (FULL) {
        const uint4 c = s[2], d = s[3];
        r.instance = c.x; r.v[0] = c.y; r.v[1] = c.z; r.v[2] = c.w;
        r.ref = d.x; r.layers = d.y; r.batch = d.z;
    }
    return r;
}

// Tiles touched by a triangle whose tile box is small: walked by its own thread.
// EMIT == false: returns the count. EMIT == true: also writes (tile, cand) pairs starting at `at`.
template <bool EMIT>
__device__ __forceinline__ uint32_t walk_tiles_small(const int* X, const int* Y, const Extent& x, const RasterTarget& tg, uint32_t cand, uint32_t at,
                                                     uint32_t* __restrict__ pair_tile, uint32_t* __restrict__ pair_cand) {
    const Edges t = make_edges(X, Y);
    uint32_t count = 0;
    for (int ty = x.ty0; ty <= x.ty1; ++ty)
        for (int tx = x.tx0; tx <= x.tx1; ++tx)
            if (cr_tile_owned(tg, tx, ty) && tile_hit(X, Y, t, x, tx, ty, tg)) {
                if (EMIT) { pair_tile[at + count] = (uint32_t)(ty * (int)tg.tiles_x + tx); pair_cand[at + count] = cand; }
                ++count;
            }
    return count;
}
// Tiles touched by a triangle with a large tile box (hull covers, big fans): walked by a whole warp, 32 tiles per step.
template <bool EMIT>
__device__ __forceinline__ uint32_t walk_tiles_warp(const int* X, const int* Y, const RasterTarget& tg, uint32_t cand, uint32_t at,
                                                    uint32_t* __restrict__ pair_tile, uint32_t* __restrict__ pair_cand) {
    const uint32_t lane = threadIdx.x & 31u;
    const Extent x = extent_of(X, Y, tg);
    const Edges t = make_edges(X, Y);
    const int w = x.tx1 - x.tx0 + 1, total = w * (x.ty1 - x.ty0 + 1);
    uint32_t running = 0;
    int ty = x.ty0 + (int)lane / w, tx = x.tx0 + (int)lane % w;   // 32 consecutive tiles of the box, advanced incrementally
    const int dy = 32 / w, dx = 32 % w;
    for (int base = 0; base < total; base += 32) {
        const bool hit = base + (int)lane < total && cr_tile_owned(tg, tx, ty) && tile_hit(X, Y, t, x, tx, ty, tg);
        const uint32_t hits = __ballot_sync(0xffffffffu, hit);
        if (EMIT && hit) {
            const uint32_t pos = at + running + __popc(hits & ((1u << lane) - 1u));
            pair_tile[pos] = (uint32_t)(ty * (int)tg.tiles_x + tx);
            pair_cand[pos] = cand;
        }
        running += __popc(hits);
        ty += dy; tx += dx;
        if (tx > x.tx1) { tx -= w; ty += 1; }
    }
    return running;
}

// ------------------------------------------------------------------------------------------ command expansion
// Live sizes of the pass as every kernel derives them from the device counters and the host's capacities.
__device__ __forceinline__ uint32_t live_candidates(const PassCounters* c, uint32_t cand_capacity) {
    const unsigned long long t = c->cand_total;
    return t <= (unsigned long long)cand_capacity ? (uint32_t)t : 0u;
}
__device__ __forceinline__ bool pairs_fit(const PassCounters* c, uint32_t pair_capacity) {
    return c->pair_total <= (unsigned long long)pair_capacity && c->pair_total < 0xFFFFFFFFull;
}

// Candidates of one command, category by category in the draw order of src/renderer.rs:275-354, from the batch's slice tables.
__device__ __forceinline__ unsigned long long expand_command(const CompactCommand& cc, const DeviceBatch* __restrict__ batches, DeviceCommand& c) {
    const DeviceBatch& b = batches[cc.batch];
    const size_t stride = (size_t)b.n_shapes + 1;
    const uint32_t* cb = b.cat_begin;
    const uint32_t shape = cc.shape;
    c.batch = cc.batch;
    c.instance_begin = cc.instance_begin;
    c.instance_count = cc.instance_count;
    c.operation = cc.operation;
    c.ref = cc.ref;
    c.layers = cc.layers;
    c._pad[0] = c._pad[1] = c._pad[2] = 0;
    unsigned long long total = 0;
#pragma unroll
    for (int cat = 0; cat < 8; ++cat) {
        const bool drawn = cc.operation == CR_OP_STENCIL ? (cat < 7 && (cat >= 2 || b.n_groups > 0)) : cat == 7;
        uint32_t slots = 0;
        if (drawn) {
            if (cat <= 2) slots = cb[(CNT_LINE_IDX + cat) * stride + shape + 1] - cb[(CNT_LINE_IDX + cat) * stride + shape];
            else if (cat <= 6) slots = (cb[cat * stride + shape + 1] - cb[cat * stride + shape]) / 3u;
            else { const uint32_t hc = b.hull_count[shape]; slots = hc >= 3 ? hc - 2 : 0u; }
        }
        c.slots[cat] = slots;
        total += (unsigned long long)slots * cc.instance_count;
        c.cat_end[cat] = (uint32_t)(total < 0xFFFFFFFFull ? total : 0xFFFFFFFFull);
        c.vbase[cat] = cb[(cat < 7 ? cat : (int)CNT_PROTO) * stride + shape];
        if (cat < 3) c.ibase[cat] = cb[(CNT_LINE_IDX + cat) * stride + shape];
    }
    return total;
}
// Up to CR_EXPAND_FUSED_MAX commands: one CTA expands them and scans their candidate counts (chunks of 1024 with a carry).
__global__ void __launch_bounds__(1024) expand_scan_kernel(const CompactCommand* __restrict__ compact, uint32_t n_commands, const DeviceBatch* __restrict__ batches,
                                                           DeviceCommand* __restrict__ commands, uint32_t* __restrict__ cmd_cand_begin, PassCounters* __restrict__ counters) {
    __shared__ unsigned long long sh_warp[32];
    __shared__ unsigned long long sh_carry;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) sh_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_commands; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        unsigned long long count = 0;
        if (i < n_commands) {
            DeviceCommand c;
            count = expand_command(compact[i], batches, c);
            commands[i] = c;
        }
        unsigned long long incl = count;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned long long y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (uint32_t)o) incl += y; }
        if (lane == 31) sh_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = sh_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned long long y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= (uint32_t)o) w += y; }
            sh_warp[lane] = w;
        }
        __syncthreads();
        const unsigned long long excl = sh_carry + (warp ? sh_warp[warp - 1] : 0ull) + (incl - count);
        if (i < n_commands) cmd_cand_begin[i] = (uint32_t)(excl < 0xFFFFFFFFull ? excl : 0xFFFFFFFFull);
        __syncthreads();
        if (threadIdx.x == 1023) sh_carry = excl + count;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const unsigned long long total = sh_carry;
        cmd_cand_begin[n_commands] = (uint32_t)(total < 0xFFFFFFFFull ? total : 0xFFFFFFFFull);
        counters->cand_total = total;
        counters->pair_total = 0;
        counters->covered = 0;
        counters->n_pairs_live = 0;
        counters->flags = 0;
        counters->clip_total = 0;
    }
}
// More commands than that: one thread per command writes its count into cmd_cand_begin (scanned afterwards by cr_scan_exclusive).
__global__ void __launch_bounds__(256) expand_kernel(const CompactCommand* __restrict__ compact, uint32_t n_commands, const DeviceBatch* __restrict__ batches,
                                                     DeviceCommand* __restrict__ commands, uint32_t* __restrict__ cmd_cand_begin, PassCounters* __restrict__ counters) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long count = 0;
    if (i < n_commands) {
        DeviceCommand c;
        count = expand_command(compact[i], batches, c);
        commands[i] = c;
        cmd_cand_begin[i] = (uint32_t)(count < 0xFFFFFFFFull ? count : 0xFFFFFFFFull);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) count += __shfl_xor_sync(0xffffffffu, count, o);
    if ((threadIdx.x & 31u) == 0 && count) atomicAdd(&counters->cand_total, count);
}

#define SETUP_THREADS 256
#ifndef SETUP_MIN_BLOCKS
#define SETUP_MIN_BLOCKS 6   // 40 registers: bin stage 0.295 -> 0.283 ms on the text scene
#endif
#define META_BIG 128u
// big[0] = number of big candidates, big[1 ...] = their candidate numbers
__global__ void __launch_bounds__(SETUP_THREADS, SETUP_MIN_BLOCKS) prim_setup_kernel(RasterScene sc, RasterTarget tg, uint32_t cand_capacity, PrimRecord* __restrict__ records,
                                                                   uint32_t* __restrict__ cand_tiles, uint32_t* __restrict__ big, uint32_t* __restrict__ clip_list,
                                                                   PassCounters* __restrict__ counters) {
    const uint32_t n = live_candidates(counters, cand_capacity);   // 0 when the capacity does not suffice: nothing is produced, the host re-submits
    if (blockIdx.x == 0 && threadIdx.x == 0 && n == 0 && counters->cand_total != 0ull) atomicOr(&counters->flags, CR_PASS_OVERFLOW_CANDS);
    // the command of the CTA's first candidate, found once by a warp-wide 32-way search; every thread walks forwards from it
    // (a CTA's 256 consecutive candidates lie in one or two commands unless the commands are tiny)
    __shared__ uint32_t sh_first_command;
    const uint32_t* cmd_begin = sc.cmd_cand_begin;
    if (threadIdx.x < 32u) {
        const uint32_t first = warp_search_last_le(cmd_begin, sc.n_commands, min(blockIdx.x * blockDim.x, n ? n - 1u : 0u), threadIdx.x);
        if (threadIdx.x == 0) sh_first_command = first;
    }
    __syncthreads();
    const uint32_t cand = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t count = 0;
    if (cand < n) {
        PrimRecord rec;
        const int built = build_record(sc, tg, cand, cmd_begin, rec, sh_first_command);
        if (built == 2) clip_list[1 + atomicAdd(clip_list, 1u)] = cand;   // clipped, counted and binned by clip_kernel
        if (built == 1) {
            const Extent x = extent_of(rec.X, rec.Y, tg);
            if (x.empty) rec.meta = 0;
            else if ((x.tx1 - x.tx0 + 1) * (x.ty1 - x.ty0 + 1) <= BIG_TILE_BOX) count = walk_tiles_small<false>(rec.X, rec.Y, x, tg, cand, 0, nullptr, nullptr);
            else { rec.meta |= META_BIG; big[1 + atomicAdd(big, 1u)] = cand; }   // counted by bin_big_kernel<false>
        }
        if (rec.meta & META_VALID) store_record(records + cand, rec);   // nobody reads the record of a candidate without tiles (more than half of them: restarts, degenerate and culled triangles)
    }
    if (cand < cand_capacity) cand_tiles[cand] = count;   // zero beyond the live count: the scan runs over the capacity
    // 64-bit total of the (tile, candidate) pairs: the scan that places them is 32-bit, so the host must see an overflow
    // (33 k full-target hull covers at 8K wrap it) instead of sizing the pair arrays from a wrapped count
    const uint32_t warp_pairs = __reduce_add_sync(0xffffffffu, count);
    if ((threadIdx.x & 31u) == 0 && warp_pairs) atomicAdd(&counters->pair_total, (unsigned long long)warp_pairs);
}

__global__ void __launch_bounds__(SETUP_THREADS) bin_emit_kernel(RasterTarget tg, uint32_t cand_capacity, uint32_t pair_capacity, const PrimRecord* __restrict__ records,
                                                                 const uint32_t* __restrict__ begin, uint32_t* __restrict__ pair_tile,
                                                                 uint32_t* __restrict__ pair_cand, PassCounters* __restrict__ counters) {
    const uint32_t n = live_candidates(counters, cand_capacity);
    const bool fit = pairs_fit(counters, pair_capacity);
    if (blockIdx.x == 0 && threadIdx.x == 0) {   // the later kernels of the pass (sort, range search, tile kernel) read these
        counters->n_pairs_live = (fit && n != 0u) ? (uint32_t)counters->pair_total : 0u;
        if (!fit) atomicOr(&counters->flags, CR_PASS_OVERFLOW_PAIRS);
    }
    if (!fit) return;
    const uint32_t cand = blockIdx.x * blockDim.x + threadIdx.x;
    if (cand >= n) return;
    const uint32_t at = begin[cand];
    if (begin[cand + 1] == at) return;
    const PrimRecord rec = load_record<false>(records + cand);
    if (!(rec.meta & META_VALID) || (rec.meta & META_BIG)) return;
    walk_tiles_small<true>(rec.X, rec.Y, extent_of(rec.X, rec.Y, tg), tg, cand, at, pair_tile, pair_cand);
}

// One warp per big candidate (grid-stride over the list built by prim_setup_kernel).
template <bool EMIT>
__global__ void __launch_bounds__(128) bin_big_kernel(RasterTarget tg, const PrimRecord* __restrict__ records, const uint32_t* __restrict__ big,
                                                      uint32_t* __restrict__ cand_tiles, uint32_t* __restrict__ pair_tile, uint32_t* __restrict__ pair_cand,
                                                      PassCounters* __restrict__ counters, uint32_t pair_capacity) {
    if (EMIT && !pairs_fit(counters, pair_capacity)) return;
    const uint32_t n_big = big[0];
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n_big; i += warps) {
        const uint32_t cand = big[1 + i];
        const PrimRecord rec = load_record<false>(records + cand);
        if (EMIT) walk_tiles_warp<true>(rec.X, rec.Y, tg, cand, cand_tiles[cand], pair_tile, pair_cand);   // cand_tiles now holds the exclusive scan
        else {
            const uint32_t count = walk_tiles_warp<false>(rec.X, rec.Y, tg, cand, 0, nullptr, nullptr);
            if ((threadIdx.x & 31u) == 0) { cand_tiles[cand] = count; if (count) atomicAdd(&counters->pair_total, (unsigned long long)count); }
        }
    }
}

// ------------------------------------------------------------------------------------------ frustum clipping
// (the rule is stated once, in words, in raster.h; oracle/raster.hpp implements the same words independently)
struct ClipVertex { float x, y, z, w; float attr[4]; };
__device__ __forceinline__ float clip_distance(const ClipVertex& v, int plane, float G) {
    switch (plane) {
        case 0: return v.w - CR_CLIP_W_MIN;
        case 1: return G * v.w - v.x;
        case 2: return G * v.w + v.x;
        case 3: return G * v.w - v.y;
        default: return G * v.w + v.y;
    }
}
// The point where the edge from `in` (distance din >= 0) to `out` (distance dout < 0) meets the plane.
__device__ __forceinline__ ClipVertex clip_intersection(const ClipVertex& in, const ClipVertex& out, float din, float dout) {
    const float t = din / (din - dout);
    ClipVertex r;
    r.x = in.x + (out.x - in.x) * t; r.y = in.y + (out.y - in.y) * t; r.z = in.z + (out.z - in.z) * t; r.w = in.w + (out.w - in.w) * t;
#pragma unroll
    for (int k = 0; k < 4; ++k) r.attr[k] = in.attr[k] + (out.attr[k] - in.attr[k]) * t;
    return r;
}
// Sutherland-Hodgman over the five planes; poly holds n corners (capacity 9). Returns the number of corners left (0: nothing).
__device__ int clip_polygon(ClipVertex* poly, int n, float G) {
    ClipVertex tmp[9];
    for (int plane = 0; plane < 5; ++plane) {
        int m = 0;
        for (int i = 0; i < n; ++i) {
            const ClipVertex& cur = poly[i];
            const ClipVertex& nxt = poly[i + 1 == n ? 0 : i + 1];
            const float dc = clip_distance(cur, plane, G), dn = clip_distance(nxt, plane, G);
            const bool cin = dc >= 0.0f, nin = dn >= 0.0f;
            if (cin) tmp[m++] = cur;
            if (cin != nin) tmp[m++] = cin ? clip_intersection(cur, nxt, dc, dn) : clip_intersection(nxt, cur, dn, dc);
        }
        n = m;
        for (int i = 0; i < n; ++i) poly[i] = tmp[i];
        if (n < 3) return 0;
    }
    return n;
}
// Viewport transform + snapping of a clip-space corner (same arithmetic as snap_vertex).
__device__ __forceinline__ SnapVertex snap_clip_vertex(const ClipVertex& c, uint32_t W, uint32_t H) {
    SnapVertex v;
    v.ok = c.w > 0.0f;
    const float invw = 1.0f / c.w;
    const float fx = ((c.x * invw) * 0.5f + 0.5f) * (float)W;
    const float fy = (0.5f - (c.y * invw) * 0.5f) * (float)H;
    if (!(cr::fabs_f(fx) <= 2097152.0f) || !(cr::fabs_f(fy) <= 2097152.0f)) v.ok = false;
    v.X = v.ok ? (int)cr::floor_f(fx * 256.0f + 0.5f) : 0;
    v.Y = v.ok ? (int)cr::floor_f(fy * 256.0f + 0.5f) : 0;
    return v;
}

// One thread per candidate of the clip list. EMIT == false: clips, writes the fan triangles' records (behind the candidates:
// records[cand_capacity + k]) and corner attributes, counts their tiles into cand_tiles[cand] and leaves a parent record.
// EMIT == true (after the scan): writes the fan triangles' (tile, record) pairs into the candidate's pair range.
template <bool EMIT>
__global__ void __launch_bounds__(128) clip_kernel(RasterScene sc, RasterTarget tg, uint32_t cand_capacity, uint32_t clip_capacity, PrimRecord* __restrict__ records,
                                                   ClipAttr* __restrict__ clip_attrs, const uint32_t* __restrict__ clip_list, uint32_t* __restrict__ cand_tiles,
                                                   uint32_t* __restrict__ pair_tile, uint32_t* __restrict__ pair_cand, PassCounters* __restrict__ counters,
                                                   uint32_t pair_capacity) {
    const uint32_t n_clip = min(clip_list[0], cand_capacity);
    if (EMIT && (!pairs_fit(counters, pair_capacity) || (counters->flags & CR_PASS_OVERFLOW_CLIP) != 0u)) return;
    const float G = cr_guard_band(tg.width, tg.height);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_clip; i += gridDim.x * blockDim.x) {
        const uint32_t cand = clip_list[1 + i];
        if (EMIT) {
            const PrimRecord parent = load_record<false>(records + cand);
            if (!(parent.meta & META_CLIP_PARENT)) continue;
            uint32_t at = cand_tiles[cand];   // the exclusive scan by now
            for (int j = 0; j < parent.X[1]; ++j) {
                const uint32_t slot = cand_capacity + (uint32_t)parent.X[0] + (uint32_t)j;
                const PrimRecord rec = load_record<false>(records + slot);
                const Extent x = extent_of(rec.X, rec.Y, tg);
                if (!x.empty) at += walk_tiles_small<true>(rec.X, rec.Y, x, tg, slot, at, pair_tile, pair_cand);
            }
            continue;
        }
        store_record(records + cand, PrimRecord{});   // whatever an earlier pass left there: not a parent unless the end of this iteration says so
        CandInfo c;
        if (!candidate_info(sc, cand, sc.cmd_cand_begin, c)) continue;
        const DeviceCommand& cmd = sc.commands[c.ci];
        const DeviceBatch& b = sc.batches[cmd.batch];
        const float* m = sc.transforms + 16 * (size_t)c.instance;
        const int n_attr = (int)((0x04332032u >> (4u * c.cat)) & 15u);   // attribute floats per category: 2,3,0,2,3,3,4,0
        ClipVertex poly[9];
        uint32_t flat_u = 0;
        float flat_f = 0.0f;
        for (int k = 0; k < 3; ++k) {
            const float* p = vertex_ptr(b, c.cat, c.v[k]);
            poly[k].x = (m[0] * p[0] + m[4] * p[1]) + m[12];
            poly[k].y = (m[1] * p[0] + m[5] * p[1]) + m[13];
            poly[k].z = clip_z(m, p[0], p[1]);
            poly[k].w = clip_w(m, p[0], p[1]);
            for (int a = 0; a < 4; ++a) poly[k].attr[a] = a < n_attr ? p[2 + a] : 0.0f;
            if (k == 0 && c.cat <= 1) { flat_u = __float_as_uint(p[2 + n_attr]); flat_f = p[3]; }   // flat attributes: the first vertex of the ORIGINAL triangle
        }
        const int n = clip_polygon(poly, 3, G);
        if (n < 3) continue;
        PrimRecord sub[CR_CLIP_MAX_TRIANGLES];
        ClipAttr att[CR_CLIP_MAX_TRIANGLES];
        uint32_t n_sub = 0;
        const bool depth_attr = tg.depth != nullptr && cmd.operation == CR_OP_COLOR;
        for (int j = 1; j + 1 < n && n_sub < CR_CLIP_MAX_TRIANGLES; ++j) {
            const ClipVertex* corner[3] = {&poly[0], &poly[j], &poly[j + 1]};
            SnapVertex sv[3];
            for (int k = 0; k < 3; ++k) sv[k] = snap_clip_vertex(*corner[k], tg.width, tg.height);
            if (!sv[0].ok || !sv[1].ok || !sv[2].ok) continue;
            PrimRecord rec;
            if (!finish_record(tg, cmd, c, sv, rec)) continue;
            rec.meta |= META_CLIPPED;
            ClipAttr& a = att[n_sub];
            for (int k = 0; k < 3; ++k) {
                a.invw[k] = 1.0f / corner[k]->w;
                for (int q = 0; q < 4; ++q) a.attr[k][q] = corner[k]->attr[q];
                if (depth_attr) a.attr[k][0] = corner[k]->z / corner[k]->w;
            }
            if (c.cat <= 1) a.attr[1][3] = flat_f;
            a.flat_u = flat_u;
            sub[n_sub++] = rec;
        }
        if (n_sub == 0) continue;
        const uint32_t base = atomicAdd(&counters->clip_total, n_sub);
        if (base + n_sub > clip_capacity) { atomicOr(&counters->flags, CR_PASS_OVERFLOW_CLIP); continue; }
        uint32_t tiles = 0;
        for (uint32_t j = 0; j < n_sub; ++j) {
            sub[j].v[0] = base + j;   // where its corners are
            store_record(records + cand_capacity + base + j, sub[j]);
            clip_attrs[base + j] = att[j];
            const Extent x = extent_of(sub[j].X, sub[j].Y, tg);
            if (!x.empty) tiles += walk_tiles_small<false>(sub[j].X, sub[j].Y, x, tg, 0u, 0u, nullptr, nullptr);
        }
        PrimRecord parent{};
        parent.meta = META_CLIP_PARENT;
        parent.X[0] = (int)base; parent.X[1] = (int)n_sub;
        store_record(records + cand, parent);
        cand_tiles[cand] = tiles;
        if (tiles) atomicAdd(&counters->pair_total, (unsigned long long)tiles);
    }
}

// ------------------------------------------------------------------------------ draw-order sharding: exchange
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t order_slot(const RasterTarget& tg, uint32_t rank) { return rank - (rank > tg.order_rank ? 1u : 0u); }

// One thread per bitmap word: bit t = this rank's slice has primitives in tile 32 w + t. Written into every rank's table.
__global__ void touched_tiles_kernel(RasterTarget tg, const uint32_t* __restrict__ tile_begin) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= tg.order_mask_words) return;
    const uint32_t n_tiles = tg.tiles_x * tg.tiles_y;
    uint32_t bits = 0;
    for (uint32_t t = 0; t < 32u; ++t) {
        const uint32_t tile = 32u * w + t;
        if (tile < n_tiles && tile_begin[tile + 1] > tile_begin[tile]) bits |= 1u << t;
    }
    const size_t at = cr_exchange_mask_offset(n_tiles) + (size_t)tg.order_rank * tg.order_mask_words + w;
    tg.exchange[at] = bits;
    for (uint32_t peer = 0; peer + 1 < tg.order_world; ++peer) tg.peer_exchange[peer][at] = bits;
}
// Runs after touched_tiles_kernel has completed: the bitmap of this rank is in place everywhere, say so (release, system scope).
__global__ void touched_ready_kernel(RasterTarget tg) {
    __threadfence_system();
    if (threadIdx.x == 0) st_release_sys(&tg.exchange[tg.order_rank], tg.order_epoch);
    else if (threadIdx.x < tg.order_world) st_release_sys(&tg.peer_exchange[threadIdx.x - 1][tg.order_rank], tg.order_epoch);
}

// ------------------------------------------------------------------------------------------- K3: tile raster
struct TilePrim {          // 128 bytes: one primitive set up for one tile (a record of the tile-ordered stream K3 bulk-loads)
    long long e0[3];        // edge functions, top-left bias folded in, at the centre of the tile's pixel (0, 0)
    int A[3], B[3];         // per-pixel steps: +256 A per row, -256 B per column
    float invw[3];
    uint32_t meta;          // PrimRecord::meta + the tile-local flags; META_VALID cleared if nothing of it can land in this tile
    float attr[3][4];       // stroke vertices: attr[0][3] holds their flat u32 (bits), attr[1][3] their flat float
    uint32_t box_mask;      // the bounding box clipped to the tile as bit masks: bits x0..x1 | (bits y0..y1) << 16 — "is my pixel in the box" is one AND + compare
    uint32_t ref_batch;     // stencil reference (8 bits) | batch << 8
    uint32_t instance, layers;
};
static_assert(sizeof(TilePrim) == 128, "TilePrim streams are moved with 16-byte bulk copies");
__device__ __forceinline__ int box_x0(uint32_t m) { return __ffs((int)(m & 0xFFFFu)) - 1; }
__device__ __forceinline__ int box_x1(uint32_t m) { return 31 - __clz((int)(m & 0xFFFFu)); }
__device__ __forceinline__ int box_y0(uint32_t m) { return __ffs((int)(m >> 16)) - 1; }
__device__ __forceinline__ int box_y1(uint32_t m) { return 31 - __clz((int)(m >> 16)); }
__device__ __forceinline__ int prim_bias(uint32_t meta, int e) { return -(int)((meta >> (META_BIAS_SHIFT + e)) & 1u); }   // top-left bias of edge e: 0 or -1
__device__ __forceinline__ uint32_t prim_ref(const TilePrim& ps) { return ps.ref_batch & 255u; }
__device__ __forceinline__ uint32_t prim_batch(const TilePrim& ps) { return ps.ref_batch >> 8; }
__device__ __forceinline__ uint32_t prim_flat_u(const TilePrim& ps) { return __float_as_uint(ps.attr[0][3]); }

__device__ __forceinline__ bool cap_test(float tx, float ty, uint32_t cap_type) {   // src/shaders.wgsl:165-189
    switch (cap_type & 15u) {
        case 0: return ty > 0.5f;
        case 1: return tx * tx + ty * ty < 0.25f;
        case 2: return 0.5f - ty > cr::fabs_f(tx);
        case 3: return ty < cr::fabs_f(tx);
        case 4: return 0.5f - ty > tx;
        case 5: return ty - 0.5f < tx;
        default: return ty < 0.0f;
    }
}
__device__ bool stroke_dashed(const Descriptor& d, float tx, float ty) {   // src/shaders.wgsl:205-231
    const uint32_t last = d.count_dashed_join >> 3;
    const float pattern_length = d.gap_end[last & 3u];
    uint32_t interval = 0;
    float gap_start, gap_end;
    float pos = cr::wgsl_mod(ty - d.phase, pattern_length);
    if (pos < 0.0f) pos = pos + pattern_length;
    for (;;) {
        gap_end = d.gap_end[interval & 3u] - pos;
        if (gap_end >= 0.0f || interval >= last) break;
        interval = interval + 1u;
    }
    gap_start = pos - d.gap_start[interval & 3u];
    if (gap_start > 0.0f) {
        const uint32_t caps = d.caps >> (interval * 8u);
        const bool start_cap = cap_test(tx, gap_start, caps >> 4u);
        const bool end_cap = cap_test(tx, gap_end, caps);
        return start_cap || end_cap;
    }
    return true;
}

// Fragment stage of the stencil pipelines (src/shaders.wgsl:233-300): perspective-correct attributes at the sample and
// the sample_mask predicate. E[] are the biased edge values at the sample.
template <typename EdgeT>   // long long, or int for primitives whose edge values fit 32 bits over the whole tile (META_E32)
__device__ __forceinline__ bool fragment_keep(const RasterScene& sc, const TilePrim& ps, uint32_t pipe, const EdgeT* E) {
    const uint32_t meta = ps.meta;
    const float e0 = (float)(E[1] - prim_bias(meta, 1)) * ps.invw[0], e1 = (float)(E[2] - prim_bias(meta, 2)) * ps.invw[1], e2 = (float)(E[0] - prim_bias(meta, 0)) * ps.invw[2];
    const float den = (e0 + e1) + e2;
    float a[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    const int needed = (pipe == P_FILL_IQ || pipe == P_STROKE_LINE) ? 2 : (pipe == P_FILL_RC ? 4 : 3);   // attributes the predicate reads
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (q < needed) a[q] = ((e0 * ps.attr[0][q] + e1 * ps.attr[1][q]) + e2 * ps.attr[2][q]) / den;
    switch (pipe) {
        case P_FILL_IQ: return a[0] * a[0] - a[1] <= 0.0f;
        case P_FILL_IC: return a[0] * a[0] * a[0] - a[1] * a[2] <= 0.0f;
        case P_FILL_RQ: return a[0] * a[0] - a[1] * a[2] <= 0.0f;
        case P_FILL_RC: return a[0] * a[0] * a[0] - a[1] * a[2] * a[3] <= 0.0f;
        case P_STROKE_LINE: {   // src/shaders.wgsl:268-285
            const uint32_t flat_u = prim_flat_u(ps);
            const Descriptor& d = reinterpret_cast<const Descriptor*>(sc.batches[prim_batch(ps)].stroke)[flat_u & 65535u];
            if ((d.count_dashed_join & 4u) != 0u) return stroke_dashed(d, a[0], a[1]);
            if ((flat_u & 65536u) != 0u) return cap_test(a[0], a[1] - ps.attr[1][3], d.caps >> 4u);
            if (a[1] < 0.0f) return cap_test(a[0], -a[1], d.caps);
            return true;
        }
        default: {              // P_STROKE_JOINT, src/shaders.wgsl:287-300
            const uint32_t flat_u = prim_flat_u(ps);
            const Descriptor& d = reinterpret_cast<const Descriptor*>(sc.batches[prim_batch(ps)].stroke)[flat_u & 65535u];
            const float radius = cr::sqrt_f(a[0] * a[0] + a[1] * a[1]);
            const uint32_t kind = d.count_dashed_join & 3u;
            bool keep = kind == 1u ? (flat_u & 65536u) != 0u : (kind == 2u ? radius <= 0.5f : true);
            if (keep && (d.count_dashed_join & 4u) != 0u) keep = stroke_dashed(d, radius, a[2] + cr::atan2_f(a[1], a[0]) / 6.28318548202514648438f);
            return keep;
        }
    }
}

// Sample positions inside a pixel in 1/256 px: the centre for 1x, the WebGPU standard pattern for 4x (same table as the oracle).
template <int S> __device__ __forceinline__ int sample_x(int k) { return S == 1 ? 128 : (k == 0 ? 96 : (k == 1 ? 224 : (k == 2 ? 32 : 160))); }
template <int S> __device__ __forceinline__ int sample_y(int k) { return S == 1 ? 128 : (k == 0 ? 32 : (k == 1 ? 96 : (k == 2 ? 160 : 224))); }

// Edge values of primitive `ps` at pixel (lx, ly) of the tile (evaluation origin of stage_primitive), in EdgeT arithmetic.
template <typename EdgeT> __device__ __forceinline__ EdgeT edge_origin(const TilePrim& ps, int e) {
    return sizeof(EdgeT) == 8 ? (EdgeT)ps.e0[e] : (EdgeT)reinterpret_cast<const int*>(&ps.e0[e])[0];   // the low word IS the value when it fits
}
// Samples of pixel (lx, ly) inside the triangle (bit q), optionally also passing the pipeline's fragment predicate.
template <int S, typename EdgeT>
__device__ __forceinline__ uint32_t pixel_hits(const RasterScene& sc, const TilePrim& ps, uint32_t pipe, bool predicate, int lx, int ly) {
    EdgeT E[3];
#pragma unroll
    for (int e = 0; e < 3; ++e) E[e] = edge_origin<EdgeT>(ps, e) + (EdgeT)ps.A[e] * (EdgeT)(ly * 256) - (EdgeT)ps.B[e] * (EdgeT)(lx * 256);
    uint32_t hit = 0;
#pragma unroll
    for (int q = 0; q < S; ++q) {
        EdgeT Es[3];
#pragma unroll
        for (int e = 0; e < 3; ++e) Es[e] = S == 1 ? E[e] : E[e] + (EdgeT)ps.A[e] * (EdgeT)sample_y<S>(q) - (EdgeT)ps.B[e] * (EdgeT)sample_x<S>(q);
        if ((Es[0] | Es[1] | Es[2]) >= 0 && (!predicate || fragment_keep(sc, ps, pipe, Es))) hit |= 1u << q;
    }
    return hit;
}
// One (primitive, row) item of a stencil run in row mode: accumulate into `out` for the pixels x0..x1 of row y.
template <int S, typename EdgeT>
__device__ __forceinline__ void stencil_row(const RasterScene& sc, const TilePrim& ps, uint32_t pipe, uint32_t kind, int delta, int y, int x0, int x1, int* out) {
    EdgeT E[3], step[3];
#pragma unroll
    for (int e = 0; e < 3; ++e) {
        step[e] = (EdgeT)ps.B[e] * (EdgeT)256;
        E[e] = edge_origin<EdgeT>(ps, e) + (EdgeT)ps.A[e] * (EdgeT)(y * 256) - step[e] * (EdgeT)x0;
    }
    for (int x = x0; x <= x1; ++x) {
#pragma unroll
        for (int q = 0; q < S; ++q) {
            EdgeT Es[3];
#pragma unroll
            for (int e = 0; e < 3; ++e) Es[e] = S == 1 ? E[e] : E[e] + (EdgeT)ps.A[e] * (EdgeT)sample_y<S>(q) - (EdgeT)ps.B[e] * (EdgeT)sample_x<S>(q);
            if ((Es[0] | Es[1] | Es[2]) >= 0) {
                if (pipe == P_FILL_SOLID || fragment_keep(sc, ps, pipe, Es)) {
                    if (kind == 0u) atomicOr(&out[(y * CR_TILE + x) * S + q], 1);
                    else atomicAdd(&out[(y * CR_TILE + x) * S + q], delta);
                }
            }
        }
        E[0] -= step[0]; E[1] -= step[1]; E[2] -= step[2];
    }
}
// Coverage mask (bit x * S + q) of row y of a cover primitive, pixels x0..x1.
template <int S, typename EdgeT>
__device__ __forceinline__ unsigned long long cover_row_mask(const TilePrim& ps, int y, int x0, int x1) {
    EdgeT E[3], step[3];
#pragma unroll
    for (int e = 0; e < 3; ++e) {
        step[e] = (EdgeT)ps.B[e] * (EdgeT)256;
        E[e] = edge_origin<EdgeT>(ps, e) + (EdgeT)ps.A[e] * (EdgeT)(y * 256) - step[e] * (EdgeT)x0;
    }
    unsigned long long mask = 0;
    for (int x = x0; x <= x1; ++x) {
#pragma unroll
        for (int q = 0; q < S; ++q) {
            EdgeT any = 0;
#pragma unroll
            for (int e = 0; e < 3; ++e) any |= S == 1 ? E[e] : E[e] + (EdgeT)ps.A[e] * (EdgeT)sample_y<S>(q) - (EdgeT)ps.B[e] * (EdgeT)sample_x<S>(q);
            if (any >= 0) mask |= 1ull << (x * S + q);
        }
        E[0] -= step[0]; E[1] -= step[1]; E[2] -= step[2];
    }
    return mask;
}

// Pixel bounding box of a primitive clipped to the tile at (tile_px, tile_py) and to the target; false if it is empty.
template <int S>
__device__ __forceinline__ bool tile_bbox(const RasterTarget& tg, const PrimRecord& rec, int tile_px, int tile_py, uint32_t& bbox) {
    const int minX = min(rec.X[0], min(rec.X[1], rec.X[2])), maxX = max(rec.X[0], max(rec.X[1], rec.X[2]));
    const int minY = min(rec.Y[0], min(rec.Y[1], rec.Y[2])), maxY = max(rec.Y[0], max(rec.Y[1], rec.Y[2]));
    const int slo = S == 1 ? 128 : 32, shi = S == 1 ? 128 : 224;   // sample offsets inside a pixel span [slo, shi]
    const int x0 = max(0, ((minX - shi + 255) >> 8) - tile_px), x1 = min(min(CR_TILE - 1, (int)tg.width - 1 - tile_px), ((maxX - slo) >> 8) - tile_px);
    const int y0 = max(0, ((minY - shi + 255) >> 8) - tile_py), y1 = min(min(CR_TILE - 1, (int)tg.height - 1 - tile_py), ((maxY - slo) >> 8) - tile_py);
    if (x0 > x1 || y0 > y1) { bbox = 0; return false; }
    bbox = (((2u << x1) - 1u) & ~((1u << x0) - 1u)) | ((((2u << y1) - 1u) & ~((1u << y0) - 1u)) << 16);   // TilePrim::box_mask
    return true;
}

// Set one primitive up for one tile (one thread per (tile, primitive) pair). The edge functions are evaluated at the
// centre of the tile's pixel (0, 0) for 1x and at its top-left corner for 4x (sample offsets are added per sample).
template <int S, bool DEPTH>
__device__ void stage_primitive(const RasterScene& sc, const RasterTarget& tg, const PrimRecord& rec, int tile_px, int tile_py, TilePrim& ps) {
    ps.meta = 0;
    ps.box_mask = 0; ps.ref_batch = 0; ps.instance = 0; ps.layers = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        ps.e0[i] = 0; ps.A[i] = 0; ps.B[i] = 0; ps.invw[i] = 0.0f;
#pragma unroll
        for (int a = 0; a < 4; ++a) ps.attr[i][a] = 0.0f;
    }
    if (!(rec.meta & META_VALID)) return;
    uint32_t bbox;
    if (!tile_bbox<S>(tg, rec, tile_px, tile_py, bbox)) return;
    const uint32_t pipe = rec.meta & 15u, cat = (rec.meta >> 8) & 7u;
    const Edges t = make_edges(rec.X, rec.Y);
    const int PX = tile_px * 256 + (S == 1 ? 128 : 0), PY = tile_py * 256 + (S == 1 ? 128 : 0);
    uint32_t bias_bits = 0;
#pragma unroll
    for (int e = 0; e < 3; ++e) {
        ps.A[e] = t.A[e]; ps.B[e] = t.B[e];
        if (t.bias[e]) bias_bits |= 1u << (META_BIAS_SHIFT + e);
        ps.e0[e] = (long long)t.A[e] * (PY - rec.Y[e]) - (long long)t.B[e] * (PX - rec.X[e]) + t.bias[e];
    }
    ps.box_mask = bbox;
    bool full = true;   // minimum of every edge function over all sample positions of the tile is still inside
    const int far = (CR_TILE - 1) * 256 + (S == 1 ? 0 : 224), near = S == 1 ? 0 : 32;   // sample offsets from the evaluation origin
#pragma unroll
    for (int e = 0; e < 3; ++e) {
        const long long emin = ps.e0[e] + (long long)t.A[e] * (t.A[e] < 0 ? far : near) - (long long)t.B[e] * (t.B[e] > 0 ? far : near);
        if (emin < 0) full = false;
    }
    ps.ref_batch = (rec.ref & 255u) | (rec.batch << 8);
    ps.instance = rec.instance;
    ps.layers = rec.layers;
    if (rec.meta & META_CLIPPED) {   // a fan triangle of frustum clipping: its corners carry interpolated 1 / w and attributes (and z / w)
        const ClipAttr* ca = sc.clip_attrs + rec.v[0];
        const bool swapped = (rec.meta & META_SWAPPED) != 0;
        const float4* q = reinterpret_cast<const float4*>(ca);   // invw[3] attr[3][4] flat_u = 16 words
        const float4 q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3];
        const float w16[16] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z, q3.w};
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int src = i == 0 ? 0 : (swapped ? 3 - i : i);
            ps.invw[i] = w16[src];
#pragma unroll
            for (int a = 0; a < 4; ++a) ps.attr[i][a] = w16[3 + 4 * src + a];
        }
        if (cat <= 1) { ps.attr[0][3] = w16[15]; ps.attr[1][3] = w16[3 + 4 * 1 + 3]; }   // flat u32 (bits) and flat float, where no stroke predicate interpolates
    } else if (pipe <= P_FILL_RC && pipe != P_FILL_SOLID) {   // pipelines with a fragment predicate need the vertex attributes
        const DeviceBatch& b = sc.batches[rec.batch];
        const float* m = sc.transforms + 16 * (size_t)rec.instance;
        const int n_attr = (int)((0x04332032u >> (4u * cat)) & 15u);   // attribute floats per category: 2,3,0,2,3,3,4,0
        const bool swapped = (rec.meta & META_SWAPPED) != 0;
        float flat_f = 0.0f;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int src = i == 0 ? 0 : (swapped ? 3 - i : i);   // stored vertex order is (0, 2, 1) when swapped
            const float* p = vertex_ptr(b, cat, rec.v[src]);
            ps.invw[i] = 1.0f / clip_w(m, p[0], p[1]);
#pragma unroll
            for (int a = 0; a < 4; ++a) ps.attr[i][a] = a < n_attr ? p[2 + a] : 0.0f;
            // flat attributes come from the first (provoking) vertex: the stroke vertices' flat u32 goes into attr[0][3], their
            // flat float (the first vertex's second attribute) into attr[1][3] — no stroke predicate interpolates a fourth attribute
            if (i == 0 && cat <= 1) { ps.attr[0][3] = p[2 + n_attr]; flat_f = p[3]; }
        }
        if (cat <= 1) ps.attr[1][3] = flat_f;
    }
    if (DEPTH && pipe == P_COLOR && !(rec.meta & META_CLIPPED)) {   // the depth test of the colour cover needs z / w of the three hull vertices (src/shaders.wgsl:72)
        const DeviceBatch& b = sc.batches[rec.batch];
        const float* m = sc.transforms + 16 * (size_t)rec.instance;
        const bool swapped = (rec.meta & META_SWAPPED) != 0;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int src = i == 0 ? 0 : (swapped ? 3 - i : i);
            const float* p = vertex_ptr(b, 7, rec.v[src]);
            ps.attr[i][0] = clip_z(m, p[0], p[1]) / clip_w(m, p[0], p[1]);
        }
    }
    // 32-bit edge arithmetic is exact for this tile if every edge value at every sample position of the tile (and +-1 for the
    // bias that fragment_keep removes again) stays inside int range
    bool fits = true;
#pragma unroll
    for (int e = 0; e < 3; ++e) {
        const long long reach = ((long long)abs(t.A[e]) + (long long)abs(t.B[e])) * (15 * 256 + 256) + 2;
        if (ps.e0[e] > 0x7fffffffLL - reach || ps.e0[e] < -0x7fffffffLL + reach) fits = false;
    }
    ps.meta = rec.meta | (full ? META_FULL : 0u) | (fits ? META_E32 : 0u) | bias_bits;
}

// Run kinds: primitives of one run commute (see the file header).
__device__ __forceinline__ uint32_t run_kind(uint32_t pipe) { return pipe <= P_STROKE_JOINT ? 0u : (pipe <= P_FILL_RC ? 1u : 2u); }

// ---- tile-ordered primitive streams
// After the sort the (tile, candidate) pairs of a tile are contiguous and in draw order. One thread per PAIR does the whole
// tile-local set-up of its primitive (edge functions at the tile origin, clipped bounding box, attribute fetch, the "covers the
// whole tile" and "fits 32 bits" flags, and whether it opens a run) and writes the 128-byte TilePrim at the pair's position: the
// tile kernel then reads each tile's primitives as ONE contiguous stream with bulk asynchronous copies (cp.async.bulk, the TMA
// engine) into shared memory, double buffered, instead of gathering records and staging them itself.
template <int S, bool DEPTH>
__global__ void __launch_bounds__(256) tile_prims_kernel(RasterScene sc, RasterTarget tg, const PrimRecord* __restrict__ records, const uint32_t* __restrict__ pair_tile,
                                                         const uint32_t* __restrict__ pair_cand, uint32_t pair_capacity, TilePrim* __restrict__ out,
                                                         const PassCounters* __restrict__ counters) {
    __shared__ uint4 sh_out[8][32 * 8];
    if (counters->flags != 0u) return;   // a capacity did not suffice: the pairs are not there
    const uint32_t n = min(counters->n_pairs_live, pair_capacity);
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if ((i & ~31u) >= n) return;   // whole warps leave together (shuffles below)
    const bool live = i < n;
    TilePrim ps;
    ps.meta = 0;
    uint32_t tile = 0xFFFFFFFFu, ref = 0, cmd = 0;
    if (live) {
        tile = pair_tile[i];
        const PrimRecord rec = load_record<true>(records + pair_cand[i]);
        stage_primitive<S, DEPTH>(sc, tg, rec, (int)(tile % tg.tiles_x) * CR_TILE, (int)(tile / tg.tiles_x) * CR_TILE, ps);
        ref = rec.ref; cmd = rec.cmd;
    }
    // does this primitive open a run? It needs its predecessor in the tile: the neighbouring lane's, or (lane 0) looked up here
    uint32_t p_meta = __shfl_up_sync(0xffffffffu, ps.meta, 1), p_ref = __shfl_up_sync(0xffffffffu, ref, 1);
    uint32_t p_cmd = __shfl_up_sync(0xffffffffu, cmd, 1), p_inst = __shfl_up_sync(0xffffffffu, ps.instance, 1), p_tile = __shfl_up_sync(0xffffffffu, tile, 1);
    if (lane == 0) {
        p_tile = 0xFFFFFFFEu;
        if (live && i > 0 && pair_tile[i - 1] == tile) {
            const PrimRecord prec = load_record<true>(records + pair_cand[i - 1]);
            uint32_t bbox;
            const bool valid = (prec.meta & META_VALID) != 0 && tile_bbox<S>(tg, prec, (int)(tile % tg.tiles_x) * CR_TILE, (int)(tile / tg.tiles_x) * CR_TILE, bbox);
            p_tile = tile; p_meta = valid ? prec.meta : 0u; p_ref = prec.ref; p_cmd = prec.cmd; p_inst = prec.instance;
        }
    }
    bool boundary = true;   // the first primitive of a tile
    if (p_tile == tile) {
        // staged-out primitives (meta == 0) join whatever run precedes them: they do nothing
        const uint32_t kp = run_kind(p_meta & 15u), kq = run_kind(ps.meta & 15u);
        const bool pv = (p_meta & META_VALID) != 0, qv = (ps.meta & META_VALID) != 0;
        if (pv && qv) boundary = kp != kq || (kq < 2u ? p_ref != ref : (p_cmd != cmd || p_inst != ps.instance));
        else boundary = qv;   // a valid primitive after a staged-out one conservatively opens a run
    }
    if (boundary) ps.meta |= META_RUN_START;
    if (ps.meta & META_VALID) ps.meta |= run_kind(ps.meta & 15u) << META_KIND_SHIFT;
    // The warp's 32 records leave through shared memory so that every store instruction writes 512 contiguous bytes (a record
    // per thread would be 16 bytes every 128). Chunk k of lane l sits at slot 8 l + (k ^ (l & 7)): conflict free both ways.
    uint4 q[8];
    q[0] = make_uint4((uint32_t)ps.e0[0], (uint32_t)((unsigned long long)ps.e0[0] >> 32), (uint32_t)ps.e0[1], (uint32_t)((unsigned long long)ps.e0[1] >> 32));
    q[1] = make_uint4((uint32_t)ps.e0[2], (uint32_t)((unsigned long long)ps.e0[2] >> 32), (uint32_t)ps.A[0], (uint32_t)ps.A[1]);
    q[2] = make_uint4((uint32_t)ps.A[2], (uint32_t)ps.B[0], (uint32_t)ps.B[1], (uint32_t)ps.B[2]);
    q[3] = make_uint4(__float_as_uint(ps.invw[0]), __float_as_uint(ps.invw[1]), __float_as_uint(ps.invw[2]), ps.meta);
#pragma unroll
    for (int k = 0; k < 3; ++k) q[4 + k] = make_uint4(__float_as_uint(ps.attr[k][0]), __float_as_uint(ps.attr[k][1]), __float_as_uint(ps.attr[k][2]), __float_as_uint(ps.attr[k][3]));
    q[7] = make_uint4(ps.box_mask, ps.ref_batch, ps.instance, ps.layers);
    uint4* const mine = sh_out[threadIdx.x >> 5];
#pragma unroll
    for (int k = 0; k < 8; ++k) mine[8u * lane + ((uint32_t)k ^ (lane & 7u))] = q[k];
    __syncwarp();
    const uint32_t warp_first = i - lane, n_out = min(32u, n - warp_first) * 8u;   // 16-byte pieces this warp owns
    uint4* const dst = reinterpret_cast<uint4*>(out + warp_first);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t piece = (uint32_t)j * 32u + lane, rec = piece >> 3, chunk = piece & 7u;
        if (piece < n_out) dst[piece] = mine[8u * rec + (chunk ^ (rec & 7u))];
    }
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0u;
}
// global -> shared bulk asynchronous copy (TMA engine, UBLKCP in SASS); completion is counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <int S, bool DEPTH, bool U8>   // samples per pixel; depth test / write on the colour cover; 8-bit unorm colour + R8 alpha layers
__global__ void __launch_bounds__(CR_TILE * CR_TILE, (S == 1 && !DEPTH) ? (U8 ? K3_MIN_BLOCKS_U8 : K3_MIN_BLOCKS) : K3_MIN_BLOCKS_MSAA) raster_tiles_kernel(RasterScene sc, RasterTarget tg, const TilePrim* __restrict__ prims,
                                                                                         const uint32_t* __restrict__ tile_begin,
                                                                                         PassCounters* __restrict__ counters) {
    __shared__ __align__(128) TilePrim sh_buf[2][RCHUNK];   // the tile's primitive stream, RCHUNK at a time: chunk c + 1 streams in (bulk copy) while chunk c is rasterised
    __shared__ __align__(8) unsigned long long sh_bar[2];   // one mbarrier per buffer: completes when the chunk's bytes have landed
    __shared__ int acc[2][CR_TILE * CR_TILE * S];   // per-sample result of a stencil run; double buffered so that one barrier per run suffices
    __shared__ uint32_t run_rows[CR_TILE * CR_TILE / 32][RCHUNK];   // per warp: prefix sums of the rows of a stencil run's primitives
    __shared__ unsigned long long cov[2][CR_TILE][CR_TILE];   // row coverage masks (bit x * S + k) of the cover primitives of one sweep; double buffered: one barrier per sweep
    const uint32_t tile = blockIdx.x;
    const uint32_t begin = tile_begin[tile], end = tile_begin[tile + 1];   // (three independent loads: one round trip)
    if (counters != nullptr && counters->flags != 0u) return;   // a capacity did not suffice: leave the attachments untouched, the host re-submits
    const bool clears = tg.clear_color != 0u || tg.clear_stencil != 0u;
    // ---- draw-order sharding: where is this rank in the tile's chain?
    __shared__ uint32_t sh_chain[3];     // predecessor rank + 1 (0: none), successor rank + 1 (0: none), some rank touches the tile
    bool continue_chain = false;         // start from the predecessor's tile state (in our attachments) instead of the pass's load / clear
    uint32_t successor = 0;              // rank + 1 the finished tile state goes to; 0: this rank finishes the tile
    if (tg.order_world > 1u) {
        if (threadIdx.x == 0) {
            for (uint32_t r = 0; r < tg.order_world; ++r)
                while (ld_acquire_sys(&tg.exchange[r]) != tg.order_epoch) { }   // every rank's bitmap of this pass has arrived
            const uint32_t n_tiles = tg.tiles_x * tg.tiles_y;
            const uint32_t* masks = tg.exchange + cr_exchange_mask_offset(n_tiles);
            uint32_t pred = 0, succ = 0, any = 0;
            for (uint32_t r = 0; r < tg.order_world; ++r) {
                const uint32_t touched = (__ldcg(&masks[(size_t)r * tg.order_mask_words + (tile >> 5)]) >> (tile & 31u)) & 1u;
                if (!touched) continue;
                any = 1u;
                if (r < tg.order_rank) pred = r + 1u;
                if (r > tg.order_rank && succ == 0u) succ = r + 1u;
            }
            if (begin != end && pred != 0u)
                while (ld_acquire_sys(&tg.exchange[cr_exchange_flag_offset() + tile]) != tg.order_epoch) { }   // the predecessor has delivered the tile
            sh_chain[0] = pred; sh_chain[1] = succ; sh_chain[2] = any;
        }
        __syncthreads();
        if (begin == end) {
            if (sh_chain[2] != 0u || !clears) return;   // another rank's slice draws here (the last one of the chain writes our copy too), or nothing to clear
        } else {
            continue_chain = sh_chain[0] != 0u;
            successor = sh_chain[1];
        }
    } else if (begin == end && (!clears || !cr_tile_owned(tg, (int)(tile % tg.tiles_x), (int)(tile / tg.tiles_x)))) return;   // nothing drawn here and nothing to clear (or not ours to clear)
    // ---- the primitive stream of this tile: prims[begin, end), RCHUNK per bulk copy; issue the first chunk now, it lands while the tile is loaded
    const uint32_t n_chunks = (end - begin + RCHUNK - 1) / RCHUNK;
    auto issue_chunk = [&](uint32_t c) {   // one thread
        const uint32_t first = begin + c * RCHUNK, bytes = min((uint32_t)RCHUNK, end - first) * (uint32_t)sizeof(TilePrim);
        mbar_expect_tx(&sh_bar[c & 1u], bytes);
        bulk_copy_g2s(&sh_buf[c & 1u][0], prims + first, bytes, &sh_bar[c & 1u]);
    };
    if (threadIdx.x == 0 && n_chunks != 0u) {
        mbar_init(&sh_bar[0], 1u);
        mbar_init(&sh_bar[1], 1u);
        fence_mbar_init();
        issue_chunk(0u);
        if (n_chunks > 1u) issue_chunk(1u);
    }
    const int tile_px = (int)(tile % tg.tiles_x) * CR_TILE, tile_py = (int)(tile / tg.tiles_x) * CR_TILE;
    const int lx = threadIdx.x & (CR_TILE - 1), ly = threadIdx.x / CR_TILE;
    const int px = tile_px + lx, py = tile_py + ly;
    const bool in_fb = px < (int)tg.width && py < (int)tg.height;
    const size_t pix = ((size_t)py * tg.width + px) * S;   // first sample of this thread's pixel
    uint32_t s[S];
    float4 col[S];
    float dep[DEPTH ? S : 1];
#pragma unroll
    for (int k = 0; k < S; ++k) { s[k] = 0; col[k] = make_float4(0.f, 0.f, 0.f, 0.f); }
    if (DEPTH) {
#pragma unroll
        for (int k = 0; k < S; ++k) dep[k] = (tg.clear_depth != 0u || !in_fb) ? tg.depth_clear_value : tg.depth[pix + k];
    }
    if (in_fb) {   // LoadOp::Load reads the attachment, LoadOp::Clear starts from zero (and the tile is written in any case)
        if (tg.clear_stencil == 0u || continue_chain) {
            if (S == 1) s[0] = __ldcg(tg.stencil + pix);
            else {
                const uint32_t packed = __ldcg(reinterpret_cast<const uint32_t*>(tg.stencil + pix));   // 4 samples = 4 bytes, 4-byte aligned
#pragma unroll
                for (int k = 0; k < S; ++k) s[k] = (packed >> (8 * k)) & 255u;
            }
        }
        if (tg.clear_color == 0u || continue_chain) {
#pragma unroll
            for (int k = 0; k < S; ++k) {
                if (U8) {
                    const uint32_t t = __ldcg(reinterpret_cast<const uint32_t*>(tg.color) + pix + k);
                    const float c0 = (float)(t & 255u) / 255.0f, c1 = (float)((t >> 8) & 255u) / 255.0f, c2 = (float)((t >> 16) & 255u) / 255.0f;
                    col[k] = tg.color_format == CR_FORMAT_BGRA8_UNORM ? make_float4(c2, c1, c0, (float)(t >> 24) / 255.0f) : make_float4(c0, c1, c2, (float)(t >> 24) / 255.0f);
                } else col[k] = __ldcg(reinterpret_cast<const float4*>(tg.color) + pix + k);
            }
        }
    }
    const uint32_t W = tg.wmask, C = tg.cmask, M = W | C;
    uint32_t covered = 0;
    const size_t layer_stride = (size_t)tg.width * tg.height * S;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < S; ++k) { acc[0][threadIdx.x * S + k] = 0; acc[1][threadIdx.x * S + k] = 0; }
    // A finished stencil run whose per-sample result has not been folded into `s` yet (block-uniform state).
    int pending = -1;            // accumulator buffer of the pending run, or -1
    uint32_t pending_kind = 0, pending_ref = 0;
    int cur = 0;                 // buffer the next stencil run accumulates into
    int cb = 0;                  // row-mask buffer the next cover sweep writes
    auto apply_pending = [&]() {
        if (pending < 0) return;
#pragma unroll
        for (int k = 0; k < S; ++k) {
            const int net = acc[pending][threadIdx.x * S + k];
            if (net != 0) {
                if (pending_kind == 0u) { if ((pending_ref & M) == (s[k] & M)) s[k] = (s[k] & ~W) | ((s[k] + 1u) & W); }        // src/renderer.rs:571-576
                else if ((pending_ref & M) <= (s[k] & M)) s[k] = (s[k] & ~W) | ((s[k] + (uint32_t)net) & W);                     // src/renderer.rs:577-582
                acc[pending][threadIdx.x * S + k] = 0;   // ready for the run after next (a barrier separates)
            }
        }
        pending = -1;
    };

    // One cover primitive on the samples `hit` of this thread's pixel: stencil test / op and blend of its pipeline.
    auto cover_apply = [&](const TilePrim& ps, uint32_t hit) {
        const uint32_t pipe = ps.meta & 15u, ref = prim_ref(ps);
#pragma unroll
        for (int q = 0; q < S; ++q) {
            if (!((hit >> q) & 1u)) continue;
            if (pipe == P_COLOR) {                                                                   // src/renderer.rs:736-754
                if ((ref & M) < (s[q] & M)) {
                    if (DEPTH) {
                        // depth of the fragment: z / w of the vertices interpolated linearly in screen space (the rasteriser's rule for
                        // depth), weights = the unbiased edge values at the sample; a depth failure KEEPS the stencil value
                        // (depth_fail_op, src/renderer.rs:442) and writes nothing
                        long long E[3];
#pragma unroll
                        for (int e = 0; e < 3; ++e)
                            E[e] = ps.e0[e] - prim_bias(ps.meta, e) + (long long)ps.A[e] * (ly * 256 + (S == 1 ? 0 : sample_y<S>(q))) - (long long)ps.B[e] * (lx * 256 + (S == 1 ? 0 : sample_x<S>(q)));
                        const float b0 = (float)E[1], b1 = (float)E[2], b2 = (float)E[0];
                        const float z = ((b0 * ps.attr[0][0] + b1 * ps.attr[1][0]) + b2 * ps.attr[2][0]) / ((b0 + b1) + b2);
                        if (!depth_passes(tg.depth_compare, z, dep[q])) continue;
                        if (tg.depth_write) dep[q] = z;
                    }
                    const float4 ic = reinterpret_cast<const float4*>(sc.colors)[ps.instance];
                    const float sa = ic.w;
                    const float sr = ic.x * sa, sg = ic.y * sa, sb = ic.z * sa;
                    if (tg.blending == CR_BLEND_PREMULTIPLIED_OVER) {
                        const float kk = 1.0f - sa;
                        col[q].x = sr + col[q].x * kk; col[q].y = sg + col[q].y * kk; col[q].z = sb + col[q].z * kk; col[q].w = sa + col[q].w * kk;
                    } else { col[q].x = sr; col[q].y = sg; col[q].z = sb; col[q].w = sa; }
                    if (U8) { col[q].x = unorm8(col[q].x); col[q].y = unorm8(col[q].y); col[q].z = unorm8(col[q].z); col[q].w = unorm8(col[q].w); }
                    covered += 1;
                }
                s[q] = s[q] & ~W;
            } else if (pipe == P_CLIP) {                                                             // src/renderer.rs:692-710
                if ((ref & W) != (s[q] & W)) s[q] = (s[q] & ~M) | (ref & M);
            } else if (pipe == P_UNCLIP) {                                                           // src/renderer.rs:711-729
                if ((ref & C) < (s[q] & C)) s[q] = (s[q] & ~M) | (ref & M);
            } else if ((ref & M) <= (s[q] & M)) {   // the three alpha-context covers share one stencil state (src/renderer.rs:761-766)
                if (pipe == P_SAVE_ALPHA) {   // the frame's alpha into the layer: f32, or R8Unorm like the reference's (src/renderer.rs:783,898)
                    const size_t at = (size_t)(ps.layers & 65535u) * layer_stride + pix + q;
                    if (U8) reinterpret_cast<uint8_t*>(tg.alpha_layers)[at] = (uint8_t)unorm8_bits(unorm8(col[q].w));
                    else reinterpret_cast<float*>(tg.alpha_layers)[at] = col[q].w;
                } else if (pipe == P_SCALE_ALPHA) {
                    const float sa = 1.0f - reinterpret_cast<const float4*>(sc.colors)[ps.instance].w;
                    col[q].w = sa + col[q].w * (1.0f - sa);
                    if (U8) col[q].w = unorm8(col[q].w);
                } else {
                    const size_t at = (size_t)(ps.layers >> 16) * layer_stride + pix + q;
                    const float saved = U8 ? (float)reinterpret_cast<const uint8_t*>(tg.alpha_layers)[at] / 255.0f : reinterpret_cast<const float*>(tg.alpha_layers)[at];
                    const float sa = (1.0f - saved) * (1.0f - reinterpret_cast<const float4*>(sc.colors)[ps.instance].w);
                    col[q].w = col[q].w - sa;
                    if (U8) col[q].w = unorm8(col[q].w);
                }
            }
        }
    };
    // PIXEL MODE for short runs: every thread walks the run's primitives for its own pixel — no shared accumulator, no
    // barrier, all 256 threads busy however few primitives the run has.
    const uint32_t my_box_bits = (1u << lx) | (1u << (16 + ly));
    auto pixel_in_box = [&](const TilePrim& ps) -> bool {
        return (ps.box_mask & my_box_bits) == my_box_bits;
    };

    if (n_chunks != 0u) __syncthreads();   // the mbarriers are initialised before anybody waits on them
    for (uint32_t c = 0; c < n_chunks; ++c) {
        const uint32_t n = min((uint32_t)RCHUNK, end - (begin + c * RCHUNK));
        const TilePrim* const sh = sh_buf[c & 1u];
        if (c >= 1u && c + 1u < n_chunks) {   // chunk c + 1 goes where chunk c - 1 was: everybody must be done with that one
            __syncthreads();
            if (threadIdx.x == 0) issue_chunk(c + 1u);
        }
        while (!mbar_try_wait(&sh_bar[c & 1u], (c >> 1) & 1u)) { }
        // ---- run boundaries (found by tile_prims_kernel): every warp builds the chunk's bit mask of run starts for itself
        // ---- execute the runs in draw order (one 64-bit mask of run starts: two 32-bit halves cost registers this kernel does not have, measured 28 % slower)
        unsigned long long starts = 1ull;   // the chunk's first primitive opens a run in any case
#pragma unroll
        for (int w = 0; w < RCHUNK / 32; ++w) {
            const uint32_t k = (uint32_t)w * 32u + lane;
            starts |= (unsigned long long)__ballot_sync(0xffffffffu, k < n && (sh[k].meta & META_RUN_START) != 0u) << (32 * w);
        }
        starts &= ~1ull;   // the first run starts at 0; every later run starts where its predecessor ends (one find-first-set per run)
        uint32_t next_a = 0;
        for (bool more = true; more;) {
            const uint32_t a = next_a;
            const uint32_t b = starts != 0ull ? (uint32_t)__ffsll((long long)starts) - 1u : n;
            more = starts != 0ull;
            starts &= starts - 1ull;
            next_a = b;
            uint32_t first = a;   // a run that starts with staged-out primitives (first of the tile / of the chunk) has nothing else
            while (first < b && !(sh[first].meta & META_VALID)) ++first;
            if (first == b) continue;
            const uint32_t kind = (sh[first].meta >> META_KIND_SHIFT) & 3u;   // run_kind of its pipeline, set by tile_prims_kernel
            apply_pending();
            if (kind < 2u && b - a <= tg.pixel_run_max) {
                // short stencil run, pixel mode: the run's net effect on this pixel's samples, applied at once
                int net[S];
#pragma unroll
                for (int q = 0; q < S; ++q) net[q] = 0;
                for (uint32_t k = a; k < b; ++k) {
                    const TilePrim& ps = sh[k];
                    const uint32_t meta = ps.meta;
                    if (!(meta & META_VALID)) continue;
                    if (!pixel_in_box(ps)) continue;
                    const uint32_t pipe = meta & 15u;
                    const int delta = kind == 0u ? 1 : ((meta & META_FRONT) ? 1 : -1);
                    const uint32_t hit = (meta & META_E32) ? pixel_hits<S, int>(sc, ps, pipe, pipe != P_FILL_SOLID, lx, ly)
                                                           : pixel_hits<S, long long>(sc, ps, pipe, pipe != P_FILL_SOLID, lx, ly);
#pragma unroll
                    for (int q = 0; q < S; ++q)
                        if ((hit >> q) & 1u) net[q] = kind == 0u ? 1 : net[q] + delta;
                }
                const uint32_t ref = prim_ref(sh[first]);
#pragma unroll
                for (int q = 0; q < S; ++q) {
                    if (net[q] != 0) {
                        if (kind == 0u) { if ((ref & M) == (s[q] & M)) s[q] = (s[q] & ~W) | ((s[q] + 1u) & W); }                  // src/renderer.rs:571-576
                        else if ((ref & M) <= (s[q] & M)) s[q] = (s[q] & ~W) | ((s[q] + (uint32_t)net[q]) & W);                      // src/renderer.rs:577-582
                    }
                }
            } else if (kind < 2u) {
                // stencil run: one work item per (primitive, row). Runs of up to ROW_SWEEP_MAX primitives lay the items out on a fixed
                // grid, 16 primitives x 16 rows per sweep (no set-up; rows outside a primitive's bounding box idle). Longer runs COMPACT
                // them: every warp scans the run's row counts for itself (two primitives per lane), item i belongs to the primitive
                // whose prefix range holds i, so that small primitives keep the threads busy instead of idling through 16-row slots.
                int* const out = acc[cur];
                uint32_t* const pre = run_rows[warp];   // inclusive prefix sums of the row counts of primitives a .. a + 63
                const bool compact = b - a > ROW_SWEEP_MAX;
                uint32_t total_items = ((b - a + CR_TILE - 1u) / CR_TILE) * (CR_TILE * CR_TILE);
                if (compact) {
                    uint32_t h[2];
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const uint32_t k = a + (uint32_t)half * 32u + lane;
                        h[half] = 0;
                        if (k < b) {
                            const uint32_t meta = sh[k].meta, box = sh[k].box_mask;
                            if (meta & META_VALID) h[half] = (uint32_t)__popc(box >> 16);
                        }
                    }
                    uint32_t s0 = h[0], s1 = h[1];
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t y0 = __shfl_up_sync(0xffffffffu, s0, o), y1 = __shfl_up_sync(0xffffffffu, s1, o);
                        if (lane >= (uint32_t)o) { s0 += y0; s1 += y1; }
                    }
                    s1 += __shfl_sync(0xffffffffu, s0, 31);
                    total_items = __shfl_sync(0xffffffffu, s1, 31);
                    pre[lane] = s0;
                    pre[32u + lane] = s1;
                    __syncwarp();
                }
                for (uint32_t i = threadIdx.x; i < total_items; i += CR_TILE * CR_TILE) {
                    uint32_t k, row;
                    if (compact) {
                        uint32_t lo = 0;   // first slot whose inclusive prefix exceeds i
#pragma unroll
                        for (uint32_t step = 32u; step != 0u; step >>= 1)
                            if (pre[lo + step - 1u] <= i) lo += step;
                        k = a + lo;
                        row = i - (lo ? pre[lo - 1u] : 0u);
                    } else {
                        k = a + (i >> 8) * CR_TILE + ((i & 255u) >> 4);
                        row = i & 15u;
                        if (k >= b) continue;
                    }
                    const TilePrim& ps = sh[k];
                    const uint32_t meta = ps.meta, box = ps.box_mask;
                    if (!(meta & META_VALID)) continue;
                    const int y = box_y0(box) + (int)row;
                    if (y > box_y1(box)) continue;
                    const int x0 = box_x0(box), x1 = box_x1(box);
                    const uint32_t pipe = meta & 15u;
                    const int delta = kind == 0u ? 1 : ((meta & META_FRONT) ? 1 : -1);
                    if (meta & META_E32) stencil_row<S, int>(sc, ps, pipe, kind, delta, y, x0, x1, out);
                    else stencil_row<S, long long>(sc, ps, pipe, kind, delta, y, x0, x1, out);
                }
                __syncthreads();
                pending = cur;
                pending_kind = kind;
                pending_ref = prim_ref(sh[first]);
                cur ^= 1;
            } else if (b - a <= tg.pixel_run_max) {
                // short cover run (the usual hull of a few triangles), pixel mode: in draw order, this thread's pixel only
                for (uint32_t k = a; k < b; ++k) {
                    const TilePrim& ps = sh[k];
                    const uint32_t meta = ps.meta;
                    if (!(meta & META_VALID)) continue;
                    if (!pixel_in_box(ps)) continue;
                    uint32_t hit = (1u << S) - 1u;
                    if (!(meta & META_FULL))
                        hit = (meta & META_E32) ? pixel_hits<S, int>(sc, ps, 0u, false, lx, ly) : pixel_hits<S, long long>(sc, ps, 0u, false, lx, ly);
                    if (hit) cover_apply(ps, hit);
                }
            } else {
                // cover run: one (command, instance) hull draw. Order dependent, so the stencil / blend ops run one thread per
                // pixel, but coverage is found first by (primitive, row) work items as row masks, 16 primitives a sweep.
                for (uint32_t base = a; base < b; base += CR_TILE) {
                    {
                        const uint32_t k = base + (threadIdx.x >> 4);
                        unsigned long long mask = 0;
                        if (k < b) {
                            const TilePrim& ps = sh[k];
                            const uint32_t meta = ps.meta;
                            const uint32_t box = ps.box_mask;
                            const int y = (int)(threadIdx.x & 15u);
                            if ((meta & META_VALID) && ((box >> (16 + y)) & 1u) != 0u) {
                                const int x0 = box_x0(box), x1 = box_x1(box);
                                if (meta & META_FULL) mask = (x1 * S + S >= 64 ? ~0ull : ((1ull << (x1 * S + S)) - 1ull)) & ~((1ull << (x0 * S)) - 1ull);
                                else mask = (meta & META_E32) ? cover_row_mask<S, int>(ps, y, x0, x1) : cover_row_mask<S, long long>(ps, y, x0, x1);
                            }
                        }
                        cov[cb][threadIdx.x >> 4][threadIdx.x & 15u] = mask;
                    }
                    __syncthreads();
                    const uint32_t count = min((uint32_t)CR_TILE, b - base);
                    for (uint32_t j = 0; j < count; ++j) {
                        const uint32_t hit = (uint32_t)(cov[cb][j][ly] >> (lx * S)) & ((1u << S) - 1u);
                        if (hit) cover_apply(sh[base + j], hit);
                    }
                    cb ^= 1;   // the next sweep writes the other buffer; this one is rewritten only after that sweep's barrier
                }
            }
        }
    }
    apply_pending();
    if (in_fb) {
        auto store_stencil = [&](uint8_t* base) {
            if (S == 1) base[pix] = (uint8_t)s[0];
            else {
                uint32_t packed = 0;
#pragma unroll
                for (int k = 0; k < S; ++k) packed |= (s[k] & 255u) << (8 * k);
                *reinterpret_cast<uint32_t*>(base + pix) = packed;
            }
        };
        if (successor == 0u) store_stencil(tg.stencil);
        uint32_t texel[U8 ? S : 1];
        if (U8) {
#pragma unroll
            for (int k = 0; k < S; ++k) {
                const uint32_t r8 = unorm8_bits(col[k].x), g8 = unorm8_bits(col[k].y), b8 = unorm8_bits(col[k].z), a8 = unorm8_bits(col[k].w);
                texel[k] = tg.color_format == CR_FORMAT_BGRA8_UNORM ? (b8 | (g8 << 8) | (r8 << 16) | (a8 << 24)) : (r8 | (g8 << 8) | (b8 << 16) | (a8 << 24));
            }
        }
        auto store_color = [&](void* base) {
#pragma unroll
            for (int k = 0; k < S; ++k) {
                if (U8) reinterpret_cast<uint32_t*>(base)[pix + k] = texel[k];
                else reinterpret_cast<float4*>(base)[pix + k] = col[k];
            }
        };
        if (DEPTH) {
#pragma unroll
            for (int k = 0; k < S; ++k) tg.depth[pix + k] = dep[k];
        }
        if (successor != 0u) {
            // draw-order sharding, not the last rank of the tile's chain: the tile state goes into the successor's attachments
            const uint32_t slot = order_slot(tg, successor - 1u);
            store_stencil(tg.peer_stencil[slot]);
            store_color(tg.peer_color[slot]);
        } else {
            store_color(tg.color);
            // tile sharding / last rank of a chain: the finished tile also goes to every other rank's attachments (P2P stores over NVLink)
            const uint32_t peers = (tg.order_world > 1u && begin == end) ? 0u : max(tg.shard_world, tg.order_world);   // a tile no slice touches is cleared by every rank itself
            for (uint32_t peer = 0; peer + 1 < peers; ++peer) {
                void* pc = tg.peer_color[peer];
                uint8_t* ps = tg.peer_stencil[peer];
                if (pc == nullptr) continue;
                store_stencil(ps);
                store_color(pc);
            }
        }
        if (tg.shard_world > 1u || tg.order_world > 1u) __threadfence_system();
    }
    if (successor != 0u) {   // every thread's stores are fenced: tell the successor (release, system scope)
        __syncthreads();
        if (threadIdx.x == 0) st_release_sys(&tg.peer_exchange[order_slot(tg, successor - 1u)][cr_exchange_flag_offset() + tile], tg.order_epoch);
    }
    // covered-sample statistic: warp reduce, one atomic per warp
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) covered += __shfl_xor_sync(0xffffffffu, covered, o);
    if (lane == 0 && covered && counters != nullptr) atomicAdd(&counters->covered, (unsigned long long)covered);
}

}  // namespace

#define BIG_GRID (148 * 8)
int cr_raster_expand(cudaStream_t stream, const CompactCommand* compact, uint32_t n_commands, const DeviceBatch* batches, DeviceCommand* commands,
                     uint32_t* cmd_cand_begin, PassCounters* counters, uint32_t* scan_scratch) {
    if (n_commands <= CR_EXPAND_FUSED_MAX) {
        expand_scan_kernel<<<1, 1024, 0, stream>>>(compact, n_commands, batches, commands, cmd_cand_begin, counters);
        g_cr_kernel_launches += 1;
    } else {
        CR_CUDA_TRY(cudaMemsetAsync(counters, 0, sizeof(PassCounters), stream));
        expand_kernel<<<(n_commands + 255) / 256, 256, 0, stream>>>(compact, n_commands, batches, commands, cmd_cand_begin, counters);
        g_cr_kernel_launches += 1;
        const int st = cr_scan_exclusive(stream, cmd_cand_begin, n_commands + 1, 1, scan_scratch);   // wraps only when cand_total >= 2^32, which no capacity admits
        if (st != CR_OK) return st;
    }
    CR_CUDA_TRY(cudaGetLastError());
    return CR_OK;
}
#define CLIP_GRID 148
int cr_raster_setup(cudaStream_t stream, const RasterScene& scene, const RasterTarget& target, uint32_t cand_capacity, PrimRecord* records, uint32_t* cand_tiles,
                    uint32_t* big_list, uint32_t* clip_list, ClipAttr* clip_attrs, uint32_t clip_capacity, PassCounters* counters) {
    if (cand_capacity == 0) return CR_OK;
    CR_CUDA_TRY(cudaMemsetAsync(big_list, 0, 4, stream));
    CR_CUDA_TRY(cudaMemsetAsync(clip_list, 0, 4, stream));
    prim_setup_kernel<<<(cand_capacity + SETUP_THREADS - 1) / SETUP_THREADS, SETUP_THREADS, 0, stream>>>(scene, target, cand_capacity, records, cand_tiles, big_list, clip_list,
                                                                                                        counters);
    bin_big_kernel<false><<<BIG_GRID, 128, 0, stream>>>(target, records, big_list, cand_tiles, nullptr, nullptr, counters, 0u);
    clip_kernel<false><<<CLIP_GRID, 128, 0, stream>>>(scene, target, cand_capacity, clip_capacity, records, clip_attrs, clip_list, cand_tiles, nullptr, nullptr, counters, 0u);
    g_cr_kernel_launches += 3;
    CR_CUDA_TRY(cudaGetLastError());
    return CR_OK;
}
int cr_raster_bin_emit(cudaStream_t stream, const RasterTarget& target, uint32_t cand_capacity, uint32_t pair_capacity, const PrimRecord* records,
                       const uint32_t* cand_pair_begin, const uint32_t* big_list, const uint32_t* clip_list, uint32_t* pair_tile, uint32_t* pair_cand,
                       PassCounters* counters) {
    if (cand_capacity == 0) return CR_OK;
    bin_emit_kernel<<<(cand_capacity + SETUP_THREADS - 1) / SETUP_THREADS, SETUP_THREADS, 0, stream>>>(target, cand_capacity, pair_capacity, records, cand_pair_begin, pair_tile,
                                                                                                      pair_cand, counters);
    bin_big_kernel<true><<<BIG_GRID, 128, 0, stream>>>(target, records, big_list, const_cast<uint32_t*>(cand_pair_begin), pair_tile, pair_cand, counters, pair_capacity);
    clip_kernel<true><<<CLIP_GRID, 128, 0, stream>>>(RasterScene{}, target, cand_capacity, 0u, const_cast<PrimRecord*>(records), nullptr, clip_list, const_cast<uint32_t*>(cand_pair_begin),
                                                     pair_tile, pair_cand, counters, pair_capacity);
    g_cr_kernel_launches += 3;
    CR_CUDA_TRY(cudaGetLastError());
    return CR_OK;
}
int cr_raster_publish_touched_tiles(cudaStream_t stream, const RasterTarget& target, const uint32_t* tile_begin) {
    if (target.order_world <= 1u) return CR_OK;
    touched_tiles_kernel<<<(target.order_mask_words + 255) / 256, 256, 0, stream>>>(target, tile_begin);
    touched_ready_kernel<<<1, 32, 0, stream>>>(target);
    g_cr_kernel_launches += 2;
    CR_CUDA_TRY(cudaGetLastError());
    return CR_OK;
}
size_t cr_tile_prim_bytes() { return sizeof(TilePrim); }
int cr_raster_tile_prims(cudaStream_t stream, const RasterScene& scene, const RasterTarget& target, const PrimRecord* records, const uint32_t* pair_tile, const uint32_t* pair_cand,
                         uint32_t pair_capacity, void* tile_prims, const PassCounters* counters) {
    if (pair_capacity == 0) return CR_OK;
    const uint32_t grid = (pair_capacity + 255u) / 256u;
    TilePrim* out = static_cast<TilePrim*>(tile_prims);
    const bool depth = target.depth != nullptr;
    if (target.samples == 4) {
        if (depth) tile_prims_kernel<4, true><<<grid, 256, 0, stream>>>(scene, target, records, pair_tile, pair_cand, pair_capacity, out, counters);
        else tile_prims_kernel<4, false><<<grid, 256, 0, stream>>>(scene, target, records, pair_tile, pair_cand, pair_capacity, out, counters);
    } else {
        if (depth) tile_prims_kernel<1, true><<<grid, 256, 0, stream>>>(scene, target, records, pair_tile, pair_cand, pair_capacity, out, counters);
        else tile_prims_kernel<1, false><<<grid, 256, 0, stream>>>(scene, target, records, pair_tile, pair_cand, pair_capacity, out, counters);
    }
    g_cr_kernel_launches += 1;
    CR_CUDA_TRY(cudaGetLastError());
    return CR_OK;
}
int cr_raster_tiles(cudaStream_t stream, const RasterScene& scene, const RasterTarget& target, const void* tile_prims, const uint32_t* tile_begin, PassCounters* counters) {
    const uint32_t n_tiles = target.tiles_x * target.tiles_y;
    if (n_tiles == 0) return CR_OK;
    const TilePrim* prims = static_cast<const TilePrim*>(tile_prims);
    const bool depth = target.depth != nullptr, u8 = target.color_format != CR_FORMAT_RGBA32F;
#define K3_LAUNCH(S, D, U) raster_tiles_kernel<S, D, U><<<n_tiles, CR_TILE * CR_TILE, 0, stream>>>(scene, target, prims, tile_begin, counters)
    if (target.samples == 4) {
        if (depth) { if (u8) K3_LAUNCH(4, true, true); else K3_LAUNCH(4, true, false); }
        else { if (u8) K3_LAUNCH(4, false, true); else K3_LAUNCH(4, false, false); }
    } else {
        if (depth) { if (u8) K3_LAUNCH(1, true, true); else K3_LAUNCH(1, true, false); }
        else { if (u8) K3_LAUNCH(1, false, true); else K3_LAUNCH(1, false, false); }
    }
#undef K3_LAUNCH
    g_cr_kernel_launches += 1;
    CR_CUDA_TRY(cudaGetLastError());
    return CR_OK;
}
