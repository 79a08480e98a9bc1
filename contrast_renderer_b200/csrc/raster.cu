// raster.cu — the binner and K3, the software stencil-then-cover tile rasteriser for sm_100a.
//
// Replaces the wgpu render pass of the reference: Shape::render draw recording (src/renderer.rs:267-355), the 13
// pipeline stencil / blend states (src/renderer.rs:565-861) and every entry point of src/shaders.wgsl.
//
// Data flow (all in HBM, no host round trips except one pair-count read-back):
//   commands --count--> candidate primitives (one per index slot / list triangle / hull strip triangle, in exact
//   draw order) --bin--> (tile, candidate) pairs --stable radix sort by tile--> per-tile ranges --K3--> framebuffer.
// K3: one CTA per 16x16 tile, one thread per pixel. The pixel's stencil byte and RGBA colour live in REGISTERS for
// the whole pass; triangles are set up cooperatively (one thread per triangle) into shared memory in chunks of 256
// and then every pixel walks the chunk in draw order (shared-memory broadcast reads, warp-uniform branches). The
// tile is read once and written once with 128-bit accesses, so overdraw costs no HBM traffic.
//
// Rasterisation contract: see the header comment of oracle/raster.hpp (written independently, same rules).
#include "device_common.cuh"
#include "prims.h"
#include "raster.h"

namespace {

#define CHUNK 256

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

__device__ __forceinline__ uint32_t slots_of(const DeviceBatch& b, uint32_t shape, int cat) {
    const size_t stride = (size_t)b.n_shapes + 1;
    if (cat <= 2) {
        const size_t row = (size_t)(CNT_LINE_IDX + cat) * stride;
        return b.cat_begin[row + shape + 1] - b.cat_begin[row + shape];
    }
    if (cat <= 6) {
        const size_t row = (size_t)cat * stride;
        return (b.cat_begin[row + shape + 1] - b.cat_begin[row + shape]) / 3u;
    }
    const uint32_t hc = b.hull_count[shape];
    return hc >= 3 ? hc - 2 : 0u;
}
__device__ __forceinline__ bool cat_drawn(const DeviceBatch& b, int cat) { return cat >= 2 || b.n_groups > 0; }   // src/renderer.rs:276

// Candidate primitives of one command: Stencil = 7 instanced draws in category order (src/renderer.rs:275-336),
// everything else = one instanced hull draw (:345-354).
__global__ void count_candidates_kernel(const DeviceBatch* __restrict__ batches, const DeviceCommand* __restrict__ commands, uint32_t n, uint32_t* __restrict__ out) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const DeviceCommand cmd = commands[c];
    const DeviceBatch& b = batches[cmd.batch];
    const uint32_t n_inst = cmd.instance_end > cmd.instance_begin ? cmd.instance_end - cmd.instance_begin : 0u;
    uint32_t total = 0;
    if (cmd.operation == CR_OP_STENCIL) {
        for (int cat = 0; cat < 7; ++cat)
            if (cat_drawn(b, cat)) total += slots_of(b, cmd.shape, cat) * n_inst;
    } else {
        total = slots_of(b, cmd.shape, 7) * n_inst;
    }
    out[c] = total;
}

struct Candidate {
    uint32_t cmd;
    uint32_t cat;      // 0..7
    uint32_t instance;
    uint32_t local;    // slot / triangle number inside the category
};
__device__ Candidate decode_candidate(const RasterScene& sc, uint32_t cand, DeviceCommand& cmd_out) {
    uint32_t lo = 0, hi = sc.n_commands;   // last command with begin <= cand
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (sc.cmd_cand_begin[mid] <= cand) lo = mid; else hi = mid; }
    Candidate k;
    k.cmd = lo;
    const DeviceCommand cmd = sc.commands[lo];
    cmd_out = cmd;
    const DeviceBatch& b = sc.batches[cmd.batch];
    uint32_t rem = cand - sc.cmd_cand_begin[lo];
    const uint32_t n_inst = cmd.instance_end - cmd.instance_begin;
    k.cat = 7;
    if (cmd.operation == CR_OP_STENCIL) {
        for (int cat = 0; cat < 7; ++cat) {
            if (!cat_drawn(b, cat)) continue;
            const uint32_t n = slots_of(b, cmd.shape, cat);
            const uint32_t tot = n * n_inst;
            if (rem < tot) { k.cat = cat; k.instance = cmd.instance_begin + rem / n; k.local = rem % n; return k; }
            rem -= tot;
        }
    }
    const uint32_t n = slots_of(b, cmd.shape, 7);
    k.instance = cmd.instance_begin + rem / n;
    k.local = rem % n;
    return k;
}

// The three vertex numbers (absolute, inside the batch-wide category array) of a candidate; false if the slot is
// not a triangle (restart inside, strip too short).
__device__ bool candidate_vertices(const DeviceBatch& b, uint32_t shape, const Candidate& k, uint32_t v[3], bool& odd) {
    const size_t stride = (size_t)b.n_shapes + 1;
    if (k.cat <= 2) {
        const size_t irow = (size_t)(CNT_LINE_IDX + k.cat) * stride;
        const uint32_t ib = b.cat_begin[irow + shape], ie = b.cat_begin[irow + shape + 1];
        if (k.local + 2 >= ie - ib) return false;
        const uint32_t* idx = b.idx[k.cat] + ib + k.local;
        const uint32_t i0 = idx[0], i1 = idx[1], i2 = idx[2];
        if (i0 == CR_RESTART || i1 == CR_RESTART || i2 == CR_RESTART) return false;
        const uint32_t vb = b.cat_begin[(size_t)k.cat * stride + shape];
        v[0] = vb + (i0 >> 1); v[1] = vb + (i1 >> 1); v[2] = vb + (i2 >> 1);
        odd = (i0 & 1u) != 0;
        return true;
    }
    if (k.cat <= 6) {
        const uint32_t vb = b.cat_begin[(size_t)k.cat * stride + shape];
        v[0] = vb + 3 * k.local; v[1] = v[0] + 1; v[2] = v[0] + 2;
        odd = false;
        return true;
    }
    const uint32_t vb = b.cat_begin[(size_t)CNT_PROTO * stride + shape];
    v[0] = vb + k.local; v[1] = v[0] + 1; v[2] = v[0] + 2;
    odd = (k.local & 1u) != 0;
    return true;
}
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

struct SnapVertex { int X, Y; float invw; bool ok; };
// vertex shader + viewport transform + 1/256 px snapping (src/shaders.wgsl:66-74)
__device__ __forceinline__ SnapVertex snap_vertex(const float* __restrict__ m, float x, float y, uint32_t W, uint32_t H) {
    SnapVertex v;
    const float cx = (m[0] * x + m[4] * y) + m[12];
    const float cy = (m[1] * x + m[5] * y) + m[13];
    const float cw = (m[3] * x + m[7] * y) + m[15];
    v.ok = cw > 0.0f;
    v.invw = 1.0f / cw;
    const float fx = ((cx * v.invw) * 0.5f + 0.5f) * (float)W;
    const float fy = (0.5f - (cy * v.invw) * 0.5f) * (float)H;
    if (!(cr::fabs_f(fx) <= 2097152.0f) || !(cr::fabs_f(fy) <= 2097152.0f)) v.ok = false;
    v.X = v.ok ? (int)cr::floor_f(fx * 256.0f + 0.5f) : 0;
    v.Y = v.ok ? (int)cr::floor_f(fy * 256.0f + 0.5f) : 0;
    return v;
}

// A triangle oriented clockwise on screen (positive doubled area in y-down pixels) with its three edge functions
// E_e(P) = A_e * (P.y - Y_e) - B_e * (P.x - X_e), inside when E_e + bias_e >= 0 (top-left rule).
struct Tri {
    int X[3], Y[3];
    int A[3], B[3];
    int bias[3];
    bool front, swapped;
};
__device__ __forceinline__ bool make_tri(const SnapVertex& a, const SnapVertex& b, const SnapVertex& c, bool odd, Tri& t) {
    if (!a.ok || !b.ok || !c.ok) return false;
    long long area2 = (long long)(b.X - a.X) * (c.Y - a.Y) - (long long)(c.X - a.X) * (b.Y - a.Y);
    if (area2 == 0) return false;
    t.front = (area2 < 0) != odd;
    t.swapped = area2 < 0;
    t.X[0] = a.X; t.Y[0] = a.Y;
    if (t.swapped) { t.X[1] = c.X; t.Y[1] = c.Y; t.X[2] = b.X; t.Y[2] = b.Y; }
    else { t.X[1] = b.X; t.Y[1] = b.Y; t.X[2] = c.X; t.Y[2] = c.Y; }
#pragma unroll
    for (int e = 0; e < 3; ++e) {
        const int n = e == 2 ? 0 : e + 1;
        t.A[e] = t.X[n] - t.X[e];
        t.B[e] = t.Y[n] - t.Y[e];
        t.bias[e] = ((t.B[e] == 0 && t.A[e] > 0) || t.B[e] < 0) ? 0 : -1;
    }
    return true;
}
__device__ __forceinline__ int floor_div256(int v) { return v >> 8; }   // arithmetic shift = floor for negatives

// Visit every tile whose sample centres can be inside the triangle (bounding box + per-edge trivial reject).
template <typename F>
__device__ __forceinline__ void for_each_tile(const Tri& t, const RasterTarget& tg, F f) {
    const int minX = min(t.X[0], min(t.X[1], t.X[2])), maxX = max(t.X[0], max(t.X[1], t.X[2]));
    const int minY = min(t.Y[0], min(t.Y[1], t.Y[2])), maxY = max(t.Y[0], max(t.Y[1], t.Y[2]));
    // pixel centres are at 256*p + 128: first centre >= minX, last centre <= maxX
    const int px0 = max(0, (minX - 128 + 255) >> 8), px1 = min((int)tg.width - 1, (maxX - 128) >> 8);
    const int py0 = max(0, (minY - 128 + 255) >> 8), py1 = min((int)tg.height - 1, (maxY - 128) >> 8);
    if (px0 > px1 || py0 > py1) return;
    const int tx0 = px0 / CR_TILE, tx1 = px1 / CR_TILE, ty0 = py0 / CR_TILE, ty1 = py1 / CR_TILE;
    for (int ty = ty0; ty <= ty1; ++ty) {
        const int ylo = max(py0, ty * CR_TILE) * 256 + 128, yhi = min(py1, ty * CR_TILE + CR_TILE - 1) * 256 + 128;
        for (int tx = tx0; tx <= tx1; ++tx) {
            const int xlo = max(px0, tx * CR_TILE) * 256 + 128, xhi = min(px1, tx * CR_TILE + CR_TILE - 1) * 256 + 128;
            bool out = false;
#pragma unroll
            for (int e = 0; e < 3; ++e) {
                const int PX = t.B[e] < 0 ? xhi : xlo, PY = t.A[e] > 0 ? yhi : ylo;   // corner maximising E_e
                const long long E = (long long)t.A[e] * (PY - t.Y[e]) - (long long)t.B[e] * (PX - t.X[e]);
                if (E + t.bias[e] < 0) out = true;
            }
            if (!out) f((uint32_t)(ty * (int)tg.tiles_x + tx));
        }
    }
}

// Geometry of a candidate as the binner needs it (positions only). Returns false if nothing can be drawn.
__device__ bool candidate_triangle(const RasterScene& sc, const RasterTarget& tg, uint32_t cand, Tri& tri) {
    DeviceCommand cmd;
    const Candidate k = decode_candidate(sc, cand, cmd);
    const DeviceBatch& b = sc.batches[cmd.batch];
    uint32_t v[3];
    bool odd;
    if (!candidate_vertices(b, cmd.shape, k, v, odd)) return false;
    const float* m = sc.transforms + 16 * (size_t)k.instance;
    SnapVertex sv[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float* p = vertex_ptr(b, k.cat, v[i]);
        sv[i] = snap_vertex(m, p[0], p[1], tg.width, tg.height);
    }
    if (!make_tri(sv[0], sv[1], sv[2], odd, tri)) return false;
    if (cmd.operation == CR_OP_COLOR) {   // cull_mode applies to the colour cover only (src/renderer.rs:743)
        if (tg.cull_mode == CR_CULL_BACK && !tri.front) return false;
        if (tg.cull_mode == CR_CULL_FRONT && tri.front) return false;
    }
    return true;
}

__global__ void __launch_bounds__(128) bin_count_kernel(RasterScene sc, RasterTarget tg, uint32_t n, uint32_t* __restrict__ cand_tiles) {
    const uint32_t cand = blockIdx.x * blockDim.x + threadIdx.x;
    if (cand >= n) return;
    Tri tri;
    uint32_t count = 0;
    if (candidate_triangle(sc, tg, cand, tri)) for_each_tile(tri, tg, [&](uint32_t) { ++count; });
    cand_tiles[cand] = count;
}
__global__ void __launch_bounds__(128) bin_emit_kernel(RasterScene sc, RasterTarget tg, uint32_t n, const uint32_t* __restrict__ begin, uint32_t* __restrict__ pair_tile,
                                                      uint32_t* __restrict__ pair_cand) {
    const uint32_t cand = blockIdx.x * blockDim.x + threadIdx.x;
    if (cand >= n) return;
    if (begin[cand + 1] == begin[cand]) return;
    Tri tri;
    if (!candidate_triangle(sc, tg, cand, tri)) return;
    uint32_t at = begin[cand];
    for_each_tile(tri, tg, [&](uint32_t tile) { pair_tile[at] = tile; pair_cand[at] = cand; ++at; });
}

// ------------------------------------------------------------------------------------------- K3: tile raster
struct PrimSetup {
    long long e0[3];        // edge functions (with top-left bias folded in) at the tile's first pixel centre
    int A[3], B[3];         // per-pixel steps are 256*A (y) and -256*B (x)
    int bias[3];
    float invw[3];
    float attr[3][4];
    uint32_t flat_u;
    float flat_f;
    uint32_t pipe;          // Pipe | front << 8 | valid << 9
    uint32_t ref;
    uint32_t instance;
    uint32_t batch;
    uint32_t layers;        // save_layer | restore_layer << 16
    uint32_t bbox;          // x0 | y0 << 8 | x1 << 16 | y1 << 24 in tile pixels
};

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

__device__ void setup_primitive(const RasterScene& sc, const RasterTarget& tg, uint32_t cand, int tile_px, int tile_py, PrimSetup& ps) {
    ps.pipe = 0;
    DeviceCommand cmd;
    const Candidate k = decode_candidate(sc, cand, cmd);
    const DeviceBatch& b = sc.batches[cmd.batch];
    uint32_t v[3];
    bool odd;
    if (!candidate_vertices(b, cmd.shape, k, v, odd)) return;
    const float* m = sc.transforms + 16 * (size_t)k.instance;
    const int n_attr = (int)((0x04332032u >> (4u * k.cat)) & 15u);   // attribute floats per category: 2,3,0,2,3,3,4,0
    SnapVertex sv[3];
    float attr[3][4];
    uint32_t flat_u = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float* p = vertex_ptr(b, k.cat, v[i]);
        sv[i] = snap_vertex(m, p[0], p[1], tg.width, tg.height);
#pragma unroll
        for (int a = 0; a < 4; ++a) attr[i][a] = a < n_attr ? p[2 + a] : 0.0f;
        if (i == 0 && k.cat <= 1) flat_u = __float_as_uint(p[2 + n_attr]);
    }
    Tri t;
    if (!make_tri(sv[0], sv[1], sv[2], odd, t)) return;
    uint32_t pipe;
    if (cmd.operation == CR_OP_STENCIL) pipe = k.cat;
    else pipe = P_CLIP + (cmd.operation - CR_OP_CLIP);   // CLIP, UNCLIP, COLOR, SAVE, SCALE, RESTORE follow the enum order
    ps.flat_u = flat_u;
    ps.flat_f = attr[0][1];
    const int i1 = t.swapped ? 2 : 1, i2 = t.swapped ? 1 : 2;
    ps.invw[0] = sv[0].invw; ps.invw[1] = sv[i1].invw; ps.invw[2] = sv[i2].invw;
#pragma unroll
    for (int a = 0; a < 4; ++a) { ps.attr[0][a] = attr[0][a]; ps.attr[1][a] = attr[i1][a]; ps.attr[2][a] = attr[i2][a]; }
    const int PX = tile_px * 256 + 128, PY = tile_py * 256 + 128;
#pragma unroll
    for (int e = 0; e < 3; ++e) {
        ps.A[e] = t.A[e]; ps.B[e] = t.B[e]; ps.bias[e] = t.bias[e];
        ps.e0[e] = (long long)t.A[e] * (PY - t.Y[e]) - (long long)t.B[e] * (PX - t.X[e]) + t.bias[e];
    }
    // pixel bounding box clipped to this tile
    const int minX = min(t.X[0], min(t.X[1], t.X[2])), maxX = max(t.X[0], max(t.X[1], t.X[2]));
    const int minY = min(t.Y[0], min(t.Y[1], t.Y[2])), maxY = max(t.Y[0], max(t.Y[1], t.Y[2]));
    const int x0 = max(0, ((minX - 128 + 255) >> 8) - tile_px), x1 = min(CR_TILE - 1, ((maxX - 128) >> 8) - tile_px);
    const int y0 = max(0, ((minY - 128 + 255) >> 8) - tile_py), y1 = min(CR_TILE - 1, ((maxY - 128) >> 8) - tile_py);
    if (x0 > x1 || y0 > y1) return;
    ps.bbox = (uint32_t)x0 | ((uint32_t)y0 << 8) | ((uint32_t)x1 << 16) | ((uint32_t)y1 << 24);
    ps.ref = cmd.ref;
    ps.instance = k.instance;
    ps.batch = cmd.batch;
    ps.layers = cmd.save_layer | (cmd.restore_layer << 16);
    ps.pipe = pipe | (t.front ? 256u : 0u) | 512u;
}

__global__ void __launch_bounds__(CR_TILE * CR_TILE) raster_tiles_kernel(RasterScene sc, RasterTarget tg, const uint32_t* __restrict__ tile_begin,
                                                                         const uint32_t* __restrict__ pair_cand, unsigned long long* __restrict__ covered_out) {
    __shared__ PrimSetup sh[CHUNK];
    const uint32_t tile = blockIdx.x;
    const uint32_t begin = tile_begin[tile], end = tile_begin[tile + 1];
    if (begin == end) return;
    const int tile_px = (int)(tile % tg.tiles_x) * CR_TILE, tile_py = (int)(tile / tg.tiles_x) * CR_TILE;
    const int lx = threadIdx.x & (CR_TILE - 1), ly = threadIdx.x / CR_TILE;
    const int px = tile_px + lx, py = tile_py + ly;
    const bool in_fb = px < (int)tg.width && py < (int)tg.height;
    const size_t pix = (size_t)py * tg.width + px;
    uint32_t s = 0;
    float4 col = make_float4(0.f, 0.f, 0.f, 0.f);
    if (in_fb) { s = tg.stencil[pix]; col = tg.color[pix]; }
    const uint32_t W = tg.wmask, C = tg.cmask, M = W | C;
    const int warp_y0 = (threadIdx.x >> 5) * (32 / CR_TILE), warp_y1 = warp_y0 + (32 / CR_TILE) - 1;
    uint32_t covered = 0;
    const size_t layer_stride = (size_t)tg.width * tg.height;

    for (uint32_t chunk = begin; chunk < end; chunk += CHUNK) {
        const uint32_t n = min((uint32_t)CHUNK, end - chunk);
        __syncthreads();
        if (threadIdx.x < n) setup_primitive(sc, tg, pair_cand[chunk + threadIdx.x], tile_px, tile_py, sh[threadIdx.x]);
        __syncthreads();
        for (uint32_t k = 0; k < n; ++k) {
            const PrimSetup& ps = sh[k];
            const uint32_t meta = ps.pipe;
            if (!(meta & 512u)) continue;
            const uint32_t bbox = ps.bbox;
            const int by0 = (bbox >> 8) & 255, by1 = bbox >> 24;
            if (by1 < warp_y0 || by0 > warp_y1) continue;                      // warp-uniform reject
            const int bx0 = bbox & 255, bx1 = (bbox >> 16) & 255;
            if (lx < bx0 || lx > bx1 || ly < by0 || ly > by1 || !in_fb) continue;
            long long E[3];
            bool inside = true;
#pragma unroll
            for (int e = 0; e < 3; ++e) {
                E[e] = ps.e0[e] + (long long)ps.A[e] * (ly * 256) - (long long)ps.B[e] * (lx * 256);
                if (E[e] < 0) inside = false;
            }
            if (!inside) continue;
            const uint32_t pipe = meta & 255u;
            const bool front = (meta & 256u) != 0;
            // ---- fragment stage: perspective-correct attributes at the sample + sample_mask predicate
            bool keep = true;
            if (pipe != P_FILL_SOLID && pipe < P_CLIP) {
                const float e0 = (float)(E[1] - ps.bias[1]) * ps.invw[0], e1 = (float)(E[2] - ps.bias[2]) * ps.invw[1], e2 = (float)(E[0] - ps.bias[0]) * ps.invw[2];
                const float den = (e0 + e1) + e2;
                float a[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) a[q] = ((e0 * ps.attr[0][q] + e1 * ps.attr[1][q]) + e2 * ps.attr[2][q]) / den;
                switch (pipe) {
                    case P_FILL_IQ: keep = a[0] * a[0] - a[1] <= 0.0f; break;
                    case P_FILL_IC: keep = a[0] * a[0] * a[0] - a[1] * a[2] <= 0.0f; break;
                    case P_FILL_RQ: keep = a[0] * a[0] - a[1] * a[2] <= 0.0f; break;
                    case P_FILL_RC: keep = a[0] * a[0] * a[0] - a[1] * a[2] * a[3] <= 0.0f; break;
                    case P_STROKE_LINE: {   // src/shaders.wgsl:268-285
                        const Descriptor& d = reinterpret_cast<const Descriptor*>(sc.batches[ps.batch].stroke)[ps.flat_u & 65535u];
                        if ((d.count_dashed_join & 4u) != 0u) keep = stroke_dashed(d, a[0], a[1]);
                        else if ((ps.flat_u & 65536u) != 0u) keep = cap_test(a[0], a[1] - ps.flat_f, d.caps >> 4u);
                        else if (a[1] < 0.0f) keep = cap_test(a[0], -a[1], d.caps);
                    } break;
                    default: {              // P_STROKE_JOINT, src/shaders.wgsl:287-300
                        const Descriptor& d = reinterpret_cast<const Descriptor*>(sc.batches[ps.batch].stroke)[ps.flat_u & 65535u];
                        const float radius = cr::sqrt_f(a[0] * a[0] + a[1] * a[1]);
                        const uint32_t kind = d.count_dashed_join & 3u;
                        keep = kind == 1u ? (ps.flat_u & 65536u) != 0u : (kind == 2u ? radius <= 0.5f : true);
                        if (keep && (d.count_dashed_join & 4u) != 0u)
                            keep = stroke_dashed(d, radius, a[2] + cr::atan2_f(a[1], a[0]) / 6.28318548202514648438f);
                    } break;
                }
            }
            if (!keep) continue;
            // ---- output merger: stencil test / op and colour blend (src/renderer.rs:571-861)
            const uint32_t ref = ps.ref;
            if (pipe <= P_STROKE_JOINT) {
                if ((ref & M) == (s & M)) s = (s & ~W) | ((s + 1u) & W);
            } else if (pipe <= P_FILL_RC) {
                if ((ref & M) <= (s & M)) s = (s & ~W) | ((front ? s + 1u : s - 1u) & W);
            } else if (pipe == P_COLOR) {
                if ((ref & M) < (s & M)) {
                    const float4 ic = reinterpret_cast<const float4*>(sc.colors)[ps.instance];
                    const float sa = ic.w;
                    const float sr = ic.x * sa, sg = ic.y * sa, sb = ic.z * sa;
                    if (tg.blending == CR_BLEND_PREMULTIPLIED_OVER) {
                        const float kk = 1.0f - sa;
                        col.x = sr + col.x * kk; col.y = sg + col.y * kk; col.z = sb + col.z * kk; col.w = sa + col.w * kk;
                    } else { col.x = sr; col.y = sg; col.z = sb; col.w = sa; }
                    covered += 1;
                }
                s = s & ~W;
            } else if (pipe == P_CLIP) {
                if ((ref & W) != (s & W)) s = (s & ~M) | (ref & M);
            } else if (pipe == P_UNCLIP) {
                if ((ref & C) < (s & C)) s = (s & ~M) | (ref & M);
            } else if ((ref & M) <= (s & M)) {   // the three alpha-context covers share one stencil state (:761-766)
                if (pipe == P_SAVE_ALPHA) {
                    tg.alpha_layers[(size_t)(ps.layers & 65535u) * layer_stride + pix] = col.w;
                } else if (pipe == P_SCALE_ALPHA) {
                    const float sa = 1.0f - reinterpret_cast<const float4*>(sc.colors)[ps.instance].w;
                    col.w = sa + col.w * (1.0f - sa);
                } else {
                    const float saved = tg.alpha_layers[(size_t)(ps.layers >> 16) * layer_stride + pix];
                    const float sa = (1.0f - saved) * (1.0f - reinterpret_cast<const float4*>(sc.colors)[ps.instance].w);
                    col.w = col.w - sa;
                }
            }
        }
    }
    if (in_fb) { tg.stencil[pix] = (uint8_t)s; tg.color[pix] = col; }
    // covered-sample statistic: warp reduce, one atomic per warp
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) covered += __shfl_xor_sync(0xffffffffu, covered, o);
    if ((threadIdx.x & 31u) == 0 && covered) atomicAdd(covered_out, (unsigned long long)covered);
}

}  // namespace

int cr_raster_count_candidates(cudaStream_t stream, const DeviceBatch* batches, const DeviceCommand* commands, uint32_t n_commands, uint32_t* cmd_cands) {
    if (n_commands == 0) return CR_OK;
    count_candidates_kernel<<<(n_commands + 255) / 256, 256, 0, stream>>>(batches, commands, n_commands, cmd_cands);
    g_cr_kernel_launches += 1;
    CR_CUDA_TRY(cudaGetLastError());
    return CR_OK;
}
int cr_raster_bin_count(cudaStream_t stream, const RasterScene& scene, const RasterTarget& target, uint32_t n_candidates, uint32_t* cand_tiles) {
    if (n_candidates == 0) return CR_OK;
    bin_count_kernel<<<(n_candidates + 127) / 128, 128, 0, stream>>>(scene, target, n_candidates, cand_tiles);
    g_cr_kernel_launches += 1;
    CR_CUDA_TRY(cudaGetLastError());
    return CR_OK;
}
int cr_raster_bin_emit(cudaStream_t stream, const RasterScene& scene, const RasterTarget& target, uint32_t n_candidates, const uint32_t* cand_pair_begin,
                       uint32_t* pair_tile, uint32_t* pair_cand) {
    if (n_candidates == 0) return CR_OK;
    bin_emit_kernel<<<(n_candidates + 127) / 128, 128, 0, stream>>>(scene, target, n_candidates, cand_pair_begin, pair_tile, pair_cand);
    g_cr_kernel_launches += 1;
    CR_CUDA_TRY(cudaGetLastError());
    return CR_OK;
}
int cr_raster_tiles(cudaStream_t stream, const RasterScene& scene, const RasterTarget& target, const uint32_t* tile_begin, const uint32_t* pair_cand,
                    unsigned long long* covered_samples) {
    const uint32_t n_tiles = target.tiles_x * target.tiles_y;
    if (n_tiles == 0) return CR_OK;
    raster_tiles_kernel<<<n_tiles, CR_TILE * CR_TILE, 0, stream>>>(scene, target, tile_begin, pair_cand, covered_samples);
    g_cr_kernel_launches += 1;
    CR_CUDA_TRY(cudaGetLastError());
    return CR_OK;
}
