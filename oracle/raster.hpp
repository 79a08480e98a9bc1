// ORACLE — TEST INFRASTRUCTURE ONLY (see ga.hpp). PARITY UNPINNED (no wgpu / Vulkan here; semantics = WebGPU spec).
// Scalar software restatement of the stencil-then-cover pass: Shape::render draw order (src/renderer.rs:267-355),
// the 13 pipeline stencil / blend states (src/renderer.rs:565-861) and the fragment entry points of
// src/shaders.wgsl. One sample at a time, one triangle at a time, in exact draw order.
//
// Rasterisation contract (shared, in words, with csrc/raster.cu — the code is written twice):
//  * clip = col0*x + col1*y + col3 of the instance mat4 (src/shaders.wgsl:20-27,72). A triangle with a vertex at w <= 0 or
//    further than 2^21 px from the origin is CLIPPED in clip space (WebGPU clips primitives to the view volume): Sutherland-
//    Hodgman against, in this order, w >= 1e-6, G w - x >= 0, G w + x >= 0, G w - y >= 0, G w + y >= 0 with the guard band
//    G = float(2^20 / max(W, H)) (integer division, at least 1); the point where an edge meets a plane is interpolated from
//    the edge's INSIDE end (in + (out - in) * (d_in / (d_in - d_out)), every component and attribute alike), so triangles
//    sharing the edge agree on it; the polygon is cut into the fan (p0, pi, pi+1), each fan triangle is snapped and
//    rasterised like any other, in the original's place; flat attributes stay those of the original first vertex. Depth is
//    not clipped (unclipped-depth semantics).
//  * depth (colour cover only, src/renderer.rs:743-745): z / w of each vertex, interpolated linearly in screen space with
//    the unbiased edge values as weights, f32 attachment, viewport depth range [0, 1]; order stencil test -> depth test;
//    stencil fail -> fail_op (Zero), depth fail -> depth_fail_op (Keep, src/renderer.rs:442), both pass -> pass_op (Zero),
//    colour write, depth write if enabled.
//  * 8-bit colour formats: every blend result is stored as unorm8 (clamp, x*255 + 0.5, truncate) and read back as c / 255;
//    the alpha layers are R8Unorm (src/renderer.rs:783,898). The float arrays below then hold exactly those c / 255 values.
//  * framebuffer coordinates: fx = (ndc.x*0.5+0.5)*W, fy = (0.5-ndc.y*0.5)*H, snapped to 1/256 px
//    (floor(v*256+0.5)); edge functions are exact 64-bit integers; top-left fill rule; zero-area => nothing.
//  * front face = counter-clockwise in NDC (src/renderer.rs:477) = negative doubled area in y-down pixels; odd
//    strip triangles flip; 0xFFFF restarts a strip (src/renderer.rs:476).
//  * attributes: perspective-correct at the sample, a = sum(E_i/w_i * a_i) / sum(E_i/w_i); flat attributes come
//    from the first vertex of the primitive.
//  * samples: pixel centre for 1x; (6,2),(14,6),(2,10),(10,14)/16 for 4x; shading is per sample
//    (@interpolate(perspective, sample), src/shaders.wgsl:35).
#pragma once
#include "tessellate.hpp"

namespace oracle {

struct RenderCommand {
    uint32_t shape;
    uint32_t instance_begin, instance_end;
    uint32_t operation;       // cr_render_operation
    uint32_t clip_depth;      // Renderer::set_clip_depth state (src/renderer.rs:932-938)
    uint32_t save_layer;      // Renderer::save_alpha_context state
    uint32_t restore_layer;   // Renderer::restore_alpha_context state
};

struct RasterShape {
    const uint8_t* vertex_buffer;
    uint64_t vertex_offsets[8];
    const uint32_t* wide_indices;   // [line | joint | solid] 32-bit copies of the u16 index buffer (0xFFFFFFFF = restart)
    uint64_t index_counts[3];       // entries per category
    const DynamicStrokeDescriptor* stroke;
    uint64_t n_groups;
};

struct Framebuffer {
    uint32_t width, height, samples;
    float* color;        // [h][w][s][4]
    uint8_t* stencil;    // [h][w][s]
    float* alpha_layers; // [layer][h][w][s]
    uint64_t covered_samples;
    float* depth = nullptr;   // [h][w][s], or null: no depth attachment
};

enum PipelineKind {
    PIPE_STROKE_LINE, PIPE_STROKE_JOINT, PIPE_FILL_SOLID, PIPE_FILL_IQ, PIPE_FILL_IC, PIPE_FILL_RQ, PIPE_FILL_RC,
    PIPE_CLIP, PIPE_UNCLIP, PIPE_COLOR, PIPE_SAVE_ALPHA, PIPE_SCALE_ALPHA, PIPE_RESTORE_ALPHA
};

struct RVertex {
    int64_t X, Y;
    float invw;
    float z;   // clip z / clip w
    float attr[4];
    uint32_t flat_u;
    bool ok;
    float clip[4];   // clip-space x, y, z, w (what frustum clipping works on)
};

struct DrawState {
    const cr_config* config;
    Framebuffer* fb;
    const DynamicStrokeDescriptor* stroke;
    uint32_t ref;
    uint32_t wmask, cmask;
    const float* color;   // instance colour (rgba) or null
    uint32_t save_layer, restore_layer;
    int y_begin, y_end;   // row band processed by this call
};

static const int SAMPLE_POS_1[1][2] = {{128, 128}};
static const int SAMPLE_POS_4[4][2] = {{96, 32}, {224, 96}, {32, 160}, {160, 224}};

inline RVertex transform_vertex(const float* m, const float pos[2], uint32_t W, uint32_t H) {
    RVertex v{};
    const float x = pos[0], y = pos[1];
    const float cx = (m[0] * x + m[4] * y) + m[12];
    const float cy = (m[1] * x + m[5] * y) + m[13];
    const float cw = (m[3] * x + m[7] * y) + m[15];
    const float cz = (m[2] * x + m[6] * y) + m[14];
    v.clip[0] = cx; v.clip[1] = cy; v.clip[2] = cz; v.clip[3] = cw;
    v.ok = cw > 0.0f;
    if (!v.ok) return v;
    v.invw = 1.0f / cw;
    v.z = cz / cw;
    const float fx = ((cx * v.invw) * 0.5f + 0.5f) * (float)W;
    const float fy = (0.5f - (cy * v.invw) * 0.5f) * (float)H;
    if (!(cr::fabs_f(fx) <= 2097152.0f) || !(cr::fabs_f(fy) <= 2097152.0f)) { v.ok = false; return v; }
    v.X = (int64_t)cr::floor_f(fx * 256.0f + 0.5f);
    v.Y = (int64_t)cr::floor_f(fy * 256.0f + 0.5f);
    return v;
}

// src/shaders.wgsl:165-189
inline bool cap_test(float tx, float ty, uint32_t cap_type) {
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
// src/shaders.wgsl:191-203
inline bool joint_test(float radius, bool bevel, uint32_t kind) {
    switch (kind) {
        case 1: return bevel;
        case 2: return radius <= 0.5f;
        default: return true;
    }
}
// src/shaders.wgsl:205-231
inline bool stroke_dashed(const DynamicStrokeDescriptor& d, float tx, float ty) {
    const uint32_t last_interval_index = d.count_dashed_join >> 3;
    const float pattern_length = d.gap_end[last_interval_index & 3u];
    uint32_t interval_index = 0;
    float gap_start, gap_end;
    float position_in_pattern = cr::wgsl_mod(ty - d.phase, pattern_length);
    if (position_in_pattern < 0.0f) position_in_pattern = position_in_pattern + pattern_length;
    for (;;) {
        gap_end = d.gap_end[interval_index & 3u] - position_in_pattern;
        if (gap_end >= 0.0f || interval_index >= last_interval_index) break;
        interval_index = interval_index + 1u;
    }
    gap_start = position_in_pattern - d.gap_start[interval_index & 3u];
    if (gap_start > 0.0f) {
        const uint32_t caps = d.caps >> (interval_index * 8u);
        const bool start_cap = cap_test(tx, gap_start, caps >> 4u);
        const bool end_cap = cap_test(tx, gap_end, caps);
        return start_cap || end_cap;
    }
    return true;
}

// Fragment predicate (src/shaders.wgsl:233-300): does this sample survive `sample_mask`?
inline bool fragment_keep(PipelineKind pipe, const DrawState& st, const float a[4], uint32_t flat_u, float flat_f) {
    switch (pipe) {
        case PIPE_FILL_IQ: return a[0] * a[0] - a[1] <= 0.0f;
        case PIPE_FILL_IC: return a[0] * a[0] * a[0] - a[1] * a[2] <= 0.0f;
        case PIPE_FILL_RQ: return a[0] * a[0] - a[1] * a[2] <= 0.0f;
        case PIPE_FILL_RC: return a[0] * a[0] * a[0] - a[1] * a[2] * a[3] <= 0.0f;
        case PIPE_STROKE_LINE: {
            const uint32_t path_index = flat_u & 65535u;
            const DynamicStrokeDescriptor& d = st.stroke[path_index];
            if ((d.count_dashed_join & 4u) != 0u) return stroke_dashed(d, a[0], a[1]);
            if ((flat_u & 65536u) != 0u) return cap_test(a[0], a[1] - flat_f, d.caps >> 4u);
            if (a[1] < 0.0f) return cap_test(a[0], -a[1], d.caps);
            return true;
        }
        case PIPE_STROKE_JOINT: {
            const float radius = cr::sqrt_f(a[0] * a[0] + a[1] * a[1]);
            const uint32_t path_index = flat_u & 65535u;
            const DynamicStrokeDescriptor& d = st.stroke[path_index];
            bool fill = joint_test(radius, (flat_u & 65536u) != 0u, d.count_dashed_join & 3u);
            const float TAU = 6.28318548202514648438f;  // acos(-1.0) * 2.0 in f32
            if (fill && (d.count_dashed_join & 4u) != 0u) fill = stroke_dashed(d, radius, a[2] + cr::atan2_f(a[1], a[0]) / TAU);
            return fill;
        }
        default: return true;
    }
}

inline bool depth_passes(uint32_t f, float z, float d) {   // wgpu::CompareFunction
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
inline float unorm8(float x) {
    x = x > 0.0f ? x : 0.0f;
    x = x < 1.0f ? x : 1.0f;
    return cr::floor_f(x * 255.0f + 0.5f) / 255.0f;
}

// Stencil test + stencil op + colour write for one covered sample (src/renderer.rs:571-861). `z`: the fragment's depth.
inline void apply_sample(PipelineKind pipe, const DrawState& st, bool front, size_t sample_index, float z) {
    Framebuffer& fb = *st.fb;
    const bool u8 = st.config->color_format != CR_FORMAT_RGBA32F;
    const uint32_t W = st.wmask, C = st.cmask, M = W | C;
    uint32_t s = fb.stencil[sample_index];
    const uint32_t ref = st.ref;
    float* px = fb.color + sample_index * 4;
    switch (pipe) {
        case PIPE_STROKE_LINE:
        case PIPE_STROKE_JOINT:
            if ((ref & M) == (s & M)) s = (s & ~W) | ((s + 1u) & W);
            break;
        case PIPE_FILL_SOLID: case PIPE_FILL_IQ: case PIPE_FILL_IC: case PIPE_FILL_RQ: case PIPE_FILL_RC:
            if ((ref & M) <= (s & M)) s = (s & ~W) | ((front ? s + 1u : s - 1u) & W);
            break;
        case PIPE_CLIP:
            if ((ref & W) != (s & W)) s = (s & ~M) | (ref & M);
            break;
        case PIPE_UNCLIP:
            if ((ref & C) < (s & C)) s = (s & ~M) | (ref & M);
            break;
        case PIPE_COLOR: {
            if ((ref & M) < (s & M)) {
                if (fb.depth) {
                    if (!depth_passes(st.config->depth_compare, z, fb.depth[sample_index])) break;   // depth_fail_op: Keep
                    if (st.config->depth_write_enabled) fb.depth[sample_index] = z;
                }
                const float a = st.color[3];
                const float src[4] = {st.color[0] * a, st.color[1] * a, st.color[2] * a, a};
                if (st.config->blending == CR_BLEND_PREMULTIPLIED_OVER) {
                    const float k = 1.0f - src[3];
                    for (int c = 0; c < 4; ++c) px[c] = src[c] + px[c] * k;
                } else {
                    for (int c = 0; c < 4; ++c) px[c] = src[c];
                }
                if (u8) for (int c = 0; c < 4; ++c) px[c] = unorm8(px[c]);
                fb.covered_samples += 1;
            }
            s = s & ~W;  // pass_op = fail_op = Zero on write_mask W
        } break;
        case PIPE_SAVE_ALPHA:
            if ((ref & M) <= (s & M)) fb.alpha_layers[(size_t)st.save_layer * fb.width * fb.height * fb.samples + sample_index] = u8 ? unorm8(px[3]) : px[3];
            break;
        case PIPE_SCALE_ALPHA:
            if ((ref & M) <= (s & M)) {
                const float sa = 1.0f - st.color[3];
                px[3] = sa + px[3] * (1.0f - sa);
                if (u8) px[3] = unorm8(px[3]);
            }
            break;
        case PIPE_RESTORE_ALPHA:
            if ((ref & M) <= (s & M)) {
                const float saved = fb.alpha_layers[(size_t)st.restore_layer * fb.width * fb.height * fb.samples + sample_index];
                const float sa = (1.0f - saved) * (1.0f - st.color[3]);
                px[3] = px[3] - sa;
                if (u8) px[3] = unorm8(px[3]);
            }
            break;
    }
    fb.stencil[sample_index] = (uint8_t)s;
}

inline bool is_top_left(int64_t dx, int64_t dy) { return (dy == 0 && dx > 0) || dy < 0; }

inline void rasterize_snapped(PipelineKind pipe, const DrawState& st, RVertex v0, RVertex v1, RVertex v2, bool odd, uint32_t flat_u, float flat_f) {
    if (!v0.ok || !v1.ok || !v2.ok) return;
    int64_t area2 = (v1.X - v0.X) * (v2.Y - v0.Y) - (v2.X - v0.X) * (v1.Y - v0.Y);
    if (area2 == 0) return;
    const bool front = (area2 < 0) != odd;
    if (pipe == PIPE_COLOR) {
        if (st.config->cull_mode == CR_CULL_BACK && !front) return;
        if (st.config->cull_mode == CR_CULL_FRONT && front) return;
    }
    if (area2 < 0) { std::swap(v1, v2); area2 = -area2; }
    Framebuffer& fb = *st.fb;
    const int64_t minX = std::min(v0.X, std::min(v1.X, v2.X)), maxX = std::max(v0.X, std::max(v1.X, v2.X));
    const int64_t minY = std::min(v0.Y, std::min(v1.Y, v2.Y)), maxY = std::max(v0.Y, std::max(v1.Y, v2.Y));
    auto floor_div256 = [](int64_t v) { return v >= 0 ? v / 256 : -((-v + 255) / 256); };
    int64_t px0 = std::max<int64_t>(0, floor_div256(minX)), px1 = std::min<int64_t>((int64_t)fb.width - 1, floor_div256(maxX));
    int64_t py0 = std::max<int64_t>(st.y_begin, floor_div256(minY)), py1 = std::min<int64_t>((int64_t)st.y_end - 1, floor_div256(maxY));
    const int (*spos)[2] = fb.samples == 4 ? SAMPLE_POS_4 : SAMPLE_POS_1;
    const RVertex* vs[3] = {&v0, &v1, &v2};
    int64_t A[3], B[3], bias[3];
    for (int e = 0; e < 3; ++e) {  // edge e goes from vs[e] to vs[(e+1)%3]
        const RVertex& a = *vs[e];
        const RVertex& b = *vs[(e + 1) % 3];
        A[e] = b.X - a.X;
        B[e] = b.Y - a.Y;
        bias[e] = is_top_left(A[e], B[e]) ? 0 : -1;
    }
    for (int64_t py = py0; py <= py1; ++py)
        for (int64_t px = px0; px <= px1; ++px)
            for (uint32_t sidx = 0; sidx < fb.samples; ++sidx) {
                const int64_t PX = px * 256 + spos[sidx][0], PY = py * 256 + spos[sidx][1];
                int64_t E[3];
                bool inside = true;
                for (int e = 0; e < 3; ++e) {
                    const RVertex& a = *vs[e];
                    E[e] = A[e] * (PY - a.Y) - B[e] * (PX - a.X);
                    if (E[e] + bias[e] < 0) inside = false;
                }
                if (!inside) continue;
                // barycentric weight of vs[i] is the edge function of the opposite edge: v0 <- E[1], v1 <- E[2], v2 <- E[0]
                const float e0 = (float)E[1] * v0.invw, e1 = (float)E[2] * v1.invw, e2 = (float)E[0] * v2.invw;
                const float den = (e0 + e1) + e2;
                float a[4];
                for (int k = 0; k < 4; ++k) a[k] = ((e0 * v0.attr[k] + e1 * v1.attr[k]) + e2 * v2.attr[k]) / den;
                if (!fragment_keep(pipe, st, a, flat_u, flat_f)) continue;
                const size_t sample_index = ((size_t)py * fb.width + (size_t)px) * fb.samples + sidx;
                float z = 0.0f;
                if (pipe == PIPE_COLOR && fb.depth) {
                    const float b0 = (float)E[1], b1 = (float)E[2], b2 = (float)E[0];
                    z = ((b0 * v0.z + b1 * v1.z) + b2 * v2.z) / ((b0 + b1) + b2);
                }
                apply_sample(pipe, st, front, sample_index, z);
            }
}

// ---- frustum clipping (see the contract at the top of this file)
struct ClipV { float x, y, z, w; float attr[4]; };
inline float clip_distance(const ClipV& v, int plane, float G) {
    switch (plane) {
        case 0: return v.w - 1.0e-6f;
        case 1: return G * v.w - v.x;
        case 2: return G * v.w + v.x;
        case 3: return G * v.w - v.y;
        default: return G * v.w + v.y;
    }
}
inline ClipV clip_meet(const ClipV& in, const ClipV& out, float din, float dout) {
    const float t = din / (din - dout);
    ClipV r;
    r.x = in.x + (out.x - in.x) * t;
    r.y = in.y + (out.y - in.y) * t;
    r.z = in.z + (out.z - in.z) * t;
    r.w = in.w + (out.w - in.w) * t;
    for (int k = 0; k < 4; ++k) r.attr[k] = in.attr[k] + (out.attr[k] - in.attr[k]) * t;
    return r;
}
inline RVertex snap_clipped(const ClipV& c, uint32_t W, uint32_t H) {
    RVertex v{};
    v.clip[0] = c.x; v.clip[1] = c.y; v.clip[2] = c.z; v.clip[3] = c.w;
    for (int k = 0; k < 4; ++k) v.attr[k] = c.attr[k];
    v.ok = c.w > 0.0f;
    if (!v.ok) return v;
    v.invw = 1.0f / c.w;
    v.z = c.z / c.w;
    const float fx = ((c.x * v.invw) * 0.5f + 0.5f) * (float)W;
    const float fy = (0.5f - (c.y * v.invw) * 0.5f) * (float)H;
    if (!(cr::fabs_f(fx) <= 2097152.0f) || !(cr::fabs_f(fy) <= 2097152.0f)) { v.ok = false; return v; }
    v.X = (int64_t)cr::floor_f(fx * 256.0f + 0.5f);
    v.Y = (int64_t)cr::floor_f(fy * 256.0f + 0.5f);
    return v;
}
inline void rasterize_triangle(PipelineKind pipe, const DrawState& st, const RVertex& v0, const RVertex& v1, const RVertex& v2, bool odd) {
    const uint32_t flat_u = v0.flat_u;
    const float flat_f = v0.attr[1];
    if (v0.ok && v1.ok && v2.ok) { rasterize_snapped(pipe, st, v0, v1, v2, odd, flat_u, flat_f); return; }
    const uint32_t W = st.fb->width, H = st.fb->height;
    const uint32_t g = 1048576u / std::max(W, H);
    const float G = (float)(g ? g : 1u);
    std::vector<ClipV> poly, next;
    for (const RVertex* v : {&v0, &v1, &v2}) {
        ClipV c{v->clip[0], v->clip[1], v->clip[2], v->clip[3], {v->attr[0], v->attr[1], v->attr[2], v->attr[3]}};
        poly.push_back(c);
    }
    for (int plane = 0; plane < 5; ++plane) {
        next.clear();
        for (size_t i = 0; i < poly.size(); ++i) {
            const ClipV& cur = poly[i];
            const ClipV& nxt = poly[(i + 1) % poly.size()];
            const float dc = clip_distance(cur, plane, G), dn = clip_distance(nxt, plane, G);
            const bool cin = dc >= 0.0f, nin = dn >= 0.0f;
            if (cin) next.push_back(cur);
            if (cin != nin) next.push_back(cin ? clip_meet(cur, nxt, dc, dn) : clip_meet(nxt, cur, dn, dc));
        }
        poly.swap(next);
        if (poly.size() < 3) return;
    }
    for (size_t j = 1; j + 1 < poly.size(); ++j)
        rasterize_snapped(pipe, st, snap_clipped(poly[0], W, H), snap_clipped(poly[j], W, H), snap_clipped(poly[j + 1], W, H), odd, flat_u, flat_f);
}

struct VertexFormat { size_t stride; int n_attr; bool has_u; };
static const VertexFormat FORMATS[8] = {
    {20, 2, true}, {24, 3, true}, {8, 0, false}, {16, 2, false}, {20, 3, false}, {20, 3, false}, {24, 4, false}, {8, 0, false}};

inline RVertex fetch(const RasterShape& sh, int category, uint64_t index, const float* m, uint32_t W, uint32_t H) {
    const VertexFormat& f = FORMATS[category];
    const uint64_t base = category == 0 ? 0 : sh.vertex_offsets[category - 1];
    const uint64_t count = (sh.vertex_offsets[category] - base) / f.stride;
    if (index >= count) { RVertex bad{}; bad.ok = false; return bad; }
    const uint8_t* p = sh.vertex_buffer + base + index * f.stride;
    float pos[2];
    std::memcpy(pos, p, 8);
    RVertex v = transform_vertex(m, pos, W, H);
    for (int k = 0; k < 4; ++k) v.attr[k] = 0.0f;
    std::memcpy(v.attr, p + 8, 4 * f.n_attr);
    v.flat_u = 0;
    if (f.has_u) std::memcpy(&v.flat_u, p + 8 + 4 * f.n_attr, 4);
    return v;
}

inline void draw_indexed_strip(PipelineKind pipe, const DrawState& st, const RasterShape& sh, int category, const uint32_t* idx, uint64_t n,
                               const float* m) {
    uint64_t strip_pos = 0;  // position inside the current strip
    for (uint64_t i = 0; i < n; ++i) {
        if (idx[i] == 0xFFFFFFFFu) { strip_pos = 0; continue; }
        if (strip_pos >= 2) {
            const RVertex a = fetch(sh, category, idx[i - 2], m, st.fb->width, st.fb->height);
            const RVertex b = fetch(sh, category, idx[i - 1], m, st.fb->width, st.fb->height);
            const RVertex c = fetch(sh, category, idx[i], m, st.fb->width, st.fb->height);
            rasterize_triangle(pipe, st, a, b, c, ((strip_pos - 2) & 1) != 0);
        }
        ++strip_pos;
    }
}
inline void draw_list(PipelineKind pipe, const DrawState& st, const RasterShape& sh, int category, const float* m) {
    const VertexFormat& f = FORMATS[category];
    const uint64_t base = sh.vertex_offsets[category - 1];
    const uint64_t count = (sh.vertex_offsets[category] - base) / f.stride;
    for (uint64_t i = 0; i + 2 < count; i += 3)
        rasterize_triangle(pipe, st, fetch(sh, category, i, m, st.fb->width, st.fb->height), fetch(sh, category, i + 1, m, st.fb->width, st.fb->height),
                           fetch(sh, category, i + 2, m, st.fb->width, st.fb->height), false);
}
inline void draw_hull(PipelineKind pipe, const DrawState& st, const RasterShape& sh, const float* m) {
    const uint64_t count = (sh.vertex_offsets[7] - sh.vertex_offsets[6]) / 8;
    for (uint64_t i = 0; i + 2 < count; ++i)
        rasterize_triangle(pipe, st, fetch(sh, 7, i, m, st.fb->width, st.fb->height), fetch(sh, 7, i + 1, m, st.fb->width, st.fb->height),
                           fetch(sh, 7, i + 2, m, st.fb->width, st.fb->height), (i & 1) != 0);
}

// One band [y_begin, y_end) of the whole command stream. Commands are validated by the caller.
inline void render_band(const cr_config& config, Framebuffer& fb, const RasterShape* shapes, const RenderCommand* cmds, size_t n_cmds,
                        const float* transforms, const float* colors, int y_begin, int y_end) {
    DrawState st{};
    st.config = &config;
    st.fb = &fb;
    st.wmask = (1u << config.winding_counter_bits) - 1u;
    st.cmask = ((1u << config.clip_nesting_counter_bits) - 1u) << config.winding_counter_bits;
    st.y_begin = y_begin;
    st.y_end = y_end;
    for (size_t c = 0; c < n_cmds; ++c) {
        const RenderCommand& cmd = cmds[c];
        const RasterShape& sh = shapes[cmd.shape];
        st.ref = cmd.clip_depth << config.winding_counter_bits;
        st.stroke = sh.stroke;
        st.save_layer = cmd.save_layer;
        st.restore_layer = cmd.restore_layer;
        if (cmd.operation == CR_OP_STENCIL) {
            // src/renderer.rs:275-336: one instanced draw per vertex category, in this order.
            if (sh.n_groups > 0) {
                if (sh.vertex_offsets[0] > 0)
                    for (uint32_t i = cmd.instance_begin; i < cmd.instance_end; ++i)
                        draw_indexed_strip(PIPE_STROKE_LINE, st, sh, 0, sh.wide_indices, sh.index_counts[0], transforms + 16 * (size_t)i);
                if (sh.vertex_offsets[0] < sh.vertex_offsets[1])
                    for (uint32_t i = cmd.instance_begin; i < cmd.instance_end; ++i)
                        draw_indexed_strip(PIPE_STROKE_JOINT, st, sh, 1, sh.wide_indices + sh.index_counts[0], sh.index_counts[1], transforms + 16 * (size_t)i);
            }
            if (sh.vertex_offsets[1] < sh.vertex_offsets[2])
                for (uint32_t i = cmd.instance_begin; i < cmd.instance_end; ++i)
                    draw_indexed_strip(PIPE_FILL_SOLID, st, sh, 2, sh.wide_indices + sh.index_counts[0] + sh.index_counts[1], sh.index_counts[2],
                                       transforms + 16 * (size_t)i);
            static const PipelineKind curve_pipes[4] = {PIPE_FILL_IQ, PIPE_FILL_IC, PIPE_FILL_RQ, PIPE_FILL_RC};
            for (int k = 0; k < 4; ++k)
                if (sh.vertex_offsets[k + 2] < sh.vertex_offsets[k + 3])
                    for (uint32_t i = cmd.instance_begin; i < cmd.instance_end; ++i) draw_list(curve_pipes[k], st, sh, k + 3, transforms + 16 * (size_t)i);
            continue;
        }
        PipelineKind pipe;
        bool needs_color = false;
        switch (cmd.operation) {
            case CR_OP_CLIP: pipe = PIPE_CLIP; break;
            case CR_OP_UNCLIP: pipe = PIPE_UNCLIP; break;
            case CR_OP_COLOR: pipe = PIPE_COLOR; needs_color = true; break;
            case CR_OP_SAVE_ALPHA_CONTEXT: pipe = PIPE_SAVE_ALPHA; break;
            case CR_OP_SCALE_ALPHA_CONTEXT: pipe = PIPE_SCALE_ALPHA; needs_color = true; break;
            default: pipe = PIPE_RESTORE_ALPHA; needs_color = true; break;
        }
        for (uint32_t i = cmd.instance_begin; i < cmd.instance_end; ++i) {
            st.color = needs_color ? colors + 4 * (size_t)i : nullptr;
            draw_hull(pipe, st, sh, transforms + 16 * (size_t)i);
        }
    }
}

}  // namespace oracle
