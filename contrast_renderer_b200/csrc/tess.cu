// tess.cu — K1 (stroke subdivision / hull emission) and K2 (fill fan + Loop-Blinn classification) for sm_100a,
// plus the scans that size the output and the per-shape convex hull.
//
// Replaces, on the device: src/curve.rs (all), src/stroke.rs (all), src/fill.rs (all), src/convex_hull.rs,
// src/vertex.rs and the buffer concatenation of Shape::from_paths (src/renderer.rs:184-209).
//
// Structure: a counting pass (one thread per Path, with a counting sink), an exclusive scan over paths, then the emit pass that
// writes every vertex / index / proto-hull point to its final address:
//   * filled paths without cubic segments: fill_segments_kernel, one thread per SEGMENT (the fill builder's only running state
//     are counts, which warp ballots provide);
//   * stroked paths and fills with cubics: tess_emit_kernel, one thread walks one Path exactly like StrokeBuilder::add_path /
//     FillBuilder::add_path do (the sequential state — arc length, previous tangent, strip cuts, the five per-type cursors —
//     lives in registers). The per-sample root solving of interpolate_normal! (src/curve.rs:228-252) only runs in the emit pass.
// Then, per Shape, convex_hull::andrew: a shared-memory bitonic sort and the two monotone chains (one lane per chain).
//
// Output layout in HBM: one array per vertex category for the WHOLE batch (all shapes), categories in
// concat_buffers! order; a shape's slice of category c is [cat_begin[c][s], cat_begin[c][s+1]). Index buffers are
// 32-bit: (vertex index within the shape's category << 1) | strip parity, CR_RESTART between strips.
#include "device_common.cuh"
#include "tess.h"
#include "prims.h"

using namespace crd;

namespace {

// ================================================================================================ curve.rs
__device__ __forceinline__ Pt mvt1(const Pt* pb, float a0) { return pb[0] * a0; }
__device__ __forceinline__ Pt mvt2(const Pt* pb, float a0, float a1) { return pb[0] * a0 + pb[1] * a1; }
__device__ __forceinline__ Pt mvt3(const Pt* pb, float a0, float a1, float a2) { return pb[0] * a0 + (pb[1] * a1 + pb[2] * a2); }
__device__ __forceinline__ Pt mvt4(const Pt* pb, float a0, float a1, float a2, float a3) {
    return pb[0] * a0 + (pb[1] * a1 + (pb[2] * a2 + pb[3] * a3));
}
// src/curve.rs:26-42
__device__ __forceinline__ void quad_power_basis(const Pt* cp, Pt* pb) {
    pb[0] = mvt1(cp, 1.0f);
    pb[1] = mvt2(cp, -2.0f, 2.0f);
    pb[2] = mvt3(cp, 1.0f, -2.0f, 1.0f);
}
__device__ __forceinline__ void cubic_power_basis(const Pt* cp, Pt* pb) {
    pb[0] = mvt1(cp, 1.0f);
    pb[1] = mvt2(cp, -3.0f, 3.0f);
    pb[2] = mvt3(cp, 3.0f, -6.0f, 3.0f);
    pb[3] = mvt4(cp, -1.0f, 3.0f, -3.0f, 1.0f);
}
// src/curve.rs:58-83
__device__ void reparametrize_cubic(const Pt* pb, float a, float b, Pt* out) {
    const float a2 = a * a, a3 = a * a * a, b2 = b * b, b3 = b * b * b;
    const float amb = a - b;
    out[0] = mvt4(pb, 1.0f, a, a2, a3);
    out[1] = mvt4(pb, 0.0f, b - a, -2.0f * a2 + 2.0f * a * b, 3.0f * a2 * b - 3.0f * a3);
    out[2] = mvt4(pb, 0.0f, 0.0f, amb * amb, -6.0f * a2 * b + 3.0f * a * b2 + 3.0f * a3);
    out[3] = mvt4(pb, 0.0f, 0.0f, 0.0f, 3.0f * a2 * b - 3.0f * a * b2 - a3 + b3);
}
// src/curve.rs:86-114
__device__ __forceinline__ Pt quad_point(const Pt* pb, float t) { return mvt3(pb, 1.0f, t, t * t); }
__device__ __forceinline__ Ln quad_d1(const Pt* pb, float t) { return join(mvt3(pb, 1.0f, t, t * t), mvt3(pb, 0.0f, 1.0f, 2.0f * t)); }
__device__ __forceinline__ Pt cubic_point(const Pt* pb, float t) { return mvt4(pb, 1.0f, t, t * t, t * t * t); }
__device__ __forceinline__ Ln cubic_d1(const Pt* pb, float t) {
    return join(mvt4(pb, 1.0f, t, t * t, t * t * t), mvt4(pb, 0.0f, 1.0f, 2.0f * t, 3.0f * (t * t)));
}
// src/curve.rs:133-144
__device__ void ippc_of(const Pt* pb, bool integral, float* ippc) {
    ippc[0] = 0.0f;
    if (!integral) ippc[0] = triple(pb[1], pb[2], pb[3]) * -1.0f;
    ippc[1] = triple(pb[0], pb[2], pb[3]) * 1.0f;
    ippc[2] = triple(pb[0], pb[1], pb[3]) * -1.0f;
    ippc[3] = triple(pb[0], pb[1], pb[2]) * 1.0f;
    const float m = cr::sqrt_f(ippc[0] * ippc[0] + ippc[1] * ippc[1] + ippc[2] * ippc[2] + ippc[3] * ippc[3]);
    const float inv = 1.0f / m;
    ippc[0] = ippc[0] * inv; ippc[1] = ippc[1] * inv; ippc[2] = ippc[2] * inv; ippc[3] = ippc[3] * inv;
}
struct Inflections { float disc; cr::Root r[3]; };
// src/curve.rs:151-190
__device__ Inflections integral_inflections(const float* ippc, bool loop_self_intersection) {
    Inflections o;
    o.r[1] = cr::no_root(); o.r[2] = cr::no_root();
    const float disc = 3.0f * (ippc[2] * ippc[2]) - 4.0f * ippc[1] * ippc[3];
    if (cr::fabs_f(ippc[1]) <= CR_ERROR_MARGIN) {
        if (cr::fabs_f(ippc[2]) <= CR_ERROR_MARGIN) { o.disc = -1.0f; o.r[0] = cr::make_root(-1.0f, 0.0f, 1.0f); }
        else { o.disc = 1.0f; o.r[0] = cr::make_root(ippc[3], 0.0f, 3.0f * ippc[2]); }
        return o;
    }
    const float d = cr::sqrt_f(disc * (disc < 0.0f ? (loop_self_intersection ? -1.0f : 0.0f) : 1.0f / 3.0f));
    o.disc = disc;
    o.r[0] = cr::make_root(ippc[2] + d, 0.0f, 2.0f * ippc[1]);
    o.r[1] = cr::make_root(ippc[2] - d, 0.0f, 2.0f * ippc[1]);
    return o;
}
// src/curve.rs:197-226
__device__ Inflections rational_inflections(const float* ippc, bool loop_self_intersection) {
    if (cr::fabs_f(ippc[0]) <= CR_ERROR_MARGIN) return integral_inflections(ippc, loop_self_intersection);
    const cr::Roots cubic = cr::solve_cubic(ippc[3] * -1.0f, ippc[2] * 3.0f, ippc[1] * -3.0f, ippc[0], CR_ERROR_MARGIN);
    Inflections o;
    o.r[0] = cubic.r[0]; o.r[1] = cubic.r[1]; o.r[2] = cubic.r[2];
    if (!loop_self_intersection) { o.disc = cubic.discriminant; return o; }
    const cr::Roots hess = cr::solve_quadratic(ippc[1] * ippc[3] - ippc[2] * ippc[2], ippc[1] * ippc[2] - ippc[0] * ippc[3],
                                               ippc[0] * ippc[2] - ippc[1] * ippc[1], CR_ERROR_MARGIN);
    if (hess.discriminant > 0.0f) {
        o.r[2] = cubic.real_root == 0 ? cubic.r[0] : (cubic.real_root == 1 ? cubic.r[1] : cubic.r[2]);
        if (hess.count == 2) { o.r[0] = hess.r[0]; o.r[1] = hess.r[1]; }
        else if (hess.count == 1) { o.r[0] = hess.r[0]; o.r[1] = cr::no_root(); }
    }
    o.disc = -hess.discriminant;
    return o;
}

// Which polynomial interpolate_normal! solves for one curve kind (src/curve.rs:318,342,368,405).
enum { K_IQ = 0, K_IC = 1, K_RQ = 2, K_RC = 3 };
struct SolvePlanes { Ln p[5]; };
template <int KIND>
__device__ __forceinline__ float first_root_in_unit(const SolvePlanes& sp, Ln normal) {
    cr::Roots r;
    if (KIND == K_IQ) r = cr::solve_linear(dot(normal, sp.p[0]), dot(normal, sp.p[1]), CR_ERROR_MARGIN);
    else if (KIND == K_IC) r = cr::solve_quadratic(dot(normal, sp.p[0]), dot(normal, sp.p[1]), dot(normal, sp.p[2]), CR_ERROR_MARGIN);
    else if (KIND == K_RQ) {
        const Ln n = rot90cw(normal);
        r = cr::solve_quadratic(dot(n, sp.p[0]), dot(n, sp.p[1]), dot(n, sp.p[2]), CR_ERROR_MARGIN);
    } else {
        const Ln n = rot90cw(normal);
        r = cr::solve_quartic(dot(n, sp.p[0]), dot(n, sp.p[1]), dot(n, sp.p[2]), dot(n, sp.p[3]), dot(n, sp.p[4]), CR_ERROR_MARGIN);
    }
    for (int k = 0; k < r.count; ++k) {
        if (r.r[k].denominator == 0.0f) continue;
        const float parameter = r.r[k].numerator.re / r.r[k].denominator;
        if (parameter >= 0.0f && parameter <= 1.0f) return parameter;
    }
    return 0.0f;
}
// The polar interpolation set-up of interpolate_normal! (src/curve.rs:230-234).
struct Polar { cr::Complex start, step; uint32_t steps; };
__device__ __forceinline__ uint32_t polar_steps(Ln st, Ln et, float angle_step, cr::Complex* range_out, cr::Complex* start_out) {
    const cr::Complex ps = cr::cplx(st.g1, st.g2), pe = cr::cplx(et.g1, et.g2);
    const cr::Complex range = cr::cdiv(pe, ps);
    *range_out = range;
    *start_out = ps;
    return cr::f32_to_usize_sat(cr::fabs_f(cr::carg(range) / angle_step) + 0.5f);
}

// ============================================================================================== output sinks
struct PathOffsets { uint32_t v[CNT_COUNT]; };   // absolute begin of this path's slice in every counter's array

template <bool EMIT>
struct Sink {
    uint32_t n[CNT_COUNT];
    uint32_t strip_start;
    // emit-only state
    TessOutput out;
    PathOffsets base;
    uint32_t shape_base[3];    // the shape's first vertex in CAT_LINE / CAT_JOINT / CAT_SOLID (index values are shape-relative)
    uint32_t solid_total;      // S.len() of this path (known from the count pass) for the fan->strip scatter
    uint32_t err;

    __device__ void init() {
#pragma unroll
        for (int i = 0; i < CNT_COUNT; ++i) n[i] = 0;
        strip_start = 0;
        err = 0;
    }
    __device__ __forceinline__ void proto(float2 p) {
        if (EMIT) {
            if (!cr::is_finite(p.x) || !cr::is_finite(p.y)) err |= CR_DEVERR_NON_FINITE;
            out.proto[base.v[CNT_PROTO] + n[CNT_PROTO]] = make_float2(cr::canon_zero(p.x), cr::canon_zero(p.y));
        }
        n[CNT_PROTO] += 1;
    }
    // Vertex2f1i + its proto_hull entry (pushed at the next cut_stroke_polygon, src/stroke.rs:125)
    __device__ __forceinline__ void line_vertex(float2 p, float side, float along, uint32_t flags) {
        if (EMIT) {
            uint32_t* w = reinterpret_cast<uint32_t*>(out.vtx[CAT_LINE]) + (size_t)(base.v[CAT_LINE] + n[CAT_LINE]) * 5;
            w[0] = __float_as_uint(p.x); w[1] = __float_as_uint(p.y); w[2] = __float_as_uint(side); w[3] = __float_as_uint(along); w[4] = flags;
        }
        n[CAT_LINE] += 1;
        proto(p);
    }
    // cut_stroke_polygon (src/stroke.rs:123-132)
    __device__ void cut() {
        const uint32_t count = n[CAT_LINE] - strip_start;
        if (count == 0) return;
        if (EMIT) {
            uint32_t* idx = out.idx[0] + base.v[CNT_LINE_IDX] + n[CNT_LINE_IDX];
            const uint32_t rel = base.v[CAT_LINE] + strip_start - shape_base[0];
            for (uint32_t i = 0; i < count; ++i) idx[i] = ((rel + i) << 1) | (i & 1u);
            idx[count] = CR_RESTART;
        }
        n[CNT_LINE_IDX] += count + 1;
        strip_start = n[CAT_LINE];
    }
    __device__ __forceinline__ void joint_vertex(float2 p, float t0, float t1, float t2, uint32_t flags) {
        if (EMIT) {
            uint32_t* w = reinterpret_cast<uint32_t*>(out.vtx[CAT_JOINT]) + (size_t)(base.v[CAT_JOINT] + n[CAT_JOINT]) * 6;
            w[0] = __float_as_uint(p.x); w[1] = __float_as_uint(p.y); w[2] = __float_as_uint(t0); w[3] = __float_as_uint(t1);
            w[4] = __float_as_uint(t2); w[5] = flags;
        }
        n[CAT_JOINT] += 1;
    }
    __device__ void joint_indices() {  // the five vertices just written
        if (EMIT) {
            uint32_t* idx = out.idx[1] + base.v[CNT_JOINT_IDX] + n[CNT_JOINT_IDX];
            const uint32_t rel = base.v[CAT_JOINT] + n[CAT_JOINT] - 5 - shape_base[1];
            for (uint32_t i = 0; i < 5; ++i) idx[i] = ((rel + i) << 1) | (i & 1u);
            idx[5] = CR_RESTART;
        }
        n[CNT_JOINT_IDX] += 6;
    }
    // path_solid_vertices.push(..): written straight to its triangle_fan_to_strip position (src/vertex.rs:28-35)
    __device__ __forceinline__ void solid(float2 p) {
        if (EMIT) {
            const uint32_t j = n[CAT_SOLID], total = solid_total;
            const uint32_t i = (j < (total + 1) / 2) ? 2 * j : 2 * (total - 1 - j) + 1;
            reinterpret_cast<float2*>(out.vtx[CAT_SOLID])[base.v[CAT_SOLID] + i] = p;
        }
        n[CAT_SOLID] += 1;
    }
    __device__ void solid_indices() {
        const uint32_t count = n[CAT_SOLID];
        if (EMIT) {
            uint32_t* idx = out.idx[2] + base.v[CNT_SOLID_IDX];
            const uint32_t rel = base.v[CAT_SOLID] - shape_base[2];
            for (uint32_t i = 0; i < count; ++i) idx[i] = ((rel + i) << 1) | (i & 1u);
            idx[count] = CR_RESTART;
        }
        n[CNT_SOLID_IDX] += count + 1;
    }
    template <int CAT, int NW>
    __device__ __forceinline__ void curve_vertex(float2 p, const float* w) {
        if (EMIT) {
            float* dst = reinterpret_cast<float*>(out.vtx[CAT]) + (size_t)(base.v[CAT] + n[CAT]) * (2 + NW);
            dst[0] = p.x; dst[1] = p.y;
#pragma unroll
            for (int k = 0; k < NW; ++k) dst[2 + k] = w[k];
        }
        n[CAT] += 1;
    }
};

// ================================================================================================ stroke.rs
struct StrokeCtx {
    float width, offset, miter_clip;
    uint32_t group;
    float length;   // length_accumulator
};
// src/stroke.rs:18-22
__device__ __forceinline__ Pt offset_cp(Pt cp, Ln tangent, float offset) { return cp + mk_pt(0.0f, tangent.g1, tangent.g2) * offset; }
// emit_stroke_vertices (src/stroke.rs:28-51)
template <bool EMIT>
__device__ __forceinline__ void emit_stroke_vertices(Sink<EMIT>& s, const StrokeCtx& c, uint32_t flags, float length, Pt pt, Ln tangent) {
    if (EMIT) {
        const float along = length / c.width;
        s.line_vertex(to_vec(offset_cp(pt, tangent, (c.offset - 0.5f) * c.width)), -0.5f, along, flags);
        s.line_vertex(to_vec(offset_cp(pt, tangent, (c.offset + 0.5f) * c.width)), 0.5f, along, flags);
    } else {
        s.n[CAT_LINE] += 2;
        s.n[CNT_PROTO] += 2;
    }
}
// emit_stroke_join (src/stroke.rs:53-121)
template <bool EMIT>
__device__ void emit_stroke_join(Sink<EMIT>& s, StrokeCtx& c, Pt cp, Ln prev, Ln next) {
    const float d = dot(prev, next);
    if (cr::fabs_f(d - 1.0f) <= CR_ERROR_MARGIN) return;
    const float side_sign = cr::rust_signum(meet(prev, next).g0);
    const float miter_clip = c.width * c.miter_clip;
    const float side_offset = (c.offset - side_sign * 0.5f) * c.width;
    const Pt pe = offset_cp(cp, prev, side_offset);
    const Pt ne = offset_cp(cp, next, side_offset);
    const Ln pl = parallel_through(prev, pe);
    const Ln nl = parallel_through(next, ne);
    const Pt x = intersect(pl, nl);
    Pt v3 = x, v4 = x;
    const bool anti = cr::fabs_f(d + 1.0f) <= CR_ERROR_MARGIN;
    if (anti || mag(join(cp, x)) > miter_clip) {
        const Ln mid = anti ? neg(rot90cw(prev)) : unit(prev + next);
        const Pt cv = offset_cp(cp, mid, -side_sign * miter_clip);
        const Ln cl = parallel_through(mid, cv);
        v3 = intersect(pl, cl);
        v4 = intersect(cl, nl);
        s.proto(to_vec(v3));
        s.proto(to_vec(v4));
    } else {
        s.proto(to_vec(v3));
    }
    if (EMIT) {
        const Ln st = prev * (1.0f / -c.width);
        const float along = c.length / c.width;
        const Pt vs[5] = {cp, pe, ne, v3, v4};
#pragma unroll
        for (int k = 0; k < 5; ++k)
            s.joint_vertex(to_vec(vs[k]), side_sign * incidence(vs[k], st), dot(join(vs[k], cp), st), along, c.group);
    } else {
        s.n[CAT_JOINT] += 5;
    }
    s.joint_indices();
    c.length += cr::acos_f(d) / (3.14159274101257324219f * 2.0f) * c.width;
    s.cut();
    emit_stroke_vertices(s, c, c.group, c.length, cp, next);
}
// get_quadratic_tangents / get_cubic_tangents (src/stroke.rs:179-202)
__device__ __forceinline__ void quad_tangents(Pt a, Pt b, Pt c, Ln& s, Ln& e) {
    s = unit(join(a, b));
    e = unit(join(b, c));
    if (cr::is_nan(s.g0) || cr::is_nan(e.g0)) { s = unit(join(a, c)); e = s; }
}
__device__ __forceinline__ void cubic_tangents(Pt a, Pt b, Pt c, Pt d, Ln& s, Ln& e) {
    s = unit(join(a, b));
    if (cr::is_nan(s.g0)) s = unit(join(a, c));
    e = unit(join(c, d));
    if (cr::is_nan(e.g0)) e = unit(join(b, d));
    if (cr::is_nan(s.g0) || cr::is_nan(e.g0)) e = unit(join(a, d));
}

// One sample of emit_curve_stroke! (src/stroke.rs:143-166).
template <bool CUBIC>
__device__ __forceinline__ void stroke_sample(Sink<true>& s, StrokeCtx& c, const Pt* pb, float t, Pt& previous_point) {  // emit pass only
    Ln tangent = CUBIC ? cubic_d1(pb, t) : quad_d1(pb, t);
    if (sqmag(tangent) == 0.0f) {
        if (t < 0.5f) t += CR_F32_EPSILON; else t -= CR_F32_EPSILON;
        tangent = CUBIC ? cubic_d1(pb, t) : quad_d1(pb, t);
    }
    tangent = unit(tangent);
    Pt pt = CUBIC ? cubic_point(pb, t) : quad_point(pb, t);
    pt = pt * (1.0f / pt.g0);
    c.length += mag(join(previous_point, pt));
    emit_stroke_vertices(s, c, c.group, c.length, pt, tangent);
    previous_point = pt;
}
// interpolate_normal! has no limit (src/curve.rs:228-252); the device keeps one: 2^22 samples per inflection-free interval keep
// every 32-bit count of a path far from wrapping. (Up to round 1 the limit was the 256-entry parameter buffer of the cubic.)
__device__ __forceinline__ uint32_t clamp_steps(uint32_t steps, uint32_t& err) {
    if (steps > CR_MAX_STEPS_PER_INTERVAL) { err |= CR_DEVERR_STEPS; return CR_MAX_STEPS_PER_INTERVAL; }
    return steps;
}
#define CR_PARAM_BUFFER 256u   // cubic parameters of one interval that are sorted in local memory; longer intervals sort in the output array

// Quadratic curve body: *_quadratic_uniform_tangent_angle (src/curve.rs:306-322,355-380) + emit_curve_stroke!.
template <bool EMIT, int KIND>
__device__ void stroke_quadratic(Sink<EMIT>& s, StrokeCtx& c, const Pt* pb, Ln st, Ln et, bool uta, float angle_step, uint32_t usp_steps, Pt start) {
    if (!uta) {
        if constexpr (EMIT) {
            Pt prev = start;
            for (uint32_t i = 1; i < usp_steps + 1; ++i) stroke_sample<false>(s, c, pb, (float)i / (float)usp_steps, prev);
        } else { s.n[CAT_LINE] += 2 * usp_steps; s.n[CNT_PROTO] += 2 * usp_steps; }
        return;
    }
    cr::Complex range, pstart;
    const uint32_t steps = clamp_steps(polar_steps(st, et, angle_step, &range, &pstart), s.err);
    const uint32_t n_params = (steps >= 2 ? steps - 1 : 0) + 1;
    if constexpr (!EMIT) { s.n[CAT_LINE] += 2 * n_params; s.n[CNT_PROTO] += 2 * n_params; return; } else {
    SolvePlanes sp;
    if (KIND == K_IQ) { sp.p[0] = dual(pb[1]); sp.p[1] = dual(pb[2]) * 2.0f; }
    else { sp.p[0] = join(pb[1], pb[0]); sp.p[1] = join(pb[2], pb[0]) * 2.0f; sp.p[2] = join(pb[2], pb[1]); }
    Pt prev = start;
    if (steps >= 2) {
        const cr::Complex pstep = cr::cpowf(range, 1.0f / (float)steps);
        for (uint32_t i = 1; i < steps; ++i) {
            const cr::Complex ip = cr::cmul(pstart, cr::cpowi(pstep, i));
            const float t = first_root_in_unit<KIND>(sp, mk_ln(0.0f, ip.re, ip.im));
            stroke_sample<false>(s, c, pb, t, prev);
        }
    }
    stroke_sample<false>(s, c, pb, 1.0f, prev);
    }
}

// Cubic curve body: cubic_uniform_tangent_angle! (src/curve.rs:254-303) + emit_curve_stroke!.
template <bool EMIT, int KIND>
__device__ void stroke_cubic(Sink<EMIT>& s, StrokeCtx& c, const Pt* pb, bool uta, float angle_step, uint32_t usp_steps, Pt start) {
    if (!uta) {
        if constexpr (EMIT) {
            Pt prev = start;
            for (uint32_t i = 1; i < usp_steps + 1; ++i) stroke_sample<true>(s, c, pb, (float)i / (float)usp_steps, prev);
        } else { s.n[CAT_LINE] += 2 * usp_steps; s.n[CNT_PROTO] += 2 * usp_steps; }
        return;
    }
    float ippc[4];
    ippc_of(pb, KIND == K_IC, ippc);
    const Inflections inf = (KIND == K_IC) ? integral_inflections(ippc, false) : rational_inflections(ippc, false);
    // split parameters: roots in [0,1], sorted, near-duplicates (< ERROR_MARGIN apart) removed
    float split[3];
    int n_split = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (inf.r[k].denominator == 0.0f) continue;
        const float p = inf.r[k].numerator.re / inf.r[k].denominator;
        if (p >= 0.0f && p <= 1.0f) {
            int j = n_split++;
            while (j > 0 && split[j - 1] > p) { split[j] = split[j - 1]; --j; }
            split[j] = p;
        }
    }
    {
        int i = 1;
        while (i < n_split) {
            if (split[i] - split[i - 1] < CR_ERROR_MARGIN) { for (int j = i; j + 1 < n_split; ++j) split[j] = split[j + 1]; --n_split; }
            else ++i;
        }
    }
    const bool cusp = cr::fabs_f(inf.disc) < CR_ERROR_MARGIN;
    float previous_split = 0.0f;
    Pt prev = start;
    float params[EMIT ? CR_PARAM_BUFFER : 1];
    (void)params;
    for (int iv = 0; iv <= n_split; ++iv) {
        float a = previous_split, b = 1.0f;
        if (iv < n_split) {
            if (cusp) { b = split[iv] - CR_F32_EPSILON; previous_split = split[iv] + CR_F32_EPSILON; }
            else { b = split[iv]; previous_split = split[iv]; }
        }
        const Ln st = unit(cubic_d1(pb, a));
        const Ln et = unit(cubic_d1(pb, b));
        cr::Complex range, pstart;
        const uint32_t steps = clamp_steps(polar_steps(st, et, angle_step, &range, &pstart), s.err);
        const uint32_t n_params = (steps >= 2 ? steps - 1 : 0) + 1;
        if constexpr (!EMIT) { s.n[CAT_LINE] += 2 * n_params; s.n[CNT_PROTO] += 2 * n_params; continue; } else {
        uint32_t np = 0;
        if (steps >= 2) {
            Pt tr[4];
            reparametrize_cubic(pb, a, b, tr);
            SolvePlanes sp;
            if (KIND == K_IC) { sp.p[0] = dual(tr[1]); sp.p[1] = dual(tr[2]) * 2.0f; sp.p[2] = dual(tr[3]) * 3.0f; }
            else {
                sp.p[0] = join(tr[1], tr[0]);
                sp.p[1] = join(tr[2], tr[0]) * 2.0f;
                sp.p[2] = join(tr[2], tr[1]) + join(tr[3], tr[0]) * 3.0f;
                sp.p[3] = join(tr[3], tr[1]) * 2.0f;
                sp.p[4] = join(tr[3], tr[2]);
            }
            const cr::Complex pstep = cr::cpowf(range, 1.0f / (float)steps);
            auto parameter = [&](uint32_t i) -> float {
                const cr::Complex ip = cr::cmul(pstart, cr::cpowi(pstep, i));
                return a + (b - a) * first_root_in_unit<KIND>(sp, mk_ln(0.0f, ip.re, ip.im));
            };
            if (steps - 1u <= CR_PARAM_BUFFER) {
                for (uint32_t i = 1; i < steps; ++i) {
                    const float mapped = parameter(i);
                    uint32_t j = np++;   // stable insertion sort (src/curve.rs:297)
                    while (j > 0 && params[j - 1] > mapped) { params[j] = params[j - 1]; --j; }
                    params[j] = mapped;
                }
            } else {
                // More parameters than the local buffer holds: they are sorted (src/curve.rs:297, stable insertion sort — they come out
                // nearly ordered) in GLOBAL memory, in the tail of the very vertex slots this interval is about to fill: its
                // steps samples own 40 bytes each in the line-vertex array, the steps - 1 floats sit in the last bytes of that
                // range, and sample i is read before its vertices are written; the write front (40 bytes per sample) never
                // reaches a parameter that has not been consumed (36 np + 4 (i + 1) >= 36 i + 40 bytes for every i < np).
                const uint32_t np_long = steps - 1u;
                float* const gp = reinterpret_cast<float*>(s.out.vtx[CAT_LINE]) + (size_t)5 * (s.base.v[CAT_LINE] + s.n[CAT_LINE] + 2u * steps) - np_long;
                for (uint32_t i = 1; i < steps; ++i) {
                    const float mapped = parameter(i);
                    uint32_t j = i - 1u;
                    while (j > 0 && gp[j - 1] > mapped) { gp[j] = gp[j - 1]; --j; }
                    gp[j] = mapped;
                }
                for (uint32_t i = 0; i < np_long; ++i) { const float t = gp[i]; stroke_sample<true>(s, c, pb, t, prev); }
            }
        }
        for (uint32_t i = 0; i < np; ++i) stroke_sample<true>(s, c, pb, params[i], prev);
        stroke_sample<true>(s, c, pb, b, prev);
        }
    }
}

struct PathView {
    float2 start;
    const uint8_t* seg_types;
    uint32_t n_segments;
    const float* seg[5];   // per-type arrays positioned at this path's first segment of that type
    uint32_t count[5];     // segments of each type in this path's slices (from the cursor tables): no read goes beyond them
    cr_stroke_options so;
};
// A segment of type T is about to be read with cursor `cur`: it must lie inside the path's slice of that type's array.
#define CR_SEGMENT_IN_SLICE(pv, T, cur) ((cur) < (pv).count[T])

// StrokeBuilder::add_path (src/stroke.rs:205-465)
template <bool EMIT>
__device__ void stroke_path(Sink<EMIT>& s, const PathView& pv) {
    StrokeCtx c;
    c.width = pv.so.width; c.offset = pv.so.offset; c.miter_clip = pv.so.miter_clip;
    c.group = pv.so.dynamic_stroke_options_group;
    c.length = 0.0f;
    const bool closed = (pv.so.flags & CR_STROKE_FLAG_CLOSED) != 0;
    const bool uta = (pv.so.flags & CR_STROKE_FLAG_UNIFORM_TANGENT_ANGLE) != 0;
    const float angle_step = pv.so.approximation.angle_step;
    const uint32_t usp = pv.so.approximation.steps;
    Pt prev_cp = from_vec(pv.start.x, pv.start.y);
    Ln first_tangent = mk_ln(0.0f, 0.0f, 0.0f), prev_tangent = mk_ln(0.0f, 0.0f, 0.0f);
    uint32_t cur[5] = {0, 0, 0, 0, 0};
    bool is_first = true;
    for (uint32_t si = 0; si < pv.n_segments; ++si) {
        const uint32_t type = pv.seg_types[si];
        Pt next_cp;
        Ln st, et;
        const float* d = nullptr;
        if (type > 4u || !((type == 0 && CR_SEGMENT_IN_SLICE(pv, 0, cur[0])) || (type == 1 && CR_SEGMENT_IN_SLICE(pv, 1, cur[1])) || (type == 2 && CR_SEGMENT_IN_SLICE(pv, 2, cur[2])) ||
                           (type == 3 && CR_SEGMENT_IN_SLICE(pv, 3, cur[3])) || (type == 4 && CR_SEGMENT_IN_SLICE(pv, 4, cur[4])))) {
            s.err |= CR_DEVERR_BAD_TABLES;   // the type stream and the per-type cursor tables disagree
            break;
        }
        switch (type) {
            case CR_SEG_LINE:
                d = pv.seg[0] + 2 * (size_t)cur[0]++;
                next_cp = from_vec(d[0], d[1]);
                st = unit(join(prev_cp, next_cp));
                et = st;
                break;
            case CR_SEG_INTEGRAL_QUADRATIC:
                d = pv.seg[1] + 4 * (size_t)cur[1];
                next_cp = from_vec(d[2], d[3]);
                quad_tangents(prev_cp, from_vec(d[0], d[1]), next_cp, st, et);
                break;
            case CR_SEG_INTEGRAL_CUBIC:
                d = pv.seg[2] + 6 * (size_t)cur[2];
                next_cp = from_vec(d[4], d[5]);
                cubic_tangents(prev_cp, from_vec(d[0], d[1]), from_vec(d[2], d[3]), next_cp, st, et);
                break;
            case CR_SEG_RATIONAL_QUADRATIC:
                d = pv.seg[3] + 5 * (size_t)cur[3];
                next_cp = from_vec(d[3], d[4]);
                quad_tangents(prev_cp, from_vec(d[1], d[2]), next_cp, st, et);
                break;
            default:
                d = pv.seg[4] + 10 * (size_t)cur[4];
                next_cp = from_vec(d[8], d[9]);
                cubic_tangents(prev_cp, from_vec(d[4], d[5]), from_vec(d[6], d[7]), next_cp, st, et);
                break;
        }
        if (cr::is_nan(st.g0) || cr::is_nan(et.g0)) continue;  // cursors of curve types are NOT advanced (quirk C.3)
        if (is_first) {
            is_first = false;
            first_tangent = st;
            if (!closed) {
                const Ln normal = rot90cw(st);
                emit_stroke_vertices(s, c, c.group, c.length - 0.5f * c.width, offset_cp(prev_cp, normal, 0.5f * cr::fabs_f(c.width)), st);
            }
            if (closed || type != CR_SEG_LINE) emit_stroke_vertices(s, c, c.group, c.length, prev_cp, st);
        } else {
            emit_stroke_join(s, c, prev_cp, prev_tangent, st);
        }
        switch (type) {
            case CR_SEG_LINE:
                c.length += mag(join(prev_cp, next_cp));
                emit_stroke_vertices(s, c, c.group, c.length, next_cp, et);
                break;
            case CR_SEG_INTEGRAL_QUADRATIC: {
                cur[1]++;
                const Pt cp[3] = {prev_cp, from_vec(d[0], d[1]), from_vec(d[2], d[3])};
                Pt pb[3];
                quad_power_basis(cp, pb);
                stroke_quadratic<EMIT, K_IQ>(s, c, pb, st, et, uta, angle_step, usp, prev_cp);
            } break;
            case CR_SEG_INTEGRAL_CUBIC: {
                cur[2]++;
                const Pt cp[4] = {prev_cp, from_vec(d[0], d[1]), from_vec(d[2], d[3]), from_vec(d[4], d[5])};
                Pt pb[4];
                cubic_power_basis(cp, pb);
                stroke_cubic<EMIT, K_IC>(s, c, pb, uta, angle_step, usp, prev_cp);
            } break;
            case CR_SEG_RATIONAL_QUADRATIC: {
                cur[3]++;
                const Pt cp[3] = {prev_cp, from_wvec(d[0], d[1], d[2]), from_vec(d[3], d[4])};
                Pt pb[3];
                quad_power_basis(cp, pb);
                stroke_quadratic<EMIT, K_RQ>(s, c, pb, st, et, uta, angle_step, usp, prev_cp);
            } break;
            default: {
                cur[4]++;
                const float2 pv2 = to_vec(prev_cp);
                const Pt cp[4] = {from_wvec(d[0], pv2.x, pv2.y), from_wvec(d[1], d[4], d[5]), from_wvec(d[2], d[6], d[7]), from_wvec(d[3], d[8], d[9])};
                Pt pb[4];
                cubic_power_basis(cp, pb);
                stroke_cubic<EMIT, K_RC>(s, c, pb, uta, angle_step, usp, prev_cp);
            } break;
        }
        prev_cp = next_cp;
        prev_tangent = et;
    }
    if (closed) {
        const Pt start = from_vec(pv.start.x, pv.start.y);
        const Ln line = join(prev_cp, start);
        const float length = mag(line);
        if (length > 0.0f) {
            const Ln tangent = line * (1.0f / length);
            emit_stroke_join(s, c, prev_cp, prev_tangent, tangent);
            c.length += length;
            emit_stroke_vertices(s, c, c.group, c.length, start, tangent);
            emit_stroke_join(s, c, start, tangent, first_tangent);
        } else {
            emit_stroke_join(s, c, start, prev_tangent, first_tangent);
        }
    } else {
        s.cut();
        emit_stroke_vertices(s, c, c.group | 0x10000u, c.length, prev_cp, prev_tangent);
        const Ln normal = rot90cw(prev_tangent);
        emit_stroke_vertices(s, c, c.group | 0x10000u, c.length + 0.5f * c.width, offset_cp(prev_cp, normal, -0.5f * cr::fabs_f(c.width)), prev_tangent);
    }
    s.cut();
}

// ================================================================================================== fill.rs
struct W4 { float k, l, m, n; };
__device__ __forceinline__ float implicit_value(W4 w) { return w.k * w.k * w.k - w.l * w.m * w.n; }
// weight_derivatives (src/fill.rs:34-49) for one column; returns the four Bernstein-blossom weights.
__device__ void weight_column(const cr::Root& r0, const cr::Root& r1, const cr::Root& r2, float* col) {
    const float n0 = r0.numerator.re, n1 = r1.numerator.re, n2 = r2.numerator.re;
    const float d0 = r0.denominator, d1 = r1.denominator, d2 = r2.denominator;
    const float p0 = n0 * n1 * n2;
    const float p1 = -d0 * n1 * n2 - n0 * d1 * n2 - n0 * n1 * d2;
    const float p2 = n0 * d1 * d2 + d0 * n1 * d2 + d0 * d1 * n2;
    const float p3 = -d0 * d1 * d2;
    col[0] = p0;
    col[1] = p0 + p1 * 1.0f / 3.0f;
    col[2] = p0 + p1 * 2.0f / 3.0f + p2 * 1.0f / 3.0f;
    col[3] = p0 + p1 + p2 + p3;
}
__device__ __forceinline__ float det3(float a0, float a1, float a2, float b0, float b1, float b2, float c0, float c1, float c2) {
    return a0 * (b1 * c2 - b2 * c1) - a1 * (b0 * c2 - b2 * c0) + a2 * (b0 * c1 - b1 * c0);
}
// weight_planes (src/fill.rs:70-85) for one weight column: plane through three (w, wx, wy, weight) points.
__device__ Ln weight_plane(const Pt* cp, const float* col) {
    float pl[4];
    for (int attempt = 0; attempt < 2; ++attempt) {
        const int third = attempt == 0 ? 2 : 3;
        const float P[4] = {cp[0].g0, cp[0].g1, cp[0].g2, col[0]};
        const float Q[4] = {cp[1].g0, cp[1].g1, cp[1].g2, col[1]};
        const float R[4] = {cp[third].g0, cp[third].g1, cp[third].g2, col[third]};
        pl[0] = det3(P[1], P[2], P[3], Q[1], Q[2], Q[3], R[1], R[2], R[3]);
        pl[1] = -det3(P[0], P[2], P[3], Q[0], Q[2], Q[3], R[0], R[2], R[3]);
        pl[2] = det3(P[0], P[1], P[3], Q[0], Q[1], Q[3], R[0], R[1], R[3]);
        pl[3] = -det3(P[0], P[1], P[2], Q[0], Q[1], Q[2], R[0], R[1], R[2]);
        if (!(pl[1] * pl[1] + pl[2] * pl[2] + pl[3] * pl[3] < CR_ERROR_MARGIN)) break;
    }
    const float sc = 1.0f / -pl[3];
    return mk_ln(pl[0] * sc, pl[1] * sc, pl[2] * sc);
}

// emit_cubic_curve_triangle! (src/fill.rs:116-132)
template <bool EMIT, bool RATIONAL>
__device__ void cubic_triangle(Sink<EMIT>& s, const float* areas, const Pt* cp, const W4* w, int skip) {
    const float area = areas[skip];
    if (!(cr::fabs_f(area) > CR_ERROR_MARGIN)) return;
    int idx[3], n = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) if (i != skip) idx[n++] = i;
    if (area < 0.0f) { const int t = idx[0]; idx[0] = idx[2]; idx[2] = t; }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const W4 ww = w[idx[k]];
        const float wv[4] = {ww.k, ww.l, ww.m, ww.n};
        if (RATIONAL) s.template curve_vertex<CAT_RC, 4>(to_vec(cp[idx[k]]), wv);
        else s.template curve_vertex<CAT_IC, 3>(to_vec(cp[idx[k]]), wv);
    }
}
// triangulate_cubic_curve_quadrilateral! (src/fill.rs:134-204)
template <bool EMIT, bool RATIONAL>
__device__ void cubic_quadrilateral(Sink<EMIT>& s, const Pt* cp, W4* w) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float sc = 1.0f / cp[j].g0;
        w[j].k *= sc; w[j].l *= sc; w[j].m *= sc; w[j].n *= sc;
    }
    float areas[4];
    areas[0] = triple(cp[1], cp[2], cp[3]);
    areas[1] = triple(cp[0], cp[2], cp[3]);
    areas[2] = triple(cp[0], cp[1], cp[3]);
    areas[3] = triple(cp[0], cp[1], cp[2]);
    const float sum = cr::fabs_f(areas[0]) + cr::fabs_f(areas[1]) + cr::fabs_f(areas[2]) + cr::fabs_f(areas[3]);
    int enclosing = -1;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float equilibrium = 0.5f * sum;
        if (cr::fabs_f(equilibrium - cr::fabs_f(areas[i])) <= CR_ERROR_MARGIN) enclosing = (enclosing == -1) ? i : -1;
    }
    if (enclosing >= 0) {
        cubic_triangle<EMIT, RATIONAL>(s, areas, cp, w, enclosing);
    } else {
        int opposite = 0;
        bool bad = false;
#pragma unroll
        for (int j = 1; j < 4; ++j) {
            const float side_of_a = areas[j];
            const float side_of_d = areas[0] * (j == 2 ? -1.0f : 1.0f);
            if (side_of_a * side_of_d < 0.0f) { if (opposite != 0) bad = true; opposite = j; }
        }
        if (opposite == 0 || bad) { s.err |= CR_DEVERR_CUBIC; return; }   // the reference panics here (src/fill.rs:174,178)
        cubic_triangle<EMIT, RATIONAL>(s, areas, cp, w, 0);
        cubic_triangle<EMIT, RATIONAL>(s, areas, cp, w, opposite);
    }
    const bool add1 = enclosing != 1 && implicit_value(w[1]) < 0.0f;
    const bool add2 = enclosing != 2 && implicit_value(w[2]) < 0.0f;
    if (add1 && add2) {
        if (areas[0] * areas[1] < 0.0f) { s.solid(to_vec(cp[2])); s.solid(to_vec(cp[1])); }
        else { s.solid(to_vec(cp[1])); s.solid(to_vec(cp[2])); }
    } else if (add1) s.solid(to_vec(cp[1]));
    else if (add2) s.solid(to_vec(cp[2]));
}
__device__ __forceinline__ Pt lerp_pt(Pt a, Pt b, float t) { return a * (1.0f - t) + b * t; }
__device__ __forceinline__ W4 lerp_w(W4 a, W4 b, float t) {
    W4 o;
    o.k = a.k * (1.0f - t) + b.k * t; o.l = a.l * (1.0f - t) + b.l * t; o.m = a.m * (1.0f - t) + b.m * t; o.n = a.n * (1.0f - t) + b.n * t;
    return o;
}
// emit_cubic_curve! (src/fill.rs:218-250)
template <bool EMIT, bool RATIONAL>
__device__ void fill_cubic(Sink<EMIT>& s, const Pt* cp) {
    Pt pb[4];
    cubic_power_basis(cp, pb);
    float ippc[4];
    ippc_of(pb, !RATIONAL, ippc);
    const Inflections inf = RATIONAL ? rational_inflections(ippc, true) : integral_inflections(ippc, true);
    // weights (src/fill.rs:51-68): columns k, l, m, n
    float col[4][4];
    if (inf.disc == 0.0f) {
        weight_column(inf.r[0], inf.r[0], inf.r[2], col[0]);
        weight_column(inf.r[0], inf.r[0], inf.r[0], col[1]);
        weight_column(inf.r[0], inf.r[0], inf.r[0], col[2]);
    } else if (inf.disc < 0.0f) {
        weight_column(inf.r[0], inf.r[1], inf.r[2], col[0]);
        weight_column(inf.r[0], inf.r[0], inf.r[1], col[1]);
        weight_column(inf.r[1], inf.r[1], inf.r[0], col[2]);
    } else {
        weight_column(inf.r[0], inf.r[1], inf.r[2], col[0]);
        weight_column(inf.r[0], inf.r[0], inf.r[0], col[1]);
        weight_column(inf.r[1], inf.r[1], inf.r[1], col[2]);
    }
    weight_column(inf.r[2], inf.r[2], inf.r[2], col[3]);
    // gradient of k^3 - l m n at control point 0 (src/fill.rs:91-96) and side normalisation (:98-114)
    const Ln p0 = weight_plane(cp, col[0]), p1 = weight_plane(cp, col[1]), p2 = weight_plane(cp, col[2]), p3 = weight_plane(cp, col[3]);
    const float wk = col[0][0], wl = col[1][0], wm = col[2][0], wn = col[3][0];
    const Ln gradient = p0 * (3.0f * wk * wk) - p1 * (wm * wn) - p2 * (wl * wn) - p3 * (wl * wm);
    const Ln tangent = cubic_d1(pb, 0.0f);
    const float flip = dot(tangent, gradient) > 0.0f ? -1.0f : 1.0f;
    W4 w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        w[j].k = col[0][j]; w[j].l = col[1][j]; w[j].m = col[2][j]; w[j].n = col[3][j];
        if (flip < 0.0f) { w[j].k *= -1.0f; w[j].l *= -1.0f; }
    }
    // find_double_point_issue (src/fill.rs:14-32)
    float param = -1.0f;
    int inside = 0;
    if (inf.disc < 0.0f) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (inf.r[k].denominator != 0.0f) {
                const float p = inf.r[k].numerator.re / inf.r[k].denominator;
                if (0.0f < p && p < 1.0f) { param = p; inside += 1; }
            }
        }
    }
    if (inside == 1) {
        // split_curve_at! (src/fill.rs:206-216) on control points and on weights
        const Pt p10 = lerp_pt(cp[0], cp[1], param), p11 = lerp_pt(cp[1], cp[2], param), p12 = lerp_pt(cp[2], cp[3], param);
        const Pt p20 = lerp_pt(p10, p11, param), p21 = lerp_pt(p11, p12, param);
        const Pt p30 = lerp_pt(p20, p21, param);
        const W4 w10 = lerp_w(w[0], w[1], param), w11 = lerp_w(w[1], w[2], param), w12 = lerp_w(w[2], w[3], param);
        const W4 w20 = lerp_w(w10, w11, param), w21 = lerp_w(w11, w12, param);
        const W4 w30 = lerp_w(w20, w21, param);
        const Pt cpa[4] = {cp[0], p10, p20, p30};
        W4 wa[4] = {w[0], w10, w20, w30};
        cubic_quadrilateral<EMIT, RATIONAL>(s, cpa, wa);
        const Pt cpb[4] = {p30, p21, p12, cp[3]};
        s.solid(to_vec(cpb[0]));
        W4 wb[4] = {w30, w21, w12, w[3]};
#pragma unroll
        for (int j = 0; j < 4; ++j) { wb[j].k *= -1.0f; wb[j].l *= -1.0f; }
        cubic_quadrilateral<EMIT, RATIONAL>(s, cpb, wb);
    } else {
        cubic_quadrilateral<EMIT, RATIONAL>(s, cp, w);
    }
    s.proto(to_vec(cp[1]));
    s.proto(to_vec(cp[2]));
    s.proto(to_vec(cp[3]));
    s.solid(to_vec(cp[3]));
}

// FillBuilder::add_path (src/fill.rs:263-367)
template <bool EMIT, bool CUBICS>
__device__ void fill_path(Sink<EMIT>& s, const PathView& pv) {   // CUBICS == false: the batch holds no cubic segment (checked on the host), their builder is compiled out
    float2 last = pv.start;
    s.solid(last);
    s.proto(last);
    uint32_t cur[5] = {0, 0, 0, 0, 0};
    for (uint32_t si = 0; si < pv.n_segments; ++si) {
        const uint32_t type = pv.seg_types[si];
        if (type > 4u || !((type == 0 && CR_SEGMENT_IN_SLICE(pv, 0, cur[0])) || (type == 1 && CR_SEGMENT_IN_SLICE(pv, 1, cur[1])) || (type == 2 && CR_SEGMENT_IN_SLICE(pv, 2, cur[2])) ||
                           (type == 3 && CR_SEGMENT_IN_SLICE(pv, 3, cur[3])) || (type == 4 && CR_SEGMENT_IN_SLICE(pv, 4, cur[4])))) {
            s.err |= CR_DEVERR_BAD_TABLES;   // the type stream and the per-type cursor tables disagree
            break;
        }
        switch (type) {
            case CR_SEG_LINE: {
                const float* d = pv.seg[0] + 2 * (size_t)cur[0]++;
                last = make_float2(d[0], d[1]);
                s.proto(last);
                s.solid(last);
            } break;
            case CR_SEG_INTEGRAL_QUADRATIC: {
                const float* d = pv.seg[1] + 4 * (size_t)cur[1]++;
                const float w0[2] = {1.0f, 1.0f}, w1[2] = {0.5f, 0.0f}, w2[2] = {0.0f, 0.0f};
                s.template curve_vertex<CAT_IQ, 2>(make_float2(d[2], d[3]), w0);
                s.template curve_vertex<CAT_IQ, 2>(make_float2(d[0], d[1]), w1);
                s.template curve_vertex<CAT_IQ, 2>(last, w2);
                s.proto(make_float2(d[0], d[1]));
                s.proto(make_float2(d[2], d[3]));
                last = make_float2(d[2], d[3]);
                s.solid(last);
            } break;
            case CR_SEG_INTEGRAL_CUBIC: {
                const float* d = pv.seg[2] + 6 * (size_t)cur[2]++;
                const Pt cp[4] = {from_vec(last.x, last.y), from_vec(d[0], d[1]), from_vec(d[2], d[3]), from_vec(d[4], d[5])};
                if (CUBICS) fill_cubic<EMIT, false>(s, cp); else s.err |= CR_DEVERR_MODE;
                last = to_vec(cp[3]);
            } break;
            case CR_SEG_RATIONAL_QUADRATIC: {
                const float* d = pv.seg[3] + 5 * (size_t)cur[3]++;
                const float weight = 1.0f / d[0];
                const float w0[3] = {1.0f, 1.0f, 1.0f}, w1[3] = {0.5f * weight, 0.0f, weight}, w2[3] = {0.0f, 0.0f, 1.0f};
                s.template curve_vertex<CAT_RQ, 3>(make_float2(d[3], d[4]), w0);
                s.template curve_vertex<CAT_RQ, 3>(make_float2(d[1], d[2]), w1);
                s.template curve_vertex<CAT_RQ, 3>(last, w2);
                s.proto(make_float2(d[1], d[2]));
                s.proto(make_float2(d[3], d[4]));
                last = make_float2(d[3], d[4]);
                s.solid(last);
            } break;
            default: {
                const float* d = pv.seg[4] + 10 * (size_t)cur[4]++;
                const Pt cp[4] = {from_wvec(d[0], last.x, last.y), from_wvec(d[1], d[4], d[5]), from_wvec(d[2], d[6], d[7]), from_wvec(d[3], d[8], d[9])};
                if (CUBICS) fill_cubic<EMIT, true>(s, cp); else s.err |= CR_DEVERR_MODE;
                last = to_vec(cp[3]);
            } break;
        }
    }
    // a filled path consumes exactly its slices (a stroked one may not: degenerate curves leave their cursor where it is, quirk C.3)
    if (cur[0] != pv.count[0] || cur[1] != pv.count[1] || cur[2] != pv.count[2] || cur[3] != pv.count[3] || cur[4] != pv.count[4]) s.err |= CR_DEVERR_BAD_TABLES;
    s.solid_indices();
}

__device__ __forceinline__ PathView load_path(const DevicePaths& P, uint32_t p) {
    PathView pv;
    pv.start = make_float2(P.start[2 * (size_t)p], P.start[2 * (size_t)p + 1]);
    const uint32_t sb = P.segment_begin[p];
    pv.seg_types = P.segment_types + sb;
    pv.n_segments = P.segment_begin[p + 1] - sb;
    const size_t stride = (size_t)P.n_paths + 1;
    pv.seg[0] = P.seg[0] + 2 * (size_t)P.type_begin[0 * stride + p];
    pv.seg[1] = P.seg[1] + 4 * (size_t)P.type_begin[1 * stride + p];
    pv.seg[2] = P.seg[2] + 6 * (size_t)P.type_begin[2 * stride + p];
    pv.seg[3] = P.seg[3] + 5 * (size_t)P.type_begin[3 * stride + p];
    pv.seg[4] = P.seg[4] + 10 * (size_t)P.type_begin[4 * stride + p];
#pragma unroll
    for (int t = 0; t < 5; ++t) pv.count[t] = P.type_begin[t * stride + p + 1] - P.type_begin[t * stride + p];
    if (P.stroke_options) pv.so = P.stroke_options[p];
    else pv.so = cr_stroke_options{};   // no stroke options: a filled Path
    return pv;
}

// The cursor tables of cr_path_soa come from the caller; everything below indexes with them. A Path of the reference cannot
// be inconsistent (its vectors carry their own lengths, src/path.rs:213-230), a C caller's tables can: check, before any
// segment is read, that the path's slice of the type stream lies inside [0, n_segments], that every per-type cursor runs
// forwards and that the per-type slice lengths add up to the path's segment count. While the path is walked every segment
// read is checked against its slice (CR_SEGMENT_IN_SLICE), so no read leaves the arrays even if the type stream disagrees
// with the tables. Any violation sets CR_DEVERR_BAD_TABLES and nothing is emitted.
__device__ bool path_tables_valid(const DevicePaths& P, uint32_t p) {
    const uint32_t sb = P.segment_begin[p], se = P.segment_begin[p + 1];
    if (sb > se || se > P.n_segments) return false;
    if (p == 0 && sb != 0) return false;
    if (p + 1 == P.n_paths && se != P.n_segments) return false;
    const size_t stride = (size_t)P.n_paths + 1;
    uint32_t total = 0;
#pragma unroll
    for (int t = 0; t < 5; ++t) {
        const uint32_t tb = P.type_begin[t * stride + p], te = P.type_begin[t * stride + p + 1];
        if (tb > te || (p == 0 && tb != 0)) return false;
        total += te - tb;
    }
    // the per-segment agreement of the type stream with these slices is checked where the segments are read (CR_SEGMENT_IN_SLICE)
    return total == se - sb;
}

// ------------------------------------------------------------------------------------------------- kernels
// Pass A: per-path output sizes. counts is [CNT_COUNT][n_paths + 1].
// MODE 0: anything. MODE 1: the batch has no stroke options at all (cr_path_soa.stroke_options == NULL, every Path is filled): the
// stroke builder is compiled out. MODE 2: filled paths of lines and quadratics only (text): the Loop-Blinn cubic builder with its
// binary64 solvers is compiled out as well (28 instead of 114 registers for the count pass: 33 -> 7 us on the text scene).
template <int MODE>
__global__ void __launch_bounds__(128) tess_count_kernel(DevicePaths P, uint32_t n_groups, uint32_t* __restrict__ counts, uint32_t* __restrict__ err) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.n_paths) return;
    const size_t stride = (size_t)P.n_paths + 1;
    if (!path_tables_valid(P, p)) {
#pragma unroll
        for (int c = 0; c < CNT_COUNT; ++c) counts[c * stride + p] = 0;
        atomicOr(err, CR_DEVERR_BAD_TABLES);
        return;
    }
    const PathView pv = load_path(P, p);
    Sink<false> s;
    s.init();
    if (MODE == 0 && (pv.so.flags & CR_STROKE_FLAG_STROKED)) {
        if (pv.so.dynamic_stroke_options_group >= n_groups) s.err |= CR_DEVERR_GROUP_OOB;
        else stroke_path<false>(s, pv);
    } else {
        fill_path<false, MODE != 2>(s, pv);
    }
#pragma unroll
    for (int c = 0; c < CNT_COUNT; ++c) counts[c * stride + p] = s.n[c];
    if (s.err) atomicOr(err, s.err);
}

// A filled path without cubic segments: tessellated by fill_segments_kernel (one thread per segment, see below).
__device__ __forceinline__ bool path_is_simple(const DevicePaths& P, uint32_t p) {
    const size_t stride = (size_t)P.n_paths + 1;
    if (P.stroke_options && (P.stroke_options[p].flags & CR_STROKE_FLAG_STROKED)) return false;
    return P.type_begin[2 * stride + p + 1] == P.type_begin[2 * stride + p] && P.type_begin[4 * stride + p + 1] == P.type_begin[4 * stride + p];
}

// Pass B: write everything. offsets is the exclusive scan of counts ([CNT_COUNT][n_paths + 1], total in the last slot).
template <int MODE>
__global__ void __launch_bounds__(128) tess_emit_kernel(DevicePaths P, const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ shape_path_begin,
                                                       uint32_t n_shapes, TessOutput out, uint32_t* __restrict__ err, uint32_t skip_simple) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.n_paths) return;
    if (*reinterpret_cast<volatile const uint32_t*>(err) & CR_DEVERR_FATAL_MASK) return;   // the count pass rejected the input (uniform): write nothing
    if (skip_simple && path_is_simple(P, p)) return;                                       // fill_segments_kernel tessellates it
    const PathView pv = load_path(P, p);
    const size_t stride = (size_t)P.n_paths + 1;
    Sink<true> s;
    s.init();
    s.out = out;
#pragma unroll
    for (int c = 0; c < CNT_COUNT; ++c) s.base.v[c] = offsets[c * stride + p];
    // shape of this path: last s with shape_path_begin[s] <= p
    uint32_t lo = 0, hi = n_shapes;
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (shape_path_begin[mid] <= p) lo = mid; else hi = mid; }
    const uint32_t first_path = shape_path_begin[lo];
    s.shape_base[0] = offsets[CAT_LINE * stride + first_path];
    s.shape_base[1] = offsets[CAT_JOINT * stride + first_path];
    s.shape_base[2] = offsets[CAT_SOLID * stride + first_path];
    s.solid_total = offsets[CAT_SOLID * stride + p + 1] - s.base.v[CAT_SOLID];
    if (MODE == 0 && (pv.so.flags & CR_STROKE_FLAG_STROKED)) stroke_path<true>(s, pv);
    else fill_path<true, MODE != 2>(s, pv);
    if (s.err) atomicOr(err, s.err);
}

// ---------------------------------------------------------------------------------- fills, one thread per SEGMENT
// FillBuilder::add_path (src/fill.rs:263-367) has no state that runs along the path except the solid-fan position and the
// per-type cursors, and both are COUNTS: segment i of a path writes solid vertex i + 1, its curve triangle at 3 x (number of
// earlier segments of its type), its proto-hull points behind 1 + lines + 2 x quadratics before it; the point it starts from is
// the end point of segment i - 1. So filled paths without cubic segments ("simple" paths: the Loop-Blinn builder emits a
// data-dependent number of vertices) are tessellated with one thread per segment: coalesced reads of the type stream and the
// per-type segment arrays, ranks by warp ballots (plus one cooperative count for a path that began before the warp), path
// records shared inside the warp by shuffles, and every thread writes at its final address. The count pass, the scan and the layout are those of the per-path kernels; paths that are
// stroked or hold cubics stay with tess_emit_kernel (which skips the simple ones when this kernel runs).
__device__ __forceinline__ uint32_t shape_of_path(const uint32_t* __restrict__ shape_path_begin, uint32_t n_shapes, uint32_t p) {
    uint32_t lo = 0, hi = n_shapes;   // last s with shape_path_begin[s] <= p
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (shape_path_begin[mid] <= p) lo = mid; else hi = mid; }
    return lo;
}
__device__ __forceinline__ void store_proto(float2* __restrict__ proto, uint32_t at, float2 p, uint32_t& err) {
    if (!cr::is_finite(p.x) || !cr::is_finite(p.y)) err |= CR_DEVERR_NON_FINITE;
    proto[at] = make_float2(cr::canon_zero(p.x), cr::canon_zero(p.y));
}
__device__ __forceinline__ uint32_t fan_to_strip_slot(uint32_t j, uint32_t total) { return (j < (total + 1) / 2) ? 2 * j : 2 * (total - 1 - j) + 1; }   // src/vertex.rs:28-35

#define FS_THREADS 256
// Everything a segment thread needs to know about its path, loaded once per CTA into shared memory (a CTA's 256 consecutive
// segments belong to a handful of consecutive paths).
struct FillPathInfo {
    uint32_t sb;                                   // segment_begin[p]
    uint32_t solid, proto, iq, rq, solid_idx;      // the path's slices in the output arrays (scan of the count pass)
    uint32_t tb0, tb1, tb3, te0, te1, te3;         // its slices of the line / integral quadratic / rational quadratic segment arrays
    uint32_t rel;                                  // shape-relative number of its first solid vertex
    uint32_t total_simple;                         // solid vertices (segments + 1) | simple << 31
};
// `shape_hint`: a shape at or before the path's (its own shape is found by walking forwards from there).
__device__ __forceinline__ FillPathInfo load_fill_path_info(const DevicePaths& P, const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ shape_path_begin,
                                                            uint32_t n_shapes, uint32_t p, bool all_simple, uint32_t shape_hint) {
    const size_t stride = (size_t)P.n_paths + 1;
    FillPathInfo f;
    f.sb = P.segment_begin[p];
    f.solid = offsets[CAT_SOLID * stride + p]; f.proto = offsets[CNT_PROTO * stride + p]; f.iq = offsets[CAT_IQ * stride + p]; f.rq = offsets[CAT_RQ * stride + p];
    f.solid_idx = offsets[CNT_SOLID_IDX * stride + p];
    f.tb0 = P.type_begin[0 * stride + p]; f.te0 = P.type_begin[0 * stride + p + 1];
    f.tb1 = P.type_begin[1 * stride + p]; f.te1 = P.type_begin[1 * stride + p + 1];
    f.tb3 = P.type_begin[3 * stride + p]; f.te3 = P.type_begin[3 * stride + p + 1];
    f.total_simple = (P.segment_begin[p + 1] - f.sb + 1u) | ((all_simple || path_is_simple(P, p)) ? 0x80000000u : 0u);
    uint32_t sh = shape_hint;
    while (sh + 1u < n_shapes && shape_path_begin[sh + 1u] <= p) ++sh;   // the CTA's paths lie in the hinted shape or the next few
    f.rel = f.solid - offsets[CAT_SOLID * stride + shape_path_begin[sh]];
    return f;
}

__global__ void __launch_bounds__(FS_THREADS) fill_segments_kernel(DevicePaths P, const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ shape_path_begin,
                                                                   uint32_t n_shapes, TessOutput out, uint32_t* __restrict__ err, uint32_t all_simple) {
    if (*reinterpret_cast<volatile const uint32_t*>(err) & CR_DEVERR_FATAL_MASK) return;   // the count pass rejected the input (uniform): write nothing
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31u;
    const uint32_t seg_threads = (P.n_segments + FS_THREADS - 1u) / FS_THREADS * FS_THREADS;   // whole CTAs of segment threads, then one thread per path for its start point
    uint32_t e = 0;
    if (tid >= seg_threads) {
        // ---- the path's first point: solid vertex 0, proto-hull point 0, index entry 0 and the strip's restart (src/fill.rs:264-266,366)
        const uint32_t p = tid - seg_threads, warp_p = min(p - lane, P.n_paths - 1u);
        const uint32_t hint = warp_search_last_le(shape_path_begin, n_shapes, warp_p, lane);   // shape of the warp's first path
        if (p >= P.n_paths) return;
        const FillPathInfo f = load_fill_path_info(P, offsets, shape_path_begin, n_shapes, p, all_simple != 0u, hint);
        if (!(f.total_simple >> 31)) return;
        const float2 start = make_float2(P.start[2 * (size_t)p], P.start[2 * (size_t)p + 1]);
        reinterpret_cast<float2*>(out.vtx[CAT_SOLID])[f.solid] = start;   // fan_to_strip_slot(0, total) == 0
        store_proto(out.proto, f.proto, start, e);
        uint32_t* idx = out.idx[2] + f.solid_idx;
        idx[0] = f.rel << 1;
        idx[f.total_simple & 0x7fffffffu] = CR_RESTART;
        if (e) atomicOr(err, e);
        return;
    }
    // ---- the warp's paths: the first one (and its shape) by a warp-wide search, then lane j loads everything about path first + j in
    // one round of independent loads; the paths that begin at or before the warp's last segment form a prefix of the lanes, and a
    // segment finds its path among them by a five-step search over the lanes (shuffles). No shared memory, no CTA barrier: the
    // warps of a CTA do not wait for each other's dependent loads.
    const uint32_t g = tid, g_warp = g - lane;
    if (g_warp >= P.n_segments) return;
    const uint32_t g_warp_last = min(g_warp + 32u, P.n_segments) - 1u;
    const uint32_t p_first = warp_search_last_le(P.segment_begin, P.n_paths, g_warp, lane);   // segment_begin[n_paths] = n_segments > g: empty paths share a value, the last one owns the segment
    const uint32_t s_first = warp_search_last_le(shape_path_begin, n_shapes, p_first, lane);
    FillPathInfo mine{};
    mine.sb = 0xFFFFFFFFu;
    if (p_first + lane < P.n_paths) mine = load_fill_path_info(P, offsets, shape_path_begin, n_shapes, p_first + lane, all_simple != 0u, s_first);
    const uint32_t n_local = (uint32_t)__popc(__ballot_sync(0xffffffffu, mine.sb <= g_warp_last));   // sb is non-decreasing: a prefix of the lanes
    const bool cached = n_local < 32u || p_first + 32u >= P.n_paths;   // else the warp's last segments may belong to paths beyond the loaded ones (runs of EMPTY paths)
    const bool live = g < P.n_segments;
    uint32_t p = 0xFFFFFFFFu, type = 255u;
    FillPathInfo f{};
    {
        uint32_t lo = 0;   // last loaded path with sb <= g
#pragma unroll
        for (uint32_t step = 16u; step != 0u; step >>= 1) {
            const uint32_t cand = lo + step;
            const uint32_t sb_cand = __shfl_sync(0xffffffffu, mine.sb, cand & 31u);
            if (cand < n_local && sb_cand <= g) lo = cand;
        }
        f.sb = __shfl_sync(0xffffffffu, mine.sb, lo);
        f.solid = __shfl_sync(0xffffffffu, mine.solid, lo); f.proto = __shfl_sync(0xffffffffu, mine.proto, lo);
        f.iq = __shfl_sync(0xffffffffu, mine.iq, lo); f.rq = __shfl_sync(0xffffffffu, mine.rq, lo);
        f.solid_idx = __shfl_sync(0xffffffffu, mine.solid_idx, lo);
        f.tb0 = __shfl_sync(0xffffffffu, mine.tb0, lo); f.tb1 = __shfl_sync(0xffffffffu, mine.tb1, lo); f.tb3 = __shfl_sync(0xffffffffu, mine.tb3, lo);
        f.te0 = __shfl_sync(0xffffffffu, mine.te0, lo); f.te1 = __shfl_sync(0xffffffffu, mine.te1, lo); f.te3 = __shfl_sync(0xffffffffu, mine.te3, lo);
        f.rel = __shfl_sync(0xffffffffu, mine.rel, lo); f.total_simple = __shfl_sync(0xffffffffu, mine.total_simple, lo);
        p = p_first + lo;
    }
    if (live) {
        if (!cached) {
            uint32_t lo = 0, hi = P.n_paths;
            while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (P.segment_begin[mid] <= g) lo = mid; else hi = mid; }
            p = lo;
            f = load_fill_path_info(P, offsets, shape_path_begin, n_shapes, p, all_simple != 0u, shape_of_path(shape_path_begin, n_shapes, p));
        }
        type = P.segment_types[g];
    } else p = 0xFFFFFFFFu;
    // ranks: how many segments of each type precede this one in its path. Inside the warp by ballots ...
    const uint32_t same = __match_any_sync(0xffffffffu, p) & ((1u << lane) - 1u);
    const uint32_t b0 = __ballot_sync(0xffffffffu, type == CR_SEG_LINE), b1 = __ballot_sync(0xffffffffu, type == CR_SEG_INTEGRAL_QUADRATIC),
                   b3 = __ballot_sync(0xffffffffu, type == CR_SEG_RATIONAL_QUADRATIC);
    uint32_t r0 = __popc(b0 & same), r1 = __popc(b1 & same), r3 = __popc(b3 & same);
    // ... and, for the one path that began before this warp's first segment, by a cooperative count over [its begin, the warp's first segment)
    const uint32_t p_lane0 = __shfl_sync(0xffffffffu, p, 0), sb_lane0 = __shfl_sync(0xffffffffu, f.sb, 0);
    if (p_lane0 != 0xFFFFFFFFu && sb_lane0 < g_warp) {
        uint32_t c0 = 0, c1 = 0, c3 = 0;
        for (uint32_t k = sb_lane0 + lane; k < g_warp; k += 32u) {
            const uint32_t t = P.segment_types[k];
            c0 += t == CR_SEG_LINE; c1 += t == CR_SEG_INTEGRAL_QUADRATIC; c3 += t == CR_SEG_RATIONAL_QUADRATIC;
        }
        c0 = __reduce_add_sync(0xffffffffu, c0); c1 = __reduce_add_sync(0xffffffffu, c1); c3 = __reduce_add_sync(0xffffffffu, c3);
        if (p == p_lane0) { r0 += c0; r1 += c1; r3 += c3; }
    }
    if (!live || !(f.total_simple >> 31)) return;
    if (type != CR_SEG_LINE && type != CR_SEG_INTEGRAL_QUADRATIC && type != CR_SEG_RATIONAL_QUADRATIC) { atomicOr(err, CR_DEVERR_BAD_TABLES); return; }
    {   // every read stays inside the path's slice of its type's array (the count pass has checked the tables; this keeps a disagreeing type stream harmless)
        const uint32_t at = type == CR_SEG_LINE ? f.tb0 + r0 : (type == CR_SEG_INTEGRAL_QUADRATIC ? f.tb1 + r1 : f.tb3 + r3);
        const uint32_t end = type == CR_SEG_LINE ? f.te0 : (type == CR_SEG_INTEGRAL_QUADRATIC ? f.te1 : f.te3);
        if (at >= end) { atomicOr(err, CR_DEVERR_BAD_TABLES); return; }
    }
    const uint32_t si = g - f.sb, total = f.total_simple & 0x7fffffffu;
    // the point this segment starts from: the path's start, or the end point of the previous segment
    float2 last;
    if (si == 0) last = make_float2(P.start[2 * (size_t)p], P.start[2 * (size_t)p + 1]);
    else {
        const uint32_t tp = P.segment_types[g - 1];
        if (tp == CR_SEG_LINE) { const float* d = P.seg[0] + 2 * (size_t)(f.tb0 + r0 - 1u); last = make_float2(d[0], d[1]); }
        else if (tp == CR_SEG_INTEGRAL_QUADRATIC) { const float* d = P.seg[1] + 4 * (size_t)(f.tb1 + r1 - 1u); last = make_float2(d[2], d[3]); }
        else { const float* d = P.seg[3] + 5 * (size_t)(f.tb3 + r3 - 1u); last = make_float2(d[3], d[4]); }
    }
    const uint32_t proto_at = f.proto + 1u + r0 + 2u * (r1 + r3);
    float2 end;
    if (type == CR_SEG_LINE) {
        const float* d = P.seg[0] + 2 * (size_t)(f.tb0 + r0);
        end = make_float2(d[0], d[1]);
        store_proto(out.proto, proto_at, end, e);
    } else if (type == CR_SEG_INTEGRAL_QUADRATIC) {   // src/fill.rs:284-299
        const float* dp = P.seg[1] + 4 * (size_t)(f.tb1 + r1);
        const float4 d = (reinterpret_cast<uintptr_t>(P.seg[1]) & 15u) == 0 ? *reinterpret_cast<const float4*>(dp) : make_float4(dp[0], dp[1], dp[2], dp[3]);   // a caller's device array may be 4-byte aligned only
        end = make_float2(d.z, d.w);
        float4* v = reinterpret_cast<float4*>(out.vtx[CAT_IQ]) + (size_t)f.iq + 3u * (size_t)r1;
        v[0] = make_float4(d.z, d.w, 1.0f, 1.0f);
        v[1] = make_float4(d.x, d.y, 0.5f, 0.0f);
        v[2] = make_float4(last.x, last.y, 0.0f, 0.0f);
        store_proto(out.proto, proto_at, make_float2(d.x, d.y), e);
        store_proto(out.proto, proto_at + 1u, end, e);
    } else {                                          // src/fill.rs:320-335
        const float* d = P.seg[3] + 5 * (size_t)(f.tb3 + r3);
        const float weight = 1.0f / d[0];
        end = make_float2(d[3], d[4]);
        float* v = reinterpret_cast<float*>(out.vtx[CAT_RQ]) + ((size_t)f.rq + 3u * (size_t)r3) * 5;
        v[0] = d[3]; v[1] = d[4]; v[2] = 1.0f; v[3] = 1.0f; v[4] = 1.0f;
        v[5] = d[1]; v[6] = d[2]; v[7] = 0.5f * weight; v[8] = 0.0f; v[9] = weight;
        v[10] = last.x; v[11] = last.y; v[12] = 0.0f; v[13] = 0.0f; v[14] = 1.0f;
        store_proto(out.proto, proto_at, make_float2(d[1], d[2]), e);
        store_proto(out.proto, proto_at + 1u, end, e);
    }
    reinterpret_cast<float2*>(out.vtx[CAT_SOLID])[f.solid + fan_to_strip_slot(si + 1u, total)] = end;
    out.idx[2][f.solid_idx + si + 1u] = ((f.rel + si + 1u) << 1) | ((si + 1u) & 1u);
    if (e) atomicOr(err, e);
}

// Per-shape slice boundaries: cat_begin[c][s] = offsets[c][shape_path_begin[s]], s in [0, n_shapes].
__global__ void shape_bounds_kernel(const uint32_t* __restrict__ offsets, uint32_t n_paths, const uint32_t* __restrict__ shape_path_begin, uint32_t n_shapes,
                                    uint32_t* __restrict__ cat_begin, uint32_t* __restrict__ max_proto, TessCapacity caps, uint32_t* __restrict__ err) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s == 0) {   // optimistic rebuild: the emit pass is already enqueued behind this kernel and writes into the previous build's arrays
        bool fits = true;
        for (int c = 0; c < CNT_COUNT; ++c) if (offsets[(size_t)c * (n_paths + 1) + n_paths] > caps.v[c]) fits = false;
        if (!fits) atomicOr(err, CR_DEVERR_CAPACITY);
    }
    if (s > n_shapes) return;
    const uint32_t p = shape_path_begin[s];
    for (int c = 0; c < CNT_COUNT; ++c) cat_begin[(size_t)c * (n_shapes + 1) + s] = offsets[(size_t)c * (n_paths + 1) + p];
    if (s < n_shapes) {
        const uint32_t q = shape_path_begin[s + 1];
        atomicMax(max_proto, offsets[(size_t)CNT_PROTO * (n_paths + 1) + q] - offsets[(size_t)CNT_PROTO * (n_paths + 1) + p]);
    }
}

// ------------------------------------------------------------------------------------ convex_hull.rs on device
// Lexicographic (x, then y) order of SafeFloat<f32, 2> (src/safe_float.rs:158-168).
__device__ __forceinline__ bool lex_less(float2 a, float2 b) { return a.x < b.x || (a.x == b.x && a.y < b.y); }
__device__ __forceinline__ float turn(float2 a, float2 b, float2 c) { return triple(from_vec(a.x, a.y), from_vec(b.x, b.y), from_vec(c.x, c.y)); }

// In-place bitonic sort of n points with virtual +inf padding (every comparator puts the minimum at the lower index, so
// comparators whose upper element is padding are no-ops and whole blocks of padding are skipped).
__device__ void block_sort_points(float2* pts, uint32_t n) {
    uint32_t log_np2 = 0;
    while ((1u << log_np2) < n) ++log_np2;
    for (uint32_t lk = 1; lk <= log_np2; ++lk) {          // k = 1 << lk: size of the bitonic blocks being merged
        for (uint32_t lj = lk; lj-- > 0;) {               // j = 1 << lj: comparator distance
            // comparators are numbered t = block * j + off; only blocks of size 2j that start below n can hold a real upper element
            const uint32_t blocks = (n + (2u << lj) - 1) >> (lj + 1);
            const uint32_t count = blocks << lj;
            const bool flip = lj + 1 == lk;
            for (uint32_t t = threadIdx.x; t < count; t += blockDim.x) {
                const uint32_t block = t >> lj, off = t & ((1u << lj) - 1u);
                const uint32_t lo = (block << (lj + 1)) + off;
                const uint32_t hi = flip ? (block << (lj + 1)) + (2u << lj) - 1u - off : lo + (1u << lj);
                if (hi < n) {
                    const float2 a = pts[lo], b = pts[hi];
                    if (lex_less(b, a)) { pts[lo] = b; pts[hi] = a; }
                }
            }
            __syncthreads();
        }
    }
}
// One monotone chain of convex_hull::andrew (src/convex_hull.rs:14-24 / :27-38): a strictly sequential stack machine
// (the pop test has a tolerance, so the result depends on the transient stack states and must be replayed exactly).
// The top two stack entries and the line through them stay in registers; `stack` may be shared or global memory.
// Returns the stack size, or HULL_OVERFLOW if it would exceed `capacity`.
#define HULL_OVERFLOW 0xFFFFFFFFu
template <int DIR>   // +1: points in ascending order (lower chain), -1: descending (upper chain)
__device__ __noinline__ uint32_t hull_chain(const float2* pts, uint32_t n, float2* stack, uint32_t capacity) {
    // This loop runs on ONE thread and is pure dependent-issue latency, so it is written for minimum instruction count:
    // the two top stack entries (a, b) and the line through them (l0 + l1 x + l2 y, == join(a, b) of device_common.cuh
    // for unit-weight points) live in registers, the point stream is walked by pointer and prefetched one ahead.
    const float2* src = DIR > 0 ? pts : pts + (n - 1);
    float2 a = *src;
    stack[0] = a;
    if (n < 2) return n;
    src += DIR;
    float2 b = *src;
    stack[1] = b;                       // no test while the stack holds fewer than two points
    uint32_t len = 2;
    float l0 = a.y * b.x - a.x * b.y, l1 = b.y - a.y, l2 = a.x - b.x;
    if (n < 3) return len;
    src += DIR;
    float2 p = *src;
    for (uint32_t k = 2; k < n; ++k) {
        const float2 next = src[k + 1 < n ? DIR : 0];   // independent of the stack: overlaps the tests below
        src += DIR;
        while (len > 1) {
            const float t = (l0 + p.x * l1) + p.y * l2;   // == (a v b) v p, src/convex_hull.rs:16-19
            if (!(t <= CR_ERROR_MARGIN)) break;
            --len;
            b = a;
            if (len > 1) {
                a = stack[len - 2];
                l0 = a.y * b.x - a.x * b.y; l1 = b.y - a.y; l2 = a.x - b.x;
            }
        }
        if (len >= capacity) return HULL_OVERFLOW;
        stack[len++] = p;
        a = b;
        b = p;
        l0 = a.y * b.x - a.x * b.y; l1 = b.y - a.y; l2 = a.x - b.x;
        p = next;
    }
    return len;
}
// ---- shared-memory sort: register-tiled bitonic network
// The network is the same data-oblivious one as block_sort_points (flip stage + half-cleaners, minimum to the lower index,
// virtual +inf padding), but each thread carries 16 elements through up to four consecutive stages in registers, so a
// merge of 2^lk elements costs 1 + ceil((lk - 1) / 4) shared-memory round trips instead of lk. Any correct sort gives the
// same array (equal keys are identical points). Element i lives at slot i + (i >> 4): the padding makes the three access
// strides used (1, 16, 256 elements) bank-conflict free for 8-byte accesses.
__device__ __forceinline__ void cswap(float2& lo, float2& hi) {
    const bool sw = hi.x < lo.x || (hi.x == lo.x && hi.y < lo.y);
    const float2 l = lo, h = hi;
    lo.x = sw ? h.x : l.x; lo.y = sw ? h.y : l.y;
    hi.x = sw ? l.x : h.x; hi.y = sw ? l.y : h.y;
}
template <int LK> __device__ __forceinline__ void reg_flip(float2 (&v)[16]) {   // first stage of the merge of 2^LK-blocks
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const int blk = t >> (LK - 1), off = t & ((1 << (LK - 1)) - 1);
        cswap(v[(blk << LK) + off], v[(blk << LK) + (1 << LK) - 1 - off]);
    }
}
template <int Q> __device__ __forceinline__ void reg_stage(float2 (&v)[16]) {   // half-cleaner at register distance 2^Q
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const int lo = ((t >> Q) << (Q + 1)) | (t & ((1 << Q) - 1));
        cswap(v[lo], v[lo + (1 << Q)]);
    }
}
#define SORT_SLOT(i) ((i) + ((i) >> 4))
// 16 elements i0 + m * s (m = 0..15) <-> registers; elements at or beyond n are +inf and never stored.
__device__ __forceinline__ void sort_load16(const float2* sm, uint32_t n, uint32_t i0, uint32_t s, uint32_t ps, float2 (&v)[16]) {
    const uint32_t p0 = SORT_SLOT(i0);
    if (i0 + 15u * s < n) {
#pragma unroll
        for (int m = 0; m < 16; ++m) v[m] = sm[p0 + m * ps];
    } else {
        const float inf = __int_as_float(0x7f800000);
#pragma unroll
        for (int m = 0; m < 16; ++m) v[m] = i0 + m * s < n ? sm[p0 + m * ps] : make_float2(inf, inf);
    }
}
__device__ __forceinline__ void sort_store16(float2* sm, uint32_t n, uint32_t i0, uint32_t s, uint32_t ps, const float2 (&v)[16]) {
    const uint32_t p0 = SORT_SLOT(i0);
    if (i0 + 15u * s < n) {
#pragma unroll
        for (int m = 0; m < 16; ++m) sm[p0 + m * ps] = v[m];
    } else {
#pragma unroll
        for (int m = 0; m < 16; ++m) if (i0 + m * s < n) sm[p0 + m * ps] = v[m];
    }
}
// Half-cleaner stages lj = top_lj .. 0 of the network on sm[0, n), four at a time: group g holds lj = 4g .. 4g+3.
__device__ void shared_half_cleaners(float2* sm, uint32_t n, int top_lj) {
    for (int g = top_lj >> 2; g >= 0; --g) {
        const int top_q = min(3, top_lj - 4 * g);
        const uint32_t s = 1u << (4 * g), ps = g == 0 ? 1u : s + (s >> 4);
        const uint32_t count = ((n + 16u * s - 1u) / (16u * s)) * s;
        for (uint32_t t = threadIdx.x; t < count; t += blockDim.x) {
            const uint32_t i0 = ((t >> (4 * g)) << (4 * g + 4)) + (t & (s - 1u));
            if (i0 >= n) continue;
            float2 v[16];
            sort_load16(sm, n, i0, s, ps, v);
            if (top_q >= 3) reg_stage<3>(v);
            if (top_q >= 2) reg_stage<2>(v);
            if (top_q >= 1) reg_stage<1>(v);
            reg_stage<0>(v);
            sort_store16(sm, n, i0, s, ps, v);
        }
        __syncthreads();
    }
}
// One stage of the network directly on global memory (shapes larger than the shared-memory capacity): comparator distance
// 2^lj, `flip` = first stage of a merge. Same comparators as block_sort_points.
__device__ void global_sort_stage(float2* pts, uint32_t n, uint32_t lj, bool flip) {
    const uint32_t blocks = (n + (2u << lj) - 1) >> (lj + 1);
    const uint32_t count = blocks << lj;
    for (uint32_t t = threadIdx.x; t < count; t += blockDim.x) {
        const uint32_t block = t >> lj, off = t & ((1u << lj) - 1u);
        const uint32_t lo = (block << (lj + 1)) + off;
        const uint32_t hi = flip ? (block << (lj + 1)) + (2u << lj) - 1u - off : lo + (1u << lj);
        if (hi < n) {
            const float2 a = pts[lo], b = pts[hi];
            if (lex_less(b, a)) { pts[lo] = b; pts[hi] = a; }
        }
    }
    __syncthreads();
}
__device__ void block_sort_points_shared(float2* sm, uint32_t n) {
    uint32_t log_n2 = 4;
    while ((1u << log_n2) < n) ++log_n2;
    // runs of 16: sorted completely in registers
    for (uint32_t t = threadIdx.x; t * 16u < n; t += blockDim.x) {
        float2 v[16];
        sort_load16(sm, n, t * 16u, 1u, 1u, v);
        reg_flip<1>(v);
        reg_flip<2>(v); reg_stage<0>(v);
        reg_flip<3>(v); reg_stage<1>(v); reg_stage<0>(v);
        reg_flip<4>(v); reg_stage<2>(v); reg_stage<1>(v); reg_stage<0>(v);
        sort_store16(sm, n, t * 16u, 1u, 1u, v);
    }
    __syncthreads();
    for (uint32_t lk = 5; lk <= log_n2; ++lk) {
        {   // flip stage: off <-> 2^lk - 1 - off inside each block; only blocks that start below n hold real upper elements
            const uint32_t half = 1u << (lk - 1);
            const uint32_t count = ((n + (2u * half) - 1u) >> lk) << (lk - 1);
            for (uint32_t t = threadIdx.x; t < count; t += blockDim.x) {
                const uint32_t blk = t >> (lk - 1), off = t & (half - 1u);
                const uint32_t lo = (blk << lk) + off, hi = (blk << lk) + 2u * half - 1u - off;
                if (hi < n) {
                    float2 a = sm[SORT_SLOT(lo)], b = sm[SORT_SLOT(hi)];
                    if (b.x < a.x || (b.x == a.x && b.y < a.y)) { sm[SORT_SLOT(lo)] = b; sm[SORT_SLOT(hi)] = a; }
                }
            }
            __syncthreads();
        }
        shared_half_cleaners(sm, n, (int)lk - 2);
    }
}

// Sort kernel: one CTA per shape; proto[begin, begin + n) is sorted in place (shared memory when it fits `cap` points).
__global__ void __launch_bounds__(512) hull_sort_kernel(float2* __restrict__ proto, const uint32_t* __restrict__ proto_begin, uint32_t cap, const uint32_t* __restrict__ err) {
    extern __shared__ float2 hull_smem[];
    if (*reinterpret_cast<volatile const uint32_t*>(err) & CR_DEVERR_FATAL_MASK) return;   // nothing was emitted
    const uint32_t s = blockIdx.x;
    const uint32_t begin = proto_begin[s], n = proto_begin[s + 1] - begin;
    if (n < 3) return;   // returned as-is, unsorted (src/convex_hull.rs:9-11)
    if (n <= cap) {
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) hull_smem[SORT_SLOT(i)] = proto[begin + i];
        __syncthreads();
        block_sort_points_shared(hull_smem, n);
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) proto[begin + i] = hull_smem[SORT_SLOT(i)];
    } else {
        // Larger than shared memory: the same network, with every stage whose comparator distance is below the chunk size C
        // done chunk by chunk in shared memory. Chunks are sorted first (merge levels up to C); each further level is its
        // flip stage and its half-cleaners of distance >= C on global memory (L2 resident), then one shared-memory pass
        // per chunk for the half-cleaners below C.
        float2* const g = proto + begin;
        const uint32_t log_c = 31u - (uint32_t)__clz((int)cap), C = 1u << log_c, n_chunks = (n + C - 1u) / C;
        uint32_t log_n2 = log_c;
        while ((1u << log_n2) < n) ++log_n2;
        for (uint32_t c = 0; c < n_chunks; ++c) {
            const uint32_t len = min(C, n - c * C);
            for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) hull_smem[SORT_SLOT(i)] = g[c * C + i];
            __syncthreads();
            block_sort_points_shared(hull_smem, len);
            for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) g[c * C + i] = hull_smem[SORT_SLOT(i)];
            __syncthreads();
        }
        for (uint32_t lk = log_c + 1u; lk <= log_n2; ++lk) {
            global_sort_stage(g, n, lk - 1u, true);
            for (uint32_t lj = lk - 2u; lj >= log_c; --lj) global_sort_stage(g, n, lj, false);
            for (uint32_t c = 0; c < n_chunks; ++c) {
                const uint32_t len = min(C, n - c * C);
                for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) hull_smem[SORT_SLOT(i)] = g[c * C + i];
                __syncthreads();
                shared_half_cleaners(hull_smem, len, (int)log_c - 1);
                for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) g[c * C + i] = hull_smem[SORT_SLOT(i)];
                __syncthreads();
            }
        }
    }
}

// ---- the two monotone chains
// One warp per chain, executed by lane 0 (the stack machine is strictly sequential, see hull_chain); the other lanes
// stream the sorted points into a double-buffered shared-memory window ahead of it. The three top stack entries (c, a, b;
// b on top) and the lines through (a, b) and (c, a) live in registers, and every point is tested against BOTH lines at
// once, together with the two lines a push would create. The common outcomes — "keep b" and "pop b once" — then cost one
// dependent test + select with no memory access on the critical path; only a point that pops two or more entries reloads
// from the shared stack. The predicate evaluations that decide anything are exactly the reference's (the speculative
// second test is only consulted after the first one failed). Chains run in their own kernel so that nothing else
// competes for the issue slots of these latency-bound warps.
__device__ __forceinline__ float2 lds_f2(uint32_t addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_f2(uint32_t addr, float2 v) { asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory"); }
struct HullLine { float l0, l1, l2; };
__device__ __forceinline__ HullLine hull_line(float2 a, float2 b) { return {a.y * b.x - a.x * b.y, b.y - a.y, a.x - b.x}; }   // == join(a, b), unit weights
__device__ __forceinline__ float hull_side(const HullLine& l, float2 p) { return (l.l0 + p.x * l.l1) + p.y * l.l2; }           // == (a v b) v p

#define HULL_STACK 512      // shared-memory chain stacks; deeper chains are redone with global stacks
#define CHAIN_WINDOW 128    // points per prefetch window
#ifndef CHAIN_LANES
#define CHAIN_LANES 2       // chains per warp, one per lane: lower and upper chain of CHAIN_LANES / 2 shapes. Measured on the text scene (kernel alone /
                            // pipelined step): 1 lane 0.492 / 1.109 ms, 2 lanes 0.532 / 0.987 ms, 4 lanes 0.604 / 1.016 ms, 8 lanes 0.741 / 1.041 ms
#endif
#define CHAIN_THREADS 128
#define CHAINS_PER_CTA (CHAIN_LANES * CHAIN_THREADS / 32)
#define CHAIN_SHAPES (CHAINS_PER_CTA / 2)
#define STACK_STRIDE (HULL_STACK + 1)      // + 1: the chains of a warp walk their stacks and windows at similar depths — keep them in different banks
#define WINDOW_STRIDE (CHAIN_WINDOW + 1)
// The chains are sequential stack machines (the pop test has a tolerance, so the result depends on every transient stack
// state): one LANE per chain. A lone lane per warp, as in round 1, left 31 of 32 lanes of every issued instruction idle — 1250
// chains took 46 % of the GPU's issue slots for half a millisecond, which is what the frames overlapping this kernel paid for.
// With CHAIN_LANES chains per warp the same instruction stream serves several chains (the common keep / replace-the-top
// outcomes are branch free; lanes that must pop deeper make the others wait), the remaining lanes stream the sorted points
// of all the warp's chains into double-buffered windows (cp.async).
__global__ void __launch_bounds__(CHAIN_THREADS) hull_chain_kernel(const float2* __restrict__ sorted, float2* __restrict__ scratch_a, float2* __restrict__ scratch_b,
                                                                   const uint32_t* __restrict__ proto_begin, uint32_t n_shapes,
                                                                   float2* __restrict__ hull_out, uint32_t* __restrict__ hull_count, const uint32_t* __restrict__ err) {
    if (*reinterpret_cast<volatile const uint32_t*>(err) & CR_DEVERR_FATAL_MASK) return;   // nothing was emitted
    extern __shared__ float2 sh_chain[];
    float2* const stacks = sh_chain;                                         // [CHAINS_PER_CTA][STACK_STRIDE]
    float2* const windows = sh_chain + CHAINS_PER_CTA * STACK_STRIDE;        // [CHAINS_PER_CTA][2][WINDOW_STRIDE]
    __shared__ uint32_t sh_len[CHAINS_PER_CTA];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    // the chain this lane describes (lanes >= CHAIN_LANES describe chain lane % CHAIN_LANES again: every lane helps prefetching)
    const uint32_t c_own = lane % CHAIN_LANES, chain_own = warp * CHAIN_LANES + c_own;
    const uint32_t s_own = blockIdx.x * CHAIN_SHAPES + (chain_own >> 1);
    uint32_t begin_own = 0, n_own = 0;
    if (s_own < n_shapes) { begin_own = proto_begin[s_own]; n_own = proto_begin[s_own + 1] - begin_own; }
    const bool descending = (c_own & 1u) != 0;   // even chain: lower (ascending points), odd chain: upper (descending)
    const bool runs = lane < CHAIN_LANES && n_own >= 3u;   // this lane executes a machine
    uint32_t n_max = 0;
#pragma unroll
    for (uint32_t c = 0; c < CHAIN_LANES; ++c) n_max = max(n_max, __shfl_sync(0xffffffffu, n_own >= 3u ? n_own : 0u, c));
    auto fetch = [&](uint32_t w) {       // stream window w of every chain of this warp into its buffer w & 1
        const uint32_t base = w * CHAIN_WINDOW;
#pragma unroll
        for (uint32_t c = 0; c < CHAIN_LANES; ++c) {
            const uint32_t n_c = __shfl_sync(0xffffffffu, n_own, c), begin_c = __shfl_sync(0xffffffffu, begin_own, c);
            if (n_c < 3u) continue;
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(windows + ((size_t)(warp * CHAIN_LANES + c) * 2u + (w & 1u)) * WINDOW_STRIDE);
#pragma unroll
            for (uint32_t i = 0; i < CHAIN_WINDOW / 32; ++i) {
                const uint32_t k = base + lane + 32u * i;
                if (k < n_c) {
                    const float2* src = sorted + begin_c + ((c & 1u) ? n_c - 1u - k : k);
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + (lane + 32u * i) * 8u), "l"(src) : "memory");
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (n_max != 0u) {
        fetch(0);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
    }
    const uint32_t stack0 = (uint32_t)__cvta_generic_to_shared(stacks + (size_t)chain_own * STACK_STRIDE);
    const uint32_t window0 = (uint32_t)__cvta_generic_to_shared(windows + (size_t)chain_own * 2u * WINDOW_STRIDE);
    const uint32_t floor2 = stack0 + 16;                 // top == floor2  <=>  two entries
    const uint32_t limit = stack0 + HULL_STACK * 8;      // pushing at `limit` would overflow
    uint32_t top = floor2;                               // address one past the top entry
    float2 a = make_float2(0.f, 0.f), b = a, c = a;
    HullLine lab = {0.f, 0.f, 0.f}, lca = lab;
    bool overflow = false;
    if (runs) {
        a = lds_f2(window0); b = lds_f2(window0 + 8);
        sts_f2(stack0, a); sts_f2(stack0 + 8, b);
        c = a; lab = hull_line(a, b); lca = lab;         // c, lca are meaningful only while the stack holds >= 3 entries
    }
    const uint32_t n_windows = (n_max + CHAIN_WINDOW - 1) / CHAIN_WINDOW;
    for (uint32_t w = 0; w < n_windows; ++w) {
        if (w + 1 < n_windows) fetch(w + 1);
        // The stack grows by at most one entry per point: one capacity check per window keeps it out of the loop.
        if (runs && !overflow && top + CHAIN_WINDOW * 8u > limit) overflow = true;
        if (runs && !overflow && w * CHAIN_WINDOW < n_own) {
            uint32_t src = window0 + (w & 1u) * (WINDOW_STRIDE * 8u);
            const uint32_t k0 = w == 0 ? 2u : 0u, k1 = min((uint32_t)CHAIN_WINDOW, n_own - w * CHAIN_WINDOW);
            src += k0 * 8u;
            float2 pn = lds_f2(src);
            for (uint32_t k = k0; k < k1; ++k) {
                const float2 p = pn;
                src += 8u;
                pn = lds_f2(k + 1 < k1 ? src : src - 8u);   // next point: independent of the stack, overlaps the tests below
                // pop while (a v b) v p <= margin (src/convex_hull.rs:16-19); a, b are the two top entries
                const float t1 = hull_side(lab, p), t2 = hull_side(lca, p);
                const bool keep = !(t1 <= CR_ERROR_MARGIN);                        // keep b:  .. c a b  ->  .. a b p
                if (keep || top == floor2 || !(t2 <= CR_ERROR_MARGIN)) {           // else pop b only (a is the last entry, or a stays):  .. c a b  ->  .. c a p
                    const uint32_t at = keep ? top : top - 8u;                     // the two common outcomes, branch free
                    sts_f2(at, p);
                    top = at + 8u;
                    c.x = keep ? a.x : c.x; c.y = keep ? a.y : c.y;
                    a.x = keep ? b.x : a.x; a.y = keep ? b.y : a.y;
                    lca.l0 = keep ? lab.l0 : lca.l0; lca.l1 = keep ? lab.l1 : lca.l1; lca.l2 = keep ? lab.l2 : lca.l2;
                    b = p;
                    lab = hull_line(a, b);                                         // one line from the selected entry (fewer instructions than both candidates + selects)
                } else {                                                           // b and a are popped: continue on the shared stack, which now ends with c
                    top -= 16;
                    b = c;
                    while (top >= floor2) {                                        // at least two entries are left
                        a = lds_f2(top - 16);
                        lab = hull_line(a, b);
                        if (!(hull_side(lab, p) <= CR_ERROR_MARGIN)) break;
                        top -= 8;
                        b = a;
                    }
                    sts_f2(top, p);
                    top += 8;
                    a = b; b = p;
                    lab = hull_line(a, b);
                    if (top > floor2) { c = lds_f2(top - 24); lca = hull_line(c, a); }
                }
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
    }
    if (lane < CHAIN_LANES) {
        uint32_t len = n_own;   // n < 3: returned as-is (src/convex_hull.rs:9-11)
        if (runs) {
            len = (top - stack0) >> 3;
            if (overflow) len = descending ? hull_chain<-1>(sorted + begin_own, n_own, scratch_b + begin_own, n_own) : hull_chain<1>(sorted + begin_own, n_own, scratch_a + begin_own, n_own);
        }
        sh_len[chain_own] = len | (overflow ? 0x80000000u : 0u);
    }
    __syncwarp();
    // triangle_fan_to_strip(andrew(..)) (src/renderer.rs:197, src/vertex.rs:28-35): the whole warp writes the hulls of its shapes
#pragma unroll 1
    for (uint32_t j = 0; j < CHAIN_LANES / 2; ++j) {
        const uint32_t chain = warp * CHAIN_LANES + 2u * j, s = blockIdx.x * CHAIN_SHAPES + (chain >> 1);
        if (s >= n_shapes) break;
        const uint32_t begin = __shfl_sync(0xffffffffu, begin_own, 2u * j), n = __shfl_sync(0xffffffffu, n_own, 2u * j);
        float2* out = hull_out + begin;
        if (n < 3u) {   // fan->strip of <= 2 points is the identity
            if (lane < n) out[lane] = sorted[begin + lane];
            if (lane == 0) hull_count[s] = n;
            continue;
        }
        const uint32_t len_a = sh_len[chain], len_b = sh_len[chain + 1];
        const float2* sa = (len_a >> 31) ? scratch_a + begin : stacks + (size_t)chain * STACK_STRIDE;
        const float2* sb = (len_b >> 31) ? scratch_b + begin : stacks + (size_t)(chain + 1) * STACK_STRIDE;
        const uint32_t la = (len_a & 0x7fffffffu) - 1, lb = (len_b & 0x7fffffffu) - 1, total = la + lb;   // hull.pop() after each chain
        for (uint32_t i = lane; i < total; i += 32u) {
            const uint32_t src = (i & 1u) == 0 ? (i >> 1) : total - 1 - (i >> 1);
            out[i] = src < la ? sa[src] : sb[src - la];
        }
        if (lane == 0) hull_count[s] = total;
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------- host launchers
int cr_tess_count(cudaStream_t stream, const DevicePaths& paths, uint32_t n_groups, uint32_t* counts, uint32_t* err_flag, bool has_cubics) {
    if (paths.n_paths == 0) return CR_OK;
    const uint32_t grid = (paths.n_paths + 127) / 128;
    if (paths.stroke_options) tess_count_kernel<0><<<grid, 128, 0, stream>>>(paths, n_groups, counts, err_flag);
    else if (has_cubics) tess_count_kernel<1><<<grid, 128, 0, stream>>>(paths, n_groups, counts, err_flag);
    else tess_count_kernel<2><<<grid, 128, 0, stream>>>(paths, n_groups, counts, err_flag);
    g_cr_kernel_launches += 1;
    CR_CUDA_TRY(cudaGetLastError());
    return CR_OK;
}
int cr_tess_shape_bounds(cudaStream_t stream, const uint32_t* offsets, uint32_t n_paths, const uint32_t* shape_path_begin, uint32_t n_shapes, uint32_t* cat_begin,
                         uint32_t* max_proto, const TessCapacity& caps, uint32_t* err_flag) {
    shape_bounds_kernel<<<(n_shapes + 1 + 127) / 128, 128, 0, stream>>>(offsets, n_paths, shape_path_begin, n_shapes, cat_begin, max_proto, caps, err_flag);
    g_cr_kernel_launches += 1;
    CR_CUDA_TRY(cudaGetLastError());
    return CR_OK;
}
int cr_tess_emit(cudaStream_t stream, const DevicePaths& paths, const uint32_t* offsets, const uint32_t* shape_path_begin, uint32_t n_shapes,
                 const TessOutput& out, uint32_t* err_flag, bool has_cubics) {
    if (paths.n_paths == 0) return CR_OK;
    // filled paths without cubics: one thread per segment (+ one per path); everything else: one thread per path
    const bool all_simple = !paths.stroke_options && !has_cubics;
    const uint32_t seg_threads = (paths.n_segments + 255u) / 256u * 256u + paths.n_paths;
    fill_segments_kernel<<<(seg_threads + 255u) / 256u, 256, 0, stream>>>(paths, offsets, shape_path_begin, n_shapes, out, err_flag, all_simple ? 1u : 0u);
    g_cr_kernel_launches += 1;
    if (!all_simple) {
        const uint32_t grid = (paths.n_paths + 127) / 128;
        if (paths.stroke_options) tess_emit_kernel<0><<<grid, 128, 0, stream>>>(paths, offsets, shape_path_begin, n_shapes, out, err_flag, 1u);
        else tess_emit_kernel<1><<<grid, 128, 0, stream>>>(paths, offsets, shape_path_begin, n_shapes, out, err_flag, 1u);
        g_cr_kernel_launches += 1;
    }
    CR_CUDA_TRY(cudaGetLastError());
    return CR_OK;
}
int cr_tess_hull(cudaStream_t stream, float2* proto, float2* scratch_a, float2* scratch_b, const uint32_t* proto_begin, uint32_t n_shapes,
                 float2* hull_out, uint32_t* hull_count, uint32_t max_points, const uint32_t* err_flag, cudaEvent_t after_sort) {
    if (n_shapes == 0) return CR_OK;
    // Sort: shared-memory capacity (in points) = the largest shape rounded up to 512 if that fits one SM's shared memory
    // (8.5 bytes per point with the bank padding), else the maximum — larger shapes sort in global memory.
    static bool attr_set = false;
    const uint32_t max_cap = 26624;
    if (!attr_set) {
        CR_CUDA_TRY(cudaFuncSetAttribute(hull_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SORT_SLOT(max_cap) * sizeof(float2))));
        attr_set = true;
    }
    const uint32_t cap = std::min<uint32_t>(max_cap, std::max<uint32_t>(512u, (max_points + 511u) / 512u * 512u));
    const uint32_t threads = std::min<uint32_t>(512u, std::max<uint32_t>(64u, cap / 16u));   // one 16-element register tile per thread
    hull_sort_kernel<<<n_shapes, threads, (size_t)SORT_SLOT(cap) * sizeof(float2), stream>>>(proto, proto_begin, cap, err_flag);
    if (after_sort) CR_CUDA_TRY(cudaEventRecord(after_sort, stream));
    const size_t chain_smem = ((size_t)CHAINS_PER_CTA * STACK_STRIDE + (size_t)CHAINS_PER_CTA * 2 * WINDOW_STRIDE) * sizeof(float2);
    CR_CUDA_TRY(cudaFuncSetAttribute(hull_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chain_smem));   // per device: set on every call
    hull_chain_kernel<<<(n_shapes + CHAIN_SHAPES - 1) / CHAIN_SHAPES, CHAIN_THREADS, chain_smem, stream>>>(proto, scratch_a, scratch_b, proto_begin, n_shapes, hull_out, hull_count, err_flag);
    g_cr_kernel_launches += 2;
    CR_CUDA_TRY(cudaGetLastError());
    return CR_OK;
}
