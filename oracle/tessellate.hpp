// ORACLE — TEST INFRASTRUCTURE ONLY (see ga.hpp). PARITY UNPINNED.
// Sequential restatement of /root/reference/src/{stroke.rs, fill.rs, convex_hull.rs, vertex.rs} and of
// Shape::from_paths / concat_buffers! / convert_dynamic_stroke_options in src/renderer.rs.
#pragma once
#include <cstring>
#include <stdexcept>
#include "../include/contrast_b200.h"
#include "curve.hpp"

namespace oracle {

// ------------------------------------------------------------------------------------------------ src/vertex.rs
#pragma pack(push, 1)
struct Vertex0 { float p[2]; };                               // 8 B
struct Vertex2f { float p[2]; float w[2]; };                  // 16 B
struct Vertex2f1i { float p[2]; float t[2]; uint32_t i; };    // 20 B
struct Vertex3f { float p[2]; float w[3]; };                  // 20 B
struct Vertex3f1i { float p[2]; float t[3]; uint32_t i; };    // 24 B
struct Vertex4f { float p[2]; float w[4]; };                  // 24 B
#pragma pack(pop)

// src/vertex.rs:28-35
template <typename T>
inline std::vector<T> triangle_fan_to_strip(const std::vector<T>& vertices) {
    std::vector<T> result;
    result.reserve(vertices.size());
    for (size_t i = 0; i < vertices.size(); ++i) result.push_back(vertices[(i & 1) == 0 ? (i >> 1) : vertices.size() - 1 - (i >> 1)]);
    return result;
}

struct NonFinite : std::runtime_error { NonFinite() : std::runtime_error("non finite") {} };
struct CubicTriangulation : std::runtime_error { CubicTriangulation() : std::runtime_error("cubic triangulation") {} };

// SafeFloat<f32, 2>::from (src/safe_float.rs:111-121)
inline Vertex0 safe2(const float v[2]) {
    if (!cr::is_finite(v[0]) || !cr::is_finite(v[1])) throw NonFinite();
    return Vertex0{{cr::canon_zero(v[0]), cr::canon_zero(v[1])}};
}
inline Vertex0 safe2(Point p) {
    float v[2];
    point_to_vec(p, v);
    return safe2(v);
}

// -------------------------------------------------------------------------------------------------- path model
struct Path {
    bool stroked = false;
    cr_stroke_options so{};
    float start[2] = {0, 0};
    std::vector<std::array<float, 2>> line_segments;
    std::vector<std::array<float, 4>> integral_quadratic_curve_segments;
    std::vector<std::array<float, 6>> integral_cubic_curve_segments;
    std::vector<std::array<float, 5>> rational_quadratic_curve_segments;   // weight, p1, p2
    std::vector<std::array<float, 10>> rational_cubic_curve_segments;      // weights[4], p1, p2, p3
    std::vector<uint8_t> segment_types;
};

// ------------------------------------------------------------------------------------------------ src/stroke.rs
struct StrokeBuilder {
    std::vector<uint16_t> line_indices, joint_indices;
    std::vector<uint32_t> line_indices_wide, joint_indices_wide;  // same entries without the `as u16` truncation (0xFFFFFFFF = restart)
    std::vector<Vertex2f1i> line_vertices;
    std::vector<Vertex3f1i> joint_vertices;
    std::vector<Vertex2f1i> path_line_vertices;
};

// src/stroke.rs:18-22
inline Point offset_control_point(Point control_point, Plane tangent, float offset) {
    Point direction = point(0.0f, tangent[1], tangent[2]);
    return control_point + direction * offset;
}
// src/stroke.rs:24-51
inline void emit_stroke_vertices(StrokeBuilder& b, const cr_stroke_options& so, uint32_t path_index, float length_accumulator,
                                 Point pt, Plane tangent) {
    const float offset_along_path = length_accumulator / so.width;
    const float sides[2] = {-0.5f, 0.5f};
    const float offs[2] = {(so.offset - 0.5f) * so.width, (so.offset + 0.5f) * so.width};
    for (int k = 0; k < 2; ++k) {
        Vertex2f1i v;
        point_to_vec(offset_control_point(pt, tangent, offs[k]), v.p);
        v.t[0] = sides[k];
        v.t[1] = offset_along_path;
        v.i = path_index;
        b.path_line_vertices.push_back(v);
    }
}
// src/stroke.rs:123-132
inline void cut_stroke_polygon(StrokeBuilder& b, std::vector<Vertex0>& proto_hull) {
    if (b.path_line_vertices.empty()) return;
    for (const Vertex2f1i& v : b.path_line_vertices) proto_hull.push_back(safe2(v.p));
    const size_t start_index = b.line_vertices.size();
    b.line_vertices.insert(b.line_vertices.end(), b.path_line_vertices.begin(), b.path_line_vertices.end());
    b.path_line_vertices.clear();
    for (size_t i = start_index; i < b.line_vertices.size(); ++i) { b.line_indices.push_back((uint16_t)i); b.line_indices_wide.push_back((uint32_t)i); }
    b.line_indices.push_back((uint16_t)0xFFFF);
    b.line_indices_wide.push_back(0xFFFFFFFFu);
}
// src/stroke.rs:53-121
inline void emit_stroke_join(StrokeBuilder& b, std::vector<Vertex0>& proto_hull, const cr_stroke_options& so, float& length_accumulator,
                             Point control_point, Plane previous_tangent, Plane next_tangent) {
    const float tangets_dot_product = inner(previous_tangent, next_tangent);
    if (cr::fabs_f(tangets_dot_product - 1.0f) <= ERROR_MARGIN) return;
    const float side_sign = cr::rust_signum(outer(previous_tangent, next_tangent)[0]);
    const float miter_clip = so.width * so.miter_clip;
    const float side_offset = (so.offset - side_sign * 0.5f) * so.width;
    const Point previous_edge_vertex = offset_control_point(control_point, previous_tangent, side_offset);
    const Point next_edge_vertex = offset_control_point(control_point, next_tangent, side_offset);
    const Plane previous_edge_tangent = parallel_through(previous_tangent, previous_edge_vertex);
    const Plane next_edge_tangent = parallel_through(next_tangent, next_edge_vertex);
    const Point intersection = line_line_intersection(previous_edge_tangent, next_edge_tangent);
    Point vertices[5] = {control_point, previous_edge_vertex, next_edge_vertex, intersection, intersection};
    const bool anti_parallel = cr::fabs_f(tangets_dot_product + 1.0f) <= ERROR_MARGIN;
    if (anti_parallel || magnitude(regressive(control_point, intersection)) > miter_clip) {
        const Plane mid_tangent = anti_parallel ? -rotate_90_degree_clockwise(previous_tangent) : signum(previous_tangent + next_tangent);
        const Point clipping_vertex = offset_control_point(control_point, mid_tangent, -side_sign * miter_clip);
        const Plane clipping_plane = parallel_through(mid_tangent, clipping_vertex);
        vertices[3] = line_line_intersection(previous_edge_tangent, clipping_plane);
        vertices[4] = line_line_intersection(clipping_plane, next_edge_tangent);
        proto_hull.push_back(safe2(vertices[3]));
        proto_hull.push_back(safe2(vertices[4]));
    } else {
        proto_hull.push_back(safe2(vertices[3]));
    }
    const Plane scaled_tangent = previous_tangent * (1.0f / -so.width);
    const size_t start_index = b.joint_vertices.size();
    const float offset_along_path = length_accumulator / so.width;
    for (const Point& vertex : vertices) {
        Vertex3f1i v;
        point_to_vec(vertex, v.p);
        v.t[0] = side_sign * regressive(vertex, scaled_tangent);
        v.t[1] = inner(regressive(vertex, control_point), scaled_tangent);
        v.t[2] = offset_along_path;
        v.i = so.dynamic_stroke_options_group;
        b.joint_vertices.push_back(v);
    }
    for (size_t i = start_index; i < b.joint_vertices.size(); ++i) { b.joint_indices.push_back((uint16_t)i); b.joint_indices_wide.push_back((uint32_t)i); }
    b.joint_indices.push_back((uint16_t)0xFFFF);
    b.joint_indices_wide.push_back(0xFFFFFFFFu);
    length_accumulator += cr::acos_f(tangets_dot_product) / (3.14159274101257324219f * 2.0f) * so.width;
    cut_stroke_polygon(b, proto_hull);
    emit_stroke_vertices(b, so, so.dynamic_stroke_options_group, length_accumulator, control_point, next_tangent);
}
// src/stroke.rs:179-202
inline void get_quadratic_tangents(const Point cp[3], Plane& s, Plane& e) {
    s = signum(regressive(cp[0], cp[1]));
    e = signum(regressive(cp[1], cp[2]));
    if (cr::is_nan(s[0]) || cr::is_nan(e[0])) {
        s = signum(regressive(cp[0], cp[2]));
        e = s;
    }
}
inline void get_cubic_tangents(const Point cp[4], Plane& s, Plane& e) {
    s = signum(regressive(cp[0], cp[1]));
    if (cr::is_nan(s[0])) s = signum(regressive(cp[0], cp[2]));
    e = signum(regressive(cp[2], cp[3]));
    if (cr::is_nan(e[0])) e = signum(regressive(cp[1], cp[3]));
    if (cr::is_nan(s[0]) || cr::is_nan(e[0])) e = signum(regressive(cp[0], cp[3]));
}

// emit_curve_stroke! (src/stroke.rs:134-168)
template <typename PB, typename PointFn, typename TangentFn>
inline void emit_curve_stroke(StrokeBuilder& b, const cr_stroke_options& so, float& length_accumulator, Point previous_control_point,
                              const PB& power_basis, const std::vector<float>& parameters, PointFn point_fn, TangentFn tangent_fn) {
    Point previous_point = previous_control_point;
    for (float t : parameters) {
        Plane tangent = tangent_fn(power_basis, t);
        if (squared_magnitude(tangent) == 0.0f) {
            if (t < 0.5f) t += F32_EPSILON;
            else t -= F32_EPSILON;
            tangent = tangent_fn(power_basis, t);
        }
        tangent = signum(tangent);
        Point pt = point_fn(power_basis, t);
        pt = pt * (1.0f / pt[0]);
        length_accumulator += magnitude(regressive(previous_point, pt));
        emit_stroke_vertices(b, so, so.dynamic_stroke_options_group, length_accumulator, pt, tangent);
        previous_point = pt;
    }
}
inline std::vector<float> uniformly_spaced_parameters(uint32_t steps) {
    std::vector<float> p;
    for (uint32_t i = 1; i < steps + 1; ++i) p.push_back((float)i / (float)steps);
    return p;
}

// StrokeBuilder::add_path (src/stroke.rs:205-465)
inline void stroke_add_path(StrokeBuilder& b, std::vector<Vertex0>& proto_hull, const Path& path) {
    const cr_stroke_options& so = path.so;
    const bool closed = (so.flags & CR_STROKE_FLAG_CLOSED) != 0;
    const bool uta = (so.flags & CR_STROKE_FLAG_UNIFORM_TANGENT_ANGLE) != 0;
    const uint32_t group = so.dynamic_stroke_options_group;
    Point previous_control_point = vec_to_point(path.start);
    Plane first_tangent = plane(0, 0, 0), previous_tangent = plane(0, 0, 0);
    size_t li = 0, iqi = 0, ici = 0, rqi = 0, rci = 0;
    float length_accumulator = 0.0f;
    bool is_first_segment = true;
    for (uint8_t segment_type : path.segment_types) {
        Point next_control_point;
        Plane segment_start_tangent, segment_end_tangent;
        switch (segment_type) {
            case CR_SEG_LINE: {
                const auto& s = path.line_segments.at(li++);
                next_control_point = vec_to_point(&s[0]);
                segment_start_tangent = signum(regressive(previous_control_point, next_control_point));
                segment_end_tangent = segment_start_tangent;
            } break;
            case CR_SEG_INTEGRAL_QUADRATIC: {
                const auto& s = path.integral_quadratic_curve_segments.at(iqi);  // peek
                next_control_point = vec_to_point(&s[2]);
                const Point cp[3] = {previous_control_point, vec_to_point(&s[0]), next_control_point};
                get_quadratic_tangents(cp, segment_start_tangent, segment_end_tangent);
            } break;
            case CR_SEG_INTEGRAL_CUBIC: {
                const auto& s = path.integral_cubic_curve_segments.at(ici);
                next_control_point = vec_to_point(&s[4]);
                const Point cp[4] = {previous_control_point, vec_to_point(&s[0]), vec_to_point(&s[2]), next_control_point};
                get_cubic_tangents(cp, segment_start_tangent, segment_end_tangent);
            } break;
            case CR_SEG_RATIONAL_QUADRATIC: {
                const auto& s = path.rational_quadratic_curve_segments.at(rqi);
                next_control_point = vec_to_point(&s[3]);
                const Point cp[3] = {previous_control_point, vec_to_point(&s[1]), next_control_point};
                get_quadratic_tangents(cp, segment_start_tangent, segment_end_tangent);
            } break;
            default: {
                const auto& s = path.rational_cubic_curve_segments.at(rci);
                next_control_point = vec_to_point(&s[8]);
                const Point cp[4] = {previous_control_point, vec_to_point(&s[4]), vec_to_point(&s[6]), next_control_point};
                get_cubic_tangents(cp, segment_start_tangent, segment_end_tangent);
            } break;
        }
        if (cr::is_nan(segment_start_tangent[0]) || cr::is_nan(segment_end_tangent[0])) continue;  // quirk C.3: no iterator advance
        if (is_first_segment) {
            is_first_segment = false;
            first_tangent = segment_start_tangent;
            if (!closed) {
                const Plane normal = rotate_90_degree_clockwise(segment_start_tangent);
                emit_stroke_vertices(b, so, group, length_accumulator - 0.5f * so.width,
                                     offset_control_point(previous_control_point, normal, 0.5f * cr::fabs_f(so.width)), segment_start_tangent);
            }
            if (closed || segment_type != CR_SEG_LINE)
                emit_stroke_vertices(b, so, group, length_accumulator, previous_control_point, segment_start_tangent);
        } else {
            emit_stroke_join(b, proto_hull, so, length_accumulator, previous_control_point, previous_tangent, segment_start_tangent);
        }
        switch (segment_type) {
            case CR_SEG_LINE: {
                length_accumulator += magnitude(regressive(previous_control_point, next_control_point));
                emit_stroke_vertices(b, so, group, length_accumulator, next_control_point, segment_end_tangent);
            } break;
            case CR_SEG_INTEGRAL_QUADRATIC: {
                const auto& s = path.integral_quadratic_curve_segments.at(iqi++);
                const auto pb = rational_quadratic_control_points_to_power_basis({previous_control_point, vec_to_point(&s[0]), vec_to_point(&s[2])});
                const auto params = uta ? integral_quadratic_uniform_tangent_angle(pb, segment_start_tangent, segment_end_tangent, so.approximation.angle_step)
                                        : uniformly_spaced_parameters(so.approximation.steps);
                emit_curve_stroke(b, so, length_accumulator, previous_control_point, pb, params, rational_quadratic_point,
                                  rational_quadratic_first_order_derivative);
            } break;
            case CR_SEG_INTEGRAL_CUBIC: {
                const auto& s = path.integral_cubic_curve_segments.at(ici++);
                const auto pb = rational_cubic_control_points_to_power_basis(
                    {previous_control_point, vec_to_point(&s[0]), vec_to_point(&s[2]), vec_to_point(&s[4])});
                const auto params = uta ? integral_cubic_uniform_tangent_angle(pb, so.approximation.angle_step)
                                        : uniformly_spaced_parameters(so.approximation.steps);
                emit_curve_stroke(b, so, length_accumulator, previous_control_point, pb, params, rational_cubic_point,
                                  rational_cubic_first_order_derivative);
            } break;
            case CR_SEG_RATIONAL_QUADRATIC: {
                const auto& s = path.rational_quadratic_curve_segments.at(rqi++);
                const auto pb = rational_quadratic_control_points_to_power_basis(
                    {previous_control_point, weighted_vec_to_point(s[0], &s[1]), vec_to_point(&s[3])});
                const auto params = uta ? rational_quadratic_uniform_tangent_angle(pb, segment_start_tangent, segment_end_tangent, so.approximation.angle_step)
                                        : uniformly_spaced_parameters(so.approximation.steps);
                emit_curve_stroke(b, so, length_accumulator, previous_control_point, pb, params, rational_quadratic_point,
                                  rational_quadratic_first_order_derivative);
            } break;
            default: {
                const auto& s = path.rational_cubic_curve_segments.at(rci++);
                float prev_vec[2];
                point_to_vec(previous_control_point, prev_vec);
                const auto pb = rational_cubic_control_points_to_power_basis({weighted_vec_to_point(s[0], prev_vec), weighted_vec_to_point(s[1], &s[4]),
                                                                              weighted_vec_to_point(s[2], &s[6]), weighted_vec_to_point(s[3], &s[8])});
                const auto params = uta ? rational_cubic_uniform_tangent_angle(pb, so.approximation.angle_step)
                                        : uniformly_spaced_parameters(so.approximation.steps);
                emit_curve_stroke(b, so, length_accumulator, previous_control_point, pb, params, rational_cubic_point,
                                  rational_cubic_first_order_derivative);
            } break;
        }
        previous_control_point = next_control_point;
        previous_tangent = segment_end_tangent;
    }
    if (closed) {
        const Point start = vec_to_point(path.start);
        const Plane line_segment = regressive(previous_control_point, start);
        const float length = magnitude(line_segment);
        if (length > 0.0f) {
            const Plane segment_tangent = line_segment * (1.0f / length);  // geometric_quotient by a scalar
            emit_stroke_join(b, proto_hull, so, length_accumulator, previous_control_point, previous_tangent, segment_tangent);
            length_accumulator += length;
            emit_stroke_vertices(b, so, group, length_accumulator, start, segment_tangent);
            emit_stroke_join(b, proto_hull, so, length_accumulator, start, segment_tangent, first_tangent);
        } else {
            emit_stroke_join(b, proto_hull, so, length_accumulator, start, previous_tangent, first_tangent);
        }
    } else {
        cut_stroke_polygon(b, proto_hull);
        emit_stroke_vertices(b, so, group | 0x10000u, length_accumulator, previous_control_point, previous_tangent);
        const Plane normal = rotate_90_degree_clockwise(previous_tangent);
        emit_stroke_vertices(b, so, group | 0x10000u, length_accumulator + 0.5f * so.width,
                             offset_control_point(previous_control_point, normal, -0.5f * cr::fabs_f(so.width)), previous_tangent);
    }
    cut_stroke_polygon(b, proto_hull);
}

// -------------------------------------------------------------------------------------------------- src/fill.rs
struct FillBuilder {
    std::vector<uint16_t> solid_indices;
    std::vector<uint32_t> solid_indices_wide;
    std::vector<Vertex0> solid_vertices;
    std::vector<Vertex2f> integral_quadratic_vertices;
    std::vector<Vertex3f> integral_cubic_vertices;
    std::vector<Vertex3f> rational_quadratic_vertices;
    std::vector<Vertex4f> rational_cubic_vertices;
};
using Weights = std::array<std::array<float, 4>, 4>;  // [control point][k, l, m, n]

// src/fill.rs:14-32
inline bool find_double_point_issue(float discriminant, const std::array<Root, 3>& roots, float& param) {
    if (discriminant < 0.0f) {
        float result = -1.0f;
        int inside = 0;
        for (const Root& root : roots) {
            if (root.denominator != 0.0f) {
                const float parameter = root.numerator.re / root.denominator;
                if (0.0f < parameter && parameter < 1.0f) {
                    result = parameter;
                    inside += 1;
                }
            }
        }
        if (inside == 1) {
            param = result;
            return true;
        }
    }
    return false;
}
// src/fill.rs:34-49
inline void weight_derivatives(Weights& weights, int column, const Root& r0, const Root& r1, const Root& r2) {
    const float n0 = r0.numerator.re, n1 = r1.numerator.re, n2 = r2.numerator.re;
    const float d0 = r0.denominator, d1 = r1.denominator, d2 = r2.denominator;
    const float pb[4] = {
        n0 * n1 * n2,
        -d0 * n1 * n2 - n0 * d1 * n2 - n0 * n1 * d2,
        n0 * d1 * d2 + d0 * n1 * d2 + d0 * d1 * n2,
        -d0 * d1 * d2,
    };
    weights[0][column] = pb[0];
    weights[1][column] = pb[0] + pb[1] * 1.0f / 3.0f;
    weights[2][column] = pb[0] + pb[1] * 2.0f / 3.0f + pb[2] * 1.0f / 3.0f;
    weights[3][column] = pb[0] + pb[1] + pb[2] + pb[3];
}
// src/fill.rs:51-68
inline Weights weights_of(float discriminant, const std::array<Root, 3>& roots) {
    Weights w{};
    if (discriminant == 0.0f) {
        weight_derivatives(w, 0, roots[0], roots[0], roots[2]);
        weight_derivatives(w, 1, roots[0], roots[0], roots[0]);
        weight_derivatives(w, 2, roots[0], roots[0], roots[0]);
    } else if (discriminant < 0.0f) {
        weight_derivatives(w, 0, roots[0], roots[1], roots[2]);
        weight_derivatives(w, 1, roots[0], roots[0], roots[1]);
        weight_derivatives(w, 2, roots[1], roots[1], roots[0]);
    } else {
        weight_derivatives(w, 0, roots[0], roots[1], roots[2]);
        weight_derivatives(w, 1, roots[0], roots[0], roots[0]);
        weight_derivatives(w, 2, roots[1], roots[1], roots[1]);
    }
    weight_derivatives(w, 3, roots[2], roots[2], roots[2]);
    return w;
}
// ppga3d: plane through three homogeneous points P v Q v R (4 cofactors of the 3x4 matrix). The overall sign is
// irrelevant: weight_planes rescales by 1 / -plane[3] (src/fill.rs:81).
inline float det3(const float a[3], const float b[3], const float c[3]) {
    return a[0] * (b[1] * c[2] - b[2] * c[1]) - a[1] * (b[0] * c[2] - b[2] * c[0]) + a[2] * (b[0] * c[1] - b[1] * c[0]);
}
inline std::array<float, 4> plane_through(const float p[4], const float q[4], const float r[4]) {
    std::array<float, 4> out;
    for (int skip = 0; skip < 4; ++skip) {
        float a[3], b[3], c[3];
        int n = 0;
        for (int k = 0; k < 4; ++k)
            if (k != skip) {
                a[n] = p[k];
                b[n] = q[k];
                c[n] = r[k];
                ++n;
            }
        const float d = det3(a, b, c);
        out[skip] = (skip & 1) ? -d : d;
    }
    return out;
}
// src/fill.rs:70-85
inline std::array<Plane, 4> weight_planes(const std::array<Point, 4>& control_points, const Weights& weights) {
    std::array<Plane, 4> planes;
    for (int i = 0; i < 4; ++i) {
        float pts[4][4];
        for (int j = 0; j < 4; ++j) {
            pts[j][0] = control_points[j][0];
            pts[j][1] = control_points[j][1];
            pts[j][2] = control_points[j][2];
            pts[j][3] = weights[j][i];
        }
        std::array<float, 4> p3 = plane_through(pts[0], pts[1], pts[2]);
        if (p3[1] * p3[1] + p3[2] * p3[2] + p3[3] * p3[3] < ERROR_MARGIN) p3 = plane_through(pts[0], pts[1], pts[3]);
        const float s = 1.0f / -p3[3];
        planes[i] = plane(p3[0] * s, p3[1] * s, p3[2] * s);
    }
    return planes;
}
// src/fill.rs:87-96
inline float implicit_curve_value(const std::array<float, 4>& w) { return w[0] * w[0] * w[0] - w[1] * w[2] * w[3]; }
inline Plane implicit_curve_gradient(const std::array<Plane, 4>& planes, const std::array<float, 4>& w) {
    return planes[0] * (3.0f * w[0] * w[0]) - planes[1] * (w[2] * w[3]) - planes[2] * (w[1] * w[3]) - planes[3] * (w[1] * w[2]);
}

struct CubicSink {
    std::vector<Vertex0>* fill_solid_vertices;
    std::vector<Vertex3f>* integral;   // exactly one of integral / rational is set
    std::vector<Vertex4f>* rational;
};
// emit_cubic_curve_triangle! (src/fill.rs:116-132)
inline void emit_cubic_curve_triangle(CubicSink& sink, const float areas[4], const std::array<Point, 4>& cp, const Weights& w, int triangle_index) {
    int idx[3], n = 0;
    for (int i = 0; i < 4; ++i)
        if (i != triangle_index) idx[n++] = i;
    const float area = areas[triangle_index];
    if (!(cr::fabs_f(area) > ERROR_MARGIN)) return;
    if (area < 0.0f) std::swap(idx[0], idx[2]);
    for (int k = 0; k < 3; ++k) {
        float v[2];
        point_to_vec(cp[idx[k]], v);
        const auto& ww = w[idx[k]];
        if (sink.integral) sink.integral->push_back(Vertex3f{{v[0], v[1]}, {ww[0], ww[1], ww[2]}});
        else sink.rational->push_back(Vertex4f{{v[0], v[1]}, {ww[0], ww[1], ww[2], ww[3]}});
    }
}
// triangulate_cubic_curve_quadrilateral! (src/fill.rs:134-204). `w` is modified in place like the macro does.
inline void triangulate_cubic_curve_quadrilateral(CubicSink& sink, const std::array<Point, 4>& cp, Weights& w) {
    for (int j = 0; j < 4; ++j) {
        const float s = 1.0f / cp[j][0];
        for (int k = 0; k < 4; ++k) w[j][k] *= s;
    }
    float areas[4];
    for (int i = 0; i < 4; ++i) {
        Point sel[3];
        int n = 0;
        for (int j = 0; j < 4; ++j)
            if (i != j) sel[n++] = cp[j];
        areas[i] = triple(sel[0], sel[1], sel[2]);
    }
    const float sum = cr::fabs_f(areas[0]) + cr::fabs_f(areas[1]) + cr::fabs_f(areas[2]) + cr::fabs_f(areas[3]);
    int enclosing = -1;
    for (int i = 0; i < 4; ++i) {
        const float equilibrium = 0.5f * sum;
        if (cr::fabs_f(equilibrium - cr::fabs_f(areas[i])) <= ERROR_MARGIN) enclosing = (enclosing == -1) ? i : -1;
    }
    if (enclosing >= 0) {
        emit_cubic_curve_triangle(sink, areas, cp, w, enclosing);
    } else {
        int opposite = 0;
        for (int j = 1; j < 4; ++j) {
            const float side_of_a = areas[j];
            const float side_of_d = areas[0] * (j == 2 ? -1.0f : 1.0f);
            if (side_of_a * side_of_d < 0.0f) {
                if (opposite != 0) throw CubicTriangulation();  // assert_eq!(opposite_triangle, 0)
                opposite = j;
            }
        }
        if (opposite == 0) throw CubicTriangulation();  // assert_ne!(opposite_triangle, 0)
        emit_cubic_curve_triangle(sink, areas, cp, w, 0);
        emit_cubic_curve_triangle(sink, areas, cp, w, opposite);
    }
    int additional = 0;
    for (int i = 1; i < 3; ++i) {
        if (enclosing != i && implicit_curve_value(w[i]) < 0.0f) {
            float v[2];
            point_to_vec(cp[i], v);
            sink.fill_solid_vertices->push_back(Vertex0{{v[0], v[1]}});
            additional += 1;
        }
    }
    if (additional == 2 && areas[0] * areas[1] < 0.0f) {
        auto& s = *sink.fill_solid_vertices;
        std::swap(s[s.size() - 2], s[s.size() - 1]);
    }
}
// split_curve_at! (src/fill.rs:206-216) for Points and for weight rows.
template <typename T, typename Lerp>
inline void split_curve_at(const std::array<T, 4>& c, float param, Lerp lerp, std::array<T, 4>& a, std::array<T, 4>& b) {
    const T p10 = lerp(c[0], c[1], param), p11 = lerp(c[1], c[2], param), p12 = lerp(c[2], c[3], param);
    const T p20 = lerp(p10, p11, param), p21 = lerp(p11, p12, param);
    const T p30 = lerp(p20, p21, param);
    a = {c[0], p10, p20, p30};
    b = {p30, p21, p12, c[3]};
}
// emit_cubic_curve! (src/fill.rs:218-250)
inline void emit_cubic_curve(std::vector<Vertex0>& proto_hull, CubicSink& sink, const std::array<Point, 4>& control_points,
                             const std::array<Point, 4>& c, float discriminant, const std::array<Root, 3>& roots) {
    Weights weights = weights_of(discriminant, roots);
    std::array<Plane, 4> planes = weight_planes(control_points, weights);
    const Plane gradient = implicit_curve_gradient(planes, weights[0]);
    // normalize_implicit_curve_side (src/fill.rs:98-114)
    const Plane tangent = rational_cubic_first_order_derivative(c, 0.0f);
    if (inner(tangent, gradient) > 0.0f) {
        for (auto& row : weights) {
            row[0] *= -1.0f;
            row[1] *= -1.0f;
        }
    }
    float param;
    if (find_double_point_issue(discriminant, roots, param)) {
        auto lerp_point = [](Point a, Point b, float t) { return a * (1.0f - t) + b * t; };
        auto lerp_row = [](std::array<float, 4> a, std::array<float, 4> b, float t) {
            std::array<float, 4> o;
            for (int k = 0; k < 4; ++k) o[k] = a[k] * (1.0f - t) + b[k] * t;
            return o;
        };
        std::array<Point, 4> cpa, cpb;
        Weights wa, wb;
        split_curve_at(control_points, param, lerp_point, cpa, cpb);
        split_curve_at(weights, param, lerp_row, wa, wb);
        triangulate_cubic_curve_quadrilateral(sink, cpa, wa);
        float v[2];
        point_to_vec(cpb[0], v);
        sink.fill_solid_vertices->push_back(Vertex0{{v[0], v[1]}});
        for (auto& row : wb) {
            row[0] *= -1.0f;
            row[1] *= -1.0f;
        }
        triangulate_cubic_curve_quadrilateral(sink, cpb, wb);
    } else {
        triangulate_cubic_curve_quadrilateral(sink, control_points, weights);
    }
    proto_hull.push_back(safe2(control_points[1]));
    proto_hull.push_back(safe2(control_points[2]));
    proto_hull.push_back(safe2(control_points[3]));
    float v[2];
    point_to_vec(control_points[3], v);
    sink.fill_solid_vertices->push_back(Vertex0{{v[0], v[1]}});
}

// FillBuilder::add_path (src/fill.rs:263-367)
inline void fill_add_path(FillBuilder& b, std::vector<Vertex0>& proto_hull, const Path& path) {
    std::vector<Vertex0> path_solid_vertices;
    path_solid_vertices.push_back(Vertex0{{path.start[0], path.start[1]}});
    proto_hull.push_back(safe2(path.start));
    size_t li = 0, iqi = 0, ici = 0, rqi = 0, rci = 0;
    for (uint8_t segment_type : path.segment_types) {
        switch (segment_type) {
            case CR_SEG_LINE: {
                const auto& s = path.line_segments.at(li++);
                proto_hull.push_back(safe2(&s[0]));
                path_solid_vertices.push_back(Vertex0{{s[0], s[1]}});
            } break;
            case CR_SEG_INTEGRAL_QUADRATIC: {
                const auto& s = path.integral_quadratic_curve_segments.at(iqi++);
                const Vertex0 last = path_solid_vertices.back();
                b.integral_quadratic_vertices.push_back(Vertex2f{{s[2], s[3]}, {1.0f, 1.0f}});
                b.integral_quadratic_vertices.push_back(Vertex2f{{s[0], s[1]}, {0.5f, 0.0f}});
                b.integral_quadratic_vertices.push_back(Vertex2f{{last.p[0], last.p[1]}, {0.0f, 0.0f}});
                proto_hull.push_back(safe2(&s[0]));
                proto_hull.push_back(safe2(&s[2]));
                path_solid_vertices.push_back(Vertex0{{s[2], s[3]}});
            } break;
            case CR_SEG_INTEGRAL_CUBIC: {
                const auto& s = path.integral_cubic_curve_segments.at(ici++);
                const Vertex0 last = path_solid_vertices.back();
                const std::array<Point, 4> control_points = {vec_to_point(last.p), vec_to_point(&s[0]), vec_to_point(&s[2]), vec_to_point(&s[4])};
                const auto power_basis = rational_cubic_control_points_to_power_basis(control_points);
                const auto ippc = inflection_point_polynomial_coefficients(power_basis, true);
                const DiscriminantAndRoots dr = integral_inflection_points(ippc, true);
                CubicSink sink{&path_solid_vertices, &b.integral_cubic_vertices, nullptr};
                emit_cubic_curve(proto_hull, sink, control_points, power_basis, dr.discriminant, dr.roots);
            } break;
            case CR_SEG_RATIONAL_QUADRATIC: {
                const auto& s = path.rational_quadratic_curve_segments.at(rqi++);
                const float weight = 1.0f / s[0];
                const Vertex0 last = path_solid_vertices.back();
                b.rational_quadratic_vertices.push_back(Vertex3f{{s[3], s[4]}, {1.0f, 1.0f, 1.0f}});
                b.rational_quadratic_vertices.push_back(Vertex3f{{s[1], s[2]}, {0.5f * weight, 0.0f, weight}});
                b.rational_quadratic_vertices.push_back(Vertex3f{{last.p[0], last.p[1]}, {0.0f, 0.0f, 1.0f}});
                proto_hull.push_back(safe2(&s[1]));
                proto_hull.push_back(safe2(&s[3]));
                path_solid_vertices.push_back(Vertex0{{s[3], s[4]}});
            } break;
            default: {
                const auto& s = path.rational_cubic_curve_segments.at(rci++);
                const Vertex0 last = path_solid_vertices.back();
                const std::array<Point, 4> control_points = {weighted_vec_to_point(s[0], last.p), weighted_vec_to_point(s[1], &s[4]),
                                                             weighted_vec_to_point(s[2], &s[6]), weighted_vec_to_point(s[3], &s[8])};
                const auto power_basis = rational_cubic_control_points_to_power_basis(control_points);
                const auto ippc = inflection_point_polynomial_coefficients(power_basis, false);
                const DiscriminantAndRoots dr = rational_inflection_points(ippc, true);
                CubicSink sink{&path_solid_vertices, nullptr, &b.rational_cubic_vertices};
                emit_cubic_curve(proto_hull, sink, control_points, power_basis, dr.discriminant, dr.roots);
            } break;
        }
    }
    const size_t start_index = b.solid_vertices.size();
    const std::vector<Vertex0> strip = triangle_fan_to_strip(path_solid_vertices);
    b.solid_vertices.insert(b.solid_vertices.end(), strip.begin(), strip.end());
    for (size_t i = start_index; i < b.solid_vertices.size(); ++i) { b.solid_indices.push_back((uint16_t)i); b.solid_indices_wide.push_back((uint32_t)i); }
    b.solid_indices.push_back((uint16_t)0xFFFF);
    b.solid_indices_wide.push_back(0xFFFFFFFFu);
}

// ------------------------------------------------------------------------------------------- src/convex_hull.rs
inline std::vector<Vertex0> andrew(const std::vector<Vertex0>& input) {
    std::vector<Vertex0> pts = input;
    if (pts.size() < 3) return pts;
    std::stable_sort(pts.begin(), pts.end(), [](const Vertex0& a, const Vertex0& b) {
        if (a.p[0] != b.p[0]) return a.p[0] < b.p[0];
        return a.p[1] < b.p[1];
    });
    std::vector<Vertex0> hull;
    hull.reserve(2 * pts.size());
    auto turn = [](const Vertex0& a, const Vertex0& b, const Vertex0& c) { return triple(vec_to_point(a.p), vec_to_point(b.p), vec_to_point(c.p)); };
    for (const Vertex0& p : pts) {
        while (hull.size() > 1 && turn(hull[hull.size() - 2], hull[hull.size() - 1], p) <= ERROR_MARGIN) hull.pop_back();
        hull.push_back(p);
    }
    hull.pop_back();
    const size_t t = hull.size() + 1;
    for (size_t k = pts.size(); k-- > 0;) {
        const Vertex0& p = pts[k];
        while (hull.size() > t && turn(hull[hull.size() - 2], hull[hull.size() - 1], p) <= ERROR_MARGIN) hull.pop_back();
        hull.push_back(p);
    }
    hull.pop_back();
    return hull;
}

// -------------------------------------------------------------------------- src/renderer.rs:18-60, 177-215
#pragma pack(push, 1)
struct DynamicStrokeDescriptor {
    float gap_start[CR_MAX_DASH_INTERVALS];
    float gap_end[CR_MAX_DASH_INTERVALS];
    uint32_t caps;
    uint32_t count_dashed_join;
    float phase;
    uint32_t _padding;
};
#pragma pack(pop)
static_assert(sizeof(DynamicStrokeDescriptor) == 48, "descriptor must be 48 bytes");

inline int convert_dynamic_stroke_options(const cr_dynamic_stroke_options& o, DynamicStrokeDescriptor& out) {
    std::memset(&out, 0, sizeof(out));
    if (o.dashed) {
        if (o.pattern_len > CR_MAX_DASH_INTERVALS) return CR_ERR_TOO_MANY_DASH_INTERVALS;
        if (o.pattern_len == 0) return CR_ERR_INVALID_ARGUMENT;
        out.count_dashed_join = ((o.pattern_len - 1) << 3) | 4u | o.join;
        out.phase = o.phase;
        for (uint32_t i = 0; i < o.pattern_len; ++i) {
            out.gap_start[i] = o.pattern[i].gap_start;
            out.gap_end[i] = o.pattern[i].gap_end;
            out.caps |= o.pattern[i].dash_start << (((i + o.pattern_len - 1) % o.pattern_len) * 8);
            out.caps |= o.pattern[i].dash_end << (i * 8 + 4);
        }
    } else {
        out.caps = o.start | (o.end << 4);
        out.count_dashed_join = o.join;
    }
    return CR_OK;
}

struct Shape {
    uint64_t vertex_offsets[8];
    uint64_t index_offsets[3];
    std::vector<uint8_t> vertex_buffer, index_buffer, stroke_buffer;
    std::vector<uint32_t> wide_indices;   // [line | joint | solid], used by the oracle rasteriser (fences off quirk C.1)
    uint64_t index_counts[3] = {0, 0, 0};
    uint64_t proto_hull_points = 0;
    uint64_t dynamic_stroke_options_count = 0;
};

template <typename T>
inline void append_bytes(std::vector<uint8_t>& dst, const std::vector<T>& src) {
    const uint8_t* p = reinterpret_cast<const uint8_t*>(src.data());
    dst.insert(dst.end(), p, p + src.size() * sizeof(T));
}

// Shape::from_paths (src/renderer.rs:177-249); returns a cr_status.
inline int shape_from_paths(const cr_dynamic_stroke_options* groups, size_t n_groups, const std::vector<Path>& paths, Shape& out) {
    std::vector<Vertex0> proto_hull;
    StrokeBuilder sb;
    FillBuilder fb;
    try {
        for (const Path& path : paths) {
            if (path.stroked) {
                if (path.so.dynamic_stroke_options_group >= n_groups) return CR_ERR_DYNAMIC_STROKE_OPTIONS_INDEX_OUT_OF_BOUNDS;
                stroke_add_path(sb, proto_hull, path);
            } else {
                fill_add_path(fb, proto_hull, path);
            }
        }
    } catch (const NonFinite&) {
        return CR_ERR_NON_FINITE;
    } catch (const CubicTriangulation&) {
        return CR_ERR_CUBIC_TRIANGULATION;
    }
    const std::vector<Vertex0> convex_hull = triangle_fan_to_strip(andrew(proto_hull));
    out.proto_hull_points = proto_hull.size();
    out.vertex_buffer.clear();
    out.index_buffer.clear();
    out.stroke_buffer.clear();
    int k = 0;
    append_bytes(out.vertex_buffer, sb.line_vertices); out.vertex_offsets[k++] = out.vertex_buffer.size();
    append_bytes(out.vertex_buffer, sb.joint_vertices); out.vertex_offsets[k++] = out.vertex_buffer.size();
    append_bytes(out.vertex_buffer, fb.solid_vertices); out.vertex_offsets[k++] = out.vertex_buffer.size();
    append_bytes(out.vertex_buffer, fb.integral_quadratic_vertices); out.vertex_offsets[k++] = out.vertex_buffer.size();
    append_bytes(out.vertex_buffer, fb.integral_cubic_vertices); out.vertex_offsets[k++] = out.vertex_buffer.size();
    append_bytes(out.vertex_buffer, fb.rational_quadratic_vertices); out.vertex_offsets[k++] = out.vertex_buffer.size();
    append_bytes(out.vertex_buffer, fb.rational_cubic_vertices); out.vertex_offsets[k++] = out.vertex_buffer.size();
    append_bytes(out.vertex_buffer, convex_hull); out.vertex_offsets[k++] = out.vertex_buffer.size();
    k = 0;
    append_bytes(out.index_buffer, sb.line_indices); out.index_offsets[k++] = out.index_buffer.size();
    append_bytes(out.index_buffer, sb.joint_indices); out.index_offsets[k++] = out.index_buffer.size();
    append_bytes(out.index_buffer, fb.solid_indices); out.index_offsets[k++] = out.index_buffer.size();
    out.wide_indices = sb.line_indices_wide;
    out.wide_indices.insert(out.wide_indices.end(), sb.joint_indices_wide.begin(), sb.joint_indices_wide.end());
    out.wide_indices.insert(out.wide_indices.end(), fb.solid_indices_wide.begin(), fb.solid_indices_wide.end());
    out.index_counts[0] = sb.line_indices_wide.size();
    out.index_counts[1] = sb.joint_indices_wide.size();
    out.index_counts[2] = fb.solid_indices_wide.size();
    std::vector<DynamicStrokeDescriptor> descs(n_groups);
    for (size_t i = 0; i < n_groups; ++i) {
        const int st = convert_dynamic_stroke_options(groups[i], descs[i]);
        if (st != CR_OK) return st;
    }
    append_bytes(out.stroke_buffer, descs);
    out.dynamic_stroke_options_count = n_groups;
    return CR_OK;
}

// Build oracle::Path objects from the C-ABI structure-of-arrays (host memory).
inline std::vector<Path> paths_from_soa(const cr_path_soa& s, uint32_t begin, uint32_t end) {
    std::vector<Path> paths;
    const uint32_t stride = s.n_paths + 1;
    for (uint32_t p = begin; p < end; ++p) {
        Path path;
        path.so = s.stroke_options ? s.stroke_options[p] : cr_stroke_options{};
        path.stroked = s.stroke_options && (path.so.flags & CR_STROKE_FLAG_STROKED);
        path.start[0] = s.start[2 * p];
        path.start[1] = s.start[2 * p + 1];
        path.segment_types.assign(s.segment_types + s.segment_begin[p], s.segment_types + s.segment_begin[p + 1]);
        auto copy = [&](auto& dst, const float* src, int t, int width) {
            for (uint32_t i = s.type_begin[t * stride + p]; i < s.type_begin[t * stride + p + 1]; ++i) {
                typename std::remove_reference<decltype(dst)>::type::value_type e;
                for (int k = 0; k < width; ++k) e[k] = src[(size_t)i * width + k];
                dst.push_back(e);
            }
        };
        copy(path.line_segments, s.line_segments, 0, 2);
        copy(path.integral_quadratic_curve_segments, s.integral_quadratic, 1, 4);
        copy(path.integral_cubic_curve_segments, s.integral_cubic, 2, 6);
        copy(path.rational_quadratic_curve_segments, s.rational_quadratic, 3, 5);
        copy(path.rational_cubic_curve_segments, s.rational_cubic, 4, 10);
        paths.push_back(std::move(path));
    }
    return paths;
}

}  // namespace oracle
