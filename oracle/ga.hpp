// ORACLE — TEST INFRASTRUCTURE ONLY. Nothing under contrast_renderer_b200/ may include, link or call this.
// PARITY UNPINNED: the reference has no tests / golden vectors and cannot be built here (no Rust toolchain);
// the closed forms below are derived from the basis declaration of the un-vendored crate
// `geometric_algebra 0.3.0` (Cargo.lock:520) — see SURVEY.md Appendix A — and pinned only by the hand-derived
// anchors in tests/test_oracle_anchors.py.
//
// ppga2d: generators e0^2 = 0, e1^2 = e2^2 = 1;  Point = (e12, e01, -e02),  Plane = (e0, e2, e1).
// A Point is (w, w*x, w*y) (src/utils.rs:111-118); a Plane is the line g0 + g1*x + g2*y = 0 whose (g1, g2) is
// the path direction rotated 90 degrees clockwise (y up) when it comes from Point v Point.
#pragma once
#include "../contrast_renderer_b200/csrc/arith/cr_arith.h"

namespace oracle {

struct Point { float g[3]; float operator[](int i) const { return g[i]; } float& operator[](int i) { return g[i]; } };
struct Plane { float g[3]; float operator[](int i) const { return g[i]; } float& operator[](int i) { return g[i]; } };

inline Point point(float a, float b, float c) { return Point{{a, b, c}}; }
inline Plane plane(float a, float b, float c) { return Plane{{a, b, c}}; }
inline Point operator*(Point p, float s) { return point(p[0] * s, p[1] * s, p[2] * s); }
inline Point operator+(Point a, Point b) { return point(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
inline Plane operator*(Plane p, float s) { return plane(p[0] * s, p[1] * s, p[2] * s); }
inline Plane operator+(Plane a, Plane b) { return plane(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
inline Plane operator-(Plane a, Plane b) { return plane(a[0] - b[0], a[1] - b[1], a[2] - b[2]); }
inline Plane operator-(Plane a) { return plane(-a[0], -a[1], -a[2]); }

// Dual is the component-wise identity between Point and Plane.
inline Plane dual(Point p) { return plane(p[0], p[1], p[2]); }

// Point v Point -> Plane (the line joining them).
inline Plane regressive(Point p, Point q) {
    return plane(p[2] * q[1] - p[1] * q[2], p[0] * q[2] - p[2] * q[0], p[1] * q[0] - p[0] * q[1]);
}
// Point v Plane -> scalar (signed, weighted incidence).
inline float regressive(Point p, Plane a) { return p[0] * a[0] + p[1] * a[1] + p[2] * a[2]; }
inline float regressive(Plane a, Point p) { return regressive(p, a); }
// Plane ^ Plane -> Point (the intersection, homogeneous).
inline Point outer(Plane a, Plane b) {
    return point(a[2] * b[1] - a[1] * b[2], a[0] * b[2] - a[2] * b[0], a[1] * b[0] - a[0] * b[1]);
}
// Plane . Plane -> scalar (dot product of the normals).
inline float inner(Plane a, Plane b) { return a[1] * b[1] + a[2] * b[2]; }
inline float squared_magnitude(Plane a) { return a[1] * a[1] + a[2] * a[2]; }
inline float magnitude(Plane a) { return cr::sqrt_f(squared_magnitude(a)); }
inline Plane signum(Plane a) { return a * (1.0f / magnitude(a)); }
// `tangent.inner_product(vertex).geometric_product(vertex).into(): Plane` (src/stroke.rs:71-75,86):
// the line through `p` parallel to `a`, scaled by -p0^2 (the scale cancels in line_line_intersection).
inline Plane parallel_through(Plane a, Point p) {
    const float t = a[1] * p[1] + a[2] * p[2];
    return plane(p[0] * t, -(p[0] * (a[1] * p[0])), -(p[0] * (a[2] * p[0])));
}

// src/utils.rs:67-118
inline Point line_line_intersection(Plane a, Plane b) {
    const Point p = outer(a, b);
    return p * (1.0f / p[0]);
}
inline Plane rotate_90_degree_clockwise(Plane v) { return plane(0.0f, v[2], -v[1]); }
inline void point_to_vec(Point p, float out[2]) { out[0] = p[1] / p[0]; out[1] = p[2] / p[0]; }
inline Point vec_to_point(const float v[2]) { return point(1.0f, v[0], v[1]); }
inline Point weighted_vec_to_point(float w, const float v[2]) { return point(w, v[0] * w, v[1] * w); }

// (A v B) v C for three Points = -det[A;B;C] (negative for counter-clockwise triangles, y up).
inline float triple(Point a, Point b, Point c) { return regressive(regressive(a, b), c); }

}  // namespace oracle
