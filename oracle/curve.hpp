// ORACLE — TEST INFRASTRUCTURE ONLY (see ga.hpp). PARITY UNPINNED.
// Sequential restatement of /root/reference/src/curve.rs. Each function cites the lines it follows.
#pragma once
#include <algorithm>
#include <array>
#include <vector>
#include "ga.hpp"

namespace oracle {

static constexpr float ERROR_MARGIN = 0.0001f;            // src/error.rs:19
static constexpr float F32_EPSILON = 1.1920929e-7f;       // f32::EPSILON

using cr::Root;

// mat_vec_transform! (src/curve.rs:12-23): power_basis[0]*at0 + (power_basis[1]*at1 + (...)).
template <size_t N>
inline Point mat_vec_transform(const std::array<Point, N>& pb, const float* at, size_t count) {
    Point acc = pb[count - 1] * at[count - 1];
    for (size_t i = count - 1; i-- > 0;) acc = pb[i] * at[i] + acc;
    return acc;
}
template <size_t N>
inline Point mvt(const std::array<Point, N>& pb, std::initializer_list<float> at) {
    return mat_vec_transform(pb, at.begin(), at.size());
}

// src/curve.rs:26-32
inline std::array<Point, 3> rational_quadratic_control_points_to_power_basis(const std::array<Point, 3>& cp) {
    return {mvt(cp, {1.0f}), mvt(cp, {-2.0f, 2.0f}), mvt(cp, {1.0f, -2.0f, 1.0f})};
}
// src/curve.rs:35-42
inline std::array<Point, 4> rational_cubic_control_points_to_power_basis(const std::array<Point, 4>& cp) {
    return {mvt(cp, {1.0f}), mvt(cp, {-3.0f, 3.0f}), mvt(cp, {3.0f, -6.0f, 3.0f}), mvt(cp, {-1.0f, 3.0f, -3.0f, 1.0f})};
}
inline float powi2(float a) { return a * a; }
inline float powi3(float a) { return a * a * a; }
// src/curve.rs:58-83
inline std::array<Point, 4> reparametrize_rational_cubic(const std::array<Point, 4>& pb, float a, float b) {
    return {
        mvt(pb, {1.0f, a, powi2(a), powi3(a)}),
        mvt(pb, {0.0f, b - a, -2.0f * powi2(a) + 2.0f * a * b, 3.0f * powi2(a) * b - 3.0f * powi3(a)}),
        mvt(pb, {0.0f, 0.0f, powi2(a - b), -6.0f * powi2(a) * b + 3.0f * a * powi2(b) + 3.0f * powi3(a)}),
        mvt(pb, {0.0f, 0.0f, 0.0f, 3.0f * powi2(a) * b - 3.0f * a * powi2(b) - powi3(a) + powi3(b)}),
    };
}
// src/curve.rs:86-95
inline Point rational_quadratic_point(const std::array<Point, 3>& pb, float t) { return mvt(pb, {1.0f, t, powi2(t)}); }
inline Plane rational_quadratic_first_order_derivative(const std::array<Point, 3>& pb, float t) {
    const Point p = mvt(pb, {1.0f, t, powi2(t)});
    const Point d1 = mvt(pb, {0.0f, 1.0f, 2.0f * t});
    return regressive(p, d1);
}
// src/curve.rs:105-114
inline Point rational_cubic_point(const std::array<Point, 4>& pb, float t) { return mvt(pb, {1.0f, t, powi2(t), powi3(t)}); }
inline Plane rational_cubic_first_order_derivative(const std::array<Point, 4>& pb, float t) {
    const Point p = mvt(pb, {1.0f, t, powi2(t), powi3(t)});
    const Point d1 = mvt(pb, {0.0f, 1.0f, 2.0f * t, 3.0f * powi2(t)});
    return regressive(p, d1);
}

// src/curve.rs:133-144
inline std::array<float, 4> inflection_point_polynomial_coefficients(const std::array<Point, 4>& pb, bool integral) {
    std::array<float, 4> ippc = {0.0f, 0.0f, 0.0f, 0.0f};
    for (int j = integral ? 1 : 0; j < 4; ++j) {
        Point sel[3];
        int n = 0;
        for (int i = 0; i < 4; ++i)
            if (i != j) sel[n++] = pb[i];
        ippc[j] = triple(sel[0], sel[1], sel[2]) * (float)(j % 2 * 2 - 1);
    }
    // ppga3d::Rotor::signum: divide by the Euclidean norm of the four components.
    const float mag = cr::sqrt_f(ippc[0] * ippc[0] + ippc[1] * ippc[1] + ippc[2] * ippc[2] + ippc[3] * ippc[3]);
    const float inv = 1.0f / mag;
    for (float& v : ippc) v = v * inv;
    return ippc;
}

struct DiscriminantAndRoots {
    float discriminant;
    std::array<Root, 3> roots;
};

// src/curve.rs:151-190
inline DiscriminantAndRoots integral_inflection_points(const std::array<float, 4>& ippc, bool loop_self_intersection) {
    const float discriminant = 3.0f * powi2(ippc[2]) - 4.0f * ippc[1] * ippc[3];
    if (cr::fabs_f(ippc[1]) <= ERROR_MARGIN) {
        if (cr::fabs_f(ippc[2]) <= ERROR_MARGIN)
            return {-1.0f, {cr::make_root(-1.0f, 0.0f, 1.0f), cr::no_root(), cr::no_root()}};
        return {1.0f, {cr::make_root(ippc[3], 0.0f, 3.0f * ippc[2]), cr::no_root(), cr::no_root()}};
    }
    const float d = cr::sqrt_f(discriminant * (discriminant < 0.0f ? (loop_self_intersection ? -1.0f : 0.0f) : 1.0f / 3.0f));
    return {discriminant,
            {cr::make_root(ippc[2] + d, 0.0f, 2.0f * ippc[1]), cr::make_root(ippc[2] - d, 0.0f, 2.0f * ippc[1]), cr::no_root()}};
}

// src/curve.rs:197-226
inline DiscriminantAndRoots rational_inflection_points(const std::array<float, 4>& ippc, bool loop_self_intersection) {
    if (cr::fabs_f(ippc[0]) <= ERROR_MARGIN) return integral_inflection_points(ippc, loop_self_intersection);
    const cr::Roots cubic = cr::solve_cubic(ippc[3] * -1.0f, ippc[2] * 3.0f, ippc[1] * -3.0f, ippc[0], ERROR_MARGIN);
    std::array<Root, 3> roots = {cubic.r[0], cubic.r[1], cubic.r[2]};
    if (!loop_self_intersection) return {cubic.discriminant, roots};
    const cr::Roots hessian = cr::solve_quadratic(ippc[1] * ippc[3] - ippc[2] * ippc[2], ippc[1] * ippc[2] - ippc[0] * ippc[3],
                                                  ippc[0] * ippc[2] - ippc[1] * ippc[1], ERROR_MARGIN);
    if (hessian.discriminant > 0.0f) {
        roots[2] = roots[cubic.real_root];
        if (hessian.count == 2) {
            roots[0] = hessian.r[0];
            roots[1] = hessian.r[1];
        } else if (hessian.count == 1) {
            roots[0] = hessian.r[0];
            roots[1] = cr::no_root();
        }
    }
    return {-hessian.discriminant, roots};
}

// interpolate_normal! (src/curve.rs:228-252). `solve(normal)` returns the candidate roots for one normal.
template <typename Solve>
inline std::vector<float> interpolate_normal(Plane start_tangent, Plane end_tangent, float angle_step, Solve solve) {
    const cr::Complex polar_start = cr::cplx(start_tangent[1], start_tangent[2]);
    const cr::Complex polar_end = cr::cplx(end_tangent[1], end_tangent[2]);
    const cr::Complex polar_range = cr::cdiv(polar_end, polar_start);
    const uint32_t steps = cr::f32_to_usize_sat(cr::fabs_f(cr::carg(polar_range) / angle_step) + 0.5f);
    std::vector<float> out;
    if (steps < 2) return out;
    const cr::Complex polar_step = cr::cpowf(polar_range, 1.0f / (float)steps);
    out.reserve(steps - 1);
    for (uint32_t i = 1; i < steps; ++i) {
        const cr::Complex interpolated = cr::cmul(polar_start, cr::cpowi(polar_step, i));
        const Plane normal = plane(0.0f, interpolated.re, interpolated.im);
        const cr::Roots sol = solve(normal);
        float parameter = 0.0f;
        for (int k = 0; k < sol.count; ++k) {
            if (sol.r[k].denominator == 0.0f) continue;
            const float candidate = sol.r[k].numerator.re / sol.r[k].denominator;
            if (candidate >= 0.0f && candidate <= 1.0f) {
                parameter = candidate;
                break;
            }
        }
        out.push_back(parameter);
    }
    return out;
}

// cubic_uniform_tangent_angle! (src/curve.rs:254-303). `make_solver(trimmed_power_basis)` returns the per-interval
// solve(normal) closure.
template <typename MakeSolver>
inline std::vector<float> cubic_uniform_tangent_angle(const std::array<Point, 4>& pb, float angle_step,
                                                      const DiscriminantAndRoots& dr, MakeSolver make_solver) {
    std::vector<float> split_parameters;
    for (const Root& root : dr.roots) {
        if (root.denominator == 0.0f) continue;
        const float parameter = root.numerator.re / root.denominator;
        if (parameter >= 0.0f && parameter <= 1.0f) split_parameters.push_back(parameter);
    }
    std::stable_sort(split_parameters.begin(), split_parameters.end());
    for (size_t i = 1; i < split_parameters.size();) {
        if (split_parameters[i] - split_parameters[i - 1] < ERROR_MARGIN) split_parameters.erase(split_parameters.begin() + i);
        else ++i;
    }
    float previous_split = 0.0f;
    std::vector<std::pair<float, float>> intervals;
    for (float split_parameter : split_parameters) {
        if (cr::fabs_f(dr.discriminant) < ERROR_MARGIN) {
            intervals.push_back({previous_split, split_parameter - F32_EPSILON});
            previous_split = split_parameter + F32_EPSILON;
        } else {
            intervals.push_back({previous_split, split_parameter});
            previous_split = split_parameter;
        }
    }
    intervals.push_back({previous_split, 1.0f});
    std::vector<float> parameters;
    for (const auto& ab : intervals) {
        const float a = ab.first, b = ab.second;
        const std::array<Point, 4> trimmed = reparametrize_rational_cubic(pb, a, b);
        const Plane start_tangent = signum(rational_cubic_first_order_derivative(pb, a));
        const Plane end_tangent = signum(rational_cubic_first_order_derivative(pb, b));
        auto solve = make_solver(trimmed);
        std::vector<float> interval_parameters = interpolate_normal(start_tangent, end_tangent, angle_step, solve);
        for (float& t : interval_parameters) t = a + (b - a) * t;
        std::stable_sort(interval_parameters.begin(), interval_parameters.end());
        parameters.insert(parameters.end(), interval_parameters.begin(), interval_parameters.end());
        parameters.push_back(b);
    }
    return parameters;
}

// src/curve.rs:306-322
inline std::vector<float> integral_quadratic_uniform_tangent_angle(const std::array<Point, 3>& pb, Plane start_tangent,
                                                                   Plane end_tangent, float angle_step) {
    const Plane planes[2] = {dual(pb[1]), dual(pb[2]) * 2.0f};
    std::vector<float> parameters = interpolate_normal(start_tangent, end_tangent, angle_step, [&](Plane normal) {
        return cr::solve_linear(inner(normal, planes[0]), inner(normal, planes[1]), ERROR_MARGIN);
    });
    parameters.push_back(1.0f);
    return parameters;
}
// src/curve.rs:325-352
inline std::vector<float> integral_cubic_uniform_tangent_angle(const std::array<Point, 4>& pb, float angle_step) {
    const std::array<float, 4> ippc = inflection_point_polynomial_coefficients(pb, true);
    const DiscriminantAndRoots dr = integral_inflection_points(ippc, false);
    return cubic_uniform_tangent_angle(pb, angle_step, dr, [](const std::array<Point, 4>& trimmed) {
        const std::array<Plane, 3> planes = {dual(trimmed[1]), dual(trimmed[2]) * 2.0f, dual(trimmed[3]) * 3.0f};
        return [planes](Plane normal) {
            return cr::solve_quadratic(inner(normal, planes[0]), inner(normal, planes[1]), inner(normal, planes[2]), ERROR_MARGIN);
        };
    });
}
// src/curve.rs:355-380
inline std::vector<float> rational_quadratic_uniform_tangent_angle(const std::array<Point, 3>& pb, Plane start_tangent,
                                                                   Plane end_tangent, float angle_step) {
    const Plane planes[3] = {regressive(pb[1], pb[0]), regressive(pb[2], pb[0]) * 2.0f, regressive(pb[2], pb[1])};
    std::vector<float> parameters = interpolate_normal(start_tangent, end_tangent, angle_step, [&](Plane normal_in) {
        const Plane normal = rotate_90_degree_clockwise(normal_in);
        return cr::solve_quadratic(inner(normal, planes[0]), inner(normal, planes[1]), inner(normal, planes[2]), ERROR_MARGIN);
    });
    parameters.push_back(1.0f);
    return parameters;
}
// src/curve.rs:383-418
inline std::vector<float> rational_cubic_uniform_tangent_angle(const std::array<Point, 4>& pb, float angle_step) {
    const std::array<float, 4> ippc = inflection_point_polynomial_coefficients(pb, false);
    const DiscriminantAndRoots dr = rational_inflection_points(ippc, false);
    return cubic_uniform_tangent_angle(pb, angle_step, dr, [](const std::array<Point, 4>& t) {
        const std::array<Plane, 5> planes = {
            regressive(t[1], t[0]),
            regressive(t[2], t[0]) * 2.0f,
            regressive(t[2], t[1]) + regressive(t[3], t[0]) * 3.0f,
            regressive(t[3], t[1]) * 2.0f,
            regressive(t[3], t[2]),
        };
        return [planes](Plane normal_in) {
            const Plane normal = rotate_90_degree_clockwise(normal_in);
            return cr::solve_quartic(inner(normal, planes[0]), inner(normal, planes[1]), inner(normal, planes[2]),
                                     inner(normal, planes[3]), inner(normal, planes[4]), ERROR_MARGIN);
        };
    });
}

}  // namespace oracle
