// ORACLE — TEST INFRASTRUCTURE ONLY. C entry points for tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// `--impl reference` legs (ctypes). PARITY UNPINNED — see ga.hpp.
#include <chrono>
#include <memory>
#include "raster.hpp"
#include <atomic>
#include <functional>
#include <thread>

// Minimal dynamic-schedule parallel for (this image's gcc ships no libgomp, so no OpenMP).
static void parallel_for(int64_t n, int threads, int64_t chunk, const std::function<void(int64_t, int)>& body) {
    if (threads <= 1 || n <= chunk) {
        for (int64_t i = 0; i < n; ++i) body(i, 0);
        return;
    }
    std::atomic<int64_t> next{0};
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([&, t]() {
            for (;;) {
                const int64_t b = next.fetch_add(chunk);
                if (b >= n) break;
                for (int64_t i = b; i < std::min(n, b + chunk); ++i) body(i, t);
            }
        });
    for (auto& th : pool) th.join();
}

using namespace oracle;

struct oracle_shape {
    Shape shape;
};

extern "C" {

int oracle_shape_from_paths(const cr_dynamic_stroke_options* groups, size_t n_groups, const cr_path_soa* soa, uint32_t path_begin,
                            uint32_t path_end, oracle_shape** out) {
    auto s = std::make_unique<oracle_shape>();
    const std::vector<Path> paths = paths_from_soa(*soa, path_begin, path_end);
    const int st = shape_from_paths(groups, n_groups, paths, s->shape);
    if (st != CR_OK) return st;
    *out = s.release();
    return CR_OK;
}
void oracle_shape_destroy(oracle_shape* s) { delete s; }
int oracle_shape_get_layout(const oracle_shape* s, cr_shape_layout* out) {
    for (int i = 0; i < 8; ++i) out->vertex_offsets[i] = s->shape.vertex_offsets[i];
    for (int i = 0; i < 3; ++i) out->index_offsets[i] = s->shape.index_offsets[i];
    out->dynamic_stroke_options_count = s->shape.dynamic_stroke_options_count;
    out->proto_hull_points = s->shape.proto_hull_points;
    return CR_OK;
}
const void* oracle_shape_vertex_buffer(const oracle_shape* s) { return s->shape.vertex_buffer.data(); }
const void* oracle_shape_index_buffer(const oracle_shape* s) { return s->shape.index_buffer.data(); }
const void* oracle_shape_stroke_buffer(const oracle_shape* s) { return s->shape.stroke_buffer.data(); }
int oracle_shape_set_dynamic_stroke_options(oracle_shape* s, size_t index, const cr_dynamic_stroke_options* o) {
    if (index >= s->shape.dynamic_stroke_options_count) return CR_ERR_DYNAMIC_STROKE_OPTIONS_INDEX_OUT_OF_BOUNDS;
    DynamicStrokeDescriptor d;
    const int st = convert_dynamic_stroke_options(*o, d);
    if (st != CR_OK) return st;
    std::memcpy(s->shape.stroke_buffer.data() + index * sizeof(d), &d, sizeof(d));
    return CR_OK;
}

// Tessellate many shapes (shape i = paths [begin[i], begin[i+1])) and report seconds spent; used as the CPU baseline.
// Results are discarded except for byte totals. threads <= 1: the reference's own sequential loop (src/renderer.rs:187).
int oracle_tessellate_batch(const cr_dynamic_stroke_options* groups, size_t n_groups, const cr_path_soa* soa, const uint32_t* begin,
                            uint32_t n_shapes, int threads, uint64_t* out_bytes, double* out_seconds) {
    const int nt = threads > 0 ? threads : 1;
    std::vector<uint64_t> bytes(nt, 0);
    std::atomic<int> status{CR_OK};
    const auto t0 = std::chrono::steady_clock::now();
    parallel_for((int64_t)n_shapes, nt, 16, [&](int64_t i, int t) {
        Shape s;
        const std::vector<Path> paths = paths_from_soa(*soa, begin[i], begin[i + 1]);
        const int st = shape_from_paths(groups, n_groups, paths, s);
        if (st != CR_OK) status = st;
        bytes[t] += s.vertex_buffer.size() + s.index_buffer.size();
    });
    const auto t1 = std::chrono::steady_clock::now();
    *out_bytes = 0;
    for (uint64_t b : bytes) *out_bytes += b;
    *out_seconds = std::chrono::duration<double>(t1 - t0).count();
    return status;
}

int oracle_render(const cr_config* config, uint32_t width, uint32_t height, oracle_shape* const* shapes, uint32_t n_shapes,
                  const RenderCommand* cmds, size_t n_cmds, const float* transforms, const float* colors, float* color, uint8_t* stencil,
                  float* alpha_layers, int threads, uint64_t* covered_samples, float* depth) {
    std::vector<RasterShape> rs(n_shapes);
    for (uint32_t i = 0; i < n_shapes; ++i) {
        const oracle_shape& s = *shapes[i];
        rs[i].vertex_buffer = s.shape.vertex_buffer.data();
        for (int k = 0; k < 8; ++k) rs[i].vertex_offsets[k] = s.shape.vertex_offsets[k];
        rs[i].wide_indices = s.shape.wide_indices.data();
        for (int k = 0; k < 3; ++k) rs[i].index_counts[k] = s.shape.index_counts[k];
        rs[i].stroke = reinterpret_cast<const DynamicStrokeDescriptor*>(s.shape.stroke_buffer.data());
        rs[i].n_groups = s.shape.dynamic_stroke_options_count;
    }
    const int n_threads = threads > 0 ? threads : 1;
    const int band = 16;
    const int n_bands = ((int)height + band - 1) / band;
    std::vector<uint64_t> covered(n_threads, 0);
    parallel_for(n_bands, n_threads, 1, [&](int64_t b, int t) {
        Framebuffer fb{width, height, config->msaa_sample_count, color, stencil, alpha_layers, 0, depth};
        render_band(*config, fb, rs.data(), cmds, n_cmds, transforms, colors, (int)b * band, std::min<int>(((int)b + 1) * band, (int)height));
        covered[t] += fb.covered_samples;
    });
    if (covered_samples) {
        *covered_samples = 0;
        for (uint64_t c : covered) *covered_samples += c;
    }
    return CR_OK;
}

int oracle_max_threads() {
    const unsigned n = std::thread::hardware_concurrency();
    return n ? (int)n : 1;
}

// ---- unit-level probes used by the anchor tests ------------------------------------------------------------
float oracle_atan2(float y, float x) { return cr::atan2_f(y, x); }
float oracle_acos(float x) { return cr::acos_f(x); }
void oracle_sincos(float a, float* s, float* c) { cr::sincos_f(a, s, c); }
float oracle_pow(float b, float e) { return cr::pow_pos_f(b, e); }
float oracle_wgsl_mod(float x, float y) { return cr::wgsl_mod(x, y); }
// coefficients ascending, degree 1..4; writes up to 4 (re, im, den) triples; returns count; *disc gets the discriminant.
int oracle_solve(int degree, const float* c, float margin, float* roots, float* disc, int* real_root) {
    cr::Roots r;
    switch (degree) {
        case 1: r = cr::solve_linear(c[0], c[1], margin); break;
        case 2: r = cr::solve_quadratic(c[0], c[1], c[2], margin); break;
        case 3: r = cr::solve_cubic(c[0], c[1], c[2], c[3], margin); break;
        default: r = cr::solve_quartic(c[0], c[1], c[2], c[3], c[4], margin); break;
    }
    for (int i = 0; i < r.count; ++i) {
        roots[3 * i] = r.r[i].numerator.re;
        roots[3 * i + 1] = r.r[i].numerator.im;
        roots[3 * i + 2] = r.r[i].denominator;
    }
    *disc = r.discriminant;
    *real_root = r.real_root;
    return r.count;
}
// kind: cr_segment_type of a curve; cp: control points incl. the start (3 or 4 xy pairs); w: weights (rational) or null.
// Returns the number of parameters written (capacity-limited).
int oracle_uniform_tangent_angle(int kind, const float* cp, const float* w, float angle_step, float* out, int capacity) {
    std::vector<float> params;
    if (kind == CR_SEG_INTEGRAL_QUADRATIC || kind == CR_SEG_RATIONAL_QUADRATIC) {
        std::array<Point, 3> pts = {vec_to_point(cp), kind == CR_SEG_RATIONAL_QUADRATIC ? weighted_vec_to_point(w[0], cp + 2) : vec_to_point(cp + 2),
                                    vec_to_point(cp + 4)};
        const Point plain[3] = {vec_to_point(cp), vec_to_point(cp + 2), vec_to_point(cp + 4)};
        Plane s, e;
        get_quadratic_tangents(plain, s, e);
        const auto pb = rational_quadratic_control_points_to_power_basis(pts);
        params = kind == CR_SEG_INTEGRAL_QUADRATIC ? integral_quadratic_uniform_tangent_angle(pb, s, e, angle_step)
                                                   : rational_quadratic_uniform_tangent_angle(pb, s, e, angle_step);
    } else {
        std::array<Point, 4> pts;
        for (int i = 0; i < 4; ++i) pts[i] = kind == CR_SEG_RATIONAL_CUBIC ? weighted_vec_to_point(w[i], cp + 2 * i) : vec_to_point(cp + 2 * i);
        const auto pb = rational_cubic_control_points_to_power_basis(pts);
        params = kind == CR_SEG_INTEGRAL_CUBIC ? integral_cubic_uniform_tangent_angle(pb, angle_step) : rational_cubic_uniform_tangent_angle(pb, angle_step);
    }
    const int n = std::min<int>((int)params.size(), capacity);
    for (int i = 0; i < n; ++i) out[i] = params[i];
    return (int)params.size();
}
// point and unit tangent normal (Plane g1, g2) of a curve at t, for invariant tests.
void oracle_curve_eval(int kind, const float* cp, const float* w, float t, float* out_xy, float* out_normal) {
    Point p;
    Plane d;
    if (kind == CR_SEG_INTEGRAL_QUADRATIC || kind == CR_SEG_RATIONAL_QUADRATIC) {
        std::array<Point, 3> pts = {vec_to_point(cp), kind == CR_SEG_RATIONAL_QUADRATIC ? weighted_vec_to_point(w[0], cp + 2) : vec_to_point(cp + 2),
                                    vec_to_point(cp + 4)};
        const auto pb = rational_quadratic_control_points_to_power_basis(pts);
        p = rational_quadratic_point(pb, t);
        d = signum(rational_quadratic_first_order_derivative(pb, t));
    } else {
        std::array<Point, 4> pts;
        for (int i = 0; i < 4; ++i) pts[i] = kind == CR_SEG_RATIONAL_CUBIC ? weighted_vec_to_point(w[i], cp + 2 * i) : vec_to_point(cp + 2 * i);
        const auto pb = rational_cubic_control_points_to_power_basis(pts);
        p = rational_cubic_point(pb, t);
        d = signum(rational_cubic_first_order_derivative(pb, t));
    }
    point_to_vec(p, out_xy);
    out_normal[0] = d[1];
    out_normal[1] = d[2];
}
// ppga2d sign conventions, exposed so that tests can pin them against the reference's own GA-free statements
// (src/utils.rs:80-101 "Expects the vertices to be ordered clockwise", src/stroke.rs:272-281 cap direction).
float oracle_ga_triple(const float* a, const float* b, const float* c) { return triple(vec_to_point(a), vec_to_point(b), vec_to_point(c)); }
void oracle_ga_join(const float* p, const float* q, float* out_plane) {
    const Plane l = regressive(vec_to_point(p), vec_to_point(q));
    out_plane[0] = l[0]; out_plane[1] = l[1]; out_plane[2] = l[2];
}
int oracle_andrew(const float* xy, uint32_t n, float* out_xy) {
    std::vector<Vertex0> pts(n);
    for (uint32_t i = 0; i < n; ++i) pts[i] = Vertex0{{cr::canon_zero(xy[2 * i]), cr::canon_zero(xy[2 * i + 1])}};
    const std::vector<Vertex0> hull = andrew(pts);
    for (size_t i = 0; i < hull.size(); ++i) { out_xy[2 * i] = hull[i].p[0]; out_xy[2 * i + 1] = hull[i].p[1]; }
    return (int)hull.size();
}

}  // extern "C"
