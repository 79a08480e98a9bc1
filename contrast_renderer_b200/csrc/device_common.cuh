// device_common.cuh — shared declarations of the sm_100a kernels: error plumbing, the packed vertex layouts of
// src/vertex.rs:1-26, and the ppga2d closed forms as __device__ code.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/contrast_b200.h"
#include "arith/cr_arith.h"

#define CR_ERROR_MARGIN 0.0001f        // src/error.rs:19
#define CR_F32_EPSILON 1.1920929e-7f   // f32::EPSILON

// ---------------------------------------------------------------------------------------- category numbering
// Vertex categories in the order of concat_buffers! (src/renderer.rs:198-207).
enum {
    CAT_LINE = 0,      // Vertex2f1i 20 B   stroke_builder.line_vertices
    CAT_JOINT = 1,     // Vertex3f1i 24 B   stroke_builder.joint_vertices
    CAT_SOLID = 2,     // Vertex0     8 B   fill_builder.solid_vertices
    CAT_IQ = 3,        // Vertex2f   16 B   integral quadratic
    CAT_IC = 4,        // Vertex3f   20 B   integral cubic
    CAT_RQ = 5,        // Vertex3f   20 B   rational quadratic
    CAT_RC = 6,        // Vertex4f   24 B   rational cubic
    CAT_HULL = 7,      // Vertex0     8 B   convex hull (strip order)
    // counters that are not vertex arrays
    CNT_PROTO = 7,     // proto_hull points (src/renderer.rs:184); shares slot 7 during count/emit (hull is produced later)
    CNT_LINE_IDX = 8,  // line index slots incl. restarts
    CNT_JOINT_IDX = 9,
    CNT_SOLID_IDX = 10,
    CNT_COUNT = 11
};
static const int kCategoryStride[8] = {20, 24, 8, 16, 20, 20, 24, 8};

#define CR_RESTART 0xFFFFFFFFu

// Device error flag bits (or-ed into one word, decoded by the host in this priority order).
#define CR_DEVERR_GROUP_OOB 1u
#define CR_DEVERR_NON_FINITE 2u
#define CR_DEVERR_STEPS 4u
#define CR_DEVERR_CUBIC 8u
#define CR_DEVERR_BAD_TABLES 16u   // cr_path_soa cursor tables are inconsistent (checked before anything is indexed with them)
#define CR_DEVERR_CAPACITY 32u     // (optimistic rebuild) an output array of the previous build is too small for this one: nothing is emitted, the host re-runs
#define CR_DEVERR_MODE 64u         // (optimistic rebuild) the batch holds a cubic segment but the kernels without the cubic builder were launched
#define CR_DEVERR_FATAL_MASK (CR_DEVERR_GROUP_OOB | CR_DEVERR_BAD_TABLES | CR_DEVERR_CAPACITY | CR_DEVERR_MODE)   // set before the emit pass, which then must not run

namespace crd {

struct Pt { float g0, g1, g2; };   // ppga2d::Point (e12, e01, -e02) = (w, w*x, w*y)
struct Ln { float g0, g1, g2; };   // ppga2d::Plane (e0, e2, e1): g0 + g1*x + g2*y = 0

__device__ __forceinline__ Pt mk_pt(float a, float b, float c) { Pt p; p.g0 = a; p.g1 = b; p.g2 = c; return p; }
__device__ __forceinline__ Ln mk_ln(float a, float b, float c) { Ln l; l.g0 = a; l.g1 = b; l.g2 = c; return l; }
__device__ __forceinline__ Pt operator*(Pt p, float s) { return mk_pt(p.g0 * s, p.g1 * s, p.g2 * s); }
__device__ __forceinline__ Pt operator+(Pt a, Pt b) { return mk_pt(a.g0 + b.g0, a.g1 + b.g1, a.g2 + b.g2); }
__device__ __forceinline__ Ln operator*(Ln p, float s) { return mk_ln(p.g0 * s, p.g1 * s, p.g2 * s); }
__device__ __forceinline__ Ln operator+(Ln a, Ln b) { return mk_ln(a.g0 + b.g0, a.g1 + b.g1, a.g2 + b.g2); }
__device__ __forceinline__ Ln operator-(Ln a, Ln b) { return mk_ln(a.g0 - b.g0, a.g1 - b.g1, a.g2 - b.g2); }
__device__ __forceinline__ Ln neg(Ln a) { return mk_ln(-a.g0, -a.g1, -a.g2); }
__device__ __forceinline__ Ln dual(Pt p) { return mk_ln(p.g0, p.g1, p.g2); }
// Point v Point
__device__ __forceinline__ Ln join(Pt p, Pt q) { return mk_ln(p.g2 * q.g1 - p.g1 * q.g2, p.g0 * q.g2 - p.g2 * q.g0, p.g1 * q.g0 - p.g0 * q.g1); }
// Point v Plane
__device__ __forceinline__ float incidence(Pt p, Ln a) { return p.g0 * a.g0 + p.g1 * a.g1 + p.g2 * a.g2; }
// Plane ^ Plane
__device__ __forceinline__ Pt meet(Ln a, Ln b) { return mk_pt(a.g2 * b.g1 - a.g1 * b.g2, a.g0 * b.g2 - a.g2 * b.g0, a.g1 * b.g0 - a.g0 * b.g1); }
// Plane . Plane
__device__ __forceinline__ float dot(Ln a, Ln b) { return a.g1 * b.g1 + a.g2 * b.g2; }
__device__ __forceinline__ float sqmag(Ln a) { return a.g1 * a.g1 + a.g2 * a.g2; }
__device__ __forceinline__ float mag(Ln a) { return cr::sqrt_f(sqmag(a)); }
__device__ __forceinline__ Ln unit(Ln a) { return a * (1.0f / mag(a)); }
// line through p parallel to a (scaled by -p0^2), src/stroke.rs:71-75
__device__ __forceinline__ Ln parallel_through(Ln a, Pt p) {
    const float t = a.g1 * p.g1 + a.g2 * p.g2;
    return mk_ln(p.g0 * t, -(p.g0 * (a.g1 * p.g0)), -(p.g0 * (a.g2 * p.g0)));
}
__device__ __forceinline__ Pt intersect(Ln a, Ln b) { const Pt p = meet(a, b); return p * (1.0f / p.g0); }
__device__ __forceinline__ Ln rot90cw(Ln v) { return mk_ln(0.0f, v.g2, -v.g1); }
__device__ __forceinline__ float2 to_vec(Pt p) { return make_float2(p.g1 / p.g0, p.g2 / p.g0); }
__device__ __forceinline__ Pt from_vec(float x, float y) { return mk_pt(1.0f, x, y); }
__device__ __forceinline__ Pt from_wvec(float w, float x, float y) { return mk_pt(w, x * w, y * w); }
__device__ __forceinline__ float triple(Pt a, Pt b, Pt c) { return incidence(c, join(a, b)); }

}  // namespace crd

// ------------------------------------------------------------------------------------------- host-side helpers
struct CrCudaError { cudaError_t err; const char* file; int line; };
void cr_set_error_message(const char* fmt, ...);
#define CR_CUDA_TRY(expr)                                                                                    \
    do {                                                                                                     \
        cudaError_t _e = (expr);                                                                             \
        if (_e != cudaSuccess) {                                                                             \
            cr_set_error_message("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return CR_ERR_CUDA;                                                                              \
        }                                                                                                    \
    } while (0)

// Last index i of a non-decreasing array with arr[i] <= key (arr[0] <= key), found by the whole warp: 32 probes per round,
// so a table of a million entries takes 4 dependent loads instead of 20.
__device__ __forceinline__ uint32_t warp_search_last_le(const uint32_t* __restrict__ arr, uint32_t n, uint32_t key, uint32_t lane) {
    uint32_t lo = 0, len = n;
    while (len > 1u) {
        const uint32_t step = (len + 31u) / 32u, pos = lo + lane * step;
        const bool le = pos < lo + len && arr[pos] <= key;
        const uint32_t hit = __ballot_sync(0xffffffffu, le);          // monotone: lanes 0 .. j
        const uint32_t j = 31u - (uint32_t)__clz((int)(hit | 1u));
        const uint32_t end = lo + len;
        lo += j * step;
        len = min(step, end - lo);
    }
    return lo;
}
