// cr_arith.h — the arithmetic contract shared by the CUDA kernels and the CPU oracle.
//
// The reference (Lichtso/contrast_renderer) delegates all of its scalar arithmetic to two things that are
// NOT in /root/reference: the platform libm (f32::atan2/acos/powf/... via Rust std) and the un-vendored crate
// `geometric_algebra 0.3.0` (Cargo.lock:520) — `epga1d::ComplexNumber` and `polynomial::{solve_linear,
// solve_quadratic, solve_cubic, solve_quartic, Root}` (call sites: src/curve.rs:201,206,230-243,318,342,368,405).
// libm results differ between glibc and CUDA by 1-2 ulp, which would flip `(x + 0.5) as usize` step counts
// (src/curve.rs:233) and therefore vertex COUNTS. This header therefore defines those functions once, using
// only IEEE-754 correctly rounded operations (+ - * / sqrt, int<->float conversions), so that g++
// (-ffp-contract=off) and nvcc (-fmad=false -prec-div=true -prec-sqrt=true) produce bit-identical results.
// Everything ABOVE this layer (curve.rs / stroke.rs / fill.rs / convex_hull.rs / renderer.rs / shaders.wgsl)
// is written twice, independently: once in oracle/ (sequential C++) and once in csrc/*.cu (kernels).
//
// Elementary functions are evaluated in binary64 with short Taylor/Horner kernels after exact range reduction
// and rounded once to binary32 (error < 0.5000001 ulp, checked against numpy in tests/test_arith.py).
#pragma once
#include <stdint.h>
#include <string.h>
#include <math.h>

#if defined(__CUDACC__)
#define CR_HD __host__ __device__ __forceinline__
#else
#define CR_HD inline
#endif

namespace cr {

// ---------------------------------------------------------------------------------------------- bit helpers
CR_HD uint32_t f32_bits(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
CR_HD float bits_f32(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}
CR_HD uint64_t f64_bits(double d) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t u; memcpy(&u, &d, 8); return u;
#endif
}
CR_HD double bits_f64(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double d; memcpy(&d, &u, 8); return d;
#endif
}
CR_HD bool is_nan(float f) { return f != f; }
CR_HD bool is_finite(float f) { return (f32_bits(f) & 0x7f800000u) != 0x7f800000u; }
CR_HD float fabs_f(float f) { return bits_f32(f32_bits(f) & 0x7fffffffu); }
CR_HD double fabs_d(double d) { return bits_f64(f64_bits(d) & 0x7fffffffffffffffull); }
CR_HD float copysign_f(float mag, float sgn) { return bits_f32((f32_bits(mag) & 0x7fffffffu) | (f32_bits(sgn) & 0x80000000u)); }
CR_HD bool sign_bit(float f) { return (f32_bits(f) >> 31) != 0; }
/// Rust `f32::signum`: 1.0 for +0.0 and positive, -1.0 for -0.0 and negative, NaN for NaN (src/stroke.rs:66).
CR_HD float rust_signum(float f) { return is_nan(f) ? f : (sign_bit(f) ? -1.0f : 1.0f); }
/// SafeFloat canonicalisation: -0.0 -> +0.0 (src/safe_float.rs:46-49,113-118).
CR_HD float canon_zero(float f) { return f32_bits(f) == 0x80000000u ? 0.0f : f; }

CR_HD float sqrt_f(float x) {
#if defined(__CUDA_ARCH__)
    return __fsqrt_rn(x);
#else
    return sqrtf(x);
#endif
}
CR_HD double sqrt_d(double x) {
#if defined(__CUDA_ARCH__)
    return __dsqrt_rn(x);
#else
    return sqrt(x);
#endif
}
/// floor for |x| < 2^52 without calling libm (exact).
CR_HD double floor_d(double x) {
    if (!(fabs_d(x) < 4503599627370496.0)) return x;
    double t = (double)(long long)x;  // trunc
    return (t > x) ? t - 1.0 : t;
}
CR_HD float trunc_f(float x) {
    if (!(fabs_f(x) < 8388608.0f)) return x;
    return copysign_f((float)(int)x, x);
}
CR_HD float floor_f(float x) {
    if (!(fabs_f(x) < 8388608.0f)) return x;
    float t = (float)(int)x;
    return (t > x) ? t - 1.0f : t;
}
/// WGSL float `%` (src/shaders.wgsl:211): x - y * trunc(x / y).
CR_HD float wgsl_mod(float x, float y) { return x - y * trunc_f(x / y); }

/// Rust `f32 as usize` (saturating, NaN -> 0), additionally capped to 2^31-1 (src/curve.rs:233).
CR_HD uint32_t f32_to_usize_sat(float f) {
    if (!(f > 0.0f)) return 0u;
    if (f >= 2147483648.0f) return 0x7fffffffu;
    return (uint32_t)f;
}

// ---------------------------------------------------------------------------------- binary64 kernels
#define CR_PI_D 3.14159265358979323846
#define CR_PIO2_HI 1.57079632679489655800e+00
#define CR_PIO2_LO 6.12323399573676603587e-17
#define CR_LN2_HI 6.93147180369123816490e-01
#define CR_LN2_LO 1.90821492927058770002e-10

/// atan on |x| <= tan(pi/12): odd Taylor series through x^19.
CR_HD double atan_core_d(double x) {
    const double x2 = x * x;
    double s = -1.0 / 19.0;
    s = s * x2 + 1.0 / 17.0;
    s = s * x2 - 1.0 / 15.0;
    s = s * x2 + 1.0 / 13.0;
    s = s * x2 - 1.0 / 11.0;
    s = s * x2 + 1.0 / 9.0;
    s = s * x2 - 1.0 / 7.0;
    s = s * x2 + 1.0 / 5.0;
    s = s * x2 - 1.0 / 3.0;
    s = s * x2 + 1.0;
    return x * s;
}
/// atan for 0 <= x <= 1.
CR_HD double atan_unit_d(double x) {
    if (x > 0.26794919243112270647) {
        const double sqrt3 = 1.73205080756887729353;
        return CR_PI_D / 6.0 + atan_core_d((sqrt3 * x - 1.0) / (sqrt3 + x));
    }
    return atan_core_d(x);
}
CR_HD double atan2_d(double y, double x) {
    if (x != x || y != y) return x + y;
    const double ax = fabs_d(x), ay = fabs_d(y);
    double r;
    if (ax == 0.0 && ay == 0.0) r = 0.0;
    else if (ay <= ax) r = atan_unit_d(ay / ax);   // inf/inf -> NaN like a degenerate input; callers never pass it
    else r = CR_PI_D / 2.0 - atan_unit_d(ax / ay);
    if ((f64_bits(x) >> 63) != 0) r = CR_PI_D - r;
    return ((f64_bits(y) >> 63) != 0) ? -r : r;
}
/// sin and cos of a finite angle; Cody-Waite reduction by pi/2, Taylor on [-pi/4, pi/4].
CR_HD void sincos_d(double a, double* s, double* c) {
    if (!(fabs_d(a) < 1.0e9)) { *s = a - a; *c = a - a; return; }  // inf/NaN -> NaN; huge -> 0 (never used)
    const double kf = floor_d(a * (2.0 / CR_PI_D) + 0.5);
    const double r = (a - kf * CR_PIO2_HI) - kf * CR_PIO2_LO;
    const long long k = (long long)kf;
    const double r2 = r * r;
    double ps = -1.0 / 1307674368000.0;        // r^15/15!
    ps = ps * r2 + 1.0 / 6227020800.0;         // 13!
    ps = ps * r2 - 1.0 / 39916800.0;           // 11!
    ps = ps * r2 + 1.0 / 362880.0;             // 9!
    ps = ps * r2 - 1.0 / 5040.0;               // 7!
    ps = ps * r2 + 1.0 / 120.0;                // 5!
    ps = ps * r2 - 1.0 / 6.0;                  // 3!
    ps = ps * r2 + 1.0;
    const double sr = r * ps;
    double pc = 1.0 / 20922789888000.0;        // r^16/16!
    pc = pc * r2 - 1.0 / 87178291200.0;        // 14!
    pc = pc * r2 + 1.0 / 479001600.0;          // 12!
    pc = pc * r2 - 1.0 / 3628800.0;            // 10!
    pc = pc * r2 + 1.0 / 40320.0;              // 8!
    pc = pc * r2 - 1.0 / 720.0;                // 6!
    pc = pc * r2 + 1.0 / 24.0;                 // 4!
    pc = pc * r2 - 1.0 / 2.0;                  // 2!
    const double cr_ = pc * r2 + 1.0;
    switch ((int)(k & 3)) {
        case 0: *s = sr; *c = cr_; break;
        case 1: *s = cr_; *c = -sr; break;
        case 2: *s = -sr; *c = -cr_; break;
        default: *s = -cr_; *c = sr; break;
    }
}
/// natural log of a positive, finite, normal binary64.
CR_HD double log_d(double x) {
    uint64_t b = f64_bits(x);
    long long e = (long long)((b >> 52) & 0x7ff) - 1023;
    double m = bits_f64((b & 0x000fffffffffffffull) | 0x3ff0000000000000ull);  // [1,2)
    if (m > 1.41421356237309504880) { m = m * 0.5; e += 1; }
    const double s = (m - 1.0) / (m + 1.0);
    const double s2 = s * s;
    double p = 1.0 / 21.0;
    p = p * s2 + 1.0 / 19.0;
    p = p * s2 + 1.0 / 17.0;
    p = p * s2 + 1.0 / 15.0;
    p = p * s2 + 1.0 / 13.0;
    p = p * s2 + 1.0 / 11.0;
    p = p * s2 + 1.0 / 9.0;
    p = p * s2 + 1.0 / 7.0;
    p = p * s2 + 1.0 / 5.0;
    p = p * s2 + 1.0 / 3.0;
    p = p * s2 + 1.0;
    return ((double)e * CR_LN2_HI + (double)e * CR_LN2_LO) + 2.0 * s * p;
}
CR_HD double exp_d(double y) {
    if (y != y) return y;
    if (y > 700.0) return bits_f64(0x7ff0000000000000ull);
    if (y < -700.0) return 0.0;
    const double kf = floor_d(y * 1.44269504088896340736 + 0.5);
    const double r = (y - kf * CR_LN2_HI) - kf * CR_LN2_LO;
    double p = 1.0 / 6227020800.0;  // 13!
    p = p * r + 1.0 / 479001600.0;
    p = p * r + 1.0 / 39916800.0;
    p = p * r + 1.0 / 3628800.0;
    p = p * r + 1.0 / 362880.0;
    p = p * r + 1.0 / 40320.0;
    p = p * r + 1.0 / 5040.0;
    p = p * r + 1.0 / 720.0;
    p = p * r + 1.0 / 120.0;
    p = p * r + 1.0 / 24.0;
    p = p * r + 1.0 / 6.0;
    p = p * r + 0.5;
    p = p * r + 1.0;
    p = p * r + 1.0;
    const long long k = (long long)kf;  // |k| <= 1011
    return p * bits_f64((uint64_t)(k + 1023) << 52);
}

// ---------------------------------------------------------------------------------- binary32 front ends
CR_HD float atan2_f(float y, float x) { return (float)atan2_d((double)y, (double)x); }
CR_HD float acos_f(float x) {
    const double xd = (double)x;
    if (!(fabs_d(xd) <= 1.0)) return bits_f32(0x7fc00000u);
    return (float)atan2_d(sqrt_d((1.0 - xd) * (1.0 + xd)), xd);
}
CR_HD void sincos_f(float a, float* s, float* c) {
    double sd, cd;
    sincos_d((double)a, &sd, &cd);
    *s = (float)sd; *c = (float)cd;
}
/// base^e for base >= 0 (magnitudes only).
CR_HD float pow_pos_f(float base, float e) {
    if (is_nan(base) || is_nan(e)) return base + e;
    if (e == 0.0f) return 1.0f;
    if (base == 0.0f) return e > 0.0f ? 0.0f : bits_f32(0x7f800000u);
    if (!is_finite(base)) return e > 0.0f ? base : 0.0f;
    return (float)exp_d((double)e * log_d((double)base));
}

// ------------------------------------------------------------------------------ epga1d::ComplexNumber
struct Complex {
    float re, im;
};
CR_HD Complex cplx(float re, float im) { Complex c; c.re = re; c.im = im; return c; }
CR_HD Complex operator+(Complex a, Complex b) { return cplx(a.re + b.re, a.im + b.im); }
CR_HD Complex operator-(Complex a, Complex b) { return cplx(a.re - b.re, a.im - b.im); }
CR_HD Complex operator-(Complex a) { return cplx(-a.re, -a.im); }
CR_HD Complex operator*(Complex a, float s) { return cplx(a.re * s, a.im * s); }
/// geometric_product of two complex numbers.
CR_HD Complex cmul(Complex a, Complex b) { return cplx(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
/// geometric_quotient a / b (src/curve.rs:232).
CR_HD Complex cdiv(Complex a, Complex b) {
    const float d = b.re * b.re + b.im * b.im;
    return cplx((a.re * b.re + a.im * b.im) / d, (a.im * b.re - a.re * b.im) / d);
}
CR_HD float cabs_f(Complex a) { return sqrt_f(a.re * a.re + a.im * a.im); }
CR_HD float carg(Complex a) { return atan2_f(a.im, a.re); }
/// Powf: polar(|z|^x, arg(z) * x)  (src/curve.rs:234).
CR_HD Complex cpowf(Complex a, float x) {
    const float m = pow_pos_f(cabs_f(a), x);
    float s, c;
    sincos_f(carg(a) * x, &s, &c);
    return cplx(m * c, m * s);
}
/// Powi: exponentiation by squaring, n >= 1 (src/curve.rs:237).
CR_HD Complex cpowi(Complex a, uint32_t n) {
    Complex x = a, y = cplx(1.0f, 0.0f);
    while (n > 1u) {
        if (n & 1u) y = cmul(x, y);
        x = cmul(x, x);
        n >>= 1;
    }
    return cmul(x, y);
}
/// principal square root (algebraic form, no transcendental calls).
CR_HD Complex csqrt(Complex a) {
    if (a.im == 0.0f) return a.re >= 0.0f ? cplx(sqrt_f(a.re), 0.0f) : cplx(0.0f, sqrt_f(-a.re));
    const float m = cabs_f(a);
    if (a.re >= 0.0f) {
        const float t = sqrt_f((m + a.re) * 0.5f);
        return cplx(t, a.im / (2.0f * t));
    }
    const float t = sqrt_f((m - a.re) * 0.5f);
    return cplx(fabs_f(a.im) / (2.0f * t), copysign_f(t, a.im));
}
/// principal cube root.
CR_HD Complex ccbrt(Complex a) {
    const float m = cabs_f(a);
    if (m == 0.0f) return cplx(0.0f, 0.0f);
    const float r = pow_pos_f(m, 1.0f / 3.0f);
    float s, c;
    sincos_f(carg(a) / 3.0f, &s, &c);
    return cplx(r * c, r * s);
}

// ------------------------------------------------------------------ geometric_algebra::polynomial contract
/// A root in homogeneous form numerator / denominator. `denominator == 0` marks "no root".
struct Root {
    Complex numerator;
    float denominator;
};
CR_HD Root make_root(float re, float im, float den) { Root r; r.numerator = cplx(re, im); r.denominator = den; return r; }
CR_HD Root no_root() { return make_root(1.0f, 0.0f, 0.0f); }  // Root::new([1.0, 0.0], 0.0) (src/curve.rs:157)

struct Roots {
    float discriminant;
    int count;        // valid entries of r[]
    int real_root;    // solve_cubic only: index of the (most) real root
    Root r[4];
};

/// 0 = c0 + c1 x.
CR_HD Roots solve_linear(float c0, float c1, float margin) {
    Roots o; o.count = 0; o.real_root = 0; o.discriminant = 0.0f;
    for (int i = 0; i < 4; ++i) o.r[i] = no_root();
    if (fabs_f(c1) <= margin) return o;
    o.discriminant = 1.0f;
    o.count = 1;
    o.r[0] = make_root(-c0, 0.0f, c1);
    return o;
}
/// 0 = c0 + c1 x + c2 x^2; discriminant > 0 <=> two distinct real roots (src/curve.rs:214).
CR_HD Roots solve_quadratic(float c0, float c1, float c2, float margin) {
    if (fabs_f(c2) <= margin) return solve_linear(c0, c1, margin);
    Roots o; o.real_root = 0;
    for (int i = 0; i < 4; ++i) o.r[i] = no_root();
    const float disc = c1 * c1 - 4.0f * c2 * c0;
    const Complex q = csqrt(cplx(disc, 0.0f));
    o.discriminant = disc;
    o.count = 2;
    o.r[0] = make_root(-c1 - q.re, -q.im, 2.0f * c2);
    o.r[1] = make_root(-c1 + q.re, q.im, 2.0f * c2);
    return o;
}
// The cubic and quartic closed forms lose most of their digits to cancellation in binary32 (root errors of 0.1 were seen
// on well-conditioned stroke quartics), so they are evaluated in binary64 complex arithmetic — built, like everything in
// this header, from correctly rounded operations only — and rounded to binary32 once at the end.
struct ComplexD {
    double re, im;
};
CR_HD ComplexD cplxd(double re, double im) { ComplexD c; c.re = re; c.im = im; return c; }
CR_HD ComplexD operator+(ComplexD a, ComplexD b) { return cplxd(a.re + b.re, a.im + b.im); }
CR_HD ComplexD operator-(ComplexD a, ComplexD b) { return cplxd(a.re - b.re, a.im - b.im); }
CR_HD ComplexD operator-(ComplexD a) { return cplxd(-a.re, -a.im); }
CR_HD ComplexD operator*(ComplexD a, double s) { return cplxd(a.re * s, a.im * s); }
CR_HD ComplexD cmul_d(ComplexD a, ComplexD b) { return cplxd(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
CR_HD ComplexD cdiv_d(ComplexD a, ComplexD b) {
    const double d = b.re * b.re + b.im * b.im;
    return cplxd((a.re * b.re + a.im * b.im) / d, (a.im * b.re - a.re * b.im) / d);
}
CR_HD bool is_zero_d(ComplexD a) { return a.re == 0.0 && a.im == 0.0; }
CR_HD double cabs_d(ComplexD a) { return sqrt_d(a.re * a.re + a.im * a.im); }
CR_HD ComplexD csqrt_d(ComplexD a) {
    if (a.im == 0.0) return a.re >= 0.0 ? cplxd(sqrt_d(a.re), 0.0) : cplxd(0.0, sqrt_d(-a.re));
    const double m = cabs_d(a);
    if (a.re >= 0.0) {
        const double t = sqrt_d((m + a.re) * 0.5);
        return cplxd(t, a.im / (2.0 * t));
    }
    const double t = sqrt_d((m - a.re) * 0.5);
    return cplxd(fabs_d(a.im) / (2.0 * t), (f64_bits(a.im) >> 63) ? -t : t);
}
CR_HD ComplexD ccbrt_d(ComplexD a) {
    const double m = cabs_d(a);
    if (!(m > 0.0)) return cplxd(0.0, 0.0);
    const double r = exp_d(log_d(m) / 3.0);
    double s, c;
    sincos_d(atan2_d(a.im, a.re) / 3.0, &s, &c);
    return cplxd(r * c, r * s);
}

/// 0 = c0 + c1 x + c2 x^2 + c3 x^3 (general cubic formula). discriminant > 0: three real roots (serpentine),
/// = 0: repeated root (cusp), < 0: one real root (loop) — the Loop-Blinn convention used by src/fill.rs:53-65.
CR_HD Roots solve_cubic(float c0, float c1, float c2, float c3, float margin) {
    if (fabs_f(c3) <= margin) {
        Roots o = solve_quadratic(c0, c1, c2, margin);
        o.real_root = 0;
        return o;
    }
    Roots o;
    for (int i = 0; i < 4; ++i) o.r[i] = no_root();
    const double a = c3, b = c2, c = c1, d = c0;
    const double d0 = b * b - 3.0 * a * c;
    const double d1 = 2.0 * b * b * b - 9.0 * a * b * c + 27.0 * a * a * d;
    const double inner = d1 * d1 - 4.0 * d0 * d0 * d0;
    o.discriminant = (float)(-inner / (27.0 * a * a));
    const ComplexD s = csqrt_d(cplxd(inner, 0.0));
    ComplexD cc = ccbrt_d((cplxd(d1, 0.0) + s) * 0.5);
    if (is_zero_d(cc)) cc = ccbrt_d((cplxd(d1, 0.0) - s) * 0.5);
    const float den = (float)(3.0 * a);
    o.count = 3;
    if (is_zero_d(cc)) {  // triple root
        for (int k = 0; k < 3; ++k) o.r[k] = make_root((float)-b, 0.0f, den);
        o.real_root = 0;
        return o;
    }
    const ComplexD xi = cplxd(-0.5, 0.86602540378443864676);
    ComplexD u = cc;
    double best = 0.0;
    o.real_root = 0;
    for (int k = 0; k < 3; ++k) {
        const ComplexD t = cdiv_d(cplxd(d0, 0.0), u);
        const ComplexD n = -(cplxd(b, 0.0) + u + t);
        o.r[k] = make_root((float)n.re, (float)n.im, den);
        if (k == 0 || fabs_d(n.im) < best) { best = fabs_d(n.im); o.real_root = k; }
        u = cmul_d(u, xi);
    }
    return o;
}
/// 0 = c0 + ... + c4 x^4 (Ferrari / general quartic formula in complex arithmetic).
/// Contract: roots are returned most-real first (stable insertion sort on |Im|). The callers take the FIRST root whose
/// real part lies in [0, 1] without looking at the imaginary part (src/curve.rs:239-247), so a complex pair must not shadow
/// a genuine real solution.
CR_HD Roots solve_quartic(float c0, float c1, float c2, float c3, float c4, float margin) {
    if (fabs_f(c4) <= margin) return solve_cubic(c0, c1, c2, c3, margin);
    Roots o; o.real_root = 0;
    for (int i = 0; i < 4; ++i) o.r[i] = no_root();
    const double a = c4, b = c3, c = c2, d = c1, e = c0;
    const double p = (8.0 * a * c - 3.0 * b * b) / (8.0 * a * a);
    const double q = (b * b * b - 4.0 * a * b * c + 8.0 * a * a * d) / (8.0 * a * a * a);
    const double d0 = c * c - 3.0 * b * d + 12.0 * a * e;
    const double d1 = 2.0 * c * c * c - 9.0 * b * c * d + 27.0 * b * b * e + 27.0 * a * d * d - 72.0 * a * c * e;
    const double inner = d1 * d1 - 4.0 * d0 * d0 * d0;
    o.discriminant = (float)(-inner / 27.0);
    const ComplexD sq = csqrt_d(cplxd(inner, 0.0));
    ComplexD qq = ccbrt_d((cplxd(d1, 0.0) + sq) * 0.5);
    if (is_zero_d(qq)) qq = ccbrt_d((cplxd(d1, 0.0) - sq) * 0.5);
    const ComplexD xi = cplxd(-0.5, 0.86602540378443864676);
    const double m23p = -2.0 / 3.0 * p;
    ComplexD S = cplxd(0.0, 0.0);
    for (int k = 0; k < 3; ++k) {  // pick the first cube root that gives S != 0
        ComplexD t = cplxd(0.0, 0.0);
        if (!is_zero_d(qq)) t = (qq + cdiv_d(cplxd(d0, 0.0), qq)) * (1.0 / (3.0 * a));
        S = csqrt_d(cplxd(m23p, 0.0) + t) * 0.5;
        if (!is_zero_d(S)) break;
        qq = cmul_d(qq, xi);
    }
    const double mb4a = -b / (4.0 * a);
    o.count = 4;
    ComplexD r[4];
    if (is_zero_d(S)) {  // depressed quartic y^4 + p y^2 + r = 0 with q == 0 and S == 0: biquadratic
        const ComplexD h = csqrt_d(cplxd(-2.0 * p, 0.0)) * 0.5;  // roots -b/4a ± sqrt(-2p)/2 (double)
        r[0] = cplxd(mb4a + h.re, h.im);
        r[1] = cplxd(mb4a - h.re, -h.im);
        r[2] = r[0];
        r[3] = r[1];
    } else {
        const ComplexD s2 = cmul_d(S, S);
        const ComplexD base = cplxd(-2.0 * p, 0.0) - s2 * 4.0;
        const ComplexD qs = cdiv_d(cplxd(q, 0.0), S);
        const ComplexD h1 = csqrt_d(base + qs) * 0.5;
        const ComplexD h2 = csqrt_d(base - qs) * 0.5;
        const ComplexD m = cplxd(mb4a, 0.0);
        r[0] = m - S + h1; r[1] = m - S - h1; r[2] = m + S + h2; r[3] = m + S - h2;
    }
    for (int i = 1; i < 4; ++i) {
        const ComplexD key = r[i];
        const double k = fabs_d(key.im);
        int j = i - 1;
        while (j >= 0 && fabs_d(r[j].im) > k) { r[j + 1] = r[j]; --j; }
        r[j + 1] = key;
    }
    for (int i = 0; i < 4; ++i) o.r[i] = make_root((float)r[i].re, (float)r[i].im, 1.0f);
    return o;
}

}  // namespace cr
