"""Host-side mirror of the reference's `utils.rs` helpers that callers of the hot path use to build instance
matrices and colours (src/utils.rs:67-225), plus the few `geometric_algebra::ppga2d` / `epga1d` operations that
`path.rs:387-815` needs. All arithmetic is binary32 like the reference's.

Conventions (closed forms of the sandwich product M X ~M derived symbolically from the basis
declaration Point = (e12, e01, -e02), Plane = (e0, e2, e1), Motor = (1, e12, e01, -e02) and pinned by
tests/test_path_constructors.py):
  * a Point is (w, w x, w y), a Plane used as a vector is (0, x, y), `dual()` is the component-wise identity;
  * a Motor is (m0, m1, m2, m3); `rotate2d(a)` = (cos a/2, sin a/2, 0, 0) rotates counter-clockwise (y up),
    `translate2d(v)` = (1, 0, -v_y / 2, v_x / 2).

The ppga3d camera motors of the demo (src/utils.rs:141-151,167-180) are out of scope; the
perspective matrix, matrix product and colour-space helpers are here because instance matrices / colours are inputs
of the hot path.
"""
from __future__ import annotations

import math
from typing import Sequence

import numpy as np

f32 = np.float32


def _v(values) -> np.ndarray:
    return np.asarray(values, dtype=np.float32)


# ------------------------------------------------------------------------------------------------- ppga2d
def vec_to_point(v) -> np.ndarray:
    """src/utils.rs:111-113"""
    return _v([1.0, v[0], v[1]])


def weighted_vec_to_point(w, v) -> np.ndarray:
    """src/utils.rs:116-118"""
    w = f32(w)
    return _v([w, f32(v[0]) * w, f32(v[1]) * w])


def point_to_vec(p) -> np.ndarray:
    """src/utils.rs:106-108"""
    return _v([p[1] / p[0], p[2] / p[0]])


def rotate_90_degree_clockwise(v) -> np.ndarray:
    """src/utils.rs:101-103"""
    return _v([0.0, v[2], -v[1]])


def rotate2d(angle: float) -> np.ndarray:
    """src/utils.rs:121-124"""
    half = f32(angle) * f32(0.5)
    return _v([np.cos(half), np.sin(half), 0.0, 0.0])


def translate2d(v) -> np.ndarray:
    """src/utils.rs:127-129"""
    return _v([1.0, 0.0, f32(-0.5) * f32(v[1]), f32(0.5) * f32(v[0])])


def motor_product(m, n) -> np.ndarray:
    """`m.geometric_product(n)` of two ppga2d Motors: apply n first, then m."""
    m, n = _v(m), _v(n)
    return _v([m[0] * n[0] - m[1] * n[1],
               m[0] * n[1] + m[1] * n[0],
               m[0] * n[2] - m[1] * n[3] + m[2] * n[0] + m[3] * n[1],
               m[0] * n[3] + m[1] * n[2] - m[2] * n[1] + m[3] * n[0]])


def motor_inverse(m) -> np.ndarray:
    """Inverse of a unit Motor = its reversal."""
    m = _v(m)
    return _v([m[0], -m[1], -m[2], -m[3]])


def motor_transform_point(m, p) -> np.ndarray:
    """`motor.transformation(point)` = M p ~M."""
    m, p = _v(m), _v(p)
    cc, ss, cs = m[0] * m[0], m[1] * m[1], m[0] * m[1]
    return _v([p[0] * (cc + ss),
               (cc - ss) * p[1] - f32(2) * cs * p[2] + f32(2) * p[0] * (m[0] * m[3] + m[1] * m[2]),
               (cc - ss) * p[2] + f32(2) * cs * p[1] - f32(2) * p[0] * (m[0] * m[2] - m[1] * m[3])])


def motor_transform_plane(m, a) -> np.ndarray:
    """`motor.transformation(plane)` = M a ~M; the vector part (a1, a2) rotates like a point's coordinates."""
    m, a = _v(m), _v(a)
    cc, ss, cs = m[0] * m[0], m[1] * m[1], m[0] * m[1]
    return _v([a[0] * (cc + ss) - f32(2) * a[1] * (m[0] * m[3] - m[1] * m[2]) + f32(2) * a[2] * (m[0] * m[2] + m[1] * m[3]),
               (cc - ss) * a[1] - f32(2) * cs * a[2],
               f32(2) * cs * a[1] + (cc - ss) * a[2]])


def rotation2d(motor) -> float:
    """src/utils.rs:132-134"""
    return float(f32(2) * np.arctan2(f32(motor[1]), f32(motor[0])))


def translation2d(motor) -> np.ndarray:
    """src/utils.rs:137-140: divide the rotor part out, read the translation."""
    m = _v(motor)
    inv = f32(1) / (m[0] * m[0] + m[1] * m[1])
    t = motor_product(m, _v([m[0] * inv, -m[1] * inv, 0.0, 0.0]))
    return _v([f32(2) * t[3], f32(-2) * t[2]])


def motor2d_to_mat3(motor) -> np.ndarray:
    """src/utils.rs:154-165: rows are the images of the x axis, the y axis and the origin, each as (x, y, w);
    `Path::transform` applies it as p' = row2 + p.x row0 + p.y row1 (src/path.rs:391-398)."""
    rows = []
    for index in (1, 2, 0):
        point = np.zeros(3, np.float32)
        point[index] = 1.0
        r = motor_transform_point(motor, point)
        rows.append([r[1], r[2], r[0]])
    return _v(rows)


def mat3_to_instance_mat4(mat3, scale: float = 1.0) -> np.ndarray:
    """Embeds the 2D transform into the instance mat4 the vertex stage multiplies positions with (column vectors
    `transform_row_0..3`, src/shaders.wgsl:13-27,66-74): clip = M (x, y, 0, 1)."""
    t = _v(mat3)
    m = np.zeros(16, np.float32)
    m[0], m[1] = t[0][0] * f32(scale), t[0][1] * f32(scale)
    m[4], m[5] = t[1][0] * f32(scale), t[1][1] * f32(scale)
    m[10] = 1.0
    m[12], m[13] = t[2][0], t[2][1]
    m[15] = 1.0
    return m


def line_line_intersection(a, b) -> np.ndarray:
    """src/utils.rs:67-70: `a.outer_product(b)` normalised to unit weight."""
    a, b = _v(a), _v(b)
    p = _v([a[2] * b[1] - a[1] * b[2], a[0] * b[2] - a[2] * b[0], a[1] * b[0] - a[0] * b[1]])
    return p * (f32(1) / p[0])


# ------------------------------------------------------------------------------------------------- epga1d
def complex_signum(z) -> np.ndarray:
    z = _v(z)
    return z * (f32(1) / np.sqrt(z[0] * z[0] + z[1] * z[1]))


def complex_mul(a, b) -> np.ndarray:
    a, b = _v(a), _v(b)
    return _v([a[0] * b[0] - a[1] * b[1], a[0] * b[1] + a[1] * b[0]])


def complex_div(a, b) -> np.ndarray:
    a, b = _v(a), _v(b)
    inv = f32(1) / (b[0] * b[0] + b[1] * b[1])
    return _v([(a[0] * b[0] + a[1] * b[1]) * inv, (a[1] * b[0] - a[0] * b[1]) * inv])


def complex_arg(z) -> np.float32:
    return np.arctan2(f32(z[1]), f32(z[0]))


def complex_powf(z, exponent) -> np.ndarray:
    z = _v(z)
    mag = np.power(np.sqrt(z[0] * z[0] + z[1] * z[1]), f32(exponent))
    ang = complex_arg(z) * f32(exponent)
    return _v([mag * np.cos(ang), mag * np.sin(ang)])


def complex_powi(z, n: int) -> np.ndarray:
    """Exponentiation by squaring."""
    x, y = _v(z), _v([1.0, 0.0])
    if n == 0:
        return y
    while n > 1:
        if n & 1:
            y = complex_mul(x, y)
        x = complex_mul(x, x)
        n >>= 1
    return complex_mul(x, y)


# ------------------------------------------------------------------------------------------------- matrices, colours
def perspective_projection(field_of_view_y: float, aspect_ratio: float, near: float, far: float) -> np.ndarray:
    """src/utils.rs:183-192 (four column vectors)."""
    height = f32(1) / np.tan(f32(field_of_view_y) * f32(0.5))
    denominator = f32(1) / (f32(near) - f32(far))
    return _v([[height / f32(aspect_ratio), 0, 0, 0], [0, height, 0, 0], [0, 0, -f32(far) * denominator, 1],
               [0, 0, f32(near) * f32(far) * denominator, 0]])


def matrix_multiplication(a, b) -> np.ndarray:
    """src/utils.rs:195-202 (column vectors: result column j = sum_k a column k * b[j][k])."""
    a, b = _v(a), _v(b)
    return _v([a[0] * b[j][0] + a[1] * b[j][1] + a[2] * b[j][2] + a[3] * b[j][3] for j in range(4)])


def srgb_to_linear(color: Sequence[float]) -> np.ndarray:
    """src/utils.rs:205-214"""
    c = _v(color).copy()
    for i in range(3):
        c[i] = np.power((c[i] + f32(0.055)) / f32(1.055), f32(2.4)) if c[i] > f32(0.04045) else c[i] / f32(12.92)
    return c


def linear_to_srgb(color: Sequence[float]) -> np.ndarray:
    """src/utils.rs:217-226"""
    c = _v(color).copy()
    for i in range(3):
        c[i] = f32(1.055) * np.power(c[i], f32(1.0 / 2.4)) - f32(0.055) if c[i] > f32(0.0031308) else f32(12.92) * c[i]
    return c


TAU = f32(2.0 * math.pi)
