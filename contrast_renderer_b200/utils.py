"""Host-side mirror of the reference's `utils.rs` helpers that callers of the hot path use to build instance
matrices and colours (src/utils.rs:67-225), plus the few `geometric_algebra::ppga2d` / `epga1d` operations that
`path.rs:387-815` needs. All arithmetic is binary32 like the reference's.

Conventions (closed forms of the sandwich product M X ~M derived symbolically from the basis
declaration Point = (e12, e01, -e02), Plane = (e0, e2, e1), Motor = (1, e12, e01, -e02) and pinned by
tests/test_path_constructors.py):
  * a Point is (w, w x, w y), a Plane used as a vector is (0, x, y), `dual()` is the component-wise identity;
  * a Motor is (m0, m1, m2, m3); `rotate2d(a)` = (cos a/2, sin a/2, 0, 0) rotates counter-clockwise (y up),
    `translate2d(v)` = (1, 0, -v_y / 2, v_x / 2).

The ppga3d camera motors of the demo (src/utils.rs:141-151,167-180: `rotate_around_axis`, `motor2d_to_motor3d`,
`motor3d_to_mat4`, plus the `Translator` / `Motor` products examples/showcase/main.rs:163-201 builds its instance matrices
with) are evaluated with a small generic geometric-product routine over the basis
Motor = (1, e32, e13, e21, e0123, e01, e02, e03), Point = (e123, -e023, e013, -e012). The crate that defines the basis is
not in the reference tree; the signs of e21, e01, e02 are FORCED by requiring `motor3d_to_mat4(motor2d_to_motor3d(m))` to
move (x, y, 0, 1) exactly like `motor2d_to_mat3(m)` moves (x, y, 1) (tests/test_path_constructors.py), the remaining ones
(e32, e13, e03, e0123) are the cyclic continuation. `save_png` writes a frame read back from the renderer.
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


# ------------------------------------------------------------------------------------------------- ppga3d (demo camera)
_METRIC = (0.0, 1.0, 1.0, 1.0)   # e0 is null


def _blade_product(a: int, b: int):
    """Geometric product of two basis blades given as bit masks over (e0, e1, e2, e3): (sign, blade)."""
    sign = 1.0
    for i in range(4):
        if (b >> i) & 1:
            if bin(a >> (i + 1)).count("1") & 1:
                sign = -sign
            if (a >> i) & 1:
                sign *= _METRIC[i]
                a ^= 1 << i
            else:
                a |= 1 << i
    return sign, a


def _blade(indices, sign=1.0):
    mask = 0
    for i in indices:
        mask |= 1 << i
    for i in range(len(indices)):
        for j in range(i + 1, len(indices)):
            if indices[i] > indices[j]:
                sign = -sign
    return mask, sign


_MOTOR3 = [_blade(()), _blade((3, 2)), _blade((1, 3)), _blade((2, 1)), _blade((0, 1, 2, 3)), _blade((0, 1)), _blade((0, 2)), _blade((0, 3))]
_POINT3 = [_blade((1, 2, 3)), _blade((0, 2, 3), -1.0), _blade((0, 1, 3)), _blade((0, 1, 2), -1.0)]


def _to_multivector(basis, values):
    out = {}
    for (mask, sign), v in zip(basis, values):
        out[mask] = out.get(mask, 0.0) + sign * float(v)
    return out


def _from_multivector(basis, x) -> np.ndarray:
    return _v([x.get(mask, 0.0) * sign for mask, sign in basis])


def _gp(x, y):
    out = {}
    for a, va in x.items():
        for b, vb in y.items():
            sign, c = _blade_product(a, b)
            if sign:
                out[c] = out.get(c, 0.0) + sign * va * vb
    return out


def _reverse(x):
    return {a: v * (-1.0) ** (bin(a).count("1") * (bin(a).count("1") - 1) // 2) for a, v in x.items()}


def rotate_around_axis(angle: float, axis) -> np.ndarray:
    """src/utils.rs:143-146: ppga3d::Rotor (cos a/2, axis sin a/2)."""
    sinus = math.sin(angle * 0.5)
    return _v([math.cos(angle * 0.5), axis[0] * sinus, axis[1] * sinus, axis[2] * sinus])


def translator3d(x: float, y: float, z: float) -> np.ndarray:
    """ppga3d::Translator::new(1, x, y, z) as a Motor (examples/showcase/main.rs:171 passes -view_distance / 2 as z)."""
    return _v([1.0, 0.0, 0.0, 0.0, 0.0, x, y, z])


def rotor3d_to_motor3d(rotor) -> np.ndarray:
    r = _v(rotor)
    return _v([r[0], r[1], r[2], r[3], 0.0, 0.0, 0.0, 0.0])


def motor3d_product(a, b) -> np.ndarray:
    """`a.geometric_product(b)` of two ppga3d motors (8 components each)."""
    return _from_multivector(_MOTOR3, _gp(_to_multivector(_MOTOR3, a), _to_multivector(_MOTOR3, b)))


def motor2d_to_motor3d(motor) -> np.ndarray:
    """src/utils.rs:149-151"""
    m = _v(motor)
    return _v([m[0], 0.0, 0.0, m[1], 0.0, -m[3], m[2], 0.0])


def motor3d_transform_point(motor, point) -> np.ndarray:
    """`motor.transformation(point)`: M P ~M on a ppga3d::Point (w, x, y, z)."""
    m = _to_multivector(_MOTOR3, motor)
    return _from_multivector(_POINT3, _gp(_gp(m, _to_multivector(_POINT3, point)), _reverse(m)))


def motor3d_to_mat4(motor) -> np.ndarray:
    """src/utils.rs:168-180: four column vectors (images of the x, y, z axes and of the origin), each as (x, y, z, w)."""
    rows = []
    for index in (1, 2, 3, 0):
        point = np.zeros(4, np.float32)
        point[index] = 1.0
        r = motor3d_transform_point(motor, point)
        rows.append([r[1], r[2], r[3], r[0]])
    return _v(rows)


# ------------------------------------------------------------------------------------------------- frame dump
def save_png(path: str, rgba) -> None:
    """Writes an [H, W, 4] frame as an 8-bit RGBA PNG. Float input is taken as the renderer's premultiplied linear colour
    (`Renderer.read_color()`): un-premultiplied, converted with `linear_to_srgb`'s curve and quantised; uint8 input is
    written as it is (`read_color_texels()` of an Rgba8Unorm target)."""
    import struct
    import zlib
    a = np.asarray(rgba)
    if a.ndim != 3 or a.shape[2] != 4:
        raise ValueError("save_png expects an [H, W, 4] array")
    if a.dtype != np.uint8:
        c = a.astype(np.float32)
        alpha = c[..., 3:4]
        rgb = np.where(alpha > 0, c[..., :3] / np.where(alpha > 0, alpha, 1), 0).clip(0, 1)
        rgb = np.where(rgb > 0.0031308, 1.055 * np.power(rgb, 1.0 / 2.4) - 0.055, 12.92 * rgb)
        a = (np.concatenate([rgb, alpha.clip(0, 1)], axis=2) * 255.0 + 0.5).astype(np.uint8)
    h, w = a.shape[:2]
    raw = b"".join(b"\x00" + a[y].tobytes() for y in range(h))

    def chunk(tag: bytes, data: bytes) -> bytes:
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 6, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


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
