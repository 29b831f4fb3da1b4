"""Vec3 / Quat / Mat4 with the reference's conventions.

Mirrors `engine/src/math.ts` of the reference (column-major mat4, ``v' = M v``,
left-handed, f64 arithmetic, f32 storage in ``Mat4.values``) so that host-side
pose evaluation produces the same palette input as `model.ts`.  Names follow the
reference API (`Quat.slerp`, `Quat.fromEuler`, `Mat4.fromQuat`, ...).

Reference: math.ts:2-4 (easeInOut), 56-232 (Quat), 234-546 (Mat4).
"""
from __future__ import annotations

import math
import numpy as np

__all__ = ["easeInOut", "Vec3", "Quat", "Mat4"]


def easeInOut(t: float) -> float:
    """Quadratic ease (math.ts:2-4)."""
    if t < 0.5:
        return 2.0 * t * t
    u = -2.0 * t + 2.0
    return 1.0 - (u * u) / 2.0


class Vec3:
    __slots__ = ("x", "y", "z")

    def __init__(self, x: float, y: float, z: float):
        self.x, self.y, self.z = float(x), float(y), float(z)

    def add(self, o: "Vec3") -> "Vec3":
        return Vec3(self.x + o.x, self.y + o.y, self.z + o.z)

    def subtract(self, o: "Vec3") -> "Vec3":
        return Vec3(self.x - o.x, self.y - o.y, self.z - o.z)

    def length(self) -> float:
        return math.sqrt(self.x * self.x + self.y * self.y + self.z * self.z)

    def normalize(self) -> "Vec3":
        n = self.length()
        if n == 0:
            return Vec3(0, 0, 0)
        return Vec3(self.x / n, self.y / n, self.z / n)

    def cross(self, o: "Vec3") -> "Vec3":
        return Vec3(self.y * o.z - self.z * o.y, self.z * o.x - self.x * o.z, self.x * o.y - self.y * o.x)

    def dot(self, o: "Vec3") -> float:
        return self.x * o.x + self.y * o.y + self.z * o.z

    def scale(self, s: float) -> "Vec3":
        return Vec3(self.x * s, self.y * s, self.z * s)

    def clone(self) -> "Vec3":
        return Vec3(self.x, self.y, self.z)

    def __repr__(self):
        return f"Vec3({self.x}, {self.y}, {self.z})"


class Quat:
    """xyzw quaternion, f64 components (math.ts:56-232)."""

    __slots__ = ("x", "y", "z", "w")

    def __init__(self, x: float, y: float, z: float, w: float):
        self.x, self.y, self.z, self.w = float(x), float(y), float(z), float(w)

    def clone(self) -> "Quat":
        return Quat(self.x, self.y, self.z, self.w)

    def add(self, o: "Quat") -> "Quat":
        return Quat(self.x + o.x, self.y + o.y, self.z + o.z, self.w + o.w)

    def multiply(self, o: "Quat") -> "Quat":
        a, b = self, o
        return Quat(
            a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
            a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x,
            a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w,
            a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z,
        )

    def conjugate(self) -> "Quat":
        return Quat(-self.x, -self.y, -self.z, self.w)

    def length(self) -> float:
        return math.sqrt(self.x * self.x + self.y * self.y + self.z * self.z + self.w * self.w)

    def normalize(self) -> "Quat":
        n = self.length()
        if n == 0:
            return Quat(0, 0, 0, 1)
        return Quat(self.x / n, self.y / n, self.z / n, self.w / n)

    def rotateVec(self, v: Vec3) -> Vec3:
        tx = 2 * (self.y * v.z - self.z * v.y)
        ty = 2 * (self.z * v.x - self.x * v.z)
        tz = 2 * (self.x * v.y - self.y * v.x)
        return Vec3(
            v.x + self.w * tx + (self.y * tz - self.z * ty),
            v.y + self.w * ty + (self.z * tx - self.x * tz),
            v.z + self.w * tz + (self.x * ty - self.y * tx),
        )

    def toArray(self):
        return [self.x, self.y, self.z, self.w]

    @staticmethod
    def slerp(a: "Quat", b: "Quat", t: float) -> "Quat":
        """Shortest-arc slerp with the 0.9995 lerp fallback (math.ts:156-189)."""
        cos = a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w
        bx, by, bz, bw = b.x, b.y, b.z, b.w
        if cos < 0:
            cos, bx, by, bz, bw = -cos, -bx, -by, -bz, -bw
        if cos > 0.9995:
            x = a.x + t * (bx - a.x)
            y = a.y + t * (by - a.y)
            z = a.z + t * (bz - a.z)
            w = a.w + t * (bw - a.w)
            inv = 1.0 / math.hypot(x, y, z, w)
            return Quat(x * inv, y * inv, z * inv, w * inv)
        theta0 = math.acos(cos)
        s = math.sin(theta0)
        theta = theta0 * t
        s0 = math.sin(theta0 - theta) / s
        s1 = math.sin(theta) / s
        return Quat(s0 * a.x + s1 * bx, s0 * a.y + s1 * by, s0 * a.z + s1 * bz, s0 * a.w + s1 * bw)

    @staticmethod
    def fromEuler(rotX: float, rotY: float, rotZ: float) -> "Quat":
        """ZXY order, left-handed (math.ts:192-206)."""
        cx, sx = math.cos(rotX * 0.5), math.sin(rotX * 0.5)
        cy, sy = math.cos(rotY * 0.5), math.sin(rotY * 0.5)
        cz, sz = math.cos(rotZ * 0.5), math.sin(rotZ * 0.5)
        w = cy * cx * cz + sy * sx * sz
        x = cy * sx * cz + sy * cx * sz
        y = sy * cx * cz - cy * sx * sz
        z = cy * cx * sz - sy * sx * cz
        return Quat(x, y, z, w).normalize()

    def __repr__(self):
        return f"Quat({self.x}, {self.y}, {self.z}, {self.w})"


class Mat4:
    """Column-major 4x4, ``values`` is a float32[16] (math.ts:234-546).

    Arithmetic is carried out in f64 on the f32-stored operands and rounded to
    f32 when written back, which is what a JS ``Float32Array`` store does.
    """

    __slots__ = ("values",)

    def __init__(self, values):
        self.values = np.asarray(values, dtype=np.float32).reshape(16)

    @staticmethod
    def identity() -> "Mat4":
        return Mat4(np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1], dtype=np.float32))

    def clone(self) -> "Mat4":
        return Mat4(self.values.copy())

    def setIdentity(self) -> "Mat4":
        self.values[:] = (1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1)
        return self

    def translateInPlace(self, tx: float, ty: float, tz: float) -> "Mat4":
        v = self.values
        v[12] = np.float32(float(v[12]) + tx)
        v[13] = np.float32(float(v[13]) + ty)
        v[14] = np.float32(float(v[14]) + tz)
        return self

    def multiply(self, other: "Mat4") -> "Mat4":
        out = np.empty(16, dtype=np.float32)
        Mat4.multiplyArrays(self.values, 0, other.values, 0, out, 0)
        return Mat4(out)

    @staticmethod
    def multiplyArrays(a, aOffset: int, b, bOffset: int, out, outOffset: int) -> None:
        """out = a * b (column-major), f64 accumulate in the reference's term order
        (math.ts:324-346), rounded to f32 on store."""
        A = a[aOffset:aOffset + 16].astype(np.float64)
        Bm = b[bOffset:bOffset + 16].astype(np.float64)
        res = np.empty(16, dtype=np.float64)
        for c in range(4):
            b0, b1, b2, b3 = Bm[c * 4], Bm[c * 4 + 1], Bm[c * 4 + 2], Bm[c * 4 + 3]
            for r in range(4):
                res[c * 4 + r] = A[r] * b0 + A[4 + r] * b1 + A[8 + r] * b2 + A[12 + r] * b3
        out[outOffset:outOffset + 16] = res.astype(np.float32)

    @staticmethod
    def fromQuat(x: float, y: float, z: float, w: float) -> "Mat4":
        """math.ts:352-384."""
        x2, y2, z2 = x + x, y + y, z + z
        xx, xy, xz = x * x2, x * y2, x * z2
        yy, yz, zz = y * y2, y * z2, z * z2
        wx, wy, wz = w * x2, w * y2, w * z2
        return Mat4(np.array([
            1 - (yy + zz), xy + wz, xz - wy, 0,
            xy - wz, 1 - (xx + zz), yz + wx, 0,
            xz + wy, yz - wx, 1 - (xx + yy), 0,
            0, 0, 0, 1], dtype=np.float64).astype(np.float32))

    @staticmethod
    def fromPositionRotation(position: Vec3, rotation: Quat) -> "Mat4":
        m = Mat4.fromQuat(rotation.x, rotation.y, rotation.z, rotation.w)
        m.values[12], m.values[13], m.values[14] = position.x, position.y, position.z
        return m

    def getPosition(self) -> Vec3:
        return Vec3(self.values[12], self.values[13], self.values[14])

    def toQuat(self) -> Quat:
        return Mat4.toQuatFromArray(self.values, 0)

    @staticmethod
    def toQuatFromArray(m, offset: int) -> Quat:
        """math.ts:406-448 (Shepperd-style branch on the trace)."""
        g = lambda i: float(m[offset + i])
        m00, m01, m02 = g(0), g(4), g(8)
        m10, m11, m12 = g(1), g(5), g(9)
        m20, m21, m22 = g(2), g(6), g(10)
        trace = m00 + m11 + m22
        if trace > 0:
            s = math.sqrt(trace + 1.0) * 2
            w, x, y, z = 0.25 * s, (m21 - m12) / s, (m02 - m20) / s, (m10 - m01) / s
        elif m00 > m11 and m00 > m22:
            s = math.sqrt(1.0 + m00 - m11 - m22) * 2
            w, x, y, z = (m21 - m12) / s, 0.25 * s, (m01 + m10) / s, (m02 + m20) / s
        elif m11 > m22:
            s = math.sqrt(1.0 + m11 - m00 - m22) * 2
            w, x, y, z = (m02 - m20) / s, (m01 + m10) / s, 0.25 * s, (m12 + m21) / s
        else:
            s = math.sqrt(1.0 + m22 - m00 - m11) * 2
            w, x, y, z = (m10 - m01) / s, (m02 + m20) / s, (m12 + m21) / s, 0.25 * s
        inv = 1.0 / math.hypot(x, y, z, w)
        return Quat(x * inv, y * inv, z * inv, w * inv)
