"""SE3 as Sophus::SE3d stores it — unit quaternion (x, y, z, w) + translation — with the
operations the reference's driver uses around update():

    SE3d(Quaterniond(w,x,y,z), Vector3d)          dense_mapping/test_monocular_mapping.cpp:333-335
    pose_curr_TWC.inverse() * pose_ref_TWC        :289-290

Arithmetic order follows Sophus @61f9a98 / Eigen (see oracle/dense_mono_oracle.cpp for the
C++ restatement the tests compare against).  Python floats are IEEE doubles, so results agree
with the oracle bit for bit.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Sequence


def _cross(a, b):
    return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])


def _qnormalized(q):
    x, y, z, w = q
    n = math.sqrt((x * x + z * z) + (y * y + w * w))
    return (x / n, y / n, z / n, w / n)


def _rotate(q, v):
    qv = (q[0], q[1], q[2])
    uv = _cross(qv, v)
    uv = (uv[0] + uv[0], uv[1] + uv[1], uv[2] + uv[2])
    c = _cross(qv, uv)
    w = q[3]
    return (v[0] + uv[0] * w + c[0], v[1] + uv[1] * w + c[1], v[2] + uv[2] * w + c[2])


@dataclass(frozen=True)
class SE3:
    q: tuple  # (x, y, z, w)
    t: tuple  # (tx, ty, tz)

    @staticmethod
    def from_quat_trans(qx, qy, qz, qw, tx, ty, tz) -> "SE3":
        """SE3d(Quaterniond(qw,qx,qy,qz), Vector3d(tx,ty,tz)) — the constructor normalises q."""
        return SE3(_qnormalized((float(qx), float(qy), float(qz), float(qw))), (float(tx), float(ty), float(tz)))

    @staticmethod
    def raw(q: Sequence[float], t: Sequence[float]) -> "SE3":
        return SE3(tuple(float(v) for v in q), tuple(float(v) for v in t))

    @staticmethod
    def identity() -> "SE3":
        return SE3((0.0, 0.0, 0.0, 1.0), (0.0, 0.0, 0.0))

    def inverse(self) -> "SE3":
        c = _qnormalized((-self.q[0], -self.q[1], -self.q[2], self.q[3]))
        nt = (self.t[0] * -1.0, self.t[1] * -1.0, self.t[2] * -1.0)
        return SE3(c, _rotate(c, nt))

    def __mul__(self, other):
        if isinstance(other, SE3):
            a, b = self.q, other.q
            ax, ay, az, aw = a
            bx, by, bz, bw = b
            q = (aw * bx + ax * bw + ay * bz - az * by,
                 aw * by + ay * bw + az * bx - ax * bz,
                 aw * bz + az * bw + ax * by - ay * bx,
                 aw * bw - ax * bx - ay * by - az * bz)
            r = _rotate(self.q, other.t)
            return SE3(_qnormalized(q), (self.t[0] + r[0], self.t[1] + r[1], self.t[2] + r[2]))
        v = tuple(float(x) for x in other)
        r = _rotate(self.q, v)
        return (r[0] + self.t[0], r[1] + self.t[1], r[2] + self.t[2])

    def unit_quaternion(self):
        return self.q

    def translation(self):
        return self.t


def relative_pose(T_WC_ref: SE3, T_WC_curr: SE3) -> SE3:
    """T_C_R = T_WC(curr)^-1 * T_WC(ref)  (ref:289-290)."""
    return T_WC_curr.inverse() * T_WC_ref
