"""Batched host-side pose evaluation for crowds: P poses of one skeleton at once.

Vectorised restatement of `Model.computeWorldMatrices` (model.ts:330-420) and of the
tween rule `slerp(start, target, easeInOut(t))` (model.ts:158-194, math.ts:2-4,156-189)
over a leading batch axis, with f32 rounding at exactly the points where the reference
stores into a Float32Array (local rotations, every Mat4), so one row of the batch is
bit-identical to what `Model.evaluatePose()` leaves in `getBoneWorldMatrices()`.
Feeds `rz_set_palettes` for K-instance crowds ("staggered VMD phase", BASELINE configs 2/4).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from .model import Bone


def ease_in_out(t: np.ndarray) -> np.ndarray:
    t = np.asarray(t, dtype=np.float64)
    u = -2.0 * t + 2.0
    return np.where(t < 0.5, 2.0 * t * t, 1.0 - (u * u) / 2.0)


def slerp_batch(a: np.ndarray, b: np.ndarray, t: np.ndarray) -> np.ndarray:
    """Quat.slerp (math.ts:156-189) on [...,4] f64 arrays; t broadcastable to [...]."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    shp = np.broadcast_shapes(a.shape[:-1], b.shape[:-1], np.shape(t))
    a = np.broadcast_to(a, shp + (4,))
    b = np.broadcast_to(b, shp + (4,))
    t = np.broadcast_to(np.asarray(t, np.float64), shp)[..., None]
    cos = np.sum(a * b, axis=-1, keepdims=True)
    neg = cos < 0
    cos = np.where(neg, -cos, cos)
    b = np.where(neg, -b, b)
    lin = a + t * (b - a)
    lin = lin * (1.0 / np.sqrt(np.sum(lin * lin, axis=-1, keepdims=True)))   # Math.hypot
    cc = np.clip(cos, -1.0, 1.0)
    th0 = np.arccos(cc)
    s = np.sin(th0)
    s = np.where(s == 0, 1.0, s)
    th = th0 * t
    sph = (np.sin(th0 - th) / s) * a + (np.sin(th) / s) * b
    return np.where(cos > 0.9995, lin, sph)


def _from_quat(q: np.ndarray) -> np.ndarray:
    """Mat4.fromQuat (math.ts:352-384) -> [...,16] float32 (column-major)."""
    x, y, z, w = (q[..., i].astype(np.float64) for i in range(4))
    x2, y2, z2 = x + x, y + y, z + z
    xx, xy, xz = x * x2, x * y2, x * z2
    yy, yz, zz = y * y2, y * z2, z * z2
    wx, wy, wz = w * x2, w * y2, w * z2
    o = np.zeros(q.shape[:-1] + (16,), dtype=np.float64)
    o[..., 0] = 1 - (yy + zz); o[..., 1] = xy + wz; o[..., 2] = xz - wy
    o[..., 4] = xy - wz; o[..., 5] = 1 - (xx + zz); o[..., 6] = yz + wx
    o[..., 8] = xz + wy; o[..., 9] = yz - wx; o[..., 10] = 1 - (xx + yy)
    o[..., 15] = 1
    return o.astype(np.float32)


def _mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Mat4.multiply on [...,16] f32 operands: f64 accumulate in the reference's term order, f32 store."""
    A = a.astype(np.float64)
    Bm = b.astype(np.float64)
    out = np.empty(np.broadcast_shapes(a.shape, b.shape), dtype=np.float64)
    for c in range(4):
        b0, b1, b2, b3 = (Bm[..., c * 4 + i] for i in range(4))
        for r in range(4):
            out[..., c * 4 + r] = A[..., r] * b0 + A[..., 4 + r] * b1 + A[..., 8 + r] * b2 + A[..., 12 + r] * b3
    return out.astype(np.float32)


def parent_first_order(bones: Sequence[Bone]) -> List[int]:
    n = len(bones)
    done = [False] * n
    order: List[int] = []
    for i in range(n):
        chain = []
        j = i
        while not done[j]:
            chain.append(j)
            done[j] = True
            p = bones[j].parentIndex
            if p < 0 or p >= n:
                break
            j = p
        order.extend(reversed(chain))
    return order


def world_matrices_batch(bones: Sequence[Bone], localRot: np.ndarray) -> np.ndarray:
    """localRot: [P,B,4] (stored as f32 like SkeletonRuntime.localRotations) -> world [P,B,16] f32."""
    lr = np.asarray(localRot, dtype=np.float32)
    P, B, _ = lr.shape
    world = np.zeros((P, B, 16), dtype=np.float32)
    ident = np.zeros(16, np.float32)
    ident[[0, 5, 10, 15]] = 1
    for i in parent_first_order(bones):
        b = bones[i]
        rotM = _from_quat(lr[:, i])
        ap = b.appendParentIndex
        if b.appendRotate and ap is not None and 0 <= ap < B:
            ratio = 1.0 if b.appendRatio is None else max(-1.0, min(1.0, b.appendRatio))
            if abs(ratio) > 1e-6:
                q = lr[:, ap].astype(np.float64).copy()
                if ratio < 0:
                    q[:, :3] = -q[:, :3]
                idq = np.broadcast_to(np.array([0.0, 0.0, 0.0, 1.0]), q.shape)
                r = slerp_batch(idq, q, abs(ratio))
                rotM = _mul(_from_quat(r), rotM)
        m1 = ident.copy()
        m1[12:15] = np.asarray(b.bindTranslation, np.float64).astype(np.float32)   # identity + translateInPlace
        local = _mul(_mul(m1[None, :], rotM), ident[None, :])
        if b.parentIndex >= 0:
            world[:, i] = _mul(world[:, b.parentIndex], local)
        else:
            world[:, i] = local
    return world


def tween_pose_batch(qa: np.ndarray, qb: np.ndarray, phase: np.ndarray) -> np.ndarray:
    """Local rotations of P instances between two keys: slerp(qa, qb, easeInOut(phase_p)), stored f32.
    qa, qb: [B,4]; phase: [P] in [0,1]  ->  [P,B,4] float32 (what updateRotationTweens writes)."""
    e = ease_in_out(np.clip(np.asarray(phase, np.float64), 0.0, 1.0))
    qa32 = np.asarray(qa, np.float32).astype(np.float64)   # tween state is Float32Array
    qb32 = np.asarray(qb, np.float32).astype(np.float64)
    q = slerp_batch(qa32[None, :, :], qb32[None, :, :], e[:, None])
    return q.astype(np.float32)
