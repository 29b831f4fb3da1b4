"""Model: static mesh/skeleton tables + the CPU skeleton runtime that produces the
per-frame bone-matrix palette (the *input* of the deform stage).

Mirrors `engine/src/model.ts` of the reference: same getters
(`getVertices/getSkinning/getSkeleton/getBoneWorldMatrices/...`), same
`rotateBones(names, quats, durationMs)` tween semantics (model.ts:246-315),
`evaluatePose()` = `updateRotationTweens` (model.ts:158-194) +
`computeWorldMatrices` (model.ts:330-420).  The reference reads
`performance.now()`; here the clock is injected (`clock: () -> ms`) so that pose
evaluation is reproducible.

New (not in the reference): vertex-morph tables and SDEF records are
kept instead of being skipped, see `VertexMorphs` / `SdefTable`.
"""
from __future__ import annotations

import math
import time
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

from .math3d import Mat4, Quat, easeInOut

VERTEX_STRIDE = 8


@dataclass
class Bone:
    name: str
    parentIndex: int
    bindTranslation: List[float]  # f64 triple (pmx-loader.ts:423)
    children: List[int] = field(default_factory=list)
    appendParentIndex: Optional[int] = None
    appendRatio: Optional[float] = None
    appendRotate: bool = False
    appendMove: bool = False


@dataclass
class Skeleton:
    bones: List[Bone]
    inverseBindMatrices: np.ndarray  # float32[16*B], column-major


@dataclass
class Skinning:
    joints: np.ndarray   # uint16[4V]
    weights: np.ndarray  # uint8[4V], sums to 255 per vertex


@dataclass
class VertexMorphs:
    """Vertex morphs in PMX (morph-major) order, group morphs already expanded.

    offsets[m]..offsets[m+1] index into vertexIndex / delta (CSR by morph).
    """
    names: List[str]
    offsets: np.ndarray      # uint32[M+1]
    vertexIndex: np.ndarray  # uint32[nnz]
    delta: np.ndarray        # float32[nnz,3]

    @property
    def count(self) -> int:
        return len(self.offsets) - 1

    @staticmethod
    def empty() -> "VertexMorphs":
        return VertexMorphs([], np.zeros(1, np.uint32), np.zeros(0, np.uint32), np.zeros((0, 3), np.float32))


@dataclass
class SdefTable:
    vertexIndex: np.ndarray  # uint32[n]
    c_r0_r1: np.ndarray      # float32[n,9]
    w0: np.ndarray           # float32[n] unquantised BDEF2 weight of bone0 (informational)

    @staticmethod
    def empty() -> "SdefTable":
        return SdefTable(np.zeros(0, np.uint32), np.zeros((0, 9), np.float32), np.zeros(0, np.float32))


def _default_clock() -> float:
    return time.perf_counter() * 1000.0


class Model:
    def __init__(self, vertexData: np.ndarray, indexData: np.ndarray, textures, materials,
                 skeleton: Skeleton, skinning: Skinning, rigidbodies=None, joints=None,
                 morphs: Optional[VertexMorphs] = None, sdef: Optional[SdefTable] = None,
                 clock: Optional[Callable[[], float]] = None):
        self.vertexData = np.ascontiguousarray(vertexData, dtype=np.float32)
        self.vertexCount = self.vertexData.size // VERTEX_STRIDE
        self.indexData = np.ascontiguousarray(indexData, dtype=np.uint32)
        self.textures = textures or []
        self.materials = materials or []
        self.skeleton = skeleton
        self.skinning = skinning
        self.rigidbodies = rigidbodies or []
        self.joints = joints or []
        self.morphs = morphs or VertexMorphs.empty()
        self.sdef = sdef or SdefTable.empty()
        self.clock = clock or _default_clock
        if len(self.skeleton.bones) == 0:
            raise ValueError("Model has no bones")  # model.ts:113-115
        n = len(self.skeleton.bones)
        # runtime skeleton (model.ts:121-145)
        self.nameIndex: Dict[str, int] = {}
        for i, b in enumerate(self.skeleton.bones):
            self.nameIndex[b.name] = i  # later duplicates win, as in the reduce()
        self.localRotations = np.zeros(n * 4, dtype=np.float32)
        self.localRotations[3::4] = 1.0
        self.localTranslations = np.zeros(n * 3, dtype=np.float32)
        self.worldMatrices = np.zeros(n * 16, dtype=np.float32)
        # tween state (model.ts:147-156); Float32Array storage is kept on purpose
        self._active = np.zeros(n, dtype=np.uint8)
        self._startQuat = np.zeros(n * 4, dtype=np.float32)
        self._targetQuat = np.zeros(n * 4, dtype=np.float32)
        self._startTimeMs = np.zeros(n, dtype=np.float32)
        self._durationMs = np.zeros(n, dtype=np.float32)
        self._order = self._parent_first_order()

    # ---- getters (model.ts:196-244) -------------------------------------------------
    def getVertices(self): return self.vertexData
    def getTextures(self): return self.textures
    def getMaterials(self): return self.materials
    def getVertexCount(self): return self.vertexCount
    def getIndices(self): return self.indexData
    def getSkeleton(self): return self.skeleton
    def getSkinning(self): return self.skinning
    def getRigidbodies(self): return self.rigidbodies
    def getJoints(self): return self.joints
    def getBoneNames(self): return [b.name for b in self.skeleton.bones]
    def getBoneWorldMatrices(self): return self.worldMatrices
    def getBoneInverseBindMatrices(self): return self.skeleton.inverseBindMatrices
    def getMorphs(self): return self.morphs
    def getSdef(self): return self.sdef

    # ---- tweens ------------------------------------------------------------------------
    def _tween_value(self, idx: int, now: float) -> Quat:
        qi = idx * 4
        startMs = float(self._startTimeMs[idx])
        dur = max(1.0, float(self._durationMs[idx]))
        t = max(0.0, min(1.0, (now - startMs) / dur))
        e = easeInOut(t)
        s = Quat(*[float(v) for v in self._startQuat[qi:qi + 4]])
        g = Quat(*[float(v) for v in self._targetQuat[qi:qi + 4]])
        return Quat.slerp(s, g, e), t

    def rotateBones(self, names: Sequence[str], quats: Sequence[Quat], durationMs: Optional[float] = None) -> None:
        """model.ts:246-315."""
        normalized = [q.normalize() for q in quats]
        now = self.clock()
        dur = durationMs if (durationMs and durationMs > 0) else 0
        rot = self.localRotations
        for i, name in enumerate(names):
            idx = self.nameIndex.get(name, -1)
            if idx < 0 or idx >= len(self.skeleton.bones):
                continue
            qi = idx * 4
            tx, ty, tz, tw = normalized[i].toArray()
            if dur == 0:
                rot[qi:qi + 4] = (tx, ty, tz, tw)
                self._active[idx] = 0
                continue
            sx, sy, sz, sw = (float(v) for v in rot[qi:qi + 4])
            if self._active[idx] == 1:
                cur, _ = self._tween_value(idx, now)
                sx, sy, sz, sw = cur.x, cur.y, cur.z, cur.w
            self._startQuat[qi:qi + 4] = (sx, sy, sz, sw)
            self._targetQuat[qi:qi + 4] = (tx, ty, tz, tw)
            self._startTimeMs[idx] = now
            self._durationMs[idx] = dur
            self._active[idx] = 1

    def updateRotationTweens(self) -> None:
        """model.ts:158-194."""
        now = self.clock()
        for i in np.nonzero(self._active == 1)[0]:
            q, t = self._tween_value(int(i), now)
            self.localRotations[i * 4:i * 4 + 4] = (q.x, q.y, q.z, q.w)
            if t >= 1:
                self._active[i] = 0

    def evaluatePose(self) -> None:
        self.updateRotationTweens()
        self.computeWorldMatrices()

    # ---- hierarchy ---------------------------------------------------------------------
    def _parent_first_order(self) -> List[int]:
        """Order in which the reference's memoised recursion (model.ts:340-419) finishes bones."""
        bones = self.skeleton.bones
        n = len(bones)
        done = [False] * n
        order: List[int] = []
        for i in range(n):
            chain = []
            j = i
            guard = 0
            while not done[j]:
                chain.append(j)
                done[j] = True
                p = bones[j].parentIndex
                if p < 0 or p >= n:   # out-of-range parent would throw in the reference; treat as root
                    break
                j = p
                guard += 1
                if guard > n:
                    break
            order.extend(reversed(chain))
        return order

    def computeWorldMatrices(self) -> None:
        """model.ts:330-420: world = parent * T(bind) * R(append o local) * T(appendMove*ratio)."""
        bones = self.skeleton.bones
        n = len(bones)
        lr = self.localRotations
        lt = self.localTranslations
        world = self.worldMatrices
        m1 = Mat4.identity()
        m2 = Mat4.identity()
        tmp = np.empty(16, dtype=np.float32)
        for i in self._order:
            b = bones[i]
            qi = i * 4
            rotateM = Mat4.fromQuat(float(lr[qi]), float(lr[qi + 1]), float(lr[qi + 2]), float(lr[qi + 3]))
            ax = ay = az = 0.0
            ap = b.appendParentIndex
            hasAppend = bool(b.appendRotate) and ap is not None and 0 <= ap < n
            if hasAppend:
                ratio = 1.0 if b.appendRatio is None else max(-1.0, min(1.0, b.appendRatio))
                if abs(ratio) > 1e-6:
                    aq = ap * 4
                    x, y, z, w = float(lr[aq]), float(lr[aq + 1]), float(lr[aq + 2]), float(lr[aq + 3])
                    absr = -ratio if ratio < 0 else ratio
                    if ratio < 0:
                        x, y, z = -x, -y, -z
                    r = Quat.slerp(Quat(0, 0, 0, 1), Quat(x, y, z, w), absr)
                    rotateM = Mat4.fromQuat(r.x, r.y, r.z, r.w).multiply(rotateM)
                    if b.appendMove:
                        ar = 1.0 if b.appendRatio is None else b.appendRatio
                        ax = float(lt[ap * 3]) * ar
                        ay = float(lt[ap * 3 + 1]) * ar
                        az = float(lt[ap * 3 + 2]) * ar
            m1.setIdentity().translateInPlace(b.bindTranslation[0], b.bindTranslation[1], b.bindTranslation[2])
            m2.setIdentity().translateInPlace(ax, ay, az)
            localM = m1.multiply(rotateM).multiply(m2)
            wo = i * 16
            if b.parentIndex >= 0:
                Mat4.multiplyArrays(world, b.parentIndex * 16, localM.values, 0, tmp, 0)
                world[wo:wo + 16] = tmp
            else:
                world[wo:wo + 16] = localM.values
