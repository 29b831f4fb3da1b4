"""PMX 2.x loader -> Model, host side of the deform path.

Follows the reference's `engine/src/pmx-loader.ts` for everything the deform
stage consumes, bit-for-bit on the integer outputs:

* header / globals / text           pmx-loader.ts:51-96, 1031-1048
* vertices + weight quantisation    pmx-loader.ts:98-189
* bones (+append fields, IK skipped) pmx-loader.ts:311-448
* inverse bind (pure translations)  pmx-loader.ts:791-824
* joints clamp + renormalise to 255 pmx-loader.ts:857-939
* index readers (unsigned vertex / signed other) pmx-loader.ts:981-1005

Where the reference *skips* data the B200 path needs, this loader keeps it
("new feature, unpinned"): SDEF C/R0/R1 (pmx-loader.ts:153-155 steps
over them) and vertex/group morphs (pmx-loader.ts:450-553 walks and discards
them).  Morph strides follow the PMX 2.0 spec (bone morph 7 floats, UV morph 4
floats) so files containing those types stay in sync.

Sections after the morphs: display frames are stepped over, rigid bodies and
joints are kept as the reference's records (pmx-loader.ts:555-789) -- the solver
that consumes them is out of scope, but the rigid bodies feed the
physics -> bone feedback plumbing (physics_bridge.py, rz_load_rigid_bodies).
"""
from __future__ import annotations

import math
import struct
from typing import List, Optional, Tuple

import numpy as np

from .model import Bone, Model, SdefTable, Skeleton, Skinning, VertexMorphs

_NAN = float("nan")


def _js_min(a: float, b: float) -> float:
    if a != a or b != b:
        return _NAN
    return a if a < b else b


def _js_max(a: float, b: float) -> float:
    if a != a or b != b:
        return _NAN
    return a if a > b else b


def _js_round(x: float) -> float:
    """JS Math.round: nearest, halves toward +inf; NaN stays NaN."""
    if x != x or math.isinf(x):
        return x
    return float(math.floor(x + 0.5))


def _to_u8(x: float) -> int:
    """Store into a Uint8Array (ToUint8: NaN/inf -> 0, else trunc mod 256)."""
    if x != x or math.isinf(x):
        return 0
    return int(math.trunc(x)) & 0xFF


def _to_u16(x: int) -> int:
    return int(x) & 0xFFFF


class PmxFormatError(Exception):
    pass


class PmxLoader:
    def __init__(self, data: bytes):
        self.buf = memoryview(data)
        self.offset = 0
        self.encoding = 0
        self.additionalVec4Count = 0
        self.vertexIndexSize = 0
        self.textureIndexSize = 0
        self.materialIndexSize = 0
        self.boneIndexSize = 0
        self.morphIndexSize = 0
        self.rigidBodyIndexSize = 0
        self.warnings: List[str] = []

    # ---- public ----------------------------------------------------------------------
    @staticmethod
    def load(path: str, clock=None) -> Model:
        """`PmxLoader.load(url)` (pmx-loader.ts:30-33); `fetch` becomes a file read."""
        with open(path, "rb") as f:
            return PmxLoader(f.read()).parse(clock=clock)

    @staticmethod
    def loadFromBuffer(data: bytes, clock=None) -> Model:
        return PmxLoader(data).parse(clock=clock)

    def parse(self, clock=None) -> Model:
        self.parseHeader()
        vtx, joints, weights, sdef = self.parseVertices()
        indices = self.parseIndices()
        textures = self.parseTextures()
        materials = self.parseMaterials()
        bones = self.parseBones()
        morphs = self.parseMorphs(vtx.shape[0])
        # display frames are stepped over, rigid bodies and joints kept (pmx-loader.ts:42-48, 555-789): they feed the
        # physics -> bone feedback (physics_bridge.py); soft failures leave the lists empty like the reference's try/catch
        rigidbodies, pjoints = [], []
        try:
            if self.skipDisplayFrames():
                rigidbodies = self.parseRigidbodies()
                pjoints = self.parseJoints()
        except (PmxFormatError, struct.error, IndexError, ValueError):
            pass
        invBind = compute_inverse_bind(bones)
        finalize_skinning(joints, weights, len(bones))
        skeleton = Skeleton(bones=bones, inverseBindMatrices=invBind)
        return Model(vtx.reshape(-1), indices, textures, materials, skeleton,
                     Skinning(joints=joints, weights=weights), rigidbodies=rigidbodies, joints=pjoints, morphs=morphs, sdef=sdef,
                     clock=clock)

    # ---- primitive readers (pmx-loader.ts:965-1053) ------------------------------------
    def _need(self, n: int):
        if self.offset + n > len(self.buf):
            raise PmxFormatError(f"Offset {self.offset} + {n} exceeds buffer bounds {len(self.buf)}")

    def u8(self) -> int:
        self._need(1)
        v = self.buf[self.offset]
        self.offset += 1
        return v

    def u16(self) -> int:
        self._need(2)
        v = struct.unpack_from("<H", self.buf, self.offset)[0]
        self.offset += 2
        return v

    def i32(self) -> int:
        self._need(4)
        v = struct.unpack_from("<i", self.buf, self.offset)[0]
        self.offset += 4
        return v

    def f32(self) -> float:
        self._need(4)
        v = struct.unpack_from("<f", self.buf, self.offset)[0]
        self.offset += 4
        return v

    def vertexIndex(self) -> int:
        """unsigned for 1/2 bytes, int32 for 4 (pmx-loader.ts:981-990)."""
        s = self.vertexIndexSize
        if s == 1:
            return self.u8()
        if s == 2:
            return self.u16()
        return self.i32()

    def index(self, size: int) -> int:
        """signed non-vertex index (pmx-loader.ts:992-1005)."""
        self._need(size if size in (1, 2) else 4)
        if size == 1:
            v = struct.unpack_from("<b", self.buf, self.offset)[0]
            self.offset += 1
            return v
        if size == 2:
            v = struct.unpack_from("<h", self.buf, self.offset)[0]
            self.offset += 2
            return v
        return self.i32()

    def text(self) -> str:
        n = self.i32()
        if n <= 0:
            return ""
        if n > 1000:
            raise PmxFormatError(f"Suspicious string length: {n} at offset {self.offset - 4}")
        self._need(n)
        raw = bytes(self.buf[self.offset:self.offset + n])
        self.offset += n
        return raw.decode("utf-16-le" if self.encoding == 0 else "utf-8", errors="replace")

    # ---- sections ----------------------------------------------------------------------
    def parseHeader(self):
        if bytes(self.buf[0:3]) != b"PMX":
            raise PmxFormatError("Not a PMX file")
        self.offset = 4
        version = self.f32()
        if version < 2.0 or version > 2.2:
            self.warnings.append(f"PMX version {version} may not be fully supported")
        g = self.u8()
        if g < 8:
            raise PmxFormatError(f"Invalid globalsCount: {g}, expected at least 8")
        (self.encoding, self.additionalVec4Count, self.vertexIndexSize, self.textureIndexSize,
         self.materialIndexSize, self.boneIndexSize, self.morphIndexSize, self.rigidBodyIndexSize) = (
            self.u8() for _ in range(8))
        for _ in range(8, g):
            self.u8()
        for _ in range(4):
            self.text()

    def parseVertices(self):
        count = self.i32()
        vtx = np.zeros((max(count, 0), 8), dtype=np.float32)
        joints = np.zeros(max(count, 0) * 4, dtype=np.uint16)
        weights = np.zeros(max(count, 0) * 4, dtype=np.uint8)
        sdef_idx: List[int] = []
        sdef_vec: List[Tuple[float, ...]] = []
        sdef_w0: List[float] = []
        bs = self.boneIndexSize
        skip_extra = self.additionalVec4Count * 16
        unpack8 = struct.Struct("<8f").unpack_from
        for i in range(count):
            self._need(32)
            vtx[i] = unpack8(self.buf, self.offset)
            self.offset += 32 + skip_extra
            t = self.u8()
            base = i * 4
            weights[base] = 255
            if t == 0:                                   # BDEF1
                j0 = self.index(bs)
                joints[base] = _to_u16(j0 if j0 >= 0 else 0)
            elif t == 1 or t == 3:                       # BDEF2 / SDEF-as-BDEF2
                j0 = self.index(bs)
                j1 = self.index(bs)
                w0f = self.f32()
                w0 = _js_max(0, _js_min(255, _js_round(w0f * 255)))
                w1 = _js_max(0, _js_min(255, 255 - w0))
                joints[base] = _to_u16(j0 if j0 >= 0 else 0)
                joints[base + 1] = _to_u16(j1 if j1 >= 0 else 0)
                weights[base] = _to_u8(w0)
                weights[base + 1] = _to_u8(w1)
                if t == 3:
                    self._need(36)
                    sdef_idx.append(i)
                    sdef_vec.append(struct.unpack_from("<9f", self.buf, self.offset))
                    sdef_w0.append(w0f)
                    self.offset += 36
            elif t == 2 or t == 4:                       # BDEF4 / QDEF-as-BDEF4
                for k in range(4):
                    j = self.index(bs)
                    joints[base + k] = _to_u16(j if j >= 0 else 0)
                wf = [self.f32() for _ in range(4)]
                w8 = quantize_bdef4(wf)
                for k in range(4):
                    weights[base + k] = w8[k]
            else:
                raise PmxFormatError(f"Invalid bone weight type: {t}")
            self.offset += 4                             # edge scale
        sdef = SdefTable(np.asarray(sdef_idx, dtype=np.uint32),
                         np.asarray(sdef_vec, dtype=np.float32).reshape(-1, 9),
                         np.asarray(sdef_w0, dtype=np.float32))
        return vtx, joints, weights, sdef

    def parseIndices(self) -> np.ndarray:
        count = self.i32()
        s = self.vertexIndexSize
        if s in (1, 2, 4) and count >= 0:
            self._need(count * s)
            dt = {1: "<u1", 2: "<u2", 4: "<i4"}[s]
            arr = np.frombuffer(self.buf, dtype=dt, count=count, offset=self.offset).astype(np.uint32)
            self.offset += count * s
            return arr
        return np.asarray([self.vertexIndex() for _ in range(count)], dtype=np.uint32)

    def parseTextures(self):
        count = self.i32()
        out = []
        for _ in range(count):
            p = self.text()
            out.append({"path": p, "name": (p.split("/")[-1] or p)})
        return out

    def parseMaterials(self):
        """Walks the material table (pmx-loader.ts:222-309).  Only the fields the
        façade exposes are kept; shading belongs to the rasteriser (out of scope)."""
        count = self.i32()
        mats = []
        ts = self.textureIndexSize
        for _ in range(count):
            name = self.text()
            self.text()
            diffuse = [self.f32() for _ in range(4)]
            specular = [self.f32() for _ in range(3)]
            shininess = self.f32()
            ambient = [self.f32() for _ in range(3)]
            flag = self.u8()
            edgeColor = [self.f32() for _ in range(4)]
            edgeSize = self.f32()
            tex = self.index(ts)
            sph = self.index(ts)
            sphMode = self.u8()
            shared = self.u8() == 1
            toon = self.u8() if shared else self.index(ts)
            self.text()
            vcount = self.i32()
            mats.append(dict(name=name, diffuse=diffuse, specular=specular, ambient=ambient, shininess=shininess,
                             diffuseTextureIndex=tex, sphereTextureIndex=sph, sphereMode=sphMode,
                             toonTextureIndex=toon, edgeFlag=flag, edgeColor=edgeColor, edgeSize=edgeSize,
                             vertexCount=vcount))
        return mats

    def parseBones(self) -> List[Bone]:
        count = self.i32()
        bs = self.boneIndexSize
        absb = []
        for _ in range(count):
            name = self.text()
            self.text()
            x, y, z = self.f32(), self.f32(), self.f32()
            parent = self.index(bs)
            self.i32()                                    # transform layer
            flags = self.u16()
            if flags & 0x0001:
                self.index(bs)
            else:
                self.offset += 12
            appendParent = appendRatio = None
            appendRotate = appendMove = False
            if flags & 0x0300:
                appendParent = self.index(bs)
                appendRatio = self.f32()
                appendRotate = bool(flags & 0x0100)
                appendMove = bool(flags & 0x0200)
            if flags & 0x0400:
                self.offset += 12
            if flags & 0x0800:
                self.offset += 24
            if flags & 0x2000:
                self.i32()
            if flags & 0x0020:                            # IK block: parsed and dropped
                self.index(bs)
                self.i32()
                self.f32()
                links = self.i32()
                for _ in range(links):
                    self.index(bs)
                    if self.u8() == 1:
                        self.offset += 24
            absb.append((name, parent, x, y, z, appendParent, appendRatio, appendRotate, appendMove))
        bones: List[Bone] = []
        for (name, parent, x, y, z, ap, ar, arot, amov) in absb:
            if 0 <= parent < count:
                p = absb[parent]
                bt = [x - p[2], y - p[3], z - p[4]]       # f64 difference of f32 reads
            else:
                bt = [x, y, z]
            bones.append(Bone(name=name, parentIndex=parent, bindTranslation=bt, appendParentIndex=ap,
                              appendRatio=ar, appendRotate=arot, appendMove=amov))
        return bones

    def skipDisplayFrames(self) -> bool:
        """pmx-loader.ts:555-601: name, english name, flag, then n x (u8 kind + bone | morph index)."""
        if self.offset + 4 > len(self.buf):
            return False
        count = self.i32()
        if count < 0 or count > 100000:
            self.offset -= 4
            return False
        for _ in range(count):
            self.text()
            self.text()
            self.u8()
            n = self.i32()
            for _j in range(n):
                kind = self.u8()
                if kind == 0:
                    self.index(self.boneIndexSize)
                elif kind == 1:
                    self.index(self.morphIndexSize)
        return True

    def parseRigidbodies(self) -> list:
        """pmx-loader.ts:603-690: the reference's Rigidbody records (model.ts:52-70 field names).  type 0 static, 1 dynamic
        (the only ones that drive bones, physics.ts:729), 2 kinematic; rotations are ZXY Euler angles in radians."""
        if self.offset + 4 > len(self.buf):
            return []
        count = self.i32()
        if count < 0 or count > 10000:
            return []
        out = []
        for _ in range(count):
            if self.offset >= len(self.buf):
                break
            name = self.text()
            englishName = self.text()
            boneIndex = self.index(self.boneIndexSize)
            group = self.u8()
            collisionMask = self.u16()
            shape = self.u8()
            f = [self.f32() for _k in range(14)]
            rtype = self.u8()
            out.append(dict(name=name, englishName=englishName, boneIndex=boneIndex, group=group, collisionMask=collisionMask, shape=shape,
                            size=f[0:3], shapePosition=f[3:6], shapeRotation=f[6:9], mass=f[9], linearDamping=f[10],
                            angularDamping=f[11], restitution=f[12], friction=f[13], type=rtype))
        return out

    def parseJoints(self) -> list:
        """pmx-loader.ts:692-789: 6-DoF spring joints between two rigid bodies (indices use rigidBodyIndexSize)."""
        if self.offset + 4 > len(self.buf):
            return []
        count = self.i32()
        if count < 0 or count > 10000:
            return []
        out = []
        for _ in range(count):
            if self.offset >= len(self.buf):
                break
            name = self.text()
            englishName = self.text()
            jtype = self.u8()
            a = self.index(self.rigidBodyIndexSize)
            b = self.index(self.rigidBodyIndexSize)
            f = [self.f32() for _k in range(24)]
            out.append(dict(name=name, englishName=englishName, type=jtype, rigidbodyIndexA=a, rigidbodyIndexB=b, position=f[0:3],
                            rotation=f[3:6], positionMin=f[6:9], positionMax=f[9:12], rotationMin=f[12:15], rotationMax=f[15:18],
                            springPosition=f[18:21], springRotation=f[21:24]))
        return out

    def parseMorphs(self, vertexCount: int) -> VertexMorphs:
        """Vertex (type 1) and group (type 0) morphs; layout as documented by the
        reference's skip code (pmx-loader.ts:462-541).  Group morphs are expanded
        into their member vertex morphs x ratio at load.  Every PMX
        morph keeps its slot (non-vertex types become empty) so morph indices used
        by VMD morph tracks / the caller stay PMX indices."""
        if self.offset + 4 > len(self.buf):
            return VertexMorphs.empty()
        count = self.i32()
        if count < 0 or count > 100000:
            self.offset -= 4
            return VertexMorphs.empty()
        names: List[str] = []
        kinds: List[int] = []
        vtx_lists: List[Tuple[np.ndarray, np.ndarray]] = []
        groups: List[List[Tuple[int, float]]] = []
        vs = self.vertexIndexSize
        for _ in range(count):
            name = self.text()
            self.text()
            self.u8()
            mtype = self.u8()
            n = self.i32()
            names.append(name)
            kinds.append(mtype)
            vi = np.zeros(0, np.uint32)
            dl = np.zeros((0, 3), np.float32)
            grp: List[Tuple[int, float]] = []
            if mtype == 1:
                rec = np.dtype([("i", {1: "<u1", 2: "<u2", 4: "<i4"}.get(vs, "<i4")), ("d", "<f4", (3,))])
                self._need(n * rec.itemsize)
                a = np.frombuffer(self.buf, dtype=rec, count=n, offset=self.offset)
                self.offset += n * rec.itemsize
                vi = a["i"].astype(np.int64)
                ok = (vi >= 0) & (vi < vertexCount)
                vi = vi[ok].astype(np.uint32)
                dl = np.ascontiguousarray(a["d"][ok], dtype=np.float32)
            elif mtype == 0:
                for _j in range(n):
                    grp.append((self.index(self.morphIndexSize), self.f32()))
            elif mtype == 2:
                self.offset += n * (self.boneIndexSize + 28)
            elif 3 <= mtype <= 7:
                self.offset += n * (vs + 16)
            elif mtype == 8:
                self.offset += n * (self.materialIndexSize + 1 + 28 * 4)
            elif mtype == 9:
                self.offset += n * (self.morphIndexSize + 4)
            elif mtype == 10:
                self.offset += n * (self.rigidBodyIndexSize + 1 + 24)
            else:
                raise PmxFormatError(f"Unknown morph type {mtype}")
            self._need(0)
            vtx_lists.append((vi, dl))
            groups.append(grp)
        # expand group morphs (one level, as MMD does; nested groups are ignored)
        for m in range(count):
            if kinds[m] != 0:
                continue
            vis, dls = [], []
            for (mi, ratio) in groups[m]:
                if 0 <= mi < count and kinds[mi] == 1:
                    vis.append(vtx_lists[mi][0])
                    dls.append((vtx_lists[mi][1].astype(np.float64) * ratio).astype(np.float32))
            if vis:
                vtx_lists[m] = (np.concatenate(vis), np.concatenate(dls))
        offsets = np.zeros(count + 1, dtype=np.uint32)
        for m in range(count):
            offsets[m + 1] = offsets[m] + len(vtx_lists[m][0])
        vi_all = np.concatenate([v for v, _ in vtx_lists]) if count else np.zeros(0, np.uint32)
        dl_all = np.concatenate([d for _, d in vtx_lists]) if count else np.zeros((0, 3), np.float32)
        return VertexMorphs(names, offsets, vi_all.astype(np.uint32), dl_all.reshape(-1, 3).astype(np.float32))


def quantize_bdef4(wf) -> List[int]:
    """BDEF4/QDEF float weights -> 4 x u8 (pmx-loader.ts:163-179)."""
    ws = [_js_max(0, _js_min(1, x)) for x in wf]
    w8 = [_js_round(x * 255) for x in ws]
    s = w8[0] + w8[1] + w8[2] + w8[3]
    out = [255, 0, 0, 0]
    if s == 0:
        return out
    scale = 255 / s if s == s else _NAN
    accum = 0.0
    for k in range(3):
        v = _js_max(0, _js_min(255, _js_round(w8[k] * scale)))
        out[k] = _to_u8(v)
        accum += v
    out[3] = _to_u8(_js_max(0, _js_min(255, 255 - accum)))
    return out


def compute_inverse_bind(bones: List[Bone]) -> np.ndarray:
    """invBind[b] = T(-bindWorld[b].t); bind world is the f32 chain of translations
    (pmx-loader.ts:791-824: Mat4.identity().translateInPlace + Mat4.multiply)."""
    n = len(bones)
    inv = np.zeros(n * 16, dtype=np.float32)
    if n == 0:
        return inv
    wt = np.zeros((n, 3), dtype=np.float32)
    done = [False] * n

    def world(i: int):
        if done[i]:
            return
        b = bones[i]
        lt = np.asarray(b.bindTranslation, dtype=np.float64).astype(np.float32)  # translateInPlace: 0 + t -> f32
        if 0 <= b.parentIndex < n:
            world(b.parentIndex)
            # (parent * local) column 3 = p.t*1 ... in f64: 1*lx + 0*ly + 0*lz + px*1, term order of math.ts:314
            p = wt[b.parentIndex].astype(np.float64)
            l64 = lt.astype(np.float64)
            wt[i] = (l64 + p).astype(np.float32)
        else:
            wt[i] = lt
        done[i] = True

    import sys
    sys.setrecursionlimit(max(sys.getrecursionlimit(), n + 100))
    for i in range(n):
        world(i)
    m = inv.reshape(n, 16)
    m[:, 0] = m[:, 5] = m[:, 10] = m[:, 15] = 1.0
    m[:, 12:15] = (0.0 - wt.astype(np.float64)).astype(np.float32) + np.float32(0.0)
    return inv


def finalize_skinning(joints: np.ndarray, weights: np.ndarray, boneCount: int) -> None:
    """Clamp joints to the bone range and renormalise weights to sum exactly 255,
    in place (pmx-loader.ts:857-939)."""
    V = joints.size // 4
    J = joints.reshape(V, 4)
    W = weights.reshape(V, 4)
    valid = J < boneCount
    sums = np.where(valid, W, 0).sum(axis=1)
    fast = valid.all(axis=1) & (sums == 255)
    for i in np.nonzero(~fast)[0]:
        j = [int(x) for x in J[i]]
        w = [int(x) for x in W[i]]
        vsum = 0
        vcount = 0
        for k in range(4):
            if j[k] >= boneCount:
                w[k] = 0
                j[k] = boneCount - 1 if boneCount > 0 else 0
            else:
                vsum += w[k]
                vcount += 1
        ok = lambda k: 0 <= j[k] < boneCount
        if vsum == 0 or vcount == 0:
            w = [255, 0, 0, 0]
            j = [0, 0, 0, 0]
        elif vsum != 255:
            scale = 255 / vsum
            accum = 0
            for k in range(3):
                if ok(k):
                    v = int(max(0, min(255, _js_round(w[k] * scale))))
                    w[k] = v
                    accum += v
                else:
                    w[k] = 0
            if ok(3):
                w[3] = max(0, min(255, 255 - accum))
            else:
                w[3] = 0
                if accum < 255:
                    for k in (2, 1, 0):
                        if ok(k) and w[k] > 0:
                            w[k] = min(255, w[k] + (255 - accum))
                            break
            fs = w[0] + w[1] + w[2] + w[3]
            if fs != 255:
                diff = 255 - fs
                mi, mw = 0, w[0]
                for k in range(1, 4):
                    if w[k] > mw and ok(k):
                        mw, mi = w[k], k
                if ok(mi):
                    w[mi] = max(0, min(255, w[mi] + diff))
        J[i] = j
        W[i] = w
