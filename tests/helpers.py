"""Test helpers: a tiny PMX 2.0 / VMD writer (our own data, exercising every weight type and
the morph table), error metrics, and independent numpy restatements used as cross-checks."""
from __future__ import annotations

import struct
from typing import List, Optional, Sequence

import numpy as np

REF_ROOT = "/root/reference"


def rel_err(out: np.ndarray, ref: np.ndarray) -> float:
    """SURVEY 8c tolerance metric: max_v |out-ref|_inf / max(|ref|_inf, 1)."""
    out = np.asarray(out, np.float64)
    ref = np.asarray(ref, np.float64)
    return float(np.abs(out - ref).max() / max(np.abs(ref).max(), 1.0))


class PmxWriter:
    def __init__(self, encoding=0, vertex_index_size=2, bone_index_size=2, morph_index_size=1, extra_vec4=0):
        self.b = bytearray()
        self.enc = encoding
        self.vs, self.bs, self.ms, self.xv = vertex_index_size, bone_index_size, morph_index_size, extra_vec4

    def text(self, s: str):
        raw = s.encode("utf-16-le" if self.enc == 0 else "utf-8")
        self.b += struct.pack("<i", len(raw)) + raw

    def idx(self, v: int, size: int, signed=True):
        fmt = {1: "b" if signed else "B", 2: "h" if signed else "H", 4: "i"}[size]
        self.b += struct.pack("<" + fmt, v)

    def header(self):
        self.b += b"PMX " + struct.pack("<f", 2.0) + bytes([8, self.enc, self.xv, self.vs, 1, 1, self.bs, self.ms, 1])
        for s in ("model", "model_en", "comment", "comment_en"):
            self.text(s)

    def vertices(self, verts: Sequence[dict]):
        self.b += struct.pack("<i", len(verts))
        for v in verts:
            self.b += struct.pack("<8f", *v["pos"], *v["nrm"], *v["uv"])
            self.b += b"\x00" * (16 * self.xv)
            t = v["type"]
            self.b += bytes([t])
            if t == 0:
                self.idx(v["bones"][0], self.bs)
            elif t in (1, 3):
                self.idx(v["bones"][0], self.bs)
                self.idx(v["bones"][1], self.bs)
                self.b += struct.pack("<f", v["w"][0])
                if t == 3:
                    self.b += struct.pack("<9f", *v["sdef"])
            elif t in (2, 4):
                for k in range(4):
                    self.idx(v["bones"][k], self.bs)
                self.b += struct.pack("<4f", *v["w"])
            self.b += struct.pack("<f", 1.0)

    def indices(self, idx: Sequence[int]):
        self.b += struct.pack("<i", len(idx))
        for i in idx:
            self.idx(i, self.vs, signed=False)

    def textures(self, names: Sequence[str]):
        self.b += struct.pack("<i", len(names))
        for n in names:
            self.text(n)

    def materials(self, n: int, vcount: int):
        self.b += struct.pack("<i", n)
        for i in range(n):
            self.text(f"mat{i}")
            self.text("")
            self.b += struct.pack("<4f3ff3f", 1, 1, 1, 1, 0, 0, 0, 5.0, 0.5, 0.5, 0.5)
            self.b += bytes([0x10]) + struct.pack("<4ff", 0, 0, 0, 1, 1.0)
            self.idx(-1, 1)
            self.idx(-1, 1)
            self.b += bytes([0, 1, 0])
            self.text("")
            self.b += struct.pack("<i", vcount)

    def bones(self, bones: Sequence[dict]):
        self.b += struct.pack("<i", len(bones))
        for bn in bones:
            self.text(bn["name"])
            self.text("")
            self.b += struct.pack("<3f", *bn["pos"])
            self.idx(bn["parent"], self.bs)
            self.b += struct.pack("<i", 0)
            flags = 0x0001 if bn.get("tail_bone") else 0
            if bn.get("append_rotate"):
                flags |= 0x0100
            if bn.get("append_move"):
                flags |= 0x0200
            if bn.get("axis"):
                flags |= 0x0400
            if bn.get("local_axis"):
                flags |= 0x0800
            if bn.get("ik"):
                flags |= 0x0020
            self.b += struct.pack("<H", flags)
            if flags & 1:
                self.idx(0, self.bs)
            else:
                self.b += struct.pack("<3f", 0, 1, 0)
            if flags & 0x0300:
                self.idx(bn["append_parent"], self.bs)
                self.b += struct.pack("<f", bn["append_ratio"])
            if flags & 0x0400:
                self.b += struct.pack("<3f", 1, 0, 0)
            if flags & 0x0800:
                self.b += struct.pack("<6f", 1, 0, 0, 0, 0, 1)
            if flags & 0x0020:
                self.idx(0, self.bs)
                self.b += struct.pack("<ifi", 10, 1.0, 2)
                self.idx(0, self.bs)
                self.b += bytes([1]) + struct.pack("<6f", *([0.0] * 6))
                self.idx(0, self.bs)
                self.b += bytes([0])

    def morphs(self, morphs: Sequence[dict]):
        self.b += struct.pack("<i", len(morphs))
        for m in morphs:
            self.text(m["name"])
            self.text("")
            self.b += bytes([1, m["type"]]) + struct.pack("<i", len(m["items"]))
            for it in m["items"]:
                if m["type"] == 1:
                    self.idx(it[0], self.vs, signed=False)
                    self.b += struct.pack("<3f", *it[1])
                elif m["type"] == 0:
                    self.idx(it[0], self.ms)
                    self.b += struct.pack("<f", it[1])
                elif m["type"] == 2:
                    self.idx(it[0], self.bs)
                    self.b += struct.pack("<7f", *([0.0] * 7))
                elif m["type"] == 3:
                    self.idx(it[0], self.vs, signed=False)
                    self.b += struct.pack("<4f", *([0.0] * 4))

    def tail(self, display=(), bodies=(), joints=()):
        """Display frames (name, flag, elements (kind 0 bone | 1 morph, index)), rigid bodies and joints: the sections after the
        morphs as pmx-loader.ts:555-789 reads them.  rbs = rigid-body index size."""
        rbs = getattr(self, "rbs", 1)
        self.b += struct.pack("<i", len(display))
        for name, elems in display:
            self.text(name)
            self.text("")
            self.b += bytes([0]) + struct.pack("<i", len(elems))
            for kind, index in elems:
                self.b += bytes([kind])
                self.idx(index, self.bs if kind == 0 else self.ms)
        self.b += struct.pack("<i", len(bodies))
        for rb in bodies:
            self.text(rb["name"])
            self.text("")
            self.idx(rb["boneIndex"], self.bs)
            self.b += bytes([rb["group"]]) + struct.pack("<H", rb["collisionMask"]) + bytes([rb["shape"]])
            self.b += struct.pack("<14f", *rb["size"], *rb["shapePosition"], *rb["shapeRotation"], rb["mass"], rb["linearDamping"],
                                  rb["angularDamping"], rb["restitution"], rb["friction"])
            self.b += bytes([rb["type"]])
        self.b += struct.pack("<i", len(joints))
        for jt in joints:
            self.text(jt["name"])
            self.text("")
            self.b += bytes([jt["type"]])
            self.idx(jt["rigidbodyIndexA"], rbs)
            self.idx(jt["rigidbodyIndexB"], rbs)
            self.b += struct.pack("<24f", *jt["position"], *jt["rotation"], *jt["positionMin"], *jt["positionMax"], *jt["rotationMin"],
                                  *jt["rotationMax"], *jt["springPosition"], *jt["springRotation"])

    def bytes(self) -> bytes:
        return bytes(self.b)


def write_vmd(keys: Sequence[tuple], morph_keys: Sequence[tuple] = ()) -> bytes:
    """keys: (boneName, frame, (x,y,z,w)); morph_keys: (name, frame, weight)."""
    b = bytearray(b"Vocaloid Motion Data 0002".ljust(30, b"\x00"))
    b += b"model".ljust(20, b"\x00")
    b += struct.pack("<I", len(keys))
    for name, frame, q in keys:
        b += name.encode("shift_jis").ljust(15, b"\x00")[:15]
        b += struct.pack("<I3f4f", frame, 0, 0, 0, *q)
        b += bytes(64)
    b += struct.pack("<I", len(morph_keys))
    for name, frame, w in morph_keys:
        b += name.encode("shift_jis").ljust(15, b"\x00")[:15] + struct.pack("<If", frame, w)
    b += struct.pack("<III", 0, 0, 0)
    return bytes(b)


def random_pmx(rng, V=300, B=12, n_morph=3, with_sdef=True, **kw):
    """Synthetic PMX bytes covering BDEF1/2/4/SDEF/QDEF, append bones, IK block, all morph kinds."""
    w = PmxWriter(**kw)
    w.header()
    verts = []
    for i in range(V):
        t = int(rng.choice([0, 1, 2, 3, 4] if with_sdef else [0, 1, 2, 4]))
        bones = [int(x) for x in rng.integers(-1, B, 4)]
        if t in (2, 4):
            ws = rng.dirichlet(np.ones(4)).astype(np.float32)
            if rng.random() < 0.15:
                ws[int(rng.integers(0, 4))] = 0.0
            if rng.random() < 0.05:
                ws[:] = 0.0
        else:
            ws = np.array([rng.uniform(-0.1, 1.1), 0, 0, 0], np.float32)
        p = rng.normal(size=3) * 3
        n = rng.normal(size=3)
        n /= np.linalg.norm(n)
        v = dict(pos=[float(x) for x in p], nrm=[float(x) for x in n], uv=[float(x) for x in rng.random(2)], type=t, bones=bones,
                 w=[float(x) for x in ws])
        if t == 3:
            c = p + rng.normal(size=3) * 0.1
            d = rng.normal(size=3) * 0.2
            v["sdef"] = [float(x) for x in np.concatenate([c, c + d, c - d])]
        verts.append(v)
    w.vertices(verts)
    w.indices([int(x) for x in rng.integers(0, V, 3 * 10)])
    w.textures(["a.png"])
    w.materials(1, 30)
    bones = []
    pos = np.zeros((B, 3))
    for i in range(B):
        parent = -1 if i == 0 else int(rng.integers(0, i))
        pos[i] = (pos[parent] if parent >= 0 else 0) + rng.normal(size=3)
        bn = dict(name=f"骨{i}", pos=[float(x) for x in pos[i]], parent=parent, tail_bone=bool(i % 2), ik=(i == 3),
                  axis=(i == 4), local_axis=(i == 5))
        if i >= 2 and i % 3 == 0:
            bn.update(append_rotate=True, append_parent=int(rng.integers(0, i)), append_ratio=float(rng.choice([-1.0, 0.25, 0.5, 1.5])))
        bones.append(bn)
    w.bones(bones)
    morphs = []
    for m in range(n_morph):
        n = int(rng.integers(1, max(2, V // 5)))
        vi = rng.choice(V, size=n, replace=False)
        morphs.append(dict(name=f"m{m}", type=1, items=[(int(v), [float(x) for x in rng.normal(0, 0.05, 3)]) for v in vi]))
    if n_morph >= 2:
        morphs.append(dict(name="grp", type=0, items=[(0, 0.5), (1, -1.0)]))
        morphs.append(dict(name="bonem", type=2, items=[(0, None)]))
        morphs.append(dict(name="uvm", type=3, items=[(0, None), (1, None)]))
    w.morphs(morphs)
    f3 = lambda: [float(np.float32(x)) for x in rng.normal(size=3)]
    bodies = [dict(name=f"rb{i}", boneIndex=int(rng.integers(-1, B)), group=int(rng.integers(0, 16)), collisionMask=int(rng.integers(0, 65536)),
                   shape=int(rng.integers(0, 3)), size=f3(), shapePosition=f3(), shapeRotation=f3(), mass=float(np.float32(rng.uniform(0, 2))),
                   linearDamping=0.5, angularDamping=0.25, restitution=0.0, friction=0.5, type=int(rng.integers(0, 3))) for i in range(5)]
    pjoints = [dict(name=f"j{i}", type=0, rigidbodyIndexA=int(rng.integers(-1, 5)), rigidbodyIndexB=int(rng.integers(0, 5)), position=f3(), rotation=f3(),
                    positionMin=f3(), positionMax=f3(), rotationMin=f3(), rotationMax=f3(), springPosition=f3(), springRotation=f3()) for i in range(3)]
    w.tail(display=[("Root", [(0, 0)]), ("表情", [(1, 0), (1, 1), (0, 2)])], bodies=bodies, joints=pjoints)
    w.rigid = (bodies, pjoints)
    random_pmx.last_rigid = (bodies, pjoints)
    return w.bytes(), verts, bones, morphs


def numpy_blend_f64(vtx8, joints, weights, skin16):
    """Independent dense restatement of engine.ts:253-272 in f64 (cross-check of the C oracle)."""
    vtx8 = np.asarray(vtx8, np.float64).reshape(-1, 8)
    V = vtx8.shape[0]
    J = np.asarray(joints).reshape(V, 4)
    W = np.asarray(weights, np.float64).reshape(V, 4) / 255.0
    s = W.sum(axis=1, keepdims=True)
    Wn = np.where(s > 1e-4, W / np.where(s > 1e-4, s, 1), np.array([1.0, 0, 0, 0]))
    M = np.asarray(skin16, np.float64).reshape(-1, 4, 4).transpose(0, 2, 1)     # [B,row,col]
    p4 = np.concatenate([vtx8[:, :3], np.ones((V, 1))], axis=1)
    pos = np.zeros((V, 3))
    nrm = np.zeros((V, 3))
    for i in range(4):
        Mi = M[J[:, i]]
        pos += np.einsum("vrc,vc->vr", Mi[:, :3, :], p4) * Wn[:, i:i + 1]
        nrm += np.einsum("vrc,vc->vr", Mi[:, :3, :3], vtx8[:, 3:6]) * Wn[:, i:i + 1]
    ln = np.linalg.norm(nrm, axis=1, keepdims=True)
    nrm = np.where(ln > 0, nrm / np.where(ln > 0, ln, 1), 0)
    return pos, nrm


def numpy_morph_sdef_f64(vtx8, joints, weights, skin16, morph=None, morphW=None, sdef=None):
    """Independent f64 restatement of the two extensions the reference does not have, vertex by vertex in plain
    numpy -- a second implementation next to oracle/rz_oracle_body.inc so that the unpinned features are at least pinned to
    each other.  Morph: p~ = p + sum_m w_m * delta_m[v] in PMX morph order, before skinning.  SDEF (two-influence vertices
    only): q = slerp(quat(M0), quat(M1), w1) with the reference's own toQuatFromArray / slerp / fromQuat (math.ts:406-448,
    156-189, 352-384), pos' = R(q)(p~ - C) + w0 M0 c0 + w1 M1 c1, n' = normalize(R(q) n); the load-time constants c0, c1 are
    float32 like in the product (rze_b200.cu / mesh_tables.h)."""
    vtx8 = np.asarray(vtx8, np.float64).reshape(-1, 8)
    V = vtx8.shape[0]
    J = np.asarray(joints).reshape(V, 4)
    W8 = np.asarray(weights).reshape(V, 4)
    M = np.asarray(skin16, np.float64).reshape(-1, 4, 4).transpose(0, 2, 1)     # [B,row,col]
    P = vtx8[:, :3].copy()
    if morph is not None and morphW is not None:
        off, vi, d3 = np.asarray(morph[0], np.int64), np.asarray(morph[1], np.int64), np.asarray(morph[2], np.float64).reshape(-1, 3)
        for m in range(len(off) - 1):                                            # PMX order; a vertex listed twice adds twice
            for e in range(off[m], off[m + 1]):
                P[vi[e]] = P[vi[e]] + np.float64(np.float32(morphW[m])) * d3[e]
    pos, nrm = numpy_blend_f64(np.concatenate([P, vtx8[:, 3:]], axis=1), joints, weights, skin16)

    def quat_of(m):
        m00, m01, m02, m10, m11, m12, m20, m21, m22 = m[0, 0], m[0, 1], m[0, 2], m[1, 0], m[1, 1], m[1, 2], m[2, 0], m[2, 1], m[2, 2]
        tr = m00 + m11 + m22
        if tr > 0:
            s_ = np.sqrt(tr + 1.0) * 2
            q = np.array([(m21 - m12) / s_, (m02 - m20) / s_, (m10 - m01) / s_, 0.25 * s_])
        elif m00 > m11 and m00 > m22:
            s_ = np.sqrt(1.0 + m00 - m11 - m22) * 2
            q = np.array([0.25 * s_, (m01 + m10) / s_, (m02 + m20) / s_, (m21 - m12) / s_])
        elif m11 > m22:
            s_ = np.sqrt(1.0 + m11 - m00 - m22) * 2
            q = np.array([(m01 + m10) / s_, 0.25 * s_, (m12 + m21) / s_, (m02 - m20) / s_])
        else:
            s_ = np.sqrt(1.0 + m22 - m00 - m11) * 2
            q = np.array([(m02 + m20) / s_, (m12 + m21) / s_, 0.25 * s_, (m10 - m01) / s_])
        return q / np.linalg.norm(q)

    def slerp(a, b, t):
        c = float(a @ b)
        if c < 0:
            c, b = -c, -b
        if c > 0.9995:
            v = a + t * (b - a)
            return v / np.linalg.norm(v)
        th0 = np.arccos(c)
        return (np.sin(th0 - th0 * t) * a + np.sin(th0 * t) * b) / np.sin(th0)

    def rot_of(q):
        x, y, z, w = q
        return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                         [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                         [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])

    if sdef is not None:
        sv, vec = np.asarray(sdef[0], np.int64), np.asarray(sdef[1], np.float32).reshape(-1, 9)
        for i, v in enumerate(sv):
            if W8[v, 2] != 0 or W8[v, 3] != 0:
                continue
            w = W8[v].astype(np.float64) / 255.0
            w = w / w.sum() if w.sum() > 1e-4 else np.array([1.0, 0, 0, 0])
            w0f, w1f = np.float32(w[0]), np.float32(w[1])
            C, R0, R1 = vec[i, :3], vec[i, 3:6], vec[i, 6:9]
            rw = w0f * R0 + w1f * R1                                             # float32 arithmetic, like the table builder
            c0 = ((C + (C + R0 - rw)) * np.float32(0.5)).astype(np.float64)
            c1 = ((C + (C + R1 - rw)) * np.float32(0.5)).astype(np.float64)
            M0, M1 = M[J[v, 0]], M[J[v, 1]]
            R = rot_of(slerp(quat_of(M0), quat_of(M1), w[1]))
            pos[v] = R @ (P[v] - C.astype(np.float64)) + w[0] * (M0[:3, :3] @ c0 + M0[:3, 3]) + w[1] * (M1[:3, :3] @ c1 + M1[:3, 3])
            n = R @ vtx8[v, 3:6]
            ln = np.linalg.norm(n)
            nrm[v] = n / ln if ln > 0 else 0
    return pos, nrm


def sdef_from_record(rec, M, p, n):
    """Evaluate ONE SDEF vertex from its 12-float device record (C, c0, c1, w0, w1, rows j0 | j1 << 16) the way the kernel's dense
    phase does, in f64: q = slerp(quat(M0), quat(M1), w1), pos = R(q)(p - C) + w0 M0 c0 + w1 M1 c1, n = normalize(R(q) n).
    M: [B,4,4] row-major skin matrices."""
    C, c0, c1 = rec[0:3].astype(np.float64), rec[3:6].astype(np.float64), rec[6:9].astype(np.float64)
    w0, w1 = float(rec[9]), float(rec[10])
    rows = int(rec[11:12].view(np.uint32)[0])
    M0, M1 = M[rows & 0xFFFF], M[rows >> 16]

    def quat_of(m):
        m00, m01, m02, m10, m11, m12, m20, m21, m22 = m[0, 0], m[0, 1], m[0, 2], m[1, 0], m[1, 1], m[1, 2], m[2, 0], m[2, 1], m[2, 2]
        tr = m00 + m11 + m22
        if tr > 0:
            s_ = np.sqrt(tr + 1.0) * 2
            q = np.array([(m21 - m12) / s_, (m02 - m20) / s_, (m10 - m01) / s_, 0.25 * s_])
        elif m00 > m11 and m00 > m22:
            s_ = np.sqrt(1.0 + m00 - m11 - m22) * 2
            q = np.array([0.25 * s_, (m01 + m10) / s_, (m02 + m20) / s_, (m21 - m12) / s_])
        elif m11 > m22:
            s_ = np.sqrt(1.0 + m11 - m00 - m22) * 2
            q = np.array([(m01 + m10) / s_, 0.25 * s_, (m12 + m21) / s_, (m02 - m20) / s_])
        else:
            s_ = np.sqrt(1.0 + m22 - m00 - m11) * 2
            q = np.array([(m02 + m20) / s_, (m12 + m21) / s_, 0.25 * s_, (m10 - m01) / s_])
        return q / np.linalg.norm(q)
    a, b = quat_of(M0), quat_of(M1)
    c = float(a @ b)
    if c < 0:
        c, b = -c, -b
    if c > 0.9995:
        q = a + w1 * (b - a)
        q /= np.linalg.norm(q)
    else:
        th0 = np.arccos(c)
        q = (np.sin(th0 - th0 * w1) * a + np.sin(th0 * w1) * b) / np.sin(th0)
    x, y, z, w = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                  [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                  [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
    pos = R @ (p - C) + w0 * (M0[:3, :3] @ c0 + M0[:3, 3]) + w1 * (M1[:3, :3] @ c1 + M1[:3, 3])
    nn = R @ n
    ln = np.linalg.norm(nn)
    return pos, (nn / ln if ln > 0 else nn * 0)
