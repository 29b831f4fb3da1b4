"""ctypes front-end of the CPU oracle (oracle/rz_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under reze-engine_b200/ imports this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = [os.path.join(_HERE, "rz_oracle.c"), os.path.join(_HERE, "rz_oracle_body.inc")]
_OUT_DIR = os.path.join(_HERE, "_build")
LIB_PATH = os.path.join(_OUT_DIR, "liborc.so")
_lib = None


def build(force: bool = False) -> str:
    os.makedirs(_OUT_DIR, exist_ok=True)
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(s) <= os.path.getmtime(LIB_PATH) for s in _SRC):
        return LIB_PATH
    cmd = ["gcc", "-O2", "-std=c99", "-ffp-contract=off", "-fPIC", "-shared", "-Wall", "-o", LIB_PATH, _SRC[0], "-lm", "-lpthread"]
    subprocess.run(cmd, check=True)
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_quantize_bdef2.argtypes = [C.c_float, C.c_void_p]
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def skin_matrices(world: np.ndarray, invBind: np.ndarray, dtype=np.float32) -> np.ndarray:
    """[B,16] column-major skin = world * invBind (engine.ts:926-928)."""
    world = np.ascontiguousarray(world, np.float32).reshape(-1, 16)
    invBind = np.ascontiguousarray(invBind, np.float32).reshape(-1, 16)
    B = world.shape[0]
    out = np.empty((B, 16), dtype=dtype)
    fn = lib().orc_skin_matrices_f32 if dtype == np.float32 else lib().orc_skin_matrices_f64
    fn(_p(world), _p(invBind), C.c_uint32(B), _p(out))
    return out


def outline_hull(pos: np.ndarray, nrm: np.ndarray, edge_size) -> np.ndarray:
    """The outline vertex shader's expansion (engine.ts:458-461), on the blend's outputs:
        expandedPos = worldPos + worldNormal * material.edgeSize * scaleFactor,   scaleFactor = 0.01
    f32, evaluated left to right like the WGSL expression.  edge_size: [V] Material.edgeSize per vertex (0 = none)."""
    pos = np.asarray(pos, np.float32)
    nrm = np.asarray(nrm, np.float32)
    e = np.asarray(edge_size, np.float32).reshape(-1, 1)
    return (pos + (nrm * e) * np.float32(0.01)).astype(np.float32)


def interleaved(pos: np.ndarray, nrm: np.ndarray, vtx8: np.ndarray) -> np.ndarray:
    """[V,8] = [pos', nrm', uv]: the reference's vertex-buffer record (model.ts:196-200, arrayStride 32 at
    engine.ts:340-347) holding the blend's outputs (engine.ts:270-273: uv is passed through)."""
    uv = np.asarray(vtx8, np.float32).reshape(-1, 8)[:, 6:8]
    return np.concatenate([np.asarray(pos, np.float32), np.asarray(nrm, np.float32), uv], axis=1)


def vertex_major_morphs(V: int, offsets, vertIdx, delta3):
    """morph-major CSR -> vertex-major (start[V+1], morphId[nnz], delta[nnz,3]) keeping PMX morph order."""
    offsets = np.asarray(offsets, np.int64)
    vertIdx = np.asarray(vertIdx, np.int64)
    delta3 = np.asarray(delta3, np.float32).reshape(-1, 3)
    M = max(len(offsets) - 1, 0)
    mid = np.repeat(np.arange(M), np.diff(offsets)) if M else np.zeros(0, np.int64)
    order = np.argsort(vertIdx, kind="stable")
    start = np.zeros(V + 1, np.uint32)
    np.add.at(start, vertIdx + 1, 1)
    start = np.cumsum(start).astype(np.uint32)
    return start, mid[order].astype(np.uint32), np.ascontiguousarray(delta3[order])


def deform(vtx8, joints, weights, skin16, morph=None, morphW=None, sdef=None, dtype=np.float32):
    """One instance through the reference blend (+ morph / SDEF extensions).
    morph = (offsets, vertIdx, delta3) morph-major; morphW = dense [M]; sdef = (vertIdx, vec9)."""
    vtx8 = np.ascontiguousarray(vtx8, np.float32).reshape(-1, 8)
    V = vtx8.shape[0]
    joints = np.ascontiguousarray(joints, np.uint16).reshape(-1)
    weights = np.ascontiguousarray(weights, np.uint8).reshape(-1)
    skin16 = np.ascontiguousarray(skin16, dtype).reshape(-1)
    vs = vm = vd = mw = None
    if morph is not None and morphW is not None and len(morph[0]) > 1:
        vs, vm, vd = vertex_major_morphs(V, *morph)
        mw = np.ascontiguousarray(morphW, np.float32).reshape(-1)
    so = sv = None
    if sdef is not None and len(sdef[0]):
        so = np.full(V, -1, np.int32)
        keep = [(i, v) for i, v in enumerate(np.asarray(sdef[0]).tolist())
                if weights[v * 4 + 2] == 0 and weights[v * 4 + 3] == 0]
        for i, v in keep:
            so[v] = i
        sv = np.ascontiguousarray(sdef[1], np.float32).reshape(-1)
    pos = np.empty((V, 3), dtype)
    nrm = np.empty((V, 3), dtype)
    fn = lib().orc_deform_range_f32 if dtype == np.float32 else lib().orc_deform_range_f64
    fn(_p(vtx8), _p(joints), _p(weights), C.c_uint32(0), C.c_uint32(V), _p(skin16), _p(vs), _p(vm), _p(vd), _p(mw), _p(so), _p(sv),
       _p(pos), _p(nrm))
    return pos, nrm


def deform_instances(vtx8, joints, weights, world, invBind, inst2pal, K, morph=None, morphW=None, sdef=None, nthreads=1,
                     out: Optional[np.ndarray] = None, ring: bool = False):
    """K instances, f32, multi-threaded (CPU baseline).  Returns out [K, 2, Vpad4*3] f32 view helpers."""
    vtx8 = np.ascontiguousarray(vtx8, np.float32).reshape(-1, 8)
    V = vtx8.shape[0]
    joints = np.ascontiguousarray(joints, np.uint16).reshape(-1)
    weights = np.ascontiguousarray(weights, np.uint8).reshape(-1)
    world = np.ascontiguousarray(world, np.float32)
    invBind = np.ascontiguousarray(invBind, np.float32).reshape(-1)
    B = invBind.size // 16
    i2p = None if inst2pal is None else np.ascontiguousarray(inst2pal, np.uint32)
    vs = vm = vd = mw = None
    M = 0
    if morph is not None and morphW is not None and len(morph[0]) > 1:
        vs, vm, vd = vertex_major_morphs(V, *morph)
        mw = np.ascontiguousarray(morphW, np.float32)
        M = mw.shape[-1]
    so = sv = None
    if sdef is not None and len(sdef[0]):
        so = np.full(V, -1, np.int32)
        for i, v in enumerate(np.asarray(sdef[0]).tolist()):
            if weights[v * 4 + 2] == 0 and weights[v * 4 + 3] == 0:
                so[v] = i
        sv = np.ascontiguousarray(sdef[1], np.float32).reshape(-1)
    nrmOff = (V * 3 + 3) // 4 * 4
    stride = 2 * nrmOff
    if out is None:
        out = np.empty((max(nthreads, 1) if ring else K, stride), np.float32)
    lib().orc_deform_instances(_p(vtx8), _p(joints), _p(weights), C.c_uint32(V), C.c_uint32(B), _p(world), _p(invBind), _p(i2p),
                               C.c_uint32(K), _p(vs), _p(vm), _p(vd), _p(mw), C.c_uint32(M), _p(so), _p(sv), _p(out),
                               C.c_size_t(stride), C.c_size_t(nrmOff), C.c_uint32(nthreads), C.c_int32(1 if ring else 0))
    return out, nrmOff


def quantize_bdef2(w0f: float):
    out = np.zeros(2, np.uint8)
    lib().orc_quantize_bdef2(C.c_float(w0f), _p(out))
    return out


def quantize_bdef4(wf):
    w = np.ascontiguousarray(wf, np.float32)
    out = np.zeros(4, np.uint8)
    lib().orc_quantize_bdef4(_p(w), _p(out))
    return out


def finalize_skinning(joints, weights, boneCount: int):
    j = np.array(joints, np.uint16).reshape(-1).copy()
    w = np.array(weights, np.uint8).reshape(-1).copy()
    lib().orc_finalize_skinning(_p(j), _p(w), C.c_uint32(j.size // 4), C.c_uint32(boneCount))
    return j, w


def inverse_bind(parent, bindTranslation) -> np.ndarray:
    parent = np.ascontiguousarray(parent, np.int32)
    bt = np.ascontiguousarray(bindTranslation, np.float64).reshape(-1)
    B = parent.size
    out = np.empty(B * 16, np.float32)
    lib().orc_inverse_bind(_p(parent), _p(bt), C.c_uint32(B), _p(out))
    return out
