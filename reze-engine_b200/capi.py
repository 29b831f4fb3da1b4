"""ctypes binding of the C ABI in `include/rze_b200.h` (librze_b200.so).

This is the Python stand-in for the N-API addon a Node host would use (the same
marshalling): numpy arrays in, status codes turned into
exceptions like the reference's `throw new Error(...)` (engine.ts:161,167,1828).
There is no fallback: if the shared library is missing or no CUDA device can run
it, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "librze_b200.so")

RZ_FLAG_SDEF = 0x1
RZ_FLAG_NO_NORMALS = 0x2
RZ_FLAG_BOUNDS = 0x4
RZ_FLAG_REORDER_VERTICES = 0x8
RZ_FLAG_OUTLINE = 0x10
RZ_FLAG_INTERLEAVED = 0x20
RZ_FLAG_DOUBLE_BUFFER = 0x40
RZ_NO_ATTRIBUTE = (1 << (8 * C.sizeof(C.c_size_t))) - 1

EXPORTS = [
    "rz_create", "rz_destroy", "rz_abi_version", "rz_load_mesh", "rz_load_morphs", "rz_load_sdef",
    "rz_set_palettes", "rz_set_palettes_device", "rz_palette_staging", "rz_load_skeleton", "rz_set_local_rotations",
    "rz_set_tweens", "rz_set_instance_clocks", "rz_load_animation", "rz_set_morph_weights", "rz_deform",
    "rz_sync", "rz_output_device_ptr", "rz_read_instance", "rz_get_vertex_order", "rz_plan_lanes", "rz_read_bounds", "rz_read_skinning",
    "rz_read_skin_matrices", "rz_get_stats", "rz_last_error",
    "rz_load_edge_size", "rz_get_output_layout", "rz_read_outline", "rz_read_interleaved",
    "rz_plan_morph_rows", "rz_plan_chunks", "rz_read_instance_async", "rz_read_wait",
    "rz_load_rigid_bodies", "rz_apply_body_transforms", "rz_plan_sdef", "rz_plan_palette_rows", "rz_plan_lanes2",
    "rz_read_world_matrices",
]


class RzError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"rze_b200 status {status}: {message}")
        self.status = status


class RzConfig(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("device", C.c_int32), ("max_instances", C.c_uint32), ("flags", C.c_uint32),
        ("stream", C.c_void_p),
        ("tune_instances_per_group", C.c_uint32), ("tune_store_mode", C.c_uint32), ("tune_threads", C.c_uint32),
        ("tune_chunks", C.c_uint32), ("tune_ctas_per_sm", C.c_uint32), ("tune_vertices_per_lane", C.c_uint32),
        ("tune_reserved", C.c_uint32 * 2),
    ]


class RzStats(C.Structure):
    _fields_ = [
        ("fps", C.c_double), ("frameTime", C.c_double), ("gpuMemory", C.c_double),
        ("vertsPerSec", C.c_double), ("algorithmicBytes", C.c_double), ("achievedGBs", C.c_double),
        ("lastDeformMs", C.c_double), ("frames", C.c_uint64), ("kernelLaunches", C.c_uint64),
        ("vertexCount", C.c_uint32), ("boneCount", C.c_uint32), ("instanceCount", C.c_uint32), ("paletteCount", C.c_uint32),
        ("morphCount", C.c_uint32), ("morphNnz", C.c_uint32), ("sdefCount", C.c_uint32), ("activeMorphs", C.c_uint32),
        ("instancesPerGroup", C.c_uint32), ("storeMode", C.c_uint32), ("ctas", C.c_uint32), ("threads", C.c_uint32),
        ("smemBytes", C.c_uint32), ("verticesPerLane", C.c_uint32), ("fastGatherPermille", C.c_uint32),
    ]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class RzOutputLayout(C.Structure):
    _fields_ = [("base", C.c_void_p), ("instanceStride", C.c_size_t), ("vertexStride", C.c_size_t), ("positionOffset", C.c_size_t),
                ("normalOffset", C.c_size_t), ("hullOffset", C.c_size_t), ("uvOffset", C.c_size_t)]


_lib = None


def load_library(path: Optional[str] = None) -> C.CDLL:
    """dlopen librze_b200.so and declare every prototype.  Raises if the library was not built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RzError(-2, f"{p} not found: build it with `python -m reze_engine_b200.build` (there is no CPU fallback)")
    lib = C.CDLL(p)
    vp, u32, i32, sz = C.c_void_p, C.c_uint32, C.c_int32, C.c_size_t
    P = C.POINTER
    lib.rz_create.argtypes = [P(RzConfig), P(vp)]
    lib.rz_destroy.argtypes = [vp]
    lib.rz_abi_version.restype = u32
    lib.rz_load_mesh.argtypes = [vp, vp, vp, vp, u32, vp, u32]
    lib.rz_load_morphs.argtypes = [vp, vp, vp, vp, u32]
    lib.rz_load_sdef.argtypes = [vp, vp, vp, u32]
    lib.rz_set_palettes.argtypes = [vp, vp, u32, vp, u32]
    lib.rz_set_palettes_device.argtypes = [vp, vp, u32, vp, u32]
    lib.rz_palette_staging.argtypes = [vp, sz, P(vp)]
    lib.rz_load_skeleton.argtypes = [vp, vp, vp, vp, vp, vp, u32]
    lib.rz_set_local_rotations.argtypes = [vp, vp, u32, vp, u32]
    lib.rz_set_tweens.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    lib.rz_set_instance_clocks.argtypes = [vp, vp, u32, vp, u32]
    lib.rz_load_animation.argtypes = [vp, vp, vp, vp, vp]
    lib.rz_set_morph_weights.argtypes = [vp, vp, vp, u32, u32]
    lib.rz_deform.argtypes = [vp, u32, u32]
    lib.rz_sync.argtypes = [vp]
    lib.rz_output_device_ptr.argtypes = [vp, P(vp), P(sz), P(sz)]
    lib.rz_read_instance.argtypes = [vp, u32, vp, vp]
    lib.rz_load_rigid_bodies.argtypes = [vp, vp, vp, vp, u32]
    lib.rz_apply_body_transforms.argtypes = [vp, vp, u32]
    lib.rz_read_instance_async.argtypes = [vp, u32, vp, vp]
    lib.rz_read_wait.argtypes = [vp]
    lib.rz_load_edge_size.argtypes = [vp, vp]
    lib.rz_get_output_layout.argtypes = [vp, P(RzOutputLayout)]
    lib.rz_read_outline.argtypes = [vp, u32, vp]
    lib.rz_read_interleaved.argtypes = [vp, u32, vp]
    lib.rz_get_vertex_order.argtypes = [vp, vp]
    lib.rz_plan_lanes.argtypes = [vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, vp, vp, vp, vp]
    lib.rz_plan_morph_rows.argtypes = [vp, u32, u32, vp, vp, vp, u32, vp, vp, vp, vp, C.c_uint64, P(C.c_uint64)]
    lib.rz_plan_chunks.argtypes = [vp, u32, u32, u32, vp, P(u32)]
    lib.rz_plan_palette_rows.argtypes = [vp, u32, u32, vp]
    lib.rz_plan_lanes2.argtypes = [vp, vp, u32, u32, u32, vp, vp, vp, vp, vp, vp, vp, vp, vp, P(u32), vp, vp, vp]
    lib.rz_plan_sdef.argtypes = [vp, u32, vp, vp, u32, u32, vp, vp, u32, vp, vp, P(u32)]
    lib.rz_read_bounds.argtypes = [vp, u32, u32, vp]
    lib.rz_read_skinning.argtypes = [vp, vp, vp]
    lib.rz_read_world_matrices.argtypes = [vp, C.c_uint32, vp]
    lib.rz_read_skin_matrices.argtypes = [vp, u32, vp]
    lib.rz_get_stats.argtypes = [vp, P(RzStats)]
    lib.rz_last_error.argtypes = [vp]
    lib.rz_last_error.restype = C.c_char_p
    for name in EXPORTS:
        fn = getattr(lib, name)
        if name not in ("rz_abi_version", "rz_last_error"):
            fn.restype = i32
    if path is None:
        _lib = lib
    return lib


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _arr(a, dtype) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=dtype)


def plan_lanes(joints, weights, B: int, mode: int = 2, lib: Optional[C.CDLL] = None) -> dict:
    """rz_plan_lanes: the load-time lane / influence-slot plan for a skinning table (host only, no device needed)."""
    lib = lib or load_library()
    j, w = _arr(joints, np.uint16).reshape(-1, 4), _arr(weights, np.uint8).reshape(-1, 4)
    V = j.shape[0]
    Vp = (V + 255) // 256 * 256
    lane_vertex = np.empty(Vp, np.uint32)
    lane_joints = np.empty((Vp, 4), np.uint16)
    lane_weights = np.empty((Vp, 4), np.float32)
    stats = np.zeros(27, np.uint64)
    st = lib.rz_plan_lanes(_ptr(j), _ptr(w), V, B, mode, _ptr(lane_vertex), _ptr(lane_joints), _ptr(lane_weights), _ptr(stats))
    if st != 0:
        raise RzError(st, lib.rz_last_error(None).decode("utf-8", "replace"))
    return {"laneVertex": lane_vertex, "laneJoints": lane_joints, "laneWeights": lane_weights, "fast": int(stats[0]), "total": int(stats[1]),
            "hist": stats[2:].reshape(5, 5).astype(np.int64)}


def plan_morph_rows(lane_vertex, V: int, offsets, vert_idx, delta3, lib: Optional[C.CDLL] = None) -> dict:
    """Device-free: the per-warp morph rows rz_load_morphs builds for the lane plan `lane_vertex` (rz_plan_morph_rows)."""
    lib = lib or load_library()
    lv = _arr(lane_vertex, np.uint32).reshape(-1)
    off = _arr(offsets, np.uint32).reshape(-1)
    vi = _arr(vert_idx, np.uint32).reshape(-1)
    d3 = _arr(delta3, np.float32).reshape(-1)
    Vp, M = lv.size, max(off.size - 1, 0)
    first = np.zeros(Vp // 32, np.uint32)
    depth = np.zeros(Vp // 32, np.uint32)
    mm = np.zeros(Vp // 32, np.uint8)
    need = C.c_uint64(0)
    st = lib.rz_plan_morph_rows(_ptr(lv), Vp, V, _ptr(off), _ptr(vi), _ptr(d3), M, _ptr(first), _ptr(depth), _ptr(mm), None, 0, C.byref(need))
    if st != 0:
        raise RzError(st, (lib.rz_last_error(None) or b"").decode())
    rows = np.zeros((need.value, 4), np.float32)
    st = lib.rz_plan_morph_rows(_ptr(lv), Vp, V, _ptr(off), _ptr(vi), _ptr(d3), M, None, None, None, _ptr(rows), need.value, None)
    if st != 0:
        raise RzError(st, (lib.rz_last_error(None) or b"").decode())
    return dict(first=first, depth=depth, morphMajor=mm, rows=rows)


def plan_lanes2(joints, weights, B: int, lib: Optional[C.CDLL] = None) -> dict:
    """rz_plan_lanes2 (the two-vertices-per-lane kernel's lane plan, device-free): groups of 32 lanes covering up to 64 vertices."""
    lib = lib or load_library()
    j, w = _arr(joints, np.uint16).reshape(-1, 4), _arr(weights, np.uint8).reshape(-1, 4)
    V = j.shape[0]
    n = C.c_uint32(0)
    stats = np.zeros(4, np.uint64)
    st = lib.rz_plan_lanes2(_ptr(j), _ptr(w), V, B, 0, None, None, None, None, None, None, None, None, _ptr(stats), C.byref(n), None, None, None)
    if st != 0:
        raise RzError(st, (lib.rz_last_error(None) or b"").decode())
    G = n.value
    out = dict(groupFirst=np.zeros(G, np.uint32), groupCount=np.zeros(G, np.uint32), groupPaired=np.zeros(G, np.uint8),
               vertA=np.zeros(G * 32, np.uint32), vertB=np.zeros(G * 32, np.uint32), laneJoints=np.zeros((G * 32, 4), np.uint16),
               wA=np.zeros((G * 32, 4), np.float32), wB=np.zeros((G * 32, 4), np.float32),
               slotA=np.zeros(G * 32, np.uint8), slotB=np.zeros(G * 32, np.uint8), laneSlots=np.zeros(G * 32, np.uint8))
    st = lib.rz_plan_lanes2(_ptr(j), _ptr(w), V, B, G, _ptr(out["groupFirst"]), _ptr(out["groupCount"]), _ptr(out["groupPaired"]),
                            _ptr(out["vertA"]), _ptr(out["vertB"]), _ptr(out["laneJoints"]), _ptr(out["wA"]), _ptr(out["wB"]), _ptr(stats),
                            C.byref(n), _ptr(out["slotA"]), _ptr(out["slotB"]), _ptr(out["laneSlots"]))
    if st != 0:
        raise RzError(st, (lib.rz_last_error(None) or b"").decode())
    out.update(fast=int(stats[0]), total=int(stats[1]), pairedWindows=int(stats[2]), fallbackWindows=int(stats[3]))
    return out


def plan_palette_rows(lane_joints, B: int, lib: Optional[C.CDLL] = None) -> np.ndarray:
    """Device-free: the bank-aware palette permutation for the gather table `lane_joints` [Vp,4] (rz_plan_palette_rows)."""
    lib = lib or load_library()
    lj = _arr(lane_joints, np.uint16).reshape(-1)
    pos = np.zeros(B, np.uint32)
    st = lib.rz_plan_palette_rows(_ptr(lj), lj.size // 4, B, _ptr(pos))
    if st != 0:
        raise RzError(st, (lib.rz_last_error(None) or b"").decode())
    return pos


def plan_sdef(lane_vertex, joints, weights, B: int, sdef_vert_idx, c_r0_r1, lib: Optional[C.CDLL] = None) -> dict:
    """Device-free: SDEF records [n,12] and per-lane descriptor words for the lane plan `lane_vertex` (rz_plan_sdef)."""
    lib = lib or load_library()
    lv = _arr(lane_vertex, np.uint32).reshape(-1)
    j, w = _arr(joints, np.uint16).reshape(-1), _arr(weights, np.uint8).reshape(-1)
    vi = _arr(sdef_vert_idx, np.uint32).reshape(-1)
    vec = _arr(c_r0_r1, np.float32).reshape(-1)
    rec = np.zeros((max(vi.size, 1), 12), np.float32)
    desc = np.zeros(lv.size, np.uint32)
    n = C.c_uint32(0)
    st = lib.rz_plan_sdef(_ptr(lv), lv.size, _ptr(j), _ptr(w), j.size // 4, B, _ptr(vi), _ptr(vec), vi.size, _ptr(rec), _ptr(desc), C.byref(n))
    if st != 0:
        raise RzError(st, (lib.rz_last_error(None) or b"").decode())
    return dict(records=rec[:n.value], desc=desc, active=n.value)


def plan_chunks(tile_depth, tiles_per_pass: int, n_chunks_target: int, lib: Optional[C.CDLL] = None) -> np.ndarray:
    """Device-free: cost-balanced chunk boundaries (rz_plan_chunks), in tiles."""
    lib = lib or load_library()
    td = _arr(tile_depth, np.uint32).reshape(-1)
    tab = np.zeros(max(n_chunks_target, 1) + 1, np.uint32)
    n = C.c_uint32(0)
    st = lib.rz_plan_chunks(_ptr(td), td.size, tiles_per_pass, n_chunks_target, _ptr(tab), C.byref(n))
    if st != 0:
        raise RzError(st, (lib.rz_last_error(None) or b"").decode())
    return tab[:n.value + 1].copy()


class DeformContext:
    """One rz_ctx: one GPU, one mesh, K instances."""

    def __init__(self, max_instances: int = 1, device: int = 0, flags: int = 0, stream: int = 0,
                 instances_per_group: int = 0, store_mode: int = 0, threads: int = 0, chunks: int = 0, ctas_per_sm: int = 0,
                 vertices_per_lane: int = 0):
        self.lib = load_library()
        cfg = RzConfig()
        cfg.struct_size = C.sizeof(RzConfig)
        cfg.device = device
        cfg.max_instances = max_instances
        cfg.flags = flags
        cfg.stream = stream or None
        cfg.tune_instances_per_group = instances_per_group
        cfg.tune_store_mode = store_mode
        cfg.tune_threads = threads
        cfg.tune_chunks = chunks
        cfg.tune_ctas_per_sm = ctas_per_sm
        cfg.tune_vertices_per_lane = vertices_per_lane
        h = C.c_void_p()
        st = self.lib.rz_create(C.byref(cfg), C.byref(h))
        if st != 0:
            raise RzError(st, (self.lib.rz_last_error(None) or b"").decode())
        self.h = h
        self.max_instances = max_instances
        self.flags = flags
        self.V = 0
        self.B = 0
        self._stage = None

    # -- plumbing
    def _check(self, st: int):
        if st != 0:
            raise RzError(st, (self.lib.rz_last_error(self.h) or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.rz_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- static tables
    def load_mesh(self, vtx8, joints, weights, invBind):
        vtx8 = _arr(vtx8, np.float32).reshape(-1)
        joints = _arr(joints, np.uint16).reshape(-1)
        weights = _arr(weights, np.uint8).reshape(-1)
        invBind = _arr(invBind, np.float32).reshape(-1)
        V, B = vtx8.size // 8, invBind.size // 16
        if joints.size != V * 4 or weights.size != V * 4:
            raise ValueError("joints/weights must hold 4 entries per vertex")
        self._check(self.lib.rz_load_mesh(self.h, _ptr(vtx8), _ptr(joints), _ptr(weights), V, _ptr(invBind), B))
        self.V, self.B = V, B

    def load_morphs(self, offsets, vertIdx, delta3):
        offsets = _arr(offsets, np.uint32).reshape(-1)
        vertIdx = _arr(vertIdx, np.uint32).reshape(-1)
        delta3 = _arr(delta3, np.float32).reshape(-1)
        M = max(offsets.size - 1, 0)
        self._check(self.lib.rz_load_morphs(self.h, _ptr(offsets) if M else None, _ptr(vertIdx) if vertIdx.size else None,
                                            _ptr(delta3) if delta3.size else None, M))

    def load_sdef(self, vertIdx, c_r0_r1):
        vertIdx = _arr(vertIdx, np.uint32).reshape(-1)
        vec = _arr(c_r0_r1, np.float32).reshape(-1)
        self._check(self.lib.rz_load_sdef(self.h, _ptr(vertIdx) if vertIdx.size else None, _ptr(vec) if vec.size else None, vertIdx.size))

    # -- per frame
    def palette_staging(self, P: int) -> np.ndarray:
        """Pinned host buffer shaped [P, B, 16] the caller fills and hands to set_palettes.  Call it before EVERY refill: the
        library alternates between two buffers and blocks until the upload that last read the returned one has completed."""
        nbytes = P * self.B * 64
        p = C.c_void_p()
        self._check(self.lib.rz_palette_staging(self.h, nbytes, C.byref(p)))
        buf = (C.c_float * (P * self.B * 16)).from_address(p.value)
        return np.frombuffer(buf, dtype=np.float32).reshape(P, self.B, 16)

    def rotation_staging(self, P: int) -> np.ndarray:
        """Pinned host buffer shaped [P, B, 4] (xyzw per bone) for set_local_rotations; same two-buffer protocol as
        palette_staging: ask before every refill."""
        nbytes = P * self.B * 16
        p = C.c_void_p()
        self._check(self.lib.rz_palette_staging(self.h, nbytes, C.byref(p)))
        buf = (C.c_float * (P * self.B * 4)).from_address(p.value)
        return np.frombuffer(buf, dtype=np.float32).reshape(P, self.B, 4)

    def set_palettes(self, world, inst_to_palette=None, K: Optional[int] = None):
        world = np.asarray(world)
        if world.dtype != np.float32 or not world.flags.c_contiguous:
            world = _arr(world, np.float32)
        P = world.size // (self.B * 16)
        i2p = None if inst_to_palette is None else _arr(inst_to_palette, np.uint32).reshape(-1)
        if K is None:
            K = P if i2p is None else i2p.size
        self._check(self.lib.rz_set_palettes(self.h, _ptr(world), P, _ptr(i2p), K))
        self.K = K

    def set_palettes_device(self, d_world_ptr: int, P: int, d_inst_to_palette_ptr: int = 0, K: Optional[int] = None):
        K = P if K is None else K
        self._check(self.lib.rz_set_palettes_device(self.h, C.c_void_p(d_world_ptr), P,
                                                    C.c_void_p(d_inst_to_palette_ptr) if d_inst_to_palette_ptr else None, K))
        self.K = K

    # -- GPU pose evaluation
    def load_skeleton(self, bones):
        """bones: sequence of model.Bone (parentIndex, bindTranslation, append*)."""
        B = len(bones)
        parent = np.asarray([b.parentIndex for b in bones], np.int32)
        bt = np.asarray([b.bindTranslation for b in bones], np.float64).astype(np.float32).reshape(-1)
        ap = np.asarray([-1 if b.appendParentIndex is None else b.appendParentIndex for b in bones], np.int32)
        ar = np.asarray([np.nan if b.appendRatio is None else b.appendRatio for b in bones], np.float32)
        rot = np.asarray([1 if b.appendRotate else 0 for b in bones], np.uint8)
        self._check(self.lib.rz_load_skeleton(self.h, _ptr(parent), _ptr(bt), _ptr(ap), _ptr(ar), _ptr(rot), B))

    def set_local_rotations(self, quats, inst_to_palette=None, K: Optional[int] = None):
        q = np.asarray(quats)
        if q.dtype != np.float32 or not q.flags.c_contiguous:
            q = _arr(q, np.float32)
        P = q.size // (self.B * 4)
        i2p = None if inst_to_palette is None else _arr(inst_to_palette, np.uint32).reshape(-1)
        if K is None:
            K = P if i2p is None else i2p.size
        self._check(self.lib.rz_set_local_rotations(self.h, _ptr(q), P, _ptr(i2p), K))
        self.K = K

    def set_tweens(self, start, target, start_ms, dur_ms, active, rest):
        a = [_arr(start, np.float32), _arr(target, np.float32), _arr(start_ms, np.float32), _arr(dur_ms, np.float32),
             _arr(active, np.uint8), _arr(rest, np.float32)]
        self._check(self.lib.rz_set_tweens(self.h, *[_ptr(x) for x in a]))

    def load_animation(self, key_offsets, key_times_ms, key_quats, rest_quat=None):
        """Keyframe tracks per bone (CSR): see rz_load_animation.  key_offsets=None unloads."""
        if key_offsets is None:
            self._check(self.lib.rz_load_animation(self.h, None, None, None, None))
            return
        off = _arr(key_offsets, np.uint32).reshape(-1)
        t = _arr(key_times_ms, np.float32).reshape(-1)
        q = _arr(key_quats, np.float32).reshape(-1)
        r = None if rest_quat is None else _arr(rest_quat, np.float32).reshape(-1)
        self._check(self.lib.rz_load_animation(self.h, _ptr(off), _ptr(t) if t.size else None, _ptr(q) if q.size else None, _ptr(r)))

    def set_instance_clocks(self, now_ms, inst_to_palette=None, K: Optional[int] = None):
        t = _arr(now_ms, np.float32).reshape(-1)
        i2p = None if inst_to_palette is None else _arr(inst_to_palette, np.uint32).reshape(-1)
        if K is None:
            K = t.size if i2p is None else i2p.size
        self._check(self.lib.rz_set_instance_clocks(self.h, _ptr(t), t.size, _ptr(i2p), K))
        self.K = K

    def load_rigid_bodies(self, bone_index, dynamic, offset_inverse):
        """Rigid bodies that may drive bones (physics.ts:560-585 gives offset_inverse, see physics_bridge.compute_body_offsets)."""
        bi = _arr(bone_index, np.int32).reshape(-1)
        dy = _arr(dynamic, np.uint8).reshape(-1)
        oi = _arr(offset_inverse, np.float32).reshape(-1)
        if dy.size != bi.size or oi.size != bi.size * 16:
            raise ValueError("dynamic / offset_inverse must match bone_index")
        self._n_bodies = bi.size
        self._check(self.lib.rz_load_rigid_bodies(self.h, _ptr(bi), _ptr(dy), _ptr(oi), bi.size))

    def apply_body_transforms(self, pos_quat):
        """Solver output [P, nBodies, 7] (x,y,z, qx,qy,qz,qw) over this frame's palettes (physics.ts:714-751 on the device)."""
        pq = _arr(pos_quat, np.float32)
        n = getattr(self, "_n_bodies", 0)
        P = pq.size // (7 * n) if n else 0
        self._check(self.lib.rz_apply_body_transforms(self.h, _ptr(pq.reshape(-1)), P))

    def set_morph_weights(self, w, active_ids, K: Optional[int] = None):
        ids = _arr(active_ids, np.uint32).reshape(-1)
        w = _arr(w, np.float32)
        Mact = ids.size
        if K is None:
            K = w.size // max(Mact, 1) if Mact else self.max_instances
        self._check(self.lib.rz_set_morph_weights(self.h, _ptr(w) if Mact else None, _ptr(ids) if Mact else None, Mact, K))

    def deform(self, first: int = 0, count: Optional[int] = None):
        self._check(self.lib.rz_deform(self.h, first, self.K - first if count is None else count))

    def sync(self):
        self._check(self.lib.rz_sync(self.h))

    # -- results
    def output_device_ptr(self) -> Tuple[int, int, int]:
        base, stride, noff = C.c_void_p(), C.c_size_t(), C.c_size_t()
        self._check(self.lib.rz_output_device_ptr(self.h, C.byref(base), C.byref(stride), C.byref(noff)))
        return base.value, stride.value, noff.value

    def read_instance(self, inst: int, normals: bool = True, out_pos: Optional[np.ndarray] = None, out_nrm: Optional[np.ndarray] = None):
        """Skinned positions / normals of one instance as [V,3] float32.  `out_pos` / `out_nrm` let the caller supply
        (e.g. pinned) destination arrays."""
        pos = out_pos if out_pos is not None else np.empty((self.V, 3), dtype=np.float32)
        nrm = None
        if normals and not (self.flags & RZ_FLAG_NO_NORMALS):
            nrm = out_nrm if out_nrm is not None else np.empty((self.V, 3), dtype=np.float32)
        self._check(self.lib.rz_read_instance(self.h, inst, _ptr(pos), _ptr(nrm)))
        return pos, nrm

    def read_instance_async(self, inst: int, out_pos: np.ndarray, out_nrm: Optional[np.ndarray] = None):
        """Queue the read-back of one instance behind the work issued so far and return at once; the (preferably pinned)
        arrays are valid after read_wait()."""
        self._check(self.lib.rz_read_instance_async(self.h, inst, _ptr(out_pos), _ptr(out_nrm)))

    def read_wait(self):
        self._check(self.lib.rz_read_wait(self.h))

    def load_edge_size(self, edge_size):
        """Per-vertex Material.edgeSize (0 = no outline) for RZ_FLAG_OUTLINE; None resets to zero."""
        e = None if edge_size is None else _arr(edge_size, np.float32).reshape(-1)
        if e is not None and e.size != self.V:
            raise ValueError("edge_size must hold one entry per vertex")
        self._check(self.lib.rz_load_edge_size(self.h, _ptr(e)))

    def output_layout(self) -> dict:
        lay = RzOutputLayout()
        self._check(self.lib.rz_get_output_layout(self.h, C.byref(lay)))
        return {k: getattr(lay, k) for k, _ in lay._fields_}

    def read_outline(self, inst: int) -> np.ndarray:
        """Outline hull positions pos' + n' * edgeSize * 0.01 of one instance, [V,3] float32 (RZ_FLAG_OUTLINE)."""
        out = np.empty((self.V, 3), dtype=np.float32)
        self._check(self.lib.rz_read_outline(self.h, inst, _ptr(out)))
        return out

    def read_interleaved(self, inst: int) -> np.ndarray:
        """The [x,y,z,nx,ny,nz,u,v] stream of one instance, [V,8] float32 (RZ_FLAG_INTERLEAVED)."""
        out = np.empty((self.V, 8), dtype=np.float32)
        self._check(self.lib.rz_read_interleaved(self.h, inst, _ptr(out)))
        return out

    def vertex_order(self) -> np.ndarray:
        """order[i] = caller vertex id stored at position i of the device planes."""
        o = np.empty(self.V, dtype=np.uint32)
        self._check(self.lib.rz_get_vertex_order(self.h, _ptr(o)))
        return o

    def read_bounds(self, first: int, count: int) -> np.ndarray:
        out = np.empty((count, 6), dtype=np.float32)
        self._check(self.lib.rz_read_bounds(self.h, first, count, _ptr(out)))
        return out

    def read_skinning(self):
        j = np.zeros(self.V * 4, dtype=np.uint16)
        w = np.zeros(self.V * 4, dtype=np.uint8)
        self._check(self.lib.rz_read_skinning(self.h, _ptr(j), _ptr(w)))
        return j, w

    def read_skin_matrices(self, palette: int) -> np.ndarray:
        out = np.empty((self.B, 12), dtype=np.float32)
        self._check(self.lib.rz_read_skin_matrices(self.h, palette, _ptr(out)))
        return out

    def read_world_matrices(self, palette: int) -> np.ndarray:
        """[B,16] column-major bone world matrices of a palette (the layout of Model.getBoneWorldMatrices())."""
        out = np.empty((self.B, 16), np.float32)
        self._check(self.lib.rz_read_world_matrices(self.h, palette, _ptr(out)))
        return out

    def stats(self) -> dict:
        s = RzStats()
        self._check(self.lib.rz_get_stats(self.h, C.byref(s)))
        return s.asdict()
