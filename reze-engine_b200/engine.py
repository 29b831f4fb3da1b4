"""Engine facade: the reference's public API surface over the B200 deform path.

Keeps the method names and call order of `engine/src/engine.ts`
(`new Engine(canvas, options)` -> `init()` -> `loadModel(path)` ->
`loadAnimation(url)` -> `runRenderLoop(cb)` -> `playAnimation(opts)`;
`rotateBones(names, quats, durationMs)`, `stopAnimation()`, `getStats()`,
`dispose()`; engine.ts:145-157, 1419-1425, 1593, 1664-1725).  What it drives is
only the stage this repo replaces: pose evaluation on the host (as the reference
does, model.ts) -> palette upload + skin matrices + fused morph/skin kernel on the
GPU (replacing engine.ts:2375-2402 and the vertex-shader blend engine.ts:245-276).
Rasterisation, camera, bloom, physics are out of scope.

Differences a caller sees, all forced by leaving the browser:
* no canvas: the first constructor argument is accepted and ignored;
* `init/loadModel/loadAnimation` are plain (synchronous) methods;
* time is injected: `clock` (ms) replaces `performance.now()` (model.ts:160,249)
  and `window.setTimeout` (engine.ts:1547,1587,1657) is an internal timer queue
  pumped by `render()`; `ManualClock` gives reproducible playback;
* `instances=K` deforms a crowd of K independent copies of the model; every
  reference method addresses instance 0 unless `instance=` is given.
"""
from __future__ import annotations

import heapq
import itertools
import time
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Union

import numpy as np

from . import capi
from .math3d import Quat, Vec3
from .model import Model
from .pmx import PmxLoader
from .vmd import VMDKeyFrame, VMDLoader


@dataclass
class EngineStats:
    """engine.ts:16-20 plus deform-stage counters."""
    fps: float = 0.0
    frameTime: float = 0.0   # ms
    gpuMemory: float = 0.0   # MB
    vertsPerSec: float = 0.0
    achievedGBs: float = 0.0
    algorithmicBytes: float = 0.0


class ManualClock:
    """Deterministic clock in milliseconds for tests and offline playback."""

    def __init__(self, start_ms: float = 0.0):
        self.now_ms = float(start_ms)

    def __call__(self) -> float:
        return self.now_ms

    def advance(self, ms: float) -> float:
        self.now_ms += ms
        return self.now_ms


class Engine:
    def __init__(self, canvas=None, options: Optional[dict] = None, *, instances: int = 1, device: int = 0,
                 clock: Optional[Callable[[], float]] = None, sdef: bool = False, bounds: bool = False, stream: int = 0,
                 gpu_pose: bool = False, crowd: bool = False, reorder_vertices: bool = False, outline: bool = False,
                 interleaved: bool = False, double_buffer: bool = False):
        o = options or {}
        # EngineOptions (engine.ts:8-14): kept so existing call sites construct unchanged; they only
        # parameterise passes this repo does not replace.
        self.ambient = o.get("ambient", 1.0)
        self.bloomIntensity = o.get("bloomIntensity", 0.12)
        self.rimLightIntensity = o.get("rimLightIntensity", 0.45)
        self.cameraDistance = o.get("cameraDistance", 26.6)
        self.cameraTarget = o.get("cameraTarget", Vec3(0, 12.5, 0))
        self.instances = int(instances)
        self.device = device
        self.clock = clock or (lambda: time.perf_counter() * 1000.0)
        # reorder_vertices: the device planes store vertices grouped by bone tuple (faster blend); draw with deviceIndexBuffer()
        self._flags = ((capi.RZ_FLAG_SDEF if sdef else 0) | (capi.RZ_FLAG_BOUNDS if bounds else 0)
                       | (capi.RZ_FLAG_REORDER_VERTICES if reorder_vertices else 0)
                       # fused consumers of the skinned stream: the outline pass' hull positions
                       # (engine.ts:431-463) as a third plane / the result in the reference's 32-byte vertex layout
                       | (capi.RZ_FLAG_OUTLINE if outline else 0) | (capi.RZ_FLAG_INTERLEAVED if interleaved else 0)
                       # two result buffers: frame n stays readable (readSkinnedAsync) while render() produces frame n+1
                       | (capi.RZ_FLAG_DOUBLE_BUFFER if double_buffer else 0))
        self._stream = stream
        self.gpu_pose = gpu_pose or crowd   # walk the bone hierarchy on the GPU (rz_set_local_rotations) instead of in Model
        # crowd mode: ONE shared skeleton runtime + animation clip, every instance plays it at its own clock offset; tweens /
        # keyframe tracks are evaluated on the device (rz_set_tweens / rz_load_animation + rz_set_instance_clocks)
        self.crowd = crowd
        self._crowd_playing = False
        self._anim_start_ms = 0.0
        self._offsets_ms = np.zeros(int(instances), np.float64)
        self.ctx: Optional[capi.DeformContext] = None
        self.currentModel: Optional[Model] = None
        self.models: List[Model] = []
        self.animationFrames: List[VMDKeyFrame] = []
        self.hasAnimation = False
        self.playingAnimation = False
        self._timers: list = []
        self._timer_ids = itertools.count(1)
        self._cancelled: set = set()
        self._animationTimeouts: List[int] = []
        self._breathingTimeout: Optional[int] = None
        self._breathingBase: Dict[str, Quat] = {}
        self._loop_running = False
        self._renderLoopCallback = None
        self._stats = EngineStats()
        self._morph_ids = np.zeros(0, np.uint32)
        self._morphTracks: Dict[str, list] = {}
        self._morphPlaying = False

    # ---- lifecycle ---------------------------------------------------------------------------
    def init(self):
        """engine.ts:157-185: acquire the device.  Raises if no sm_100 GPU / library (no fallback)."""
        self.ctx = capi.DeformContext(max_instances=self.instances, device=self.device, flags=self._flags, stream=self._stream)
        return self

    def dispose(self):
        """engine.ts:1692-1701."""
        self.stopRenderLoop()
        self.stopAnimation()
        self._stopBreathing()
        if self.ctx:
            self.ctx.close()
            self.ctx = None

    # ---- timers (window.setTimeout stand-in) ------------------------------------------------
    def _setTimeout(self, fn: Callable[[], None], delayMs: float) -> int:
        tid = next(self._timer_ids)
        heapq.heappush(self._timers, (self.clock() + max(0.0, delayMs), tid, fn))
        return tid

    def _clearTimeout(self, tid: Optional[int]):
        if tid is not None:
            self._cancelled.add(tid)

    def _pumpTimers(self):
        now = self.clock()
        while self._timers and self._timers[0][0] <= now:
            _, tid, fn = heapq.heappop(self._timers)
            if tid in self._cancelled:
                self._cancelled.discard(tid)
                continue
            fn()

    # ---- assets ------------------------------------------------------------------------------
    def loadModel(self, path: Union[str, Model]):
        """engine.ts:1704-1721 -> setupModelBuffers (engine.ts:1728-1832): parse, then upload the static
        tables through the C ABI.  Instances share the mesh; each gets its own skeleton runtime."""
        if self.ctx is None:
            raise RuntimeError("Engine.init() must be called before loadModel")
        model = path if isinstance(path, Model) else PmxLoader.load(path, clock=self.clock)
        model.clock = self.clock
        self.currentModel = model
        self.models = [model]
        for _ in range(1, 1 if self.crowd else self.instances):
            self.models.append(Model(model.vertexData, model.indexData, model.textures, model.materials, model.skeleton,
                                     model.skinning, morphs=model.morphs, sdef=model.sdef, clock=self.clock))
        sk = model.getSkinning()
        self.ctx.load_mesh(model.getVertices(), sk.joints, sk.weights, model.getBoneInverseBindMatrices())
        if model.morphs.count:
            self.ctx.load_morphs(model.morphs.offsets, model.morphs.vertexIndex, model.morphs.delta)
        if model.sdef.vertexIndex.size:
            self.ctx.load_sdef(model.sdef.vertexIndex, model.sdef.c_r0_r1)
        if self._flags & capi.RZ_FLAG_OUTLINE:
            self.ctx.load_edge_size(self.vertexEdgeSizes(model))
        if self.gpu_pose:
            self.ctx.load_skeleton(model.getSkeleton().bones)
            self._rot_stage = np.zeros((self.instances, len(model.getSkeleton().bones), 4), np.float32)
        return model

    def loadAnimation(self, url: str):
        """engine.ts:1419-1423.  Extension: the clip's vertex-morph track (which the reference never reads,
        vmd-loader.ts stops after the bone table) is kept and played back by playAnimation."""
        self.animationFrames, morphFrames = VMDLoader.loadWithMorphs(url)
        self.hasAnimation = True
        self._morphTracks = {}
        for mf in morphFrames:
            self._morphTracks.setdefault(mf.morphName, []).append((mf.frame / 30.0 * 1000.0, float(mf.weight)))
        for v in self._morphTracks.values():
            v.sort(key=lambda tw: tw[0])

    def _evalMorphTracks(self, tau_ms: np.ndarray):
        """Morph weights [K, n] at per-instance clip times (MMD rule: linear interpolation between keys, held outside)."""
        names = self.currentModel.morphs.names
        ids, cols = [], []
        for name, keys in self._morphTracks.items():
            if name not in names:
                continue
            t = np.asarray([k[0] for k in keys], np.float64)
            w = np.asarray([k[1] for k in keys], np.float64)
            ids.append(names.index(name))
            cols.append(np.interp(tau_ms, t, w))
        if not ids:
            return None, None
        return np.asarray(ids, np.uint32), np.stack(cols, axis=1).astype(np.float32)

    # ---- bone API ----------------------------------------------------------------------------
    def rotateBones(self, bones: Sequence[str], rotations: Sequence[Quat], durationMs: Optional[float] = None,
                    instance: Optional[int] = 0):
        """engine.ts:1723-1725 -> model.ts:246-315.  instance=None addresses every instance."""
        targets = self.models if instance is None else self.models[instance:instance + 1]
        for m in targets:
            m.rotateBones(bones, rotations, durationMs)

    def setMorphWeights(self, weights, morph_names_or_ids: Sequence[Union[str, int]]):
        """New: per-instance weights [K, len(ids)] of the named vertex morphs."""
        names = self.currentModel.morphs.names
        ids = [names.index(x) if isinstance(x, str) else int(x) for x in morph_names_or_ids]
        w = np.ascontiguousarray(weights, dtype=np.float32).reshape(self.instances, len(ids))
        self.ctx.set_morph_weights(w, np.asarray(ids, np.uint32), K=self.instances)

    # ---- animation playback (engine.ts:1425-1662) ----------------------------------------------
    def setInstanceOffsets(self, offsets_ms):
        """Crowd mode: instance k plays the shared animation `offsets_ms[k]` milliseconds behind instance 0."""
        self._offsets_ms = np.ascontiguousarray(offsets_ms, dtype=np.float64).reshape(self.instances)

    def _playCrowd(self):
        """Crowd playback: the clip becomes per-bone keyframe tracks on the device (same rule as the timers below
        produce when they fire on time: key at t=0 instant, others tween from the previous key with ease + slerp)."""
        bones = self.currentModel.getSkeleton().bones
        idx = self.currentModel.nameIndex
        per: List[list] = [[] for _ in bones]
        for kf in self.animationFrames:
            for bf in kf.boneFrames:
                i = idx.get(bf.boneName, -1)
                if i >= 0:
                    per[i].append((kf.time * 1000.0, bf.rotation))
        off = np.zeros(len(bones) + 1, np.uint32)
        times, quats = [], []
        for i, keys in enumerate(per):
            keys.sort(key=lambda tq: tq[0])
            off[i + 1] = off[i] + len(keys)
            for t, q in keys:
                times.append(t)
                quats.append(q.toArray())
        self.ctx.load_animation(off, np.asarray(times, np.float32), np.asarray(quats, np.float32).reshape(-1, 4))
        self._anim_start_ms = self.clock()
        self._crowd_playing = True

    def playAnimation(self, options: Optional[dict] = None, instance: Optional[int] = 0):
        if not self.animationFrames:
            return
        self.stopAnimation()
        self._stopBreathing()
        self.playingAnimation = True
        self._anim_start_ms = self.clock()
        self._morphPlaying = bool(self._morphTracks) and self.currentModel is not None and self.currentModel.morphs.count > 0
        if self.crowd:
            return self._playCrowd()
        opts = options or {}
        bb = opts.get("breathBones")
        enableBreath = bb is not None
        breathBones: List[str] = []
        breathRanges: Optional[Dict[str, float]] = None
        if enableBreath and bb:
            if isinstance(bb, dict):
                breathBones, breathRanges = list(bb.keys()), dict(bb)
            else:
                breathBones = list(bb)
        breathDuration = opts.get("breathDuration", 4000)

        byBone: Dict[str, list] = {}
        for kf in self.animationFrames:
            for bf in kf.boneFrames:
                byBone.setdefault(bf.boneName, []).append((kf.time, bf.rotation))
        for v in byBone.values():
            v.sort(key=lambda tr: tr[0])

        rot = lambda names, quats, dur: self.rotateBones(names, quats, dur, instance=instance)
        if self.currentModel:
            t0 = [(n, v[0][1]) for n, v in byBone.items() if v and v[0][0] == 0]
            have0 = {n for n, _ in t0}
            if t0:
                rot([n for n, _ in t0], [q for _, q in t0], 0)
            reset = [b.name for b in self.currentModel.getSkeleton().bones if b.name not in have0]
            if reset:
                rot(reset, [Quat(0, 0, 0, 1)] * len(reset), 0)
        for name, keys in byBone.items():
            for i, (t, q) in enumerate(keys):
                if t == 0:
                    continue
                prev = keys[i - 1] if i > 0 else None
                durationMs = t * 1000 if i == 0 else (t - prev[0]) * 1000
                delayMs = prev[0] * 1000 if prev is not None else 0
                if delayMs <= 0:
                    rot([name], [q], durationMs)
                else:
                    self._animationTimeouts.append(
                        self._setTimeout(lambda n=name, qq=q, d=durationMs: rot([n], [qq], d), delayMs))
        if enableBreath and self.currentModel:
            maxTime = max((kf.time for kf in self.animationFrames), default=0)
            last: Dict[str, Quat] = {}
            for b in breathBones:
                keys = byBone.get(b)
                if keys:
                    for (t, q) in reversed(keys):
                        if t <= maxTime:
                            last[b] = q
                            break
            self._breathingTimeout = self._setTimeout(
                lambda: self._startBreathing(breathBones, last, breathRanges, breathDuration, instance), maxTime * 1000 + 200)

    def stopAnimation(self):
        for t in self._animationTimeouts:
            self._clearTimeout(t)
        self._animationTimeouts = []
        self.playingAnimation = False
        self._morphPlaying = False
        if self.crowd and self._crowd_playing and self.ctx:
            self.ctx.load_animation(None, None, None)
            self._crowd_playing = False

    def _stopBreathing(self):
        self._clearTimeout(self._breathingTimeout)
        self._breathingTimeout = None
        self._breathingBase.clear()

    def _startBreathing(self, bones, baseRotations, ranges, durationMs, instance):
        """engine.ts:1609-1662."""
        if not self.currentModel:
            return
        for b in bones:
            if b in baseRotations:
                self._breathingBase[b] = baseRotations[b]
        half = durationMs / 2

        def animate(inhale: bool):
            names, quats = [], []
            for b in bones:
                base = self._breathingBase.get(b)
                if base is None:
                    continue
                r = (ranges or {}).get(b, 0.02)
                names.append(b)
                quats.append(base.multiply(Quat.fromEuler(r if inhale else -r, 0, 0)))
            if names:
                self.rotateBones(names, quats, half, instance=instance)
            self._breathingTimeout = self._setTimeout(lambda: animate(not inhale), half)

        animate(False)

    # ---- frame ---------------------------------------------------------------------------------
    def render(self):
        """One frame of the replaced stage (engine.ts:2124 -> updateModelPose 2375-2391 -> draws)."""
        if self.ctx is None or self.currentModel is None:
            return
        self._pumpTimers()
        B = len(self.currentModel.skeleton.bones)
        if self._morphPlaying:
            tau = self.clock() - self._anim_start_ms - (self._offsets_ms if self.crowd else np.zeros(self.instances))
            ids, w = self._evalMorphTracks(np.maximum(tau, 0.0))
            if ids is not None:
                self.ctx.set_morph_weights(w, ids, K=self.instances)
        if self.crowd:
            m = self.currentModel
            now = self.clock()
            if self._crowd_playing:
                self.ctx.set_instance_clocks((now - self._anim_start_ms - self._offsets_ms).astype(np.float32), K=self.instances)
            else:
                # shared rotateBones tweens, evaluated per instance on the device; times are sent relative to `now`
                # (f32 has ~0.1 ms resolution only near zero)
                self.ctx.set_tweens(m._startQuat, m._targetQuat, (m._startTimeMs.astype(np.float64) - now).astype(np.float32),
                                    m._durationMs, m._active, m.localRotations)
                self.ctx.set_instance_clocks((-self._offsets_ms).astype(np.float32), K=self.instances)
                # (the host copy of the tween state is NOT advanced: a finished tween keeps evaluating to its target, and
                #  instances that lag behind still have to play it)
        elif self.gpu_pose:
            # host: tweens only (model.ts:158-194); device: hierarchy + append + skin matrices (model.ts:330-420)
            for k, m in enumerate(self.models):
                m.updateRotationTweens()
                self._rot_stage[k] = m.localRotations.reshape(B, 4)
            self.ctx.set_local_rotations(self._rot_stage, K=self.instances)
        else:
            # rz_palette_staging hands out two pinned buffers alternately and waits until the upload that last read the
            # one it returns has completed: ask before EVERY refill (the previous frame's copy may still be in flight)
            stage = self.ctx.palette_staging(self.instances)
            for k, m in enumerate(self.models):
                m.evaluatePose()
                stage[k] = m.getBoneWorldMatrices().reshape(B, 16)
            self.ctx.set_palettes(stage, K=self.instances)
        self.ctx.deform()

    def runRenderLoop(self, callback: Optional[Callable[[], None]] = None, frames: Optional[int] = None,
                      frame_ms: Optional[float] = None):
        """engine.ts:1668-1682.  rAF does not exist here: runs `frames` frames (or until stopRenderLoop()
        is called from the callback); `frame_ms` advances a ManualClock between frames."""
        self._renderLoopCallback = callback
        self._loop_running = True
        n = 0
        while self._loop_running and (frames is None or n < frames):
            self.render()
            if self._renderLoopCallback:
                self._renderLoopCallback()
            if frame_ms is not None and hasattr(self.clock, "advance"):
                self.clock.advance(frame_ms)
            n += 1
        self._loop_running = False

    def stopRenderLoop(self):
        self._loop_running = False
        self._renderLoopCallback = None

    def getStats(self) -> EngineStats:
        """engine.ts:1664-1666."""
        if self.ctx:
            s = self.ctx.stats()
            self._stats = EngineStats(s["fps"], s["frameTime"], s["gpuMemory"], s["vertsPerSec"], s["achievedGBs"], s["algorithmicBytes"])
        return EngineStats(**self._stats.__dict__)

    # ---- results (the reference hands these straight to the rasteriser) ---------------------------
    def deviceIndexBuffer(self) -> np.ndarray:
        """The model's triangle list (Model.getIndices(), engine.ts:1762-1771 uploads it as the index buffer) re-expressed
        against the device planes: identical to getIndices() unless reorder_vertices was requested."""
        idx = np.asarray(self.currentModel.getIndices(), np.uint32)
        if not (self._flags & capi.RZ_FLAG_REORDER_VERTICES):
            return idx
        order = self.ctx.vertex_order()
        inv = np.empty_like(order)
        inv[order] = np.arange(order.size, dtype=np.uint32)
        return inv[idx]

    @staticmethod
    def vertexEdgeSizes(model: Model) -> np.ndarray:
        """Material.edgeSize per vertex: the reference draws the outline of material m over m's slice of the index buffer
        when (edgeFlag & 0x10) and edgeSize > 0 (engine.ts:2016-2046), expanding by material.edgeSize (engine.ts:458-461).
        A vertex shared by two outlined materials takes the larger size."""
        idx = np.asarray(model.getIndices(), np.int64)
        edge = np.zeros(model.getVertexCount(), np.float32)
        start = 0
        for m in model.getMaterials():
            get = m.get if isinstance(m, dict) else (lambda k, _m=m: getattr(_m, k))
            n = int(get("vertexCount"))
            if (int(get("edgeFlag")) & 0x10) and float(get("edgeSize")) > 0:
                sl = idx[start:start + n]
                edge[sl] = np.maximum(edge[sl], np.float32(get("edgeSize")))
            start += n
        return edge

    def readOutline(self, instance: int = 0) -> np.ndarray:
        """Outline hull positions of one instance (Engine(outline=True)), [V,3] float32."""
        return self.ctx.read_outline(instance)

    def readInterleaved(self, instance: int = 0) -> np.ndarray:
        """One instance in the reference's vertex-buffer layout [x,y,z,nx,ny,nz,u,v] (Engine(interleaved=True))."""
        return self.ctx.read_interleaved(instance)

    def readSkinnedAsync(self, instance: int, out_pos: np.ndarray, out_nrm: Optional[np.ndarray] = None):
        """Queue the read-back of the frame just rendered and return at once; the arrays are valid after readWait()."""
        self.ctx.read_instance_async(instance, out_pos, out_nrm)

    def readWait(self):
        self.ctx.read_wait()

    def readSkinned(self, instance: int = 0):
        """Skinned positions and normals of one instance, [V,3] float32 each."""
        return self.ctx.read_instance(instance)


class MultiDeviceEngine:
    """A crowd over several GPUs from ONE process (SURVEY 8b row 1 "device list", 8e): one Engine — one rz_ctx, one CUDA
    stream — per entry of `devices`, the K instances split into contiguous ranges by `sharding.instance_range` (the same
    rule the one-process-per-GPU bench uses).  Every entry point of the C ABI is asynchronous on its device, so one host
    thread launches the frame on all devices before it waits for any; no data crosses between devices, results stay where
    they were produced (`readSkinned` routes to the owning shard).  The reference is one process with one GPUDevice
    (engine.ts:158-163): this is what its `Engine` becomes when the device is a list.  `devices` may name a device more
    than once (logical shards on one GPU: how the sharding is tested on a single-GPU box).

    Mirrors ts/engine.ts MultiDeviceEngine; the per-shard keyword arguments are Engine's."""

    def __init__(self, canvas=None, options: Optional[dict] = None, *, devices: Sequence[int], instances: int = 1, **engine_kwargs):
        from . import sharding
        if not devices:
            raise ValueError("devices must name at least one CUDA device")
        self.instances = int(instances)
        self.shards: List[tuple] = []                       # (engine, first, count)
        for g, dev in enumerate(devices):
            first, last = sharding.instance_range(self.instances, len(devices), g)
            if last > first:
                self.shards.append((Engine(canvas, options, instances=last - first, device=int(dev), **engine_kwargs), first, last - first))

    def _each(self):
        return (e for e, _, _ in self.shards)

    def init(self):
        for e in self._each():
            e.init()
        return self

    def dispose(self):
        for e in self._each():
            e.dispose()

    def loadModel(self, path):
        model = None
        for e in self._each():
            model = e.loadModel(path)
        return model

    def loadAnimation(self, url: str):
        for e in self._each():
            e.loadAnimation(url)

    def setInstanceOffsets(self, offsets_ms):
        off = np.ascontiguousarray(offsets_ms, dtype=np.float64).reshape(self.instances)
        for e, first, count in self.shards:
            e.setInstanceOffsets(off[first:first + count])

    def setMorphWeights(self, weights, morph_names_or_ids):
        w = np.ascontiguousarray(weights, dtype=np.float32).reshape(self.instances, -1)
        for e, first, count in self.shards:
            e.setMorphWeights(w[first:first + count], morph_names_or_ids)

    def playAnimation(self, options: Optional[dict] = None):
        for e in self._each():
            e.playAnimation(options, instance=None)

    def stopAnimation(self):
        for e in self._each():
            e.stopAnimation()

    def shardOf(self, instance: int):
        for e, first, count in self.shards:
            if first <= instance < first + count:
                return e, instance - first
        raise IndexError(instance)

    def rotateBones(self, bones, rotations, durationMs=None, instance: Optional[int] = None):
        """instance=None addresses every instance of every shard (the default here: a crowd call)."""
        if instance is None:
            for e in self._each():
                e.rotateBones(bones, rotations, durationMs, instance=None)
        else:
            e, local = self.shardOf(instance)
            e.rotateBones(bones, rotations, durationMs, instance=0 if e.crowd else local)

    def render(self):
        for e in self._each():                                # every device is launched before any is waited for
            e.render()

    def sync(self):
        for e in self._each():
            e.ctx.sync()

    def readSkinned(self, instance: int = 0):
        e, local = self.shardOf(instance)
        return e.readSkinned(local)

    def getStats(self) -> List[EngineStats]:
        return [e.getStats() for e in self._each()]
