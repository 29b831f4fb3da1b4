// engine.ts — TypeScript Engine for Node.js hosts: the reference's Engine surface (engine/src/engine.ts:145-157 constructor
// + init, 1419-1423 loadAnimation, 1425-1591 playAnimation, 1593-1599 stopAnimation, 1609-1662 breathing, 1664-1666 getStats,
// 1668-1690 runRenderLoop / stopRenderLoop, 1692-1701 dispose, 1704-1721 loadModel, 1723-1725 rotateBones, 2124 render) over
// the B200 deform path through the N-API addon (napi/rze_b200_napi.cc -> include/rze_b200.h).
//
// It is written to sit NEXT TO the reference's own sources (engine/src/): the CPU side — PmxLoader, VMDLoader, Model with its
// tween + hierarchy evaluation, Quat / Vec3 — is imported unchanged; only what the reference did on the GPU for this path
// (setupModelBuffers, the per-frame palette upload + skin-matrix pass, the vertex-shader blend) goes to `rz`.
//
// SOURCE ONLY in this repository: no JS runtime / tsc exists in the build image or on the GPU box, so this file is not
// compiled or executed here.  Every behaviour below is implemented a second time in Python against the same C ABI
// (reze-engine_b200/engine.py) and tested there; tests/test_host.py checks that every `rz.*` call made here is exported by the
// N-API shim with that name and that the reference's public methods for the path all exist.
//
//   reference                                              here
//   ---------                                              ----
//   init(): navigator.gpu adapter / device (157-185)       rz.create(device, instances, flags)
//   setupModelBuffers(model)  (1728-1832)                  rz.loadMesh(getVertices(), joints, weights, inverseBind) [+ loadSkeleton]
//   updateModelPose(): writeBuffer + compute (2375-2402)   rz.setPalettes(world, P, null, K)  |  rz.setLocalRotations  |  rz.setInstanceClocks
//   vertex-shader blend, re-run per draw (245-276)         rz.deform(0, K)              (once per frame, result materialised)
//   playAnimation timers (1451-1553)                       same timers on an injectable clock; crowd mode: rz.loadAnimation
import { Model } from "./model"
import { PmxLoader } from "./pmx-loader"
import { VMDLoader, VMDKeyFrame } from "./vmd-loader"
import { Quat, Vec3 } from "./math"
// eslint-disable-next-line @typescript-eslint/no-var-requires
const rz = require("./rze_b200.node")

export type EngineOptions = {
  // the reference's options (engine.ts:8-14): accepted so call sites construct unchanged; they parameterise passes this
  // engine does not replace
  ambient?: number; bloomIntensity?: number; rimLightIntensity?: number; cameraDistance?: number; cameraTarget?: Vec3
  // additions
  instances?: number        // crowd size K (default 1): K independent instances of the loaded model
  device?: number           // CUDA device ordinal (one Engine per GPU; see MultiDeviceEngine below for a device list)
  sdef?: boolean            // evaluate SDEF vertices spherically instead of as BDEF2 (needs loadSdef)
  bounds?: boolean          // also produce one AABB per instance
  outline?: boolean         // also produce the outline pass' hull positions (engine.ts:458-461) as a third plane
  interleaved?: boolean     // result in the reference's own 32-byte vertex layout [pos, nrm, uv] (engine.ts:340-347)
  doubleBuffer?: boolean    // two result buffers: frame n is read / drawn while frame n+1 is deformed
  gpuPose?: boolean         // host evaluates the tweens only; hierarchy + append + skin matrices run on the device
  crowd?: boolean           // one shared animation state, instance k plays it offsets[k] ms behind (GPU pose evaluation)
  clock?: () => number      // replaces performance.now() (model.ts:160,249) for reproducible playback
}
export interface EngineStats { fps: number; frameTime: number; gpuMemory: number; vertsPerSec?: number; achievedGBs?: number; algorithmicBytes?: number }

// rz_config.flags (include/rze_b200.h)
const RZ_FLAG_SDEF = 0x1, RZ_FLAG_BOUNDS = 0x4, RZ_FLAG_OUTLINE = 0x10, RZ_FLAG_INTERLEAVED = 0x20, RZ_FLAG_DOUBLE_BUFFER = 0x40

type Timer = { due: number; id: number; fn: () => void }
type BoneKey = { time: number; rotation: Quat }
// the private runtime state of the reference's Model (model.ts:53-68, 85, 91), read for the GPU-pose modes.  TypeScript's
// `private` is compile-time only; INTEGRATION.md shows the two getters a maintainer would add instead of this cast.
type ModelRuntime = {
  runtimeSkeleton: { localRotations: Float32Array }
  rotTweenState: { active: Uint8Array; startQuat: Float32Array; targetQuat: Float32Array; startTimeMs: Float32Array; durationMs: Float32Array }
  updateRotationTweens(): void
}
const runtimeOf = (m: Model) => m as unknown as ModelRuntime

export class Engine {
  private ctx: unknown = null
  private models: Model[] = []
  private currentModel: Model | null = null
  private readonly K: number
  private readonly clock: () => number
  private B = 0
  private V = 0
  private rotStage: Float32Array | null = null
  private world: Float32Array | null = null
  // timers: window.setTimeout (engine.ts:1547, 1587, 1657) on the injectable clock, pumped by render()
  private timers: Timer[] = []
  private nextTimer = 1
  private animationTimeouts: number[] = []
  private breathingTimeout: number | null = null
  private breathingBaseRotations = new Map<string, Quat>()
  private animationFrames: VMDKeyFrame[] = []
  private playingAnimation = false
  private running = false
  private renderLoopCallback: (() => void) | null = null
  // crowd mode
  private offsetsMs: Float64Array
  private animStartMs = 0
  private crowdPlaying = false
  private morphNames: string[] = []
  private bodyCount = 0

  constructor(_canvas: unknown, private options: EngineOptions = {}) {
    this.K = options.instances ?? 1
    this.clock = options.clock ?? (() => performance.now())
    this.offsetsMs = new Float64Array(this.K)
  }

  // engine.ts:157-185
  public async init() {
    const o = this.options
    const flags = (o.sdef ? RZ_FLAG_SDEF : 0) | (o.bounds ? RZ_FLAG_BOUNDS : 0) | (o.outline ? RZ_FLAG_OUTLINE : 0) |
      (o.interleaved ? RZ_FLAG_INTERLEAVED : 0) | (o.doubleBuffer ? RZ_FLAG_DOUBLE_BUFFER : 0)
    this.ctx = rz.create(o.device ?? 0, this.K, flags)   // throws when no sm_100 device is present: there is no CPU path
  }

  // engine.ts:1704-1721 + setupModelBuffers 1728-1832
  public async loadModel(path: string) {
    const model = await PmxLoader.load(path)
    this.currentModel = model
    this.models = [model]
    if (!this.options.crowd) for (let k = 1; k < this.K; k++) this.models.push(await PmxLoader.load(path))   // per-instance tween state
    const sk = model.getSkinning()
    const inv = model.getBoneInverseBindMatrices()
    this.B = inv.length / 16
    this.V = model.getVertexCount()
    rz.loadMesh(this.ctx, model.getVertices(), sk.joints, sk.weights, inv)
    if (this.options.outline) rz.loadEdgeSize(this.ctx, Engine.vertexEdgeSizes(model))
    if (this.options.gpuPose || this.options.crowd) {
      const bones = model.getSkeleton().bones
      const parent = new Int32Array(this.B), bind = new Float32Array(this.B * 3)
      const appendParent = new Int32Array(this.B).fill(-1), appendRatio = new Float32Array(this.B).fill(1), appendRotate = new Uint8Array(this.B)
      bones.forEach((b, i) => {
        parent[i] = b.parentIndex
        bind.set(b.bindTranslation, i * 3)
        if (b.appendRotate && b.appendParentIndex !== undefined && b.appendParentIndex >= 0) {
          appendParent[i] = b.appendParentIndex
          appendRatio[i] = b.appendRatio === undefined ? 1 : b.appendRatio   // "undefined" ratio means 1 (model.ts:360)
          appendRotate[i] = 1
        }
      })
      rz.loadSkeleton(this.ctx, parent, bind, appendParent, appendRatio, appendRotate)
      this.rotStage = new Float32Array(this.K * this.B * 4)
    } else {
      this.world = new Float32Array(this.K * this.B * 16)
    }
    const bodies = model.getRigidbodies()
    if (bodies.length) this.loadRigidBodies(bodies)
  }

  // ---- tables the reference's loader steps over (pmx-loader.ts:141-155, 450-553); a host that parses them hands them in
  public loadMorphs(offsets: Uint32Array, vertexIndex: Uint32Array, delta: Float32Array, names: string[] = []) {
    rz.loadMorphs(this.ctx, offsets, vertexIndex, delta)
    this.morphNames = names
  }
  public loadSdef(vertexIndex: Uint32Array, cR0R1: Float32Array) { rz.loadSdef(this.ctx, vertexIndex, cR0R1) }
  // per-instance weights [K][ids.length] of the given vertex morphs (names from loadMorphs, or indices)
  public setMorphWeights(weights: Float32Array, morphs: (string | number)[]) {
    const ids = Uint32Array.from(morphs.map((m) => (typeof m === "string" ? this.morphNames.indexOf(m) : m)))
    rz.setMorphWeights(this.ctx, weights, ids, this.K)
  }

  // Material.edgeSize per vertex: material m outlines its slice of the index buffer when (edgeFlag & 0x10) && edgeSize > 0
  // (engine.ts:2016-2046); tested restatement: reze-engine_b200/engine.py Engine.vertexEdgeSizes
  static vertexEdgeSizes(model: Model): Float32Array {
    const idx = model.getIndices(), edge = new Float32Array(model.getVertexCount())
    let start = 0
    for (const m of model.getMaterials()) {
      if ((m.edgeFlag & 0x10) !== 0 && m.edgeSize > 0)
        for (let i = start; i < start + m.vertexCount; i++) edge[idx[i]] = Math.max(edge[idx[i]], m.edgeSize)
      start += m.vertexCount
    }
    return edge
  }

  // ---- physics -> bone feedback: the solver stays in JS (physics.ts), its output patches the palette on the device
  // (called by loadModel; call it again once Physics has computed bodyOffsetMatrixInverse, physics.ts:571-592)
  public loadRigidBodies(bodies: ReturnType<Model["getRigidbodies"]>) {
    const n = bodies.length
    const bone = new Int32Array(n), dynamic = new Uint8Array(n), offInv = new Float32Array(n * 16)
    bodies.forEach((rb, i) => {
      bone[i] = rb.boneIndex
      dynamic[i] = rb.type === 1 ? 1 : 0                          // RigidbodyType.Dynamic (physics.ts:10-14)
      offInv.set(rb.bodyOffsetMatrixInverse.values, i * 16)       // physics.ts:560-585
    })
    this.bodyCount = rz.loadRigidBodies(this.ctx, bone, dynamic, offInv)
  }
  // after this frame's pose update, before deform: the solver's world transforms, 7 floats (pos xyz, rot xyzw) per palette per body
  public applyBodyTransforms(posQuat: Float32Array, palettes = this.K) { rz.applyBodyTransforms(this.ctx, posQuat, palettes, this.bodyCount) }
  // bone world matrices of a pose evaluated on the device, for the kinematic half of Physics.step (syncFromBones, physics.ts:649-703)
  public readWorldMatrices(palette: number, world: Float32Array) { rz.readWorldMatrices(this.ctx, palette, world) }

  // ---- bone API (engine.ts:1723-1725 -> model.ts:246-315); instance = null addresses every instance
  public rotateBones(bones: string[], rotations: Quat[], durationMs?: number, instance: number | null = 0) {
    const targets = instance === null ? this.models : this.models.slice(instance, instance + 1)
    for (const m of targets) m.rotateBones(bones, rotations, durationMs)
  }

  // ---- timers on the injectable clock
  private setTimeout(fn: () => void, delayMs: number): number {
    const id = this.nextTimer++
    this.timers.push({ due: this.clock() + Math.max(0, delayMs), id, fn })
    return id
  }
  private clearTimeout(id: number | null) { if (id !== null) this.timers = this.timers.filter((t) => t.id !== id) }
  private pumpTimers() {
    for (;;) {                                                    // a timer may schedule another one that is already due
      const now = this.clock()
      const due = this.timers.filter((t) => t.due <= now).sort((a, b) => a.due - b.due || a.id - b.id)
      if (!due.length) return
      const t = due[0]
      this.timers = this.timers.filter((x) => x.id !== t.id)
      t.fn()
    }
  }

  // ---- animation (engine.ts:1419-1423, 1425-1591)
  public async loadAnimation(url: string) { this.animationFrames = await VMDLoader.load(url) }

  // crowd mode: instance k plays the shared animation offsetsMs[k] milliseconds behind instance 0 ("staggered phase")
  public setInstanceOffsets(offsetsMs: ArrayLike<number>) { this.offsetsMs = Float64Array.from(offsetsMs) }

  private keysByBone(): Map<string, BoneKey[]> {
    const byBone = new Map<string, BoneKey[]>()
    for (const kf of this.animationFrames)
      for (const bf of kf.boneFrames) {
        if (!byBone.has(bf.boneName)) byBone.set(bf.boneName, [])
        byBone.get(bf.boneName)!.push({ time: kf.time, rotation: bf.rotation })
      }
    for (const keys of byBone.values()) keys.sort((a, b) => a.time - b.time)
    return byBone
  }

  // crowd playback: the clip becomes per-bone key tracks on the device, evaluated per instance clock with the same rule the
  // timers below produce when they fire on time (key at t = 0 instant, key i reached from key i-1 by ease + slerp)
  private playCrowd(byBone: Map<string, BoneKey[]>) {
    const names = this.currentModel!.getBoneNames()
    const off = new Uint32Array(this.B + 1)
    const times: number[] = [], quats: number[] = []
    names.forEach((name, i) => {
      const keys = byBone.get(name) ?? []
      off[i + 1] = off[i] + keys.length
      for (const k of keys) { times.push(k.time * 1000); quats.push(k.rotation.x, k.rotation.y, k.rotation.z, k.rotation.w) }
    })
    rz.loadAnimation(this.ctx, off, Float32Array.from(times), Float32Array.from(quats), null)
    this.animStartMs = this.clock()
    this.crowdPlaying = true
  }

  public playAnimation(options?: { breathBones?: string[] | Record<string, number>; breathDuration?: number }, instance: number | null = 0) {
    if (this.animationFrames.length === 0) return
    this.stopAnimation()
    this.stopBreathing()
    this.playingAnimation = true
    this.animStartMs = this.clock()
    const byBone = this.keysByBone()
    if (this.options.crowd) { this.playCrowd(byBone); return }

    const enableBreath = options?.breathBones !== undefined && options.breathBones !== null
    let breathBones: string[] = []
    let breathRotationRanges: Record<string, number> | undefined = undefined
    if (enableBreath && options!.breathBones) {
      if (Array.isArray(options!.breathBones)) breathBones = options!.breathBones
      else { breathBones = Object.keys(options!.breathBones); breathRotationRanges = options!.breathBones }
    }
    const breathDuration = options?.breathDuration ?? 4000
    const rot = (names: string[], quats: Quat[], dur: number) => this.rotateBones(names, quats, dur, instance)

    if (this.currentModel) {
      // keys at t = 0 apply instantly, every other bone of the model resets to identity (engine.ts:1474-1505)
      const time0: { boneName: string; rotation: Quat }[] = []
      const bonesWithTime0 = new Set<string>()
      for (const [boneName, keys] of byBone.entries())
        if (keys.length > 0 && keys[0].time === 0) { time0.push({ boneName, rotation: keys[0].rotation }); bonesWithTime0.add(boneName) }
      if (time0.length > 0) rot(time0.map((r) => r.boneName), time0.map((r) => r.rotation), 0)
      const reset = this.currentModel.getSkeleton().bones.filter((b) => !bonesWithTime0.has(b.name)).map((b) => b.name)
      if (reset.length > 0) rot(reset, new Array(reset.length).fill(new Quat(0, 0, 0, 1)), 0)
    }
    // key i > 0: a tween of duration t(i) - t(i-1) started at wall time t(i-1) (engine.ts:1529-1553)
    for (const [boneName, keys] of byBone.entries())
      for (let i = 0; i < keys.length; i++) {
        const key = keys[i], prev = i > 0 ? keys[i - 1] : null
        if (key.time === 0) continue
        const durationMs = i === 0 ? key.time * 1000 : (key.time - prev!.time) * 1000
        const delayMs = (prev ? prev.time : 0) * 1000
        if (delayMs <= 0) rot([boneName], [key.rotation], durationMs)
        else this.animationTimeouts.push(this.setTimeout(() => rot([boneName], [key.rotation], durationMs), delayMs))
      }
    // breathing after the clip's last key + 200 ms (engine.ts:1555-1590)
    if (enableBreath && this.currentModel) {
      let maxTime = 0
      for (const kf of this.animationFrames) if (kf.time > maxTime) maxTime = kf.time
      const last = new Map<string, Quat>()
      for (const bone of breathBones) {
        const keys = byBone.get(bone)
        if (!keys) continue
        for (let i = keys.length - 1; i >= 0; i--) if (keys[i].time <= maxTime) { last.set(bone, keys[i].rotation); break }
      }
      this.breathingTimeout = this.setTimeout(() => this.startBreathing(breathBones, last, breathRotationRanges, breathDuration, instance), maxTime * 1000 + 200)
    }
  }

  // engine.ts:1593-1599
  public stopAnimation() {
    for (const id of this.animationTimeouts) this.clearTimeout(id)
    this.animationTimeouts = []
    this.playingAnimation = false
    if (this.options.crowd && this.crowdPlaying && this.ctx) {
      rz.loadAnimation(this.ctx, null, null, null, null)          // unload: the shared tween table drives the crowd again
      this.crowdPlaying = false
    }
  }

  // engine.ts:1601-1607
  private stopBreathing() {
    this.clearTimeout(this.breathingTimeout)
    this.breathingTimeout = null
    this.breathingBaseRotations.clear()
  }

  // engine.ts:1609-1662
  private startBreathing(bones: string[], baseRotations: Map<string, Quat>, rotationRanges: Record<string, number> | undefined, durationMs = 4000,
                         instance: number | null = 0) {
    if (!this.currentModel) return
    for (const bone of bones) { const base = baseRotations.get(bone); if (base) this.breathingBaseRotations.set(bone, base) }
    const halfCycleMs = durationMs / 2
    const animate = (isInhale: boolean) => {
      if (!this.currentModel) return
      const names: string[] = [], quats: Quat[] = []
      for (const bone of bones) {
        const base = this.breathingBaseRotations.get(bone)
        if (!base) continue
        const rotation = rotationRanges?.[bone] ?? 0.02
        names.push(bone)
        quats.push(base.multiply(Quat.fromEuler(isInhale ? rotation : -rotation, 0, 0)))
      }
      if (names.length > 0) this.rotateBones(names, quats, halfCycleMs, instance)
      this.breathingTimeout = this.setTimeout(() => animate(!isInhale), halfCycleMs)
    }
    animate(false)                                                 // start from the exhale side (engine.ts:1660-1661)
  }

  // ---- one frame of the replaced stage (engine.ts:2124 -> updateModelPose 2375-2402 -> the draws' vertex stage)
  public render() {
    if (!this.ctx || !this.currentModel) return
    this.pumpTimers()
    const K = this.K, B = this.B
    if (this.options.crowd) {
      const now = this.clock()
      const clocks = new Float32Array(K)
      if (this.crowdPlaying) {
        for (let k = 0; k < K; k++) clocks[k] = now - this.animStartMs - this.offsetsMs[k]
      } else {
        // shared rotateBones tweens evaluated per instance on the device; times relative to `now` (f32 resolution)
        const rt = runtimeOf(this.currentModel)
        const tw = rt.rotTweenState
        const startRel = new Float32Array(B)
        for (let b = 0; b < B; b++) startRel[b] = tw.startTimeMs[b] - now
        rz.setTweens(this.ctx, tw.startQuat, tw.targetQuat, startRel, tw.durationMs, tw.active, rt.runtimeSkeleton.localRotations)
        for (let k = 0; k < K; k++) clocks[k] = -this.offsetsMs[k]
      }
      rz.setInstanceClocks(this.ctx, clocks, K, null, K)
    } else if (this.options.gpuPose) {
      // host: tweens only (model.ts:158-194); device: hierarchy + append + skin matrices (model.ts:330-420)
      this.models.forEach((m, k) => {
        const rt = runtimeOf(m)
        rt.updateRotationTweens()
        this.rotStage!.set(rt.runtimeSkeleton.localRotations, k * B * 4)
      })
      rz.setLocalRotations(this.ctx, this.rotStage, K, null, K)
    } else {
      // exactly the reference's feed: getBoneWorldMatrices() of every instance (engine.ts:2383-2389)
      this.models.forEach((m, k) => { m.evaluatePose(); this.world!.set(m.getBoneWorldMatrices(), k * B * 16) })
      rz.setPalettes(this.ctx, this.world, K, null, K)
    }
    rz.deform(this.ctx, 0, K)
  }

  // engine.ts:1668-1690 (requestAnimationFrame does not exist in Node: setImmediate paces the loop)
  public runRenderLoop(callback?: () => void) {
    this.renderLoopCallback = callback || null
    this.running = true
    const loop = () => {
      if (!this.running) return
      this.render()
      if (this.renderLoopCallback) this.renderLoopCallback()
      setImmediate(loop)
    }
    setImmediate(loop)
  }
  public stopRenderLoop() { this.running = false; this.renderLoopCallback = null }

  // engine.ts:1664-1666
  public getStats(): EngineStats { return rz.getStats(this.ctx) }

  // ---- results (the reference hands these straight to the rasteriser, engine.ts:270-274)
  public sync() { rz.sync(this.ctx) }
  public readSkinned(instance: number, pos: Float32Array | null, nrm: Float32Array | null) { rz.readInstance(this.ctx, instance, pos, nrm) }
  // frame n comes back while frame n+1 is being deformed: the arrays are valid after readWait()
  public readSkinnedAsync(instance: number, pos: Float32Array | null, nrm: Float32Array | null) { rz.readInstanceAsync(this.ctx, instance, pos, nrm) }
  public readWait() { rz.readWait(this.ctx) }
  public readOutline(instance: number, hull: Float32Array) { rz.readOutline(this.ctx, instance, hull) }
  public readInterleaved(instance: number, vtx8: Float32Array) { rz.readInterleaved(this.ctx, instance, vtx8) }
  public readBounds(first: number, count: number, minmax: Float32Array) { rz.readBounds(this.ctx, first, count, minmax) }
  public getOutputLayout() { return rz.getOutputLayout(this.ctx) }          // device pointer + strides for zero-copy consumers
  public getVertexOrder(order: Uint32Array) { rz.getVertexOrder(this.ctx, order) }

  // engine.ts:1692-1701
  public dispose() {
    this.stopRenderLoop()
    this.stopAnimation()
    this.stopBreathing()
    if (this.ctx) { rz.destroy(this.ctx); this.ctx = null }
  }
}

// A crowd over several GPUs from ONE process (SURVEY 8e): one Engine (= one rz_ctx, one CUDA stream) per device, instances
// split into contiguous ranges exactly like reze-engine_b200/sharding.py instance_range; every call is asynchronous on its
// device, so one host thread keeps all of them busy.  No data crosses between devices: outputs stay where they were made.
export class MultiDeviceEngine {
  readonly shards: { engine: Engine; first: number; count: number }[] = []
  constructor(canvas: unknown, options: EngineOptions & { devices: number[] }) {
    const K = options.instances ?? 1, G = options.devices.length
    const base = Math.floor(K / G), rem = K % G
    let first = 0
    options.devices.forEach((device, g) => {
      const count = base + (g < rem ? 1 : 0)
      if (count > 0) this.shards.push({ engine: new Engine(canvas, { ...options, device, instances: count }), first, count })
      first += count
    })
  }
  async init() { for (const s of this.shards) await s.engine.init() }
  async loadModel(path: string) { for (const s of this.shards) await s.engine.loadModel(path) }
  async loadAnimation(url: string) { for (const s of this.shards) await s.engine.loadAnimation(url) }
  setInstanceOffsets(offsetsMs: ArrayLike<number>) {
    for (const s of this.shards) s.engine.setInstanceOffsets(Array.prototype.slice.call(offsetsMs, s.first, s.first + s.count))
  }
  playAnimation(options?: Parameters<Engine["playAnimation"]>[0]) { for (const s of this.shards) s.engine.playAnimation(options, null) }
  stopAnimation() { for (const s of this.shards) s.engine.stopAnimation() }
  rotateBones(bones: string[], rotations: Quat[], durationMs?: number) { for (const s of this.shards) s.engine.rotateBones(bones, rotations, durationMs, null) }
  render() { for (const s of this.shards) s.engine.render() }        // every device is launched before any is waited for
  sync() { for (const s of this.shards) s.engine.sync() }
  shardOf(instance: number) { return this.shards.find((s) => instance >= s.first && instance < s.first + s.count)! }
  readSkinned(instance: number, pos: Float32Array | null, nrm: Float32Array | null) {
    const s = this.shardOf(instance)
    s.engine.readSkinned(instance - s.first, pos, nrm)
  }
  getStats(): EngineStats[] { return this.shards.map((s) => s.engine.getStats()) }
  dispose() { for (const s of this.shards) s.engine.dispose() }
}
