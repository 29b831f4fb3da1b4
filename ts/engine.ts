// engine.ts — TypeScript facade for Node.js hosts: the reference's Engine surface (engine/src/engine.ts:145-157,
// 1419-1425, 1593, 1664-1725) over the B200 deform path via the N-API addon (napi/rze_b200_napi.cc).
//
// SOURCE ONLY in this repository: no JS runtime / tsc exists in the build image or on the GPU box, so this file is not
// compiled or executed here.  The behaviour it mirrors is implemented and tested in Python
// (reze-engine_b200/engine.py), against the same C ABI.  A host keeps using the reference's own loaders and skeleton
// runtime (PmxLoader, VMDLoader, Model from "reze-engine"): only the GPU stage is swapped.
//
//   reference                                        here
//   ---------                                        ----
//   init(): navigator.gpu adapter/device             rz.create(device, instances, flags)
//   setupModelBuffers(model)  (engine.ts:1728-1832)  rz.loadMesh(getVertices(), joints, weights, inverseBind)
//   updateModelPose(): writeBuffer + compute pass    rz.setPalettes(worldMatrices, P, null, K)
//   vertex-shader blend per draw (engine.ts:245-276) rz.deform(0, K)            (once per frame, materialised)
import { Model, PmxLoader, VMDLoader, Quat, Vec3 } from "reze-engine"   // the reference's CPU side, unchanged
// eslint-disable-next-line @typescript-eslint/no-var-requires
const rz = require("./rze_b200.node")

export type EngineOptions = {
  ambient?: number; bloomIntensity?: number; rimLightIntensity?: number; cameraDistance?: number; cameraTarget?: Vec3
  instances?: number        // new: crowd size K (default 1)
  device?: number           // new: CUDA device ordinal (one process per GPU)
  sdef?: boolean            // new: evaluate SDEF spherically instead of as BDEF2
  outline?: boolean         // new: also produce the outline pass' hull positions (engine.ts:458-461) as a third plane
  interleaved?: boolean     // new: result in the reference's own 32-byte vertex layout [pos, nrm, uv] (engine.ts:340-347)
  doubleBuffer?: boolean    // new: two result buffers, frame n is read / drawn while frame n+1 is deformed
  clock?: () => number      // new: replaces performance.now() (model.ts:160,249) for reproducible playback
}
export interface EngineStats { fps: number; frameTime: number; gpuMemory: number; vertsPerSec?: number; achievedGBs?: number }

export class Engine {
  private ctx: unknown = null
  private models: Model[] = []
  private world!: Float32Array
  private timers: { due: number; id: number; fn: () => void }[] = []
  private nextTimer = 1
  private running = false
  private animationFrames: ReturnType<typeof VMDLoader.loadFromBuffer> = []
  private readonly clock: () => number
  private readonly K: number

  constructor(_canvas: unknown, private options: EngineOptions = {}) {
    this.K = options.instances ?? 1
    this.clock = options.clock ?? (() => performance.now())
  }

  async init() {
    const o = this.options   // rz_config.flags: RZ_FLAG_SDEF 0x1, RZ_FLAG_OUTLINE 0x10, RZ_FLAG_INTERLEAVED 0x20
    this.ctx = rz.create(o.device ?? 0, this.K, (o.sdef ? 0x1 : 0) | (o.outline ? 0x10 : 0) | (o.interleaved ? 0x20 : 0) | (o.doubleBuffer ? 0x40 : 0))
  }

  async loadModel(path: string) {
    const model = await PmxLoader.load(path)
    this.models = [model]
    for (let k = 1; k < this.K; k++) this.models.push(await PmxLoader.load(path))
    const sk = model.getSkinning()
    rz.loadMesh(this.ctx, model.getVertices(), sk.joints, sk.weights, model.getBoneInverseBindMatrices())
    this.world = new Float32Array(this.K * model.getBoneInverseBindMatrices().length)
    if (this.options.outline) rz.loadEdgeSize(this.ctx, Engine.vertexEdgeSizes(model))
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

  async loadAnimation(url: string) { this.animationFrames = await VMDLoader.load(url) }
  rotateBones(bones: string[], rotations: Quat[], durationMs?: number, instance = 0) {
    this.models[instance]?.rotateBones(bones, rotations, durationMs)
  }

  private setTimeout(fn: () => void, delayMs: number) {
    const id = this.nextTimer++
    this.timers.push({ due: this.clock() + Math.max(0, delayMs), id, fn })
    return id
  }

  render() {
    const now = this.clock()
    const due = this.timers.filter((t) => t.due <= now).sort((a, b) => a.due - b.due)
    this.timers = this.timers.filter((t) => t.due > now)
    for (const t of due) t.fn()
    const n = this.world.length / this.K
    this.models.forEach((m, k) => { m.evaluatePose(); this.world.set(m.getBoneWorldMatrices(), k * n) })
    rz.setPalettes(this.ctx, this.world, this.K, null, this.K)
    rz.deform(this.ctx, 0, this.K)
  }

  runRenderLoop(callback?: () => void) {
    this.running = true
    const loop = () => { if (!this.running) return; this.render(); callback?.(); setImmediate(loop) }
    setImmediate(loop)
  }
  stopRenderLoop() { this.running = false }
  getStats(): EngineStats { return rz.getStats(this.ctx) }
  readSkinned(instance: number, pos: Float32Array, nrm: Float32Array | null) { rz.readInstance(this.ctx, instance, pos, nrm) }
  // frame n comes back while frame n+1 is being deformed: resolve after the NEXT render() (or call readWait yourself)
  readSkinnedAsync(instance: number, pos: Float32Array, nrm: Float32Array | null) { rz.readInstanceAsync(this.ctx, instance, pos, nrm) }
  readWait() { rz.readWait(this.ctx) }
  readOutline(instance: number, hull: Float32Array) { rz.readOutline(this.ctx, instance, hull) }
  readInterleaved(instance: number, vtx8: Float32Array) { rz.readInterleaved(this.ctx, instance, vtx8) }
  getOutputLayout() { return rz.getOutputLayout(this.ctx) }
  dispose() { this.stopRenderLoop(); this.ctx = null }
  // playAnimation / stopAnimation / breathing: identical scheduling to engine.ts:1425-1662 on top of this.setTimeout;
  // see reze-engine_b200/engine.py (playAnimation, _startBreathing) for the tested restatement.
}
