/*
 * rze_b200.h — C ABI of the B200-native per-frame vertex-deformation path
 * (vertex-morph accumulation fused with BDEF1/2/4/SDEF linear-blend skinning of
 * positions + normals against a per-frame bone-matrix palette).
 *
 * This is the drop-in boundary for the deform stage of AmyangXYZ/reze-engine.
 * Each entry point names the reference interface it replaces (file:line under
 * the reference's engine/src/).  Plain pointers and sizes only; every input is
 * caller-owned and copied before the call returns; outputs are library-owned.
 * A rz_ctx is NOT thread-safe: one host thread drives it (the reference is
 * single-threaded: one rAF callback, engine.ts:1671-1681).
 *
 * All functions return 0 (RZ_OK) on success or a negative rz_status; the text of
 * the last failure is available from rz_last_error().  The reference convention
 * "fatal => throw new Error" (engine.ts:161,167,1828) maps to: the N-API / ctypes
 * shim raises when status != 0.
 */
#ifndef RZE_B200_H
#define RZE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RZE_B200_ABI_VERSION 3

typedef struct rz_ctx rz_ctx;

typedef enum rz_status {
  RZ_OK = 0,
  RZ_ERR_INVALID_ARG = -1,
  RZ_ERR_NO_DEVICE = -2,   /* no CUDA device / sm_100 kernel image not loadable: never a CPU fallback */
  RZ_ERR_CUDA = -3,
  RZ_ERR_OOM = -4,
  RZ_ERR_STATE = -5        /* call order violated (e.g. deform before load_mesh / set_palettes) */
} rz_status;

/* rz_config.flags */
#define RZ_FLAG_SDEF        0x1u  /* evaluate SDEF vertices spherically; default (0) = BDEF2, exactly as the
                                     reference treats them (pmx-loader.ts:141-155) */
#define RZ_FLAG_NO_NORMALS  0x2u  /* positions only — the reference's depth-only blend (engine.ts:692-715) */
#define RZ_FLAG_BOUNDS      0x4u  /* also produce one AABB per instance (fused consumer) */
#define RZ_FLAG_REORDER_VERTICES 0x8u /* opt-in: the device planes store vertices sorted by bone tuple (like a vertex-cache
                                     optimiser reorders a mesh); ~20 % faster.  rz_get_vertex_order() gives the order so the
                                     caller can remap its index buffer once; rz_read_instance still returns caller order */

#define RZ_FLAG_OUTLINE     0x10u /* fused consumer: a third plane per instance with the outline hull position
                                     pos' + normalize(n') * edgeSize * 0.01 — what the reference's outline vertex shader
                                     feeds the rasteriser (engine.ts:431-463, expansion 458-461); edgeSize per vertex from
                                     rz_load_edge_size (0 = no outline: hull == pos') */
#define RZ_FLAG_INTERLEAVED 0x20u /* fused consumer: the result leaves as ONE stream of 8 f32 per vertex
                                     [x,y,z,nx,ny,nz,u,v] per instance — the reference's own vertex-buffer layout
                                     (arrayStride 32, engine.ts:340-347; model.ts:196-200), so an unmodified pipeline can
                                     draw instance k with an identity palette.  uv is passed through from rz_load_mesh.
                                     Not combinable with RZ_FLAG_NO_NORMALS / RZ_FLAG_OUTLINE */

#define RZ_FLAG_DOUBLE_BUFFER 0x40u /* two result buffers: every palette update (rz_set_palettes*, rz_set_local_rotations,
                                     rz_set_instance_clocks = a new frame) flips to the other one, so a consumer keeps
                                     reading frame n (rz_read_instance_async, or a renderer holding the pointer of
                                     rz_get_output_layout) while frame n+1 is written.  Doubles the result memory */

typedef struct rz_config {
  uint32_t struct_size;    /* sizeof(rz_config), for forward compatibility */
  int32_t  device;         /* CUDA device ordinal this context owns (one process per GPU) */
  uint32_t max_instances;  /* K: independent character instances resident on this device (>=1) */
  uint32_t flags;          /* RZ_FLAG_* */
  void*    stream;         /* optional cudaStream_t to launch on (e.g. the caller's current stream);
                              NULL = the context creates its own non-blocking stream */
  uint32_t tune_instances_per_group; /* 0 = auto; kernel tuning knob (instances sharing one vertex pass) */
  uint32_t tune_store_mode;          /* 0 = auto; sub-batch size of the two-vertices-per-lane kernel (instances whose gathers are in
                                        flight together and whose stores form one commit group); ignored by the other kernels */
  uint32_t tune_threads;             /* 0 = auto; threads per CTA: 256, 512, 768 or 1024 (must be a compiled launch shape) */
  uint32_t tune_chunks;              /* 0 = auto; vertex chunks per instance group (work-item granularity) */
  uint32_t tune_ctas_per_sm;         /* 0 = auto; persistent CTAs per SM */
  uint32_t tune_vertices_per_lane;   /* 0 = auto (two on the plain planar path, one wherever morphs / SDEF / a fused consumer
                                        run); 1 = always the one-vertex-per-lane kernel; 2 = insist on the two-vertex kernel */
  uint32_t tune_reserved[2];
} rz_config;

typedef struct rz_stats {
  /* EngineStats of the reference (engine.ts:16-20): same three fields, same units */
  double fps;               /* frames per second over the last second of rz_deform calls */
  double frameTime;         /* ms, mean device time of the last <=60 rz_deform calls */
  double gpuMemory;         /* MB of device memory held by this context */
  /* additions */
  double vertsPerSec;       /* skinned vertices per second, last rz_deform */
  double algorithmicBytes;  /* compulsory DRAM bytes of the last rz_deform */
  double achievedGBs;       /* algorithmicBytes / device time */
  double lastDeformMs;      /* device time of the last rz_deform (CUDA events on the ctx stream) */
  uint64_t frames;          /* rz_deform calls so far */
  uint64_t kernelLaunches;  /* kernels launched by this context so far */
  uint32_t vertexCount, boneCount, instanceCount, paletteCount;
  uint32_t morphCount, morphNnz, sdefCount, activeMorphs;
  uint32_t instancesPerGroup, storeMode, ctas, threads; /* launch shape actually used */
  uint32_t smemBytes;
  uint32_t verticesPerLane;    /* 2: the last rz_deform ran the two-vertices-per-lane kernel (deform2_kernel.cuh), else 1 */
  uint32_t fastGatherPermille; /* share of the warp-level palette-row gathers whose lane pairs were packed onto the
                                * shared-memory broadcast fast path at load time (pair packing) */
} rz_stats;

/* ---- lifecycle: replaces Engine.init() (engine.ts:157-185) / dispose() (engine.ts:1692-1701) ---- */
int32_t rz_create(const rz_config* cfg, rz_ctx** out);
int32_t rz_destroy(rz_ctx* ctx);
uint32_t rz_abi_version(void);

/* ---- static tables: replaces setupModelBuffers (engine.ts:1728-1832) ----
 * vtx8    : V x [x,y,z,nx,ny,nz,u,v] f32, exactly Model.getVertices() (model.ts:196-200)
 * joints  : V x 4 u16, weights: V x 4 u8 (UNORM8), exactly Model.getSkinning() (model.ts:47-50)
 * invBind : B x 16 f32 column-major, exactly Model.getBoneInverseBindMatrices() (model.ts:321-323)
 * joints >= B are rejected (the reference's loader guarantees joints < B, pmx-loader.ts:865-876). */
int32_t rz_load_mesh(rz_ctx* ctx, const float* vtx8, const uint16_t* joints, const uint8_t* weights,
                     uint32_t V, const float* invBind, uint32_t B);

/* Vertex morphs, morph-major CSR in PMX order (layout documented by pmx-loader.ts:483-488):
 * morph m owns entries morphOffsets[m] .. morphOffsets[m+1]-1 of (vertIdx, delta3).  M may be 0. */
int32_t rz_load_morphs(rz_ctx* ctx, const uint32_t* morphOffsets /* M+1 */, const uint32_t* vertIdx /* nnz */,
                       const float* delta3 /* 3*nnz */, uint32_t M);

/* SDEF records the reference steps over (pmx-loader.ts:153-155): per SDEF vertex C, R0, R1.
 * Only used when RZ_FLAG_SDEF is set; the vertex must carry exactly two influences. */
int32_t rz_load_sdef(rz_ctx* ctx, const uint32_t* vertIdx /* n */, const float* c_r0_r1 /* 9*n */, uint32_t n);

/* Outline width per vertex for RZ_FLAG_OUTLINE: Material.edgeSize (model.ts:24, uploaded at engine.ts:2030) of the
 * material the vertex is drawn with when that material has an outline ((edgeFlag & 0x10) && edgeSize > 0,
 * engine.ts:2024), else 0.  NULL resets every vertex to 0.  The 0.01 scale factor of engine.ts:459 is applied here. */
int32_t rz_load_edge_size(rz_ctx* ctx, const float* edgeSize /* V or NULL */);

/* ---- per frame: replaces queue.writeBuffer(worldMatrixBuffer) + computeSkinMatrices
 * (engine.ts:2383-2402, shader engine.ts:920-929) ----
 * world : P x B x 16 f32 column-major, exactly Model.getBoneWorldMatrices() (model.ts:317-319), P palettes.
 * instToPalette : K entries < P, or NULL for identity (then P must be >= K).
 * skin = world * invBind is evaluated on the device. */
int32_t rz_set_palettes(rz_ctx* ctx, const float* world, uint32_t P, const uint32_t* instToPalette, uint32_t K);
/* (Large identity-mapped uploads are pipelined: the matrices travel in ~16 MB blocks on an internal copy stream and
 * rz_deform consumes them block by block — wait, skin matrices, deform that block's instances — so the PCIe transfer
 * overlaps the deform.  Results are bit-identical; environment RZ_NO_PIPELINE=1 disables it.) */
/* same, `world` / `instToPalette` are device pointers on this context's device (zero-copy producers).
 * The skin-matrix pass over `d_world` is issued by the next rz_deform as part of that frame's CUDA graph (or by whichever
 * other entry point needs the matrices first: rz_sync, rz_read_skin_matrices, rz_apply_body_transforms): keep `d_world`
 * unchanged until then. */
int32_t rz_set_palettes_device(rz_ctx* ctx, const float* d_world, uint32_t P, const uint32_t* d_instToPalette, uint32_t K);
/* Pinned host staging the caller fills directly and passes to rz_set_palettes / rz_set_local_rotations /
 * rz_set_instance_clocks (saves one host copy; uploads from it are asynchronous).  OWNERSHIP: the library keeps TWO staging
 * buffers and hands them out alternately; the call blocks until the upload that last read the buffer it returns has
 * completed, so the producer refills one buffer while the other is still in flight.  Call it before EVERY refill — a
 * pointer obtained earlier must not be written again (its upload may still be running).  A returned pointer stays valid
 * until the second-next call or rz_destroy.  Arguments that do not point into a staging buffer are copied through
 * library-owned pinned memory (never through the caller's staging). */
int32_t rz_palette_staging(rz_ctx* ctx, size_t bytes, void** host_ptr);

/* ---- GPU pose evaluation (replaces Model.evaluatePose, model.ts:325-328, + the palette upload for crowds) ----
 * rz_load_skeleton: the static part of Skeleton (model.ts:31-45) the hierarchy walk needs.
 *   parent[B] (-1 / out of range = root), bindTranslation[3B] (Bone.bindTranslation as f32, pmx-loader.ts:423),
 *   appendParent[B] / appendRatio[B] / appendRotate[B] may be NULL (no append bones); the ratio is clamped to [-1,1]
 *   and ignored unless appendRotate != 0 and the parent index is valid (model.ts:356-361). */
int32_t rz_load_skeleton(rz_ctx* ctx, const int32_t* parent, const float* bindTranslation, const int32_t* appendParent,
                         const float* appendRatio, const uint8_t* appendRotate, uint32_t B);
/* Local bone rotations (SkeletonRuntime.localRotations, model.ts:55: xyzw per bone) of P poses; the device walks the
 * hierarchy (model.ts:330-420), multiplies by invBind and fills the palettes.  Same (P, instToPalette, K) meaning as
 * rz_set_palettes; 4x less host->device traffic than world matrices. */
int32_t rz_set_local_rotations(rz_ctx* ctx, const float* quats /* P*B*4 */, uint32_t P, const uint32_t* instToPalette, uint32_t K);
/* Rotation tween state shared by the crowd (RotationTweenState, model.ts:62-68) + rest rotations of inactive bones. */
int32_t rz_set_tweens(rz_ctx* ctx, const float* startQuat /* 4B */, const float* targetQuat /* 4B */, const float* startTimeMs /* B */,
                      const float* durationMs /* B */, const uint8_t* active /* B */, const float* restQuat /* 4B */);
/* Evaluate the tweens (model.ts:158-194: slerp + quadratic ease) at P clock values and fill the palettes: instance k
 * plays the shared animation at time nowMs[instToPalette[k]] ("staggered phase" crowds).  Host->device traffic: 4*P bytes. */
int32_t rz_set_instance_clocks(rz_ctx* ctx, const float* nowMs /* P */, uint32_t P, const uint32_t* instToPalette, uint32_t K);
/* Keyframe tracks of one clip (what loadAnimation + playAnimation schedule, engine.ts:1419-1423, 1451-1553): bone b owns keys
 * keyOffsets[b] .. keyOffsets[b+1]-1 (times in ms, ascending; quaternions xyzw, normalised on load like rotateBones does).
 * While a clip is loaded rz_set_instance_clocks evaluates it instead of the tween table: a key at t = 0 applies instantly,
 * bones without one start from identity, key i is reached from key i-1 by slerp + quadratic ease over [t(i-1), t(i)], the
 * last key is held; bones without keys take restQuat (NULL = identity).  keyOffsets == NULL unloads the clip. */
int32_t rz_load_animation(rz_ctx* ctx, const uint32_t* keyOffsets /* B+1 */, const float* keyTimesMs, const float* keyQuats /* 4*nKeys */,
                          const float* restQuat /* 4B or NULL */);

/* ---- physics -> bone feedback, the matrix plumbing of Physics.step (physics.ts:534-569, 714-751); the solver itself
 * (Bullet / Ammo wasm) stays with the caller ----
 * rz_load_rigid_bodies: per body the bone it is attached to (RigidBody.boneIndex; < 0 or >= B: none), whether it is
 *   RigidbodyType.Dynamic (only those drive bones), and bodyOffsetMatrixInverse (physics.ts:560-585, column-major).
 * rz_apply_body_transforms: after this frame's palette update and before rz_deform; posQuat = per palette, per body the
 *   solver's world transform (origin x,y,z + rotation x,y,z,w).  Every bone driven by dynamic bodies gets
 *   boneWorld = fromPositionRotation(pos, rot) x bodyOffsetMatrixInverse, bodies in index order, NaN / >= 1e6 results skipped
 *   (physics.ts:744-748); other bones — including the children of driven bones — keep the matrices of the pose evaluation,
 *   exactly like the reference's in-place edit. */
int32_t rz_load_rigid_bodies(rz_ctx* ctx, const int32_t* boneIndex /* n */, const uint8_t* dynamic /* n */,
                             const float* bodyOffsetInverse /* 16*n */, uint32_t n);
int32_t rz_apply_body_transforms(rz_ctx* ctx, const float* posQuat /* P*n*7 */, uint32_t P);

/* Per-instance weights of the active morphs: w[k*M_active + a] scales morph activeIds[a] for instance k.
 * K must match the instance count in use; M_active = 0 disables morphing. */
int32_t rz_set_morph_weights(rz_ctx* ctx, const float* w, const uint32_t* activeIds, uint32_t M_active, uint32_t K);

/* ---- the deform pass: replaces the vertex-shader blend (engine.ts:245-276; 431-463; 692-715) ----
 * Deforms instances [firstInstance, firstInstance+count) asynchronously on the context's stream. */
int32_t rz_deform(rz_ctx* ctx, uint32_t firstInstance, uint32_t count);
int32_t rz_sync(rz_ctx* ctx);

/* ---- results (the reference feeds the rasteriser; nothing is stored, engine.ts:270-274) ----
 * Device layout per instance k: pos plane V x 3 f32 at  base + k*instanceStride,
 *                               nrm plane V x 3 f32 at  base + k*instanceStride + normalOffset.  */
int32_t rz_output_device_ptr(rz_ctx* ctx, void** base, size_t* instanceStride, size_t* normalOffset);
/* Full description of the device result (all sizes in bytes).  Attribute a of vertex i of instance k lives at
 * base + k*instanceStride + aOffset + i*vertexStride.  Planar (default): vertexStride 12, one plane per attribute.
 * RZ_FLAG_INTERLEAVED: vertexStride 32, offsets 0 / 12 / 24 inside the vertex.  An attribute that is not produced has
 * offset RZ_NO_ATTRIBUTE. */
#define RZ_NO_ATTRIBUTE ((size_t)-1)
typedef struct rz_output_layout {
  void*  base;
  size_t instanceStride;
  size_t vertexStride;
  size_t positionOffset;   /* always 0 */
  size_t normalOffset;     /* RZ_NO_ATTRIBUTE with RZ_FLAG_NO_NORMALS */
  size_t hullOffset;       /* outline hull plane, RZ_FLAG_OUTLINE only */
  size_t uvOffset;         /* RZ_FLAG_INTERLEAVED only */
} rz_output_layout;
int32_t rz_get_output_layout(rz_ctx* ctx, rz_output_layout* out);
int32_t rz_read_instance(rz_ctx* ctx, uint32_t inst, float* pos3 /* 3V or NULL */, float* nrm3 /* 3V or NULL */);
/* Asynchronous variant: the copies are queued behind the work issued so far on an internal read stream and the call
 * returns at once; rz_read_wait blocks until they have landed.  pos3 / nrm3 should be page-locked.  A later rz_deform that
 * would overwrite the buffer being read waits for the copy on the device (never with RZ_FLAG_DOUBLE_BUFFER, where the
 * next frame goes to the other buffer).  Planar layout in caller order only. */
int32_t rz_read_instance_async(rz_ctx* ctx, uint32_t inst, float* pos3 /* 3V or NULL */, float* nrm3 /* 3V or NULL */);
int32_t rz_read_wait(rz_ctx* ctx);
/* outline hull positions of one instance (RZ_FLAG_OUTLINE), caller vertex order */
int32_t rz_read_outline(rz_ctx* ctx, uint32_t inst, float* hull3 /* 3V */);
/* the interleaved stream of one instance (RZ_FLAG_INTERLEAVED), caller vertex order: 8 f32 per vertex */
int32_t rz_read_interleaved(rz_ctx* ctx, uint32_t inst, float* vtx8 /* 8V */);
/* order[i] = caller vertex id stored at position i of the device planes (identity unless RZ_FLAG_REORDER_VERTICES) */
int32_t rz_get_vertex_order(rz_ctx* ctx, uint32_t* order /* V */);

/* ---- diagnostics: the load-time lane / influence-slot plan rz_load_mesh applies (pair packing), computed on
 * the host for the given tables; needs no context and no device (CPU tests, tooling).  Vp = V rounded up to 256.
 * mode: 0 natural lane order, 1 lanes sorted, 2 pair packing (the library default).
 * laneVertex [Vp]: vertex evaluated by warp*32+lane (0xFFFFFFFF = padding); laneJoints / laneWeights [Vp*4]: the
 * influence table the kernel sees; stats [27]: fast, total gather instructions of packed warps, then the 5x5 histogram
 * of packed warps by [slot count N][mixed slots m].  Any output may be NULL. */
int32_t rz_plan_lanes(const uint16_t* joints, const uint8_t* weights, uint32_t V, uint32_t B, uint32_t mode,
                      uint32_t* laneVertex, uint16_t* laneJoints, float* laneWeights, uint64_t* stats);
/* The per-warp morph rows rz_load_morphs builds for a given lane plan, device-free like rz_plan_lanes:
 * warp w owns rows rowFirst[w] .. +rowDepth[w]-1, entry u of lane l at rows[(rowFirst[w] + 32*u + l)*4 .. +3] =
 * (dx, dy, dz, morph id bits); morphMajor[w] = 1 when a row holds ONE morph for the whole warp.  laneVertex as returned by
 * rz_plan_lanes.  rows may be NULL (size query: *rowsNeeded entries of 4 floats). */
int32_t rz_plan_morph_rows(const uint32_t* laneVertex /* Vp */, uint32_t Vp, uint32_t V, const uint32_t* morphOffsets /* M+1 */,
                           const uint32_t* vertIdx, const float* delta3, uint32_t M, uint32_t* rowFirst /* Vp/32 */,
                           uint32_t* rowDepth /* Vp/32 */, uint8_t* morphMajor /* Vp/32 */, float* rows, uint64_t rowsCapacity,
                           uint64_t* rowsNeeded);
/* The lane plan of the two-vertices-per-lane kernel (deform2_kernel.cuh), device-free like rz_plan_lanes.
 * Windows of 64 consecutive vertices become one group of 32 lanes x 2 vertices sharing one list of <= 4 palette rows per lane
 * (groupPaired = 1) or, where no such pairing exists, two ordinary groups of 32 lanes x 1 vertex.  Per group: first output vertex,
 * vertex count; per lane (group*32 + lane): the two vertices (0xFFFFFFFF = none), the four rows it gathers, the shader-normalised
 * weight of every row for vertex A and for vertex B (0 where the row is not that vertex' bone), the staging slot (0..63) each
 * side writes — A-slots of a group are distinct modulo 32, likewise B-slots (bank-conflict-free staging) — and the number of
 * rows the lane uses.  stats[4]: gather instructions on the fast path, all gather instructions, paired windows, fallback
 * windows.  Call with every output NULL to learn *nGroups.  Any output may be NULL. */
int32_t rz_plan_lanes2(const uint16_t* joints, const uint8_t* weights, uint32_t V, uint32_t B, uint32_t groupCapacity,
                       uint32_t* groupFirst, uint32_t* groupCount, uint8_t* groupPaired, uint32_t* laneVertA, uint32_t* laneVertB,
                       uint16_t* laneJoints, float* laneWeightsA, float* laneWeightsB, uint64_t* stats, uint32_t* nGroups,
                       uint8_t* laneSlotA, uint8_t* laneSlotB, uint8_t* laneSlots);
/* The bank-aware palette permutation rz_load_mesh applies: bonePos[b] = palette row of bone b, chosen so
 * that bones gathered by the same warp instruction sit in different 16-byte bank groups.  laneJoints as returned by
 * rz_plan_lanes. */
int32_t rz_plan_palette_rows(const uint16_t* laneJoints /* Vp*4 */, uint32_t Vp, uint32_t B, uint32_t* bonePos /* B */);
/* The SDEF records and per-warp descriptor lists rz_load_sdef builds, device-free: records receives 12
 * floats per evaluated SDEF vertex (C, c0, c1, w0, w1, then the two bone ids as one u32: j0 | j1 << 16 — palette rows once a
 * mesh is loaded), desc [Vp] the descriptor word of every lane (table index | output slot << 24, 0xFFFFFFFF = none),
 * nActive the number of records (vertices with a third or fourth influence are not SDEF and are dropped). */
int32_t rz_plan_sdef(const uint32_t* laneVertex /* Vp */, uint32_t Vp, const uint16_t* joints, const uint8_t* weights, uint32_t V,
                     uint32_t B, const uint32_t* sdefVertIdx /* n */, const float* c_r0_r1 /* 9*n */, uint32_t n,
                     float* records /* 12*n or NULL */, uint32_t* desc /* Vp or NULL */, uint32_t* nActive);
/* The cost-balanced chunk boundaries rz_deform uses when morphs are active: tileDepth[t] = deepest row list among the
 * warps of 256-vertex tile t; tab receives nChunks+1 tile indices (room for nChunksTarget+1). */
int32_t rz_plan_chunks(const uint32_t* tileDepth, uint32_t nTiles, uint32_t tilesPerPass, uint32_t nChunksTarget,
                       uint32_t* tab, uint32_t* nChunks);
int32_t rz_read_bounds(rz_ctx* ctx, uint32_t firstInstance, uint32_t count, float* minmax6 /* 6*count */);
/* Integer view of the tables the KERNEL consumes, read back from the device records and mapped to caller order (parity
 * tests): joints = the palette rows the kernel gathers mapped back to bone ids, weights = the kernel's pre-normalised f32
 * weights (engine.ts:255-258 applied at load) re-quantised round(w*255).  Bit-identical to the input of rz_load_mesh
 * whenever the weights of a vertex sum to 255 — the loader's invariant (pmx-loader.ts:892-938); for other inputs the
 * NORMALISED weights come back (what the shader rule makes of them; all-zero -> 255,0,0,0).  Slots whose weight is zero
 * gather a borrowed row on the device (they cannot influence the result); the caller's own joint index is reported there. */
int32_t rz_read_skinning(rz_ctx* ctx, uint16_t* joints /* 4V */, uint8_t* weights /* 4V */);
/* device-computed skin matrices of palette p as 3x4 row-major (12 floats per bone) */
int32_t rz_read_skin_matrices(rz_ctx* ctx, uint32_t palette, float* skin3x4 /* 12*B */);

/* Bone WORLD matrices of palette p (16 floats per bone, column-major, exactly the layout of Model.getBoneWorldMatrices(),
 * model.ts:317-319) recovered from the device's skin matrices: world = skin * invBind^-1.  For hosts whose pose is
 * evaluated on the device (rz_set_local_rotations / rz_set_instance_clocks) but that need the matrices back — the kinematic
 * half of Physics.step moves bodies to their bones from exactly these (syncFromBones, physics.ts:649-703). */
int32_t rz_read_world_matrices(rz_ctx* ctx, uint32_t palette, float* world16 /* 16*B */);

/* ---- stats: replaces Engine.getStats() (engine.ts:1664-1666) ---- */
int32_t rz_get_stats(rz_ctx* ctx, rz_stats* out);

/* Last error text for ctx (or for the calling thread when ctx == NULL, e.g. after a failed rz_create). */
const char* rz_last_error(rz_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* RZE_B200_H */
