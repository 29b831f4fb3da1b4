// rze_b200.cu — C ABI (include/rze_b200.h) over the sm_100a deform kernels.
//
// Host side of the deform stage: table preprocessing that replaces setupModelBuffers
// (engine.ts:1728-1832), the per-frame palette upload (engine.ts:2383-2389) + skin-matrix
// pass (engine.ts:2393-2402), and the launch logic for the fused deform kernel.
// There is no CPU implementation behind this ABI: without a CUDA device that can run the
// sm_100a image, rz_create fails with RZ_ERR_NO_DEVICE.
#include "../../include/rze_b200.h"
#include "deform_kernel.cuh"
#include "deform2_kernel.cuh"
#include "aux_kernels.cuh"
#include "kernel_table.h"
#include "lane_plan.h"
#include "lane_plan2.h"
#include "mesh_tables.h"

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <cstdlib>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

using namespace rz;

// the ctypes / N-API mirrors of these structs rely on the layout
static_assert(sizeof(rz_config) == 56, "rz_config layout changed: update capi.py / napi shim");
static_assert(sizeof(rz_stats) == 136, "rz_stats layout changed: update capi.py / napi shim");

namespace {

thread_local std::string g_last_error;

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
};

struct rz_ctx_impl {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool ownStream = false;
  uint32_t flags = 0, maxK = 1;
  uint32_t tuneI = 0, tuneStore = 0, tuneThreads = 0, tuneChunks = 0, tuneCtas = 0, tuneVpl = 0;
  int numSM = 0, maxSmemOptin = 0;
  std::string err;

  // host copies of the caller's tables (caller order)
  uint32_t V = 0, B = 0;
  std::vector<float> h_vtx8;
  std::vector<uint16_t> h_joints;
  std::vector<uint8_t> h_weights;
  std::vector<float> h_invBind;
  uint32_t M = 0;
  std::vector<uint32_t> h_moff, h_mvert;
  std::vector<float> h_mdelta;
  std::vector<uint32_t> h_sdefVert;
  std::vector<float> h_sdefVec;
  std::vector<float> h_edgeSize;          // [V] Material.edgeSize per vertex (RZ_FLAG_OUTLINE), empty = all zero
  DevBuf d_edge, d_uv;
  size_t hullOffF = 0;
  bool tablesDirty = true;

  // device tables (processing order)
  uint32_t Vp = 0, nTiles = 0;
  std::vector<uint32_t> procToVertex;   // processing index -> caller vertex id (or ~0u for padding)
  uint64_t packFastSlots = 0, packTotalSlots = 0;   // pair packing: warp x influence-slot gathers on the fast path / all
  std::vector<uint8_t> procSlotMap;     // [Vp][4] device influence slot holding the caller's influence k (pair packing permutes them)
  std::vector<uint32_t> vorder, vinv;    // stored position -> caller vertex id and its inverse (identity unless RZ_FLAG_REORDER_VERTICES)
  std::vector<uint32_t> bonePos, boneAt; // palette row of bone b (bank-aware permutation) and its inverse
  DevBuf d_bonePos;
  // experiment knobs (environment RZ_PERM / RZ_COLOR / RZ_LAYOUT, see DESIGN.md "tuning knobs"):
  //   permMode   2: lane PAIRS + influence slots packed so that aligned lane pairs gather the same palette rows (default);
  //              1: lanes of a warp sorted by (influence count, bones), caller's slot order; 0: natural lane order
  //   colorMode  1: bank-aware palette permutation (8-colouring of the bone co-occurrence graph); 0: identity
  //   layoutMode 0: palette rows of 48 B ([B][3] float4); 1: [3][B] float4 + co-occurrence clustering (measured slower)
  int permMode = 2, colorMode = 1, layoutMode = 0;       // vertex ordering inside a tile / bank-aware palette permutation
  DevBuf d_rec0, d_rec1, d_rec2, d_meta, d_mrange, d_ments, d_sdefIdx, d_sdefTab, d_invBind;
  // two-vertices-per-lane table of the plain path (lane_plan2.h / deform2_kernel.cuh); built when the output layout is the
  // plain planar one (no fused consumer flags); vplMode: RZ_VPL=0 disables it (A/B measurements)
  DevBuf d_rec2v;
  uint32_t vgCount = 0;                   // vertex groups (32 lanes each)
  bool vpl2Ready = false;
  int vplMode = 1;
  std::vector<uint32_t> p2VertA, p2VertB; // stored vertex evaluated on side A / B of lane L (~0u: none)
  uint64_t p2FastSlots = 0, p2Slots = 0;
  uint32_t morphNnz = 0, sdefActive = 0;
  std::vector<uint32_t> tileMorphMax;     // [nTiles] most morph entries on one vertex of the tile (chunk balancing)
  DevBuf d_chunkTab;
  struct ChunkKey { uint32_t tilesPerPass = 0, target = 0; } chunkKey;
  uint32_t chunkCount = 0;

  // skeleton for GPU pose evaluation
  bool haveSkeleton = false, haveTweens = false, haveAnimation = false;
  DevBuf d_trStart, d_trMs, d_trQ, d_trRest;
  uint32_t nLevels = 0;
  DevBuf d_skParent, d_skBindT, d_skAppendParent, d_skAppendRatio, d_skLevelBones, d_skLevelStart, d_skChainStart, d_skChainBones;
  bool useChains = false;
  bool packedMeta = false;
  // physics -> bone feedback (rz_load_rigid_bodies / rz_apply_body_transforms)
  DevBuf d_rbBones, d_rbStart, d_rbIds, d_rbOffInv, d_rbPosQuat;
  uint32_t rbBones = 0, rbBodies = 0;
  DevBuf d_twStart, d_twTarget, d_twRest, d_twStartMs, d_twDurMs, d_twActive, d_localRot, d_localRot2, d_nowMs, d_twAux, d_invBindSoA;
  // local-rotation uploads alternate between two device buffers on the copy stream, so frame n+1's rotations travel while
  // frame n is still being deformed; evRotFree[x]: the pose kernel that last read buffer x has been issued
  uint32_t rotCur = 0;
  cudaEvent_t evRotFree[2] = {nullptr, nullptr}, evRotCopied = nullptr;
  bool rotFreeValid[2] = {false, false};

  // per-frame
  uint32_t P = 0, K = 0;
  DevBuf d_quat;
  DevBuf d_world, d_skin, d_inst2pal, d_mwIn, d_mwIds, d_mwDense, d_out, d_bounds, d_counter;
  // RZ_FLAG_DOUBLE_BUFFER: a second result buffer; every palette update (= a new frame) flips `cur`, so a consumer --
  // rz_read_instance_async on the read stream, or a renderer holding rz_get_output_layout's pointer -- keeps reading
  // frame n while frame n+1 is being written
  DevBuf d_out2;
  uint32_t cur = 0;
  cudaStream_t readStream = nullptr;
  cudaEvent_t evReadReady = nullptr, evReadDone[2] = {nullptr, nullptr};   // per result buffer
  bool readPending[2] = {false, false};
  bool haveInst2pal = false, palettesSet = false;
  uint32_t Mact = 0, Mpad = 4;
  // Pinned host memory.  Every buffer carries the event of the last asynchronous copy that reads it; nobody (library or
  // caller) rewrites a buffer before that event has completed (pinned_acquire).
  //   stage[2]   : caller-visible palette staging, handed out alternately by rz_palette_staging (the producer fills one
  //                while the upload of the other is in flight)
  //   big        : library-owned copy of a pageable `world` / `quats` argument (never the caller's staging)
  //   scratch[4] : ring of small library-owned buffers for index / weight / clock uploads; a ring so that consecutive
  //                calls of one frame do not wait for each other's copies (which sit behind the previous deform)
  struct PinnedBuf { void* p = nullptr; size_t bytes = 0; cudaEvent_t done = nullptr; bool pending = false; };
  PinnedBuf stage[2], big, scratch[4];
  uint32_t stageCur = 0, scratchCur = 0;
  bool noPipeline = false;            // RZ_NO_PIPELINE (read once at rz_create)
  uint32_t pipelineBlock = 0;         // RZ_PIPELINE_BLOCK: palettes per upload block (0 = ~16 MB)
  size_t instStrideF = 0, nrmOffF = 0;

  // pipelined palette upload (rz_set_palettes with host matrices): blocks of palettes travel on a copy stream while the
  // blocks before them are already being deformed; a block's skin-matrix pass is issued by rz_deform right before the
  // deform launch of that block's instances
  cudaStream_t copyStream = nullptr;
  cudaEvent_t evWorldFree = nullptr;     // main stream: the last skin-matrix pass that reads d_world has been issued
  bool worldFreeValid = false;
  // palettes whose skin matrices are not computed yet: blocks of a pipelined upload (ev = completion of the block's copy on
  // the copy stream) or a whole resident set (ev = nullptr: rz_set_palettes_device / unpipelined rz_set_palettes; the pass
  // is issued by the next rz_deform, inside its CUDA graph, or by whichever other entry point needs the matrices first)
  struct PendBlock { uint32_t pal0, n; cudaEvent_t ev; const float* src; };
  std::vector<PendBlock> pend;
  std::vector<cudaEvent_t> evPool;

  // CUDA graphs of recorded frames (rz_deform)
  struct GraphKey {
    const void* fn; uint32_t grid, nt; size_t smem; DeformParams prm; Deform2Params prm2; int feat; uint32_t first, count, P;
    const void* invBind; const void* bonePos; int layoutMode;
    uint32_t nPend; const void* pendSrc[4]; uint32_t pendPal0[4], pendN[4];
  };
  struct GraphEntry { GraphKey key; cudaGraphExec_t exec = nullptr; uint32_t nodes = 0; };
  std::vector<GraphEntry> graphs;
  bool useGraphs = true;               // RZ_NO_GRAPH=1: direct launches (debugging, profilers that dislike graphs)
  uint64_t graphLaunches = 0;

  // stats
  cudaEvent_t evStart = nullptr, evStop = nullptr;
  bool evPending = false;
  double lastMs = 0, lastAlgBytes = 0;
  uint64_t lastVerts = 0;
  std::vector<double> msRing;
  std::vector<double> frameStamps;
  uint64_t frames = 0, launches = 0;
  size_t devBytes = 0;
  uint32_t usedI = 0, usedStore = 0, usedCtas = 0, usedThreads = 0, usedSmem = 0, usedVpl = 1;
  // launch-shape cache: the selection below (lookup + occupancy query) only depends on these
  struct ShapeKey { int feat = -1; uint32_t B = 0, Mpad = 0, countClass = 0; bool v2Allowed = false; } shapeKey;
  bool shapeV2 = false;
  KernelEntry shapeKe{nullptr, 0, 0, 0, 0, 0, 0};
  size_t shapeSmem = 0;
  int shapeOcc = 0;
};

int fail(rz_ctx_impl* c, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c) c->err = buf;
  g_last_error = buf;
  return code;
}

#define CU_TRY(c, expr)                                                                              \
  do {                                                                                               \
    cudaError_t e_ = (expr);                                                                         \
    if (e_ != cudaSuccess)                                                                           \
      return fail((c), e_ == cudaErrorMemoryAllocation ? RZ_ERR_OOM : RZ_ERR_CUDA, "%s: %s (%s:%d)", #expr, \
                  cudaGetErrorString(e_), __FILE__, __LINE__);                                       \
  } while (0)

int dev_reserve(rz_ctx_impl* c, DevBuf& b, size_t bytes) {
  if (bytes <= b.bytes && b.p) return RZ_OK;
  if (b.p) {
    cudaFree(b.p);
    c->devBytes -= b.bytes;
    b.p = nullptr;
    b.bytes = 0;
  }
  if (bytes == 0) bytes = 16;
  cudaError_t e = cudaMalloc(&b.p, bytes);
  if (e != cudaSuccess) {
    b.p = nullptr;
    return fail(c, e == cudaErrorMemoryAllocation ? RZ_ERR_OOM : RZ_ERR_CUDA, "cudaMalloc(%zu bytes): %s", bytes,
                cudaGetErrorString(e));
  }
  b.bytes = bytes;
  c->devBytes += bytes;
  return RZ_OK;
}

void dev_free(rz_ctx_impl* c, DevBuf& b) {
  if (b.p) {
    cudaFree(b.p);
    c->devBytes -= b.bytes;
  }
  b.p = nullptr;
  b.bytes = 0;
}

// Wait until the last asynchronous copy sourced from `b` has completed, then make sure it holds `bytes`.
int pinned_acquire(rz_ctx_impl* c, rz_ctx_impl::PinnedBuf& b, size_t bytes) {
  if (b.pending) {
    CU_TRY(c, cudaEventSynchronize(b.done));
    b.pending = false;
  }
  if (!b.done) CU_TRY(c, cudaEventCreateWithFlags(&b.done, cudaEventDisableTiming));
  if (bytes <= b.bytes && b.p) return RZ_OK;
  if (b.p) cudaFreeHost(b.p);
  b.p = nullptr;
  b.bytes = 0;
  CU_TRY(c, cudaMallocHost(&b.p, std::max<size_t>(bytes, 4096)));
  b.bytes = std::max<size_t>(bytes, 4096);
  return RZ_OK;
}
// The copies queued on `s` so far are the last readers of `b`.
int pinned_release(rz_ctx_impl* c, rz_ctx_impl::PinnedBuf& b, cudaStream_t s) {
  CU_TRY(c, cudaEventRecord(b.done, s));
  b.pending = true;
  return RZ_OK;
}
void pinned_free(rz_ctx_impl::PinnedBuf& b) {
  if (b.pending) cudaEventSynchronize(b.done);
  if (b.done) cudaEventDestroy(b.done);
  if (b.p) cudaFreeHost(b.p);
  b = rz_ctx_impl::PinnedBuf();
}
// the caller's staging buffer that holds [src, src+bytes), or nullptr
rz_ctx_impl::PinnedBuf* staging_of(rz_ctx_impl* c, const void* src, size_t bytes) {
  const char* s0 = reinterpret_cast<const char*>(src);
  for (auto& b : c->stage)
    if (b.p && s0 >= (char*)b.p && s0 + bytes <= (char*)b.p + b.bytes) return &b;
  return nullptr;
}
// H2D copy of a small pageable array through the scratch ring (no stream synchronisation)
int upload_small(rz_ctx_impl* c, void* dst, const void* src, size_t bytes, const void* src2 = nullptr, void* dst2 = nullptr, size_t bytes2 = 0) {
  rz_ctx_impl::PinnedBuf& b = c->scratch[c->scratchCur];
  c->scratchCur = (c->scratchCur + 1) % 4;
  int rc;
  if ((rc = pinned_acquire(c, b, bytes + bytes2))) return rc;
  memcpy(b.p, src, bytes);
  CU_TRY(c, cudaMemcpyAsync(dst, b.p, bytes, cudaMemcpyHostToDevice, c->stream));
  if (bytes2) {
    memcpy((char*)b.p + bytes, src2, bytes2);
    CU_TRY(c, cudaMemcpyAsync(dst2, (char*)b.p + bytes, bytes2, cudaMemcpyHostToDevice, c->stream));
  }
  return pinned_release(c, b, c->stream);
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// cudaFuncAttributeMaxDynamicSharedMemorySize is a property of the FUNCTION on a device, shared by every context of the
// process: it is only ever raised (a smaller request of another context must not undercut a launch already planned), and
// the driver is only called when it actually rises -- not once per launch.
cudaError_t raise_smem_limit(int device, const void* fn, size_t bytes) {
  static std::mutex mu;
  static std::unordered_map<uint64_t, size_t> cur;
  const uint64_t key = (uint64_t)reinterpret_cast<uintptr_t>(fn) * 64u + (uint64_t)(device & 63);
  std::lock_guard<std::mutex> lk(mu);
  auto it = cur.find(key);
  if (it != cur.end() && it->second >= bytes) return cudaSuccess;
  if (bytes <= 48 * 1024 && it == cur.end()) { cur[key] = 48 * 1024; return cudaSuccess; }   // the default limit
  const cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) cur[key] = bytes;
  return e;
}

KernelEntry lookup_kernel(int feat, int I, int NT, int MINB) {
  switch (feat) {
#define RZ_CASE(f) case f: return lookup_feat_##f(I, NT, MINB);
    RZ_FEAT_LIST(RZ_CASE)
#undef RZ_CASE
    default: break;
  }
  KernelEntry none{nullptr, 0, 0, 0, 0, 0, feat};
  return none;
}

// smallest compiled feature set that covers `need` (GPAL / NONRM must match exactly: they change data paths)
int resolve_feat(int need) {
  static const int compiled[] = {
#define RZ_V(f) f,
      RZ_FEAT_LIST(RZ_V)
#undef RZ_V
  };
  int best = -1, bestBits = 99;
  const int exact = FEAT_GPAL | FEAT_NONRM | FEAT_HULL | FEAT_ILV;
  for (int cnd : compiled) {
    if ((cnd & need) != need) continue;
    if ((cnd & exact) != (need & exact)) continue;
    const int bits = __builtin_popcount((unsigned)cnd);
    if (bits < bestBits) { bestBits = bits; best = cnd; }
  }
  return best;
}

size_t smem_needed(int I, int NT, int feat, uint32_t B, uint32_t Mpad, int nbuf = 2) {
  size_t s = kCtrlBytes;
  if (!(feat & FEAT_GPAL)) s += (size_t)I * B * 48;
  if (feat & FEAT_MORPH) s += (size_t)I * Mpad * 4;
  if ((feat & FEAT_SDEF) && !(feat & FEAT_GPAL)) s += (size_t)I * B * 16;   // per-bone quaternions
  // SDEF skews every per-instance region by 16 bytes (deform_kernel.cuh kSkew): palette, quaternions, staging
  if (feat & FEAT_SDEF) s += (size_t)(I - 1) * 16 * ((feat & FEAT_GPAL) ? 0 : 2) + (size_t)nbuf * (I - 1) * 16 * (NT / 32);
  const size_t vtxBytes = (feat & FEAT_ILV) ? 32 : (((feat & FEAT_NONRM) ? 1 : 2) + ((feat & FEAT_HULL) ? 1 : 0)) * 12;
  s += (size_t)nbuf * I * NT * vtxBytes;                                // warp-private staging, laid out like the output
  return s;
}

// deform2_kernel: control block, I palettes, per warp NBUF sub-batch buffers of 64 vertices x (bytes per vertex of the layout)
size_t smem_needed2(int I, int NT, uint32_t B, int SB, int NBUF, int out2) {
  static const size_t instBytes[5] = {64 * 24, 64 * 12, 64 * 36, 64 * 32, 64 * 24};   // OUT2_PLANAR, _NONRM, _HULL, _ILV, _BOUNDS
  return kCtrlBytes + (size_t)I * B * 48 + (size_t)(NT / 32) * NBUF * SB * instBytes[out2];
}
KernelEntry lookup_v2(int out2, int I, int NT, int MINB, int SB) {
  switch (out2) {
    case OUT2_PLANAR: return lookup_v2_out0(I, NT, MINB, SB);
    case OUT2_NONRM: return lookup_v2_out1(I, NT, MINB, SB);
    case OUT2_HULL: return lookup_v2_out2(I, NT, MINB, SB);
    case OUT2_BOUNDS: return lookup_v2_out4(I, NT, MINB, SB);
    default: return lookup_v2_out3(I, NT, MINB, SB);
  }
}

// How to cut a launch into work items.  Items are pulled from an atomic queue by `grid` persistent CTAs, so the launch ends
// when the last CTA finishes.  Few, long items leave a tail (683 instance groups on 148 CTAs: 5 rounds for 4.6 rounds of
// work); many short ones pay the fixed cost of an item (palette staging + two CTA-wide barriers, o ~ 2.5 us) over and over.
// So: the first `coarse` instance groups are one item each, the rest are cut into `chunks` pieces that level the tail
// ("long items first").  Both are chosen by the estimate  rounds1 * (T + o) + ceil(rest * c / grid) * (T / c + o),
// T = one whole instance group on one CTA (measured on B200: within ~1 % of an exhaustive sweep, profiles/r02_chunk_sweep.txt).
struct ItemPlan { uint32_t coarse, chunks; };
ItemPlan pick_items(uint32_t nGroups, uint32_t grid, uint32_t V, int I, uint32_t maxChunks, bool twoLevel = true) {
  const double T = (double)V * I * 0.6e-9, o = 2.5e-6 / std::max(T, 1e-9);
  ItemPlan best{0, 1};
  double bestCost = 1e300;
  static const uint32_t cs[] = {1, 2, 3, 4, 5, 6, 8, 10, 12, 16};
  // (one-vertex-per-lane kernels: uniform items only -- the two-level plan measured slower there, profiles/r02_chunk_sweep.txt)
  const uint32_t fullRounds = twoLevel ? nGroups / std::max(grid, 1u) : 0u;
  for (uint32_t r1 = 0; r1 <= fullRounds; ++r1) {
    const uint32_t rest = nGroups - r1 * grid;
    for (uint32_t c : cs) {
      if (c > std::max(1u, maxChunks)) break;
      if (rest == 0 && c > 1) break;
      const double cost = r1 * (1.0 + o) + std::ceil((double)rest * c / grid) * (1.0 / c + o);
      if (cost < bestCost * 0.9995) { bestCost = cost; best = ItemPlan{r1 * grid, c}; }
    }
  }
  return best;
}

// ---- table preprocessing ------------------------------------------------------------------------
// Tile = 256 consecutive vertices (work-item granularity).  A warp owns 32 consecutive vertices; which LANE evaluates
// which of them is chosen at load time (see below).  Staging writes of a warp always hit 32 distinct banks
// (3*slot mod 32 is a bijection on the 32 slots).
int rebuild_tables(rz_ctx_impl* c) {
  const uint32_t V = c->V, B = c->B;
  const uint32_t nTiles = (V + kTile - 1) / kTile;
  const uint32_t Vp = nTiles * kTile;
  c->nTiles = nTiles;
  c->Vp = Vp;

  // ---- stored vertex order ------------------------------------------------------------------------------------
  // Default: the caller's order (drop-in: output vertex i == input vertex i).  With RZ_FLAG_REORDER_VERTICES the library
  // stores vertices sorted by their bone tuple, so that the lanes of a warp (and in particular every aligned lane pair)
  // gather the SAME palette rows: the shared-memory broadcast fast path then serves almost every gather instruction
  // (profiles/r01_ubench_lds_row_fetch.txt).  The caller remaps its index buffer once with rz_get_vertex_order().
  c->vorder.resize(V);
  c->vinv.resize(V);
  for (uint32_t v = 0; v < V; ++v) c->vorder[v] = v;
  if (c->flags & RZ_FLAG_REORDER_VERTICES) {
    // key = the vertex' SET of influencing bones (non-zero weights, ascending ids): vertices that gather the same rows
    // become neighbours whatever order their influences are listed in (the slot order is the packer's to choose, below)
    auto tuple_key = [&](uint32_t v) {
      const uint8_t* w = &c->h_weights[(size_t)v * 4];
      const uint16_t* j = &c->h_joints[(size_t)v * 4];
      uint16_t b[4];
      uint32_t n = 0;
      for (uint32_t k = 0; k < 4; ++k) if (w[k]) b[n++] = j[k];
      if (n == 0) b[n++] = j[0];
      std::sort(b, b + n);
      uint64_t key = (uint64_t)n << 60;
      for (uint32_t k = 0; k < n; ++k) key |= (uint64_t)(b[k] & 0x7FFF) << (45 - 15 * k);
      return key;
    };
    std::vector<uint64_t> keys(V);
    for (uint32_t v = 0; v < V; ++v) keys[v] = tuple_key(v);
    std::stable_sort(c->vorder.begin(), c->vorder.end(), [&](uint32_t x, uint32_t y) { return keys[x] < keys[y]; });
  }
  for (uint32_t i = 0; i < V; ++i) c->vinv[c->vorder[i]] = i;
  std::vector<float> vt_s;
  std::vector<uint16_t> jt_s;
  std::vector<uint8_t> wt_s;
  const float* VT = c->h_vtx8.data();
  const uint16_t* JT = c->h_joints.data();
  const uint8_t* WT = c->h_weights.data();
  if (c->flags & RZ_FLAG_REORDER_VERTICES) {
    vt_s.resize((size_t)V * 8); jt_s.resize((size_t)V * 4); wt_s.resize((size_t)V * 4);
    for (uint32_t i = 0; i < V; ++i) {
      const uint32_t v = c->vorder[i];
      memcpy(&vt_s[(size_t)i * 8], &c->h_vtx8[(size_t)v * 8], 32);
      memcpy(&jt_s[(size_t)i * 4], &c->h_joints[(size_t)v * 4], 8);
      memcpy(&wt_s[(size_t)i * 4], &c->h_weights[(size_t)v * 4], 4);
    }
    VT = vt_s.data(); JT = jt_s.data(); WT = wt_s.data();
  }

  // vertex-major view of the morph table in stored vertex order (mesh_tables.h)
  std::vector<uint32_t> mcount, mstart;
  std::vector<F4> ments;
  const uint32_t nnz = c->M ? c->h_moff[c->M] : 0;
  morphs_by_vertex(V, c->M, c->h_moff.data(), c->h_mvert.data(), c->h_mdelta.data(), c->vinv.data(), mcount, mstart, ments);
  c->morphNnz = nnz;
  c->tileMorphMax.assign(nTiles, 0);       // filled with the row depth per warp once the lane plan is known (`mell` below)
  c->chunkKey = {};

  // SDEF (only when enabled): which vertices take the spherical path.  The table itself needs the palette rows and is
  // built further down, after the bank-aware permutation.
  std::vector<int32_t> sdefOf(V, -1);     // stored vertex -> record n of the caller's rz_load_sdef arrays
  c->sdefActive = 0;
  if (c->flags & RZ_FLAG_SDEF) {
    for (size_t n = 0; n < c->h_sdefVert.size(); ++n) {
      const uint32_t v = c->vinv[c->h_sdefVert[n]];
      const uint8_t* w = &WT[(size_t)v * 4];
      if (w[2] != 0 || w[3] != 0) continue;   // not a two-influence vertex: stays linear
      sdefOf[v] = (int32_t)n;
      }
  }

  std::vector<float4> rec0(Vp), rec1(Vp), rec2(Vp);
  std::vector<uint32_t> metaArr(Vp);
  std::vector<float> edgeArr((c->flags & RZ_FLAG_OUTLINE) ? Vp : 0, 0.f);
  std::vector<float2> uvArr((c->flags & RZ_FLAG_INTERLEAVED) ? Vp : 0, make_float2(0.f, 0.f));
  std::vector<uint2> mrange(Vp / 32);     // per warp: (first entry, depth) of its lane-interleaved morph entries
  std::vector<uint32_t> sdefIdx(Vp, ~0u);   // "no SDEF vertex": a feature kernel compiled with SDEF may run without the flag
  c->procToVertex.assign(Vp, ~0u);

  const bool packMeta = B <= 4096;  // palette rows fit 12 bits: meta rides in the joint words (deform_kernel.cuh)
  c->packedMeta = packMeta;

  // which lane evaluates which vertex, and the influence table the device sees (lane_plan.h: pair packing)
  LanePlan plan;
  {
    // SDEF vertices are planned like any other: their spherical blend runs in the kernel's dense phase from the table
    // below (own bone rows + weights), so the linear slots are the packer's to arrange
    plan_lanes(JT, WT, nullptr, V, B, kTile, c->permMode, plan);
  }
  const std::vector<uint32_t>& procVertex = plan.procVertex;
  const std::vector<uint32_t>& procSlot = plan.procSlot;
  const std::vector<uint16_t>& gatherJ = plan.gatherJ;
  const std::vector<float>& devW = plan.devW;
  const std::vector<uint8_t>& devN = plan.devN;
  c->procSlotMap = plan.slotMap;
  c->packFastSlots = plan.fastSlots;
  c->packTotalSlots = plan.totalSlots;

  // ---- two vertices per lane (lane_plan2.h): the plain path's own table, when this context can run the plain path at all
  LanePlan2 plan2;
  // (the AABB rides on the two-vertex kernel with the planar layout only: with another layout the context runs deform_kernel)
  const bool vpl2 = c->vplMode != 0 && c->layoutMode == 0 &&
                    !((c->flags & RZ_FLAG_BOUNDS) && (c->flags & (RZ_FLAG_NO_NORMALS | RZ_FLAG_OUTLINE | RZ_FLAG_INTERLEAVED)));
  if (vpl2) plan_lanes2(JT, WT, V, B, plan2);
  c->vpl2Ready = false;

  // ---- bank-aware palette permutation (mesh_tables.h): one permutation serves both tables; it follows the gather pattern
  // of the kernel that will run most (the two-vertex plain path unless morphs / SDEF send every launch to the feature kernel)
  const bool colourByV2 = vpl2 && c->M == 0 && !((c->flags & RZ_FLAG_SDEF) && !c->h_sdefVert.empty());
  if (colourByV2) plan_palette_rows(plan2.gatherJ.data(), plan2.nGroups * 32, B, c->colorMode, c->layoutMode, c->bonePos);
  else plan_palette_rows(gatherJ.data(), Vp, B, c->colorMode, c->layoutMode, c->bonePos);
  c->boneAt.assign(B, 0);
  for (uint32_t b = 0; b < B; ++b) c->boneAt[c->bonePos[b]] = b;

  // ---- SDEF records + per-warp descriptor lists (mesh_tables.h)
  SdefTables sdt;
  if (c->flags & RZ_FLAG_SDEF) {
    if (!build_sdef_tables(sdefOf.data(), c->h_sdefVec.data(), JT, WT, c->bonePos.data(), procVertex.data(), procSlot.data(), V, Vp, sdt))
      return fail(c, RZ_ERR_INVALID_ARG, "rz_load_sdef: more than 2^24 SDEF vertices");
    c->sdefActive = sdt.active;
    sdefIdx = sdt.desc;
  }
  const std::vector<F4>& sdefTab = sdt.tab;

  for (uint32_t p = 0; p < Vp; ++p) {
    const uint32_t v = procVertex[p], slot = procSlot[p], ni = devN[p];
    const uint16_t* j = &gatherJ[(size_t)p * 4];
    const float* w = &devW[(size_t)p * 4];
    const uint32_t q0 = c->bonePos[j[0]], q1 = c->bonePos[j[1]], q2 = c->bonePos[j[2]], q3 = c->bonePos[j[3]];   // palette rows
    const bool real = v != ~0u;
    const bool hasMorph = real && mcount[v] != 0, hasSdef = real && sdefOf[v] >= 0;
    uint32_t meta = slot | (ni << kMetaNinfShift) | (real ? kMetaValid : 0u) | (hasMorph ? kMetaMorph : 0u) | (hasSdef ? kMetaSdef : 0u);
    uint32_t j01 = q0 | (q1 << 16), j23 = q2 | (q3 << 16);
    if (packMeta) {
      const uint32_t m11 = (slot & 31u) | (ni << 5) | (real ? 0x100u : 0u) | (hasMorph ? 0x200u : 0u) | (hasSdef ? 0x400u : 0u);
      j01 = q0 | (q1 << 12) | ((m11 & 0xFFu) << 24);
      j23 = q2 | (q3 << 12) | ((m11 >> 8) << 24);
    }
    float j01f, j23f;
    memcpy(&j01f, &j01, 4);
    memcpy(&j23f, &j23, 4);
    if (real) {
      const float* x = &VT[(size_t)v * 8];
      rec0[p] = make_float4(x[0], x[1], x[2], w[0]);
      rec1[p] = make_float4(x[3], x[4], x[5], w[1]);
      // engine.ts:459-460: worldNormal * material.edgeSize * scaleFactor, scaleFactor = 0.01
      if (!edgeArr.empty() && !c->h_edgeSize.empty()) edgeArr[p] = c->h_edgeSize[c->vorder[v]] * 0.01f;
      if (!uvArr.empty()) uvArr[p] = make_float2(x[6], x[7]);
      c->procToVertex[p] = v;
    } else {
      // padding: a harmless rigid vertex parked on an unused output slot of this warp
      rec0[p] = make_float4(0.f, 0.f, 0.f, w[0]);
      rec1[p] = make_float4(0.f, 0.f, 0.f, w[1]);
    }
    rec2[p] = make_float4(w[2], w[3], j01f, j23f);
    metaArr[p] = meta;
  }

  // ---- records of the two-vertex table: 6 float4 planes, SoA over lanes (deform2_kernel.cuh kRec2Planes)
  std::vector<float4> rec2v;
  if (vpl2) {
    const uint32_t L = plan2.nGroups * 32;
    rec2v.assign((size_t)kRec2Planes * L, make_float4(0.f, 0.f, 0.f, 0.f));
    for (uint32_t p = 0; p < L; ++p) {
      const uint32_t g = p / 32, vA = plan2.vertA[p], vB = plan2.vertB[p];
      const uint16_t* j = &plan2.gatherJ[(size_t)p * 4];
      const float* wa = &plan2.wA[(size_t)p * 4];
      const float* wb = &plan2.wB[(size_t)p * 4];
      const uint32_t r01 = c->bonePos[j[0]] | (c->bonePos[j[1]] << 16), r23 = c->bonePos[j[2]] | (c->bonePos[j[3]] << 16);
      const uint32_t meta = (uint32_t)plan2.slotA[p] | ((uint32_t)plan2.slotB[p] << kM2SlotB) | ((uint32_t)plan2.laneN[p] << kM2N) |
                            (vA != ~0u ? kM2HasA : 0u) | (vB != ~0u ? kM2HasB : 0u) | (plan2.groupCount[g] << kM2Cnt);
      const uint32_t first = plan2.groupFirst[g];
      float4 q0 = make_float4(0.f, 0.f, 0.f, wa[0]), q1 = make_float4(0.f, 0.f, 0.f, wa[1]);
      float4 q2 = make_float4(0.f, 0.f, 0.f, wb[0]), q3 = make_float4(0.f, 0.f, 0.f, wb[1]);
      if (vA != ~0u) { const float* x = &VT[(size_t)vA * 8]; q0.x = x[0]; q0.y = x[1]; q0.z = x[2]; q1.x = x[3]; q1.y = x[4]; q1.z = x[5]; }
      if (vB != ~0u) { const float* x = &VT[(size_t)vB * 8]; q2.x = x[0]; q2.y = x[1]; q2.z = x[2]; q3.x = x[3]; q3.y = x[4]; q3.z = x[5]; }
      float4 q5;
      memcpy(&q5.x, &r01, 4); memcpy(&q5.y, &r23, 4); memcpy(&q5.z, &meta, 4); memcpy(&q5.w, &first, 4);
      rec2v[p] = q0; rec2v[(size_t)L + p] = q1; rec2v[2 * (size_t)L + p] = q2; rec2v[3 * (size_t)L + p] = q3;
      rec2v[4 * (size_t)L + p] = make_float4(wa[2], wa[3], wb[2], wb[3]);
      rec2v[5 * (size_t)L + p] = q5;
      float4 q6 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c->flags & RZ_FLAG_INTERLEAVED) {                  // texture coordinates, passed through (engine.ts:273)
        if (vA != ~0u) { q6.x = VT[(size_t)vA * 8 + 6]; q6.y = VT[(size_t)vA * 8 + 7]; }
        if (vB != ~0u) { q6.z = VT[(size_t)vB * 8 + 6]; q6.w = VT[(size_t)vB * 8 + 7]; }
      } else if ((c->flags & RZ_FLAG_OUTLINE) && !c->h_edgeSize.empty()) {   // engine.ts:459-460: edgeSize * 0.01
        if (vA != ~0u) q6.x = c->h_edgeSize[c->vorder[vA]] * 0.01f;
        if (vB != ~0u) q6.y = c->h_edgeSize[c->vorder[vB]] * 0.01f;
      }
      rec2v[6 * (size_t)L + p] = q6;
    }
    c->vgCount = plan2.nGroups;
    c->p2VertA = plan2.vertA; c->p2VertB = plan2.vertB;
    c->p2FastSlots = plan2.fastSlots; c->p2Slots = plan2.slots;
  }

  // ---- morph rows, lane-interleaved per warp (mesh_tables.h)
  MorphRows mrows;
  build_morph_rows(procVertex.data(), Vp, mcount, mstart, ments, mrows);
  for (uint32_t w = 0; w < Vp / 32; ++w) {
    mrange[w] = make_uint2(mrows.first[w], mrows.depth[w]);
    c->tileMorphMax[w * 32 / kTile] = std::max(c->tileMorphMax[w * 32 / kTile], mrows.depth[w]);
  }
  const std::vector<F4>& mell = mrows.rows;

  int rc;
  if ((rc = dev_reserve(c, c->d_rec0, (size_t)Vp * 16))) return rc;
  if ((rc = dev_reserve(c, c->d_rec1, (size_t)Vp * 16))) return rc;
  if ((rc = dev_reserve(c, c->d_rec2, (size_t)Vp * 16))) return rc;
  if ((rc = dev_reserve(c, c->d_meta, (size_t)Vp * 4))) return rc;
  if ((rc = dev_reserve(c, c->d_mrange, (size_t)(Vp / 32) * 8))) return rc;
  if ((rc = dev_reserve(c, c->d_ments, mell.size() * 16))) return rc;
  if ((rc = dev_reserve(c, c->d_sdefIdx, (size_t)Vp * 4))) return rc;
  if ((rc = dev_reserve(c, c->d_sdefTab, std::max<size_t>(sdefTab.size(), 3) * 16))) return rc;
  if ((rc = dev_reserve(c, c->d_invBind, (size_t)B * 64))) return rc;
  if ((rc = dev_reserve(c, c->d_bonePos, (size_t)B * 4))) return rc;
  CU_TRY(c, cudaMemcpyAsync(c->d_rec0.p, rec0.data(), (size_t)Vp * 16, cudaMemcpyHostToDevice, c->stream));
  CU_TRY(c, cudaMemcpyAsync(c->d_rec1.p, rec1.data(), (size_t)Vp * 16, cudaMemcpyHostToDevice, c->stream));
  CU_TRY(c, cudaMemcpyAsync(c->d_rec2.p, rec2.data(), (size_t)Vp * 16, cudaMemcpyHostToDevice, c->stream));
  CU_TRY(c, cudaMemcpyAsync(c->d_meta.p, metaArr.data(), (size_t)Vp * 4, cudaMemcpyHostToDevice, c->stream));
  CU_TRY(c, cudaMemcpyAsync(c->d_mrange.p, mrange.data(), (size_t)(Vp / 32) * 8, cudaMemcpyHostToDevice, c->stream));
  CU_TRY(c, cudaMemcpyAsync(c->d_ments.p, mell.data(), mell.size() * 16, cudaMemcpyHostToDevice, c->stream));
  CU_TRY(c, cudaMemcpyAsync(c->d_sdefIdx.p, sdefIdx.data(), (size_t)Vp * 4, cudaMemcpyHostToDevice, c->stream));
  if (!sdefTab.empty())
    CU_TRY(c, cudaMemcpyAsync(c->d_sdefTab.p, sdefTab.data(), sdefTab.size() * 16, cudaMemcpyHostToDevice, c->stream));
  CU_TRY(c, cudaMemcpyAsync(c->d_invBind.p, c->h_invBind.data(), (size_t)B * 64, cudaMemcpyHostToDevice, c->stream));
  std::vector<float> ibSoA((size_t)B * 16);                     // [4][B] float4: column c of bone b at c*B + b (pose kernels)
  for (uint32_t b = 0; b < B; ++b)
    for (uint32_t col = 0; col < 4; ++col) memcpy(&ibSoA[((size_t)col * B + b) * 4], &c->h_invBind[(size_t)b * 16 + col * 4], 16);
  if ((rc = dev_reserve(c, c->d_invBindSoA, (size_t)B * 64))) return rc;
  CU_TRY(c, cudaMemcpyAsync(c->d_invBindSoA.p, ibSoA.data(), (size_t)B * 64, cudaMemcpyHostToDevice, c->stream));
  CU_TRY(c, cudaMemcpyAsync(c->d_bonePos.p, c->bonePos.data(), (size_t)B * 4, cudaMemcpyHostToDevice, c->stream));
  if (!edgeArr.empty()) {
    if ((rc = dev_reserve(c, c->d_edge, (size_t)Vp * 4))) return rc;
    CU_TRY(c, cudaMemcpyAsync(c->d_edge.p, edgeArr.data(), (size_t)Vp * 4, cudaMemcpyHostToDevice, c->stream));
  }
  if (!uvArr.empty()) {
    if ((rc = dev_reserve(c, c->d_uv, (size_t)Vp * 8))) return rc;
    CU_TRY(c, cudaMemcpyAsync(c->d_uv.p, uvArr.data(), (size_t)Vp * 8, cudaMemcpyHostToDevice, c->stream));
  }
  if (vpl2) {
    if ((rc = dev_reserve(c, c->d_rec2v, rec2v.size() * 16))) return rc;
    CU_TRY(c, cudaMemcpyAsync(c->d_rec2v.p, rec2v.data(), rec2v.size() * 16, cudaMemcpyHostToDevice, c->stream));
    c->vpl2Ready = true;
  }
  CU_TRY(c, cudaStreamSynchronize(c->stream));   // the std::vectors above go out of scope

  // output planes: pos plane then normal plane, both 16-byte aligned (TMA bulk stores)
  const bool nrm = !(c->flags & RZ_FLAG_NO_NORMALS);
  c->nrmOffF = align_up((size_t)V * 3, 4);
  c->instStrideF = nrm ? 2 * c->nrmOffF : c->nrmOffF;
  c->hullOffF = 0;
  if (c->flags & RZ_FLAG_OUTLINE) { c->hullOffF = 2 * c->nrmOffF; c->instStrideF = 3 * c->nrmOffF; }
  if (c->flags & RZ_FLAG_INTERLEAVED) { c->nrmOffF = 3; c->instStrideF = (size_t)V * 8; }   // one stream of 8 f32 per vertex
  if ((rc = dev_reserve(c, c->d_out, (size_t)c->maxK * c->instStrideF * 4))) return rc;
  if ((c->flags & RZ_FLAG_DOUBLE_BUFFER) && (rc = dev_reserve(c, c->d_out2, (size_t)c->maxK * c->instStrideF * 4))) return rc;
  if (c->flags & RZ_FLAG_BOUNDS)
    if ((rc = dev_reserve(c, c->d_bounds, (size_t)c->maxK * 24))) return rc;
  if ((rc = dev_reserve(c, c->d_counter, 16))) return rc;
  c->tablesDirty = false;
  return RZ_OK;
}

int ensure_dense_weights(rz_ctx_impl* c) {
  // dense [maxK][Mpad] table; zero when no weights were set
  const uint32_t Mpad = (uint32_t)align_up(std::max<uint32_t>(c->M, 1), 4);
  if (Mpad != c->Mpad || !c->d_mwDense.p) {
    c->Mpad = Mpad;
    int rc;
    if ((rc = dev_reserve(c, c->d_mwDense, (size_t)c->maxK * Mpad * 4))) return rc;
    CU_TRY(c, cudaMemsetAsync(c->d_mwDense.p, 0, (size_t)c->maxK * Mpad * 4, c->stream));
    c->Mact = 0;
  }
  return RZ_OK;
}

inline float* out_base(rz_ctx_impl* c) { return reinterpret_cast<float*>((c->cur && c->d_out2.p) ? c->d_out2.p : c->d_out.p); }

// a palette update starts a new frame
inline void begin_frame(rz_ctx_impl* c) {
  if (c->flags & RZ_FLAG_DOUBLE_BUFFER) c->cur ^= 1u;
}

void launch_skin_block(rz_ctx_impl* c, const float* d_world, uint32_t pal0, uint32_t n) {
  const uint32_t threads = n * c->B;
  const size_t w4 = (size_t)pal0 * c->B * 4, s4 = (size_t)pal0 * c->B * 3;        // float4 offsets of the block
  skin_matrices_kernel<<<(threads + 127) / 128, 128, 0, c->stream>>>(reinterpret_cast<const float4*>(d_world) + w4,
                                                                    reinterpret_cast<const float4*>(c->d_invBind.p),
                                                                    reinterpret_cast<float4*>(c->d_skin.p) + s4,
                                                                    reinterpret_cast<const uint32_t*>(c->d_bonePos.p), n, c->B, (uint32_t)c->layoutMode);
  c->launches++;
}

// skin-matrix passes of every uploaded block that has none yet (each waits for its block's copy)
int flush_pending(rz_ctx_impl* c) {
  if (c->pend.empty()) return RZ_OK;
  bool readsWorld = false;
  for (const auto& b : c->pend) {
    if (b.ev) CU_TRY(c, cudaStreamWaitEvent(c->stream, b.ev, 0));
    launch_skin_block(c, b.src, b.pal0, b.n);
    readsWorld |= b.src == c->d_world.p;
  }
  c->pend.clear();
  CU_TRY(c, cudaGetLastError());
  if (readsWorld) {
    CU_TRY(c, cudaEventRecord(c->evWorldFree, c->stream));
    c->worldFreeValid = true;
  }
  return RZ_OK;
}

}  // namespace

struct rz_ctx : rz_ctx_impl {};

extern "C" {

uint32_t rz_abi_version(void) { return RZE_B200_ABI_VERSION; }

const char* rz_last_error(rz_ctx* ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }

int32_t rz_create(const rz_config* cfg, rz_ctx** out) {
  if (!cfg || !out) return fail(nullptr, RZ_ERR_INVALID_ARG, "rz_create: null argument");
  if (cfg->struct_size < offsetof(rz_config, tune_instances_per_group))
    return fail(nullptr, RZ_ERR_INVALID_ARG, "rz_create: struct_size %u too small", cfg->struct_size);
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, RZ_ERR_NO_DEVICE, "no CUDA device (%s); this library has no CPU path", cudaGetErrorString(e));
  if (cfg->device < 0 || cfg->device >= ndev)
    return fail(nullptr, RZ_ERR_INVALID_ARG, "device %d out of range (have %d)", cfg->device, ndev);
  if (cfg->max_instances == 0) return fail(nullptr, RZ_ERR_INVALID_ARG, "max_instances must be >= 1");
  if ((cfg->flags & RZ_FLAG_NO_NORMALS) && (cfg->flags & (RZ_FLAG_OUTLINE | RZ_FLAG_INTERLEAVED)))
    return fail(nullptr, RZ_ERR_INVALID_ARG, "RZ_FLAG_NO_NORMALS cannot be combined with RZ_FLAG_OUTLINE / RZ_FLAG_INTERLEAVED (both need the normal)");
  if ((cfg->flags & RZ_FLAG_OUTLINE) && (cfg->flags & RZ_FLAG_INTERLEAVED))
    return fail(nullptr, RZ_ERR_INVALID_ARG, "RZ_FLAG_OUTLINE and RZ_FLAG_INTERLEAVED are separate output layouts: pick one");
  CU_TRY(nullptr, cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  CU_TRY(nullptr, cudaGetDeviceProperties(&prop, cfg->device));
  {
    // the fat binary only holds sm_100a SASS: make sure it is loadable here, loudly
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, reinterpret_cast<const void*>(&skin_matrices_kernel));
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(nullptr, RZ_ERR_NO_DEVICE, "device %d (%s, sm_%d%d) cannot run the sm_100a kernels: %s", cfg->device,
                  prop.name, prop.major, prop.minor, cudaGetErrorString(e));
    }
  }
  rz_ctx* c = new rz_ctx();
  c->device = cfg->device;
  c->flags = cfg->flags;
  c->maxK = cfg->max_instances;
  c->numSM = prop.multiProcessorCount;
  cudaDeviceGetAttribute(&c->maxSmemOptin, cudaDevAttrMaxSharedMemoryPerBlockOptin, cfg->device);
  if (cfg->struct_size >= sizeof(rz_config)) {
    c->tuneI = cfg->tune_instances_per_group;
    c->tuneStore = cfg->tune_store_mode;
    c->tuneThreads = cfg->tune_threads;
    c->tuneChunks = cfg->tune_chunks;
    c->tuneCtas = cfg->tune_ctas_per_sm;
    c->tuneVpl = cfg->tune_vertices_per_lane;
  }
  if (cfg->stream) {
    c->stream = reinterpret_cast<cudaStream_t>(cfg->stream);
  } else {
    e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete c; return fail(nullptr, RZ_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
    c->ownStream = true;
  }
  if (const char* e1 = getenv("RZ_PERM")) c->permMode = atoi(e1);        // experiment knobs (tuning)
  if (const char* e2 = getenv("RZ_COLOR")) c->colorMode = atoi(e2);
  if (const char* e3 = getenv("RZ_LAYOUT")) c->layoutMode = atoi(e3);
  if (const char* e4 = getenv("RZ_VPL")) c->vplMode = atoi(e4);
  c->noPipeline = getenv("RZ_NO_PIPELINE") != nullptr;
  c->useGraphs = getenv("RZ_NO_GRAPH") == nullptr;
  if (const char* eb = getenv("RZ_PIPELINE_BLOCK")) c->pipelineBlock = (uint32_t)std::max(1, atoi(eb));
  e = cudaEventCreate(&c->evStart);
  if (e == cudaSuccess) e = cudaEventCreate(&c->evStop);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->copyStream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->readStream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->evReadReady, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->evReadDone[0], cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->evReadDone[1], cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->evWorldFree, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->evRotFree[0], cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->evRotFree[1], cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->evRotCopied, cudaEventDisableTiming);
  if (e != cudaSuccess) {
    const int code = fail(nullptr, e == cudaErrorMemoryAllocation ? RZ_ERR_OOM : RZ_ERR_CUDA, "rz_create: stream / event creation failed: %s", cudaGetErrorString(e));
    rz_destroy(c);
    return code;
  }
  *out = c;
  return RZ_OK;
}

int32_t rz_destroy(rz_ctx* c) {
  if (!c) return RZ_OK;
  cudaSetDevice(c->device);
  if (c->copyStream) cudaStreamSynchronize(c->copyStream);
  if (c->readStream) { cudaStreamSynchronize(c->readStream); cudaStreamDestroy(c->readStream); }
  if (c->evReadReady) cudaEventDestroy(c->evReadReady);
  for (cudaEvent_t e : c->evReadDone) if (e) cudaEventDestroy(e);
  cudaStreamSynchronize(c->stream);
  for (auto& g : c->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
  for (cudaEvent_t e : c->evPool) cudaEventDestroy(e);
  if (c->evWorldFree) cudaEventDestroy(c->evWorldFree);
  for (cudaEvent_t e : c->evRotFree) if (e) cudaEventDestroy(e);
  if (c->evRotCopied) cudaEventDestroy(c->evRotCopied);
  if (c->copyStream) cudaStreamDestroy(c->copyStream);
  DevBuf* bufs[] = {&c->d_rec0, &c->d_rec1, &c->d_rec2, &c->d_rec2v, &c->d_meta, &c->d_mrange, &c->d_ments, &c->d_sdefIdx, &c->d_sdefTab,
                    &c->d_invBind, &c->d_bonePos, &c->d_world, &c->d_skin, &c->d_inst2pal, &c->d_mwIn, &c->d_mwIds, &c->d_mwDense,
                    &c->d_out, &c->d_out2, &c->d_bounds, &c->d_counter, &c->d_skParent, &c->d_skBindT, &c->d_skAppendParent, &c->d_skAppendRatio,
                    &c->d_skLevelBones, &c->d_skLevelStart, &c->d_skChainStart, &c->d_skChainBones, &c->d_twStart, &c->d_twTarget, &c->d_twRest, &c->d_twStartMs, &c->d_twDurMs,
                    &c->d_twActive, &c->d_localRot, &c->d_localRot2, &c->d_nowMs, &c->d_twAux, &c->d_invBindSoA, &c->d_rbBones, &c->d_rbStart, &c->d_rbIds, &c->d_rbOffInv, &c->d_rbPosQuat, &c->d_trStart, &c->d_trMs, &c->d_trQ, &c->d_trRest, &c->d_quat, &c->d_chunkTab, &c->d_edge, &c->d_uv};
  for (DevBuf* b : bufs) dev_free(c, *b);
  for (auto& b : c->stage) pinned_free(b);
  for (auto& b : c->scratch) pinned_free(b);
  pinned_free(c->big);
  if (c->evStart) cudaEventDestroy(c->evStart);
  if (c->evStop) cudaEventDestroy(c->evStop);
  if (c->ownStream) cudaStreamDestroy(c->stream);
  delete c;
  return RZ_OK;
}

int32_t rz_load_mesh(rz_ctx* c, const float* vtx8, const uint16_t* joints, const uint8_t* weights, uint32_t V,
                     const float* invBind, uint32_t B) {
  if (!c) return fail(nullptr, RZ_ERR_INVALID_ARG, "rz_load_mesh: null ctx");
  if (!vtx8 || !joints || !weights || !invBind) return fail(c, RZ_ERR_INVALID_ARG, "rz_load_mesh: null table");
  if (V == 0 || B == 0) return fail(c, RZ_ERR_INVALID_ARG, "rz_load_mesh: V and B must be > 0 (the reference throws 'Model has no bones')");
  if (B > 65535) return fail(c, RZ_ERR_INVALID_ARG, "rz_load_mesh: B=%u exceeds the joint range (u16 ids 0..65534; 65535 is reserved)", B);
  for (size_t i = 0; i < (size_t)V * 4; ++i)
    if (joints[i] >= B) return fail(c, RZ_ERR_INVALID_ARG, "rz_load_mesh: joint %u of vertex %zu >= bone count %u", joints[i], i / 4, B);
  CU_TRY(c, cudaSetDevice(c->device));
  c->V = V;
  c->B = B;
  c->h_vtx8.assign(vtx8, vtx8 + (size_t)V * 8);
  c->h_joints.assign(joints, joints + (size_t)V * 4);
  c->h_weights.assign(weights, weights + (size_t)V * 4);
  c->h_invBind.assign(invBind, invBind + (size_t)B * 16);
  c->M = 0;
  c->h_moff.clear(); c->h_mvert.clear(); c->h_mdelta.clear();
  c->h_sdefVert.clear(); c->h_sdefVec.clear();
  c->h_edgeSize.clear();
  c->rbBones = c->rbBodies = 0;
  c->palettesSet = false;
  c->haveSkeleton = false;
  c->haveTweens = false;
  c->haveAnimation = false;
  c->Mact = 0;
  c->Mpad = 0;
  c->tablesDirty = true;
  return rebuild_tables(c);
}

int32_t rz_load_morphs(rz_ctx* c, const uint32_t* off, const uint32_t* vertIdx, const float* delta3, uint32_t M) {
  if (!c) return fail(nullptr, RZ_ERR_INVALID_ARG, "rz_load_morphs: null ctx");
  if (c->V == 0) return fail(c, RZ_ERR_STATE, "rz_load_morphs before rz_load_mesh");
  if (M && (!off || off[0] != 0)) return fail(c, RZ_ERR_INVALID_ARG, "rz_load_morphs: morphOffsets[0] must be 0");
  const uint32_t nnz = M ? off[M] : 0;
  if (nnz && (!vertIdx || !delta3)) return fail(c, RZ_ERR_INVALID_ARG, "rz_load_morphs: null entries");
  for (uint32_t m = 0; m < M; ++m)
    if (off[m + 1] < off[m]) return fail(c, RZ_ERR_INVALID_ARG, "rz_load_morphs: offsets not monotone at %u", m);
  for (uint32_t e = 0; e < nnz; ++e)
    if (vertIdx[e] >= c->V) return fail(c, RZ_ERR_INVALID_ARG, "rz_load_morphs: vertex index %u >= V=%u", vertIdx[e], c->V);
  CU_TRY(c, cudaSetDevice(c->device));
  c->M = M;
  if (M) c->h_moff.assign(off, off + M + 1); else c->h_moff.clear();
  c->h_mvert.assign(vertIdx, vertIdx + nnz);
  c->h_mdelta.assign(delta3, delta3 + (size_t)nnz * 3);
  c->Mact = 0;
  c->Mpad = 0;   // forces the dense table to be rebuilt
  c->tablesDirty = true;
  return rebuild_tables(c);
}

int32_t rz_load_sdef(rz_ctx* c, const uint32_t* vertIdx, const float* vec9, uint32_t n) {
  if (!c) return fail(nullptr, RZ_ERR_INVALID_ARG, "rz_load_sdef: null ctx");
  if (c->V == 0) return fail(c, RZ_ERR_STATE, "rz_load_sdef before rz_load_mesh");
  if (n && (!vertIdx || !vec9)) return fail(c, RZ_ERR_INVALID_ARG, "rz_load_sdef: null table");
  for (uint32_t i = 0; i < n; ++i)
    if (vertIdx[i] >= c->V) return fail(c, RZ_ERR_INVALID_ARG, "rz_load_sdef: vertex index %u >= V=%u", vertIdx[i], c->V);
  CU_TRY(c, cudaSetDevice(c->device));
  c->h_sdefVert.assign(vertIdx, vertIdx + n);
  c->h_sdefVec.assign(vec9, vec9 + (size_t)n * 9);
  c->tablesDirty = true;
  return rebuild_tables(c);
}

static int set_palettes_common(rz_ctx* c, const float* d_world, uint32_t P, uint32_t K) {
  int rc;
  if ((rc = dev_reserve(c, c->d_skin, (size_t)P * c->B * 48))) return rc;
  c->pend.clear();                                           // superseded
  begin_frame(c);
  c->pend.push_back({0, P, nullptr, d_world});               // skin-matrix pass: issued lazily (flush_pending / rz_deform's graph)
  c->P = P;
  c->K = K;
  c->palettesSet = true;
  return RZ_OK;
}

int32_t rz_palette_staging(rz_ctx* c, size_t bytes, void** host_ptr) {
  if (!c || !host_ptr) return fail(c, RZ_ERR_INVALID_ARG, "rz_palette_staging: null argument");
  CU_TRY(c, cudaSetDevice(c->device));
  // two buffers, handed out alternately: the one returned now was last uploaded two calls ago; wait for THAT copy only,
  // so the producer refills one buffer while the upload of the other is still in flight
  c->stageCur ^= 1u;
  rz_ctx_impl::PinnedBuf& b = c->stage[c->stageCur];
  int rc = pinned_acquire(c, b, bytes);
  if (rc) return rc;
  *host_ptr = b.p;
  return RZ_OK;
}

int32_t rz_set_palettes(rz_ctx* c, const float* world, uint32_t P, const uint32_t* inst2pal, uint32_t K) {
  if (!c) return fail(nullptr, RZ_ERR_INVALID_ARG, "rz_set_palettes: null ctx");
  if (c->V == 0) return fail(c, RZ_ERR_STATE, "rz_set_palettes before rz_load_mesh");
  if (!world || P == 0 || K == 0) return fail(c, RZ_ERR_INVALID_ARG, "rz_set_palettes: world/P/K must be non-zero");
  if (K > c->maxK) return fail(c, RZ_ERR_INVALID_ARG, "rz_set_palettes: K=%u exceeds max_instances=%u", K, c->maxK);
  if (!inst2pal && P < K) return fail(c, RZ_ERR_INVALID_ARG, "rz_set_palettes: identity mapping needs P >= K (P=%u K=%u)", P, K);
  if (inst2pal)
    for (uint32_t k = 0; k < K; ++k)
      if (inst2pal[k] >= P) return fail(c, RZ_ERR_INVALID_ARG, "rz_set_palettes: instToPalette[%u]=%u >= P=%u", k, inst2pal[k], P);
  CU_TRY(c, cudaSetDevice(c->device));
  const size_t bytes = (size_t)P * c->B * 64;
  int rc;
  if ((rc = dev_reserve(c, c->d_world, bytes))) return rc;
  const char* src = reinterpret_cast<const char*>(world);
  rz_ctx_impl::PinnedBuf* srcBuf = staging_of(c, world, bytes);
  if (!srcBuf) {
    // pageable source: through the library's own pinned copy (waits for the upload that last read it)
    if ((rc = pinned_acquire(c, c->big, bytes))) return rc;
    memcpy(c->big.p, world, bytes);
    src = reinterpret_cast<const char*>(c->big.p);
    srcBuf = &c->big;
  }
  // Large identity-mapped uploads (a crowd with one palette per instance, as the reference's per-model upload scales) are
  // PIPELINED: blocks of ~16 MB go out on the copy stream, each followed by an event; rz_deform then alternates
  // "wait for block b, skin matrices of block b, deform the instances of block b", so the PCIe transfer of block b+1
  // overlaps the deform of block b instead of preceding the whole frame (measured: 4.6 -> 2.9 ms at K=4096, B=512).
  const bool noPipe = c->noPipeline;
  const size_t palBytes = (size_t)c->B * 64;
  uint32_t blk = (uint32_t)std::max<size_t>(64, ((size_t)16 << 20) / palBytes);
  if (c->pipelineBlock) blk = c->pipelineBlock;                    // palettes per block (tests)
  if (!inst2pal && !noPipe && P >= 2 * blk) {
    if ((rc = dev_reserve(c, c->d_skin, (size_t)P * c->B * 48))) return rc;
    if (c->worldFreeValid) CU_TRY(c, cudaStreamWaitEvent(c->copyStream, c->evWorldFree, 0));   // d_world is no longer being read
    c->pend.clear();
    const uint32_t nb = (P + blk - 1) / blk;
    while (c->evPool.size() < nb) {
      cudaEvent_t e;
      CU_TRY(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      c->evPool.push_back(e);
    }
    for (uint32_t b = 0; b < nb; ++b) {
      const uint32_t pal0 = b * blk, n = std::min(blk, P - pal0);
      CU_TRY(c, cudaMemcpyAsync(reinterpret_cast<char*>(c->d_world.p) + (size_t)pal0 * palBytes, src + (size_t)pal0 * palBytes,
                                (size_t)n * palBytes, cudaMemcpyHostToDevice, c->copyStream));
      CU_TRY(c, cudaEventRecord(c->evPool[b], c->copyStream));
      c->pend.push_back({pal0, n, c->evPool[b], reinterpret_cast<const float*>(c->d_world.p)});
    }
    if ((rc = pinned_release(c, *srcBuf, c->copyStream))) return rc;
    c->haveInst2pal = false;
    c->P = P;
    c->K = K;
    c->palettesSet = true;
    begin_frame(c);
    return RZ_OK;
  }
  if (!c->pend.empty() && c->pend.back().ev) CU_TRY(c, cudaStreamWaitEvent(c->stream, c->pend.back().ev, 0));    // an unconsumed pipelined upload still targets d_world
  CU_TRY(c, cudaMemcpyAsync(c->d_world.p, src, bytes, cudaMemcpyHostToDevice, c->stream));
  if ((rc = pinned_release(c, *srcBuf, c->stream))) return rc;
  if (inst2pal) {
    if ((rc = dev_reserve(c, c->d_inst2pal, (size_t)c->maxK * 4))) return rc;
    if ((rc = upload_small(c, c->d_inst2pal.p, inst2pal, (size_t)K * 4))) return rc;
  }
  c->haveInst2pal = inst2pal != nullptr;
  return set_palettes_common(c, reinterpret_cast<const float*>(c->d_world.p), P, K);
}

int32_t rz_set_palettes_device(rz_ctx* c, const float* d_world, uint32_t P, const uint32_t* d_inst2pal, uint32_t K) {
  if (!c) return fail(nullptr, RZ_ERR_INVALID_ARG, "rz_set_palettes_device: null ctx");
  if (c->V == 0) return fail(c, RZ_ERR_STATE, "rz_set_palettes_device before rz_load_mesh");
  if (!d_world || P == 0 || K == 0) return fail(c, RZ_ERR_INVALID_ARG, "rz_set_palettes_device: world/P/K must be non-zero");
  if (K > c->maxK) return fail(c, RZ_ERR_INVALID_ARG, "rz_set_palettes_device: K=%u exceeds max_instances=%u", K, c->maxK);
  if (!d_inst2pal && P < K) return fail(c, RZ_ERR_INVALID_ARG, "rz_set_palettes_device: identity mapping needs P >= K");
  CU_TRY(c, cudaSetDevice(c->device));
  int rc;
  if (d_inst2pal) {
    if ((rc = dev_reserve(c, c->d_inst2pal, (size_t)c->maxK * 4))) return rc;
    CU_TRY(c, cudaMemcpyAsync(c->d_inst2pal.p, d_inst2pal, (size_t)K * 4, cudaMemcpyDeviceToDevice, c->stream));
  }
  c->haveInst2pal = d_inst2pal != nullptr;
  return set_palettes_common(c, d_world, P, K);
}


// ---- GPU pose evaluation ------------------------------------------------------------------------------------------
static int upload(rz_ctx* c, DevBuf& b, const void* src, size_t bytes) {
  int rc;
  if ((rc = dev_reserve(c, b, bytes))) return rc;
  CU_TRY(c, cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, c->stream));
  return RZ_OK;
}

int32_t rz_load_skeleton(rz_ctx* c, const int32_t* parent, const float* bindT, const int32_t* appendParent, const float* appendRatio,
                         const uint8_t* appendRotate, uint32_t B) {
  if (!c) return fail(nullptr, RZ_ERR_INVALID_ARG, "rz_load_skeleton: null ctx");
  if (c->V == 0) return fail(c, RZ_ERR_STATE, "rz_load_skeleton before rz_load_mesh");
  if (!parent || !bindT) return fail(c, RZ_ERR_INVALID_ARG, "rz_load_skeleton: null table");
  if (B != c->B) return fail(c, RZ_ERR_INVALID_ARG, "rz_load_skeleton: B=%u differs from the mesh's bone count %u", B, c->B);
  CU_TRY(c, cudaSetDevice(c->device));
  // depth of every bone (parents first, like the reference's memoised recursion model.ts:340-419); cycles are rejected
  std::vector<int32_t> par(B), depth(B, -1);
  for (uint32_t b = 0; b < B; ++b) par[b] = (parent[b] >= 0 && (uint32_t)parent[b] < B) ? parent[b] : -1;
  for (uint32_t b = 0; b < B; ++b) {
    if (depth[b] >= 0) continue;
    std::vector<uint32_t> chain;
    int32_t cur = (int32_t)b;
    while (cur >= 0 && depth[cur] < 0) {
      chain.push_back((uint32_t)cur);
      if (chain.size() > B) return fail(c, RZ_ERR_INVALID_ARG, "rz_load_skeleton: parent cycle through bone %u", b);
      cur = par[cur];
    }
    int32_t d = cur >= 0 ? depth[cur] + 1 : 0;
    for (size_t i = chain.size(); i-- > 0;) depth[chain[i]] = d++;
  }
  uint32_t nLevels = 0;
  for (uint32_t b = 0; b < B; ++b) nLevels = std::max<uint32_t>(nLevels, (uint32_t)depth[b] + 1);
  std::vector<uint32_t> levelStart(nLevels + 1, 0), levelBones(B);
  for (uint32_t b = 0; b < B; ++b) levelStart[(uint32_t)depth[b] + 1]++;
  for (uint32_t L = 0; L < nLevels; ++L) levelStart[L + 1] += levelStart[L];
  {
    std::vector<uint32_t> fill(levelStart.begin(), levelStart.end() - 1);
    for (uint32_t b = 0; b < B; ++b) levelBones[fill[(uint32_t)depth[b]]++] = b;
  }
  std::vector<int32_t> ap(B, -1);
  std::vector<float> ar(B, 0.f);
  for (uint32_t b = 0; b < B; ++b) {
    const bool rot = appendRotate && appendRotate[b] && appendParent && appendParent[b] >= 0 && (uint32_t)appendParent[b] < B;
    if (!rot) continue;
    float r = appendRatio ? appendRatio[b] : 1.0f;
    if (r != r) r = 1.0f;                                   // "undefined" ratio means 1 (model.ts:360)
    ap[b] = appendParent[b];
    ar[b] = std::max(-1.0f, std::min(1.0f, r));
  }
  int rc;
  if ((rc = upload(c, c->d_skParent, par.data(), (size_t)B * 4))) return rc;
  if ((rc = upload(c, c->d_skBindT, bindT, (size_t)B * 12))) return rc;
  if ((rc = upload(c, c->d_skAppendParent, ap.data(), (size_t)B * 4))) return rc;
  if ((rc = upload(c, c->d_skAppendRatio, ar.data(), (size_t)B * 4))) return rc;
  if ((rc = upload(c, c->d_skLevelBones, levelBones.data(), (size_t)B * 4))) return rc;
  if ((rc = upload(c, c->d_skLevelStart, levelStart.data(), (size_t)(nLevels + 1) * 4))) return rc;
  // ancestor chains (root .. bone) for the barrier-free evaluation; skipped for pathologically deep skeletons
  {
    size_t total = 0;
    for (uint32_t b = 0; b < B; ++b) total += (size_t)depth[b] + 1;
    c->useChains = total <= (size_t)B * 96;
    if (c->useChains) {
      std::vector<uint32_t> chainStart(B + 1, 0), chainBones(total);
      for (uint32_t b = 0; b < B; ++b) chainStart[b + 1] = chainStart[b] + (uint32_t)depth[b] + 1;
      for (uint32_t b = 0; b < B; ++b) {
        uint32_t i = chainStart[b + 1];
        for (int32_t cur = (int32_t)b; cur >= 0; cur = par[cur]) chainBones[--i] = (uint32_t)cur;
      }
      if ((rc = upload(c, c->d_skChainStart, chainStart.data(), (size_t)(B + 1) * 4))) return rc;
      if ((rc = upload(c, c->d_skChainBones, chainBones.data(), total * 4))) return rc;
      CU_TRY(c, cudaStreamSynchronize(c->stream));
    }
  }
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  c->nLevels = nLevels;
  c->haveSkeleton = true;
  return RZ_OK;
}

extern "C++" {
template <int MODE>
static int launch_pose(rz_ctx* c, uint32_t P) {
  begin_frame(c);
  c->pend.clear();   // the palettes are about to be produced on the device: an unconsumed host upload is superseded
  const size_t smem = (size_t)c->B * 64;
  if (smem > (size_t)c->maxSmemOptin)
    return fail(c, RZ_ERR_INVALID_ARG, "GPU pose evaluation supports up to %d bones (B=%u): use rz_set_palettes", c->maxSmemOptin / 64, c->B);
  int rc;
  if ((rc = dev_reserve(c, c->d_skin, (size_t)P * c->B * 48))) return rc;
  PoseSkeleton sk;
  sk.parent = reinterpret_cast<const int32_t*>(c->d_skParent.p);
  sk.bindT = reinterpret_cast<const float*>(c->d_skBindT.p);
  sk.appendParent = reinterpret_cast<const int32_t*>(c->d_skAppendParent.p);
  sk.appendRatio = reinterpret_cast<const float*>(c->d_skAppendRatio.p);
  sk.levelBones = reinterpret_cast<const uint32_t*>(c->d_skLevelBones.p);
  sk.levelStart = reinterpret_cast<const uint32_t*>(c->d_skLevelStart.p);
  sk.nLevels = c->nLevels;
  sk.B = c->B;
  PoseTweens tw;
  tw.start = reinterpret_cast<const float4*>(c->d_twStart.p);
  tw.target = reinterpret_cast<const float4*>(c->d_twTarget.p);
  tw.rest = reinterpret_cast<const float4*>(MODE == 2 ? c->d_trRest.p : c->d_twRest.p);
  tw.startMs = reinterpret_cast<const float*>(c->d_twStartMs.p);
  tw.durMs = reinterpret_cast<const float*>(c->d_twDurMs.p);
  tw.active = reinterpret_cast<const uint8_t*>(c->d_twActive.p);
  PoseTracks tr;
  tr.keyStart = reinterpret_cast<const uint32_t*>(c->d_trStart.p);
  tr.keyMs = reinterpret_cast<const float*>(c->d_trMs.p);
  tr.keyQ = reinterpret_cast<const float4*>(c->d_trQ.p);
  // pointer jumping: ceil(log2(depth)) rounds of one product per bone instead of depth products (aux_kernels.cuh)
  static const int poseAlgo = getenv("RZ_POSE") ? atoi(getenv("RZ_POSE")) : 2;          // 0 levels, 1 chains, 2 jumping
  const size_t smemJump = (size_t)c->B * (c->B <= 1024 ? 68 : 104);
  if (poseAlgo >= 2 && c->nLevels > 4 && smemJump <= (size_t)c->maxSmemOptin && (MODE != 1 || c->d_twAux.p)) {
    uint32_t rounds = 0;
    while ((1u << rounds) < c->nLevels) ++rounds;
    CU_TRY(c, raise_smem_limit(c->device, reinterpret_cast<const void*>(&pose_jump_kernel<MODE>), smemJump));
    const int threads = c->B <= 1024 ? (int)std::max<uint32_t>(64u, (c->B + 31u) / 32u * 32u) : 512;   // one bone per thread when possible
    pose_jump_kernel<MODE><<<P, threads, smemJump, c->stream>>>(sk, tw, tr, reinterpret_cast<const float4*>(c->d_twAux.p),
                                                                reinterpret_cast<const float4*>(c->rotCur ? c->d_localRot2.p : c->d_localRot.p),
                                                                reinterpret_cast<const float*>(c->d_nowMs.p),
                                                                reinterpret_cast<const float4*>(c->d_invBind.p),
                                                                reinterpret_cast<const float4*>(c->d_invBindSoA.p),
                                                                reinterpret_cast<const uint32_t*>(c->d_bonePos.p),
                                                                reinterpret_cast<float4*>(c->d_skin.p), rounds, (uint32_t)c->layoutMode);
    CU_TRY(c, cudaGetLastError());
    c->launches++;
    return RZ_OK;
  }
  if (c->useChains && poseAlgo >= 1) {
    CU_TRY(c, raise_smem_limit(c->device, reinterpret_cast<const void*>(&pose_chain_kernel<MODE>), smem));
    pose_chain_kernel<MODE><<<P, 256, smem, c->stream>>>(sk, tw, tr, reinterpret_cast<const uint32_t*>(c->d_skChainStart.p),
                                                         reinterpret_cast<const uint32_t*>(c->d_skChainBones.p),
                                                         reinterpret_cast<const float4*>(c->rotCur ? c->d_localRot2.p : c->d_localRot.p),
                                                         reinterpret_cast<const float*>(c->d_nowMs.p),
                                                         reinterpret_cast<const float4*>(c->d_invBind.p),
                                                         reinterpret_cast<const uint32_t*>(c->d_bonePos.p),
                                                         reinterpret_cast<float4*>(c->d_skin.p), (uint32_t)c->layoutMode);
    CU_TRY(c, cudaGetLastError());
    c->launches++;
    return RZ_OK;
  }
  CU_TRY(c, raise_smem_limit(c->device, reinterpret_cast<const void*>(&pose_kernel<MODE>), smem));
  pose_kernel<MODE><<<P, 128, smem, c->stream>>>(sk, tw, tr, reinterpret_cast<const float4*>(c->rotCur ? c->d_localRot2.p : c->d_localRot.p),
                                                 reinterpret_cast<const float*>(c->d_nowMs.p),
                                                 reinterpret_cast<const float4*>(c->d_invBind.p),
                                                 reinterpret_cast<const uint32_t*>(c->d_bonePos.p),
                                                 reinterpret_cast<float4*>(c->d_skin.p), nullptr, (uint32_t)c->layoutMode);
  CU_TRY(c, cudaGetLastError());
  c->launches++;
  return RZ_OK;
}
}  // extern "C++"

static int set_mapping(rz_ctx* c, const uint32_t* inst2pal, uint32_t P, uint32_t K, const char* who) {
  if (P == 0 || K == 0) return fail(c, RZ_ERR_INVALID_ARG, "%s: P and K must be non-zero", who);
  if (K > c->maxK) return fail(c, RZ_ERR_INVALID_ARG, "%s: K=%u exceeds max_instances=%u", who, K, c->maxK);
  if (!inst2pal && P < K) return fail(c, RZ_ERR_INVALID_ARG, "%s: identity mapping needs P >= K (P=%u K=%u)", who, P, K);
  if (inst2pal) {
    for (uint32_t k = 0; k < K; ++k)
      if (inst2pal[k] >= P) return fail(c, RZ_ERR_INVALID_ARG, "%s: instToPalette[%u]=%u >= P=%u", who, k, inst2pal[k], P);
    int rc;
    if ((rc = dev_reserve(c, c->d_inst2pal, (size_t)c->maxK * 4))) return rc;
    if ((rc = upload_small(c, c->d_inst2pal.p, inst2pal, (size_t)K * 4))) return rc;
  }
  c->haveInst2pal = inst2pal != nullptr;
  return RZ_OK;
}

int32_t rz_set_local_rotations(rz_ctx* c, const float* quats, uint32_t P, const uint32_t* inst2pal, uint32_t K) {
  if (!c) return fail(nullptr, RZ_ERR_INVALID_ARG, "rz_set_local_rotations: null ctx");
  if (!c->haveSkeleton) return fail(c, RZ_ERR_STATE, "rz_set_local_rotations before rz_load_skeleton");
  if (!quats) return fail(c, RZ_ERR_INVALID_ARG, "rz_set_local_rotations: null quats");
  CU_TRY(c, cudaSetDevice(c->device));
  int rc;
  if ((rc = set_mapping(c, inst2pal, P, K, "rz_set_local_rotations"))) return rc;
  const size_t bytes = (size_t)P * c->B * 16;
  // the rotations go to the device buffer the previous frame did NOT use, on the copy stream: the transfer overlaps the
  // previous frame's deform instead of queueing behind it; the pose kernel waits for the copy, the copy for the pose
  // kernel that last read this buffer (two frames ago)
  c->rotCur ^= 1u;
  DevBuf& rot = c->rotCur ? c->d_localRot2 : c->d_localRot;
  if (bytes > rot.bytes || !rot.p) {
    CU_TRY(c, cudaStreamSynchronize(c->stream));             // growing: nothing may still read the old allocation
    c->rotFreeValid[c->rotCur] = false;
  }
  if ((rc = dev_reserve(c, rot, bytes))) return rc;
  const char* src = reinterpret_cast<const char*>(quats);
  rz_ctx_impl::PinnedBuf* srcBuf = staging_of(c, quats, bytes);
  if (!srcBuf) {
    if ((rc = pinned_acquire(c, c->big, bytes))) return rc;
    memcpy(c->big.p, quats, bytes);
    src = reinterpret_cast<const char*>(c->big.p);
    srcBuf = &c->big;
  }
  if (c->rotFreeValid[c->rotCur]) CU_TRY(c, cudaStreamWaitEvent(c->copyStream, c->evRotFree[c->rotCur], 0));
  CU_TRY(c, cudaMemcpyAsync(rot.p, src, bytes, cudaMemcpyHostToDevice, c->copyStream));
  CU_TRY(c, cudaEventRecord(c->evRotCopied, c->copyStream));
  if ((rc = pinned_release(c, *srcBuf, c->copyStream))) return rc;
  CU_TRY(c, cudaStreamWaitEvent(c->stream, c->evRotCopied, 0));
  if ((rc = launch_pose<0>(c, P))) return rc;
  CU_TRY(c, cudaEventRecord(c->evRotFree[c->rotCur], c->stream));
  c->rotFreeValid[c->rotCur] = true;
  c->P = P; c->K = K; c->palettesSet = true;
  return RZ_OK;
}

int32_t rz_load_rigid_bodies(rz_ctx* c, const int32_t* boneIndex, const uint8_t* dynamic, const float* offsetInverse, uint32_t n) {
  if (!c) return fail(nullptr, RZ_ERR_INVALID_ARG, "rz_load_rigid_bodies: null ctx");
  if (c->V == 0) return fail(c, RZ_ERR_STATE, "rz_load_rigid_bodies before rz_load_mesh");
  if (n && (!boneIndex || !dynamic || !offsetInverse)) return fail(c, RZ_ERR_INVALID_ARG, "rz_load_rigid_bodies: null table");
  CU_TRY(c, cudaSetDevice(c->device));
  // bone -> its dynamic bodies in index order (the reference applies them sequentially, the last valid one wins)
  std::vector<std::vector<uint32_t>> per(c->B);
  for (uint32_t i = 0; i < n; ++i)
    if (dynamic[i] && boneIndex[i] >= 0 && (uint32_t)boneIndex[i] < c->B) per[(uint32_t)boneIndex[i]].push_back(i);
  std::vector<uint32_t> bones, start(1, 0), ids;
  for (uint32_t b = 0; b < c->B; ++b)
    if (!per[b].empty()) {
      bones.push_back(b);
      ids.insert(ids.end(), per[b].begin(), per[b].end());
      start.push_back((uint32_t)ids.size());
    }
  c->rbBones = (uint32_t)bones.size();
  c->rbBodies = n;
  int rc;
  if (c->rbBones) {
    if ((rc = upload(c, c->d_rbBones, bones.data(), bones.size() * 4))) return rc;
    if ((rc = upload(c, c->d_rbStart, start.data(), start.size() * 4))) return rc;
    if ((rc = upload(c, c->d_rbIds, ids.data(), ids.size() * 4))) return rc;
    if ((rc = upload(c, c->d_rbOffInv, offsetInverse, (size_t)n * 64))) return rc;
    CU_TRY(c, cudaStreamSynchronize(c->stream));    // pageable sources
  }
  return RZ_OK;
}

int32_t rz_apply_body_transforms(rz_ctx* c, const float* posQuat, uint32_t P) {
  if (!c) return fail(nullptr, RZ_ERR_INVALID_ARG, "rz_apply_body_transforms: null ctx");
  if (!c->palettesSet) return fail(c, RZ_ERR_STATE, "rz_apply_body_transforms before this frame's palettes were set");
  if (P != c->P) return fail(c, RZ_ERR_INVALID_ARG, "rz_apply_body_transforms: P=%u but the frame has %u palettes", P, c->P);
  if (c->rbBones == 0) return RZ_OK;                  // nothing drives a bone
  if (!posQuat) return fail(c, RZ_ERR_INVALID_ARG, "rz_apply_body_transforms: null transforms");
  CU_TRY(c, cudaSetDevice(c->device));
  int rc;
  if ((rc = flush_pending(c))) return rc;            // a pipelined upload computes its skin matrices lazily: do it now
  const size_t bytes = (size_t)P * c->rbBodies * 28;
  if ((rc = dev_reserve(c, c->d_rbPosQuat, bytes))) return rc;
  if ((rc = upload_small(c, c->d_rbPosQuat.p, posQuat, bytes))) return rc;
  const uint32_t n = P * c->rbBones;
  apply_bodies_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(
      reinterpret_cast<const uint32_t*>(c->d_rbBones.p), reinterpret_cast<const uint32_t*>(c->d_rbStart.p),
      reinterpret_cast<const uint32_t*>(c->d_rbIds.p), reinterpret_cast<const float*>(c->d_rbOffInv.p),
      reinterpret_cast<const float*>(c->d_rbPosQuat.p), reinterpret_cast<const float4*>(c->d_invBind.p),
      reinterpret_cast<const uint32_t*>(c->d_bonePos.p), reinterpret_cast<float4*>(c->d_skin.p), c->rbBones, c->rbBodies, P, c->B,
      (uint32_t)c->layoutMode);
  CU_TRY(c, cudaGetLastError());
  c->launches++;
  return RZ_OK;
}

int32_t rz_set_tweens(rz_ctx* c, const float* startQ, const float* targetQ, const float* startMs, const float* durMs,
                      const uint8_t* active, const float* restQ) {
  if (!c) return fail(nullptr, RZ_ERR_INVALID_ARG, "rz_set_tweens: null ctx");
  if (!c->haveSkeleton) return fail(c, RZ_ERR_STATE, "rz_set_tweens before rz_load_skeleton");
  if (!startQ || !targetQ || !startMs || !durMs || !active || !restQ) return fail(c, RZ_ERR_INVALID_ARG, "rz_set_tweens: null table");
  CU_TRY(c, cudaSetDevice(c->device));
  const size_t B = c->B;
  int rc;
  if ((rc = upload(c, c->d_twStart, startQ, B * 16))) return rc;
  if ((rc = upload(c, c->d_twTarget, targetQ, B * 16))) return rc;
  if ((rc = upload(c, c->d_twRest, restQ, B * 16))) return rc;
  if ((rc = upload(c, c->d_twStartMs, startMs, B * 4))) return rc;
  if ((rc = upload(c, c->d_twDurMs, durMs, B * 4))) return rc;
  if ((rc = upload(c, c->d_twActive, active, B))) return rc;
  // per-bone slerp constants (functions of the two keys only, math.ts:156-189): angle, 1/sin(angle), lerp-branch flag,
  // hemisphere sign -- every pose of the crowd reuses them (aux_kernels.cuh pose_jump_kernel)
  std::vector<float> aux(B * 4);
  for (size_t b = 0; b < B; ++b) {
    const float* a = startQ + b * 4;
    const float* t = targetQ + b * 4;
    float cs = a[0] * t[0] + a[1] * t[1] + a[2] * t[2] + a[3] * t[3];
    float sign = 1.f;
    if (cs < 0.f) { cs = -cs; sign = -1.f; }
    const bool lerp = cs > 0.9995f;
    const float th0 = lerp ? 0.f : acosf(std::min(cs, 1.0f));
    aux[b * 4] = th0; aux[b * 4 + 1] = lerp ? 0.f : 1.0f / sinf(th0); aux[b * 4 + 2] = lerp ? 1.f : 0.f; aux[b * 4 + 3] = sign;
  }
  if ((rc = upload(c, c->d_twAux, aux.data(), B * 16))) return rc;
  CU_TRY(c, cudaStreamSynchronize(c->stream));    // pageable sources
  c->haveTweens = true;
  return RZ_OK;
}

int32_t rz_set_instance_clocks(rz_ctx* c, const float* nowMs, uint32_t P, const uint32_t* inst2pal, uint32_t K) {
  if (!c) return fail(nullptr, RZ_ERR_INVALID_ARG, "rz_set_instance_clocks: null ctx");
  if (!c->haveTweens && !c->haveAnimation) return fail(c, RZ_ERR_STATE, "rz_set_instance_clocks before rz_set_tweens / rz_load_animation");
  if (!nowMs) return fail(c, RZ_ERR_INVALID_ARG, "rz_set_instance_clocks: null clocks");
  CU_TRY(c, cudaSetDevice(c->device));
  int rc;
  if ((rc = set_mapping(c, inst2pal, P, K, "rz_set_instance_clocks"))) return rc;
  if ((rc = dev_reserve(c, c->d_nowMs, (size_t)P * 4))) return rc;
  if (rz_ctx_impl::PinnedBuf* sb = staging_of(c, nowMs, (size_t)P * 4)) {
    CU_TRY(c, cudaMemcpyAsync(c->d_nowMs.p, nowMs, (size_t)P * 4, cudaMemcpyHostToDevice, c->stream));
    if ((rc = pinned_release(c, *sb, c->stream))) return rc;
  } else if ((rc = upload_small(c, c->d_nowMs.p, nowMs, (size_t)P * 4))) {
    return rc;
  }
  if ((rc = c->haveAnimation ? launch_pose<2>(c, P) : launch_pose<1>(c, P))) return rc;
  c->P = P; c->K = K; c->palettesSet = true;
  return RZ_OK;
}

int32_t rz_load_animation(rz_ctx* c, const uint32_t* keyOffsets, const float* keyTimesMs, const float* keyQuats, const float* restQuat) {
  if (!c) return fail(nullptr, RZ_ERR_INVALID_ARG, "rz_load_animation: null ctx");
  if (!c->haveSkeleton) return fail(c, RZ_ERR_STATE, "rz_load_animation before rz_load_skeleton");
  CU_TRY(c, cudaSetDevice(c->device));
  if (!keyOffsets) {                       // unload: rz_set_instance_clocks evaluates the tween table again
    c->haveAnimation = false;
    return RZ_OK;
  }
  const uint32_t B = c->B, n = keyOffsets[B];
  if (keyOffsets[0] != 0) return fail(c, RZ_ERR_INVALID_ARG, "rz_load_animation: keyOffsets[0] must be 0");
  if (n && (!keyTimesMs || !keyQuats)) return fail(c, RZ_ERR_INVALID_ARG, "rz_load_animation: null keys");
  for (uint32_t b = 0; b < B; ++b) {
    if (keyOffsets[b + 1] < keyOffsets[b]) return fail(c, RZ_ERR_INVALID_ARG, "rz_load_animation: offsets not monotone at bone %u", b);
    for (uint32_t k = keyOffsets[b] + 1; k < keyOffsets[b + 1]; ++k)
      if (keyTimesMs[k] < keyTimesMs[k - 1]) return fail(c, RZ_ERR_INVALID_ARG, "rz_load_animation: key times of bone %u not ascending", b);
  }
  // normalise like Model.rotateBones does (model.ts:248; zero-length -> identity, math.ts:96-100)
  std::vector<float> q((size_t)std::max<uint32_t>(n, 1) * 4, 0.f);
  for (uint32_t k = 0; k < n; ++k) {
    const double x = keyQuats[k * 4], y = keyQuats[k * 4 + 1], z = keyQuats[k * 4 + 2], w = keyQuats[k * 4 + 3];
    const double len = std::sqrt(x * x + y * y + z * z + w * w);
    if (len == 0) { q[k * 4 + 3] = 1.f; continue; }
    q[k * 4] = (float)(x / len); q[k * 4 + 1] = (float)(y / len); q[k * 4 + 2] = (float)(z / len); q[k * 4 + 3] = (float)(w / len);
  }
  std::vector<float> rest((size_t)B * 4, 0.f);
  for (uint32_t b = 0; b < B; ++b) {
    if (restQuat) memcpy(&rest[(size_t)b * 4], restQuat + (size_t)b * 4, 16);
    else rest[(size_t)b * 4 + 3] = 1.f;   // playAnimation resets bones without keys to identity (engine.ts:1489-1505)
  }
  int rc;
  if ((rc = upload(c, c->d_trStart, keyOffsets, (size_t)(B + 1) * 4))) return rc;
  if ((rc = upload(c, c->d_trMs, n ? keyTimesMs : rest.data(), (size_t)std::max<uint32_t>(n, 1) * 4))) return rc;
  if ((rc = upload(c, c->d_trQ, q.data(), q.size() * 4))) return rc;
  if ((rc = upload(c, c->d_trRest, rest.data(), (size_t)B * 16))) return rc;   // the clip's own rest rotations (d_twRest belongs to rz_set_tweens)
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  c->haveAnimation = true;
  return RZ_OK;
}

int32_t rz_set_morph_weights(rz_ctx* c, const float* w, const uint32_t* ids, uint32_t Mact, uint32_t K) {
  if (!c) return fail(nullptr, RZ_ERR_INVALID_ARG, "rz_set_morph_weights: null ctx");
  if (c->V == 0) return fail(c, RZ_ERR_STATE, "rz_set_morph_weights before rz_load_mesh");
  if (K == 0 || K > c->maxK) return fail(c, RZ_ERR_INVALID_ARG, "rz_set_morph_weights: K=%u out of range (max %u)", K, c->maxK);
  if (Mact && (!w || !ids)) return fail(c, RZ_ERR_INVALID_ARG, "rz_set_morph_weights: null weights/ids");
  for (uint32_t a = 0; a < Mact; ++a)
    if (ids[a] >= c->M) return fail(c, RZ_ERR_INVALID_ARG, "rz_set_morph_weights: morph id %u >= M=%u", ids[a], c->M);
  CU_TRY(c, cudaSetDevice(c->device));
  int rc;
  if ((rc = ensure_dense_weights(c))) return rc;
  if (Mact == 0) {
    CU_TRY(c, cudaMemsetAsync(c->d_mwDense.p, 0, (size_t)c->maxK * c->Mpad * 4, c->stream));
    c->Mact = 0;
    return RZ_OK;
  }
  const size_t wBytes = (size_t)K * Mact * 4, idBytes = (size_t)Mact * 4;
  if ((rc = dev_reserve(c, c->d_mwIn, wBytes))) return rc;
  if ((rc = dev_reserve(c, c->d_mwIds, idBytes))) return rc;
  if ((rc = upload_small(c, c->d_mwIn.p, w, wBytes, ids, c->d_mwIds.p, idBytes))) return rc;
  morph_weights_kernel<<<K, 128, 0, c->stream>>>(reinterpret_cast<const float*>(c->d_mwIn.p),
                                                 reinterpret_cast<const uint32_t*>(c->d_mwIds.p),
                                                 reinterpret_cast<float*>(c->d_mwDense.p), K, Mact, c->Mpad);
  CU_TRY(c, cudaGetLastError());
  c->launches++;
  c->Mact = Mact;
  return RZ_OK;
}

int32_t rz_deform(rz_ctx* c, uint32_t first, uint32_t count) {
  if (!c) return fail(nullptr, RZ_ERR_INVALID_ARG, "rz_deform: null ctx");
  if (c->V == 0) return fail(c, RZ_ERR_STATE, "rz_deform before rz_load_mesh");
  if (!c->palettesSet) return fail(c, RZ_ERR_STATE, "rz_deform before rz_set_palettes");
  if (count == 0 || first >= c->K || count > c->K - first)
    return fail(c, RZ_ERR_INVALID_ARG, "rz_deform: instance range [%u,+%u) outside K=%u", first, count, c->K);
  CU_TRY(c, cudaSetDevice(c->device));
  int rc;
  if (c->tablesDirty && (rc = rebuild_tables(c))) return rc;

  int need = 0;
  if (c->M && c->morphNnz && c->Mact) need |= FEAT_MORPH;
  if ((c->flags & RZ_FLAG_SDEF) && c->sdefActive) need |= FEAT_SDEF;
  if (c->flags & RZ_FLAG_BOUNDS) need |= FEAT_BOUNDS;
  if (c->flags & RZ_FLAG_NO_NORMALS) need |= FEAT_NONRM;
  if (c->flags & RZ_FLAG_OUTLINE) need |= FEAT_HULL;
  if (c->flags & RZ_FLAG_INTERLEAVED) need |= FEAT_ILV;
  if ((rc = ensure_dense_weights(c))) return rc;
  const uint32_t Mpad = c->Mpad;

  // ---- pick the launch shape
  const size_t smemMax = (size_t)c->maxSmemOptin;
  if (smem_needed(1, 256, need, c->B, Mpad) > smemMax) need |= FEAT_GPAL;   // palette does not fit: gather from global
  const int feat = resolve_feat(need);
  if (feat < 0) return fail(c, RZ_ERR_INVALID_ARG, "rz_deform: feature combination 0x%x is not built", need);
  // a pipelined upload (rz_set_palettes) is consumed block by block below; feature sets with per-launch side kernels or
  // count-dependent tables take the whole upload first
  bool uploadInFlight = false;                                      // blocks of a pipelined rz_set_palettes upload
  for (const auto& b : c->pend) uploadInFlight |= b.ev != nullptr;
  const bool pipelined = uploadInFlight && !(feat & (FEAT_MORPH | FEAT_SDEF | FEAT_BOUNDS | FEAT_GPAL));
  if (uploadInFlight && !pipelined && (rc = flush_pending(c))) return rc;
  KernelEntry ke{nullptr, 0, 0, 0, 0, 0, feat};
  size_t smem = 0;
  int occ = 0;
  const uint32_t countClass = std::min<uint32_t>(count, 8u);        // shapes are only restricted by count when count < I <= 8
  // the plain planar path runs the two-vertices-per-lane kernel (deform2_kernel.cuh) whenever its table exists
  // -- in every output layout, as long as no morph / SDEF is active; the per-instance AABB with the planar layout only
  const bool v2Allowed = ((feat & ~(FEAT_NONRM | FEAT_HULL | FEAT_ILV)) == 0 || feat == FEAT_BOUNDS) && c->vpl2Ready && c->tuneVpl != 1;
  const int out2 = (feat & FEAT_ILV) ? OUT2_ILV : (feat & FEAT_HULL) ? OUT2_HULL : (feat & FEAT_NONRM) ? OUT2_NONRM
                   : (feat & FEAT_BOUNDS) ? OUT2_BOUNDS : OUT2_PLANAR;
  bool v2 = false;
  const bool cached = c->shapeKe.fn && c->shapeKey.feat == feat && c->shapeKey.B == c->B && c->shapeKey.Mpad == Mpad &&
                      c->shapeKey.countClass == countClass && c->shapeKey.v2Allowed == v2Allowed;
  if (cached) { ke = c->shapeKe; smem = c->shapeSmem; occ = c->shapeOcc; v2 = c->shapeV2; }
  const bool tuned = c->tuneI || c->tuneThreads;                    // an explicitly requested shape is taken as it is
  auto try_shape = [&](int I, int NT, int MINB) -> bool {
    if (!tuned && (uint32_t)I > count && I > 1) return false;       // never wider than the instance range
    KernelEntry e = lookup_kernel(feat, I, NT, MINB);
    if (!e.fn) return false;
    const size_t sm = smem_needed(e.I, e.NT, feat, c->B, Mpad, e.NB);
    if (sm > smemMax) return false;
    if (raise_smem_limit(c->device, e.fn, sm) != cudaSuccess) { cudaGetLastError(); return false; }
    int o = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, e.fn, e.NT, sm) != cudaSuccess || o < 1) { cudaGetLastError(); return false; }
    ke = e; smem = sm; occ = o; v2 = false;
    return true;
  };
  auto try_shape2 = [&](int I, int NT, int MINB, int SB) -> bool {
    if (!v2Allowed || (!tuned && (uint32_t)I > count && I > 1)) return false;
    KernelEntry e = lookup_v2(out2, I, NT, MINB, SB);
    if (!e.fn) return false;
    const size_t sm = smem_needed2(e.I, e.NT, c->B, e.SB, e.NB, out2);
    if (sm > smemMax) return false;
    if (raise_smem_limit(c->device, e.fn, sm) != cudaSuccess) { cudaGetLastError(); return false; }
    int o = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, e.fn, e.NT, sm) != cudaSuccess || o < 1) { cudaGetLastError(); return false; }
    ke = e; smem = sm; occ = o; v2 = true;
    return true;
  };
  if (cached) {
    // nothing to do
  } else if (tuned) {
    const int I = c->tuneI ? (int)c->tuneI : 2, NT = c->tuneThreads ? (int)c->tuneThreads : 256;
    // (tune_store_mode doubles as the sub-batch size of the two-vertex kernel; 0 = first compiled)
    if (try_shape2(I, NT, (int)c->tuneCtas, (int)c->tuneStore) || try_shape2(I, NT, 0, (int)c->tuneStore)) {
      // the requested shape exists for the two-vertex kernel
    } else if (c->tuneVpl == 2) {
      return fail(c, RZ_ERR_INVALID_ARG, "rz_deform: two-vertices-per-lane shape I=%d threads=%d ctas/SM=%u sub-batch=%u is not built, does not fit (B=%u) or "
                  "the context's flags rule the plain path out", I, NT, c->tuneCtas, c->tuneStore, c->B);
    } else if (!try_shape(I, NT, (int)c->tuneCtas) && !try_shape(I, NT, 0))
      return fail(c, RZ_ERR_INVALID_ARG, "rz_deform: requested launch shape I=%d threads=%d ctas/SM=%u is not built or does not fit (B=%u)",
                  I, NT, c->tuneCtas, c->B);
  } else {
    // preference order (measured on B200, profiles/): most resident warps first, then wider instance groups
    static const int pref[][3] = {{6, 512, 1}, {4, 512, 1}, {3, 256, 2}, {2, 512, 2}, {2, 256, 3}, {3, 512, 1}, {2, 256, 2}, {2, 512, 1}, {4, 256, 1},
                                  {1, 256, 4}, {1, 256, 2}, {1, 512, 2}};
    // feature kernels are compiled for fewer shapes; preference per feature set, measured on B200
    // (profiles/r01_config3_split_*.jsonl, profiles/r01_fused_consumers_K2048.jsonl):
    //   morph / SDEF: the wide group wins despite a few spilled registers (more (vertex, instance) pairs per SDEF dense
    //     pass, morph rows fetched once per 4 instances);
    //   AABB: 6 registers per instance for the bounds, I=6 spills -> I=4; outline hull: 36 B of staging per thread, the
    //     double-buffered I=2 x 2 CTAs shape beats the wide single-buffered ones; interleaved: I=4 x 768; positions only: I=6
    static const int prefWide4[][3] = {{4, 512, 1}, {4, 768, 1}, {2, 256, 2}, {2, 512, 1}, {1, 256, 2}};
    static const int prefHull[][3] = {{2, 256, 2}, {4, 768, 1}, {2, 512, 1}, {4, 512, 1}, {1, 256, 2}};
    static const int prefIlv[][3] = {{4, 768, 1}, {4, 512, 1}, {2, 256, 2}, {2, 512, 1}, {1, 256, 2}};
    static const int prefWide6[][3] = {{6, 512, 1}, {4, 768, 1}, {4, 512, 1}, {2, 256, 2}, {2, 512, 1}};
    const int (*prefLite)[3] = prefWide4;
    if (feat & FEAT_HULL) prefLite = prefHull;
    else if (feat & FEAT_ILV) prefLite = prefIlv;
    else if (feat == FEAT_NONRM) prefLite = prefWide6;
    bool ok = false;
    if (v2Allowed) {
      // two-vertex kernel: 8 warps x 64 vertices with the widest palette stage first (measured on B200, profiles/r02_*)
      static const int pref2[][4] = {{4, 512, 1, 2}, {4, 384, 1, 2}, {4, 512, 1, 1}, {3, 512, 1, 1}, {2, 256, 2, 2}, {2, 512, 1, 1}, {3, 256, 1, 3},
                                     {2, 256, 1, 2}, {1, 256, 2, 1}};
      // AABB: 6 accumulator registers per instance and an unrolled sub-batch loop: the narrower group wins (0.917 ms at
      // K = 2048 against 0.954 for 4 x 512 and 0.951 for the one-vertex kernel; profiles/r02_fused_consumers_K2048.jsonl)
      if (out2 == OUT2_BOUNDS && try_shape2(3, 512, 1, 1)) ok = true;
      // outline hull (36 B of staging per vertex-instance): 3 x 512 x SB 1 runs 1.179 ms at K = 2048 against 1.255 ms for 4 x 384;
      // interleaved: 6 x 512 x SB 1 1.016 ms against 1.049 ms for 4 x 512 (same run; consumer shape sweep of the final build)
      if (out2 == OUT2_HULL && try_shape2(3, 512, 1, 1)) ok = true;
      if (out2 == OUT2_ILV && (try_shape2(6, 512, 1, 1) || try_shape2(4, 384, 1, 2))) ok = true;
      if (!ok)
        for (const auto& p : pref2) if (try_shape2(p[0], p[1], p[2], p[3])) { ok = true; break; }
    }
    if (!ok && feat != 0) {
      for (int q = 0; q < 5 && !ok; ++q) ok = try_shape(prefLite[q][0], prefLite[q][1], prefLite[q][2]);
    }
    if (!ok) {
      for (const auto& p : pref) {
        if (try_shape(p[0], p[1], p[2])) {
          if (occ >= p[2]) { ok = true; break; }                   // the shape only pays off at its intended occupancy
        }
      }
    }
    if (!ok) {
      for (const auto& p : pref) if (try_shape(p[0], p[1], p[2])) { ok = true; break; }
      if (!ok) for (const auto& p : pref) if (try_shape(p[0], p[1], 0)) { ok = true; break; }
    }
    if (!ok) return fail(c, RZ_ERR_INVALID_ARG, "rz_deform: no kernel shape fits B=%u (smem limit %zu)", c->B, smemMax);
  }
  if (c->tuneCtas && (int)c->tuneCtas < occ) occ = (int)c->tuneCtas;
  CU_TRY(c, raise_smem_limit(c->device, ke.fn, smem));      // (a no-op unless another device / a first use needs it)
  if (!cached) {
    c->shapeKe = ke; c->shapeSmem = smem; c->shapeOcc = occ; c->shapeV2 = v2;
    c->shapeKey.v2Allowed = v2Allowed;
    c->shapeKey.feat = feat; c->shapeKey.B = c->B; c->shapeKey.Mpad = Mpad; c->shapeKey.countClass = countClass;
  }
  DeformParams prm;
  memset(&prm, 0, sizeof prm);
  prm.rec0 = reinterpret_cast<const float4*>(c->d_rec0.p);
  prm.rec1 = reinterpret_cast<const float4*>(c->d_rec1.p);
  prm.rec2 = reinterpret_cast<const float4*>(c->d_rec2.p);
  prm.meta = reinterpret_cast<const uint32_t*>(c->d_meta.p);
  prm.mrange = reinterpret_cast<const uint2*>(c->d_mrange.p);
  prm.ments = reinterpret_cast<const float4*>(c->d_ments.p);
  prm.sdefIdx = reinterpret_cast<const uint32_t*>(c->d_sdefIdx.p);
  prm.sdefTab = reinterpret_cast<const float4*>(c->d_sdefTab.p);
  prm.skin = reinterpret_cast<const float*>(c->d_skin.p);
  prm.quat = reinterpret_cast<const float4*>(c->d_quat.p);
  prm.inst2pal = c->haveInst2pal ? reinterpret_cast<const uint32_t*>(c->d_inst2pal.p) : nullptr;
  prm.mweights = reinterpret_cast<const float*>(c->d_mwDense.p);
  prm.edge = reinterpret_cast<const float*>(c->d_edge.p);
  prm.uv = reinterpret_cast<const float2*>(c->d_uv.p);
  prm.hullOffF = c->hullOffF;
  prm.out = out_base(c);
  prm.bounds = reinterpret_cast<float*>(c->d_bounds.p);
  prm.instStrideF = c->instStrideF;
  prm.nrmOffF = c->nrmOffF;
  prm.V = c->V; prm.B = c->B; prm.nTiles = c->nTiles; prm.Mpad = Mpad;
  prm.counter = reinterpret_cast<uint32_t*>(c->d_counter.p);
  prm.packedMeta = c->packedMeta ? 1u : 0u;
  prm.posStride = c->layoutMode ? 16u : 48u;
  prm.rowStride = c->layoutMode ? c->B * 16u : 16u;
  Deform2Params prm2;
  memset(&prm2, 0, sizeof prm2);
  if (v2) {
    prm2.rec = reinterpret_cast<const float4*>(c->d_rec2v.p);
    prm2.skin = prm.skin; prm2.inst2pal = prm.inst2pal; prm2.out = prm.out; prm2.bounds = prm.bounds;
    prm2.instStrideF = c->instStrideF; prm2.nrmOffF = c->nrmOffF; prm2.hullOffF = c->hullOffF;
    prm2.lanes = c->vgCount * 32; prm2.nVG = c->vgCount; prm2.V = c->V; prm2.B = c->B;
    prm2.counter = prm.counter;
  }
  uint32_t gridUsed = 0;
  // ---- one launch over the instance range [f, f+n): plan (host work, may synchronise once per launch shape) ...
  struct RangeLaunch { DeformParams prm; Deform2Params prm2; uint32_t grid; };
  auto plan_range = [&](uint32_t f, uint32_t n, RangeLaunch& out) -> int {
    if (v2) {
      prm2.K0 = f; prm2.Kcount = n;
      prm2.nGroups = (n + ke.I - 1) / ke.I;
      const uint32_t vgPerPass = ke.NT / 32;                        // one vertex group (64 vertices) per warp per pass
      const uint32_t nPasses = (c->vgCount + vgPerPass - 1) / vgPerPass;
      uint32_t grid = (uint32_t)(c->numSM * occ);
      ItemPlan ip = c->tuneChunks ? ItemPlan{0, c->tuneChunks} : pick_items(prm2.nGroups, grid, c->V, ke.I, nPasses);
      const uint32_t nChunks = std::max(1u, std::min(ip.chunks, nPasses));
      prm2.vgPerChunk = (nPasses + nChunks - 1) / nChunks * vgPerPass;
      prm2.nChunks = (c->vgCount + prm2.vgPerChunk - 1) / prm2.vgPerChunk;
      prm2.nCoarse = std::min(ip.coarse, prm2.nGroups);
      prm2.nItems = prm2.nCoarse + (prm2.nGroups - prm2.nCoarse) * prm2.nChunks;
      grid = std::min(grid, prm2.nItems);
      gridUsed = std::max(gridUsed, grid);
      memset(&out.prm, 0, sizeof out.prm);
      out.prm2 = prm2;
      out.grid = grid;
      return RZ_OK;
    }
    prm.K0 = f; prm.Kcount = n;
    prm.nGroups = (n + ke.I - 1) / ke.I;
    const uint32_t tilesPerPass = ke.NT / kTile;
    const uint32_t nPasses = (c->nTiles + tilesPerPass - 1) / tilesPerPass;
    uint32_t grid = (uint32_t)(c->numSM * occ);
    const bool morphTab = (feat & FEAT_MORPH) && c->morphNnz;       // cost-balanced chunk table below: uniform, finer items
    ItemPlan ip = c->tuneChunks ? ItemPlan{0, c->tuneChunks}
                : morphTab ? ItemPlan{0, (grid * 8 + prm.nGroups - 1) / prm.nGroups}
                           : pick_items(prm.nGroups, grid, c->V, ke.I, nPasses, false);
    uint32_t nChunks = std::max(1u, std::min(ip.chunks, nPasses));
    const uint32_t passesPerChunk = (nPasses + nChunks - 1) / nChunks;
    prm.tilesPerChunk = passesPerChunk * tilesPerPass;
    prm.nChunks = (c->nTiles + prm.tilesPerChunk - 1) / prm.tilesPerChunk;
    prm.chunkTab = nullptr;
    if ((feat & FEAT_MORPH) && c->morphNnz && nChunks > 1) {
      // Morph passes are latency chains (record -> entries -> weights: one L2 round trip per kMorphBatch entries of the
      // deepest vertex), several times longer than a plain pass, and PMX morphs cluster on the face: uniform chunks would
      // leave a few very long items (measured: +70 % on config 3).  Chunk boundaries are placed so that every item carries
      // the same estimated time instead; the table only depends on the launch shape and is cached.
      if (c->chunkKey.tilesPerPass != tilesPerPass || c->chunkKey.target != nChunks || !c->d_chunkTab.p) {
        static const float tripCost = getenv("RZ_MORPH_TRIP_COST") ? (float)atof(getenv("RZ_MORPH_TRIP_COST")) : 0.5f;
        std::vector<uint32_t> tab;
        build_chunk_table(c->tileMorphMax.data(), c->nTiles, tilesPerPass, nChunks, tripCost, kMorphPF, kMorphBatch, tab);
        int rcc;
        if ((rcc = dev_reserve(c, c->d_chunkTab, tab.size() * 4))) return rcc;
        CU_TRY(c, cudaMemcpyAsync(c->d_chunkTab.p, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice, c->stream));
        CU_TRY(c, cudaStreamSynchronize(c->stream));                  // once per launch shape; `tab` goes out of scope
        c->chunkKey.tilesPerPass = tilesPerPass; c->chunkKey.target = nChunks;
        c->chunkCount = (uint32_t)tab.size() - 1;
      }
      prm.chunkTab = reinterpret_cast<const uint32_t*>(c->d_chunkTab.p);
      prm.nChunks = c->chunkCount;
    }
    prm.nCoarse = std::min(ip.coarse, prm.nGroups);
    prm.nItems = prm.nCoarse + (prm.nGroups - prm.nCoarse) * prm.nChunks;
    grid = std::min(grid, prm.nItems);
    gridUsed = std::max(gridUsed, grid);
    out.prm = prm;
    memset(&out.prm2, 0, sizeof out.prm2);
    out.grid = grid;
    return RZ_OK;
  };
  // ... and issue (stream operations only: also runs under stream capture)
  auto issue_range = [&](RangeLaunch& r) -> int {
    CU_TRY(c, cudaMemsetAsync(c->d_counter.p, 0, 4, c->stream));
    void* args[] = {v2 ? (void*)&r.prm2 : (void*)&r.prm};
    CU_TRY(c, cudaLaunchKernel(ke.fn, dim3(r.grid), dim3(ke.NT), args, smem, c->stream));
    c->launches++;
    return RZ_OK;
  };
  auto launch_range = [&](uint32_t f, uint32_t n) -> int {
    RangeLaunch r;
    int rcl;
    if ((rcl = plan_range(f, n, r))) return rcl;
    return issue_range(r);
  };

  if (feat & FEAT_SDEF) {
    if ((rc = dev_reserve(c, c->d_quat, (size_t)c->P * c->B * 16))) return rc;
    prm.quat = reinterpret_cast<const float4*>(c->d_quat.p);
  }
  if ((feat & FEAT_BOUNDS) && !c->d_bounds.p) {
    // the smallest compiled superset of the requested features may carry the AABB reduction although RZ_FLAG_BOUNDS
    // is off: give it somewhere to write (rz_read_bounds still requires the flag)
    if ((rc = dev_reserve(c, c->d_bounds, (size_t)c->maxK * 24))) return rc;
    prm.bounds = reinterpret_cast<float*>(c->d_bounds.p);
  }
  // an asynchronous read-back still copying from the buffer this frame writes (with two buffers: the read of two frames ago)
  if (c->readPending[c->cur]) CU_TRY(c, cudaStreamWaitEvent(c->stream, c->evReadDone[c->cur], 0));
  // the side kernels of a frame, ahead of the deform launch (stream operations only)
  auto issue_side = [&]() {
    if (feat & FEAT_SDEF) {
      // rotation of every skin matrix as a quaternion, once per (palette, bone) instead of once per SDEF vertex-instance
      const uint32_t n = c->P * c->B;
      skin_quats_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(reinterpret_cast<const float4*>(c->d_skin.p),
                                                                 reinterpret_cast<float4*>(c->d_quat.p), c->P, c->B, c->layoutMode ? 1u : 0u);
      c->launches++;
    }
    if (feat & FEAT_BOUNDS) {
      bounds_reset_kernel<<<(count * 6 + 127) / 128, 128, 0, c->stream>>>(reinterpret_cast<int*>(c->d_bounds.p) + (size_t)first * 6, count * 6);
      c->launches++;
    }
  };
  if (pipelined) {
    CU_TRY(c, cudaEventRecord(c->evStart, c->stream));
    issue_side();
    // identity mapping: instance k reads palette k.  Blocks that do not touch [first, first+count) stay pending; parts of
    // the range whose blocks were consumed by an earlier call are launched as they are.
    std::vector<rz_ctx_impl::PendBlock> keep;
    uint32_t cursor = first;
    const uint32_t end = first + count;
    for (const auto& b : c->pend) {                                 // ascending pal0
      const uint32_t lo = std::max(first, b.pal0), hi = std::min(end, b.pal0 + b.n);
      if (lo >= hi) { keep.push_back(b); continue; }
      if (lo > cursor && (rc = launch_range(cursor, lo - cursor))) return rc;
      if (b.ev) CU_TRY(c, cudaStreamWaitEvent(c->stream, b.ev, 0));
      launch_skin_block(c, b.src, b.pal0, b.n);
      CU_TRY(c, cudaEventRecord(c->evWorldFree, c->stream));
      c->worldFreeValid = true;
      if ((rc = launch_range(lo, hi - lo))) return rc;
      cursor = hi;
    }
    if (cursor < end && (rc = launch_range(cursor, end - cursor))) return rc;
    c->pend.swap(keep);
  } else {
    // ---- the frame as ONE CUDA graph: [lazy skin-matrix passes] [skin quaternions] [AABB reset] counter reset, deform.
    // Recorded once per distinct frame description (every pointer, size and launch dimension below is part of the key)
    // and replayed with a single cudaGraphLaunch afterwards: a steady-state frame costs one driver call, whatever the
    // feature set ("launches pre-recorded in a CUDA graph").
    RangeLaunch r;
    if ((rc = plan_range(first, count, r))) return rc;
    rz_ctx_impl::GraphKey key;
    memset(&key, 0, sizeof key);
    key.fn = ke.fn; key.grid = r.grid; key.nt = (uint32_t)ke.NT; key.smem = smem; key.prm = r.prm; key.prm2 = r.prm2; key.feat = feat;
    key.first = first; key.count = count; key.P = c->P;
    key.invBind = c->d_invBind.p; key.bonePos = c->d_bonePos.p; key.layoutMode = c->layoutMode;
    key.nPend = (uint32_t)std::min<size_t>(c->pend.size(), 4);
    bool readsWorld = false;
    for (uint32_t i = 0; i < key.nPend; ++i) { key.pendSrc[i] = c->pend[i].src; key.pendPal0[i] = c->pend[i].pal0; key.pendN[i] = c->pend[i].n; }
    for (const auto& b : c->pend) readsWorld |= b.src == c->d_world.p;
    const bool useGraph = c->useGraphs && c->pend.size() <= 4;
    auto issue_frame = [&]() -> int {
      for (const auto& b : c->pend) launch_skin_block(c, b.src, b.pal0, b.n);
      issue_side();
      return issue_range(r);
    };
    CU_TRY(c, cudaEventRecord(c->evStart, c->stream));
    if (!useGraph) {
      if ((rc = issue_frame())) return rc;
    } else {
      rz_ctx_impl::GraphEntry* hit = nullptr;
      for (auto& g : c->graphs)
        if (memcmp(&g.key, &key, sizeof key) == 0) { hit = &g; break; }
      if (!hit) {
        const uint64_t launches0 = c->launches;
        cudaGraph_t graph = nullptr;
        CU_TRY(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
        const int rci = issue_frame();
        const cudaError_t ee = cudaStreamEndCapture(c->stream, &graph);
        if (rci) { if (graph) cudaGraphDestroy(graph); return rci; }
        if (ee != cudaSuccess) return fail(c, RZ_ERR_CUDA, "rz_deform: stream capture failed: %s", cudaGetErrorString(ee));
        rz_ctx_impl::GraphEntry ge;
        ge.key = key;
        ge.nodes = (uint32_t)(c->launches - launches0);
        const cudaError_t ei = cudaGraphInstantiate(&ge.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ei != cudaSuccess) return fail(c, RZ_ERR_CUDA, "rz_deform: cudaGraphInstantiate: %s", cudaGetErrorString(ei));
        c->launches = launches0;
        if (c->graphs.size() >= 8) {                                  // small cache, oldest out
          cudaGraphExecDestroy(c->graphs.front().exec);
          c->graphs.erase(c->graphs.begin());
        }
        c->graphs.push_back(ge);
        hit = &c->graphs.back();
      }
      CU_TRY(c, cudaGraphLaunch(hit->exec, c->stream));
      c->launches += hit->nodes;
      c->graphLaunches++;
    }
    c->pend.clear();
    if (readsWorld) {
      CU_TRY(c, cudaEventRecord(c->evWorldFree, c->stream));
      c->worldFreeValid = true;
    }
  }
  CU_TRY(c, cudaEventRecord(c->evStop, c->stream));
  c->evPending = true;
  c->frames++;
  c->usedI = ke.I; c->usedStore = v2 ? 3 : 2; c->usedVpl = v2 ? 2 : 1; c->usedCtas = gridUsed; c->usedThreads = ke.NT; c->usedSmem = (uint32_t)smem;
  c->lastVerts = (uint64_t)count * c->V;
  // compulsory DRAM bytes: outputs + mesh + palettes + invBind + morph entries/weights + sdef records
  // (fused consumers add their own compulsory bytes: +12 B per vertex-instance for the hull plane, 32 B instead of 24 B
  //  for the interleaved stream, +4 / +8 B per vertex for the edge sizes / texture coordinates read)
  const double planes = (feat & FEAT_ILV) ? 32.0 / 12.0 : ((feat & FEAT_NONRM) ? 1.0 : 2.0) + ((feat & FEAT_HULL) ? 1.0 : 0.0);
  c->lastAlgBytes = (double)count * c->V * 12.0 * planes + ((feat & FEAT_HULL) ? 4.0 * c->V : 0.0) + ((feat & FEAT_ILV) ? 8.0 * c->V : 0.0) + (double)c->V * 36.0 + (double)c->P * c->B * 64.0 + (double)c->B * 64.0 +
                    ((feat & FEAT_MORPH) ? (double)c->morphNnz * 16.0 + (double)count * c->Mact * 4.0 : 0.0) +
                    ((feat & FEAT_SDEF) ? (double)c->sdefActive * 36.0 : 0.0);
  const double now = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
  c->frameStamps.push_back(now);
  while (!c->frameStamps.empty() && now - c->frameStamps.front() > 1.0) c->frameStamps.erase(c->frameStamps.begin());
  return RZ_OK;
}

int32_t rz_sync(rz_ctx* c) {
  if (!c) return fail(nullptr, RZ_ERR_INVALID_ARG, "rz_sync: null ctx");
  CU_TRY(c, cudaSetDevice(c->device));
  int rc;
  if (c->V && c->palettesSet && (rc = flush_pending(c))) return rc;   // "everything issued so far" includes a lazy skin-matrix pass
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  return RZ_OK;
}

int32_t rz_output_device_ptr(rz_ctx* c, void** base, size_t* stride, size_t* nrmOff) {
  if (!c || !base) return fail(c, RZ_ERR_INVALID_ARG, "rz_output_device_ptr: null argument");
  if (c->V == 0) return fail(c, RZ_ERR_STATE, "rz_output_device_ptr before rz_load_mesh");
  *base = out_base(c);
  if (stride) *stride = c->instStrideF * 4;
  if (nrmOff) *nrmOff = (c->flags & RZ_FLAG_NO_NORMALS) ? 0 : c->nrmOffF * 4;   // (interleaved: 12, see rz_get_output_layout)
  return RZ_OK;
}

// one interleaved instance (RZ_FLAG_INTERLEAVED) to the host in CALLER vertex order, 8 floats per vertex
static int fetch_interleaved(rz_ctx* c, uint32_t inst, std::vector<float>& tmp) {
  tmp.resize((size_t)c->V * 8);
  const float* src = out_base(c) + (size_t)inst * c->instStrideF;
  CU_TRY(c, cudaMemcpyAsync(tmp.data(), src, (size_t)c->V * 32, cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  if (c->flags & RZ_FLAG_REORDER_VERTICES) {
    std::vector<float> t2((size_t)c->V * 8);
    for (uint32_t i = 0; i < c->V; ++i) memcpy(&t2[(size_t)c->vorder[i] * 8], &tmp[(size_t)i * 8], 32);
    tmp.swap(t2);
  }
  return RZ_OK;
}

int32_t rz_read_interleaved(rz_ctx* c, uint32_t inst, float* vtx8) {
  if (!c || !vtx8) return fail(c, RZ_ERR_INVALID_ARG, "rz_read_interleaved: null argument");
  if (c->V == 0) return fail(c, RZ_ERR_STATE, "rz_read_interleaved before rz_load_mesh");
  if (!(c->flags & RZ_FLAG_INTERLEAVED)) return fail(c, RZ_ERR_STATE, "rz_read_interleaved: context was created without RZ_FLAG_INTERLEAVED");
  if (inst >= c->maxK) return fail(c, RZ_ERR_INVALID_ARG, "rz_read_interleaved: instance %u >= max_instances %u", inst, c->maxK);
  CU_TRY(c, cudaSetDevice(c->device));
  std::vector<float> tmp;
  int rc;
  if ((rc = fetch_interleaved(c, inst, tmp))) return rc;
  memcpy(vtx8, tmp.data(), (size_t)c->V * 32);
  return RZ_OK;
}

int32_t rz_read_outline(rz_ctx* c, uint32_t inst, float* hull3) {
  if (!c || !hull3) return fail(c, RZ_ERR_INVALID_ARG, "rz_read_outline: null argument");
  if (c->V == 0) return fail(c, RZ_ERR_STATE, "rz_read_outline before rz_load_mesh");
  if (!(c->flags & RZ_FLAG_OUTLINE)) return fail(c, RZ_ERR_STATE, "rz_read_outline: context was created without RZ_FLAG_OUTLINE");
  if (inst >= c->maxK) return fail(c, RZ_ERR_INVALID_ARG, "rz_read_outline: instance %u >= max_instances %u", inst, c->maxK);
  CU_TRY(c, cudaSetDevice(c->device));
  const float* src = out_base(c) + (size_t)inst * c->instStrideF + c->hullOffF;
  if (c->flags & RZ_FLAG_REORDER_VERTICES) {
    std::vector<float> tmp((size_t)c->V * 3);
    CU_TRY(c, cudaMemcpyAsync(tmp.data(), src, (size_t)c->V * 12, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    for (uint32_t i = 0; i < c->V; ++i) memcpy(hull3 + (size_t)c->vorder[i] * 3, &tmp[(size_t)i * 3], 12);
    return RZ_OK;
  }
  CU_TRY(c, cudaMemcpyAsync(hull3, src, (size_t)c->V * 12, cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  return RZ_OK;
}

int32_t rz_get_output_layout(rz_ctx* c, rz_output_layout* out) {
  if (!c || !out) return fail(c, RZ_ERR_INVALID_ARG, "rz_get_output_layout: null argument");
  if (c->V == 0) return fail(c, RZ_ERR_STATE, "rz_get_output_layout before rz_load_mesh");
  const bool ilv = (c->flags & RZ_FLAG_INTERLEAVED) != 0;
  out->base = out_base(c);
  out->instanceStride = c->instStrideF * 4;
  out->vertexStride = ilv ? 32 : 12;
  out->positionOffset = 0;
  out->normalOffset = (c->flags & RZ_FLAG_NO_NORMALS) ? RZ_NO_ATTRIBUTE : c->nrmOffF * 4;
  out->hullOffset = (c->flags & RZ_FLAG_OUTLINE) ? c->hullOffF * 4 : RZ_NO_ATTRIBUTE;
  out->uvOffset = ilv ? 24 : RZ_NO_ATTRIBUTE;
  return RZ_OK;
}

int32_t rz_load_edge_size(rz_ctx* c, const float* edgeSize) {
  if (!c) return fail(nullptr, RZ_ERR_INVALID_ARG, "rz_load_edge_size: null ctx");
  if (c->V == 0) return fail(c, RZ_ERR_STATE, "rz_load_edge_size before rz_load_mesh");
  if (!(c->flags & RZ_FLAG_OUTLINE)) return fail(c, RZ_ERR_STATE, "rz_load_edge_size: context was created without RZ_FLAG_OUTLINE");
  CU_TRY(c, cudaSetDevice(c->device));
  if (edgeSize) c->h_edgeSize.assign(edgeSize, edgeSize + c->V); else c->h_edgeSize.clear();
  c->tablesDirty = true;
  return rebuild_tables(c);
}

int32_t rz_read_instance(rz_ctx* c, uint32_t inst, float* pos3, float* nrm3) {
  if (!c) return fail(nullptr, RZ_ERR_INVALID_ARG, "rz_read_instance: null ctx");
  if (c->V == 0) return fail(c, RZ_ERR_STATE, "rz_read_instance before rz_load_mesh");
  if (inst >= c->maxK) return fail(c, RZ_ERR_INVALID_ARG, "rz_read_instance: instance %u >= max_instances %u", inst, c->maxK);
  if (nrm3 && (c->flags & RZ_FLAG_NO_NORMALS)) return fail(c, RZ_ERR_STATE, "rz_read_instance: context was created with RZ_FLAG_NO_NORMALS");
  CU_TRY(c, cudaSetDevice(c->device));
  if (c->flags & RZ_FLAG_INTERLEAVED) {
    std::vector<float> tmp;
    int rc;
    if ((rc = fetch_interleaved(c, inst, tmp))) return rc;
    for (uint32_t i = 0; i < c->V; ++i) {
      if (pos3) memcpy(pos3 + (size_t)i * 3, &tmp[(size_t)i * 8], 12);
      if (nrm3) memcpy(nrm3 + (size_t)i * 3, &tmp[(size_t)i * 8 + 3], 12);
    }
    return RZ_OK;
  }
  const float* src = out_base(c) + (size_t)inst * c->instStrideF;
  if (c->flags & RZ_FLAG_REORDER_VERTICES) {
    // device planes are in the library's stored order: hand them back in the caller's vertex order
    std::vector<float> tmp((size_t)c->V * 3);
    for (int plane = 0; plane < 2; ++plane) {
      float* dst = plane ? nrm3 : pos3;
      if (!dst) continue;
      CU_TRY(c, cudaMemcpyAsync(tmp.data(), src + (plane ? c->nrmOffF : 0), (size_t)c->V * 12, cudaMemcpyDeviceToHost, c->stream));
      CU_TRY(c, cudaStreamSynchronize(c->stream));
      for (uint32_t i = 0; i < c->V; ++i) memcpy(dst + (size_t)c->vorder[i] * 3, &tmp[(size_t)i * 3], 12);
    }
    return RZ_OK;
  }
  if (pos3) CU_TRY(c, cudaMemcpyAsync(pos3, src, (size_t)c->V * 12, cudaMemcpyDeviceToHost, c->stream));
  if (nrm3) CU_TRY(c, cudaMemcpyAsync(nrm3, src + c->nrmOffF, (size_t)c->V * 12, cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  return RZ_OK;
}

int32_t rz_read_instance_async(rz_ctx* c, uint32_t inst, float* pos3, float* nrm3) {
  if (!c) return fail(nullptr, RZ_ERR_INVALID_ARG, "rz_read_instance_async: null ctx");
  if (c->V == 0) return fail(c, RZ_ERR_STATE, "rz_read_instance_async before rz_load_mesh");
  if (inst >= c->maxK) return fail(c, RZ_ERR_INVALID_ARG, "rz_read_instance_async: instance %u >= max_instances %u", inst, c->maxK);
  if (c->flags & (RZ_FLAG_REORDER_VERTICES | RZ_FLAG_INTERLEAVED))
    return fail(c, RZ_ERR_STATE, "rz_read_instance_async: needs the planar layout in caller order (use rz_read_instance / rz_read_interleaved)");
  if (nrm3 && (c->flags & RZ_FLAG_NO_NORMALS)) return fail(c, RZ_ERR_STATE, "rz_read_instance_async: context was created with RZ_FLAG_NO_NORMALS");
  CU_TRY(c, cudaSetDevice(c->device));
  const float* src = out_base(c) + (size_t)inst * c->instStrideF;
  CU_TRY(c, cudaEventRecord(c->evReadReady, c->stream));                 // everything issued so far (the frame's deform) ...
  CU_TRY(c, cudaStreamWaitEvent(c->readStream, c->evReadReady, 0));      // ... precedes the copies on the read stream
  if (pos3) CU_TRY(c, cudaMemcpyAsync(pos3, src, (size_t)c->V * 12, cudaMemcpyDeviceToHost, c->readStream));
  if (nrm3) CU_TRY(c, cudaMemcpyAsync(nrm3, src + c->nrmOffF, (size_t)c->V * 12, cudaMemcpyDeviceToHost, c->readStream));
  CU_TRY(c, cudaEventRecord(c->evReadDone[c->cur], c->readStream));
  c->readPending[c->cur] = true;
  return RZ_OK;
}

int32_t rz_read_wait(rz_ctx* c) {
  if (!c) return fail(nullptr, RZ_ERR_INVALID_ARG, "rz_read_wait: null ctx");
  CU_TRY(c, cudaSetDevice(c->device));
  for (int b = 0; b < 2; ++b) {
    if (c->readPending[b]) CU_TRY(c, cudaEventSynchronize(c->evReadDone[b]));
    c->readPending[b] = false;
  }
  return RZ_OK;
}

int32_t rz_get_vertex_order(rz_ctx* c, uint32_t* order) {
  if (!c || !order) return fail(c, RZ_ERR_INVALID_ARG, "rz_get_vertex_order: null argument");
  if (c->V == 0) return fail(c, RZ_ERR_STATE, "rz_get_vertex_order before rz_load_mesh");
  memcpy(order, c->vorder.data(), (size_t)c->V * 4);
  return RZ_OK;
}

int32_t rz_plan_lanes(const uint16_t* joints, const uint8_t* weights, uint32_t V, uint32_t B, uint32_t mode, uint32_t* laneVertex,
                      uint16_t* laneJoints, float* laneWeights, uint64_t* stats) {
  if (!joints || !weights || V == 0 || B == 0) return fail(nullptr, RZ_ERR_INVALID_ARG, "rz_plan_lanes: null or empty tables");
  for (size_t i = 0; i < (size_t)V * 4; ++i)
    if (joints[i] >= B) return fail(nullptr, RZ_ERR_INVALID_ARG, "rz_plan_lanes: joint %u >= B=%u", (unsigned)joints[i], B);
  LanePlan plan;
  plan_lanes(joints, weights, nullptr, V, B, kTile, (int)mode, plan);
  if (laneVertex) memcpy(laneVertex, plan.procVertex.data(), (size_t)plan.Vp * 4);
  if (laneJoints) memcpy(laneJoints, plan.gatherJ.data(), (size_t)plan.Vp * 8);
  if (laneWeights) memcpy(laneWeights, plan.devW.data(), (size_t)plan.Vp * 16);
  if (stats) {
    stats[0] = plan.fastSlots; stats[1] = plan.totalSlots;
    for (int n = 0; n < 5; ++n) for (int m = 0; m < 5; ++m) stats[2 + n * 5 + m] = plan.hist[n][m];
  }
  return RZ_OK;
}

int32_t rz_plan_morph_rows(const uint32_t* laneVertex, uint32_t Vp, uint32_t V, const uint32_t* morphOffsets, const uint32_t* vertIdx,
                           const float* delta3, uint32_t M, uint32_t* rowFirst, uint32_t* rowDepth, uint8_t* morphMajor, float* rows,
                           uint64_t rowsCapacity, uint64_t* rowsNeeded) {
  if (!laneVertex || Vp == 0 || Vp % 32 || V == 0) return fail(nullptr, RZ_ERR_INVALID_ARG, "rz_plan_morph_rows: bad lane table");
  if (M && (!morphOffsets || morphOffsets[0] != 0)) return fail(nullptr, RZ_ERR_INVALID_ARG, "rz_plan_morph_rows: morphOffsets[0] must be 0");
  const uint32_t nnz = M ? morphOffsets[M] : 0;
  for (uint32_t e = 0; e < nnz; ++e)
    if (vertIdx[e] >= V) return fail(nullptr, RZ_ERR_INVALID_ARG, "rz_plan_morph_rows: vertex index %u >= V=%u", vertIdx[e], V);
  for (uint32_t p = 0; p < Vp; ++p)
    if (laneVertex[p] != ~0u && laneVertex[p] >= V) return fail(nullptr, RZ_ERR_INVALID_ARG, "rz_plan_morph_rows: lane vertex out of range");
  std::vector<uint32_t> mcount, mstart;
  std::vector<F4> ments;
  morphs_by_vertex(V, M, morphOffsets, vertIdx, delta3, nullptr, mcount, mstart, ments);
  MorphRows mr;
  build_morph_rows(laneVertex, Vp, mcount, mstart, ments, mr);
  if (rowFirst) memcpy(rowFirst, mr.first.data(), (size_t)(Vp / 32) * 4);
  if (rowDepth) memcpy(rowDepth, mr.depth.data(), (size_t)(Vp / 32) * 4);
  if (morphMajor) memcpy(morphMajor, mr.morphMajor.data(), (size_t)(Vp / 32));
  if (rowsNeeded) *rowsNeeded = mr.rows.size();
  if (rows) {
    if (rowsCapacity < mr.rows.size()) return fail(nullptr, RZ_ERR_INVALID_ARG, "rz_plan_morph_rows: rows buffer too small");
    memcpy(rows, mr.rows.data(), mr.rows.size() * 16);
  }
  return RZ_OK;
}

int32_t rz_plan_lanes2(const uint16_t* joints, const uint8_t* weights, uint32_t V, uint32_t B, uint32_t groupCapacity, uint32_t* groupFirst,
                       uint32_t* groupCount, uint8_t* groupPaired, uint32_t* laneVertA, uint32_t* laneVertB, uint16_t* laneJoints,
                       float* laneWeightsA, float* laneWeightsB, uint64_t* stats, uint32_t* nGroups, uint8_t* laneSlotA, uint8_t* laneSlotB,
                       uint8_t* laneSlots) {
  if (!joints || !weights || V == 0 || B == 0 || !nGroups) return fail(nullptr, RZ_ERR_INVALID_ARG, "rz_plan_lanes2: null or empty tables");
  for (size_t i = 0; i < (size_t)V * 4; ++i)
    if (joints[i] >= B) return fail(nullptr, RZ_ERR_INVALID_ARG, "rz_plan_lanes2: joint %u >= B=%u", (unsigned)joints[i], B);
  LanePlan2 plan;
  plan_lanes2(joints, weights, V, B, plan);
  *nGroups = plan.nGroups;
  if (stats) { stats[0] = plan.fastSlots; stats[1] = plan.slots; stats[2] = plan.pairedWindows; stats[3] = plan.fallbackWindows; }
  if (!groupFirst && !laneVertA) return RZ_OK;              // size query
  if (groupCapacity < plan.nGroups) return fail(nullptr, RZ_ERR_INVALID_ARG, "rz_plan_lanes2: room for %u groups, need %u", groupCapacity, plan.nGroups);
  const size_t L = (size_t)plan.nGroups * 32;
  if (groupFirst) memcpy(groupFirst, plan.groupFirst.data(), (size_t)plan.nGroups * 4);
  if (groupCount) memcpy(groupCount, plan.groupCount.data(), (size_t)plan.nGroups * 4);
  if (groupPaired) memcpy(groupPaired, plan.groupPaired.data(), plan.nGroups);
  if (laneVertA) memcpy(laneVertA, plan.vertA.data(), L * 4);
  if (laneVertB) memcpy(laneVertB, plan.vertB.data(), L * 4);
  if (laneJoints) memcpy(laneJoints, plan.gatherJ.data(), L * 8);
  if (laneWeightsA) memcpy(laneWeightsA, plan.wA.data(), L * 16);
  if (laneWeightsB) memcpy(laneWeightsB, plan.wB.data(), L * 16);
  if (laneSlotA) memcpy(laneSlotA, plan.slotA.data(), L);
  if (laneSlotB) memcpy(laneSlotB, plan.slotB.data(), L);
  if (laneSlots) memcpy(laneSlots, plan.laneN.data(), L);
  return RZ_OK;
}

int32_t rz_plan_palette_rows(const uint16_t* laneJoints, uint32_t Vp, uint32_t B, uint32_t* bonePos) {
  if (!laneJoints || Vp == 0 || Vp % 32 || B == 0 || !bonePos) return fail(nullptr, RZ_ERR_INVALID_ARG, "rz_plan_palette_rows: bad argument");
  for (size_t i = 0; i < (size_t)Vp * 4; ++i)
    if (laneJoints[i] >= B) return fail(nullptr, RZ_ERR_INVALID_ARG, "rz_plan_palette_rows: joint %u >= B=%u", (unsigned)laneJoints[i], B);
  std::vector<uint32_t> pos;
  plan_palette_rows(laneJoints, Vp, B, 1, 0, pos);
  memcpy(bonePos, pos.data(), (size_t)B * 4);
  return RZ_OK;
}

int32_t rz_plan_sdef(const uint32_t* laneVertex, uint32_t Vp, const uint16_t* joints, const uint8_t* weights, uint32_t V,
                     uint32_t B, const uint32_t* sdefVertIdx, const float* c_r0_r1, uint32_t n, float* records, uint32_t* desc, uint32_t* nActive) {
  if (!laneVertex || Vp == 0 || Vp % 32 || !joints || !weights || V == 0 || B == 0)
    return fail(nullptr, RZ_ERR_INVALID_ARG, "rz_plan_sdef: bad lane / skinning tables");
  if (n && (!sdefVertIdx || !c_r0_r1)) return fail(nullptr, RZ_ERR_INVALID_ARG, "rz_plan_sdef: null SDEF arrays");
  std::vector<int32_t> sdefOf(V, -1);
  for (uint32_t i = 0; i < n; ++i) {
    if (sdefVertIdx[i] >= V) return fail(nullptr, RZ_ERR_INVALID_ARG, "rz_plan_sdef: vertex index %u >= V=%u", sdefVertIdx[i], V);
    const uint8_t* w = &weights[(size_t)sdefVertIdx[i] * 4];
    if (w[2] != 0 || w[3] != 0) continue;                 // not a two-influence vertex: stays linear (as rz_load_sdef)
    sdefOf[sdefVertIdx[i]] = (int32_t)i;
  }
  std::vector<uint32_t> ident(B);
  for (uint32_t b = 0; b < B; ++b) ident[b] = b;         // records carry bone ids here (no palette permutation without a mesh)
  std::vector<uint32_t> laneSlotV(Vp, 0);               // a warp owns 32 consecutive vertices: output slot = vertex mod 32
  for (uint32_t p = 0; p < Vp; ++p) {
    if (laneVertex[p] != ~0u && laneVertex[p] >= V) return fail(nullptr, RZ_ERR_INVALID_ARG, "rz_plan_sdef: lane vertex out of range");
    laneSlotV[p] = laneVertex[p] == ~0u ? 0u : (laneVertex[p] & 31u);
  }
  const uint32_t* laneSlot = laneSlotV.data();
  SdefTables t;
  if (!build_sdef_tables(sdefOf.data(), c_r0_r1, joints, weights, ident.data(), laneVertex, laneSlot, V, Vp, t))
    return fail(nullptr, RZ_ERR_INVALID_ARG, "rz_plan_sdef: more than 2^24 SDEF vertices");
  if (records) memcpy(records, t.tab.data(), t.tab.size() * 16);
  if (desc) memcpy(desc, t.desc.data(), (size_t)Vp * 4);
  if (nActive) *nActive = t.active;
  return RZ_OK;
}

int32_t rz_plan_chunks(const uint32_t* tileDepth, uint32_t nTiles, uint32_t tilesPerPass, uint32_t nChunksTarget, uint32_t* tab,
                       uint32_t* nChunks) {
  if (!tileDepth || nTiles == 0 || tilesPerPass == 0 || !tab || !nChunks)
    return fail(nullptr, RZ_ERR_INVALID_ARG, "rz_plan_chunks: bad argument");
  std::vector<uint32_t> t;
  build_chunk_table(tileDepth, nTiles, tilesPerPass, nChunksTarget, 0.5f, kMorphPF, kMorphBatch, t);
  memcpy(tab, t.data(), t.size() * 4);
  *nChunks = (uint32_t)t.size() - 1;
  return RZ_OK;
}

int32_t rz_read_bounds(rz_ctx* c, uint32_t first, uint32_t count, float* minmax6) {
  if (!c || !minmax6) return fail(c, RZ_ERR_INVALID_ARG, "rz_read_bounds: null argument");
  if (!(c->flags & RZ_FLAG_BOUNDS) || !c->d_bounds.p) return fail(c, RZ_ERR_STATE, "rz_read_bounds: context was created without RZ_FLAG_BOUNDS");
  if (first >= c->maxK || count > c->maxK - first) return fail(c, RZ_ERR_INVALID_ARG, "rz_read_bounds: range outside max_instances");
  CU_TRY(c, cudaSetDevice(c->device));
  std::vector<int> tmp((size_t)count * 6);
  CU_TRY(c, cudaMemcpyAsync(tmp.data(), reinterpret_cast<const int*>(c->d_bounds.p) + (size_t)first * 6, tmp.size() * 4,
                            cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  for (size_t i = 0; i < tmp.size(); ++i) {
    int v = tmp[i];
    if (v < 0) v ^= 0x7FFFFFFF;
    memcpy(&minmax6[i], &v, 4);
  }
  return RZ_OK;
}

// One vertex' (bone, weight) terms as the device holds them -> the caller's four slots.  Term order on the device is the
// planner's; the caller's order is recovered by matching bone ids against the caller's joint list (the permutation the
// library applied), every device term being consumed exactly once.  Returns false when a device term matches no slot.
static bool terms_to_caller_slots(const uint16_t* callerJ, const uint16_t* bone, const float* w, int nTerms, uint16_t* outJ, uint8_t* outW) {
  bool used[8] = {false, false, false, false, false, false, false, false};
  int left = 0;
  for (int t = 0; t < nTerms; ++t) left += w[t] != 0.f;
  for (int k = 0; k < 4; ++k) {
    outJ[k] = callerJ[k];                                   // zero-weight slot: the device spends nothing on it
    outW[k] = 0;
  }
  // pass 1: caller slots whose bone matches an unused device term
  for (int k = 0; k < 4 && left; ++k)
    for (int t = 0; t < nTerms; ++t) {
      if (used[t] || w[t] == 0.f || bone[t] != callerJ[k]) continue;
      const long q = lrintf(w[t] * 255.0f);
      outJ[k] = bone[t];
      outW[k] = (uint8_t)std::max(0l, std::min(255l, q));
      used[t] = true;
      --left;
      break;
    }
  return left == 0;
}

int32_t rz_read_skinning(rz_ctx* c, uint16_t* joints, uint8_t* weights) {
  if (!c) return fail(nullptr, RZ_ERR_INVALID_ARG, "rz_read_skinning: null ctx");
  if (c->V == 0) return fail(c, RZ_ERR_STATE, "rz_read_skinning before rz_load_mesh");
  CU_TRY(c, cudaSetDevice(c->device));
  // Everything below is derived from the records the KERNELS read: the pre-normalised f32 weights are re-quantised to UNORM8
  // and the palette rows mapped back to bone ids; the only host-side knowledge used is the permutation the library itself
  // applied (lane, influence slot, palette row).  Both device tables are decoded -- the one-vertex-per-lane records every
  // feature kernel runs on and, when present, the two-vertices-per-lane records of the plain path -- and must agree.
  const uint32_t V = c->V;
  std::vector<uint16_t> J1((size_t)V * 4), J2;
  std::vector<uint8_t> W1((size_t)V * 4), W2;
  {
    std::vector<float4> rec0(c->Vp), rec1(c->Vp), rec2(c->Vp);
    CU_TRY(c, cudaMemcpyAsync(rec0.data(), c->d_rec0.p, (size_t)c->Vp * 16, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaMemcpyAsync(rec1.data(), c->d_rec1.p, (size_t)c->Vp * 16, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaMemcpyAsync(rec2.data(), c->d_rec2.p, (size_t)c->Vp * 16, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    for (uint32_t p = 0; p < c->Vp; ++p) {
      const uint32_t sv = c->procToVertex[p];
      if (sv == ~0u) continue;
      const uint32_t v = c->vorder[sv];                     // stored position -> caller vertex id
      uint32_t j01, j23;
      memcpy(&j01, &rec2[p].z, 4);
      memcpy(&j23, &rec2[p].w, 4);
      if (c->packedMeta) {
        j01 = (j01 & 0xFFFu) | (((j01 >> 12) & 0xFFFu) << 16);
        j23 = (j23 & 0xFFFu) | (((j23 >> 12) & 0xFFFu) << 16);
      }
      const uint16_t dj[4] = {(uint16_t)c->boneAt[j01 & 0xFFFF], (uint16_t)c->boneAt[j01 >> 16], (uint16_t)c->boneAt[j23 & 0xFFFF],
                              (uint16_t)c->boneAt[j23 >> 16]};
      const float dw[4] = {rec0[p].w, rec1[p].w, rec2[p].x, rec2[p].y};
      for (uint32_t k = 0; k < 4; ++k) {
        // device slot that holds the caller's influence k (kNoSlot: its weight is zero, no slot was spent on it)
        const uint8_t sl = c->procSlotMap[(size_t)p * 4 + k];
        const long q = sl == kNoSlot ? 0 : lrintf(dw[sl] * 255.0f);
        W1[(size_t)v * 4 + k] = (uint8_t)std::max(0l, std::min(255l, q));
        // a zero-weight slot gathers a BORROWED palette row on the device (lane_plan.h: the row of a neighbouring lane, so
        // that the unconditional gather costs no extra shared-memory wavefront); it never influences the result and the
        // caller's own index is reported there
        J1[(size_t)v * 4 + k] = (sl != kNoSlot && q != 0) ? dj[sl] : c->h_joints[(size_t)v * 4 + k];
      }
    }
  }
  if (c->vpl2Ready) {
    const size_t L = (size_t)c->vgCount * 32;
    std::vector<float4> rec((size_t)kRec2Planes * L);
    CU_TRY(c, cudaMemcpyAsync(rec.data(), c->d_rec2v.p, rec.size() * 16, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    J2.assign((size_t)V * 4, 0);
    W2.assign((size_t)V * 4, 0);
    std::vector<uint8_t> seen(V, 0);
    for (size_t p = 0; p < L; ++p) {
      uint32_t r01, r23, meta;
      memcpy(&r01, &rec[5 * L + p].x, 4); memcpy(&r23, &rec[5 * L + p].y, 4); memcpy(&meta, &rec[5 * L + p].z, 4);
      const uint16_t bone[4] = {(uint16_t)c->boneAt[r01 & 0xFFFF], (uint16_t)c->boneAt[r01 >> 16], (uint16_t)c->boneAt[r23 & 0xFFFF],
                                (uint16_t)c->boneAt[r23 >> 16]};
      for (int side = 0; side < 2; ++side) {
        const uint32_t sv = side ? c->p2VertB[p] : c->p2VertA[p];
        if (((meta & (side ? kM2HasB : kM2HasA)) != 0) != (sv != ~0u))
          return fail(c, RZ_ERR_STATE, "rz_read_skinning: two-vertex record %zu disagrees with the lane plan about side %d", p, side);
        if (sv == ~0u) continue;
        const uint32_t v = c->vorder[sv];
        const float w[4] = {side ? rec[2 * L + p].w : rec[p].w, side ? rec[3 * L + p].w : rec[L + p].w,
                            side ? rec[4 * L + p].z : rec[4 * L + p].x, side ? rec[4 * L + p].w : rec[4 * L + p].y};
        if (!terms_to_caller_slots(&c->h_joints[(size_t)v * 4], bone, w, 4, &J2[(size_t)v * 4], &W2[(size_t)v * 4]))
          return fail(c, RZ_ERR_STATE, "rz_read_skinning: vertex %u carries a device term for a bone it does not list", v);
        seen[v]++;
      }
    }
    for (uint32_t v = 0; v < V; ++v)
      if (seen[v] != 1) return fail(c, RZ_ERR_STATE, "rz_read_skinning: vertex %u is evaluated %u times by the two-vertex table", v, (unsigned)seen[v]);
    if (J1 != J2 || W1 != W2) {
      for (uint32_t v = 0; v < V; ++v)
        if (memcmp(&J1[(size_t)v * 4], &J2[(size_t)v * 4], 8) || memcmp(&W1[(size_t)v * 4], &W2[(size_t)v * 4], 4))
          return fail(c, RZ_ERR_STATE, "rz_read_skinning: the one-vertex and two-vertex device tables disagree at vertex %u", v);
    }
  }
  if (joints) memcpy(joints, J1.data(), (size_t)V * 8);
  if (weights) memcpy(weights, W1.data(), (size_t)V * 4);
  return RZ_OK;
}

int32_t rz_read_skin_matrices(rz_ctx* c, uint32_t palette, float* skin3x4) {
  if (!c || !skin3x4) return fail(c, RZ_ERR_INVALID_ARG, "rz_read_skin_matrices: null argument");
  if (!c->palettesSet) return fail(c, RZ_ERR_STATE, "rz_read_skin_matrices before rz_set_palettes");
  if (palette >= c->P) return fail(c, RZ_ERR_INVALID_ARG, "rz_read_skin_matrices: palette %u >= P=%u", palette, c->P);
  CU_TRY(c, cudaSetDevice(c->device));
  {
    int rcf;
    if ((rcf = flush_pending(c))) return rcf;                 // a pipelined upload computes its skin matrices lazily
  }
  std::vector<float> tmp((size_t)c->B * 12);
  CU_TRY(c, cudaMemcpyAsync(tmp.data(), reinterpret_cast<const float*>(c->d_skin.p) + (size_t)palette * c->B * 12, (size_t)c->B * 48,
                            cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  for (uint32_t b = 0; b < c->B; ++b) {   // un-permute and un-pair (deform_kernel.cuh kRowF4) back to 3x4 row-major
    float t[12];
    if (c->layoutMode) { for (int r = 0; r < 3; ++r) memcpy(t + r * 4, tmp.data() + ((size_t)r * c->B + c->bonePos[b]) * 4, 16); }
    else memcpy(t, tmp.data() + (size_t)c->bonePos[b] * 12, 48);
    float* o = skin3x4 + (size_t)b * 12;
    o[0] = t[0]; o[4] = t[1]; o[1] = t[2]; o[5] = t[3];
    o[2] = t[4]; o[6] = t[5]; o[3] = t[6]; o[7] = t[7];
    o[8] = t[8]; o[9] = t[9]; o[10] = t[10]; o[11] = t[11];
  }
  return RZ_OK;
}

int32_t rz_read_world_matrices(rz_ctx* c, uint32_t palette, float* world16) {
  if (!c || !world16) return fail(c, RZ_ERR_INVALID_ARG, "rz_read_world_matrices: null argument");
  if (!c->palettesSet) return fail(c, RZ_ERR_STATE, "rz_read_world_matrices before this frame's palettes were set");
  if (palette >= c->P) return fail(c, RZ_ERR_INVALID_ARG, "rz_read_world_matrices: palette %u >= P=%u", palette, c->P);
  // The device keeps skin = world * invBind (rows 0-2); world = skin * invBind^-1.  invBind is affine (the reference's is a
  // pure translation, pmx-loader.ts:791-824), so its inverse is closed-form; evaluated in f64 on the host, B small products.
  std::vector<float> skin((size_t)c->B * 12);
  int rc = rz_read_skin_matrices(c, palette, skin.data());
  if (rc) return rc;
  for (uint32_t b = 0; b < c->B; ++b) {
    const float* ib = &c->h_invBind[(size_t)b * 16];            // column-major
    double A[3][3], t[3];
    for (int r = 0; r < 3; ++r) { for (int col = 0; col < 3; ++col) A[r][col] = ib[col * 4 + r]; t[r] = ib[12 + r]; }
    const double det = A[0][0] * (A[1][1] * A[2][2] - A[1][2] * A[2][1]) - A[0][1] * (A[1][0] * A[2][2] - A[1][2] * A[2][0]) +
                       A[0][2] * (A[1][0] * A[2][1] - A[1][1] * A[2][0]);
    if (det == 0.0) return fail(c, RZ_ERR_INVALID_ARG, "rz_read_world_matrices: inverse bind matrix of bone %u is singular", b);
    double Ai[3][3];
    Ai[0][0] = (A[1][1] * A[2][2] - A[1][2] * A[2][1]) / det; Ai[0][1] = (A[0][2] * A[2][1] - A[0][1] * A[2][2]) / det; Ai[0][2] = (A[0][1] * A[1][2] - A[0][2] * A[1][1]) / det;
    Ai[1][0] = (A[1][2] * A[2][0] - A[1][0] * A[2][2]) / det; Ai[1][1] = (A[0][0] * A[2][2] - A[0][2] * A[2][0]) / det; Ai[1][2] = (A[0][2] * A[1][0] - A[0][0] * A[1][2]) / det;
    Ai[2][0] = (A[1][0] * A[2][1] - A[1][1] * A[2][0]) / det; Ai[2][1] = (A[0][1] * A[2][0] - A[0][0] * A[2][1]) / det; Ai[2][2] = (A[0][0] * A[1][1] - A[0][1] * A[1][0]) / det;
    double ti[3];
    for (int r = 0; r < 3; ++r) ti[r] = -(Ai[r][0] * t[0] + Ai[r][1] * t[1] + Ai[r][2] * t[2]);
    const float* s = &skin[(size_t)b * 12];                     // 3x4 row-major
    float* o = world16 + (size_t)b * 16;
    for (int r = 0; r < 3; ++r) {
      for (int col = 0; col < 3; ++col) o[col * 4 + r] = (float)(s[r * 4] * Ai[0][col] + s[r * 4 + 1] * Ai[1][col] + s[r * 4 + 2] * Ai[2][col]);
      o[12 + r] = (float)(s[r * 4] * ti[0] + s[r * 4 + 1] * ti[1] + s[r * 4 + 2] * ti[2] + s[r * 4 + 3]);
    }
    o[3] = o[7] = o[11] = 0.f; o[15] = 1.f;
  }
  return RZ_OK;
}

int32_t rz_get_stats(rz_ctx* c, rz_stats* out) {
  if (!c || !out) return fail(c, RZ_ERR_INVALID_ARG, "rz_get_stats: null argument");
  CU_TRY(c, cudaSetDevice(c->device));
  if (c->evPending) {
    CU_TRY(c, cudaEventSynchronize(c->evStop));
    float ms = 0.f;
    CU_TRY(c, cudaEventElapsedTime(&ms, c->evStart, c->evStop));
    c->lastMs = ms;
    c->msRing.push_back(ms);
    if (c->msRing.size() > 60) c->msRing.erase(c->msRing.begin());   // 60-sample window like engine.ts:2423-2445
    c->evPending = false;
  }
  memset(out, 0, sizeof *out);
  double sum = 0;
  for (double m : c->msRing) sum += m;
  out->frameTime = c->msRing.empty() ? 0.0 : sum / (double)c->msRing.size();
  out->fps = (double)c->frameStamps.size();
  out->gpuMemory = (double)c->devBytes / (1024.0 * 1024.0);
  out->lastDeformMs = c->lastMs;
  out->algorithmicBytes = c->lastAlgBytes;
  if (c->lastMs > 0) {
    out->vertsPerSec = (double)c->lastVerts / (c->lastMs * 1e-3);
    out->achievedGBs = c->lastAlgBytes / (c->lastMs * 1e-3) / 1e9;
  }
  out->frames = c->frames;
  out->kernelLaunches = c->launches;
  out->vertexCount = c->V; out->boneCount = c->B; out->instanceCount = c->K; out->paletteCount = c->P;
  out->morphCount = c->M; out->morphNnz = c->morphNnz; out->sdefCount = c->sdefActive; out->activeMorphs = c->Mact;
  out->instancesPerGroup = c->usedI; out->storeMode = c->usedStore; out->ctas = c->usedCtas; out->threads = c->usedThreads;
  out->smemBytes = c->usedSmem;
  out->fastGatherPermille = c->packTotalSlots ? (uint32_t)(c->packFastSlots * 1000 / c->packTotalSlots) : 0u;
  if (c->usedVpl == 2 && c->p2Slots) out->fastGatherPermille = (uint32_t)(c->p2FastSlots * 1000 / c->p2Slots);
  out->verticesPerLane = c->usedVpl;
  return RZ_OK;
}

}  // extern "C"
