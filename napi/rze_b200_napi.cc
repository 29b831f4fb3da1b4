// rze_b200_napi.cc — thin N-API addon over the C ABI (include/rze_b200.h).
//
// Pure marshalling: every JS call maps 1:1 onto an rz_* entry point; typed arrays are passed through without
// copies (the C ABI copies what it keeps).  A non-zero status becomes a thrown JS Error carrying rz_last_error(),
// which is the reference's convention for fatal conditions (engine.ts:161,167,1828: `throw new Error(...)`).
//
// The C ABI takes plain pointers and trusts the sizes it is told; a JS caller can hand over a typed array of any
// length, so THIS layer checks every array length against the counts the library will read or write (V, B, M, P, K,
// from rz_get_stats and the call's own arguments) and throws a RangeError before the pointer crosses the boundary.
// Arrays handed to readInstanceAsync are kept alive (napi_ref) until readWait.
//
// NOT BUILT IN THIS IMAGE: no Node.js / node_api.h exists here or on the GPU box, so this file is
// compiled only where <node_api.h> is available:
//     g++ -O2 -fPIC -shared -I$(node -p "require('node-addon-api').include_dir") -Iinclude napi/rze_b200_napi.cc
//         -Lreze-engine_b200/lib -lrze_b200 -o rze_b200.node          (one command line)
// The tested boundary is the C ABI itself (tests/ drive it through ctypes with the same argument marshalling); the CPU
// suite compiles this file against a stand-in <node_api.h> with -Wall -Wextra -Werror (tests/test_host.py).
#if __has_include(<node_api.h>)
#include <node_api.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../include/rze_b200.h"

namespace {

#define NAPI_OK(env, call)                                         \
  do {                                                             \
    if ((call) != napi_ok) {                                       \
      napi_throw_error((env), nullptr, "N-API call failed: " #call); \
      return nullptr;                                              \
    }                                                              \
  } while (0)

// what the JS side holds: the context + the typed arrays an asynchronous read-back still writes into
struct Handle {
  rz_ctx* c = nullptr;
  std::vector<napi_ref> held;
};

Handle* unwrap(napi_env env, napi_value v) {
  void* p = nullptr;
  if (napi_get_value_external(env, v, &p) != napi_ok || !p || !static_cast<Handle*>(p)->c) {
    napi_throw_type_error(env, nullptr, "expected a live deform context");
    return nullptr;
  }
  return static_cast<Handle*>(p);
}

bool check(napi_env env, rz_ctx* c, int32_t st) {
  if (st == RZ_OK) return true;
  napi_throw_error(env, nullptr, rz_last_error(c));
  return false;
}

// `have` elements were passed where the call needs `need`: throws and returns false when too short
bool need_len(napi_env env, const char* what, size_t have, size_t need) {
  if (have >= need) return true;
  char msg[160];
  snprintf(msg, sizeof msg, "%s: typed array has %zu elements, the call needs %zu", what, have, need);
  napi_throw_range_error(env, nullptr, msg);
  return false;
}
bool exact_len(napi_env env, const char* what, size_t have, size_t need) {
  if (have == need) return true;
  char msg[160];
  snprintf(msg, sizeof msg, "%s: typed array has %zu elements, expected exactly %zu", what, have, need);
  napi_throw_range_error(env, nullptr, msg);
  return false;
}

template <typename T>
T* typed(napi_env env, napi_value v, size_t* n, napi_typedarray_type want) {
  napi_typedarray_type t;
  size_t len = 0, off = 0;
  void* data = nullptr;
  napi_value ab;
  if (napi_get_typedarray_info(env, v, &t, &len, &data, &ab, &off) != napi_ok || t != want) {
    napi_throw_type_error(env, nullptr, "wrong typed-array type");
    return nullptr;
  }
  if (n) *n = len;
  return static_cast<T*>(data);
}

bool is_nullish(napi_env env, napi_value v) {
  napi_valuetype vt = napi_undefined;
  napi_typeof(env, v, &vt);
  return vt == napi_null || vt == napi_undefined;
}

uint32_t u32(napi_env env, napi_value v) {
  uint32_t x = 0;
  napi_get_value_uint32(env, v, &x);
  return x;
}

// V, B, ... of the loaded mesh (0 before rz_load_mesh)
bool stats_of(napi_env env, rz_ctx* c, rz_stats* s) { return check(env, c, rz_get_stats(c, s)); }

void finalize(napi_env env, void* data, void*) {
  Handle* h = static_cast<Handle*>(data);
  if (h->c) rz_destroy(h->c);
  for (napi_ref r : h->held) napi_delete_reference(env, r);
  delete h;
}

#define ARGS(n)                                                                 \
  size_t argc = (n);                                                            \
  napi_value a[(n)];                                                            \
  NAPI_OK(env, napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));        \
  if (argc < (n)) { napi_throw_type_error(env, nullptr, "too few arguments"); return nullptr; }

// create(device, maxInstances, flags) -> external          Engine.init() (engine.ts:157-185)
napi_value Create(napi_env env, napi_callback_info info) {
  ARGS(3)
  rz_config cfg;
  std::memset(&cfg, 0, sizeof cfg);
  cfg.struct_size = sizeof cfg;
  cfg.device = (int32_t)u32(env, a[0]);
  cfg.max_instances = u32(env, a[1]);
  cfg.flags = u32(env, a[2]);
  rz_ctx* c = nullptr;
  if (!check(env, nullptr, rz_create(&cfg, &c))) return nullptr;
  Handle* h = new Handle();
  h->c = c;
  napi_value ext;
  if (napi_create_external(env, h, finalize, nullptr, &ext) != napi_ok) {
    rz_destroy(c);
    delete h;
    napi_throw_error(env, nullptr, "napi_create_external failed");
    return nullptr;
  }
  return ext;
}

// destroy(ctx): Engine.dispose() (engine.ts:1692-1701); the handle is dead afterwards (the GC finaliser then only frees it)
napi_value Destroy(napi_env env, napi_callback_info info) {
  ARGS(1)
  Handle* h = unwrap(env, a[0]);
  if (!h) return nullptr;
  rz_read_wait(h->c);
  for (napi_ref r : h->held) napi_delete_reference(env, r);
  h->held.clear();
  rz_ctx* c = h->c;
  h->c = nullptr;
  check(env, nullptr, rz_destroy(c));
  return nullptr;
}

// loadMesh(ctx, Float32Array vtx8, Uint16Array joints, Uint8Array weights, Float32Array invBind)
//   exactly Model.getVertices() / getSkinning() / getBoneInverseBindMatrices() of the reference (model.ts:196-200, 47-50, 321)
napi_value LoadMesh(napi_env env, napi_callback_info info) {
  ARGS(5)
  Handle* h = unwrap(env, a[0]);
  size_t nv = 0, nj = 0, nw = 0, nb = 0;
  const float* v = typed<float>(env, a[1], &nv, napi_float32_array);
  const uint16_t* j = typed<uint16_t>(env, a[2], &nj, napi_uint16_array);
  const uint8_t* w = typed<uint8_t>(env, a[3], &nw, napi_uint8_array);
  const float* ib = typed<float>(env, a[4], &nb, napi_float32_array);
  if (!h || !v || !j || !w || !ib) return nullptr;
  if (nv % 8 || nb % 16) { napi_throw_range_error(env, nullptr, "loadMesh: vertices must hold 8 floats per vertex, invBind 16 per bone"); return nullptr; }
  const size_t V = nv / 8;
  if (!exact_len(env, "loadMesh joints", nj, 4 * V) || !exact_len(env, "loadMesh weights", nw, 4 * V)) return nullptr;
  check(env, h->c, rz_load_mesh(h->c, v, j, w, (uint32_t)V, ib, (uint32_t)(nb / 16)));
  return nullptr;
}

// loadMorphs(ctx, Uint32Array offsets /*M+1*/, Uint32Array vertexIndex /*nnz*/, Float32Array delta /*3 nnz*/)
napi_value LoadMorphs(napi_env env, napi_callback_info info) {
  ARGS(4)
  Handle* h = unwrap(env, a[0]);
  size_t no = 0, ni = 0, nd = 0;
  const uint32_t* off = typed<uint32_t>(env, a[1], &no, napi_uint32_array);
  const uint32_t* idx = typed<uint32_t>(env, a[2], &ni, napi_uint32_array);
  const float* d = typed<float>(env, a[3], &nd, napi_float32_array);
  if (!h || !off || !idx || !d) return nullptr;
  const uint32_t M = no ? (uint32_t)(no - 1) : 0;
  const size_t nnz = M ? off[M] : 0;
  if (!need_len(env, "loadMorphs vertexIndex", ni, nnz) || !need_len(env, "loadMorphs delta", nd, 3 * nnz)) return nullptr;
  check(env, h->c, rz_load_morphs(h->c, off, idx, d, M));
  return nullptr;
}

// loadSdef(ctx, Uint32Array vertexIndex /*n*/, Float32Array c_r0_r1 /*9 n*/)
napi_value LoadSdef(napi_env env, napi_callback_info info) {
  ARGS(3)
  Handle* h = unwrap(env, a[0]);
  size_t ni = 0, nv = 0;
  const uint32_t* idx = typed<uint32_t>(env, a[1], &ni, napi_uint32_array);
  const float* vec = typed<float>(env, a[2], &nv, napi_float32_array);
  if (!h || !idx || !vec) return nullptr;
  if (!need_len(env, "loadSdef c_r0_r1", nv, 9 * ni)) return nullptr;
  check(env, h->c, rz_load_sdef(h->c, idx, vec, (uint32_t)ni));
  return nullptr;
}

// loadEdgeSize(ctx, Float32Array|null edgeSize /*V: Material.edgeSize of the material drawing the vertex, 0 = no outline*/)
napi_value LoadEdgeSize(napi_env env, napi_callback_info info) {
  ARGS(2)
  Handle* h = unwrap(env, a[0]);
  if (!h) return nullptr;
  size_t n = 0;
  const float* e = nullptr;
  if (!is_nullish(env, a[1])) {
    e = typed<float>(env, a[1], &n, napi_float32_array);
    rz_stats s;
    if (!e || !stats_of(env, h->c, &s) || !need_len(env, "loadEdgeSize", n, s.vertexCount)) return nullptr;
  }
  check(env, h->c, rz_load_edge_size(h->c, e));
  return nullptr;
}

// loadSkeleton(ctx, Int32Array parent /*B*/, Float32Array bindTranslation /*3B*/, Int32Array|null appendParent,
//              Float32Array|null appendRatio, Uint8Array|null appendRotate): the static part of Skeleton (model.ts:31-45)
napi_value LoadSkeleton(napi_env env, napi_callback_info info) {
  ARGS(6)
  Handle* h = unwrap(env, a[0]);
  size_t np = 0, nb = 0, n3 = 0, n4 = 0, n5 = 0;
  const int32_t* parent = typed<int32_t>(env, a[1], &np, napi_int32_array);
  const float* bind = typed<float>(env, a[2], &nb, napi_float32_array);
  if (!h || !parent || !bind) return nullptr;
  const int32_t* ap = is_nullish(env, a[3]) ? nullptr : typed<int32_t>(env, a[3], &n3, napi_int32_array);
  const float* ar = is_nullish(env, a[4]) ? nullptr : typed<float>(env, a[4], &n4, napi_float32_array);
  const uint8_t* arot = is_nullish(env, a[5]) ? nullptr : typed<uint8_t>(env, a[5], &n5, napi_uint8_array);
  if (!exact_len(env, "loadSkeleton bindTranslation", nb, 3 * np)) return nullptr;
  if ((ap && !need_len(env, "loadSkeleton appendParent", n3, np)) || (ar && !need_len(env, "loadSkeleton appendRatio", n4, np)) ||
      (arot && !need_len(env, "loadSkeleton appendRotate", n5, np)))
    return nullptr;
  check(env, h->c, rz_load_skeleton(h->c, parent, bind, ap, ar, arot, (uint32_t)np));
  return nullptr;
}

// (world | quats | clocks, P, instToPalette|null, K): shared argument checking of the three palette feeds
template <typename Fn>
napi_value palette_feed(napi_env env, napi_callback_info info, const char* what, size_t perPalettePerBone, Fn fn) {
  ARGS(5)
  Handle* h = unwrap(env, a[0]);
  size_t nw = 0, ni = 0;
  const float* data = typed<float>(env, a[1], &nw, napi_float32_array);
  if (!h || !data) return nullptr;
  const uint32_t P = u32(env, a[2]), K = u32(env, a[4]);
  const uint32_t* i2p = is_nullish(env, a[3]) ? nullptr : typed<uint32_t>(env, a[3], &ni, napi_uint32_array);
  rz_stats s;
  if (!stats_of(env, h->c, &s)) return nullptr;
  const size_t per = perPalettePerBone ? perPalettePerBone * s.boneCount : 1;
  if (!need_len(env, what, nw, (size_t)P * per)) return nullptr;
  if (i2p && !need_len(env, "instToPalette", ni, K)) return nullptr;
  check(env, h->c, fn(h->c, data, P, i2p, K));
  return nullptr;
}
// setPalettes(ctx, Float32Array world /*P*B*16, Model.getBoneWorldMatrices() per palette*/, P, Uint32Array|null instToPalette, K)
//   replaces queue.writeBuffer(worldMatrixBuffer) + computeSkinMatrices (engine.ts:2383-2402)
napi_value SetPalettes(napi_env env, napi_callback_info info) { return palette_feed(env, info, "setPalettes world", 16, rz_set_palettes); }
// setLocalRotations(ctx, Float32Array quats /*P*B*4, SkeletonRuntime.localRotations*/, P, instToPalette|null, K): pose on the device
napi_value SetLocalRotations(napi_env env, napi_callback_info info) { return palette_feed(env, info, "setLocalRotations quats", 4, rz_set_local_rotations); }
// setInstanceClocks(ctx, Float32Array nowMs /*P*/, P, instToPalette|null, K): tween table / clip evaluated per clock value
napi_value SetInstanceClocks(napi_env env, napi_callback_info info) { return palette_feed(env, info, "setInstanceClocks nowMs", 0, rz_set_instance_clocks); }

// setTweens(ctx, start /*4B*/, target /*4B*/, startMs /*B*/, durMs /*B*/, Uint8Array active /*B*/, rest /*4B*/): RotationTweenState (model.ts:62-68)
napi_value SetTweens(napi_env env, napi_callback_info info) {
  ARGS(7)
  Handle* h = unwrap(env, a[0]);
  size_t n1 = 0, n2 = 0, n3 = 0, n4 = 0, n5 = 0, n6 = 0;
  const float* sq = typed<float>(env, a[1], &n1, napi_float32_array);
  const float* tq = typed<float>(env, a[2], &n2, napi_float32_array);
  const float* sm = typed<float>(env, a[3], &n3, napi_float32_array);
  const float* dm = typed<float>(env, a[4], &n4, napi_float32_array);
  const uint8_t* act = typed<uint8_t>(env, a[5], &n5, napi_uint8_array);
  const float* rq = typed<float>(env, a[6], &n6, napi_float32_array);
  rz_stats s;
  if (!h || !sq || !tq || !sm || !dm || !act || !rq || !stats_of(env, h->c, &s)) return nullptr;
  const size_t B = s.boneCount;
  if (!need_len(env, "setTweens start", n1, 4 * B) || !need_len(env, "setTweens target", n2, 4 * B) || !need_len(env, "setTweens startMs", n3, B) ||
      !need_len(env, "setTweens durationMs", n4, B) || !need_len(env, "setTweens active", n5, B) || !need_len(env, "setTweens rest", n6, 4 * B))
    return nullptr;
  check(env, h->c, rz_set_tweens(h->c, sq, tq, sm, dm, act, rq));
  return nullptr;
}

// loadAnimation(ctx, Uint32Array keyOffsets /*B+1*/ | null (unload), Float32Array keyTimesMs, Float32Array keyQuats /*4 per key*/,
//               Float32Array|null restQuat /*4B*/): the clip playAnimation schedules (engine.ts:1451-1553), evaluated per clock value
napi_value LoadAnimation(napi_env env, napi_callback_info info) {
  ARGS(5)
  Handle* h = unwrap(env, a[0]);
  if (!h) return nullptr;
  if (is_nullish(env, a[1])) {
    check(env, h->c, rz_load_animation(h->c, nullptr, nullptr, nullptr, nullptr));
    return nullptr;
  }
  size_t no = 0, nt = 0, nq = 0, nr = 0;
  const uint32_t* off = typed<uint32_t>(env, a[1], &no, napi_uint32_array);
  const float* t = typed<float>(env, a[2], &nt, napi_float32_array);
  const float* q = typed<float>(env, a[3], &nq, napi_float32_array);
  const float* rest = is_nullish(env, a[4]) ? nullptr : typed<float>(env, a[4], &nr, napi_float32_array);
  rz_stats s;
  if (!off || !t || !q || !stats_of(env, h->c, &s)) return nullptr;
  if (!exact_len(env, "loadAnimation keyOffsets", no, (size_t)s.boneCount + 1)) return nullptr;
  const size_t n = off[s.boneCount];
  if (!need_len(env, "loadAnimation keyTimesMs", nt, n) || !need_len(env, "loadAnimation keyQuats", nq, 4 * n)) return nullptr;
  if (rest && !need_len(env, "loadAnimation restQuat", nr, 4 * (size_t)s.boneCount)) return nullptr;
  check(env, h->c, rz_load_animation(h->c, off, t, q, rest));
  return nullptr;
}

// setMorphWeights(ctx, Float32Array w /*K*M_active*/, Uint32Array activeIds /*M_active*/, K)
napi_value SetMorphWeights(napi_env env, napi_callback_info info) {
  ARGS(4)
  Handle* h = unwrap(env, a[0]);
  size_t nw = 0, ni = 0;
  const float* w = typed<float>(env, a[1], &nw, napi_float32_array);
  const uint32_t* ids = typed<uint32_t>(env, a[2], &ni, napi_uint32_array);
  if (!h || !w || !ids) return nullptr;
  const uint32_t K = u32(env, a[3]);
  if (!need_len(env, "setMorphWeights w", nw, (size_t)K * ni)) return nullptr;
  check(env, h->c, rz_set_morph_weights(h->c, w, ids, (uint32_t)ni, K));
  return nullptr;
}

napi_value Deform(napi_env env, napi_callback_info info) {
  ARGS(3)
  Handle* h = unwrap(env, a[0]);
  if (h) check(env, h->c, rz_deform(h->c, u32(env, a[1]), u32(env, a[2])));
  return nullptr;
}

napi_value Sync(napi_env env, napi_callback_info info) {
  ARGS(1)
  Handle* h = unwrap(env, a[0]);
  if (h) check(env, h->c, rz_sync(h->c));
  return nullptr;
}

// (pos /*3V*/, nrm|null /*3V*/) of a read-back; false = threw
bool read_targets(napi_env env, Handle* h, napi_value vpos, napi_value vnrm, float** pos, float** nrm) {
  size_t np = 0, nn = 0;
  *pos = is_nullish(env, vpos) ? nullptr : typed<float>(env, vpos, &np, napi_float32_array);
  *nrm = is_nullish(env, vnrm) ? nullptr : typed<float>(env, vnrm, &nn, napi_float32_array);
  rz_stats s;
  if (!stats_of(env, h->c, &s)) return false;
  if (*pos && !need_len(env, "read-back positions", np, 3 * (size_t)s.vertexCount)) return false;
  if (*nrm && !need_len(env, "read-back normals", nn, 3 * (size_t)s.vertexCount)) return false;
  return true;
}

// readInstance(ctx, inst, Float32Array|null pos /*3V*/, Float32Array|null nrm /*3V*/)
napi_value ReadInstance(napi_env env, napi_callback_info info) {
  ARGS(4)
  Handle* h = unwrap(env, a[0]);
  float *pos, *nrm;
  if (!h || !read_targets(env, h, a[2], a[3], &pos, &nrm)) return nullptr;
  check(env, h->c, rz_read_instance(h->c, u32(env, a[1]), pos, nrm));
  return nullptr;
}

// readInstanceAsync(ctx, inst, pos, nrm|null): queued behind the frame's deform, returns at once; readWait(ctx) blocks until the
// arrays are filled.  With RZ_FLAG_DOUBLE_BUFFER (0x40) the next frame is deformed meanwhile.  The arrays are referenced here
// until readWait, so the GC cannot free memory the copy engine still writes.
napi_value ReadInstanceAsync(napi_env env, napi_callback_info info) {
  ARGS(4)
  Handle* h = unwrap(env, a[0]);
  float *pos, *nrm;
  if (!h || !read_targets(env, h, a[2], a[3], &pos, &nrm)) return nullptr;
  for (int i = 2; i < 4; ++i) {
    if (is_nullish(env, a[i])) continue;
    napi_ref r;
    NAPI_OK(env, napi_create_reference(env, a[i], 1, &r));
    h->held.push_back(r);
  }
  check(env, h->c, rz_read_instance_async(h->c, u32(env, a[1]), pos, nrm));
  return nullptr;
}

napi_value ReadWait(napi_env env, napi_callback_info info) {
  ARGS(1)
  Handle* h = unwrap(env, a[0]);
  if (!h) return nullptr;
  const bool ok = check(env, h->c, rz_read_wait(h->c));
  for (napi_ref r : h->held) napi_delete_reference(env, r);
  h->held.clear();
  (void)ok;
  return nullptr;
}

// readOutline(ctx, inst, Float32Array hull /*3V*/): the outline pass' expanded positions (engine.ts:458-461), RZ_FLAG_OUTLINE
napi_value ReadOutline(napi_env env, napi_callback_info info) {
  ARGS(3)
  Handle* h = unwrap(env, a[0]);
  size_t n = 0;
  float* hull = typed<float>(env, a[2], &n, napi_float32_array);
  rz_stats s;
  if (!h || !hull || !stats_of(env, h->c, &s) || !need_len(env, "readOutline", n, 3 * (size_t)s.vertexCount)) return nullptr;
  check(env, h->c, rz_read_outline(h->c, u32(env, a[1]), hull));
  return nullptr;
}

// readInterleaved(ctx, inst, Float32Array vtx8 /*8V*/): one instance in the reference's vertex-buffer layout, RZ_FLAG_INTERLEAVED
napi_value ReadInterleaved(napi_env env, napi_callback_info info) {
  ARGS(3)
  Handle* h = unwrap(env, a[0]);
  size_t n = 0;
  float* v = typed<float>(env, a[2], &n, napi_float32_array);
  rz_stats s;
  if (!h || !v || !stats_of(env, h->c, &s) || !need_len(env, "readInterleaved", n, 8 * (size_t)s.vertexCount)) return nullptr;
  check(env, h->c, rz_read_interleaved(h->c, u32(env, a[1]), v));
  return nullptr;
}

// readBounds(ctx, first, count, Float32Array minmax /*6 per instance*/), RZ_FLAG_BOUNDS
napi_value ReadBounds(napi_env env, napi_callback_info info) {
  ARGS(4)
  Handle* h = unwrap(env, a[0]);
  size_t n = 0;
  float* mm = typed<float>(env, a[3], &n, napi_float32_array);
  if (!h || !mm) return nullptr;
  const uint32_t first = u32(env, a[1]), count = u32(env, a[2]);
  if (!need_len(env, "readBounds", n, 6 * (size_t)count)) return nullptr;
  check(env, h->c, rz_read_bounds(h->c, first, count, mm));
  return nullptr;
}

// readWorldMatrices(ctx, palette, Float32Array world /*16B, column-major*/): bone world matrices of a pose evaluated on the
// device, for host code that needs them back (Physics.step's kinematic bodies, physics.ts:649-703)
napi_value ReadWorldMatrices(napi_env env, napi_callback_info info) {
  ARGS(3)
  Handle* h = unwrap(env, a[0]);
  size_t n = 0;
  float* w = typed<float>(env, a[2], &n, napi_float32_array);
  rz_stats s;
  if (!h || !w || !stats_of(env, h->c, &s) || !need_len(env, "readWorldMatrices", n, 16 * (size_t)s.boneCount)) return nullptr;
  check(env, h->c, rz_read_world_matrices(h->c, u32(env, a[1]), w));
  return nullptr;
}

// getVertexOrder(ctx, Uint32Array order /*V*/): stored position -> caller vertex id (RZ_FLAG_REORDER_VERTICES)
napi_value GetVertexOrder(napi_env env, napi_callback_info info) {
  ARGS(2)
  Handle* h = unwrap(env, a[0]);
  size_t n = 0;
  uint32_t* o = typed<uint32_t>(env, a[1], &n, napi_uint32_array);
  rz_stats s;
  if (!h || !o || !stats_of(env, h->c, &s) || !need_len(env, "getVertexOrder", n, s.vertexCount)) return nullptr;
  check(env, h->c, rz_get_vertex_order(h->c, o));
  return nullptr;
}

// getOutputLayout(ctx) -> {base: BigInt device pointer, instanceStride, vertexStride, positionOffset, normalOffset, hullOffset, uvOffset}
//   (bytes; an attribute that is not produced is -1) — for zero-copy consumers (CUDA / Vulkan interop)
napi_value GetOutputLayout(napi_env env, napi_callback_info info) {
  ARGS(1)
  Handle* h = unwrap(env, a[0]);
  rz_output_layout l;
  if (!h || !check(env, h->c, rz_get_output_layout(h->c, &l))) return nullptr;
  napi_value o, base;
  NAPI_OK(env, napi_create_object(env, &o));
  NAPI_OK(env, napi_create_bigint_uint64(env, (uint64_t)(uintptr_t)l.base, &base));
  NAPI_OK(env, napi_set_named_property(env, o, "base", base));
  auto put = [&](const char* k, size_t v) {
    napi_value n;
    napi_create_double(env, v == RZ_NO_ATTRIBUTE ? -1.0 : (double)v, &n);
    napi_set_named_property(env, o, k, n);
  };
  put("instanceStride", l.instanceStride); put("vertexStride", l.vertexStride); put("positionOffset", l.positionOffset);
  put("normalOffset", l.normalOffset); put("hullOffset", l.hullOffset); put("uvOffset", l.uvOffset);
  return o;
}

// loadRigidBodies(ctx, Int32Array boneIndex /*n*/, Uint8Array dynamic /*n*/, Float32Array bodyOffsetMatrixInverse /*16 n*/)
//   and applyBodyTransforms(ctx, Float32Array posQuat /*P*n*7*/, P): physics.ts:714-751 on the device, the solver stays in JS
napi_value LoadRigidBodies(napi_env env, napi_callback_info info) {
  ARGS(4)
  Handle* h = unwrap(env, a[0]);
  size_t nb = 0, nd = 0, no = 0;
  const int32_t* bi = typed<int32_t>(env, a[1], &nb, napi_int32_array);
  const uint8_t* dy = typed<uint8_t>(env, a[2], &nd, napi_uint8_array);
  const float* oi = typed<float>(env, a[3], &no, napi_float32_array);
  if (!h || !bi || !dy || !oi) return nullptr;
  if (!exact_len(env, "loadRigidBodies dynamic", nd, nb) || !exact_len(env, "loadRigidBodies bodyOffsetMatrixInverse", no, 16 * nb)) return nullptr;
  h->held.shrink_to_fit();
  check(env, h->c, rz_load_rigid_bodies(h->c, bi, dy, oi, (uint32_t)nb));
  napi_value n;                                             // body count back to the caller (applyBodyTransforms sizes against it)
  NAPI_OK(env, napi_create_double(env, (double)nb, &n));
  return n;
}

// applyBodyTransforms(ctx, Float32Array posQuat /*P*n*7*/, P, n): n = the body count given to loadRigidBodies
napi_value ApplyBodyTransforms(napi_env env, napi_callback_info info) {
  ARGS(4)
  Handle* h = unwrap(env, a[0]);
  size_t n = 0;
  const float* pq = typed<float>(env, a[1], &n, napi_float32_array);
  if (!h || !pq) return nullptr;
  const uint32_t P = u32(env, a[2]), bodies = u32(env, a[3]);
  if (!exact_len(env, "applyBodyTransforms posQuat", n, (size_t)P * bodies * 7)) return nullptr;
  check(env, h->c, rz_apply_body_transforms(h->c, pq, P));
  return nullptr;
}

// getStats(ctx) -> EngineStats (engine.ts:16-20: fps, frameTime, gpuMemory) + the path's own counters
napi_value GetStats(napi_env env, napi_callback_info info) {
  ARGS(1)
  Handle* h = unwrap(env, a[0]);
  rz_stats s;
  if (!h || !stats_of(env, h->c, &s)) return nullptr;
  napi_value o;
  NAPI_OK(env, napi_create_object(env, &o));
  auto put = [&](const char* k, double v) {
    napi_value n;
    napi_create_double(env, v, &n);
    napi_set_named_property(env, o, k, n);
  };
  put("fps", s.fps); put("frameTime", s.frameTime); put("gpuMemory", s.gpuMemory);
  put("vertsPerSec", s.vertsPerSec); put("achievedGBs", s.achievedGBs); put("algorithmicBytes", s.algorithmicBytes);
  put("vertexCount", s.vertexCount); put("boneCount", s.boneCount); put("instanceCount", s.instanceCount);
  put("morphCount", s.morphCount); put("sdefCount", s.sdefCount); put("kernelLaunches", (double)s.kernelLaunches);
  return o;
}

napi_value Init(napi_env env, napi_value exports) {
#define RZ_FN(name, fn) {name, nullptr, fn, nullptr, nullptr, nullptr, napi_default, nullptr}
  const napi_property_descriptor d[] = {
      RZ_FN("create", Create), RZ_FN("destroy", Destroy), RZ_FN("loadMesh", LoadMesh), RZ_FN("loadMorphs", LoadMorphs),
      RZ_FN("loadSdef", LoadSdef), RZ_FN("loadEdgeSize", LoadEdgeSize), RZ_FN("loadSkeleton", LoadSkeleton),
      RZ_FN("setPalettes", SetPalettes), RZ_FN("setLocalRotations", SetLocalRotations), RZ_FN("setTweens", SetTweens),
      RZ_FN("setInstanceClocks", SetInstanceClocks), RZ_FN("loadAnimation", LoadAnimation), RZ_FN("setMorphWeights", SetMorphWeights),
      RZ_FN("deform", Deform), RZ_FN("sync", Sync), RZ_FN("readInstance", ReadInstance), RZ_FN("readInstanceAsync", ReadInstanceAsync),
      RZ_FN("readWait", ReadWait), RZ_FN("readOutline", ReadOutline), RZ_FN("readInterleaved", ReadInterleaved),
      RZ_FN("readBounds", ReadBounds), RZ_FN("readWorldMatrices", ReadWorldMatrices), RZ_FN("getVertexOrder", GetVertexOrder),
      RZ_FN("getOutputLayout", GetOutputLayout), RZ_FN("loadRigidBodies", LoadRigidBodies), RZ_FN("applyBodyTransforms", ApplyBodyTransforms),
      RZ_FN("getStats", GetStats),
  };
#undef RZ_FN
  napi_define_properties(env, exports, sizeof d / sizeof d[0], d);
  return exports;
}

}  // namespace

NAPI_MODULE(NODE_GYP_MODULE_NAME, Init)
#endif  // __has_include(<node_api.h>)
