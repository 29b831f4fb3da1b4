// rze_b200_napi.cc — thin N-API addon over the C ABI (include/rze_b200.h).
//
// Pure marshalling: every JS call maps 1:1 onto an rz_* entry point; typed arrays are passed through without
// copies (the C ABI copies what it keeps).  A non-zero status becomes a thrown JS Error carrying rz_last_error(),
// which is the reference's convention for fatal conditions (engine.ts:161,167,1828: `throw new Error(...)`).
//
// NOT BUILT IN THIS IMAGE: no Node.js / node_api.h exists here or on the GPU box, so this file is
// compiled only where <node_api.h> is available:
//     g++ -O2 -fPIC -shared -I$(node -p "require('node-addon-api').include_dir") -Iinclude napi/rze_b200_napi.cc
//         -Lreze-engine_b200/lib -lrze_b200 -o rze_b200.node          (one command line)
// The tested boundary is the C ABI itself (tests/ drive it through ctypes with the same argument marshalling).
#if __has_include(<node_api.h>)
#include <node_api.h>

#include <cstdint>
#include <cstring>
#include <string>

#include "../include/rze_b200.h"

namespace {

#define NAPI_OK(env, call)                                         \
  do {                                                             \
    if ((call) != napi_ok) {                                       \
      napi_throw_error((env), nullptr, "N-API call failed: " #call); \
      return nullptr;                                              \
    }                                                              \
  } while (0)

rz_ctx* unwrap(napi_env env, napi_value v) {
  void* p = nullptr;
  if (napi_get_value_external(env, v, &p) != napi_ok || !p) {
    napi_throw_type_error(env, nullptr, "expected a deform context");
    return nullptr;
  }
  return static_cast<rz_ctx*>(p);
}

bool check(napi_env env, rz_ctx* c, int32_t st) {
  if (st == RZ_OK) return true;
  napi_throw_error(env, nullptr, rz_last_error(c));
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

uint32_t u32(napi_env env, napi_value v) {
  uint32_t x = 0;
  napi_get_value_uint32(env, v, &x);
  return x;
}

void finalize(napi_env, void* data, void*) { rz_destroy(static_cast<rz_ctx*>(data)); }

// create({device, maxInstances, flags}) -> external
napi_value Create(napi_env env, napi_callback_info info) {
  size_t argc = 3;
  napi_value a[3];
  NAPI_OK(env, napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
  rz_config cfg;
  std::memset(&cfg, 0, sizeof cfg);
  cfg.struct_size = sizeof cfg;
  cfg.device = (int32_t)u32(env, a[0]);
  cfg.max_instances = u32(env, a[1]);
  cfg.flags = u32(env, a[2]);
  rz_ctx* c = nullptr;
  if (!check(env, nullptr, rz_create(&cfg, &c))) return nullptr;
  napi_value ext;
  NAPI_OK(env, napi_create_external(env, c, finalize, nullptr, &ext));
  return ext;
}

// loadMesh(ctx, Float32Array vtx8, Uint16Array joints, Uint8Array weights, Float32Array invBind)
//   exactly Model.getVertices() / getSkinning() / getBoneInverseBindMatrices() of the reference (model.ts:196-200, 47-50, 321)
napi_value LoadMesh(napi_env env, napi_callback_info info) {
  size_t argc = 5;
  napi_value a[5];
  NAPI_OK(env, napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
  rz_ctx* c = unwrap(env, a[0]);
  size_t nv = 0, nj = 0, nw = 0, nb = 0;
  const float* v = typed<float>(env, a[1], &nv, napi_float32_array);
  const uint16_t* j = typed<uint16_t>(env, a[2], &nj, napi_uint16_array);
  const uint8_t* w = typed<uint8_t>(env, a[3], &nw, napi_uint8_array);
  const float* ib = typed<float>(env, a[4], &nb, napi_float32_array);
  if (!c || !v || !j || !w || !ib) return nullptr;
  check(env, c, rz_load_mesh(c, v, j, w, (uint32_t)(nv / 8), ib, (uint32_t)(nb / 16)));
  return nullptr;
}

napi_value LoadMorphs(napi_env env, napi_callback_info info) {
  size_t argc = 4;
  napi_value a[4];
  NAPI_OK(env, napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
  rz_ctx* c = unwrap(env, a[0]);
  size_t no = 0, ni = 0, nd = 0;
  const uint32_t* off = typed<uint32_t>(env, a[1], &no, napi_uint32_array);
  const uint32_t* idx = typed<uint32_t>(env, a[2], &ni, napi_uint32_array);
  const float* d = typed<float>(env, a[3], &nd, napi_float32_array);
  if (!c || !off) return nullptr;
  check(env, c, rz_load_morphs(c, off, idx, d, no ? (uint32_t)(no - 1) : 0));
  return nullptr;
}

napi_value LoadSdef(napi_env env, napi_callback_info info) {
  size_t argc = 3;
  napi_value a[3];
  NAPI_OK(env, napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
  rz_ctx* c = unwrap(env, a[0]);
  size_t ni = 0, nv = 0;
  const uint32_t* idx = typed<uint32_t>(env, a[1], &ni, napi_uint32_array);
  const float* vec = typed<float>(env, a[2], &nv, napi_float32_array);
  if (!c) return nullptr;
  check(env, c, rz_load_sdef(c, idx, vec, (uint32_t)ni));
  return nullptr;
}

// setPalettes(ctx, Float32Array world /*P*B*16, Model.getBoneWorldMatrices() per palette*/, P, Uint32Array|null instToPalette, K)
//   replaces queue.writeBuffer(worldMatrixBuffer) + computeSkinMatrices (engine.ts:2383-2402)
napi_value SetPalettes(napi_env env, napi_callback_info info) {
  size_t argc = 5;
  napi_value a[5];
  NAPI_OK(env, napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
  rz_ctx* c = unwrap(env, a[0]);
  size_t nw = 0, ni = 0;
  const float* world = typed<float>(env, a[1], &nw, napi_float32_array);
  napi_valuetype vt;
  napi_typeof(env, a[3], &vt);
  const uint32_t* i2p = (vt == napi_null || vt == napi_undefined) ? nullptr : typed<uint32_t>(env, a[3], &ni, napi_uint32_array);
  if (!c || !world) return nullptr;
  check(env, c, rz_set_palettes(c, world, u32(env, a[2]), i2p, u32(env, a[4])));
  return nullptr;
}

napi_value SetMorphWeights(napi_env env, napi_callback_info info) {
  size_t argc = 4;
  napi_value a[4];
  NAPI_OK(env, napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
  rz_ctx* c = unwrap(env, a[0]);
  size_t nw = 0, ni = 0;
  const float* w = typed<float>(env, a[1], &nw, napi_float32_array);
  const uint32_t* ids = typed<uint32_t>(env, a[2], &ni, napi_uint32_array);
  if (!c) return nullptr;
  check(env, c, rz_set_morph_weights(c, w, ids, (uint32_t)ni, u32(env, a[3])));
  return nullptr;
}

napi_value Deform(napi_env env, napi_callback_info info) {
  size_t argc = 3;
  napi_value a[3];
  NAPI_OK(env, napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
  rz_ctx* c = unwrap(env, a[0]);
  if (c) check(env, c, rz_deform(c, u32(env, a[1]), u32(env, a[2])));
  return nullptr;
}

napi_value Sync(napi_env env, napi_callback_info info) {
  size_t argc = 1;
  napi_value a[1];
  NAPI_OK(env, napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
  rz_ctx* c = unwrap(env, a[0]);
  if (c) check(env, c, rz_sync(c));
  return nullptr;
}

// readInstance(ctx, inst, Float32Array pos /*3V*/, Float32Array|null nrm /*3V*/)
napi_value ReadInstance(napi_env env, napi_callback_info info) {
  size_t argc = 4;
  napi_value a[4];
  NAPI_OK(env, napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
  rz_ctx* c = unwrap(env, a[0]);
  size_t np = 0, nn = 0;
  float* pos = typed<float>(env, a[2], &np, napi_float32_array);
  napi_valuetype vt;
  napi_typeof(env, a[3], &vt);
  float* nrm = (vt == napi_null || vt == napi_undefined) ? nullptr : typed<float>(env, a[3], &nn, napi_float32_array);
  if (c) check(env, c, rz_read_instance(c, u32(env, a[1]), pos, nrm));
  return nullptr;
}

// loadEdgeSize(ctx, Float32Array|null edgeSize /*V: Material.edgeSize of the material drawing the vertex, 0 = no outline*/)
napi_value LoadEdgeSize(napi_env env, napi_callback_info info) {
  size_t argc = 2;
  napi_value a[2];
  NAPI_OK(env, napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
  rz_ctx* c = unwrap(env, a[0]);
  napi_valuetype vt;
  napi_typeof(env, a[1], &vt);
  size_t n = 0;
  const float* e = (vt == napi_null || vt == napi_undefined) ? nullptr : typed<float>(env, a[1], &n, napi_float32_array);
  if (c) check(env, c, rz_load_edge_size(c, e));
  return nullptr;
}

// readOutline(ctx, inst, Float32Array hull /*3V*/): the outline pass' expanded positions (engine.ts:458-461), RZ_FLAG_OUTLINE
napi_value ReadOutline(napi_env env, napi_callback_info info) {
  size_t argc = 3;
  napi_value a[3];
  NAPI_OK(env, napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
  rz_ctx* c = unwrap(env, a[0]);
  size_t n = 0;
  float* hull = typed<float>(env, a[2], &n, napi_float32_array);
  if (c && hull) check(env, c, rz_read_outline(c, u32(env, a[1]), hull));
  return nullptr;
}

// readInterleaved(ctx, inst, Float32Array vtx8 /*8V*/): one instance in the reference's vertex-buffer layout, RZ_FLAG_INTERLEAVED
napi_value ReadInterleaved(napi_env env, napi_callback_info info) {
  size_t argc = 3;
  napi_value a[3];
  NAPI_OK(env, napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
  rz_ctx* c = unwrap(env, a[0]);
  size_t n = 0;
  float* v = typed<float>(env, a[2], &n, napi_float32_array);
  if (c && v) check(env, c, rz_read_interleaved(c, u32(env, a[1]), v));
  return nullptr;
}

// getOutputLayout(ctx) -> {base: BigInt device pointer, instanceStride, vertexStride, positionOffset, normalOffset, hullOffset, uvOffset}
//   (bytes; an attribute that is not produced is -1) — for zero-copy consumers (CUDA / Vulkan interop)
napi_value GetOutputLayout(napi_env env, napi_callback_info info) {
  size_t argc = 1;
  napi_value a[1];
  NAPI_OK(env, napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
  rz_ctx* c = unwrap(env, a[0]);
  rz_output_layout l;
  if (!c || !check(env, c, rz_get_output_layout(c, &l))) return nullptr;
  napi_value o, base;
  NAPI_OK(env, napi_create_object(env, &o));
  napi_create_bigint_uint64(env, (uint64_t)(uintptr_t)l.base, &base);
  napi_set_named_property(env, o, "base", base);
  auto put = [&](const char* k, size_t v) {
    napi_value n;
    napi_create_double(env, v == RZ_NO_ATTRIBUTE ? -1.0 : (double)v, &n);
    napi_set_named_property(env, o, k, n);
  };
  put("instanceStride", l.instanceStride); put("vertexStride", l.vertexStride); put("positionOffset", l.positionOffset);
  put("normalOffset", l.normalOffset); put("hullOffset", l.hullOffset); put("uvOffset", l.uvOffset);
  return o;
}

// readInstanceAsync(ctx, inst, Float32Array pos, Float32Array|null nrm): queued behind the frame's deform, returns at once;
// readWait(ctx) blocks until the arrays are filled.  With RZ_FLAG_DOUBLE_BUFFER (0x40) the next frame is deformed meanwhile.
// (The arrays must stay alive until readWait; a Promise-returning wrapper belongs in ts/engine.ts.)
napi_value ReadInstanceAsync(napi_env env, napi_callback_info info) {
  size_t argc = 4;
  napi_value a[4];
  NAPI_OK(env, napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
  rz_ctx* c = unwrap(env, a[0]);
  size_t np = 0, nn = 0;
  float* pos = typed<float>(env, a[2], &np, napi_float32_array);
  napi_valuetype vt;
  napi_typeof(env, a[3], &vt);
  float* nrm = (vt == napi_null || vt == napi_undefined) ? nullptr : typed<float>(env, a[3], &nn, napi_float32_array);
  if (c) check(env, c, rz_read_instance_async(c, u32(env, a[1]), pos, nrm));
  return nullptr;
}

napi_value ReadWait(napi_env env, napi_callback_info info) {
  size_t argc = 1;
  napi_value a[1];
  NAPI_OK(env, napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
  rz_ctx* c = unwrap(env, a[0]);
  if (c) check(env, c, rz_read_wait(c));
  return nullptr;
}

// loadRigidBodies(ctx, Int32Array boneIndex, Uint8Array dynamic, Float32Array bodyOffsetMatrixInverse /*16 per body*/)
//   and applyBodyTransforms(ctx, Float32Array posQuat /*P*n*7*/, P): physics.ts:714-751 on the device, the solver stays in JS
napi_value LoadRigidBodies(napi_env env, napi_callback_info info) {
  size_t argc = 4;
  napi_value a[4];
  NAPI_OK(env, napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
  rz_ctx* c = unwrap(env, a[0]);
  size_t nb = 0, nd = 0, no = 0;
  const int32_t* bi = typed<int32_t>(env, a[1], &nb, napi_int32_array);
  const uint8_t* dy = typed<uint8_t>(env, a[2], &nd, napi_uint8_array);
  const float* oi = typed<float>(env, a[3], &no, napi_float32_array);
  if (c && bi && dy && oi) check(env, c, rz_load_rigid_bodies(c, bi, dy, oi, (uint32_t)nb));
  return nullptr;
}

napi_value ApplyBodyTransforms(napi_env env, napi_callback_info info) {
  size_t argc = 3;
  napi_value a[3];
  NAPI_OK(env, napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
  rz_ctx* c = unwrap(env, a[0]);
  size_t n = 0;
  const float* pq = typed<float>(env, a[1], &n, napi_float32_array);
  if (c && pq) check(env, c, rz_apply_body_transforms(c, pq, u32(env, a[2])));
  return nullptr;
}

// getStats(ctx) -> {fps, frameTime, gpuMemory, vertsPerSec, achievedGBs}  (EngineStats, engine.ts:16-20 + additions)
napi_value GetStats(napi_env env, napi_callback_info info) {
  size_t argc = 1;
  napi_value a[1];
  NAPI_OK(env, napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
  rz_ctx* c = unwrap(env, a[0]);
  rz_stats s;
  if (!c || !check(env, c, rz_get_stats(c, &s))) return nullptr;
  napi_value o;
  NAPI_OK(env, napi_create_object(env, &o));
  auto put = [&](const char* k, double v) {
    napi_value n;
    napi_create_double(env, v, &n);
    napi_set_named_property(env, o, k, n);
  };
  put("fps", s.fps); put("frameTime", s.frameTime); put("gpuMemory", s.gpuMemory);
  put("vertsPerSec", s.vertsPerSec); put("achievedGBs", s.achievedGBs); put("algorithmicBytes", s.algorithmicBytes);
  return o;
}

napi_value Init(napi_env env, napi_value exports) {
  const napi_property_descriptor d[] = {
      {"create", nullptr, Create, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"loadMesh", nullptr, LoadMesh, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"loadMorphs", nullptr, LoadMorphs, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"loadSdef", nullptr, LoadSdef, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"setPalettes", nullptr, SetPalettes, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"setMorphWeights", nullptr, SetMorphWeights, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"deform", nullptr, Deform, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"sync", nullptr, Sync, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"readInstance", nullptr, ReadInstance, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"getStats", nullptr, GetStats, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"loadEdgeSize", nullptr, LoadEdgeSize, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"readOutline", nullptr, ReadOutline, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"readInterleaved", nullptr, ReadInterleaved, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"getOutputLayout", nullptr, GetOutputLayout, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"readInstanceAsync", nullptr, ReadInstanceAsync, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"loadRigidBodies", nullptr, LoadRigidBodies, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"applyBodyTransforms", nullptr, ApplyBodyTransforms, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"readWait", nullptr, ReadWait, nullptr, nullptr, nullptr, napi_default, nullptr},
  };
  napi_define_properties(env, exports, sizeof d / sizeof d[0], d);
  return exports;
}

}  // namespace

NAPI_MODULE(NODE_GYP_MODULE_NAME, Init)
#endif  // __has_include(<node_api.h>)
