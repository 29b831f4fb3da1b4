/*
 * rz_oracle.c — CPU restatement of the reference's deform path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may build, load or call this file.  The
 * product (reze-engine_b200/) never does; it has no CPU path at all.
 *
 * Pinning status:
 *   - loader integers (joints/weights quantisation, clamp + renormalise to 255, inverse
 *     bind): PINNED bit-exactly by the reference's own dump web/app/tutorial/model.json
 *     of 塞尔凯特.pmx (tests/test_loader.py, run where /root/reference exists;
 *     digests committed under tests/golden/).
 *   - float blend (engine.ts:253-272): the reference stores no golden skinned output and
 *     cannot be executed here (TypeScript + WGSL, no JS runtime / WebGPU in the image), so
 *     float parity is oracle-vs-kernel with this file following the WGSL statement by
 *     statement; anchored by the analytic T-pose identity (bind pose => output == input).
 *   - vertex morphs and SDEF: NOT in the reference (pmx-loader.ts:141-155, 450-553 skip the
 *     data) => "parity unpinned"; semantics are the ones SURVEY 8c specifies.
 *
 * Plain C99, no FMA contraction (build with -ffp-contract=off), f32 and f64 variants.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define REAL float
#define SUF _f32
#define SQRT sqrtf
#define ACOS acosf
#define SIN sinf
#include "rz_oracle_body.inc"
#undef REAL
#undef SUF
#undef SQRT
#undef ACOS
#undef SIN

#define REAL double
#define SUF _f64
#define SQRT sqrt
#define ACOS acos
#define SIN sin
#include "rz_oracle_body.inc"
#undef REAL
#undef SUF
#undef SQRT
#undef ACOS
#undef SIN

/* ---- loader integer paths ------------------------------------------------------------------- */
static double js_round(double x) { return (x != x) ? x : floor(x + 0.5); }          /* Math.round */
static double js_min(double a, double b) { return (a != a || b != b) ? NAN : (a < b ? a : b); }
static double js_max(double a, double b) { return (a != a || b != b) ? NAN : (a > b ? a : b); }
static uint8_t to_u8(double x) { return (x != x || isinf(x)) ? 0 : (uint8_t)((long long)trunc(x) & 0xFF); }

/* BDEF2 / SDEF weight (pmx-loader.ts:145-151) */
void orc_quantize_bdef2(float w0f, uint8_t out[2]) {
  const double w0 = js_max(0, js_min(255, js_round((double)w0f * 255)));
  const double w1 = js_max(0, js_min(255, 255 - w0));
  out[0] = to_u8(w0);
  out[1] = to_u8(w1);
}

/* BDEF4 / QDEF weights (pmx-loader.ts:163-179) */
void orc_quantize_bdef4(const float wf[4], uint8_t out[4]) {
  double w8[4], sum = 0;
  for (int k = 0; k < 4; ++k) {
    w8[k] = js_round(js_max(0, js_min(1, (double)wf[k])) * 255);
    sum += w8[k];
  }
  out[0] = 255; out[1] = out[2] = out[3] = 0;
  if (sum == 0) return;
  const double scale = 255 / sum;
  double accum = 0;
  for (int k = 0; k < 3; ++k) {
    const double v = js_max(0, js_min(255, js_round(w8[k] * scale)));
    out[k] = to_u8(v);
    accum += v;
  }
  out[3] = to_u8(js_max(0, js_min(255, 255 - accum)));
}

/* toModel clamp + renormalise (pmx-loader.ts:857-939), in place */
void orc_finalize_skinning(uint16_t* joints, uint8_t* weights, uint32_t V, uint32_t boneCount) {
  for (uint32_t v = 0; v < V; ++v) {
    uint16_t* j = joints + (size_t)v * 4;
    uint8_t* w = weights + (size_t)v * 4;
    int validSum = 0, validCount = 0;
    for (int k = 0; k < 4; ++k) {
      if (j[k] >= boneCount) {
        w[k] = 0;
        j[k] = boneCount > 0 ? (uint16_t)(boneCount - 1) : 0;
      } else {
        validSum += w[k];
        validCount++;
      }
    }
#define OK(k) (j[k] < boneCount)
    if (validSum == 0 || validCount == 0) {
      w[0] = 255; w[1] = w[2] = w[3] = 0;
      j[0] = j[1] = j[2] = j[3] = 0;
    } else if (validSum != 255) {
      const double scale = 255.0 / validSum;
      int accum = 0;
      for (int k = 0; k < 3; ++k) {
        if (OK(k)) {
          double r = js_round(w[k] * scale);
          int vv = (int)(r < 0 ? 0 : (r > 255 ? 255 : r));
          w[k] = (uint8_t)vv;
          accum += vv;
        } else {
          w[k] = 0;
        }
      }
      if (OK(3)) {
        int r = 255 - accum;
        w[3] = (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
      } else {
        w[3] = 0;
        if (accum < 255)
          for (int k = 2; k >= 0; --k)
            if (OK(k) && w[k] > 0) {
              int r = w[k] + (255 - accum);
              w[k] = (uint8_t)(r > 255 ? 255 : r);
              break;
            }
      }
      const int fs = w[0] + w[1] + w[2] + w[3];
      if (fs != 255) {
        const int diff = 255 - fs;
        int mi = 0, mw = w[0];
        for (int k = 1; k < 4; ++k)
          if (w[k] > mw && OK(k)) { mw = w[k]; mi = k; }
        if (OK(mi)) {
          int r = w[mi] + diff;
          w[mi] = (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
        }
      }
    }
#undef OK
  }
}

/* computeInverseBind (pmx-loader.ts:791-824): bind world = f32 chain of parent-relative translations,
 * invBind = T(-world.t).  bindTranslation is the f64 triple the loader stores (pmx-loader.ts:423). */
void orc_inverse_bind(const int32_t* parent, const double* bindTranslation, uint32_t B, float* inv16) {
  float* wt = (float*)calloc((size_t)B * 3, sizeof(float));
  uint8_t* done = (uint8_t*)calloc(B, 1);
  uint32_t* stack = (uint32_t*)malloc((size_t)(B + 1) * sizeof(uint32_t));
  for (uint32_t i = 0; i < B; ++i) {
    uint32_t sp = 0, cur = i;
    while (!done[cur]) {
      stack[sp++] = cur;
      const int32_t p = parent[cur];
      if (p < 0 || (uint32_t)p >= B || sp > B) break;
      cur = (uint32_t)p;
    }
    while (sp) {
      const uint32_t b = stack[--sp];
      if (done[b]) continue;
      const int32_t p = parent[b];
      for (int k = 0; k < 3; ++k) {
        const float lt = (float)bindTranslation[(size_t)b * 3 + k];     /* translateInPlace: 0 + t -> f32 */
        wt[(size_t)b * 3 + k] = (p >= 0 && (uint32_t)p < B) ? (float)((double)lt + (double)wt[(size_t)p * 3 + k]) : lt;
      }
      done[b] = 1;
    }
  }
  memset(inv16, 0, (size_t)B * 16 * sizeof(float));
  for (uint32_t b = 0; b < B; ++b) {
    float* m = inv16 + (size_t)b * 16;
    m[0] = m[5] = m[10] = m[15] = 1.0f;
    for (int k = 0; k < 3; ++k) m[12 + k] = (float)(0.0 + -(double)wt[(size_t)b * 3 + k]);
  }
  free(wt); free(done); free(stack);
}

/* ---- multi-instance driver (CPU baseline): K instances x V vertices over `nthreads` pthreads ------------- */
typedef struct {
  const float* vtx8; const uint16_t* joints; const uint8_t* weights; uint32_t V, B, K;
  const float* world; /* P x B x 16 */ const float* invBind; const uint32_t* inst2pal;
  const uint32_t* vmStart; const uint32_t* vmMorph; const float* vmDelta; const float* morphW; uint32_t M;
  const int32_t* sdefOf; const float* sdefVec9;
  float* out; size_t instStrideF, nrmOffF;
  uint32_t k0, k1;
  int32_t ringSlot; /* >= 0: every instance of this worker is written to this slot (timing runs) */
} orc_job;

static void* orc_worker(void* arg) {
  orc_job* j = (orc_job*)arg;
  float* skin = (float*)malloc((size_t)j->B * 16 * sizeof(float));
  for (uint32_t k = j->k0; k < j->k1; ++k) {
    const uint32_t p = j->inst2pal ? j->inst2pal[k] : k;
    orc_skin_matrices_f32(j->world + (size_t)p * j->B * 16, j->invBind, j->B, skin);
    float* pos = j->out + (size_t)(j->ringSlot >= 0 ? (uint32_t)j->ringSlot : k) * j->instStrideF;
    orc_deform_range_f32(j->vtx8, j->joints, j->weights, 0, j->V, skin, j->vmStart, j->vmMorph, j->vmDelta,
                         j->morphW ? j->morphW + (size_t)k * j->M : NULL, j->sdefOf, j->sdefVec9, pos, pos + j->nrmOffF);
  }
  free(skin);
  return NULL;
}

/* out layout identical to the product's: per instance pos plane then normal plane.
 * ring != 0: `out` holds only `nthreads` instance slots, worker t overwrites slot t (CPU-baseline timing without
 * first-touch page faults on a K-instance buffer). */
void orc_deform_instances(const float* vtx8, const uint16_t* joints, const uint8_t* weights, uint32_t V, uint32_t B,
                          const float* world, const float* invBind, const uint32_t* inst2pal, uint32_t K,
                          const uint32_t* vmStart, const uint32_t* vmMorph, const float* vmDelta, const float* morphW, uint32_t M,
                          const int32_t* sdefOf, const float* sdefVec9,
                          float* out, size_t instStrideF, size_t nrmOffF, uint32_t nthreads, int32_t ring) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > K) nthreads = K;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
  orc_job* jobs = (orc_job*)malloc(sizeof(orc_job) * nthreads);
  for (uint32_t t = 0; t < nthreads; ++t) {
    orc_job j = {vtx8, joints, weights, V, B, K, world, invBind, inst2pal, vmStart, vmMorph, vmDelta, morphW, M,
                 sdefOf, sdefVec9, out, instStrideF, nrmOffF, (uint32_t)((uint64_t)K * t / nthreads),
                 (uint32_t)((uint64_t)K * (t + 1) / nthreads), ring ? (int32_t)t : -1};
    jobs[t] = j;
    if (nthreads == 1) orc_worker(&jobs[t]);
    else pthread_create(&th[t], NULL, orc_worker, &jobs[t]);
  }
  if (nthreads > 1)
    for (uint32_t t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  free(th); free(jobs);
}
