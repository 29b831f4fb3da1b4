// mesh_tables.h — load-time table builders of the deform path that need no device: the bank-aware palette permutation,
// the per-warp morph rows, the cost-balanced chunk table.  Pure host C++ (no CUDA types): used by rebuild_tables / rz_deform
// (rze_b200.cu) and, through rz_plan_morph_rows / rz_plan_chunks, by the CPU test-suite.  See DESIGN.md sections 3 and 4.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <unordered_map>
#include <vector>

namespace rz {

struct F4 { float x, y, z, w; };   // layout of CUDA's float4

// ---- bank-aware palette permutation -------------------------------------------------------------------------------
// A warp-wide LDS.128 costs max(2, distinct chunks / 4, 2 x chunks per 16-byte bank group) cycles on sm_100
// (profiles/r01_ubench_lds128.txt).  Rows are 48 B, so the bank group of row r of bone b is (3*pos(b)+r) mod 8:
// bones gathered by the same warp instruction should sit at positions that differ mod 8.  Greedy colouring of the
// bone co-occurrence graph into 8 classes, then class c occupies palette rows c, c+8, c+16, ...
// gatherJ [Vp][4]: the bones each processing lane gathers (lane_plan.h).  bonePos[b] = palette row of bone b (a permutation).
inline void plan_palette_rows(const uint16_t* gatherJ, uint32_t Vp, uint32_t B, int colorMode, int layoutMode, std::vector<uint32_t>& bonePos) {
  bonePos.resize(B);
  for (uint32_t b = 0; b < B; ++b) bonePos[b] = b;
  if (colorMode && B >= 16) {
    std::unordered_map<uint64_t, uint32_t> pairW;
    std::vector<uint32_t> seen;
    for (uint32_t w0 = 0; w0 < Vp; w0 += 32) {
      for (uint32_t k = 0; k < 4; ++k) {
        seen.clear();
        for (uint32_t l = 0; l < 32; ++l) {
          const uint32_t b = gatherJ[(size_t)(w0 + l) * 4 + k];
          if (std::find(seen.begin(), seen.end(), b) == seen.end()) seen.push_back(b);
        }
        for (size_t a = 0; a < seen.size(); ++a)
          for (size_t b2 = a + 1; b2 < seen.size(); ++b2) {
            const uint32_t lo = std::min(seen[a], seen[b2]), hi = std::max(seen[a], seen[b2]);
            pairW[((uint64_t)lo << 32) | hi] += 1;
          }
      }
    }
    std::vector<std::vector<std::pair<uint32_t, uint32_t>>> adj(B);
    std::vector<uint64_t> tot(B, 0);
    for (const auto& kv : pairW) {
      const uint32_t lo = (uint32_t)(kv.first >> 32), hi = (uint32_t)kv.first;
      adj[lo].push_back({hi, kv.second});
      adj[hi].push_back({lo, kv.second});
      tot[lo] += kv.second;
      tot[hi] += kv.second;
    }
    std::vector<uint32_t> order(B);
    for (uint32_t b = 0; b < B; ++b) order[b] = b;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return tot[a] > tot[b]; });
    uint32_t cap[8], cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (uint32_t cl = 0; cl < 8; ++cl) cap[cl] = (B - cl + 7) / 8;
    std::vector<int> cls(B, -1);
    for (uint32_t b : order) {
      uint64_t cost[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (const auto& e : adj[b]) if (cls[e.first] >= 0) cost[cls[e.first]] += e.second;
      int best = -1;
      for (int cl = 0; cl < 8; ++cl) {
        if (cnt[cl] >= cap[cl]) continue;
        if (best < 0 || cost[cl] < cost[best] || (cost[cl] == cost[best] && cnt[cl] < cnt[best])) best = cl;
      }
      cls[b] = best;
      cnt[best]++;
    }
    uint32_t next[8];
    for (uint32_t cl = 0; cl < 8; ++cl) next[cl] = cl;
    for (uint32_t b = 0; b < B; ++b) { bonePos[b] = next[cls[b]]; next[cls[b]] += 8; }
    if (layoutMode == 1) {
      // [3][B] layout: 8 consecutive palette rows share one 128-byte line per chunk, so bones gathered together should be
      // NEIGHBOURS: greedy chaining -- start a group of 8 with the heaviest unplaced bone, then keep appending the unplaced
      // bone with the largest co-occurrence weight to the bones already in the group.
      std::vector<char> placed(B, 0);
      std::vector<uint64_t> gain(B, 0);
      uint32_t pos = 0;
      size_t oi = 0;
      while (pos < B) {
        while (oi < B && placed[order[oi]]) ++oi;
        uint32_t seed = order[oi];
        std::vector<uint32_t> touched;
        uint32_t cur = seed;
        for (uint32_t g = 0; g < 8 && pos < B; ++g) {
          placed[cur] = 1;
          bonePos[cur] = pos++;
          for (const auto& e : adj[cur]) if (!placed[e.first]) { if (!gain[e.first]) touched.push_back(e.first); gain[e.first] += e.second; }
          uint32_t best = ~0u;
          for (uint32_t t : touched) if (!placed[t] && (best == ~0u || gain[t] > gain[best])) best = t;
          if (best == ~0u) {               // nothing related left: take the next heaviest bone
            size_t oj = oi;
            while (oj < B && placed[order[oj]]) ++oj;
            if (oj >= B) break;
            best = order[oj];
          }
          cur = best;
        }
        for (uint32_t t : touched) gain[t] = 0;
      }
    }
  }
}

// ---- morph rows ---------------------------------------------------------------------------------------------------
struct MorphRows {
  std::vector<uint32_t> first, depth;    // [Vp/32] per warp: first row entry and number of rows
  std::vector<uint8_t> morphMajor;       // [Vp/32] 1: row = one morph for the whole warp, 0: compact per-lane lists
  std::vector<F4> rows;                  // entry u of lane l of warp w at first[w] + u*32 + l = (dx, dy, dz, morph id bits)
};

// Vertex-major view of the caller's morph-major CSR (PMX morph order is kept inside every vertex):
// vertex v owns entries start[v] .. start[v]+count[v]-1 of `ents` = (delta, morph id bits).  storedOf[caller vertex] = stored index.
inline void morphs_by_vertex(uint32_t V, uint32_t M, const uint32_t* moff, const uint32_t* mvert, const float* mdelta, const uint32_t* storedOf,
                             std::vector<uint32_t>& count, std::vector<uint32_t>& start, std::vector<F4>& ents) {
  count.assign(V, 0);
  start.assign((size_t)V + 1, 0);
  const uint32_t nnz = M ? moff[M] : 0;
  for (uint32_t e = 0; e < nnz; ++e) count[storedOf ? storedOf[mvert[e]] : mvert[e]]++;
  for (uint32_t v = 0; v < V; ++v) start[v + 1] = start[v] + count[v];
  ents.assign(std::max<uint32_t>(nnz, 1), F4{0.f, 0.f, 0.f, 0.f});
  std::vector<uint32_t> fill(start.begin(), start.end() - 1);
  for (uint32_t m = 0; m < M; ++m)
    for (uint32_t e = moff[m]; e < moff[m + 1]; ++e) {
      const uint32_t v = storedOf ? storedOf[mvert[e]] : mvert[e];
      F4 r{mdelta[(size_t)e * 3], mdelta[(size_t)e * 3 + 1], mdelta[(size_t)e * 3 + 2], 0.f};
      memcpy(&r.w, &m, 4);
      ents[fill[v]++] = r;
    }
}

// ---- morph entries, lane-interleaved per warp (ELL): entry u of lane l sits at first + u*32 + l.  One LDG.128 per depth
// step is then a single coalesced 512-byte request for the warp instead of 32 scattered 16-byte ones (the L1 tag stage
// serialises those: measured +0.16 ms on config 3), and the loop bound is warp-uniform.  Within a lane the entries keep
// PMX morph order.  Two row formats, chosen per warp, identical to the kernel:
//   MORPH-MAJOR (default): row u = one morph for the whole warp, delta 0 on lanes that morph does not touch.  PMX morphs
//     are spatially coherent (the 32 vertices of a warp see the same morphs), so this costs no extra rows, and every lane
//     of a row looks up the SAME weight: a shared-memory broadcast instead of a 3-4-way bank conflict per lookup
//     (ncu on config 3: 39 % of the shared-memory wavefronts were conflicts).  fma(w, 0, p) == p: results unchanged.
//   COMPACT (fallback when the union of morphs is > 1.5x the deepest vertex): entry u = the lane's own u-th morph,
//     padded with (delta 0, morph 0).
// procVertex [Vp]: stored vertex evaluated by processing lane p (~0u: padding), from the lane plan.
inline void build_morph_rows(const uint32_t* procVertex, uint32_t Vp, const std::vector<uint32_t>& mcount, const std::vector<uint32_t>& mstart,
                             const std::vector<F4>& ments, MorphRows& out) {
  out.first.assign(Vp / 32, 0); out.depth.assign(Vp / 32, 0); out.morphMajor.assign(Vp / 32, 0);
  std::vector<F4>& mell = out.rows;
  mell.clear();
  for (uint32_t w0 = 0; w0 < Vp; w0 += 32) {
    uint32_t deep = 0;
    bool dup = false;
    std::vector<uint32_t> ids;
    for (uint32_t l = 0; l < 32; ++l) {
      const uint32_t v = procVertex[w0 + l];
      if (v == ~0u) continue;
      deep = std::max(deep, mcount[v]);
      for (uint32_t u = 0; u < mcount[v]; ++u) {
        uint32_t m;
        memcpy(&m, &ments[mstart[v] + u].w, 4);
        if (u && m == ids.back()) dup = true;      // one morph lists this vertex twice: keep both entries (compact rows)
        ids.push_back(m);
      }
    }
    const uint32_t first = (uint32_t)mell.size();
    if (!deep) { out.first[w0 / 32] = first; out.depth[w0 / 32] = 0; continue; }
    std::sort(ids.begin(), ids.end());
    ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
    const bool morphMajor = !dup && ids.size() <= (size_t)deep + deep / 2 + 1;
    const uint32_t rows = morphMajor ? (uint32_t)ids.size() : deep;
    out.first[w0 / 32] = first; out.depth[w0 / 32] = rows; out.morphMajor[w0 / 32] = morphMajor ? 1 : 0;
    mell.resize((size_t)first + (size_t)rows * 32, F4{0.f, 0.f, 0.f, 0.f});
    if (morphMajor)
      for (uint32_t r = 0; r < rows; ++r) {
        float idBits;
        memcpy(&idBits, &ids[r], 4);
        for (uint32_t l = 0; l < 32; ++l) mell[(size_t)first + (size_t)r * 32 + l].w = idBits;
      }
    for (uint32_t l = 0; l < 32; ++l) {
      const uint32_t v = procVertex[w0 + l];
      if (v == ~0u) continue;
      uint32_t r = 0;
      for (uint32_t u = 0; u < mcount[v]; ++u) {
        const F4 e = ments[mstart[v] + u];
        if (morphMajor) {
          uint32_t m;
          memcpy(&m, &e.w, 4);
          while (ids[r] != m) ++r;                                  // both ascending
          mell[(size_t)first + (size_t)r * 32 + l].x = e.x;
          mell[(size_t)first + (size_t)r * 32 + l].y = e.y;
          mell[(size_t)first + (size_t)r * 32 + l].z = e.z;
        } else {
          mell[(size_t)first + (size_t)u * 32 + l] = e;
        }
      }
    }
  }
  if (mell.empty()) mell.push_back(F4{0.f, 0.f, 0.f, 0.f});
}

// ---- SDEF tables ----------------------------------------------------------------------------------------------------
// One 48-byte record per SDEF vertex and, per warp, the descriptor list the kernel's dense phase walks (deform_kernel.cuh):
//   record  = (C.xyz, c0.x) (c0.yz, c1.xy) (c1.z, w0, w1, palette rows j0 | j1 << 16)
//             rw = w0*R0 + w1*R1, r0 = C + R0 - rw, r1 = C + R1 - rw, c0 = (C + r0)/2, c1 = (C + r1)/2      
//             w0, w1 = the reference's shader-time normalisation of the two unorm8 weights (engine.ts:255-258), f32
//   desc[w*32 + l] = table index | output slot << 24 of the l-th SDEF vertex among warp w's 32 outputs, ~0u beyond.
// sdefOf [V]: index into vec9 (the caller's C, R0, R1) of stored vertex v or -1; JT / WT: joints and unorm8 weights in stored
// vertex order; bonePos: palette row of every bone.  Returns false when there are more than 2^24 SDEF vertices.
struct SdefTables {
  std::vector<F4> tab;              // 3 per SDEF vertex
  std::vector<uint32_t> desc;       // [Vp]
  uint32_t active = 0;
};
inline bool build_sdef_tables(const int32_t* sdefOf, const float* vec9, const uint16_t* JT, const uint8_t* WT, const uint32_t* bonePos,
                              const uint32_t* procVertex, const uint32_t* procSlot, uint32_t V, uint32_t Vp, SdefTables& out) {
  out.tab.clear();
  out.desc.assign(Vp, ~0u);
  out.active = 0;
  std::vector<int32_t> recOf(V, -1);
  for (uint32_t v = 0; v < V; ++v) {
    if (sdefOf[v] < 0) continue;
    const uint8_t* w = &WT[(size_t)v * 4];
    const float* s = &vec9[(size_t)sdefOf[v] * 9];
    float w0 = (float)w[0] / 255.0f, w1 = (float)w[1] / 255.0f;        // weights exactly as the kernel derives them
    const float ws = w0 + w1 + 0.f + 0.f;
    if (ws > 0.0001f) { const float inv = 1.0f / ws; w0 *= inv; w1 *= inv; } else { w0 = 1.f; w1 = 0.f; }
    float C[3] = {s[0], s[1], s[2]}, c0[3], c1[3];
    for (int k = 0; k < 3; ++k) {
      const float R0 = s[3 + k], R1 = s[6 + k];
      const float rw = w0 * R0 + w1 * R1;
      const float r0 = C[k] + R0 - rw, r1 = C[k] + R1 - rw;
      c0[k] = (C[k] + r0) * 0.5f;
      c1[k] = (C[k] + r1) * 0.5f;
    }
    const uint32_t rows = bonePos[JT[(size_t)v * 4]] | (bonePos[JT[(size_t)v * 4 + 1]] << 16);
    float rowsF;
    memcpy(&rowsF, &rows, 4);
    recOf[v] = (int32_t)(out.tab.size() / 3);
    out.tab.push_back(F4{C[0], C[1], C[2], c0[0]});
    out.tab.push_back(F4{c0[1], c0[2], c1[0], c1[1]});
    out.tab.push_back(F4{c1[2], w0, w1, rowsF});
    out.active++;
  }
  if (out.tab.size() / 3 >= (1u << 24)) return false;
  for (uint32_t w0 = 0; w0 < Vp; w0 += 32) {
    uint32_t n = 0;
    for (uint32_t l = 0; l < 32; ++l) {
      const uint32_t v = procVertex[w0 + l];
      if (v == ~0u || recOf[v] < 0) continue;
      out.desc[w0 + n++] = (uint32_t)recOf[v] | (procSlot[w0 + l] << 24);
    }
  }
  return true;
}

// ---- cost-balanced chunks -------------------------------------------------------------------------------------------
// Morph passes are latency chains (record -> rows -> weights: one L2 round trip per `batch` rows beyond the `prefetch`ed
// ones), several times longer than a plain pass, and PMX morphs cluster on the face: uniform chunks would leave a few very
// long work items.  Boundaries are placed so that every item carries the same estimated time: a pass costs 1 (+ 1/4 when it
// has morph rows) + tripCost per round trip, in units of a plain pass.  tileDepth [nTiles]: deepest morph row list of the
// tile's warps.  tab: nChunks+1 tile boundaries, each a multiple of tilesPerPass (the last one = nTiles).
inline void build_chunk_table(const uint32_t* tileDepth, uint32_t nTiles, uint32_t tilesPerPass, uint32_t nChunksTarget, float tripCost,
                              uint32_t prefetch, uint32_t batch, std::vector<uint32_t>& tab) {
  const uint32_t nPasses = (nTiles + tilesPerPass - 1) / tilesPerPass;
  nChunksTarget = std::max(1u, std::min(nChunksTarget, nPasses));
  std::vector<float> cost(nPasses);
  double total = 0;
  for (uint32_t p = 0; p < nPasses; ++p) {
    uint32_t deep = 0;
    for (uint32_t t = p * tilesPerPass; t < std::min(nTiles, (p + 1) * tilesPerPass); ++t) deep = std::max(deep, tileDepth[t]);
    const uint32_t trips = deep > prefetch ? (deep - prefetch + batch - 1) / batch : 0;
    cost[p] = 1.0f + (deep ? 0.25f : 0.f) + tripCost * (float)trips;
    total += cost[p];
  }
  tab.clear();
  tab.push_back(0);
  double acc = 0;
  for (uint32_t p = 0; p < nPasses; ++p) {
    acc += cost[p];
    if (acc >= total * (double)tab.size() / (double)nChunksTarget && p + 1 < nPasses) tab.push_back((p + 1) * tilesPerPass);
  }
  tab.push_back(nTiles);
}

}  // namespace rz
