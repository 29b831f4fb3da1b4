// lane_plan2.h — load-time planning for TWO vertices per lane: the table deform2_kernel (deform2_kernel.cuh) runs on.
// Pure host C++, no CUDA; also reachable through the device-free diagnostic entry rz_plan_lanes2 (CPU tests).
//
// Why: the deform kernel is bound by the shared-memory pipe, and 20 of its ~32 shared-memory cycles per warp-instance
// are palette gathers -- one gather instruction per influence slot per warp of 32 vertices.  If a lane evaluates two
// vertices that share ONE list of <= 4 palette rows (each vertex with its own weight per row, zero where the bone is
// not its own), a warp covers 64 vertices with the same gather instructions.
//
// Plan, per window of 64 consecutive vertices (a group's outputs stay contiguous for the TMA bulk stores):
//   level 1  bottleneck perfect matching of the window's vertices (Edmonds' blossom algorithm, thresholds on |A u B|):
//            which two vertices share a lane so that the largest union of bone sets -- the number of gather
//            instructions the warp needs -- is as small as possible, and <= 4 (the record has four row slots);
//   level 2  the 32 lanes, each now carrying a union set, go through the existing pair packer (lane_plan.h) so that
//            aligned lane pairs read the same rows (the shared-memory broadcast fast path).
// A window that cannot be matched with unions <= 4, or that holds a vertex listing one bone twice (its two products
// must stay separate to remain bit-faithful), becomes two ordinary one-vertex-per-lane groups of 32 vertices.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "lane_plan.h"

namespace rz {

// fn(begin, end) over [0, n) on a few host threads (load-time planning only; every range writes its own part of the plan, so
// the result does not depend on the thread count; RZ_PLAN_THREADS=1 keeps it on the calling thread)
template <class F>
inline void plan_parallel_for(uint32_t n, uint32_t minPerThread, F fn) {
  unsigned nt = std::thread::hardware_concurrency();
  if (const char* e = getenv("RZ_PLAN_THREADS")) nt = (unsigned)std::max(1, atoi(e));
  nt = std::max(1u, std::min(std::min(nt, 16u), n / std::max(1u, minPerThread)));
  if (nt <= 1) { fn(0u, n); return; }
  std::vector<std::thread> th;
  const uint32_t per = (n + nt - 1) / nt;
  for (unsigned t = 0; t < nt; ++t) {
    const uint32_t b = t * per, e2 = std::min(n, b + per);
    if (b >= e2) break;
    th.emplace_back([=]() { fn(b, e2); });
  }
  for (auto& t : th) t.join();
}

struct LanePlan2 {
  uint32_t V = 0, nGroups = 0;
  std::vector<uint32_t> groupFirst;   // [nGroups] first output vertex of the group
  std::vector<uint32_t> groupCount;   // [nGroups] vertices it covers (<= 64 paired, <= 32 single)
  std::vector<uint8_t> groupPaired;   // [nGroups] 1: two vertices per lane
  // per lane p = group*32 + lane
  std::vector<uint32_t> vertA, vertB; // vertex ids (~0u: none); output slot inside the group = vertex - groupFirst
  std::vector<uint16_t> gatherJ;      // [4] bone whose palette row slot s gathers (never out of range)
  std::vector<float> wA, wB;          // [4] shader-normalised weight of slot s for vertex A / B (0 where not its bone)
  uint64_t slots = 0, fastSlots = 0;  // gather instructions (group x slot) of packed groups / of those on the fast path
  uint32_t pairedWindows = 0, fallbackWindows = 0;
  // staging slots: where in the group's 64-vertex staging area the lane writes its vertex A / B.  A real vertex sits at
  // (vertex - groupFirst); a lane without a vertex on a side gets one of the slots no vertex owns (never drained).  Sides are
  // chosen so that the 32 A-slots of a group are distinct modulo 32, likewise the B-slots: 12-byte staging records then hit
  // 32 distinct banks per store instruction (3 * slot mod 32 is a bijection on a residue system).
  std::vector<uint8_t> slotA, slotB;
  std::vector<uint8_t> laneN;         // slots the lane uses (highest slot with a non-zero weight on either side + 1, >= 1)
};

inline void plan_lanes2(const uint16_t* JT, const uint8_t* WT, uint32_t V, uint32_t B, LanePlan2& out) {
  out = LanePlan2();
  out.V = V;
  struct VSet { int n; uint16_t b[4]; float w[4]; bool dup; };
  std::vector<VSet> S(V);
  for (uint32_t v = 0; v < V; ++v) {
    // weights exactly as the reference's vertex shader derives them (engine.ts:255-258), like lane_plan.h shader_weights
    const uint8_t* w8 = &WT[(size_t)v * 4];
    float w[4];
    for (int k = 0; k < 4; ++k) w[k] = (float)w8[k] / 255.0f;
    const float wsum = w[0] + w[1] + w[2] + w[3];
    if (wsum > 0.0001f) { const float inv = 1.0f / wsum; for (int k = 0; k < 4; ++k) w[k] = w[k] * inv; }
    else { w[0] = 1.f; w[1] = w[2] = w[3] = 0.f; }
    VSet& s = S[v];
    s.n = 0; s.dup = false;
    for (int k = 0; k < 4; ++k) {
      if (w[k] == 0.f) continue;
      const uint16_t bone = JT[(size_t)v * 4 + k];
      for (int t = 0; t < s.n; ++t) if (s.b[t] == bone) s.dup = true;
      if (s.n < 4) { s.b[s.n] = bone; s.w[s.n] = w[k]; }
      ++s.n;
    }
  }
  auto union_size = [&](const VSet& a, const VSet& b2) {
    int u = a.n;
    for (int j = 0; j < b2.n; ++j) { bool in = false; for (int i = 0; i < a.n; ++i) in |= a.b[i] == b2.b[j]; u += !in; }
    return u;
  };

  // ---- level 1: who shares a lane
  struct LaneV { uint32_t a, b; };                        // ~0u: none
  std::vector<LaneV> lanes;                              // 32 per group
  auto push_group = [&](uint32_t first, uint32_t count, bool paired) {
    out.groupFirst.push_back(first); out.groupCount.push_back(count); out.groupPaired.push_back(paired ? 1 : 0);
  };
  for (uint32_t w0 = 0; w0 < V; w0 += 64) {
    const uint32_t n = std::min(64u, V - w0);
    bool ok = true;
    int nmin = 1;
    for (uint32_t i = 0; i < n; ++i) { ok &= !S[w0 + i].dup; nmin = std::max(nmin, S[w0 + i].n); }
    BlossomMatcher<64> bm;
    const int nodes = (int)((n + 1) & ~1u);               // an odd window gets one null node that pairs with anybody
    bm.n = nodes;
    int thr = nmin;
    if (ok) {
      // nodes in key order (similar sets become neighbours: the matcher's greedy start pairs them first)
      std::vector<uint32_t> ord(n);
      for (uint32_t i = 0; i < n; ++i) ord[i] = w0 + i;
      auto key = [&](uint32_t v) {
        uint16_t sb[4] = {0xFFFF, 0xFFFF, 0xFFFF, 0xFFFF};
        for (int t = 0; t < S[v].n && t < 4; ++t) sb[t] = S[v].b[t];
        std::sort(sb, sb + 4);
        return ((uint64_t)sb[0] << 48) | ((uint64_t)sb[1] << 32) | ((uint64_t)sb[2] << 16) | sb[3];
      };
      std::stable_sort(ord.begin(), ord.end(), [&](uint32_t x, uint32_t y) { return key(x) < key(y); });
      std::vector<int> un((size_t)nodes * nodes, 0);
      for (int x = 0; x < nodes; ++x)
        for (int y = 0; y < nodes; ++y)
          un[(size_t)x * nodes + y] = (x == y) ? 99 : ((uint32_t)x >= n || (uint32_t)y >= n) ? ((uint32_t)x >= n ? ((uint32_t)y >= n ? 99 : S[ord[y]].n) : S[ord[x]].n)
                                                                                            : union_size(S[ord[x]], S[ord[y]]);
      bool found = false;
      for (; thr <= 4 && !found; ++thr) {
        for (int x = 0; x < nodes; ++x) for (int y = 0; y < nodes; ++y) bm.adj[x][y] = un[(size_t)x * nodes + y] <= thr;
        found = bm.solve() == nodes / 2;
      }
      ok = found;
      if (ok) {
        push_group(w0, n, true);
        out.pairedWindows++;
        size_t made = 0;
        for (int x = 0; x < nodes; ++x) {
          const int y = bm.match[x];
          if (y < x) continue;
          LaneV l;
          l.a = (uint32_t)x < n ? ord[x] : ~0u;
          l.b = (uint32_t)y < n ? ord[y] : ~0u;
          if (l.a == ~0u) std::swap(l.a, l.b);
          lanes.push_back(l);
          ++made;
        }
        for (; made < 32; ++made) lanes.push_back(LaneV{~0u, ~0u});
      }
    }
    if (!ok) {
      out.fallbackWindows++;
      for (uint32_t h = 0; h < n; h += 32) {
        const uint32_t cnt = std::min(32u, n - h);
        push_group(w0 + h, cnt, false);
        for (uint32_t l = 0; l < 32; ++l) lanes.push_back(LaneV{l < cnt ? w0 + h + l : ~0u, ~0u});
      }
    }
  }
  out.nGroups = (uint32_t)out.groupFirst.size();

  // ---- level 2: the existing pair packer on a virtual table, one virtual vertex per lane carrying the lane's bone set
  // (any non-zero weights: the packer only looks at which weights are non-zero; a lane without vertices copies the set
  // of the group's first lane so that it adds no row the warp does not gather anyway)
  const uint32_t VL = out.nGroups * 32;
  std::vector<uint16_t> vj((size_t)VL * 4, 0);
  std::vector<uint8_t> vw((size_t)VL * 4, 0);
  auto lane_set = [&](const LaneV& l, uint16_t bones[8]) {
    int n = 0;
    for (uint32_t v : {l.a, l.b}) {
      if (v == ~0u) continue;
      for (int t = 0; t < std::min(S[v].n, 4); ++t) {
        bool in = false;
        for (int i = 0; i < n; ++i) in |= bones[i] == S[v].b[t];
        if (!in && n < 8) bones[n++] = S[v].b[t];
      }
    }
    return n;
  };
  for (uint32_t g = 0; g < out.nGroups; ++g)
    for (uint32_t l = 0; l < 32; ++l) {
      const uint32_t p = g * 32 + l;
      LaneV src = lanes[p];
      if (src.a == ~0u) src = lanes[(size_t)g * 32];
      uint16_t bones[8];
      int n = lane_set(src, bones);
      // a fallback lane whose single vertex lists a bone twice keeps the caller's four slots verbatim
      if (!out.groupPaired[g] && src.a != ~0u && S[src.a].dup) {
        for (int k = 0; k < 4; ++k) { vj[(size_t)p * 4 + k] = JT[(size_t)src.a * 4 + k]; vw[(size_t)p * 4 + k] = WT[(size_t)src.a * 4 + k]; }
        continue;
      }
      n = std::min(n, 4);
      for (int k = 0; k < n; ++k) { vj[(size_t)p * 4 + k] = bones[k]; vw[(size_t)p * 4 + k] = 60; }
      if (n == 0) { vj[(size_t)p * 4] = 0; vw[(size_t)p * 4] = 255; }
    }
  LanePlan vp;
  plan_lanes(vj.data(), vw.data(), nullptr, VL, B, 32, 2, vp);
  out.slots = vp.totalSlots;
  out.fastSlots = vp.fastSlots;

  // ---- real vertices and weights into the planned lanes
  out.vertA.assign(VL, ~0u); out.vertB.assign(VL, ~0u);
  out.gatherJ.assign((size_t)VL * 4, 0);
  out.wA.assign((size_t)VL * 4, 0.f); out.wB.assign((size_t)VL * 4, 0.f);
  for (uint32_t p = 0; p < VL; ++p) {
    const uint32_t src = vp.procVertex[p];                // virtual vertex (= lane of level 1) evaluated by lane p
    for (int k = 0; k < 4; ++k) out.gatherJ[(size_t)p * 4 + k] = vp.gatherJ[(size_t)p * 4 + k];
    if (src == ~0u) continue;
    const LaneV l = lanes[src];
    out.vertA[p] = l.a; out.vertB[p] = l.b;
    const bool verbatim = !out.groupPaired[src / 32] && l.a != ~0u && S[l.a].dup;
    for (int side = 0; side < 2; ++side) {
      const uint32_t v = side ? l.b : l.a;
      if (v == ~0u) continue;
      float* w = side ? &out.wB[(size_t)p * 4] : &out.wA[(size_t)p * 4];
      if (verbatim) {                                     // the packer left this warp in the caller's slot order
        for (int k = 0; k < 4; ++k) w[k] = vp.devW[(size_t)p * 4 + k];
        continue;
      }
      for (int t = 0; t < std::min(S[v].n, 4); ++t)
        for (int k = 0; k < 4; ++k)
          if (vp.devW[(size_t)p * 4 + k] != 0.f && vp.gatherJ[(size_t)p * 4 + k] == S[v].b[t]) { w[k] = S[v].w[t]; break; }
    }
  }

  // ---- staging slots and sides
  out.slotA.assign(VL, 0); out.slotB.assign(VL, 0); out.laneN.assign(VL, 1);
  // (the quarter-warp hill-climb below is 60 % of the planning time: groups are independent, a few host threads share them)
  plan_parallel_for(out.nGroups, 64, [&](uint32_t gBegin, uint32_t gEnd) {
  for (uint32_t g = gBegin; g < gEnd; ++g) {
    const uint32_t first = out.groupFirst[g], base = g * 32;
    int owner[64];                                          // slot -> lane * 2 + side, -1: free
    for (int s2 = 0; s2 < 64; ++s2) owner[s2] = -1;
    int slotOf[32][2];
    for (uint32_t l = 0; l < 32; ++l)
      for (int side = 0; side < 2; ++side) {
        const uint32_t v = side ? out.vertB[base + l] : out.vertA[base + l];
        slotOf[l][side] = v == ~0u ? -1 : (int)(v - first);
        if (v != ~0u) owner[v - first] = (int)l * 2 + side;
      }
    if (!out.groupPaired[g]) {
      // one vertex per lane, all on side A (the kernel skips side B of such a group): real vertices own slots 0..count-1
      // (< 32), empty lanes take the remaining low slots, the unused B sides the high half -- distinct modulo 32 as they are
      int lowFree = 0;
      for (uint32_t l = 0; l < 32; ++l) {
        const uint32_t p = base + l;
        if (slotOf[l][0] < 0) {
          while (owner[lowFree] >= 0) ++lowFree;
          slotOf[l][0] = lowFree;
          owner[lowFree] = (int)l * 2;
        }
        out.slotA[p] = (uint8_t)slotOf[l][0];
        out.slotB[p] = (uint8_t)(32 + l);
        int n = 1;
        for (int k = 0; k < 4; ++k) if (out.wA[(size_t)p * 4 + k] != 0.f) n = k + 1;
        out.laneN[p] = (uint8_t)n;
      }
      continue;
    }
    int nextFree = 0;
    for (uint32_t l = 0; l < 32; ++l)
      for (int side = 0; side < 2; ++side) {
        if (slotOf[l][side] >= 0) continue;
        while (owner[nextFree] >= 0) ++nextFree;            // 64 slots, 32 lanes x 2: there is always one left
        slotOf[l][side] = nextFree;
        owner[nextFree] = (int)l * 2 + side;
      }
    // Every slot now has exactly one lane partner (the lane's other slot) and one residue partner (slot ^ 32): the union
    // of the two pairings is a set of even cycles that alternate between them, so colouring alternately along each cycle
    // gives every lane one slot of each colour AND every residue class one slot of each colour.
    int colour[64];
    for (int s2 = 0; s2 < 64; ++s2) colour[s2] = -1;
    for (int s0 = 0; s0 < 64; ++s0) {
      if (colour[s0] >= 0) continue;
      int cur = s0, col = 0;
      for (;;) {
        colour[cur] = col;
        const int ln = owner[cur] >> 1, sd = owner[cur] & 1;
        const int mate = slotOf[ln][sd ^ 1];                // lane partner: the other colour
        colour[mate] = col ^ 1;
        const int nxt = mate ^ 32;                          // its residue partner: back to `col`
        if (nxt == s0 || colour[nxt] >= 0) break;
        cur = nxt;
      }
    }
    for (uint32_t l = 0; l < 32; ++l) {
      const uint32_t p = base + l;
      if (colour[slotOf[l][0]] == 1) {                      // swap the lane's sides
        std::swap(out.vertA[p], out.vertB[p]);
        for (int k = 0; k < 4; ++k) std::swap(out.wA[(size_t)p * 4 + k], out.wB[(size_t)p * 4 + k]);
        std::swap(slotOf[l][0], slotOf[l][1]);
      }
      out.slotA[p] = (uint8_t)slotOf[l][0];
      out.slotB[p] = (uint8_t)slotOf[l][1];
      int n = 1;
      for (int k = 0; k < 4; ++k) if (out.wA[(size_t)p * 4 + k] != 0.f || out.wB[(size_t)p * 4 + k] != 0.f) n = k + 1;
      out.laneN[p] = (uint8_t)n;
    }

    // ---- quarter-warps.  The interleaved output stages 32-byte records with two STS.128 per vertex; the hardware serves a
    // 128-bit store one quarter-warp (8 lanes) at a time, and with the half-swapping order of deform2_kernel.cuh the 8 stores
    // of a quarter-warp hit 8 distinct 16-byte bank groups iff its 8 slots are distinct MODULO 8.  Aligned lane pairs must
    // stay together (they share palette rows), but the 16 pairs may sit anywhere: hill-climb over swaps of pairs between
    // the four quarter-warps until no swap lowers the number of colliding slots (A side + B side).
    {
      int pos[16];                                          // pos[i]: pair currently at pair position i
      for (int i = 0; i < 16; ++i) pos[i] = i;
      auto residue = [&](int pair, int lane01, int side) { return (int)((side ? out.slotB : out.slotA)[base + 2 * pair + lane01] & 7u); };
      auto bin_cost = [&](const int* members) {             // colliding slots of one quarter-warp (4 pairs)
        int cost = 0;
        for (int side = 0; side < 2; ++side) {
          int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
          for (int m = 0; m < 4; ++m) for (int h = 0; h < 2; ++h) cnt[residue(members[m], h, side)]++;
          for (int c2 = 0; c2 < 8; ++c2) cost += cnt[c2] > 1 ? cnt[c2] - 1 : 0;
        }
        return cost;
      };
      auto total_cost = [&](const int* q) { return bin_cost(q) + bin_cost(q + 4) + bin_cost(q + 8) + bin_cost(q + 12); };
      auto climb = [&](int* q) {
        bool improved = true;
        for (int sweep = 0; sweep < 12 && improved; ++sweep) {
          improved = false;
          for (int i = 0; i < 16; ++i)
            for (int j = i + 1; j < 16; ++j) {
              if (i / 4 == j / 4) continue;
              const int before = bin_cost(&q[i / 4 * 4]) + bin_cost(&q[j / 4 * 4]);
              if (before == 0) continue;
              std::swap(q[i], q[j]);
              const int after = bin_cost(&q[i / 4 * 4]) + bin_cost(&q[j / 4 * 4]);
              if (after < before) improved = true; else std::swap(q[i], q[j]);
            }
        }
        return total_cost(q);
      };
      int best = climb(pos);
      uint32_t rs = 0x9E3779B9u ^ (g * 2654435761u);        // a few deterministic restarts from shuffled orders
      for (int restart = 0; restart < 4 && best > 0; ++restart) {
        int q[16];
        for (int i = 0; i < 16; ++i) q[i] = i;
        for (int i = 15; i > 0; --i) { rs = rs * 1664525u + 1013904223u; std::swap(q[i], q[(rs >> 8) % (uint32_t)(i + 1)]); }
        const int c2 = climb(q);
        if (c2 < best) { best = c2; for (int i = 0; i < 16; ++i) pos[i] = q[i]; }
      }
      bool moved = false;
      for (int i = 0; i < 16; ++i) moved |= pos[i] != i;
      if (moved) {
        uint32_t vA[32], vB[32];
        uint16_t gj[32][4];
        float wa[32][4], wb[32][4];
        uint8_t sA[32], sB[32], ln[32];
        for (int l = 0; l < 32; ++l) {
          const uint32_t p = base + l;
          vA[l] = out.vertA[p]; vB[l] = out.vertB[p]; sA[l] = out.slotA[p]; sB[l] = out.slotB[p]; ln[l] = out.laneN[p];
          for (int k = 0; k < 4; ++k) { gj[l][k] = out.gatherJ[(size_t)p * 4 + k]; wa[l][k] = out.wA[(size_t)p * 4 + k]; wb[l][k] = out.wB[(size_t)p * 4 + k]; }
        }
        for (int i = 0; i < 16; ++i)
          for (int h = 0; h < 2; ++h) {
            const uint32_t p = base + 2 * i + h;
            const int src = 2 * pos[i] + h;
            out.vertA[p] = vA[src]; out.vertB[p] = vB[src]; out.slotA[p] = sA[src]; out.slotB[p] = sB[src]; out.laneN[p] = ln[src];
            for (int k = 0; k < 4; ++k) { out.gatherJ[(size_t)p * 4 + k] = gj[src][k]; out.wA[(size_t)p * 4 + k] = wa[src][k]; out.wB[(size_t)p * 4 + k] = wb[src][k]; }
          }
      }
    }
  }
  });
}

}  // namespace rz
