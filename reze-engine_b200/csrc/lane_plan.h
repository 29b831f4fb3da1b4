// lane_plan.h — load-time planning of which lane of a warp evaluates which vertex and which influence slot holds which
// (bone, weight) pair.  Pure host C++ (no CUDA): used by rebuild_tables (rze_b200.cu) and, through rz_plan_lanes, by the
// CPU test-suite.  See DESIGN.md section 4, "pair packing".
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

namespace rz {

constexpr uint8_t kNoSlot = 0xFF;    // LanePlan::slotMap: the caller's influence has zero weight and occupies no device slot

struct LanePlan {
  uint32_t V = 0, Vp = 0;
  std::vector<uint32_t> procVertex;   // [Vp] vertex evaluated by processing index p = warp*32 + lane (~0u: padding)
  std::vector<uint32_t> procSlot;     // [Vp] position of that vertex among the warp's 32 consecutive outputs
  std::vector<uint16_t> gatherJ;      // [Vp][4] bone whose palette row influence slot s gathers (never out of range)
  std::vector<float> devW;            // [Vp][4] weight of slot s (shader-normalised, engine.ts:255-258)
  std::vector<uint8_t> devN;          // [Vp] highest slot with a non-zero weight + 1
  std::vector<uint8_t> slotMap;       // [Vp][4] device slot that holds the caller's influence k (kNoSlot: zero weight, none)
  uint64_t fastSlots = 0, totalSlots = 0;   // packed warps: gather instructions on the broadcast fast path / all
  uint32_t hist[5][5] = {};           // packed warps by [slot count N][mixed slots m*]
};

// Maximum matching in a general graph on at most MAXN nodes (Edmonds' blossom algorithm, O(n^3)): the lane planners ask
// "is there a PERFECT matching whose every pair satisfies a threshold?" for increasing thresholds (bottleneck matching).
template <int MAXN>
struct BlossomMatcher {
  int n = MAXN;
  bool adj[MAXN][MAXN];
  int match[MAXN], par[MAXN], base[MAXN], q[2 * MAXN];
  bool used[MAXN], blossom[MAXN];
  int lca(int a, int b2) {
    bool seen[MAXN] = {false};
    for (;;) { a = base[a]; seen[a] = true; if (match[a] < 0) break; a = par[match[a]]; }
    for (;;) { b2 = base[b2]; if (seen[b2]) return b2; b2 = par[match[b2]]; }
  }
  void mark_path(int v, int bb, int child) {
    while (base[v] != bb) {
      blossom[base[v]] = blossom[base[match[v]]] = true;
      par[v] = child; child = match[v]; v = par[match[v]];
    }
  }
  int find_path(int root) {
    for (int i = 0; i < n; ++i) { used[i] = false; par[i] = -1; base[i] = i; }
    int qh = 0, qt = 0;
    used[root] = true; q[qt++] = root;
    while (qh < qt) {
      const int v = q[qh++];
      for (int to = 0; to < n; ++to) {
        if (!adj[v][to] || base[v] == base[to] || match[v] == to) continue;
        if (to == root || (match[to] >= 0 && par[match[to]] >= 0)) {
          const int cb = lca(v, to);
          for (int i = 0; i < n; ++i) blossom[i] = false;
          mark_path(v, cb, to); mark_path(to, cb, v);
          for (int i = 0; i < n; ++i)
            if (blossom[base[i]]) { base[i] = cb; if (!used[i]) { used[i] = true; q[qt++] = i; } }
        } else if (par[to] < 0) {
          par[to] = v;
          if (match[to] < 0) return to;
          used[match[to]] = true; q[qt++] = match[to];
        }
      }
    }
    return -1;
  }
  // maximum matching; returns the number of matched pairs
  int solve() {
    for (int i = 0; i < n; ++i) match[i] = -1;
    for (int i = 0; i < n; ++i)                       // greedy start: nearest unmatched neighbour in key order
      if (match[i] < 0) for (int j = i + 1; j < n; ++j) if (match[j] < 0 && adj[i][j]) { match[i] = j; match[j] = i; break; }
    for (int i = 0; i < n; ++i)
      if (match[i] < 0) {
        int v = find_path(i);
        while (v >= 0) { const int pv = par[v], ppv = match[pv]; match[v] = pv; match[pv] = v; v = ppv; }
      }
    int m = 0;
    for (int i = 0; i < n; ++i) if (match[i] >= 0) ++m;
    return m / 2;
  }
};

// JT [V][4] joints, WT [V][4] unorm8 weights, isSdef [V] or nullptr, kTileVerts = tile granularity the vertex count is
// padded to.  permMode: 0 natural lane order, 1 lanes sorted by (influence count, bones), 2 pair packing.
inline void plan_lanes(const uint16_t* JT, const uint8_t* WT, const uint8_t* isSdef, uint32_t V, uint32_t B, uint32_t tileVerts,
                       int permMode, LanePlan& out) {
  const uint32_t Vp = (V + tileVerts - 1) / tileVerts * tileVerts;
  out.V = V; out.Vp = Vp;
  out.fastSlots = out.totalSlots = 0;
  memset(out.hist, 0, sizeof(out.hist));
  auto ninf_of = [&](uint32_t v) -> uint32_t {
    const uint8_t* w = &WT[(size_t)v * 4];
    if ((uint32_t)w[0] + w[1] + w[2] + w[3] == 0) return 1;   // shader rule: sum <= 1e-4 -> (1,0,0,0)
    uint32_t n = 1;
    for (uint32_t k = 0; k < 4; ++k) if (w[k]) n = k + 1;
    return n;
  };
  const bool classSort = permMode != 0;

  // lane order inside a warp: by influence count, then by bone ids, so that quarter-warps (the unit the shared-memory
  // pipe serves per wavefront) are homogeneous: same influence count => whole quarters skip the zero-weight gathers,
  // same bones => same address => no bank conflict.
  auto key2_of = [&](uint32_t v) -> uint64_t {
    const uint16_t* j = &JT[(size_t)v * 4];
    const uint32_t n = ninf_of(v);
    uint64_t k = (uint64_t)((isSdef && isSdef[v]) ? 1 : 0) << 63 | (uint64_t)n << 60;
    k |= (uint64_t)(j[0] & 0x7FFF) << 45;
    k |= (uint64_t)(n > 1 ? (j[1] & 0x7FFF) : 0) << 30;
    k |= (uint64_t)(n > 2 ? (j[2] & 0x7FFF) : 0) << 15;
    k |= (uint64_t)(n > 3 ? (j[3] & 0x7FFF) : 0);
    return k;
  };
  // Per processing index p (= warp*32 + lane): the vertex it evaluates, and the influence table the DEVICE sees --
  // devJ[p][s] bone gathered for influence slot s (kBorrow: weight is zero, any row will do, chosen further down),
  // devW[p][s] its weight, devN[p] = highest slot with a non-zero weight + 1.
  constexpr uint16_t kBorrow = 0xFFFFu;
  std::vector<uint32_t>& procVertex = out.procVertex;
  std::vector<uint32_t>& procSlot = out.procSlot;
  std::vector<float>& devW = out.devW;
  std::vector<uint8_t>& devN = out.devN;
  procVertex.assign(Vp, ~0u); procSlot.assign(Vp, 0);
  std::vector<uint16_t> devJ((size_t)Vp * 4, kBorrow);
  devW.assign((size_t)Vp * 4, 0.f);
  devN.assign(Vp, 1);
  out.slotMap.assign((size_t)Vp * 4, 0);
  for (uint32_t p = 0; p < Vp; ++p) for (uint32_t k = 0; k < 4; ++k) out.slotMap[(size_t)p * 4 + k] = (uint8_t)k;

  // weights exactly as the reference's vertex shader derives them (engine.ts:255-258): unorm8 -> f32, sum in slot order,
  // renormalise when the sum exceeds 1e-4 else (1,0,0,0).  IEEE f32 on the host == on the device.
  auto shader_weights = [&](uint32_t v, float w[4]) {
    const uint8_t* w8 = &WT[(size_t)v * 4];
    for (int k = 0; k < 4; ++k) w[k] = (float)w8[k] / 255.0f;
    const float wsum = w[0] + w[1] + w[2] + w[3];
    if (wsum > 0.0001f) {
      const float inv = 1.0f / wsum;
      for (int k = 0; k < 4; ++k) w[k] = w[k] * inv;
    } else {
      w[0] = 1.f; w[1] = w[2] = w[3] = 0.f;
    }
  };

  // ---- (A) caller's slot order: every warp keeps its 32 CONSECUTIVE output vertices (its results leave as one contiguous
  // 384-byte TMA bulk store per plane), only the lane order inside the warp is chosen: by influence count, then bones.
  auto fill_plain = [&](uint32_t w0) {
    uint32_t vs[32];
    uint32_t n = 0;
    for (uint32_t l = 0; l < 32; ++l) if (w0 + l < V) vs[n++] = w0 + l;
    if (classSort) std::stable_sort(vs, vs + n, [&](uint32_t a, uint32_t b2) { return key2_of(a) < key2_of(b2); });
    for (uint32_t lane = 0; lane < 32; ++lane) {
      const uint32_t p = w0 + lane;
      for (uint32_t k = 0; k < 4; ++k) { devJ[(size_t)p * 4 + k] = kBorrow; devW[(size_t)p * 4 + k] = 0.f; out.slotMap[(size_t)p * 4 + k] = (uint8_t)k; }
      if (lane < n) {
        const uint32_t v = vs[lane];
        procVertex[p] = v; procSlot[p] = v - w0;
        float w[4];
        shader_weights(v, w);
        const uint32_t ni = ninf_of(v);
        for (uint32_t k = 0; k < ni; ++k) { devJ[(size_t)p * 4 + k] = JT[(size_t)v * 4 + k]; devW[(size_t)p * 4 + k] = w[k]; }
        devN[p] = (uint8_t)ni;
      } else {
        procVertex[p] = ~0u; procSlot[p] = lane;               // padding: unused output slots n..31 in lane order
        devW[(size_t)p * 4] = 1.f; devN[p] = 1;
      }
    }
  };

  // ---- (B) pair packing.  A warp-wide gather of one 48-byte palette row per lane costs 6.75 shared-memory cycles when
  // every ALIGNED LANE PAIR (2l, 2l+1) reads the same row and 12 otherwise -- one mixed pair is enough to lose it
  // (profiles/r01_ubench_lds_row_fetch.txt), and these gathers are what bounds the kernel.  Three load-time freedoms buy
  // the fast case without touching the arithmetic of any vertex: which lane evaluates which of the warp's 32 vertices,
  // in which influence SLOT a vertex keeps each of its (bone, weight) pairs (the blend is a sum), and which row a
  // zero-weight slot gathers.  For two vertices with bone sets A and B and the warp's slot count N = max |set|, m(A,B) is
  // the least number of slots in which the two lanes must read different rows (0 iff |A u B| <= N: the lanes then share
  // one slot list, each with weight 0 where the bone is not its own).  The warp needs a perfect matching of its 32 lanes
  // minimising the largest m (bottleneck matching: thresholds 0..N, Edmonds' blossom algorithm for each); the mixed slots
  // of all pairs are then parked in the LAST m* slots, so N - m* gather instructions of the warp run at the fast rate.
  using PairMatcher = BlossomMatcher<32>;
  struct LaneSet { uint32_t v; int n; uint16_t b[4]; float w[4]; uint8_t src[4]; uint64_t key; };
  uint64_t packStat[5] = {0, 0, 0, 0, 0};   // warp-slots: total, fast
  auto fill_packed = [&](uint32_t w0) -> bool {
    LaneSet ls[32];
    int N = 1;
    for (uint32_t l = 0; l < 32; ++l) {
      LaneSet& e = ls[l];
      e.v = (w0 + l < V) ? w0 + l : ~0u;
      e.n = 0; e.key = 0;
      if (e.v == ~0u) continue;
      if ((isSdef && isSdef[e.v])) return false;                              // SDEF reads slots 0/1 by position
      float w[4];
      shader_weights(e.v, w);
      for (uint32_t k = 0; k < 4; ++k)
        if (w[k] != 0.f) {
          const uint16_t bone = JT[(size_t)e.v * 4 + k];
          for (int t = 0; t < e.n; ++t) if (e.b[t] == bone) return false;   // a bone listed twice: keep the caller's slots
          e.b[e.n] = bone; e.w[e.n] = w[k]; e.src[e.n] = (uint8_t)k; ++e.n;
        }
      if (e.n == 0) return false;
      N = std::max(N, e.n);
      uint16_t sb[4];
      for (int t = 0; t < e.n; ++t) sb[t] = e.b[t];
      std::sort(sb, sb + e.n);
      for (int t = 0; t < e.n; ++t) e.key |= (uint64_t)(sb[t] & 0x7FFF) << (45 - 15 * t);
    }
    // lanes in key order (similar sets become neighbours: greedy start of the matcher, homogeneous quarter-warps)
    int ord[32];
    for (int i = 0; i < 32; ++i) ord[i] = i;
    std::stable_sort(ord, ord + 32, [&](int x, int y) {
      const bool px = ls[x].v == ~0u, py = ls[y].v == ~0u;
      if (px != py) return py;
      return ls[x].key < ls[y].key;
    });
    auto common = [&](const LaneSet& A, const LaneSet& Bv) { int cc = 0; for (int i = 0; i < A.n; ++i) for (int j = 0; j < Bv.n; ++j) if (A.b[i] == Bv.b[j]) ++cc; return cc; };
    int mneed[32][32];
    for (int x = 0; x < 32; ++x)
      for (int y = x + 1; y < 32; ++y) {
        const LaneSet &A = ls[ord[x]], &Bv = ls[ord[y]];
        const int cc = common(A, Bv);
        int m = 0;
        for (; m < N; ++m) {
          const int a = std::max(A.n - m, 0), b2 = std::max(Bv.n - m, 0);
          const int Lmin = cc >= std::max(a, b2) ? std::max(a, b2) : cc + std::max(a - cc, 0) + std::max(b2 - cc, 0);
          if (Lmin <= N - m) break;
        }
        mneed[x][y] = mneed[y][x] = m;
      }
    PairMatcher pm;
    int mstar = 0;
    for (; mstar <= N; ++mstar) {
      for (int x = 0; x < 32; ++x) for (int y = 0; y < 32; ++y) pm.adj[x][y] = x != y && mneed[x][y] <= mstar;
      if (pm.solve() == 16) break;
    }
    if (mstar > N) return false;   // cannot happen (threshold N admits every pair)
    packStat[0] += (uint64_t)N; packStat[1] += (uint64_t)(N - mstar);
    out.hist[N][mstar]++;
    const int cap = N - mstar;
    // bones that many lanes of the warp share go to the low slots everywhere (fewer distinct rows per gather instruction)
    auto freq = [&](uint16_t bone) { int f = 0; for (int l = 0; l < 32; ++l) for (int t = 0; t < ls[l].n; ++t) if (ls[l].b[t] == bone) ++f; return f; };
    struct PairOut { int la, lb; uint16_t L[4]; int nL; uint64_t key; };
    PairOut po[16];
    int np = 0;
    for (int x = 0; x < 32; ++x) {
      const int y = pm.match[x];
      if (y < x) continue;
      PairOut& o = po[np++];
      o.la = ord[x]; o.lb = ord[y]; o.nL = 0; o.key = 0;
      const LaneSet &A = ls[o.la], &Bv = ls[o.lb];
      auto inL = [&](uint16_t bone) { for (int t = 0; t < o.nL; ++t) if (o.L[t] == bone) return true; return false; };
      auto has = [&](const LaneSet& S, uint16_t bone) { for (int t = 0; t < S.n; ++t) if (S.b[t] == bone) return true; return false; };
      // shared list: common bones first (each one serves both lanes), then whatever either lane cannot park in its m* own slots,
      // then -- capacity permitting -- the rest (a shared slot is never worse than a private one)
      for (int t = 0; t < A.n && o.nL < cap; ++t) if (has(Bv, A.b[t])) o.L[o.nL++] = A.b[t];
      auto outside = [&](const LaneSet& S) { int r = 0; for (int t = 0; t < S.n; ++t) if (!inL(S.b[t])) ++r; return r; };
      for (int t = 0; t < A.n && outside(A) > mstar && o.nL < cap; ++t) if (!inL(A.b[t])) o.L[o.nL++] = A.b[t];
      for (int t = 0; t < Bv.n && outside(Bv) > mstar && o.nL < cap; ++t) if (!inL(Bv.b[t])) o.L[o.nL++] = Bv.b[t];
      if (outside(A) > mstar || outside(Bv) > mstar) return false;   // cannot happen (mneed <= m*)
      for (int t = 0; t < A.n && o.nL < cap; ++t) if (!inL(A.b[t])) o.L[o.nL++] = A.b[t];
      for (int t = 0; t < Bv.n && o.nL < cap; ++t) if (!inL(Bv.b[t])) o.L[o.nL++] = Bv.b[t];
      std::stable_sort(o.L, o.L + o.nL, [&](uint16_t p1, uint16_t p2) { const int f1 = freq(p1), f2 = freq(p2); return f1 != f2 ? f1 > f2 : p1 < p2; });
      for (int t = 0; t < o.nL; ++t) o.key |= (uint64_t)(o.L[t] & 0x7FFF) << (45 - 15 * t);
      if (A.v == ~0u && Bv.v == ~0u) o.key = ~0ull;
    }
    int pord[16];
    for (int i = 0; i < 16; ++i) pord[i] = i;
    std::stable_sort(pord, pord + 16, [&](int x, int y) { return po[x].key < po[y].key; });
    uint32_t padSlot = 0;
    bool slotUsed[32] = {false};
    for (uint32_t l = 0; l < 32; ++l) if (w0 + l < V) slotUsed[l] = true;
    for (int pi = 0; pi < 16; ++pi) {
      const PairOut& o = po[pord[pi]];
      const int lanes[2] = {o.la, o.lb};
      for (int h = 0; h < 2; ++h) {
        const LaneSet& S = ls[lanes[h]];
        const LaneSet& T = ls[lanes[h ^ 1]];
        const uint32_t p = w0 + (uint32_t)pi * 2 + (uint32_t)h;
        uint16_t* dj = &devJ[(size_t)p * 4];
        float* dw = &devW[(size_t)p * 4];
        uint8_t* sm = &out.slotMap[(size_t)p * 4];
        for (int k = 0; k < 4; ++k) { dj[k] = kBorrow; dw[k] = 0.f; sm[k] = kNoSlot; }
        if (S.v == ~0u) {
          while (padSlot < 32 && slotUsed[padSlot]) ++padSlot;
          procVertex[p] = ~0u; procSlot[p] = padSlot; slotUsed[padSlot] = true;
          // a rigid unit-weight dummy on whatever row its partner reads in slot 0
          for (int t = 0; t < o.nL; ++t) dj[t] = o.L[t];
          dw[0] = 1.f; devN[p] = 1;
          continue;
        }
        procVertex[p] = S.v; procSlot[p] = S.v - w0;
        bool placed[4] = {false, false, false, false};
        int hi = 0;
        for (int t = 0; t < o.nL; ++t) {                   // shared slots: the pair's common row, own weight or zero
          dj[t] = o.L[t];
          for (int u = 0; u < S.n; ++u) if (S.b[u] == o.L[t]) { dw[t] = S.w[u]; placed[u] = true; sm[S.src[u]] = (uint8_t)t; hi = t + 1; }
        }
        int fs = cap;                                       // private slots: the lane's remaining bones
        for (int u = 0; u < S.n; ++u)
          if (!placed[u]) { dj[fs] = S.b[u]; dw[fs] = S.w[u]; sm[S.src[u]] = (uint8_t)fs; hi = fs + 1; ++fs; }
        // leftover private slots mirror the partner's bone there (weight 0): one more shared row for free
        (void)T;
        devN[p] = (uint8_t)std::max(hi, 1);
      }
    }
    // private slots still unassigned: take the partner's row when it has one
    for (uint32_t pi = 0; pi < 16; ++pi)
      for (int k = 0; k < 4; ++k) {
        uint16_t& ja = devJ[(size_t)(w0 + 2 * pi) * 4 + k];
        uint16_t& jb = devJ[(size_t)(w0 + 2 * pi + 1) * 4 + k];
        if (ja == kBorrow && jb != kBorrow) ja = jb;
        else if (jb == kBorrow && ja != kBorrow) jb = ja;
      }
    // slot-map entries of zero-weight caller slots: point at a device slot that is not one of the vertex' own
    return true;
  };

  for (uint32_t w0 = 0; w0 < Vp; w0 += 32) {
    if (!(permMode >= 2 && fill_packed(w0))) fill_plain(w0);
  }
  out.fastSlots = packStat[1];
  out.totalSlots = packStat[0];

  // joints the kernel gathers: a slot still marked kBorrow (zero weight, no partner row) takes the row of an ACTIVE lane
  // of its own warp (same quarter-warp if possible; both lanes of a pair then pick the same one), so the unconditional
  // gather adds no shared-memory wavefront.
  std::vector<uint16_t>& gatherJ = out.gatherJ;
  gatherJ.assign((size_t)Vp * 4, 0);
  for (uint32_t w0 = 0; w0 < Vp; w0 += 32) {
    for (uint32_t k = 0; k < 4; ++k) {
      int firstWarp = -1, firstQ[4] = {-1, -1, -1, -1};
      for (uint32_t l = 0; l < 32; ++l) {
        if (devJ[(size_t)(w0 + l) * 4 + k] == kBorrow) continue;
        if (firstWarp < 0) firstWarp = (int)l;
        if (firstQ[l / 8] < 0) firstQ[l / 8] = (int)l;
      }
      for (uint32_t l = 0; l < 32; ++l) {
        uint16_t j = devJ[(size_t)(w0 + l) * 4 + k];
        if (j == kBorrow) {
          const int src = firstQ[l / 8] >= 0 ? firstQ[l / 8] : firstWarp;
          if (src >= 0) j = devJ[(size_t)(w0 + src) * 4 + k];
          else j = procVertex[w0 + l] != ~0u ? JT[(size_t)procVertex[w0 + l] * 4 + k] : 0;
          if (j >= B) j = 0;
        }
        gatherJ[(size_t)(w0 + l) * 4 + k] = j;
      }
    }
  }

}

}  // namespace rz
