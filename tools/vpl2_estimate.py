"""What would "two vertices per lane" buy?  (CPU only; a planning tool for the next round)

The deform kernel is bound by the shared-memory pipe: a warp-wide gather of one 48-byte palette row per lane costs 6.75
cycles when every aligned lane pair reads the same row and 12 otherwise (profiles/r01_ubench_lds_row_fetch.txt), and a
warp of 32 vertices needs one such gather per influence slot.  If a lane evaluated TWO vertices that share one slot list,
a warp would cover 64 vertices with the same gathers.  This script measures how many gathers that needs on a given mesh:

  1. inside every 64-vertex window, vertices are paired greedily so that the union of their bone sets stays small
     (<= 4 bones is required: the record has four slots; windows that cannot be paired that way count as "fallback");
  2. every pair becomes one virtual vertex carrying the union, and the library's own pair packer (rz_plan_lanes, mode 2)
     plans the virtual mesh -- its histogram gives slots per warp and how many of them land on the fast path.

Output: shared-memory cycles spent on gathers per 64 vertices, today (two warps) and with two vertices per lane."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from reze_engine_b200 import capi, synth  # noqa: E402

FAST, SLOW = 6.75, 12.0


def gather_cycles(hist):
    """hist[N][m] = warps with N slots of which m are mixed (slow); returns (warps, total cycles, slot gathers, fast share)."""
    warps = cyc = slots = fast = 0
    for N in range(5):
        for m in range(5):
            w = int(hist[N][m])
            warps += w
            cyc += w * ((N - m) * FAST + m * SLOW)
            slots += w * N
            fast += w * (N - m)
    return warps, cyc, slots, fast / max(slots, 1)


def bone_sets(J, W):
    out = []
    for j, w in zip(J, W):
        s = frozenset(int(b) for b, x in zip(j, w) if x > 0)
        out.append(s if s else frozenset([int(j[0])]))
    return out


def pair_windows(S, V):
    """Greedy pairing inside 64-vertex windows; returns (pairs as (a, b|None, union), windows that needed > 4 slots)."""
    pairs, failed = [], 0
    for w0 in range(0, V, 64):
        vs = list(range(w0, min(V, w0 + 64)))
        order = sorted(vs, key=lambda v: sorted(S[v]))
        used, win, bad = set(), [], False
        for i, v in enumerate(order):
            if v in used:
                continue
            best, bu = None, 99
            for u in order[i + 1:]:
                if u in used:
                    continue
                un = len(S[v] | S[u])
                if un < bu:
                    bu, best = un, u
                if un == len(S[v]):
                    break
            used.add(v)
            if best is None:
                win.append((v, None, S[v]))
                continue
            used.add(best)
            win.append((v, best, S[v] | S[best]))
            bad |= bu > 4
        if bad:
            failed += 1
            # fallback: the window runs one vertex per lane (two warps' worth of lanes, each with its own set)
            win = [(v, None, S[v]) for v in vs]
        pairs.append(win)
    return pairs, failed


def analyse(name, J, W, B):
    V = len(J)
    plan1 = capi.plan_lanes(J, W, B, 2)
    w1, c1, s1, f1 = gather_cycles(plan1["hist"])
    # un-packed warps (a bone listed twice, ...) are not in the histogram: count them at the slow rate with their slot count
    S = bone_sets(J, W)
    pairs, failed = pair_windows(S, V)
    vj, vw = [], []
    for win in pairs:
        # pad a window to whole warps with copies of its first lane (adds no bone the warp does not already gather)
        lanes = win + [win[0]] * ((-len(win)) % 32)
        # a fallback window has up to 64 single-vertex lanes = two warps of the virtual mesh
        for _, _, un in lanes:
            bones = sorted(un)[:4]
            j = bones + [bones[0] if bones else 0] * (4 - len(bones))
            w = [255 // max(len(bones), 1)] * len(bones) + [0] * (4 - len(bones))
            if bones:
                w[0] += 255 - sum(w)
            else:
                w = [255, 0, 0, 0]
            vj.append(j)
            vw.append(w)
    vj, vw = np.array(vj, np.uint16), np.array(vw, np.uint8)
    pad = (-len(vj)) % 32
    if pad:
        vj = np.concatenate([vj, np.zeros((pad, 4), np.uint16)])
        vw = np.concatenate([vw, np.tile(np.array([[255, 0, 0, 0]], np.uint8), (pad, 1))])
    plan2 = capi.plan_lanes(vj, vw, B, 2)
    w2, c2, s2, f2 = gather_cycles(plan2["hist"])
    n64 = (V + 63) // 64
    # (warps the packer leaves in the caller's slot order -- a vertex listing one bone twice -- are not in its histogram:
    #  `packed_warps` < `warps` means the cycle figures of that side are a lower bound)
    row = dict(mesh=name, V=V, windows=n64, fallback_windows=failed,
               today=dict(slot_gathers_per_64=s1 / n64, fast_share=f1, gather_cycles_per_64=c1 / n64, packed_warps=w1, warps=(V + 31) // 32),
               two_per_lane=dict(slot_gathers_per_64=s2 / n64, fast_share=f2, gather_cycles_per_64=c2 / n64, packed_warps=w2, warps=len(vj) // 32),
               staging_cycles_per_64=24.0)
    t1, t2 = c1 / n64 + 24.0, c2 / n64 + 24.0
    row["lsu_cycles_per_64"] = dict(today=t1, two_per_lane=t2, saving=1 - t2 / t1)
    # the real planner (csrc/lane_plan2.h: bottleneck blossom matching instead of the greedy pairing above)
    r = capi.plan_lanes2(J, W, B)
    c3 = r["fast"] * FAST + (r["total"] - r["fast"]) * SLOW
    row["rz_plan_lanes2"] = dict(slot_gathers_per_64=r["total"] / n64, fast_share=r["fast"] / max(r["total"], 1), gather_cycles_per_64=c3 / n64,
                                 paired_windows=r["pairedWindows"], fallback_windows=r["fallbackWindows"],
                                 lsu_saving=1 - (c3 / n64 + 24.0) / t1)
    print(json.dumps(row))
    return row


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--verts", type=int, default=100_000)
    ap.add_argument("--bones", type=int, default=512)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r01_two_vertices_per_lane_estimate.jsonl"))
    a = ap.parse_args()
    rows = []
    wl = synth.make_workload(a.verts, a.bones)
    rows.append(analyse("synthetic headline mesh", wl.joints.reshape(-1, 4), wl.weights.reshape(-1, 4), a.bones))
    for nm in ("serqet2", "serqet"):
        p = os.path.join(ROOT, "tests", "golden", "_local", nm + ".npz")
        if os.path.exists(p):
            z = np.load(p)
            rows.append(analyse(nm + ".pmx", z["joints"].reshape(-1, 4), z["weights"].reshape(-1, 4), z["invBind"].size // 16))
    with open(a.out, "w") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
