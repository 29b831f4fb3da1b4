"""Randomised parity sweep (GPU box): random mesh sizes, bone counts, instance counts, feature flags, launch shapes and
instance sub-ranges through the C ABI against the CPU oracle.  Complements tests/ (fixed cases) with combinations nobody
wrote down.  Prints one line per case, exits non-zero on the first mismatch.  Writes gpurun_out/fuzz.jsonl."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import oracle as orc  # noqa: E402
from reze_engine_b200 import capi, synth  # noqa: E402

TOL = 1e-5
LITE = [(0, 0), (1, 256), (2, 256), (2, 512), (4, 512)]
FULL = LITE + [(3, 256), (6, 512), (4, 768), (3, 768), (2, 1024), (6, 256)]
V2 = [(0, 0, 0), (4, 512, 2), (4, 384, 2), (6, 384, 2), (6, 512, 1), (6, 256, 3), (5, 512, 1), (3, 256, 3), (2, 256, 2), (1, 256, 1), (4, 256, 4)]   # (I, threads, sub-batch)


def rel(a, b):
    return float(np.abs(np.asarray(a, np.float64) - b).max() / max(np.abs(b).max(), 1.0))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=60)
    ap.add_argument("--seed", type=int, default=2026)
    ap.add_argument("--seconds", type=float, default=150.0)
    ap.add_argument("--plain-share", type=float, default=0.4, help="share of cases forced onto the plain planar path (two-vertices-per-lane kernel)")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "fuzz.jsonl"))
    a = ap.parse_args()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    rng = np.random.default_rng(a.seed)
    t0 = time.time()
    worst = 0.0
    for case in range(a.cases):
        if time.time() - t0 > a.seconds:
            break
        V = int(rng.choice([1, 3, 31, 33, 255, 257, 777, 2049, 6000, 20011]))
        B = int(rng.choice([1, 2, 17, 64, 349, 512, 1500, 5000]))
        K = int(rng.choice([1, 2, 5, 7, 12, 33]))
        M = int(rng.choice([0, 0, 3, 20]))
        sdef = bool(rng.integers(0, 2))
        flags = 0
        layout = int(rng.integers(0, 4))                    # 0 planar, 1 outline, 2 interleaved, 3 positions only
        force_plain = a.plain_share > 0 and rng.random() < a.plain_share   # the plain planar path (two-vertex kernel) often enough
        if force_plain:
            M, sdef, layout = 0, False, 0
        if layout == 1:
            flags |= capi.RZ_FLAG_OUTLINE
        elif layout == 2:
            flags |= capi.RZ_FLAG_INTERLEAVED
        elif layout == 3:
            flags |= capi.RZ_FLAG_NO_NORMALS
        if rng.integers(0, 2) and not force_plain:
            flags |= capi.RZ_FLAG_BOUNDS
        if sdef:
            flags |= capi.RZ_FLAG_SDEF
        if rng.integers(0, 4) == 0:
            flags |= capi.RZ_FLAG_REORDER_VERTICES
        if rng.integers(0, 3) == 0:
            flags |= capi.RZ_FLAG_DOUBLE_BUFFER
        # host uploads with one palette per instance are pipelined in blocks of this many palettes (rze_b200.cu)
        os.environ["RZ_PIPELINE_BLOCK"] = str(int(rng.choice([1, 2, 3, 64])))
        plain = M == 0 and not sdef and not (flags & ~(capi.RZ_FLAG_REORDER_VERTICES | capi.RZ_FLAG_DOUBLE_BUFFER))
        I, nt = (FULL if plain else LITE)[int(rng.integers(0, len(FULL if plain else LITE)))]
        first = int(rng.integers(0, K))
        count = int(rng.integers(1, K - first + 1))
        # the plain planar path: half the cases on the two-vertices-per-lane kernel in one of its shapes, half on the one-vertex kernel
        vpl, sb = 0, 0
        if plain and B <= 4000:
            if rng.integers(0, 2):
                vpl = 2
                I, nt, sb = V2[int(rng.integers(0, len(V2)))]
            else:
                vpl = 1
        wl = synth.make_workload(V, B, M=M, sdef=sdef, seed=int(rng.integers(1, 1 << 30)))
        P = K if rng.integers(0, 3) == 0 else int(rng.integers(1, K + 1))     # one palette per instance a third of the time
        world0 = synth.make_palettes(wl.bones, P, rng)          # a first frame that must leave no trace
        world = synth.make_palettes(wl.bones, P, rng)
        i2p = None if P == K and rng.integers(0, 4) else rng.integers(0, P, K).astype(np.uint32)
        if i2p is None and P < K:
            i2p = rng.integers(0, P, K).astype(np.uint32)
        mw = rng.uniform(-0.3, 1.0, (K, max(M, 1))).astype(np.float32)
        edge = rng.uniform(0, 2, V).astype(np.float32)
        row = dict(case=case, V=V, B=B, K=K, P=P, M=M, sdef=sdef, flags=flags, I=I, nt=nt, first=first, count=count,
                   pipe_block=os.environ["RZ_PIPELINE_BLOCK"], identity=i2p is None)
        try:
            with capi.DeformContext(max_instances=K, flags=flags, instances_per_group=I, threads=nt, store_mode=sb, vertices_per_lane=vpl) as ctx:
                ctx.load_mesh(wl.vtx8, wl.joints, wl.weights, wl.invBind)
                if M:
                    ctx.load_morphs(wl.morphs.offsets, wl.morphs.vertexIndex, wl.morphs.delta)
                if sdef:
                    ctx.load_sdef(wl.sdef.vertexIndex, wl.sdef.c_r0_r1)
                if flags & capi.RZ_FLAG_OUTLINE:
                    ctx.load_edge_size(edge)
                try:
                    ctx.set_palettes(world0, i2p, K=K)
                    if M:
                        ctx.set_morph_weights(mw, np.arange(M), K=K)
                    ctx.deform()
                    ctx.set_palettes(world, i2p, K=K)
                    ctx.deform(first, count)
                except capi.RzError as e:
                    if "not built" in str(e) or "does not fit" in str(e):
                        row["skipped"] = str(e)[:90]
                        print(json.dumps(row), flush=True)
                        continue
                    raise
                err = 0.0
                for k in range(first, first + count):
                    p = k if i2p is None else int(i2p[k])
                    rp, rn = orc.deform(wl.vtx8, wl.joints, wl.weights, orc.skin_matrices(world[p], wl.invBind),
                                        morph=(wl.morphs.offsets, wl.morphs.vertexIndex, wl.morphs.delta) if M else None,
                                        morphW=mw[k, :M] if M else None, sdef=(wl.sdef.vertexIndex, wl.sdef.c_r0_r1) if sdef else None)
                    gp, gn = ctx.read_instance(k, normals=layout != 3)
                    err = max(err, rel(gp, rp))
                    if gn is not None:
                        err = max(err, rel(gn, rn))
                    if layout == 1:
                        err = max(err, rel(ctx.read_outline(k), orc.outline_hull(rp, rn, edge)))
                    if layout == 2:
                        st = ctx.read_interleaved(k)
                        assert np.array_equal(st[:, :3], gp) and np.array_equal(st[:, 3:6], gn) and np.array_equal(st[:, 6:], wl.vtx8.reshape(-1, 8)[:, 6:])
                    if flags & capi.RZ_FLAG_BOUNDS:
                        bb = ctx.read_bounds(k, 1)[0]
                        assert np.array_equal(bb[:3], gp.min(axis=0)) and np.array_equal(bb[3:], gp.max(axis=0)), "bounds"
                j, w = ctx.read_skinning()
                assert np.array_equal(j, wl.joints.reshape(-1)) and np.array_equal(w, wl.weights.reshape(-1)), "integer tables"
                s = ctx.stats()
                row.update(err=err, usedI=s["instancesPerGroup"], usedThreads=s["threads"], vpl=s["verticesPerLane"])
                assert vpl == 0 or s["verticesPerLane"] == vpl, "kernel variant"
                worst = max(worst, err)
                assert err <= TOL, f"parity {err}"
        except Exception as e:  # noqa: BLE001
            row["FAILED"] = repr(e)[:300]
            print(json.dumps(row), flush=True)
            with open(a.out, "a") as f:
                f.write(json.dumps(row) + "\n")
            sys.exit(1)
        print(json.dumps(row), flush=True)
        with open(a.out, "a") as f:
            f.write(json.dumps(row) + "\n")
    print(json.dumps(dict(done=True, worst_rel_err=worst, seconds=time.time() - t0)), flush=True)


if __name__ == "__main__":
    main()
