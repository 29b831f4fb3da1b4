"""BASELINE.json configs 1-4 on one B200: parity on sampled instances (vs the CPU oracle) + device timing.
Writes gpurun_out/configs.jsonl (copied to profiles/ by hand).  Config 0 (CPU plumbing) is the -m "not gpu" test suite."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from reze_engine_b200 import capi, crowd, synth  # noqa: E402
from reze_engine_b200.model import Bone  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out", "configs.jsonl")
LOCAL = os.path.join(ROOT, "tests", "golden", "_local")


def rel_err(a, b):
    return float(np.abs(np.asarray(a, np.float64) - b).max() / max(np.abs(b).max(), 1.0))


def timed(ctx, stream, fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.median(ms))


def emit(row):
    print(json.dumps(row), flush=True)
    with open(OUT, "a") as f:
        f.write(json.dumps(row) + "\n")


def bones_from(z):
    out = []
    for i in range(len(z["parents"])):
        ap, ar = int(z["appendParent"][i]), float(z["appendRatio"][i])
        out.append(Bone(name=str(z["names"][i]), parentIndex=int(z["parents"][i]), bindTranslation=[float(x) for x in z["bindTranslation"][i]],
                        appendParentIndex=None if ap < 0 else ap, appendRatio=None if np.isnan(ar) else ar,
                        appendRotate=bool(z["appendRotate"][i]), appendMove=bool(z["appendMove"][i])))
    return out


def main():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    sh = stream.cuda_stream

    # ---- configs 1 + 2: the shipped PMX + pool.vmd (needs tests/golden/_local, generated from the reference assets)
    p = os.path.join(LOCAL, "serqet2.npz")
    if os.path.exists(p):
        z = np.load(p)
        vmd = np.load(os.path.join(LOCAL, "pool_vmd.npz"))
        bones = bones_from(z)
        B, V = len(bones), z["vtx8"].size // 8
        names = [b.name for b in bones]
        idx = {n: i for i, n in enumerate(names)}
        qa = np.tile(np.array([0, 0, 0, 1], np.float64), (B, 1))
        qb = qa.copy()
        for nm, row in zip(vmd["names"], vmd["data"]):            # keys at frame 0 -> start, last key -> target (pool.vmd: frames {0,45})
            i = idx.get(str(nm))
            if i is None:
                continue
            q = row[1:5] / np.linalg.norm(row[1:5])
            if row[0] == 0:
                qa[i] = q
            qb[i] = q
        for name, K, frames in (("config1: serqet2.pmx + pool.vmd, 46 frames, K=1", 1, 46), ("config2: serqet2.pmx crowd K=1024, staggered phase", 1024, 1)):
            with capi.DeformContext(max_instances=K, stream=sh) as ctx:
                ctx.load_mesh(z["vtx8"], z["joints"], z["weights"], z["invBind"])
                ctx.load_skeleton(bones)
                ctx.set_tweens(qa, qb, np.zeros(B, np.float32), np.full(B, 1500.0, np.float32), np.ones(B, np.uint8), qa)
                if K == 1:
                    worst = 0.0
                    t0 = time.perf_counter()
                    for f in range(frames):
                        ctx.set_instance_clocks(np.array([f * 1000.0 / 30.0], np.float32))
                        ctx.deform()
                        if f % 15 == 0 or f == frames - 1:
                            lr = crowd.tween_pose_batch(qa, qb, np.array([min(1.0, f / 45.0)]))
                            world = crowd.world_matrices_batch(bones, lr)
                            rp, rn = orc.deform(z["vtx8"], z["joints"], z["weights"], orc.skin_matrices(world[0], z["invBind"]))
                            gp, gn = ctx.read_instance(0)
                            worst = max(worst, rel_err(gp, rp), rel_err(gn, rn))
                    ctx.sync()
                    wall = (time.perf_counter() - t0) / frames * 1e3
                    ms = timed(ctx, stream, lambda: (ctx.set_instance_clocks(np.array([700.0], np.float32)), ctx.deform()))
                    emit(dict(config=name, V=V, B=B, K=1, frame_ms_device=ms, frame_ms_wall_with_checks=wall, verts_per_s=V / ms * 1e3, max_rel_err=worst,
                              stats=ctx.stats()))
                else:
                    phase = (np.arange(K) * synth.GOLDEN) % 1.0
                    clk = (phase * 1500.0).astype(np.float32)
                    ms = timed(ctx, stream, lambda: (ctx.set_instance_clocks(clk), ctx.deform()))
                    worst = 0.0
                    for k in (0, 333, 1023):
                        lr = crowd.tween_pose_batch(qa, qb, np.array([phase[k]]))
                        world = crowd.world_matrices_batch(bones, lr)
                        rp, rn = orc.deform(z["vtx8"], z["joints"], z["weights"], orc.skin_matrices(world[0], z["invBind"]))
                        gp, gn = ctx.read_instance(k)
                        worst = max(worst, rel_err(gp, rp), rel_err(gn, rn))
                    s = ctx.stats()
                    emit(dict(config=name, V=V, B=B, K=K, frame_ms_device=ms, verts_per_s=K * V / ms * 1e3, algorithmic_GBs=s["algorithmicBytes"] / ms / 1e6,
                              max_rel_err=worst, stats=s))
    else:
        emit(dict(config="config1/2", skipped="tests/golden/_local not present"))

    # ---- config 3: high-poly mesh, 64 active vertex morphs, SDEF on
    wl = synth.make_workload(200_000, 512, M=64, sdef=True)
    K = 256
    rng = np.random.default_rng(3)
    world = synth.make_palettes(wl.bones, K, rng)
    mw = rng.uniform(0, 1, (K, 64)).astype(np.float32)
    dw = torch.from_numpy(world).cuda()
    with capi.DeformContext(max_instances=K, stream=sh, flags=capi.RZ_FLAG_SDEF) as ctx:
        ctx.load_mesh(wl.vtx8, wl.joints, wl.weights, wl.invBind)
        ctx.load_morphs(wl.morphs.offsets, wl.morphs.vertexIndex, wl.morphs.delta)
        ctx.load_sdef(wl.sdef.vertexIndex, wl.sdef.c_r0_r1)
        ctx.set_palettes_device(dw.data_ptr(), K)
        ctx.set_morph_weights(mw, np.arange(64), K=K)
        ms = timed(ctx, stream, ctx.deform)
        worst = 0.0
        for k in (0, 100, 255):
            rp, rn = orc.deform(wl.vtx8, wl.joints, wl.weights, orc.skin_matrices(world[k], wl.invBind),
                                morph=(wl.morphs.offsets, wl.morphs.vertexIndex, wl.morphs.delta), morphW=mw[k], sdef=(wl.sdef.vertexIndex, wl.sdef.c_r0_r1))
            gp, gn = ctx.read_instance(k)
            worst = max(worst, rel_err(gp, rp), rel_err(gn, rn))
        s = ctx.stats()
        emit(dict(config="config3: V=200k, B=512, M=64 morphs (nnz=%d), SDEF on (%d verts), K=256" % (s["morphNnz"], s["sdefCount"]), frame_ms_device=ms,
                  verts_per_s=K * wl.V / ms * 1e3, algorithmic_GBs=s["algorithmicBytes"] / ms / 1e6, max_rel_err=worst, stats=s))

    # ---- config 4, one GPU's share: V=100k, K=8192 instances (19.7 GB of output)
    wl = synth.make_workload(100_000, 512)
    K, P = 8192, 1024
    world = synth.make_palettes(wl.bones, P, np.random.default_rng(4))
    dw = torch.from_numpy(world).cuda()
    i2p = (torch.arange(K, dtype=torch.int32, device="cuda") % P).contiguous()
    with capi.DeformContext(max_instances=K, stream=sh) as ctx:
        ctx.load_mesh(wl.vtx8, wl.joints, wl.weights, wl.invBind)
        ctx.set_palettes_device(dw.data_ptr(), P, i2p.data_ptr(), K)
        ms = timed(ctx, stream, ctx.deform, iters=5)
        worst = 0.0
        for k in (0, 4097, 8191):
            rp, rn = orc.deform(wl.vtx8, wl.joints, wl.weights, orc.skin_matrices(world[k % P], wl.invBind))
            gp, gn = ctx.read_instance(k)
            worst = max(worst, rel_err(gp, rp), rel_err(gn, rn))
        s = ctx.stats()
        emit(dict(config="config4 (one GPU's share): V=100k, B=512, K=8192, P=1024", frame_ms_device=ms, verts_per_s=K * wl.V / ms * 1e3,
                  algorithmic_GBs=s["algorithmicBytes"] / ms / 1e6, max_rel_err=worst, stats=s))


if __name__ == "__main__":
    main()
