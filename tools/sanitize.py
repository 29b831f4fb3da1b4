"""Small deform runs for compute-sanitizer (memcheck / racecheck / synccheck): every feature path on tiny shapes."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from reze_engine_b200 import capi, synth  # noqa: E402

wl = synth.make_workload(1500, 40, M=6, sdef=True)
rng = np.random.default_rng(1)
K = 7
world = synth.make_palettes(wl.bones, K, rng)
mw = rng.uniform(0, 1, (K, 6)).astype(np.float32)
edge = rng.uniform(0, 2, wl.V).astype(np.float32)
for flags, I, nt in ((0, 2, 256), (capi.RZ_FLAG_SDEF | capi.RZ_FLAG_BOUNDS, 4, 512), (capi.RZ_FLAG_NO_NORMALS, 1, 256), (0, 2, 512),
                     (capi.RZ_FLAG_SDEF, 2, 256), (capi.RZ_FLAG_SDEF | capi.RZ_FLAG_BOUNDS | capi.RZ_FLAG_OUTLINE, 2, 256),
                     (capi.RZ_FLAG_SDEF | capi.RZ_FLAG_INTERLEAVED, 4, 512), (capi.RZ_FLAG_OUTLINE, 1, 256), (capi.RZ_FLAG_INTERLEAVED, 2, 512)):
    with capi.DeformContext(max_instances=K, flags=flags, instances_per_group=I, threads=nt) as ctx:
        ctx.load_mesh(wl.vtx8, wl.joints, wl.weights, wl.invBind)
        ctx.load_morphs(wl.morphs.offsets, wl.morphs.vertexIndex, wl.morphs.delta)
        ctx.load_sdef(wl.sdef.vertexIndex, wl.sdef.c_r0_r1)
        ctx.load_skeleton(wl.bones)
        if flags & capi.RZ_FLAG_OUTLINE:
            ctx.load_edge_size(edge)
        ctx.set_palettes(world)
        ctx.set_morph_weights(mw, np.arange(6), K=K)
        ctx.deform()
        ctx.sync()
        qa, qb, ph = synth.make_crowd_tween(wl.B, K, rng)
        ident = np.tile(np.array([0, 0, 0, 1], np.float32), (wl.B, 1))
        ctx.set_tweens(qa, qb, np.zeros(wl.B, np.float32), np.full(wl.B, 1000.0, np.float32), np.ones(wl.B, np.uint8), ident)
        ctx.set_instance_clocks((ph * 1000).astype(np.float32))
        ctx.deform()
        ctx.sync()
        print("ok", flags, ctx.stats()["instancesPerGroup"], ctx.stats()["threads"])

# ---- round 2: the two-vertices-per-lane kernel of the plain path (sub-batch staging ring, two-level item plan, ragged tail),
# launched through the frame's CUDA graph and directly
for I, nt, sb in ((4, 512, 2), (6, 384, 2), (6, 512, 1), (3, 256, 3), (1, 256, 1)):
    for Vx in (1500, 67):
        wx = synth.make_workload(Vx, 40)
        with capi.DeformContext(max_instances=K, instances_per_group=I, threads=nt, store_mode=sb, vertices_per_lane=2) as ctx:
            ctx.load_mesh(wx.vtx8, wx.joints, wx.weights, wx.invBind)
            ctx.set_palettes(synth.make_palettes(wx.bones, K, rng))
            ctx.deform()
            ctx.deform(2, 3)
            ctx.sync()
            assert ctx.stats()["verticesPerLane"] == 2
    print("ok two-vertex kernel", I, nt, sb)
# ... and in its other output layouts (positions only, outline hull, interleaved with the half-swapped store order, planar + AABB)
for flags in (capi.RZ_FLAG_NO_NORMALS, capi.RZ_FLAG_OUTLINE, capi.RZ_FLAG_INTERLEAVED, capi.RZ_FLAG_BOUNDS):
    for I, nt, sb in ((0, 0, 0), (4, 384, 2), (1, 256, 1)):
        wx = synth.make_workload(1003, 40)
        with capi.DeformContext(max_instances=K, flags=flags, instances_per_group=I, threads=nt, store_mode=sb, vertices_per_lane=2) as ctx:
            ctx.load_mesh(wx.vtx8, wx.joints, wx.weights, wx.invBind)
            if flags & capi.RZ_FLAG_OUTLINE:
                ctx.load_edge_size(rng.uniform(0, 2, wx.V).astype(np.float32))
            ctx.set_palettes(synth.make_palettes(wx.bones, K, rng))
            ctx.deform()
            ctx.deform(1, 5)
            ctx.sync()
            assert ctx.stats()["verticesPerLane"] == 2
    print("ok two-vertex kernel, layout flags", flags)

# ---- the round's later additions: pointer-jumping / chain / level pose kernels, pipelined palette upload, double-buffered
# results with asynchronous read-back, physics feedback, SDEF + outline on a palette that does not fit shared memory
import torch  # noqa: E402
from reze_engine_b200 import physics_bridge as pb  # noqa: E402

os.environ.setdefault("RZ_POSE", "2")               # pointer jumping (the default); run again with RZ_POSE=1 / 0 for the fallbacks
os.environ["RZ_PIPELINE_BLOCK"] = "2"
pin = lambda: torch.empty((wl.V, 3), dtype=torch.float32, pin_memory=True).numpy()
hp, hn = pin(), pin()
with capi.DeformContext(max_instances=K, flags=capi.RZ_FLAG_DOUBLE_BUFFER) as ctx:
    ctx.load_mesh(wl.vtx8, wl.joints, wl.weights, wl.invBind)
    ctx.load_skeleton(wl.bones)
    n = 9
    bone_index = rng.integers(-1, wl.B, n).astype(np.int32)
    dyn = np.ones(n, np.uint8)
    _, inv = pb.compute_body_offsets(np.asarray(wl.invBind, np.float32).reshape(-1, 16), bone_index, rng.normal(0, 2, (n, 3)), rng.uniform(-1, 1, (n, 3)))
    ctx.load_rigid_bodies(bone_index, dyn, inv)
    pq = np.concatenate([rng.normal(0, 3, (K, n, 3)), np.tile([0, 0, 0, 1.0], (K, n, 1))], axis=2).astype(np.float32)
    for f in range(3):
        ctx.set_palettes(world)                       # pipelined in blocks of 2 palettes
        ctx.apply_body_transforms(pq)
        ctx.deform(0, 3)
        ctx.deform(3, K - 3)
        ctx.read_wait()
        ctx.read_instance_async(f, hp, hn)
    ctx.read_wait()
    lr = np.tile(np.array([0, 0, 0, 1], np.float32), (K, wl.B, 1))
    ctx.set_local_rotations(lr)
    ctx.apply_body_transforms(pq)
    ctx.deform()
    qa, qb, ph = synth.make_crowd_tween(wl.B, K, rng)
    ident = np.tile(np.array([0, 0, 0, 1], np.float32), (wl.B, 1))
    ctx.set_tweens(qa, qb, np.zeros(wl.B, np.float32), np.full(wl.B, 1000.0, np.float32), np.ones(wl.B, np.uint8), ident)
    ctx.set_instance_clocks((ph * 1000).astype(np.float32))
    ctx.deform()
    ctx.sync()
    print("ok pose/pipeline/double-buffer/physics", ctx.stats()["kernelLaunches"])
big = synth.make_workload(900, 5000, M=3, sdef=True, seed=3)
wb = synth.make_palettes(big.bones, 3, rng)
with capi.DeformContext(max_instances=3, flags=capi.RZ_FLAG_SDEF | capi.RZ_FLAG_OUTLINE | capi.RZ_FLAG_BOUNDS) as ctx:
    ctx.load_mesh(big.vtx8, big.joints, big.weights, big.invBind)
    ctx.load_morphs(big.morphs.offsets, big.morphs.vertexIndex, big.morphs.delta)
    ctx.load_sdef(big.sdef.vertexIndex, big.sdef.c_r0_r1)
    ctx.load_edge_size(rng.uniform(0, 2, big.V).astype(np.float32))
    ctx.set_palettes(wb)
    ctx.set_morph_weights(rng.uniform(0, 1, (3, 3)).astype(np.float32), np.arange(3), K=3)
    ctx.deform()
    ctx.sync()
    print("ok global-palette sdef+outline+bounds", ctx.stats()["threads"])
