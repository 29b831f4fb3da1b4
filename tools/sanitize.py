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
