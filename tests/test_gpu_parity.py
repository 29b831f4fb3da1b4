"""GPU parity (-m gpu): the CUDA path through the C ABI vs the CPU oracle on the same inputs.

Tolerance (SURVEY 8c / BASELINE.json): floats  max_v |out-ref|_inf / max(|ref|_inf, 1) <= 1e-5 against the f32
oracle; integer tables bit-exact.
"""
import os

import numpy as np
import pytest

from helpers import random_pmx, rel_err, write_vmd
from reze_engine_b200 import Engine, PmxLoader, Quat, VMDLoader, capi, crowd, synth
from reze_engine_b200.engine import ManualClock
from reze_engine_b200.model import Bone

pytestmark = pytest.mark.gpu
TOL = 1e-5
LOCAL = os.path.join(os.path.dirname(__file__), "golden", "_local")


def oracle_instance(orc, wl, world_p, morphW=None, sdef=False):
    skin = orc.skin_matrices(world_p, wl.invBind)
    morph = (wl.morphs.offsets, wl.morphs.vertexIndex, wl.morphs.delta) if morphW is not None else None
    sd = (wl.sdef.vertexIndex, wl.sdef.c_r0_r1) if sdef else None
    return orc.deform(wl.vtx8, wl.joints, wl.weights, skin, morph=morph, morphW=morphW, sdef=sd)


def check_all(orc, ctx, wl, world, i2p, K, morphW=None, sdef=False, instances=None):
    for k in (range(K) if instances is None else instances):
        p = k if i2p is None else int(i2p[k])
        rp, rn = oracle_instance(orc, wl, world[p], None if morphW is None else morphW[k], sdef)
        gp, gn = ctx.read_instance(k)
        assert rel_err(gp, rp) <= TOL, (k, rel_err(gp, rp))
        assert rel_err(gn, rn) <= TOL, (k, rel_err(gn, rn))


@pytest.fixture(scope="module")
def wl_small():
    return synth.make_workload(5000, 64, M=8, sdef=True)


SHAPES = [(1, 256, 4), (2, 256, 3), (2, 256, 2), (3, 256, 2), (4, 256, 1), (1, 512, 2), (2, 512, 2), (2, 512, 1), (3, 512, 1),
          (4, 512, 1), (2, 768, 1), (3, 768, 1), (4, 768, 1), (2, 1024, 1), (3, 1024, 1), (6, 256, 1), (6, 512, 1)]     # csrc/kernel_table.h RZ_SHAPES_FULL


@pytest.mark.parametrize("I,nt,mb", SHAPES)
def test_every_launch_shape_matches_oracle(rzlib, orc, wl_small, I, nt, mb):
    wl = wl_small
    K, P = 11, 4                                    # partial last group for every I; shared palettes
    world = synth.make_palettes(wl.bones, P, np.random.default_rng(1))
    i2p = (np.arange(K) * 3) % P
    with capi.DeformContext(max_instances=K, instances_per_group=I, threads=nt, ctas_per_sm=mb, vertices_per_lane=1) as ctx:
        ctx.load_mesh(wl.vtx8, wl.joints, wl.weights, wl.invBind)
        ctx.set_palettes(world, i2p)
        ctx.deform()
        s = ctx.stats()
        assert (s["instancesPerGroup"], s["threads"], s["verticesPerLane"]) == (I, nt, 1)
        check_all(orc, ctx, wl, world, i2p, K)
        # instances that share a palette are bit-identical
        a, b = ctx.read_instance(0), ctx.read_instance(4)
        assert i2p[0] == i2p[4] and np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


SHAPES_V2 = [(6, 384, 1, 2), (6, 512, 1, 1), (6, 256, 1, 3), (6, 256, 1, 2), (5, 512, 1, 1), (4, 384, 1, 2), (4, 512, 1, 2), (4, 512, 1, 1),
             (4, 256, 1, 2), (4, 256, 1, 4), (3, 256, 2, 3), (3, 256, 1, 3), (3, 512, 1, 1), (2, 256, 2, 2), (2, 256, 1, 2), (2, 512, 1, 1),
             (1, 256, 2, 1)]                                                                                    # csrc/kernel_table.h RZ_SHAPES_V2


@pytest.mark.parametrize("I,nt,mb,sb", SHAPES_V2)
def test_every_two_vertex_launch_shape_matches_oracle(rzlib, orc, wl_small, I, nt, mb, sb):
    """The plain path's default kernel (two vertices per lane, deform2_kernel.cuh) in every compiled shape: partial last
    instance group, shared palettes, sub-range launches, and bit-identical to itself across instances sharing a palette.
    Against the one-vertex kernel the results differ at most by the order of the <= 4 blend terms."""
    wl = wl_small
    K, P = 11, 4
    world = synth.make_palettes(wl.bones, P, np.random.default_rng(1))
    i2p = (np.arange(K) * 3) % P
    with capi.DeformContext(max_instances=K, instances_per_group=I, threads=nt, ctas_per_sm=mb, store_mode=sb, vertices_per_lane=2) as ctx:
        ctx.load_mesh(wl.vtx8, wl.joints, wl.weights, wl.invBind)
        j, w = ctx.read_skinning()                   # decodes BOTH device tables and insists that they agree
        assert np.array_equal(j, wl.joints.reshape(-1)) and np.array_equal(w, wl.weights.reshape(-1))
        ctx.set_palettes(world, i2p)
        ctx.deform()
        s = ctx.stats()
        assert (s["instancesPerGroup"], s["threads"], s["verticesPerLane"]) == (I, nt, 2)
        check_all(orc, ctx, wl, world, i2p, K)
        a, b = ctx.read_instance(0), ctx.read_instance(4)
        assert i2p[0] == i2p[4] and np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        ref = [ctx.read_instance(k) for k in range(K)]
        ctx.set_palettes(world[::-1].copy(), i2p)     # other poses, then only a sub-range is re-deformed with the first ones
        ctx.deform()
        ctx.set_palettes(world, i2p)
        ctx.deform(3, 5)
        for k in range(3, 8):
            got = ctx.read_instance(k)
            assert np.array_equal(got[0], ref[k][0]) and np.array_equal(got[1], ref[k][1])
    with capi.DeformContext(max_instances=K, vertices_per_lane=1) as one:
        one.load_mesh(wl.vtx8, wl.joints, wl.weights, wl.invBind)
        one.set_palettes(world, i2p)
        one.deform()
        assert one.stats()["verticesPerLane"] == 1
        for k in (0, 10):
            g1 = one.read_instance(k)
            assert rel_err(ref[k][0], g1[0]) <= 2e-6 and rel_err(ref[k][1], g1[1]) <= 2e-6


@pytest.mark.parametrize("V", [1, 3, 31, 33, 63, 64, 65, 127, 129, 255, 256, 257, 1001, 1022])
def test_ragged_vertex_counts(rzlib, orc, V):
    wl = synth.make_workload(V, 16, seed=100 + V)
    world = synth.make_palettes(wl.bones, 3, np.random.default_rng(2))
    for I, nt in ((0, 0), (1, 256), (2, 512), (2, 768)):
        with capi.DeformContext(max_instances=3, instances_per_group=I, threads=nt) as ctx:
            ctx.load_mesh(wl.vtx8, wl.joints, wl.weights, wl.invBind)
            ctx.set_palettes(world)
            ctx.deform()
            check_all(orc, ctx, wl, world, None, 3)


def test_integer_tables_bit_exact_and_skin_matrices(rzlib, orc, wl_small):
    wl = wl_small
    world = synth.make_palettes(wl.bones, 2, np.random.default_rng(3))
    with capi.DeformContext(max_instances=2) as ctx:
        ctx.load_mesh(wl.vtx8, wl.joints, wl.weights, wl.invBind)
        j, w = ctx.read_skinning()
        assert np.array_equal(j, wl.joints.reshape(-1)) and np.array_equal(w, wl.weights.reshape(-1))
        ctx.set_palettes(world)
        got = ctx.read_skin_matrices(1)                       # [B,12] rows
        ref = orc.skin_matrices(world[1], wl.invBind).reshape(-1, 4, 4).transpose(0, 2, 1)[:, :3, :].reshape(-1, 12)
        assert rel_err(got, ref) <= 1e-6


def test_tpose_identity(rzlib, wl_small):
    wl = wl_small
    ident = np.tile(np.array([0, 0, 0, 1], np.float32), (1, wl.B, 1))
    world = crowd.world_matrices_batch(wl.bones, ident)
    with capi.DeformContext(max_instances=1) as ctx:
        ctx.load_mesh(wl.vtx8, wl.joints, wl.weights, wl.invBind)
        ctx.set_palettes(world)
        ctx.deform()
        p, n = ctx.read_instance(0)
        assert rel_err(p, wl.vtx8[:, :3]) <= 1e-6 and rel_err(n, wl.vtx8[:, 3:6]) <= 1e-6


def test_morphs(rzlib, orc, wl_small):
    wl = wl_small
    K = 6
    rng = np.random.default_rng(4)
    world = synth.make_palettes(wl.bones, K, rng)
    active = np.array([5, 0, 3], np.uint32)
    w = rng.uniform(0, 1, (K, 3)).astype(np.float32)
    dense = np.zeros((K, wl.morphs.count), np.float32)
    dense[:, active] = w
    for I, nt in ((1, 256), (4, 512), (2, 256), (2, 512)):
        with capi.DeformContext(max_instances=K, instances_per_group=I, threads=nt) as ctx:
            ctx.load_mesh(wl.vtx8, wl.joints, wl.weights, wl.invBind)
            ctx.load_morphs(wl.morphs.offsets, wl.morphs.vertexIndex, wl.morphs.delta)
            ctx.set_palettes(world)
            ctx.deform()                                    # no weights yet: the pinned (M = 0) path
            check_all(orc, ctx, wl, world, None, K)
            ctx.set_morph_weights(w, active, K=K)
            ctx.deform()
            assert ctx.stats()["activeMorphs"] == 3
            check_all(orc, ctx, wl, world, None, K, morphW=dense)
            ctx.set_morph_weights(None, [], K=K)            # back to zero weights
            ctx.deform()
            check_all(orc, ctx, wl, world, None, K)


def test_morph_row_formats(rzlib, orc):
    """The per-warp morph rows come in two formats chosen at load time (morph-major when the warp's vertices share their
    morphs, compact otherwise); both must equal the oracle, including a morph that lists one vertex twice and morphs
    scattered over random vertices."""
    rng = np.random.default_rng(90)
    wl = synth.make_workload(3000, 40, seed=90)
    V, M = wl.V, 48
    offs, vidx, deltas = [0], [], []
    for m in range(M):
        if m % 3 == 0:                              # coherent block (morph-major rows)
            a = int(rng.integers(0, V - 400))
            v = np.arange(a, a + int(rng.integers(50, 400)))
        else:                                       # scattered (compact rows in most warps)
            v = np.sort(rng.choice(V, int(rng.integers(5, 200)), replace=False))
        if m == 7:
            v = np.concatenate([v, v[:3]])          # this morph lists three vertices twice
        vidx.append(v.astype(np.uint32))
        deltas.append(rng.normal(0, 0.05, (v.size, 3)).astype(np.float32))
        offs.append(offs[-1] + v.size)
    offs = np.array(offs, np.uint32)
    vidx = np.concatenate(vidx)
    deltas = np.concatenate(deltas)
    K = 5
    world = synth.make_palettes(wl.bones, K, rng)
    w = rng.uniform(-0.5, 1.0, (K, M)).astype(np.float32)
    for I, nt in ((0, 0), (1, 256), (2, 256), (4, 512)):
        with capi.DeformContext(max_instances=K, instances_per_group=I, threads=nt) as ctx:
            ctx.load_mesh(wl.vtx8, wl.joints, wl.weights, wl.invBind)
            ctx.load_morphs(offs, vidx, deltas)
            ctx.set_palettes(world)
            ctx.set_morph_weights(w, np.arange(M), K=K)
            ctx.deform()
            assert ctx.stats()["morphNnz"] == vidx.size
            for k in range(K):
                rp, rn = orc.deform(wl.vtx8, wl.joints, wl.weights, orc.skin_matrices(world[k], wl.invBind), morph=(offs, vidx, deltas), morphW=w[k])
                gp, gn = ctx.read_instance(k)
                assert rel_err(gp, rp) <= TOL and rel_err(gn, rn) <= TOL, (I, k, rel_err(gp, rp))


def test_sdef_on_and_compat_off(rzlib, orc, wl_small):
    wl = wl_small
    K = 5
    rng = np.random.default_rng(5)
    world = synth.make_palettes(wl.bones, K, rng)
    dense = rng.uniform(0, 1, (K, wl.morphs.count)).astype(np.float32)
    with capi.DeformContext(max_instances=K, flags=capi.RZ_FLAG_SDEF) as ctx:
        ctx.load_mesh(wl.vtx8, wl.joints, wl.weights, wl.invBind)
        ctx.load_sdef(wl.sdef.vertexIndex, wl.sdef.c_r0_r1)
        ctx.set_palettes(world)
        ctx.deform()
        assert ctx.stats()["sdefCount"] == wl.sdef.vertexIndex.size
        check_all(orc, ctx, wl, world, None, K, sdef=True)
        ctx.load_morphs(wl.morphs.offsets, wl.morphs.vertexIndex, wl.morphs.delta)
        ctx.set_morph_weights(dense, np.arange(wl.morphs.count), K=K)
        ctx.deform()
        check_all(orc, ctx, wl, world, None, K, morphW=dense, sdef=True)
    with capi.DeformContext(max_instances=K) as ctx:          # compat mode: SDEF records ignored => BDEF2 (reference behaviour)
        ctx.load_mesh(wl.vtx8, wl.joints, wl.weights, wl.invBind)
        ctx.load_sdef(wl.sdef.vertexIndex, wl.sdef.c_r0_r1)
        ctx.set_palettes(world)
        ctx.deform()
        check_all(orc, ctx, wl, world, None, K, sdef=False)


def _all_sdef(wl, rng):
    """Every vertex that may be SDEF (third and fourth weight zero) becomes SDEF: up to 32 per warp, so the kernel's dense
    phase needs several rounds per pass (32 / I pairs fit one round)."""
    W = wl.weights.reshape(-1, 4)
    idx = np.nonzero((W[:, 2] == 0) & (W[:, 3] == 0))[0].astype(np.uint32)
    pos = wl.vtx8.reshape(-1, 8)[idx, :3]
    C = pos + rng.normal(0, 0.1, pos.shape)
    vec = np.concatenate([C, C + rng.normal(0, 0.2, pos.shape), C - rng.normal(0, 0.2, pos.shape)], axis=1).astype(np.float32)
    return idx, vec


@pytest.mark.parametrize("I,nt", [(1, 256), (2, 256), (2, 512), (4, 512)])
def test_sdef_dense_phase_many_rounds_and_feature_sets(rzlib, orc, wl_small, I, nt):
    wl = wl_small
    K = 5                                           # partial last group for I = 2, 4
    rng = np.random.default_rng(50 + I)
    world = synth.make_palettes(wl.bones, K, rng)
    world[1] = world[0]                             # two instances, one pose
    dense = rng.uniform(0, 1, (K, wl.morphs.count)).astype(np.float32)
    dense[1] = dense[0]
    idx, vec = _all_sdef(wl, rng)
    assert idx.size > wl.V // 2
    sd = (idx, vec)
    for flags in (0, capi.RZ_FLAG_BOUNDS, capi.RZ_FLAG_NO_NORMALS):
        with capi.DeformContext(max_instances=K, flags=capi.RZ_FLAG_SDEF | flags, instances_per_group=I, threads=nt) as ctx:
            ctx.load_mesh(wl.vtx8, wl.joints, wl.weights, wl.invBind)
            ctx.load_sdef(idx, vec)
            ctx.load_morphs(wl.morphs.offsets, wl.morphs.vertexIndex, wl.morphs.delta)
            ctx.set_palettes(world)
            ctx.set_morph_weights(dense, np.arange(wl.morphs.count), K=K)
            ctx.deform()
            s = ctx.stats()
            assert (s["instancesPerGroup"], s["threads"], s["sdefCount"]) == (I, nt, idx.size)
            for k in range(K):
                rp, rn = orc.deform(wl.vtx8, wl.joints, wl.weights, orc.skin_matrices(world[k], wl.invBind),
                                    morph=(wl.morphs.offsets, wl.morphs.vertexIndex, wl.morphs.delta), morphW=dense[k], sdef=sd)
                gp, gn = ctx.read_instance(k, normals=not (flags & capi.RZ_FLAG_NO_NORMALS))
                assert rel_err(gp, rp) <= TOL, (flags, k, rel_err(gp, rp))
                if gn is not None:
                    assert rel_err(gn, rn) <= TOL, (flags, k, rel_err(gn, rn))
                if flags & capi.RZ_FLAG_BOUNDS:
                    bb = ctx.read_bounds(k, 1)[0]
                    assert np.array_equal(bb[:3], gp.min(axis=0)) and np.array_equal(bb[3:], gp.max(axis=0))
            a, b = ctx.read_instance(0, normals=False), ctx.read_instance(1, normals=False)
            assert np.array_equal(a[0], b[0])


def test_sdef_edge_cases(rzlib, orc):
    """Bind pose (all quaternions equal: the lerp branch of slerp), one-bone SDEF records, the same bone twice, and the
    palette-in-global path (B too large for shared memory)."""
    rng = np.random.default_rng(61)
    wl = synth.make_workload(2000, 24, seed=61, sdef=True)
    W = wl.weights.reshape(-1, 4).copy()
    J = wl.joints.reshape(-1, 4).copy()
    idx, vec = _all_sdef(wl, rng)
    W[idx[0]] = [255, 0, 0, 0]                      # w1 = 0: slerp at t = 0
    W[idx[1]] = [0, 255, 0, 0]                      # w0 = 0
    J[idx[2], 1] = J[idx[2], 0]                     # the same bone in both slots
    wl.weights, wl.joints = W, J
    ident = np.tile(np.eye(4, dtype=np.float32).T.reshape(1, 1, 16), (1, wl.B, 1))
    bind = ident.copy()
    bind[0, :, 12:15] = -wl.invBind.reshape(-1, 16)[:, 12:15]        # world = inverse of the (pure translation) inverse bind
    world = np.concatenate([bind, synth.make_palettes(wl.bones, 2, rng)], axis=0)
    with capi.DeformContext(max_instances=3, flags=capi.RZ_FLAG_SDEF) as ctx:
        ctx.load_mesh(wl.vtx8, J, W, wl.invBind)
        ctx.load_sdef(idx, vec)
        ctx.set_palettes(world)
        ctx.deform()
        for k in range(3):
            rp, rn = orc.deform(wl.vtx8, J, W, orc.skin_matrices(world[k], wl.invBind), sdef=(idx, vec))
            gp, gn = ctx.read_instance(k)
            assert rel_err(gp, rp) <= TOL and rel_err(gn, rn) <= TOL, (k, rel_err(gp, rp), rel_err(gn, rn))
        gp, _ = ctx.read_instance(0)                # bind pose: SDEF is the identity too
        assert np.abs(gp - wl.vtx8.reshape(-1, 8)[:, :3]).max() <= 2e-5
    big = synth.make_workload(3000, 6000, seed=78)
    idx, vec = _all_sdef(big, rng)
    world = synth.make_palettes(big.bones, 3, rng)
    with capi.DeformContext(max_instances=3, flags=capi.RZ_FLAG_SDEF) as ctx:
        ctx.load_mesh(big.vtx8, big.joints, big.weights, big.invBind)
        ctx.load_sdef(idx, vec)
        ctx.set_palettes(world, np.array([2, 0, 1], np.uint32))
        ctx.deform()
        for k, p in enumerate((2, 0, 1)):
            rp, rn = orc.deform(big.vtx8, big.joints, big.weights, orc.skin_matrices(world[p], big.invBind), sdef=(idx, vec))
            gp, gn = ctx.read_instance(k)
            assert rel_err(gp, rp) <= TOL and rel_err(gn, rn) <= TOL, (k, rel_err(gp, rp), rel_err(gn, rn))


@pytest.mark.parametrize("full", [False, True])
def test_outline_hull_plane(rzlib, orc, wl_small, full):
    """RZ_FLAG_OUTLINE: third plane = pos' + n' * edgeSize * 0.01 (engine.ts:458-461), fused into the deform pass; plain and with
    morphs + SDEF + bounds; positions / normals are untouched by the extra plane."""
    wl = wl_small
    K = 5
    rng = np.random.default_rng(70)
    world = synth.make_palettes(wl.bones, K, rng)
    edge = np.where(rng.uniform(size=wl.V) < 0.7, rng.uniform(0.2, 2.0, wl.V), 0.0).astype(np.float32)
    dense = rng.uniform(0, 1, (K, wl.morphs.count)).astype(np.float32) if full else None
    flags = capi.RZ_FLAG_OUTLINE | ((capi.RZ_FLAG_SDEF | capi.RZ_FLAG_BOUNDS) if full else 0)
    for I, nt in ((0, 0), (1, 256), (2, 512)):
        with capi.DeformContext(max_instances=K, flags=flags, instances_per_group=I, threads=nt) as ctx:
            ctx.load_mesh(wl.vtx8, wl.joints, wl.weights, wl.invBind)
            if full:
                ctx.load_morphs(wl.morphs.offsets, wl.morphs.vertexIndex, wl.morphs.delta)
                ctx.load_sdef(wl.sdef.vertexIndex, wl.sdef.c_r0_r1)
            ctx.set_palettes(world)
            if full:
                ctx.set_morph_weights(dense, np.arange(wl.morphs.count), K=K)
            ctx.deform()                                    # no edge sizes loaded: hull == pos'
            gp, _ = ctx.read_instance(2)
            assert np.array_equal(ctx.read_outline(2), gp)
            ctx.load_edge_size(edge)
            ctx.deform()
            lay = ctx.output_layout()
            assert lay["vertexStride"] == 12 and lay["hullOffset"] == 2 * lay["normalOffset"] and lay["uvOffset"] == capi.RZ_NO_ATTRIBUTE
            assert lay["instanceStride"] >= 3 * wl.V * 12
            for k in range(K):
                rp, rn = oracle_instance(orc, wl, world[k], None if dense is None else dense[k], sdef=full)
                gp, gn = ctx.read_instance(k)
                gh = ctx.read_outline(k)
                assert rel_err(gp, rp) <= TOL and rel_err(gn, rn) <= TOL
                assert rel_err(gh, orc.outline_hull(rp, rn, edge)) <= TOL, (k, rel_err(gh, orc.outline_hull(rp, rn, edge)))
                # against its own position / normal the hull is exact up to one fma rounding
                assert np.abs(gh - (gp + gn * (edge * np.float32(0.01))[:, None])).max() <= 4e-6
                assert np.array_equal(gh[edge == 0], gp[edge == 0])
    with pytest.raises(capi.RzError):
        capi.DeformContext(max_instances=1, flags=capi.RZ_FLAG_OUTLINE | capi.RZ_FLAG_NO_NORMALS)
    with pytest.raises(capi.RzError):
        capi.DeformContext(max_instances=1, flags=capi.RZ_FLAG_OUTLINE | capi.RZ_FLAG_INTERLEAVED)


@pytest.mark.parametrize("full", [False, True])
def test_interleaved_vertex_stream(rzlib, orc, wl_small, full):
    """RZ_FLAG_INTERLEAVED: the result leaves in the reference's vertex-buffer layout (8 f32 per vertex, engine.ts:340-347):
    bit-identical to the planar result, uv passed through; ragged vertex counts included."""
    rng = np.random.default_rng(71)
    for V in (wl_small.V, 1001, 33):
        wl = wl_small if V == wl_small.V else synth.make_workload(V, 20, M=4, sdef=True, seed=V)
        K = 5
        world = synth.make_palettes(wl.bones, K, rng)
        dense = rng.uniform(0, 1, (K, wl.morphs.count)).astype(np.float32) if full else None
        extra = (capi.RZ_FLAG_SDEF | capi.RZ_FLAG_BOUNDS) if full else 0
        outs = {}
        for flags in (0, capi.RZ_FLAG_INTERLEAVED):
            with capi.DeformContext(max_instances=K, flags=flags | extra, instances_per_group=2, threads=256, vertices_per_lane=1) as ctx:   # same one-vertex kernel arithmetic in both layouts: comparable bit for bit
                ctx.load_mesh(wl.vtx8, wl.joints, wl.weights, wl.invBind)
                if full:
                    ctx.load_morphs(wl.morphs.offsets, wl.morphs.vertexIndex, wl.morphs.delta)
                    ctx.load_sdef(wl.sdef.vertexIndex, wl.sdef.c_r0_r1)
                ctx.set_palettes(world)
                if full:
                    ctx.set_morph_weights(dense, np.arange(wl.morphs.count), K=K)
                ctx.deform()
                outs[flags] = [ctx.read_instance(k) for k in range(K)]
                if flags:
                    lay = ctx.output_layout()
                    assert (lay["vertexStride"], lay["normalOffset"], lay["uvOffset"], lay["instanceStride"]) == (32, 12, 24, wl.V * 32)
                    for k in range(K):
                        st = ctx.read_interleaved(k)
                        rp, rn = oracle_instance(orc, wl, world[k], None if dense is None else dense[k], sdef=full)
                        ref = orc.interleaved(rp, rn, wl.vtx8)
                        assert rel_err(st[:, :3], ref[:, :3]) <= TOL and rel_err(st[:, 3:6], ref[:, 3:6]) <= TOL
                        assert np.array_equal(st[:, 6:], ref[:, 6:])                      # uv: bit-exact pass-through
                        assert np.array_equal(st[:, :3], outs[flags][k][0]) and np.array_equal(st[:, 3:6], outs[flags][k][1])
                    if full:
                        bb = ctx.read_bounds(0, K)
                        for k in range(K):
                            assert np.array_equal(bb[k, :3], outs[flags][k][0].min(axis=0))
        for k in range(K):                                                                 # same arithmetic, different layout
            assert np.array_equal(outs[0][k][0], outs[capi.RZ_FLAG_INTERLEAVED][k][0])
            assert np.array_equal(outs[0][k][1], outs[capi.RZ_FLAG_INTERLEAVED][k][1])


@pytest.mark.parametrize("flags,morph,sdef", [
    (capi.RZ_FLAG_BOUNDS, True, False),                              # runs the morph+SDEF+bounds kernel without SDEF tables
    (capi.RZ_FLAG_OUTLINE, True, False),                             # ... the outline superset without RZ_FLAG_BOUNDS
    (capi.RZ_FLAG_INTERLEAVED | capi.RZ_FLAG_SDEF, True, True),      # ... the interleaved superset without RZ_FLAG_BOUNDS
    (capi.RZ_FLAG_SDEF, False, True),                                # SDEF without morphs
    (capi.RZ_FLAG_NO_NORMALS | capi.RZ_FLAG_BOUNDS, False, False),
])
def test_feature_supersets_run_clean(rzlib, orc, wl_small, flags, morph, sdef):
    """rz_deform launches the smallest COMPILED feature set that covers the request; the extra features of that kernel
    (SDEF phase without SDEF records, AABB reduction without RZ_FLAG_BOUNDS, ...) must be inert."""
    wl = wl_small
    K = 6
    rng = np.random.default_rng(int(flags) + 100)
    world = synth.make_palettes(wl.bones, K, rng)
    dense = rng.uniform(0, 1, (K, wl.morphs.count)).astype(np.float32) if morph else None
    for I, nt in ((0, 0), (2, 256), (4, 512)):
        with capi.DeformContext(max_instances=K, flags=flags, instances_per_group=I, threads=nt) as ctx:
            ctx.load_mesh(wl.vtx8, wl.joints, wl.weights, wl.invBind)
            if morph:
                ctx.load_morphs(wl.morphs.offsets, wl.morphs.vertexIndex, wl.morphs.delta)
            if sdef:
                ctx.load_sdef(wl.sdef.vertexIndex, wl.sdef.c_r0_r1)
            ctx.set_palettes(world)
            if morph:
                ctx.set_morph_weights(dense, np.arange(wl.morphs.count), K=K)
            ctx.deform()
            ctx.sync()
            for k in range(K):
                rp, rn = oracle_instance(orc, wl, world[k], None if dense is None else dense[k], sdef=sdef)
                gp, gn = ctx.read_instance(k, normals=not (flags & capi.RZ_FLAG_NO_NORMALS))
                assert rel_err(gp, rp) <= TOL, (I, k, rel_err(gp, rp))
                if gn is not None:
                    assert rel_err(gn, rn) <= TOL


@pytest.mark.parametrize("B", [48, 5000])            # palette in shared memory / too large for it (gathered from global)
def test_every_flag_combination_runs_and_matches(rzlib, orc, B):
    """Every combination of the public flags (output layout x bounds x SDEF, with and without active morphs) resolves to a
    compiled kernel and equals the oracle, for a palette that fits shared memory and for one that does not."""
    wl = synth.make_workload(700, B, M=5, sdef=True, seed=B)
    K = 3
    rng = np.random.default_rng(B)
    world = synth.make_palettes(wl.bones, K, rng)
    dense = rng.uniform(0, 1, (K, wl.morphs.count)).astype(np.float32)
    edge = rng.uniform(0, 2, wl.V).astype(np.float32)
    for layout in (0, capi.RZ_FLAG_NO_NORMALS, capi.RZ_FLAG_OUTLINE, capi.RZ_FLAG_INTERLEAVED):
        for bounds in (0, capi.RZ_FLAG_BOUNDS):
            for sd in (0, capi.RZ_FLAG_SDEF):
                for morph in (False, True):
                    flags = layout | bounds | sd
                    with capi.DeformContext(max_instances=K, flags=flags) as ctx:
                        ctx.load_mesh(wl.vtx8, wl.joints, wl.weights, wl.invBind)
                        ctx.load_sdef(wl.sdef.vertexIndex, wl.sdef.c_r0_r1)
                        if morph:
                            ctx.load_morphs(wl.morphs.offsets, wl.morphs.vertexIndex, wl.morphs.delta)
                        if layout == capi.RZ_FLAG_OUTLINE:
                            ctx.load_edge_size(edge)
                        ctx.set_palettes(world)
                        if morph:
                            ctx.set_morph_weights(dense, np.arange(wl.morphs.count), K=K)
                        ctx.deform()
                        for k in range(K):
                            rp, rn = oracle_instance(orc, wl, world[k], dense[k] if morph else None, sdef=bool(sd))
                            gp, gn = ctx.read_instance(k, normals=layout != capi.RZ_FLAG_NO_NORMALS)
                            assert rel_err(gp, rp) <= TOL, (flags, morph, k, rel_err(gp, rp))
                            if gn is not None:
                                assert rel_err(gn, rn) <= TOL, (flags, morph, k)
                            if layout == capi.RZ_FLAG_OUTLINE:
                                assert rel_err(ctx.read_outline(k), orc.outline_hull(rp, rn, edge)) <= TOL
                            if bounds:
                                bb = ctx.read_bounds(k, 1)[0]
                                assert np.array_equal(bb[:3], gp.min(axis=0)) and np.array_equal(bb[3:], gp.max(axis=0))


def test_bounds_and_positions_only(rzlib, orc, wl_small):
    wl = wl_small
    K = 9
    world = synth.make_palettes(wl.bones, K, np.random.default_rng(6))
    with capi.DeformContext(max_instances=K, flags=capi.RZ_FLAG_BOUNDS) as ctx:
        ctx.load_mesh(wl.vtx8, wl.joints, wl.weights, wl.invBind)
        ctx.set_palettes(world)
        ctx.deform()
        bb = ctx.read_bounds(0, K)
        for k in range(K):
            p, _ = ctx.read_instance(k)
            assert np.array_equal(bb[k, :3], p.min(axis=0)) and np.array_equal(bb[k, 3:], p.max(axis=0))
        check_all(orc, ctx, wl, world, None, K)
    with capi.DeformContext(max_instances=K, flags=capi.RZ_FLAG_NO_NORMALS) as ctx:
        ctx.load_mesh(wl.vtx8, wl.joints, wl.weights, wl.invBind)
        ctx.set_palettes(world)
        ctx.deform()
        for k in range(K):
            rp, _ = oracle_instance(orc, wl, world[k])
            gp, gn = ctx.read_instance(k)
            assert gn is None and rel_err(gp, rp) <= TOL


def test_reordered_vertex_storage_is_transparent_to_host_readers(rzlib, orc, wl_small):
    """RZ_FLAG_REORDER_VERTICES: device planes hold vertices sorted by bone tuple; host reads and integer tables stay in
    caller order, the order table is a permutation, morphs + SDEF keep following their vertices."""
    import torch
    wl = wl_small
    K = 5
    rng = np.random.default_rng(14)
    world = synth.make_palettes(wl.bones, K, rng)
    dense = rng.uniform(0, 1, (K, wl.morphs.count)).astype(np.float32)
    with capi.DeformContext(max_instances=K, flags=capi.RZ_FLAG_REORDER_VERTICES | capi.RZ_FLAG_SDEF) as ctx:
        ctx.load_mesh(wl.vtx8, wl.joints, wl.weights, wl.invBind)
        ctx.load_morphs(wl.morphs.offsets, wl.morphs.vertexIndex, wl.morphs.delta)
        ctx.load_sdef(wl.sdef.vertexIndex, wl.sdef.c_r0_r1)
        order = ctx.vertex_order()
        assert sorted(order.tolist()) == list(range(wl.V)) and not np.array_equal(order, np.arange(wl.V))
        j, w = ctx.read_skinning()
        assert np.array_equal(j, wl.joints.reshape(-1)) and np.array_equal(w, wl.weights.reshape(-1))
        ctx.set_palettes(world)
        ctx.set_morph_weights(dense, np.arange(wl.morphs.count), K=K)
        ctx.deform()
        check_all(orc, ctx, wl, world, None, K, morphW=dense, sdef=True)
        # the device plane really is permuted: position i of the plane is caller vertex order[i]
        base, stride, noff = ctx.output_device_ptr()
        gp, _ = ctx.read_instance(2)
        class _Dev:   # view the library-owned plane through the CUDA array interface
            __cuda_array_interface__ = {"shape": (wl.V, 3), "typestr": "<f4", "data": (base + 2 * stride, False), "version": 2}
        ctx.sync()
        plane = torch.as_tensor(_Dev(), device="cuda").cpu().numpy()
        assert np.array_equal(plane, gp[order])


def test_reordered_storage_with_fused_consumers(rzlib, orc, wl_small):
    """RZ_FLAG_REORDER_VERTICES together with the outline plane / the interleaved stream: edge sizes and texture coordinates
    follow their vertices, host readers stay in caller order."""
    wl = wl_small
    K = 3
    rng = np.random.default_rng(15)
    world = synth.make_palettes(wl.bones, K, rng)
    edge = rng.uniform(0.0, 2.0, wl.V).astype(np.float32)
    with capi.DeformContext(max_instances=K, flags=capi.RZ_FLAG_REORDER_VERTICES | capi.RZ_FLAG_OUTLINE) as ctx:
        ctx.load_mesh(wl.vtx8, wl.joints, wl.weights, wl.invBind)
        ctx.load_edge_size(edge)
        ctx.set_palettes(world)
        ctx.deform()
        for k in range(K):
            rp, rn = oracle_instance(orc, wl, world[k])
            gp, gn = ctx.read_instance(k)
            assert rel_err(gp, rp) <= TOL and rel_err(gn, rn) <= TOL
            assert rel_err(ctx.read_outline(k), orc.outline_hull(rp, rn, edge)) <= TOL
    with capi.DeformContext(max_instances=K, flags=capi.RZ_FLAG_REORDER_VERTICES | capi.RZ_FLAG_INTERLEAVED) as ctx:
        ctx.load_mesh(wl.vtx8, wl.joints, wl.weights, wl.invBind)
        ctx.set_palettes(world)
        ctx.deform()
        for k in range(K):
            rp, rn = oracle_instance(orc, wl, world[k])
            st = ctx.read_interleaved(k)
            ref = orc.interleaved(rp, rn, wl.vtx8)
            assert rel_err(st[:, :6], ref[:, :6]) <= TOL and np.array_equal(st[:, 6:], ref[:, 6:])
            gp, gn = ctx.read_instance(k)
            assert np.array_equal(gp, st[:, :3]) and np.array_equal(gn, st[:, 3:6])


def test_sub_range_deform_and_device_palettes(rzlib, orc, wl_small):
    import torch
    wl = wl_small
    K = 8
    world = synth.make_palettes(wl.bones, K, np.random.default_rng(8))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    dw = torch.from_numpy(world).cuda()
    with capi.DeformContext(max_instances=K, stream=stream.cuda_stream) as ctx:
        ctx.load_mesh(wl.vtx8, wl.joints, wl.weights, wl.invBind)
        ctx.set_palettes_device(dw.data_ptr(), K)
        ctx.deform(2, 3)
        ctx.sync()
        check_all(orc, ctx, wl, world, None, K, instances=[2, 3, 4])
        base, stride, noff = ctx.output_device_ptr()
        assert stride % 16 == 0 and noff % 16 == 0 and noff >= wl.V * 12


def test_pipelined_palette_upload(rzlib, orc, wl_small, monkeypatch):
    """rz_set_palettes with host matrices and one palette per instance uploads in blocks on a copy stream; rz_deform consumes
    them block by block (wait, skin matrices, deform).  Same bits as the unpipelined path, for whole-range and partial-range
    deforms in any order, repeated uploads, readers that need the skin matrices early, and feature sets that cannot pipeline."""
    wl = wl_small
    K = 37                                          # 8 + 8 + 8 + 8 + 5
    rng = np.random.default_rng(33)
    world = synth.make_palettes(wl.bones, K, rng)
    world2 = synth.make_palettes(wl.bones, K, rng)

    def run(pipelined: bool, flags=0):
        if pipelined:
            monkeypatch.setenv("RZ_PIPELINE_BLOCK", "8")
            monkeypatch.delenv("RZ_NO_PIPELINE", raising=False)
        else:
            monkeypatch.setenv("RZ_NO_PIPELINE", "1")
        outs = []
        with capi.DeformContext(max_instances=K, flags=flags) as ctx:
            ctx.load_mesh(wl.vtx8, wl.joints, wl.weights, wl.invBind)
            if flags & capi.RZ_FLAG_SDEF:
                ctx.load_sdef(wl.sdef.vertexIndex, wl.sdef.c_r0_r1)
            ctx.set_palettes(world)
            ctx.deform()
            outs.append([ctx.read_instance(k) for k in range(K)])
            ctx.set_palettes(world2)                # partial ranges, out of order, with a hole consumed earlier
            ctx.deform(10, 9)
            ctx.deform(0, 30)
            ctx.deform(28, 9)
            outs.append([ctx.read_instance(k) for k in range(K)])
            ctx.set_palettes(world2)                # superseded before it is consumed
            ctx.set_palettes(world)
            sm = ctx.read_skin_matrices(K - 1)      # needs the skin matrices before any deform
            ctx.deform()
            outs.append([ctx.read_instance(k) for k in range(K)])
            outs.append(sm)
        return outs

    for flags in (0, capi.RZ_FLAG_INTERLEAVED, capi.RZ_FLAG_SDEF):
        a, b = run(True, flags), run(False, flags)
        for fa, fb in zip(a[:3], b[:3]):
            for (pa, na), (pb, nb) in zip(fa, fb):
                assert np.array_equal(pa, pb) and np.array_equal(na, nb)
        assert np.array_equal(a[3], b[3])
        if flags == 0:
            for k in (0, 7, 8, 31, 36):
                rp, rn = oracle_instance(orc, wl, world[k])
                assert rel_err(a[0][k][0], rp) <= TOL and rel_err(a[0][k][1], rn) <= TOL
                rp, rn = oracle_instance(orc, wl, world2[k])
                assert rel_err(a[1][k][0], rp) <= TOL and rel_err(a[1][k][1], rn) <= TOL


@pytest.mark.parametrize("double", [False, True])
def test_async_read_back_keeps_the_frame_it_was_issued_for(rzlib, orc, wl_small, double):
    """rz_read_instance_async returns at once; the frame it was issued for arrives even when the next frame is uploaded and
    deformed before rz_read_wait -- into the other result buffer with RZ_FLAG_DOUBLE_BUFFER, after the copy otherwise."""
    import torch
    wl = wl_small
    K = 6
    rng = np.random.default_rng(44)
    frames = [synth.make_palettes(wl.bones, K, rng) for _ in range(4)]
    pin = lambda: torch.empty((wl.V, 3), dtype=torch.float32, pin_memory=True).numpy()
    bufs = [(pin(), pin()) for _ in range(2)]
    with capi.DeformContext(max_instances=K, flags=capi.RZ_FLAG_DOUBLE_BUFFER if double else 0) as ctx:
        ctx.load_mesh(wl.vtx8, wl.joints, wl.weights, wl.invBind)
        bases = []
        got = []
        for f, world in enumerate(frames):
            ctx.set_palettes(world)
            ctx.deform()
            bases.append(ctx.output_device_ptr()[0])
            ctx.read_wait()                                   # frame f-1 has landed in bufs[(f-1) & 1]
            if f:
                got.append((bufs[(f - 1) & 1][0].copy(), bufs[(f - 1) & 1][1].copy()))
            ctx.read_instance_async(f % K, *bufs[f & 1])
        ctx.read_wait()
        got.append((bufs[(len(frames) - 1) & 1][0].copy(), bufs[(len(frames) - 1) & 1][1].copy()))
        assert (len(set(bases)) == 2 and bases[0] == bases[2] and bases[1] == bases[3]) if double else len(set(bases)) == 1
        for f, world in enumerate(frames):
            rp, rn = oracle_instance(orc, wl, world[f % K])
            assert rel_err(got[f][0], rp) <= TOL and rel_err(got[f][1], rn) <= TOL, f
        # the synchronous reader sees the current frame
        gp, gn = ctx.read_instance(2)
        rp, rn = oracle_instance(orc, wl, frames[-1][2])
        assert rel_err(gp, rp) <= TOL and rel_err(gn, rn) <= TOL
    with capi.DeformContext(max_instances=1, flags=capi.RZ_FLAG_INTERLEAVED) as ctx:
        ctx.load_mesh(wl.vtx8, wl.joints, wl.weights, wl.invBind)
        with pytest.raises(capi.RzError):
            ctx.read_instance_async(0, bufs[0][0], bufs[0][1])


def test_physics_feedback_patches_the_palette_like_the_reference(rzlib, orc):
    """rz_load_rigid_bodies + rz_apply_body_transforms (physics.ts:714-751 on the device) against the host restatement:
    skin matrices of the driven bones, untouched bones bit-identical, and the deformed mesh through the oracle -- with host
    palettes (plain and pipelined upload) and with GPU pose evaluation."""
    from reze_engine_b200 import physics_bridge as pb
    rng = np.random.default_rng(77)
    wl = synth.make_workload(3000, 40, seed=77)
    ib = np.asarray(wl.invBind, np.float32).reshape(-1, 16)
    n, P = 14, 6
    bone_index = rng.integers(-1, wl.B, n).astype(np.int32)
    bone_index[1] = bone_index[0] = max(int(bone_index[0]), 0)
    dynamic = (rng.random(n) < 0.6).astype(np.uint8)
    dynamic[0] = dynamic[1] = 1
    off, inv = pb.compute_body_offsets(ib, bone_index, rng.normal(0, 3, (n, 3)), rng.uniform(-1.5, 1.5, (n, 3)))
    world = synth.make_palettes(wl.bones, P, rng)
    pos = rng.normal(0, 5, (P, n, 3))
    quat = rng.normal(size=(P, n, 4))
    quat /= np.linalg.norm(quat, axis=2, keepdims=True)
    pos[2, 1] = np.nan                                    # one invalid transform: the earlier body of that bone stays
    pq = np.concatenate([pos, quat], axis=2).astype(np.float32)
    want = world.copy()
    for p in range(P):
        pb.apply_bodies_to_bones(want[p], bone_index, dynamic, inv, pos[p], quat[p])
    driven = sorted({int(b) for b, d in zip(bone_index, dynamic) if d and b >= 0})
    for block in ("", "2"):
        if block:
            os.environ["RZ_PIPELINE_BLOCK"] = block
        try:
            with capi.DeformContext(max_instances=P) as ctx:
                ctx.load_mesh(wl.vtx8, wl.joints, wl.weights, wl.invBind)
                ctx.load_rigid_bodies(bone_index, dynamic, inv)
                ctx.set_palettes(world)
                plain = [ctx.read_skin_matrices(p) for p in range(P)]
                ctx.set_palettes(world)
                ctx.apply_body_transforms(pq)
                ctx.deform()
                for p in range(P):
                    sm = ctx.read_skin_matrices(p)
                    ref = orc.skin_matrices(want[p], wl.invBind).reshape(-1, 4, 4).transpose(0, 2, 1)[:, :3, :].reshape(-1, 12)
                    scale = max(np.abs(ref).max(), 1.0)
                    assert np.abs(sm - ref).max() / scale <= 2e-6, (p, np.abs(sm - ref).max())
                    others = [b for b in range(wl.B) if b not in driven]
                    assert np.array_equal(sm[others], plain[p][others])
                    rp, rn = orc.deform(wl.vtx8, wl.joints, wl.weights, orc.skin_matrices(want[p], wl.invBind))
                    gp, gn = ctx.read_instance(p)
                    assert rel_err(gp, rp) <= TOL and rel_err(gn, rn) <= TOL, (p, rel_err(gp, rp))
        finally:
            os.environ.pop("RZ_PIPELINE_BLOCK", None)
    # with GPU pose evaluation: the driven bones are patched after the hierarchy walk, their children keep the posed matrices
    lr = np.tile(np.array([0, 0, 0, 1], np.float32), (P, wl.B, 1))
    lr[:, :, :3] = rng.normal(0, 0.2, (P, wl.B, 3))
    lr /= np.linalg.norm(lr, axis=2, keepdims=True)
    posed = crowd.world_matrices_batch(wl.bones, lr.astype(np.float64))
    want = posed.copy()
    for p in range(P):
        pb.apply_bodies_to_bones(want[p], bone_index, dynamic, inv, pos[p], quat[p])
    with capi.DeformContext(max_instances=P) as ctx:
        ctx.load_mesh(wl.vtx8, wl.joints, wl.weights, wl.invBind)
        ctx.load_skeleton(wl.bones)
        ctx.load_rigid_bodies(bone_index, dynamic, inv)
        ctx.set_local_rotations(lr)
        ctx.apply_body_transforms(pq)
        ctx.deform()
        for p in range(P):
            rp, rn = orc.deform(wl.vtx8, wl.joints, wl.weights, orc.skin_matrices(want[p], wl.invBind))
            gp, gn = ctx.read_instance(p)
            assert rel_err(gp, rp) <= TOL and rel_err(gn, rn) <= TOL, (p, rel_err(gp, rp))
        with pytest.raises(capi.RzError):
            ctx.apply_body_transforms(pq[:2])            # P must match the frame


def test_huge_bone_count_uses_global_palette_path(rzlib, orc):
    wl = synth.make_workload(3000, 6000, seed=77)
    world = synth.make_palettes(wl.bones, 2, np.random.default_rng(9))
    with capi.DeformContext(max_instances=2) as ctx:
        ctx.load_mesh(wl.vtx8, wl.joints, wl.weights, wl.invBind)
        ctx.set_palettes(world)
        ctx.deform()
        check_all(orc, ctx, wl, world, None, 2)


def test_weight_edge_cases(rzlib, orc):
    wl = synth.make_workload(600, 10, seed=5)
    W = wl.weights.copy()
    W[0] = [0, 0, 0, 0]            # sum <= 1e-4 -> (1,0,0,0) (engine.ts:257-258)
    W[1] = [128, 0, 127, 0]        # zero weight in the middle
    W[2] = [10, 20, 30, 40]        # sum != 255 is renormalised by the shader rule
    W[3] = [0, 0, 0, 255]
    wl.weights = W
    world = synth.make_palettes(wl.bones, 1, np.random.default_rng(10))
    with capi.DeformContext(max_instances=1) as ctx:
        ctx.load_mesh(wl.vtx8, wl.joints, W, wl.invBind)
        ctx.set_palettes(world)
        ctx.deform()
        check_all(orc, ctx, wl, world, None, 1)


def test_error_paths(rzlib, wl_small):
    wl = wl_small
    with capi.DeformContext(max_instances=2) as ctx:
        with pytest.raises(capi.RzError) as e:
            ctx.lib.rz_deform(ctx.h, 0, 1) and None
            ctx._check(ctx.lib.rz_deform(ctx.h, 0, 1))
        assert e.value.status == -5
        bad = wl.joints.copy()
        bad[7, 2] = wl.B
        with pytest.raises(capi.RzError) as e:
            ctx.load_mesh(wl.vtx8, bad, wl.weights, wl.invBind)
        assert e.value.status == -1 and "joint" in str(e.value)
        ctx.load_mesh(wl.vtx8, wl.joints, wl.weights, wl.invBind)
        with pytest.raises(capi.RzError):
            ctx.set_palettes(np.zeros((3, wl.B, 16), np.float32))          # K=3 > max_instances
        with pytest.raises(capi.RzError):
            ctx.set_palettes(np.zeros((1, wl.B, 16), np.float32), [0, 1])  # palette index out of range
        with pytest.raises(capi.RzError):
            ctx.load_morphs([0, 1], [wl.V], [0, 0, 0])


@pytest.mark.parametrize("gpu_pose", [False, True])
def test_engine_facade_drives_the_path(rzlib, orc, tmp_path, gpu_pose):
    rng = np.random.default_rng(12)
    data, *_ = random_pmx(rng, V=700, B=12)
    pmx_path = tmp_path / "m.pmx"
    pmx_path.write_bytes(data)
    q1 = Quat(0, 0.3, 0, 0.95).normalize()
    vmd_path = tmp_path / "a.vmd"
    vmd_path.write_bytes(write_vmd([("骨1", 0, (0, 0, 0, 1)), ("骨1", 30, q1.toArray()), ("骨6", 15, (0.2, 0, 0, 0.98))]))
    clock = ManualClock()
    eng = Engine(None, {"ambient": 1.0, "bloomIntensity": 0.1}, instances=2, clock=clock, sdef=True, gpu_pose=gpu_pose).init()
    model = eng.loadModel(str(pmx_path))
    eng.loadAnimation(str(vmd_path))
    eng.runRenderLoop(frames=1)                        # T pose frame
    p, n = eng.readSkinned(0)
    assert rel_err(p, model.getVertices().reshape(-1, 8)[:, :3]) <= 1e-6
    eng.playAnimation()
    eng.rotateBones(["骨2"], [Quat(0, 0, 0.4, 0.9)], 0, instance=1)      # instance 1 diverges
    eng.setMorphWeights(np.array([[0.7], [0.2]], np.float32), ["m0"])
    frames = []
    eng.runRenderLoop(lambda: frames.append(clock()), frames=4, frame_ms=250.0)
    assert frames == [0.0, 250.0, 500.0, 750.0]
    st = eng.getStats()
    assert st.frameTime > 0 and st.gpuMemory > 0
    for k in range(2):
        m = eng.models[k]
        if gpu_pose:
            m.computeWorldMatrices()          # the host never walked the hierarchy in this mode: do it for the reference
        skin = orc.skin_matrices(m.getBoneWorldMatrices(), m.getBoneInverseBindMatrices())
        dense = np.zeros(m.morphs.count, np.float32)
        dense[0] = [0.7, 0.2][k]
        rp, rn = orc.deform(m.getVertices(), m.skinning.joints, m.skinning.weights, skin,
                            morph=(m.morphs.offsets, m.morphs.vertexIndex, m.morphs.delta), morphW=dense,
                            sdef=(m.sdef.vertexIndex, m.sdef.c_r0_r1))
        gp, gn = eng.readSkinned(k)
        assert rel_err(gp, rp) <= TOL and rel_err(gn, rn) <= TOL
    a, b = eng.readSkinned(0)[0], eng.readSkinned(1)[0]
    assert not np.array_equal(a, b)
    eng.dispose()


def test_gpu_pose_evaluation_local_rotations(rzlib, orc):
    """hierarchy walk + append rotation + skin matrices on the device vs the host Model restatement."""
    rng = np.random.default_rng(31)
    data, *_ = random_pmx(rng, V=600, B=12)                       # has append bones with ratios -1 / 0.25 / 0.5 / 1.5(clamped)
    m = PmxLoader.loadFromBuffer(data, clock=ManualClock())
    cases = [(m.skeleton.bones, m.getVertices(), m.skinning.joints, m.skinning.weights, m.getBoneInverseBindMatrices())]
    wl = synth.make_workload(3000, 200, seed=9)
    cases.append((wl.bones, wl.vtx8, wl.joints, wl.weights, wl.invBind))
    # a 300-bone spine (depth 300): too deep for ancestor chains -> exercises the level-ordered kernel
    wl2 = synth.make_workload(1500, 300, seed=10)
    from reze_engine_b200.pmx import compute_inverse_bind
    for i, b in enumerate(wl2.bones):
        b.parentIndex = i - 1
    inv2 = compute_inverse_bind(wl2.bones)
    cases.append((wl2.bones, wl2.vtx8, wl2.joints, wl2.weights, inv2))
    # more bones than threads: the ping-pong path of the pointer-jumping kernel (1500) and, beyond its shared-memory budget,
    # the chain-walk fallback (2600)
    for nb in (1500, 2600):
        wl3 = synth.make_workload(800, nb, seed=nb)
        cases.append((wl3.bones, wl3.vtx8, wl3.joints, wl3.weights, wl3.invBind))
    for bones, vtx, J, W, inv in cases:
        B, P = len(bones), 5
        qa, qb, _ = synth.make_crowd_tween(B, P, rng)
        lr = crowd.tween_pose_batch(qa, qb, rng.uniform(0, 1, P))          # [P,B,4] f32
        world = crowd.world_matrices_batch(bones, lr)                       # bit-exact host evaluation (test_host.py)
        i2p = np.array([4, 0, 2, 2, 1, 3, 0], np.uint32)
        with capi.DeformContext(max_instances=7) as ctx:
            ctx.load_mesh(vtx, J, W, inv)
            ctx.load_skeleton(bones)
            ctx.set_local_rotations(lr, i2p)
            ref = orc.skin_matrices(world[3], inv).reshape(-1, 4, 4).transpose(0, 2, 1)[:, :3, :].reshape(-1, 12)
            assert rel_err(ctx.read_skin_matrices(3), ref) <= TOL
            ctx.deform()
            for k in range(7):
                rp, rn = orc.deform(vtx, J, W, orc.skin_matrices(world[i2p[k]], inv))
                gp, gn = ctx.read_instance(k)
                assert rel_err(gp, rp) <= TOL and rel_err(gn, rn) <= TOL


def test_gpu_pose_evaluation_tweens_with_instance_clocks(rzlib, orc):
    """Crowd playback: one shared tween state, instance k evaluated at its own clock (staggered phase)."""
    rng = np.random.default_rng(32)
    wl = synth.make_workload(2000, 96, seed=12)
    B = wl.B
    clock = ManualClock(1000.0)
    model = wl.model(clock=clock)
    names = [b.name for b in wl.bones]
    quats = [Quat(*rng.normal(size=4)) for _ in range(B)]
    model.rotateBones(names[:60], quats[:60], 800)                       # 60 bones tweening, 20 set instantly, 16 at rest
    model.rotateBones(names[60:80], quats[60:80], 0)
    offsets = np.array([0.0, 100.0, 399.5, 800.0, 2000.0, -50.0], np.float32)
    with capi.DeformContext(max_instances=6) as ctx:
        ctx.load_mesh(wl.vtx8, wl.joints, wl.weights, wl.invBind)
        ctx.load_skeleton(wl.bones)
        ctx.set_tweens(model._startQuat, model._targetQuat, model._startTimeMs, model._durationMs, model._active, model.localRotations)
        ctx.set_instance_clocks(1000.0 + offsets)
        ctx.deform()
        for k, off in enumerate(offsets):
            ref = wl.model(clock=ManualClock(1000.0 + float(off)))
            for arr in ("_startQuat", "_targetQuat", "_startTimeMs", "_durationMs", "_active", "localRotations"):
                getattr(ref, arr)[:] = getattr(model, arr)
            ref.evaluatePose()
            rp, rn = orc.deform(wl.vtx8, wl.joints, wl.weights, orc.skin_matrices(ref.getBoneWorldMatrices(), wl.invBind))
            gp, gn = ctx.read_instance(k)
            assert rel_err(gp, rp) <= TOL and rel_err(gn, rn) <= TOL, k


def test_engine_crowd_playback_matches_timer_driven_reference_playback(rzlib, orc, tmp_path):
    """BASELINE config 2 through the facade: K instances play one VMD at staggered offsets, keyframe tracks evaluated
    on the device, vs the reference-style playback (timers + rotateBones tweens on the host) sampled at the same times."""
    rng = np.random.default_rng(41)
    data, *_ = random_pmx(rng, V=500, B=12, n_morph=0, with_sdef=False)
    pmx_path = tmp_path / "m.pmx"
    pmx_path.write_bytes(data)
    qs = [Quat(*rng.normal(size=4)).normalize() for _ in range(5)]
    vmd_path = tmp_path / "a.vmd"
    vmd_path.write_bytes(write_vmd([("骨1", 0, qs[0].toArray()), ("骨1", 15, qs[1].toArray()), ("骨1", 30, qs[2].toArray()),
                                    ("骨6", 15, qs[3].toArray()), ("骨9", 45, qs[4].toArray()), ("ghost", 10, (0, 0, 0, 1))]))
    step = 50.0
    # reference-style playback, one instance, sampled every 50 ms (all key times are multiples of 50 ms)
    rclock = ManualClock()
    ref = Engine(None, None, instances=1, clock=rclock).init()
    rmodel = ref.loadModel(str(pmx_path))
    ref.loadAnimation(str(vmd_path))
    ref.playAnimation()
    samples = {}
    for f in range(0, 45):
        ref.render()
        samples[f * step] = rmodel.getBoneWorldMatrices().copy()
        rclock.advance(step)
    ref.dispose()
    # crowd playback
    offsets = np.array([0.0, 100.0, 250.0, 500.0, 1000.0, 1700.0])
    clock = ManualClock()
    eng = Engine(None, None, instances=6, clock=clock, crowd=True).init()
    model = eng.loadModel(str(pmx_path))
    eng.loadAnimation(str(vmd_path))
    eng.setInstanceOffsets(offsets)
    eng.playAnimation()
    vt, J, W, inv = model.getVertices(), model.skinning.joints, model.skinning.weights, model.getBoneInverseBindMatrices()
    for f in range(0, 40, 3):
        clock.now_ms = f * step
        eng.render()
        for k, off in enumerate(offsets):
            tau = max(0.0, f * step - off)               # before its start an instance shows the t=0 pose
            rp, rn = orc.deform(vt, J, W, orc.skin_matrices(samples[tau], inv))
            gp, gn = eng.readSkinned(k)
            assert rel_err(gp, rp) <= TOL and rel_err(gn, rn) <= TOL, (f, k)
    # stop: back to the shared rotateBones pose, still per-instance clocks
    eng.stopAnimation()
    eng.rotateBones(["骨2"], [qs[0]], 400)
    clock.advance(200.0)
    eng.render()
    a, b = eng.readSkinned(0)[0], eng.readSkinned(2)[0]      # instance 2 is 250 ms behind: its tween has not started
    assert not np.array_equal(a, b)
    assert rel_err(b, eng.readSkinned(5)[0]) <= 1e-6
    eng.dispose()


def test_vmd_morph_track_playback(rzlib, orc, tmp_path):
    """the VMD morph table drives per-instance morph weights (linear between keys), crowd offsets included."""
    rng = np.random.default_rng(43)
    data, *_ = random_pmx(rng, V=400, B=8, n_morph=3, with_sdef=False)
    (tmp_path / "m.pmx").write_bytes(data)
    (tmp_path / "a.vmd").write_bytes(write_vmd([("骨1", 0, (0, 0, 0, 1)), ("骨1", 30, (0, 0.3, 0, 0.95))],
                                               [("m0", 0, 0.0), ("m0", 30, 1.0), ("m2", 15, 0.5), ("nope", 3, 1.0)]))
    clock = ManualClock()
    eng = Engine(None, None, instances=3, clock=clock, crowd=True).init()
    model = eng.loadModel(str(tmp_path / "m.pmx"))
    eng.loadAnimation(str(tmp_path / "a.vmd"))
    eng.setInstanceOffsets([0.0, 250.0, 2000.0])
    eng.playAnimation()
    clock.now_ms = 500.0
    eng.render()
    names = model.morphs.names
    for k, tau in enumerate((500.0, 250.0, 0.0)):
        dense = np.zeros(model.morphs.count, np.float32)
        dense[names.index("m0")] = min(1.0, tau / 1000.0)
        dense[names.index("m2")] = 0.5                      # single key: held
        e = min(1.0, tau / 1000.0)
        from reze_engine_b200.math3d import easeInOut
        q = Quat.slerp(Quat(0, 0, 0, 1), Quat(0, 0.3, 0, 0.95).normalize(), easeInOut(e))
        ref = Engine(None, None, instances=1, clock=ManualClock()).init()
        rm = ref.loadModel(str(tmp_path / "m.pmx"))
        rm.rotateBones(["骨1"], [q], 0)
        rm.evaluatePose()
        rp, rn = orc.deform(rm.getVertices(), rm.skinning.joints, rm.skinning.weights,
                            orc.skin_matrices(rm.getBoneWorldMatrices(), rm.getBoneInverseBindMatrices()),
                            morph=(rm.morphs.offsets, rm.morphs.vertexIndex, rm.morphs.delta), morphW=dense)
        ref.dispose()
        gp, gn = eng.readSkinned(k)
        assert rel_err(gp, rp) <= TOL and rel_err(gn, rn) <= TOL, k
    eng.dispose()


def _load_local(name):
    path = os.path.join(LOCAL, name)
    if not os.path.exists(path):
        pytest.skip("tests/golden/_local not generated (needs the reference assets; see make_fixtures.py)")
    return np.load(path, allow_pickle=False)


def _bones_from(z):
    bones = []
    for i in range(len(z["parents"])):
        ap = int(z["appendParent"][i])
        ar = float(z["appendRatio"][i])
        bones.append(Bone(name=str(z["names"][i]), parentIndex=int(z["parents"][i]), bindTranslation=[float(x) for x in z["bindTranslation"][i]],
                          appendParentIndex=None if ap < 0 and not bool(z["appendRotate"][i]) and not bool(z["appendMove"][i]) else ap,
                          appendRatio=None if np.isnan(ar) else ar, appendRotate=bool(z["appendRotate"][i]), appendMove=bool(z["appendMove"][i])))
    return bones


def test_engine_reordered_storage_draws_the_same_triangles(rzlib, tmp_path):
    """Engine(reorder_vertices=True): same host-visible result as the default engine; the remapped index buffer addresses the
    permuted device planes so every triangle corner lands on the same skinned position."""
    rng = np.random.default_rng(47)
    data, *_ = random_pmx(rng, V=900, B=14, n_morph=2, with_sdef=True)
    (tmp_path / "m.pmx").write_bytes(data)
    outs = []
    for reorder in (False, True):
        eng = Engine(None, None, instances=2, clock=ManualClock(), sdef=True, reorder_vertices=reorder).init()
        model = eng.loadModel(str(tmp_path / "m.pmx"))
        eng.rotateBones(["骨1", "骨3"], [Quat.fromEuler(0.2, -0.4, 0.1), Quat.fromEuler(-0.3, 0.1, 0.5)], 0)
        eng.render()
        pos, nrm = eng.readSkinned(1)
        didx = eng.deviceIndexBuffer()
        order = eng.ctx.vertex_order()
        assert np.array_equal(order[didx], np.asarray(model.getIndices(), np.uint32))
        outs.append((pos, nrm, order))
        eng.dispose()
    assert np.array_equal(outs[0][2], np.arange(900)) and not np.array_equal(outs[1][2], np.arange(900))
    assert rel_err(outs[1][0], outs[0][0]) <= 2e-6 and rel_err(outs[1][1], outs[0][1]) <= 2e-6


@pytest.mark.parametrize("key", ["serqet", "serqet2"])
def test_real_model_poses(rzlib, orc, key):
    """BASELINE configs 0/1: the shipped PMX in T pose, the tutorial pose (canvas4.tsx:16-17: 腰 = (0,0.3,0,1), 首 = id)
    and the end pose of pool.vmd, skinned on the GPU vs the oracle; integer tables bit-exact."""
    from reze_engine_b200.model import Model, Skeleton, Skinning
    z = _load_local(f"{key}.npz")
    bones = _bones_from(z)
    clock = ManualClock()
    model = Model(z["vtx8"], np.zeros(0, np.uint32), [], [], Skeleton(bones, z["invBind"]), Skinning(z["joints"], z["weights"]), clock=clock)
    V, B = model.vertexCount, len(bones)
    vmd = _load_local("pool_vmd.npz")
    poses = []
    model.evaluatePose()
    poses.append(model.getBoneWorldMatrices().copy())
    model.rotateBones(["腰", "首"], [Quat(0, 0.3, 0, 1), Quat(0, 0, 0, 1)], 0)
    model.evaluatePose()
    poses.append(model.getBoneWorldMatrices().copy())
    last = {}
    for nm, row in zip(vmd["names"], vmd["data"]):
        last[str(nm)] = Quat(*row[1:5])
    model.rotateBones(list(last.keys()), list(last.values()), 0)
    model.evaluatePose()
    poses.append(model.getBoneWorldMatrices().copy())
    world = np.stack(poses).reshape(3, B, 16)
    with capi.DeformContext(max_instances=3) as ctx:
        ctx.load_mesh(z["vtx8"], z["joints"], z["weights"], z["invBind"])
        j, w = ctx.read_skinning()
        assert np.array_equal(j, z["joints"]) and np.array_equal(w, z["weights"])
        ctx.set_palettes(world)
        ctx.deform()
        vt = z["vtx8"].reshape(-1, 8)
        p0, n0 = ctx.read_instance(0)
        assert rel_err(p0, vt[:, :3]) <= 1e-6
        for k in range(3):
            skin = orc.skin_matrices(world[k], z["invBind"])
            rp, rn = orc.deform(z["vtx8"], z["joints"], z["weights"], skin)
            gp, gn = ctx.read_instance(k)
            assert rel_err(gp, rp) <= TOL and rel_err(gn, rn) <= TOL
        assert rel_err(ctx.read_instance(1)[0], p0) > 1e-3        # the pose really moved vertices


def test_headline_size_properties(rzlib, orc):
    """BASELINE headline shape (V=100k, B=512) at a K that keeps the test short: oracle on sampled instances,
    idempotence, and bit-identical outputs for instances sharing a palette."""
    import zlib
    wl = synth.make_workload(100_000, 512)
    K, P = 256, 64
    world = synth.make_palettes(wl.bones, P, np.random.default_rng(21))
    i2p = (np.arange(K) * 7) % P
    with capi.DeformContext(max_instances=K) as ctx:
        ctx.load_mesh(wl.vtx8, wl.joints, wl.weights, wl.invBind)
        ctx.set_palettes(world, i2p)
        ctx.deform()
        check_all(orc, ctx, wl, world, i2p, K, instances=[0, 101, 255])
        crc1 = [zlib.crc32(ctx.read_instance(k)[0].tobytes()) for k in (3, 67, 131, 200)]
        ctx.deform()
        crc2 = [zlib.crc32(ctx.read_instance(k)[0].tobytes()) for k in (3, 67, 131, 200)]
        assert crc1 == crc2                                           # idempotent
        same = [k for k in range(K) if i2p[k] == i2p[3]]
        assert len(same) >= 4
        ref = ctx.read_instance(3)
        for k in same[1:4]:
            got = ctx.read_instance(k)
            assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])


@pytest.mark.parametrize("layout", ["positions_only", "outline", "interleaved", "bounds"])
@pytest.mark.parametrize("I,nt,sb", [(0, 0, 0), (4, 512, 2), (4, 384, 2), (6, 512, 1), (2, 256, 2), (1, 256, 1)])
def test_two_vertex_kernel_in_every_output_layout(rzlib, orc, layout, I, nt, sb):
    """The fused consumers without morph / SDEF run the two-vertices-per-lane kernel as well (deform2_kernel OUT2_*):
    positions only (engine.ts:692-715), outline hull plane (engine.ts:458-461), interleaved 32-byte stream (engine.ts:340-347),
    planar + per-instance AABB — ragged vertex count, partial instance group, sub-range, vs the oracle; uv bit-exact; AABB
    = min / max of the positions read back, bit for bit; same layout as the one-vertex kernel."""
    wl = synth.make_workload(1003, 48, seed=19)
    K = 7
    rng = np.random.default_rng(23)
    world = synth.make_palettes(wl.bones, K, rng)
    edge = np.where(rng.uniform(size=wl.V) < 0.7, rng.uniform(0.2, 2.0, wl.V), 0.0).astype(np.float32)
    flags = {"positions_only": capi.RZ_FLAG_NO_NORMALS, "outline": capi.RZ_FLAG_OUTLINE, "interleaved": capi.RZ_FLAG_INTERLEAVED,
             "bounds": capi.RZ_FLAG_BOUNDS}[layout]
    got = {}
    for vpl in (2, 1):
        with capi.DeformContext(max_instances=K, flags=flags, instances_per_group=I if vpl == 2 else 0, threads=nt if vpl == 2 else 0,
                                store_mode=sb if vpl == 2 else 0, vertices_per_lane=vpl) as ctx:
            ctx.load_mesh(wl.vtx8, wl.joints, wl.weights, wl.invBind)
            if layout == "outline":
                ctx.load_edge_size(edge)
            ctx.set_palettes(world)
            ctx.deform()
            ctx.set_palettes(world)
            ctx.deform(2, 4)
            assert ctx.stats()["verticesPerLane"] == vpl
            lay = {k: v for k, v in ctx.output_layout().items() if k != "base"}
            assert lay == got.setdefault("layout", lay)
            for k in range(K):
                rp, rn = oracle_instance(orc, wl, world[k])
                gp, gn = ctx.read_instance(k, normals=layout != "positions_only")
                assert rel_err(gp, rp) <= TOL, (vpl, k)
                if layout != "positions_only":
                    assert rel_err(gn, rn) <= TOL, (vpl, k)
                if layout == "outline":
                    assert rel_err(ctx.read_outline(k), orc.outline_hull(rp, rn, edge)) <= TOL
                if layout == "bounds":
                    bb = ctx.read_bounds(k, 1)[0]
                    assert np.array_equal(bb[:3], gp.min(axis=0)) and np.array_equal(bb[3:], gp.max(axis=0)), (vpl, k)
                if layout == "interleaved":
                    st = ctx.read_interleaved(k)
                    assert np.array_equal(st[:, :3], gp) and np.array_equal(st[:, 3:6], gn)
                    assert np.array_equal(st[:, 6:], wl.vtx8.reshape(-1, 8)[:, 6:])             # uv: bit-exact pass-through
