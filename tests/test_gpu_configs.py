"""GPU parity (-m gpu) at the sizes BASELINE.json's `configs` name — the driver-run counterpart of tools/run_configs.py.

config 1  塞尔凯特2.pmx + pool.vmd, every one of the 46 frames (0..45 @ 30 fps), K = 1
config 2  the same mesh as a 1024-instance crowd, staggered clip phase, GPU pose evaluation
config 3  V = 200 000, B = 512, 64 active vertex morphs, SDEF on, K = 256
headline  V = 100 000, B = 512 at a K that holds several partial / full instance groups

Reference behaviour followed: the vertex-shader blend engine.ts:253-272 (oracle/rz_oracle_body.inc), the clip schedule
engine.ts:1451-1553 (key at t = 0 instantly, key i reached from key i-1 by a tween of their distance) driven through
Model.rotateBones / evaluatePose (model.ts:158-194, 246-315, 330-420).  Tolerance: 1e-5 relative against
the f32 oracle; the f32-vs-f64 oracle gap is asserted alongside so the tolerance is seen to have room.
"""
import os

import numpy as np
import pytest

from helpers import rel_err
from reze_engine_b200 import Quat, capi, synth
from reze_engine_b200.engine import ManualClock
from reze_engine_b200.model import Bone, Model, Skeleton, Skinning

pytestmark = pytest.mark.gpu
TOL = 1e-5
LOCAL = os.path.join(os.path.dirname(__file__), "golden", "_local")


def _local(name):
    path = os.path.join(LOCAL, name)
    if not os.path.exists(path):
        pytest.skip("tests/golden/_local not generated (needs the reference assets; see make_fixtures.py)")
    return np.load(path, allow_pickle=False)


def _bones(z):
    out = []
    for i in range(len(z["parents"])):
        ap, ar = int(z["appendParent"][i]), float(z["appendRatio"][i])
        out.append(Bone(name=str(z["names"][i]), parentIndex=int(z["parents"][i]), bindTranslation=[float(x) for x in z["bindTranslation"][i]],
                        appendParentIndex=None if ap < 0 and not bool(z["appendRotate"][i]) and not bool(z["appendMove"][i]) else ap,
                        appendRatio=None if np.isnan(ar) else ar, appendRotate=bool(z["appendRotate"][i]), appendMove=bool(z["appendMove"][i])))
    return out


def _clip_tracks(vmd, bones):
    """pool.vmd as per-bone key tracks (CSR over bones): times in ms, quaternions xyzw — what Engine.loadAnimation hands to
    rz_load_animation."""
    idx = {b.name: i for i, b in enumerate(bones)}
    per = [[] for _ in bones]
    for nm, row in zip(vmd["names"], vmd["data"]):
        i = idx.get(str(nm))
        if i is not None:
            per[i].append((float(row[0]) * 1000.0, row[1:5]))
    off, times, quats = [0], [], []
    for keys in per:
        keys.sort(key=lambda k: k[0])
        for t, q in keys:
            times.append(t)
            quats.append(q)
        off.append(len(times))
    return (np.asarray(off, np.uint32), np.asarray(times, np.float32), np.asarray(quats, np.float32).reshape(-1, 4)), per


def _play_reference_style(model, clock, per, names):
    """engine.ts:1451-1553 on a manual clock: keys at t = 0 are applied instantly (bones without one reset to identity), every
    later key i is a tween of duration t(i) - t(i-1) that the reference starts from a timer at wall time t(i-1).  Returns
    a function that advances the model to clip time t (ms, non-decreasing calls) and returns its world matrices."""
    ident = Quat(0, 0, 0, 1)
    pending = []                                          # (start time, bone name, target, duration)
    for i, keys in enumerate(per):
        if not keys:
            continue
        if keys[0][0] == 0.0:
            model.rotateBones([names[i]], [Quat(*[float(x) for x in keys[0][1]])], 0)
        else:
            model.rotateBones([names[i]], [ident], 0)
        for k, (t, q) in enumerate(keys):
            if t == 0.0:
                continue
            t_prev = keys[k - 1][0] if k > 0 else 0.0
            pending.append((t_prev, names[i], Quat(*[float(x) for x in q]), t - t_prev))
    pending.sort(key=lambda p: p[0])
    state = {"next": 0}

    def at(t_ms):
        while state["next"] < len(pending) and pending[state["next"]][0] <= t_ms:
            t0, nm, q, dur = pending[state["next"]]
            clock.now_ms = t0
            model.rotateBones([nm], [q], dur)
            state["next"] += 1
        clock.now_ms = t_ms
        model.evaluatePose()
        return model.getBoneWorldMatrices().copy()
    return at


def test_config1_real_pmx_every_frame_of_pool_vmd(rzlib, orc):
    """BASELINE config 1: 塞尔凯特2.pmx (28 842 verts, 349 bones) through all 46 frames of pool.vmd.  Two feeds per frame:
    (a) world matrices from the host Model exactly as the reference uploads them (rz_set_palettes), (b) the clip evaluated
    on the device (rz_load_animation + rz_set_instance_clocks).  Both vs the oracle fed with the host Model's matrices."""
    z = _local("serqet2.npz")
    vmd = _local("pool_vmd.npz")
    bones = _bones(z)
    B = len(bones)
    names = [b.name for b in bones]
    clock = ManualClock()
    model = Model(z["vtx8"], np.zeros(0, np.uint32), [], [], Skeleton(bones, z["invBind"]), Skinning(z["joints"], z["weights"]), clock=clock)
    tracks, per = _clip_tracks(vmd, bones)
    assert sum(len(k) for k in per) == 68 and sum(1 for k in per if k) == 36          # 68 keys on 36 bones
    at = _play_reference_style(model, clock, per, names)
    worst_a = worst_b = 0.0
    moved = 0.0
    with capi.DeformContext(max_instances=1) as host_fed, capi.DeformContext(max_instances=1) as dev_fed:
        for ctx in (host_fed, dev_fed):
            ctx.load_mesh(z["vtx8"], z["joints"], z["weights"], z["invBind"])
        dev_fed.load_skeleton(bones)
        dev_fed.load_animation(*tracks)
        rest = z["vtx8"].reshape(-1, 8)[:, :3]
        for f in range(46):
            t_ms = f * 1000.0 / 30.0
            world = at(t_ms).reshape(1, B, 16)
            rp, rn = orc.deform(z["vtx8"], z["joints"], z["weights"], orc.skin_matrices(world[0], z["invBind"]))
            host_fed.set_palettes(world)
            host_fed.deform()
            gp, gn = host_fed.read_instance(0)
            worst_a = max(worst_a, rel_err(gp, rp), rel_err(gn, rn))
            dev_fed.set_instance_clocks(np.array([t_ms], np.float32))
            dev_fed.deform()
            dp, dn = dev_fed.read_instance(0)
            worst_b = max(worst_b, rel_err(dp, rp), rel_err(dn, rn))
            moved = max(moved, rel_err(rp, rest))
    assert worst_a <= TOL, worst_a
    assert worst_b <= TOL, worst_b
    assert moved > 1e-2                                  # the clip really moves the mesh


def test_config2_real_pmx_crowd_1024_staggered_phase(rzlib, orc):
    """BASELINE config 2: 1024 instances of 塞尔凯特2.pmx, each at its own phase of pool.vmd (golden-ratio stagger), pose
    evaluated on the device from one clock value per instance.  Sampled instances cover the first and last instance of
    the first full group and of the trailing partial group for every instance-group width the kernel is built with
    (1024 = 170*6 + 4 = 256*4 = 341*3 + 1), plus random ones; each vs the oracle fed by the host Model at that clip time."""
    z = _local("serqet2.npz")
    vmd = _local("pool_vmd.npz")
    bones = _bones(z)
    B = len(bones)
    names = [b.name for b in bones]
    tracks, per = _clip_tracks(vmd, bones)
    K = 1024
    phase = (np.arange(K) * synth.GOLDEN) % 1.0
    clk = (phase * 1500.0).astype(np.float32)            # pool.vmd spans frames 0..45 = 1500 ms
    rng = np.random.default_rng(2)
    sample = sorted({0, 5, 6, 511, 1019, 1020, 1022, 1023, *[int(x) for x in rng.integers(0, K, 4)]})
    with capi.DeformContext(max_instances=K) as ctx:
        ctx.load_mesh(z["vtx8"], z["joints"], z["weights"], z["invBind"])
        ctx.load_skeleton(bones)
        ctx.load_animation(*tracks)
        ctx.set_instance_clocks(clk)
        ctx.deform()
        worst = 0.0
        for k in sample:
            clock = ManualClock()
            model = Model(z["vtx8"], np.zeros(0, np.uint32), [], [], Skeleton(bones, z["invBind"]), Skinning(z["joints"], z["weights"]), clock=clock)
            world = _play_reference_style(model, clock, per, names)(float(clk[k]))
            rp, rn = orc.deform(z["vtx8"], z["joints"], z["weights"], orc.skin_matrices(world.reshape(B, 16), z["invBind"]))
            gp, gn = ctx.read_instance(k)
            e = max(rel_err(gp, rp), rel_err(gn, rn))
            assert e <= TOL, (k, e)
            worst = max(worst, e)
        # staggered phases really differ; equal clocks are bit-identical
        a, b = ctx.read_instance(0)[0], ctx.read_instance(1)[0]
        assert rel_err(a, b) > 1e-3
        clk2 = clk.copy()
        clk2[1023] = clk2[0]
        ctx.set_instance_clocks(clk2)
        ctx.deform()
        assert np.array_equal(ctx.read_instance(1023)[0], ctx.read_instance(0)[0])
        s = ctx.stats()
        assert s["instanceCount"] == K and s["vertexCount"] == 28842 and s["boneCount"] == 349


def test_config3_highpoly_64_morphs_sdef_at_size(rzlib, orc):
    """BASELINE config 3 at size: V = 200 000, B = 512, 64 active vertex morphs, SDEF on, K = 256.  Sampled instances
    (first / last of a group, middle, last) against the f32 oracle; the f64 oracle bounds how much of the tolerance f32
    arithmetic itself uses.  Then the size-independent properties: zero morph weights + compat mode reproduce the plain
    path bit for bit on non-SDEF vertices, and the deform is idempotent."""
    wl = synth.make_workload(200_000, 512, M=64, sdef=True)
    K = 256
    rng = np.random.default_rng(3)
    world = synth.make_palettes(wl.bones, K, rng)
    mw = rng.uniform(0, 1, (K, 64)).astype(np.float32)
    morph = (wl.morphs.offsets, wl.morphs.vertexIndex, wl.morphs.delta)
    sd = (wl.sdef.vertexIndex, wl.sdef.c_r0_r1)
    sample = [0, 3, 4, 100, 255]
    with capi.DeformContext(max_instances=K, flags=capi.RZ_FLAG_SDEF) as ctx:
        ctx.load_mesh(wl.vtx8, wl.joints, wl.weights, wl.invBind)
        ctx.load_morphs(*morph)
        ctx.load_sdef(*sd)
        ctx.set_palettes(world)
        ctx.set_morph_weights(mw, np.arange(64), K=K)
        ctx.deform()
        s = ctx.stats()
        assert s["morphNnz"] > 300_000 and s["sdefCount"] > 15_000 and s["activeMorphs"] == 64
        worst = gap = 0.0
        for k in sample:
            skin = orc.skin_matrices(world[k], wl.invBind)
            rp, rn = orc.deform(wl.vtx8, wl.joints, wl.weights, skin, morph=morph, morphW=mw[k], sdef=sd)
            gp, gn = ctx.read_instance(k)
            e = max(rel_err(gp, rp), rel_err(gn, rn))
            assert e <= TOL, (k, e)
            worst = max(worst, e)
            if k in (0, 255):
                skin64 = orc.skin_matrices(world[k], wl.invBind, dtype=np.float64)
                dp, dn = orc.deform(wl.vtx8, wl.joints, wl.weights, skin64, morph=morph, morphW=mw[k], sdef=sd, dtype=np.float64)
                gap = max(gap, rel_err(rp, dp), rel_err(rn, dn))
                assert max(rel_err(gp, dp), rel_err(gn, dn)) <= TOL
        assert gap <= 5e-6, gap
        before = ctx.read_instance(100)
        ctx.deform()
        after = ctx.read_instance(100)
        assert np.array_equal(before[0], after[0]) and np.array_equal(before[1], after[1])
    # zero weights + compat mode (SDEF as BDEF2, pmx-loader.ts:141-155) == the pinned plain path
    with capi.DeformContext(max_instances=2) as plain, capi.DeformContext(max_instances=2) as compat:
        plain.load_mesh(wl.vtx8, wl.joints, wl.weights, wl.invBind)
        compat.load_mesh(wl.vtx8, wl.joints, wl.weights, wl.invBind)
        compat.load_morphs(*morph)
        compat.load_sdef(*sd)
        for c in (plain, compat):
            c.set_palettes(world[:2])
        compat.set_morph_weights(np.zeros((2, 64), np.float32), np.arange(64), K=2)
        plain.deform()
        compat.deform()
        for k in range(2):
            a, b = plain.read_instance(k), compat.read_instance(k)
            assert rel_err(b[0], a[0]) <= 1e-6 and rel_err(b[1], a[1]) <= 1e-6


def test_headline_shape_partial_groups_and_full_palette_set(rzlib, orc):
    """BASELINE headline shape (V = 100 000, B = 512) with one palette PER instance as the bench runs it, K = 1030 so that
    every compiled group width ends in a partial group; first / last instance of the first and last group vs the oracle,
    and a checksum of checksums over all instances is reproducible across two frames."""
    import zlib
    wl = synth.make_workload(100_000, 512)
    K = 1030
    world = synth.make_palettes(wl.bones, K, np.random.default_rng(11))
    with capi.DeformContext(max_instances=K) as ctx:
        ctx.load_mesh(wl.vtx8, wl.joints, wl.weights, wl.invBind)
        ctx.set_palettes(world)
        ctx.deform()
        for k in (0, 5, 6, 515, 1024, 1025, 1029):
            rp, rn = orc.deform(wl.vtx8, wl.joints, wl.weights, orc.skin_matrices(world[k], wl.invBind))
            gp, gn = ctx.read_instance(k)
            assert rel_err(gp, rp) <= TOL and rel_err(gn, rn) <= TOL, k
        crc = [zlib.crc32(ctx.read_instance(k)[0].tobytes()) for k in range(0, K, 41)]
        ctx.set_palettes(world)
        ctx.deform()
        assert crc == [zlib.crc32(ctx.read_instance(k)[0].tobytes()) for k in range(0, K, 41)]
        assert len(set(crc)) == len(crc)                              # every instance has its own pose


def test_multi_device_engine_two_logical_shards_equal_the_unsharded_crowd(rzlib, tmp_path):
    """Single-process multi-device (SURVEY 8b row 1 'device list', 8e): MultiDeviceEngine(devices=[0, 0]) — two contexts with
    their own streams, instance ranges from sharding.instance_range — plays the same staggered-phase crowd as one Engine
    with all K instances and must reproduce it BIT FOR BIT, instance by instance (instances are independent; no collective).
    Also through the raw C ABI: two contexts deforming the two halves of a palette set vs one context deforming all."""
    from helpers import random_pmx, write_vmd
    from reze_engine_b200 import Engine, MultiDeviceEngine, sharding
    rng = np.random.default_rng(52)
    data, *_ = random_pmx(rng, V=1500, B=20, n_morph=0, with_sdef=False)
    (tmp_path / "m.pmx").write_bytes(data)
    qs = [Quat(*rng.normal(size=4)).normalize() for _ in range(4)]
    (tmp_path / "a.vmd").write_bytes(write_vmd([("骨1", 0, qs[0].toArray()), ("骨1", 20, qs[1].toArray()), ("骨4", 10, qs[2].toArray()),
                                                ("骨7", 30, qs[3].toArray())]))
    K = 21
    offsets = (np.arange(K) * 61.8) % 1000.0
    outs = []
    for devices in (None, [0, 0], [0, 0, 0]):
        clock = ManualClock()
        eng = (Engine(None, None, instances=K, clock=clock, crowd=True) if devices is None
               else MultiDeviceEngine(None, None, devices=devices, instances=K, clock=clock, crowd=True)).init()
        eng.loadModel(str(tmp_path / "m.pmx"))
        eng.loadAnimation(str(tmp_path / "a.vmd"))
        eng.setInstanceOffsets(offsets)
        eng.playAnimation()
        frames = []
        for f in range(3):
            clock.now_ms = 400.0 * f + 150.0
            eng.render()
            frames.append([eng.readSkinned(k) for k in range(K)])
        if devices is not None:
            assert [(first, count) for _, first, count in eng.shards] == [
                (sharding.instance_range(K, len(devices), g)[0], sharding.instance_range(K, len(devices), g)[1] - sharding.instance_range(K, len(devices), g)[0])
                for g in range(len(devices))]
            assert len({id(e.ctx) for e, _, _ in eng.shards}) == len(devices)
        outs.append(frames)
        eng.dispose()
    for sharded in outs[1:]:
        for fa, fb in zip(outs[0], sharded):
            for (pa, na), (pb, nb) in zip(fa, fb):
                assert np.array_equal(pa, pb) and np.array_equal(na, nb)
    assert rel_err(outs[0][0][0][0], outs[0][2][0][0]) > 1e-3            # the clip moves the mesh between the sampled frames

    wl = synth.make_workload(20_000, 128)
    K = 40
    world = synth.make_palettes(wl.bones, K, np.random.default_rng(8))
    with capi.DeformContext(max_instances=K) as whole:
        whole.load_mesh(wl.vtx8, wl.joints, wl.weights, wl.invBind)
        whole.set_palettes(world)
        whole.deform()
        ctxs = []
        for g in range(2):
            first, last = sharding.instance_range(K, 2, g)
            c = capi.DeformContext(max_instances=last - first)
            c.load_mesh(wl.vtx8, wl.joints, wl.weights, wl.invBind)
            c.set_palettes(world[first:last])
            ctxs.append((c, first, last))
        for c, _, _ in ctxs:                                            # both launched before either is read
            c.deform()
        for c, first, last in ctxs:
            for k in range(first, last, 7):
                a, b = whole.read_instance(k), c.read_instance(k - first)
                assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
            c.close()


def test_world_matrices_come_back_from_a_device_evaluated_pose(rzlib):
    """rz_read_world_matrices: the bone world matrices of a pose evaluated ON THE DEVICE (local rotations in, hierarchy +
    append on the GPU) equal what Model.evaluatePose leaves in getBoneWorldMatrices() (model.ts:317-319, 330-420) — the input
    the kinematic half of Physics.step needs back on the host (syncFromBones, physics.ts:649-703)."""
    z = _local("serqet2.npz")
    bones = _bones(z)
    B = len(bones)
    clock = ManualClock()
    model = Model(z["vtx8"], np.zeros(0, np.uint32), [], [], Skeleton(bones, z["invBind"]), Skinning(z["joints"], z["weights"]), clock=clock)
    rng = np.random.default_rng(9)
    names = [b.name for b in bones]
    pick = rng.choice(B, 60, replace=False)
    model.rotateBones([names[i] for i in pick], [Quat(*rng.normal(size=4)).normalize() for _ in pick], 0)
    model.evaluatePose()
    ref = model.getBoneWorldMatrices().reshape(B, 16).copy()
    with capi.DeformContext(max_instances=2) as ctx:
        ctx.load_mesh(z["vtx8"], z["joints"], z["weights"], z["invBind"])
        ctx.load_skeleton(bones)
        rot = np.stack([model.localRotations.reshape(B, 4), np.tile(np.array([0, 0, 0, 1], np.float32), (B, 1))])
        ctx.set_local_rotations(rot)
        got = ctx.read_world_matrices(0)
        assert rel_err(got, ref) <= 2e-6, rel_err(got, ref)
        rest = ctx.read_world_matrices(1)                                # identity pose: world = bind translation chain
        model.rotateBones(names, [Quat(0, 0, 0, 1)] * B, 0)
        model.evaluatePose()
        assert rel_err(rest, model.getBoneWorldMatrices().reshape(B, 16)) <= 2e-6
        # and from host-uploaded world matrices it is the identity map
        ctx.set_palettes(np.stack([ref, ref]))
        assert rel_err(ctx.read_world_matrices(1), ref) <= 2e-6
