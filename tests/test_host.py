"""Host logic (no GPU): pose runtime, tween/timer semantics of the facade, oracle cross-checks,
and the C-ABI library surface (loads, exports every declared symbol, fails loudly without a device)."""
import math
import os
import re

import numpy as np
import pytest

from helpers import numpy_blend_f64, random_pmx, rel_err, write_vmd
from reze_engine_b200 import Engine, Model, PmxLoader, Quat, VMDLoader, capi, crowd, synth
from reze_engine_b200.engine import ManualClock
from reze_engine_b200.math3d import Mat4, Vec3, easeInOut

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_math_conventions():
    # column-major, v' = M v, translation in 12..14 (math.ts:303-320, 472-477)
    t = Mat4.identity().translateInPlace(1, 2, 3)
    r = Mat4.fromQuat(*Quat.fromEuler(0, math.pi / 2, 0).toArray())
    m = t.multiply(r)
    assert np.allclose(m.values[12:15], [1, 2, 3])
    q = m.toQuat()
    assert abs(abs(q.y) - math.sin(math.pi / 4)) < 1e-6
    assert easeInOut(0) == 0 and easeInOut(1) == 1 and easeInOut(0.25) == 0.125 and easeInOut(0.75) == 0.875
    a, b = Quat(0, 0, 0, 1), Quat(0, math.sin(0.4), 0, math.cos(0.4))
    h = Quat.slerp(a, b, 0.5)
    assert abs(h.y - math.sin(0.2)) < 1e-12
    n = Quat.slerp(a, Quat(0, -math.sin(0.4), 0, -math.cos(0.4)), 0.5)      # shortest arc flips b
    assert abs(n.y - math.sin(0.2)) < 1e-12
    c = Quat.slerp(a, Quat(0, 1e-3, 0, 1).normalize(), 0.5)                  # lerp branch above 0.9995
    assert abs(c.length() - 1) < 1e-12


def test_tpose_world_is_bind_and_skin_is_identity(orc):
    wl = synth.make_workload(500, 40)
    m = wl.model(clock=ManualClock())
    m.evaluatePose()
    skin = orc.skin_matrices(m.getBoneWorldMatrices(), wl.invBind)
    assert np.array_equal(skin, np.tile(np.eye(4, dtype=np.float32).reshape(-1), (40, 1)))   # exactly identity
    pos, nrm = orc.deform(wl.vtx8, wl.joints, wl.weights, skin)
    assert rel_err(pos, wl.vtx8[:, :3]) < 1e-6


def test_rotate_bones_tween_semantics():
    clock = ManualClock()
    wl = synth.make_workload(64, 8)
    m = wl.model(clock=clock)
    target = Quat(0, 0.6, 0, 1)
    m.rotateBones(["bone3", "nope"], [target, target], 1000)
    clock.advance(250)
    m.evaluatePose()
    e = easeInOut(0.25)
    want = Quat.slerp(Quat(0, 0, 0, 1), target.normalize(), e)
    got = m.localRotations[12:16]
    assert np.allclose(got, np.float32([want.x, want.y, want.z, want.w]))
    # retarget mid-tween: start becomes the interpolated value *now* (model.ts:275-301)
    m.rotateBones(["bone3"], [Quat(0.5, 0, 0, 1)], 500)
    assert np.allclose(m._startQuat[12:16], np.float32([want.x, want.y, want.z, want.w]), atol=1e-7)
    clock.advance(500)
    m.evaluatePose()
    assert m._active[3] == 0
    assert np.allclose(m.localRotations[12:16], np.float32(Quat(0.5, 0, 0, 1).normalize().toArray()))
    # duration 0 writes immediately and cancels
    m.rotateBones(["bone2"], [Quat(0, 0, 1, 0)], 0)
    assert m.localRotations[8:12].tolist() == [0, 0, 1, 0]


def test_append_rotation_and_batch_evaluator_bit_exact():
    rng = np.random.default_rng(5)
    data, *_ = random_pmx(rng, V=50, B=12)
    m = PmxLoader.loadFromBuffer(data, clock=ManualClock())
    B = len(m.skeleton.bones)
    assert any(b.appendRotate for b in m.skeleton.bones)
    qa, qb = synth.make_pose_keys(B, rng)
    lr = crowd.tween_pose_batch(qa, qb, np.array([0.0, 0.3, 0.77, 1.0]))
    batch = crowd.world_matrices_batch(m.skeleton.bones, lr)
    for p in range(4):
        m.localRotations[:] = lr[p].reshape(-1)
        m.computeWorldMatrices()
        assert np.array_equal(m.worldMatrices.view(np.uint32), batch[p].reshape(-1).view(np.uint32))
    # a negative ratio conjugates the append parent's rotation (model.ts:372-377)
    b = next(b for b in m.skeleton.bones if b.appendRotate and b.appendRatio == -1.0) if any(
        bb.appendRotate and bb.appendRatio == -1.0 for bb in m.skeleton.bones) else None
    if b is not None:
        assert max(-1.0, min(1.0, b.appendRatio)) == -1.0


def test_play_animation_schedules_like_the_reference():
    clock = ManualClock()
    rng = np.random.default_rng(2)
    data, *_ = random_pmx(rng, V=30, B=6, n_morph=0, with_sdef=False)
    eng = Engine(None, {"ambient": 1.0}, clock=clock)
    model = PmxLoader.loadFromBuffer(data, clock=clock)
    eng.currentModel, eng.models = model, [model]          # no GPU here: drive the host side only
    q1, q2 = Quat(0, 0.3, 0, 0.95).normalize(), Quat(0.2, 0, 0, 0.98).normalize()
    vmd = write_vmd([("骨1", 0, (0, 0, 0, 1)), ("骨1", 30, q1.toArray()), ("骨1", 45, q2.toArray()), ("骨2", 15, q1.toArray())])
    eng.animationFrames = VMDLoader.loadFromBuffer(vmd)
    model.rotateBones(["骨4"], [q2], 0)
    eng.playAnimation({"breathBones": {"骨1": 0.05}, "breathDuration": 1000})
    assert model.localRotations[16:20].tolist() == [0, 0, 0, 1]          # bones without a t=0 key are reset
    assert model._active[1] == 1 and model._durationMs[1] == 1000        # key 0 -> 30 frames tween starts now
    assert model._active[2] == 1 and model._durationMs[2] == 500         # first key at t>0: duration = t
    clock.advance(1000); eng._pumpTimers(); model.evaluatePose()
    assert np.allclose(model.localRotations[4:8], np.float32(q1.toArray()), atol=1e-6)
    assert model._active[1] == 1 and model._durationMs[1] == 500         # 30 -> 45 frames, scheduled at t=1s
    clock.advance(500); eng._pumpTimers(); model.evaluatePose()
    assert np.allclose(model.localRotations[4:8], np.float32(q2.toArray()), atol=1e-6)
    clock.advance(200); eng._pumpTimers()                                 # maxTime + 200 ms: breathing starts (exhale first)
    assert model._active[1] == 1 and model._durationMs[1] == 500
    want = q2.multiply(Quat.fromEuler(-0.05, 0, 0))
    assert np.allclose(model._targetQuat[4:8], np.float32(want.normalize().toArray()), atol=1e-6)
    eng.stopAnimation()
    assert not eng.playingAnimation


def test_vmd_morph_tracks_are_kept_and_interpolated(tmp_path):
    clock = ManualClock()
    data, *_ = random_pmx(np.random.default_rng(8), V=40, B=5, n_morph=3, with_sdef=False)
    eng = Engine(None, None, instances=2, clock=clock)
    eng.currentModel = PmxLoader.loadFromBuffer(data, clock=clock)
    p = tmp_path / "a.vmd"
    p.write_bytes(write_vmd([("骨1", 0, (0, 0, 0, 1))], [("m1", 0, 0.2), ("m1", 30, 1.0), ("m1", 60, 0.0), ("unknown", 5, 1.0)]))
    eng.loadAnimation(str(p))
    assert np.allclose(eng._morphTracks["m1"], [(0.0, 0.2), (1000.0, 1.0), (2000.0, 0.0)])       # weights are f32 in the file
    ids, w = eng._evalMorphTracks(np.array([500.0, 2500.0]))
    assert ids.tolist() == [eng.currentModel.morphs.names.index("m1")]
    assert np.allclose(w[:, 0], [0.6, 0.0])


def test_oracle_against_independent_numpy_restatement(orc):
    wl = synth.make_workload(3000, 48)
    world = synth.make_palettes(wl.bones, 2, np.random.default_rng(4))[1]
    skin64 = orc.skin_matrices(world, wl.invBind, np.float64)
    p64, n64 = orc.deform(wl.vtx8, wl.joints, wl.weights, skin64, dtype=np.float64)
    pn, nn = numpy_blend_f64(wl.vtx8, wl.joints, wl.weights, skin64)
    assert rel_err(p64, pn) < 1e-12 and rel_err(n64, nn) < 1e-12
    p32, n32 = orc.deform(wl.vtx8, wl.joints, wl.weights, orc.skin_matrices(world, wl.invBind))
    assert rel_err(p32, p64) < 2e-6 and rel_err(n32, n64) < 2e-6        # f32 oracle vs f64 oracle (reported in DESIGN.md)


def test_oracle_morph_and_sdef_against_independent_numpy_restatement(orc):
    """The two features the reference cannot pin (vertex morphs, SDEF) are at least pinned by two independent
    restatements of SURVEY 8c agreeing: the C oracle (f64 variant) and a vertex-by-vertex numpy one."""
    from helpers import numpy_morph_sdef_f64
    wl = synth.make_workload(1500, 32, M=6, sdef=True, seed=9)
    rng = np.random.default_rng(9)
    world = synth.make_palettes(wl.bones, 2, rng)[1]
    mw = rng.uniform(-0.5, 1.0, wl.morphs.count).astype(np.float32)
    skin64 = orc.skin_matrices(world, wl.invBind, np.float64)
    morph = (wl.morphs.offsets, wl.morphs.vertexIndex, wl.morphs.delta)
    sd = (wl.sdef.vertexIndex, wl.sdef.c_r0_r1)
    for m, s_ in ((morph, None), (None, sd), (morph, sd)):
        po, no = orc.deform(wl.vtx8, wl.joints, wl.weights, skin64, morph=m, morphW=mw if m else None, sdef=s_, dtype=np.float64)
        pn, nn = numpy_morph_sdef_f64(wl.vtx8, wl.joints, wl.weights, skin64, morph=m, morphW=mw if m else None, sdef=s_)
        assert rel_err(po, pn) < 1e-9 and rel_err(no, nn) < 1e-9, (m is not None, s_ is not None, rel_err(po, pn), rel_err(no, nn))
    assert wl.sdef.vertexIndex.size > 50 and wl.morphs.vertexIndex.size > 50


def test_oracle_morph_and_sdef_reduce_to_pinned_path(orc):
    wl = synth.make_workload(2000, 32, M=6, sdef=True)
    world = synth.make_palettes(wl.bones, 1, np.random.default_rng(9))[0]
    skin = orc.skin_matrices(world, wl.invBind)
    base = orc.deform(wl.vtx8, wl.joints, wl.weights, skin)
    morph = (wl.morphs.offsets, wl.morphs.vertexIndex, wl.morphs.delta)
    zero = orc.deform(wl.vtx8, wl.joints, wl.weights, skin, morph=morph, morphW=np.zeros(6, np.float32))
    assert np.array_equal(zero[0], base[0]) and np.array_equal(zero[1], base[1])      # M = 0 equals the pinned path
    # SDEF with both bones carrying the same rigid transform equals the linear blend up to rounding
    same = np.tile(world[5], (wl.B, 1))
    inv0 = np.tile(np.eye(4, dtype=np.float32).reshape(-1), (wl.B, 1))
    sk = orc.skin_matrices(same, inv0)
    lin = orc.deform(wl.vtx8, wl.joints, wl.weights, sk)
    sd = orc.deform(wl.vtx8, wl.joints, wl.weights, sk, sdef=(wl.sdef.vertexIndex, wl.sdef.c_r0_r1))
    assert rel_err(sd[0], lin[0]) < 5e-6 and rel_err(sd[1], lin[1]) < 5e-6


def test_vertex_edge_sizes_follow_the_outline_draw_rule(orc):
    """Engine.vertexEdgeSizes: a vertex takes Material.edgeSize of the material slice of the index buffer that draws it, only
    when that material is outlined ((edgeFlag & 0x10) and edgeSize > 0, engine.ts:2024); the oracle's hull follows
    engine.ts:458-461 term by term."""
    from reze_engine_b200 import Engine
    from reze_engine_b200.model import Bone, Model, Skeleton, Skinning
    V = 12
    vtx = np.zeros((V, 8), np.float32)
    idx = np.array([0, 1, 2, 3, 4, 5, 5, 6, 7, 8, 9, 10], np.uint32)            # vertex 11 is drawn by nothing, 5 by two materials
    mats = [dict(vertexCount=3, edgeFlag=0x10, edgeSize=1.5), dict(vertexCount=3, edgeFlag=0x00, edgeSize=2.0),
            dict(vertexCount=3, edgeFlag=0x10, edgeSize=0.5), dict(vertexCount=3, edgeFlag=0x10, edgeSize=0.0)]
    bones = [Bone(name="root", parentIndex=-1, bindTranslation=[0.0, 0.0, 0.0])]
    sk = Skeleton(bones=bones, inverseBindMatrices=np.tile(np.eye(4, dtype=np.float32).reshape(1, 16), (1, 1)))
    skin = Skinning(joints=np.zeros(V * 4, np.uint16), weights=np.tile(np.array([255, 0, 0, 0], np.uint8), V))
    model = Model(vtx.reshape(-1), idx, [], mats, sk, skin)
    e = Engine.vertexEdgeSizes(model)
    assert e.dtype == np.float32 and e.tolist() == [1.5, 1.5, 1.5, 0, 0, 0.5, 0.5, 0.5, 0, 0, 0, 0]
    pos = np.arange(V * 3, dtype=np.float32).reshape(V, 3)
    nrm = np.tile(np.array([0.0, 0.6, 0.8], np.float32), (V, 1))
    hull = orc.outline_hull(pos, nrm, e)
    assert hull.dtype == np.float32 and np.array_equal(hull[3], pos[3])
    assert np.allclose(hull[0] - pos[0], [0, 0.009, 0.012], atol=1e-6)
    st = orc.interleaved(pos, nrm, vtx)
    assert st.shape == (V, 8) and np.array_equal(st[:, :3], pos) and np.array_equal(st[:, 6:], vtx[:, 6:])


def _random_bodies(rng, bones, inv_bind, n):
    from reze_engine_b200 import physics_bridge as pb
    B = len(bones)
    bone_index = rng.integers(-1, B, n).astype(np.int32)
    bone_index[1] = bone_index[0] = max(int(bone_index[0]), 0)          # two bodies on one bone
    dynamic = (rng.random(n) < 0.6).astype(np.uint8)
    dynamic[0] = dynamic[1] = 1
    shape_pos = rng.normal(0, 3, (n, 3))
    shape_rot = rng.uniform(-1.5, 1.5, (n, 3))
    off, inv = pb.compute_body_offsets(inv_bind, bone_index, shape_pos, shape_rot)
    return bone_index, dynamic, off, inv


def test_physics_bridge_follows_the_reference_feedback_rule():
    """physics.ts:560-585 / 714-751 restated on the host: offsets invert cleanly, only DYNAMIC bodies with a valid bone write,
    bodies apply in index order (the last valid one wins), NaN / huge matrices are skipped, children are left alone."""
    from reze_engine_b200 import physics_bridge as pb
    rng = np.random.default_rng(21)
    wl = synth.make_workload(64, 24, seed=21)
    ib = np.asarray(wl.invBind, np.float32).reshape(-1, 16)
    n = 10
    bone_index, dynamic, off, inv = _random_bodies(rng, wl.bones, ib, n)
    for i in range(n):
        prod = Mat4(off[i]).multiply(Mat4(inv[i])).values.reshape(4, 4)
        assert np.abs(prod - np.eye(4)).max() < 2e-5
    world = synth.make_palettes(wl.bones, 1, rng)[0].copy()
    before = world.copy()
    pos = rng.normal(0, 5, (n, 3))
    quat = rng.normal(size=(n, 4))
    quat /= np.linalg.norm(quat, axis=1, keepdims=True)
    pos[1] = np.nan                                                     # the second body of bone b0 is invalid -> body 0 stays
    wrote = pb.apply_bodies_to_bones(world, bone_index, dynamic, inv, pos, quat)
    driven = {int(b) for b, d in zip(bone_index, dynamic) if d and b >= 0}
    assert wrote >= len(driven)
    for b in range(len(wl.bones)):
        if b not in driven:
            assert np.array_equal(world[b], before[b])                  # static / kinematic bodies and children: untouched
    b0 = int(bone_index[0])
    last = max(i for i in range(n) if dynamic[i] and bone_index[i] == b0 and i != 1)
    want = Mat4.fromPositionRotation(Vec3(*pos[last]), Quat(*quat[last])).multiply(Mat4(inv[last])).values
    assert np.array_equal(world[b0], want)
    # a body sitting exactly where the bone's bind shape is reproduces the bone's own world matrix
    one = np.array([b0], np.int32)
    sp, sr = rng.normal(0, 2, (1, 3)), rng.uniform(-1, 1, (1, 3))
    o1, i1 = pb.compute_body_offsets(ib, one, sp, sr)
    node = Mat4(before[b0]).multiply(Mat4(o1[0]))                       # nodeWorld = boneWorld x bodyOffset (physics.ts:600-603)
    w2 = before.copy()
    pb.apply_bodies_to_bones(w2, one, np.ones(1, np.uint8), i1, [list(node.getPosition().__dict__.values())] if hasattr(node.getPosition(), "__dict__") else [[node.values[12], node.values[13], node.values[14]]],
                             [node.toQuat().toArray()])
    assert np.abs(w2[b0] - before[b0]).max() < 5e-5


def test_bench_reference_arm_prints_one_contract_line(orc):
    """`bench.py --impl reference` (the CPU port of the reference arithmetic, the arm the driver times next to the GPU
    path): exactly one JSON line on stdout with the contract's keys, no GPU needed."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                          "--verts", "3000", "--instances", "16", "--bones", "32", "--cpu-seconds", "0.3"],
                         capture_output=True, text=True, timeout=300, check=True).stdout.strip().splitlines()
    assert len(out) == 1
    r = json.loads(out[0])
    assert r["impl"] == "reference" and r["metric"] == "skinned_vertices_per_sec" and r["unit"] == "verts/s" and r["higher_is_better"] is True
    assert r["n_gpus"] == 1 and r["steps"] == 2 and r["warmup"] == 1 and r["value"] > 0 and r["vs_baseline"] is None
    assert r["cpu_baseline"]["kind"] == "port" and r["cpu_baseline"]["cores"] >= 1 and r["cpu_baseline"]["value"] == r["value"]
    assert r["e2e"] == {"value": r["value"], "unit": "verts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert r["config"]["V"] == 3000 and r["config"]["K_per_gpu"] == 16 and "workload" in r["config"]


def test_header_is_plain_c_and_matches_the_ctypes_mirrors(tmp_path):
    """include/rze_b200.h is the drop-in boundary: it must compile as plain C99 (what cgo / N-API / ctypes generators consume),
    and the struct layouts the Python mirror declares must be the compiler's."""
    import ctypes as C
    import subprocess
    src = tmp_path / "abi.c"
    src.write_text('''
#include <stdio.h>
#include <stddef.h>
#include "rze_b200.h"
int main(void) {
  printf("%zu %zu %zu ", sizeof(rz_config), sizeof(rz_stats), sizeof(rz_output_layout));
  printf("%zu %zu %zu %zu ", offsetof(rz_config, stream), offsetof(rz_config, tune_ctas_per_sm), offsetof(rz_stats, frames), offsetof(rz_stats, fastGatherPermille));
  printf("%zu %zu %zu\\n", offsetof(rz_output_layout, vertexStride), offsetof(rz_output_layout, hullOffset), offsetof(rz_output_layout, uvOffset));
  printf("%u %u %u %u %u %u %u %d\\n", RZ_FLAG_SDEF, RZ_FLAG_NO_NORMALS, RZ_FLAG_BOUNDS, RZ_FLAG_REORDER_VERTICES, RZ_FLAG_OUTLINE,
         RZ_FLAG_INTERLEAVED, RZ_FLAG_DOUBLE_BUFFER, RZE_B200_ABI_VERSION);
  return 0;
}
''')
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    nums = [int(x) for x in out]
    assert nums[:3] == [C.sizeof(capi.RzConfig), C.sizeof(capi.RzStats), C.sizeof(capi.RzOutputLayout)]
    assert nums[3:7] == [capi.RzConfig.stream.offset, capi.RzConfig.tune_ctas_per_sm.offset, capi.RzStats.frames.offset,
                         capi.RzStats.fastGatherPermille.offset]
    assert nums[7:10] == [capi.RzOutputLayout.vertexStride.offset, capi.RzOutputLayout.hullOffset.offset, capi.RzOutputLayout.uvOffset.offset]
    assert nums[10:17] == [capi.RZ_FLAG_SDEF, capi.RZ_FLAG_NO_NORMALS, capi.RZ_FLAG_BOUNDS, capi.RZ_FLAG_REORDER_VERTICES, capi.RZ_FLAG_OUTLINE,
                           capi.RZ_FLAG_INTERLEAVED, capi.RZ_FLAG_DOUBLE_BUFFER]
    assert nums[17] == 3


def test_napi_shim_type_checks_against_the_c_abi(rzlib, tmp_path):
    """napi/rze_b200_napi.cc cannot be linked or run here (no Node.js), but it must at least COMPILE: against a minimal
    stand-in for <node_api.h> (tests/mock_node_api, public N-API signatures) and the real include/rze_b200.h, warnings as
    errors; and every rz_* symbol it references must be exported by the library."""
    import subprocess
    obj = tmp_path / "shim.o"
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-fPIC", "-c", "-I", os.path.join(ROOT, "tests", "mock_node_api"),
                    "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "napi", "rze_b200_napi.cc"), "-o", str(obj)], check=True)
    syms = subprocess.run(["nm", "-u", str(obj)], capture_output=True, text=True, check=True).stdout.split()
    used = sorted({x for x in syms if x.startswith("rz_")})
    assert len(used) >= 18 and set(used) <= set(capi.EXPORTS), set(used) - set(capi.EXPORTS)
    defined = subprocess.run(["nm", "--defined-only", str(obj)], capture_output=True, text=True, check=True).stdout
    assert "napi_register_module_v1" in defined


def test_typescript_facade_matches_the_shim_and_the_reference_api():
    """ts/engine.ts cannot be compiled here (no tsc / Node.js).  What can be checked statically: every `rz.<name>(` call it
    makes is a function the N-API shim registers under that name (and with that many arguments), every public method of the
    reference's Engine that belongs to the replaced path exists with the reference's name (engine/src/engine.ts:157, 1419,
    1425, 1593, 1664, 1668, 1684, 1692, 1704, 1723, 2124), and dispose() releases the context."""
    ts = open(os.path.join(ROOT, "ts", "engine.ts")).read()
    ts = re.sub(r"//[^\n]*", "", ts)                                          # (comments quote calls in short form)
    shim = open(os.path.join(ROOT, "napi", "rze_b200_napi.cc")).read()
    registered = dict(re.findall(r'RZ_FN\("(\w+)",\s*(\w+)\)', shim))
    assert len(registered) >= 27
    arity = {}
    for js_name, fn in registered.items():
        m = re.search(r"napi_value %s\(napi_env env, napi_callback_info info\)\s*\{(?:\s*return palette_feed\(|\s*ARGS\((\d+)\))" % fn, shim)
        assert m, fn
        arity[js_name] = int(m.group(1)) if m.group(1) else 5               # the three palette feeds share a 5-argument helper
    calls = re.findall(r"\brz\.(\w+)\(", ts)
    assert calls and set(calls) <= set(registered), set(calls) - set(registered)
    for need in ("create", "destroy", "loadMesh", "loadMorphs", "loadSdef", "loadEdgeSize", "loadSkeleton", "setPalettes", "setLocalRotations", "setTweens",
                 "setInstanceClocks", "loadAnimation", "setMorphWeights", "deform", "sync", "readInstance", "readInstanceAsync", "readWait", "getStats",
                 "loadRigidBodies", "applyBodyTransforms", "readWorldMatrices"):
        assert need in calls, need
    # argument counts of the calls (top-level commas inside the parentheses)
    for m in re.finditer(r"\brz\.(\w+)\(", ts):
        depth, i, nargs, seen = 1, m.end(), 0, False
        while depth:
            ch = ts[i]
            if ch in "([{":
                depth += 1
            elif ch in ")]}":
                depth -= 1
            elif ch == "," and depth == 1:
                nargs += 1
            if depth and not ch.isspace():
                seen = True
            i += 1
        nargs = nargs + 1 if seen else 0
        assert nargs == arity[m.group(1)], (m.group(1), nargs, arity[m.group(1)])
    for method in ("async init()", "async loadModel(path: string)", "async loadAnimation(url: string)", "playAnimation(", "stopAnimation()", "rotateBones(",
                   "render()", "runRenderLoop(", "stopRenderLoop()", "getStats()", "dispose()"):
        assert "public " + method in ts, method
    body = ts[ts.index("public dispose()"):]
    assert "rz.destroy(this.ctx)" in body[:body.index("\n  }")]
    assert "class MultiDeviceEngine" in ts


def test_capi_exports_every_declared_symbol(rzlib):
    from reze_engine_b200 import capi
    hdr = open(os.path.join(ROOT, "include", "rze_b200.h")).read()
    declared = set(re.findall(r"\b(rz_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(capi.EXPORTS)
    for name in declared:
        assert hasattr(rzlib, name), name
    assert rzlib.rz_abi_version() == 3
    import ctypes as C
    assert C.sizeof(capi.RzConfig) == 56 and C.sizeof(capi.RzStats) == 136


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_lane_plan_preserves_every_influence_and_packs_pairs(rzlib, mode):
    """rz_plan_lanes (host only): whatever lane / slot a (bone, weight) pair ends up in, every vertex keeps exactly its
    shader-normalised non-zero influences (engine.ts:255-258); zero-weight slots stay in range; the fast-path count the
    library reports is really there (aligned lane pairs gather the same row in those slots)."""
    from reze_engine_b200 import capi
    wl = synth.make_workload(5000, 96, seed=5)
    J, W = wl.joints.copy(), wl.weights.copy()
    W[7] = 0                                   # weight sum 0 -> (1,0,0,0) on joint 0 of the vertex
    J[11], W[11] = (3, 3, 9, 0), (100, 55, 100, 0)   # a bone listed twice keeps the caller's slots
    W[13] = (0, 200, 0, 55)                    # interior zero weights
    V, B = J.shape[0], 96
    plan = capi.plan_lanes(J, W, B, mode)
    lv, lj, lw = plan["laneVertex"], plan["laneJoints"], plan["laneWeights"]
    Vp = lv.size
    assert Vp % 256 == 0 and Vp >= V and lj.max() < B
    real = lv != 0xFFFFFFFF
    assert np.array_equal(np.sort(lv[real]), np.arange(V))
    for w0 in range(0, Vp, 32):                # a warp owns 32 consecutive vertices
        mine = lv[w0:w0 + 32]
        assert set(mine[mine != 0xFFFFFFFF].tolist()) == set(range(w0, min(w0 + 32, V)))
    wf = W.astype(np.float32) / np.float32(255.0)
    ssum = (wf[:, 0] + wf[:, 1]) + wf[:, 2] + wf[:, 3]
    with np.errstate(divide="ignore", invalid="ignore"):
        norm = np.where(ssum[:, None] > 1e-4, wf * (np.float32(1.0) / ssum)[:, None], np.array([1, 0, 0, 0], np.float32)).astype(np.float32)
    for p in np.nonzero(real)[0]:
        v = int(lv[p])
        want = sorted((int(J[v, k]), float(norm[v, k])) for k in range(4) if norm[v, k] != 0)
        got = sorted((int(lj[p, s]), float(lw[p, s])) for s in range(4) if lw[p, s] != 0)
        assert want == got, (p, v, want, got)
        if mode < 2:
            assert all(lw[p, k] == norm[v, k] for k in range(4))
    pad = ~real
    assert np.all(lw[pad, 0] == 1.0) and np.all(lw[pad, 1:] == 0.0)
    # fast-path accounting
    coherent = 0
    for w0 in range(0, Vp, 32):
        n = int(max(1, max(int(np.max(np.nonzero(lw[p])[0]) + 1) for p in range(w0, w0 + 32))))
        for s in range(n):
            coherent += bool(np.all(lj[w0:w0 + 32:2, s] == lj[w0 + 1:w0 + 32:2, s]))
    if mode == 2:
        assert plan["total"] > 0 and coherent >= plan["fast"] > 0.5 * plan["total"]
        assert plan["hist"].sum() > 0
    else:
        assert plan["total"] == 0


def _check_plan2(J, W, B, r):
    V = J.shape[0]
    gf, gc, gp = r["groupFirst"].astype(np.int64), r["groupCount"].astype(np.int64), r["groupPaired"]
    assert gf[0] == 0 and np.array_equal(gf[1:], (gf + gc)[:-1]) and gf[-1] + gc[-1] == V          # groups tile the vertex range
    assert (gc[gp == 1] <= 64).all() and (gc[gp == 0] <= 32).all() and (gc > 0).all()
    wf = W.astype(np.float32) / np.float32(255.0)
    ssum = (wf[:, 0] + wf[:, 1]) + wf[:, 2] + wf[:, 3]
    seen = np.zeros(V, np.int64)
    lj = r["laneJoints"]
    assert lj.max() < B
    for side, wkey in (("vertA", "wA"), ("vertB", "wB")):
        vv, ww = r[side], r[wkey]
        for p in np.nonzero(vv != 0xFFFFFFFF)[0]:
            v, g = int(vv[p]), p // 32
            assert gf[g] <= v < gf[g] + gc[g]                                                        # outputs of a group are contiguous
            assert side == "vertA" or gp[g] == 1
            seen[v] += 1
            if ssum[v] > 1e-4:
                inv = np.float32(1.0) / ssum[v]
                want = sorted((int(J[v, k]), float(wf[v, k] * inv)) for k in range(4) if wf[v, k] != 0)
            else:
                want = [(int(J[v, 0]), 1.0)]
            got = sorted((int(lj[p, k]), float(ww[p, k])) for k in range(4) if ww[p, k] != 0)
            assert want == got, (v, want, got)
    assert (seen == 1).all()
    assert not r["wA"][r["vertA"] == 0xFFFFFFFF].any() and not r["wB"][r["vertB"] == 0xFFFFFFFF].any()   # a missing side blends nothing
    # staging slots (deform2_kernel.cuh): a real vertex sits at (vertex - groupFirst); the 64 slots of a group are used exactly
    # once; A-slots are distinct modulo 32 and so are B-slots (12-byte records: 3*slot mod 32 is then a bijection on the banks)
    sa, sb, ln = r["slotA"].astype(np.int64), r["slotB"].astype(np.int64), r["laneSlots"]
    for g in range(len(gf)):
        a, b = sa[g * 32:(g + 1) * 32], sb[g * 32:(g + 1) * 32]
        assert sorted(np.concatenate([a, b]).tolist()) == list(range(64))
        assert len(set((a % 32).tolist())) == 32 and len(set((b % 32).tolist())) == 32
    for side, skey in (("vertA", sa), ("vertB", sb)):
        vv = r[side]
        real = vv != 0xFFFFFFFF
        assert np.array_equal(skey[real], vv[real].astype(np.int64) - gf[np.nonzero(real)[0] // 32])
        assert (skey[~real] >= gc[np.nonzero(~real)[0] // 32]).all()                                  # dummies are never drained
    used = np.maximum((r["wA"] != 0), (r["wB"] != 0))
    want_n = np.where(used.any(axis=1), 4 - np.argmax(used[:, ::-1], axis=1), 1)
    assert np.array_equal(ln.astype(np.int64), want_n)


def test_two_vertices_per_lane_plan_does_not_depend_on_the_thread_count(rzlib, monkeypatch):
    """The slot / quarter-warp stage of rz_plan_lanes2 runs on a few host threads (groups are independent): the plan is the
    same table, bit for bit, whatever RZ_PLAN_THREADS says."""
    wl = synth.make_workload(12_000, 128, seed=9)
    J, W = wl.joints.reshape(-1, 4).copy(), wl.weights.reshape(-1, 4).copy()
    plans = []
    for nt in ("1", "3", "8"):
        monkeypatch.setenv("RZ_PLAN_THREADS", nt)
        plans.append(capi.plan_lanes2(J, W, wl.B, rzlib))
    for other in plans[1:]:
        assert other.keys() == plans[0].keys()
        for k, v in plans[0].items():
            if isinstance(v, np.ndarray):
                assert np.array_equal(v, other[k]), k
            else:
                assert v == other[k], k


def test_two_vertices_per_lane_plan(rzlib):
    """rz_plan_lanes2 (groundwork for the next kernel generation): every vertex evaluated exactly once, by a lane of the group
    that owns its 64-vertex window; both vertices of a lane find all their (bone, weight) pairs among the lane's four rows;
    and on the benchmark mesh the plan needs a third fewer gather instructions than today's, most of them on the fast path."""
    wl = synth.make_workload(20_000, 256, seed=6)
    J, W = wl.joints.reshape(-1, 4).copy(), wl.weights.reshape(-1, 4).copy()
    W[7] = 0                                            # weight sum 0
    J[200], W[200] = (3, 3, 9, 0), (100, 55, 100, 0)    # a bone listed twice: its window falls back to one vertex per lane
    r = capi.plan_lanes2(J, W, wl.B, rzlib)
    _check_plan2(J, W, wl.B, r)
    g200 = int(np.searchsorted(r["groupFirst"], 200, side="right") - 1)
    assert r["groupPaired"][g200] == 0
    one = capi.plan_lanes(J, W, wl.B, 2, rzlib)
    assert r["pairedWindows"] > 8 * r["fallbackWindows"]
    assert r["total"] < 0.75 * one["total"] and r["fast"] > 0.85 * r["total"]
    # fast-path accounting: aligned lane pairs of a packed group really gather the same row in that many slots
    lj = r["laneJoints"].reshape(-1, 32, 4)
    used = ((r["wA"] != 0) | (r["wB"] != 0)).reshape(-1, 32, 4)
    coherent = 0
    for g in range(lj.shape[0]):
        n = int(max(1, used[g].any(axis=0).nonzero()[0].max(initial=0) + 1))
        for s_ in range(n):
            coherent += bool(np.all(lj[g, 0::2, s_] == lj[g, 1::2, s_]))
    assert coherent >= r["fast"]


def _emulate_lanes(vtx8, skin16, lane_vertex, lane_joints, lane_weights):
    """What the deform kernel computes from a lane plan, in numpy f64: M = sum_s w_s * skin[row_s], pos = M [p,1], n = norm(M3 n)."""
    M = np.asarray(skin16, np.float64).reshape(-1, 4, 4).transpose(0, 2, 1)[:, :3, :]           # [B,3,4]
    vt = np.asarray(vtx8, np.float64).reshape(-1, 8)
    V = vt.shape[0]
    pos, nrm = np.full((V, 3), np.nan), np.full((V, 3), np.nan)
    real = np.nonzero(lane_vertex != 0xFFFFFFFF)[0]
    v = lane_vertex[real].astype(np.int64)
    Mb = np.einsum("ls,lsrc->lrc", lane_weights[real].astype(np.float64), M[lane_joints[real].astype(np.int64)])
    pos[v] = np.einsum("lrc,lc->lr", Mb[:, :, :3], vt[v, :3]) + Mb[:, :, 3]
    n = np.einsum("lrc,lc->lr", Mb[:, :, :3], vt[v, 3:6])
    ln = np.linalg.norm(n, axis=1, keepdims=True)
    nrm[v] = np.where(ln > 0, n / np.where(ln > 0, ln, 1), 0)
    return pos, nrm


def test_lane_plans_reproduce_the_blend_when_emulated(rzlib, orc):
    """End-to-end check of the load-time plans without a GPU: the arithmetic the kernel performs on a plan (blend the gathered
    rows with the lane's weights, transform) equals the oracle's blend -- for today's plan (all three modes) and for the
    two-vertices-per-lane groundwork plan."""
    wl = synth.make_workload(6000, 64, seed=12)
    J, W = wl.joints.reshape(-1, 4), wl.weights.reshape(-1, 4)
    world = synth.make_palettes(wl.bones, 1, np.random.default_rng(12))[0]
    skin = orc.skin_matrices(world, wl.invBind, np.float64)
    rp, rn = orc.deform(wl.vtx8, wl.joints, wl.weights, skin, dtype=np.float64)
    for mode in (0, 1, 2):
        pl = capi.plan_lanes(J, W, wl.B, mode, rzlib)
        gp, gn = _emulate_lanes(wl.vtx8, skin, pl["laneVertex"], pl["laneJoints"], pl["laneWeights"])
        assert rel_err(gp, rp) < 2e-6 and rel_err(gn, rn) < 2e-6, mode        # (the lane weights are f32)
    p2 = capi.plan_lanes2(J, W, wl.B, rzlib)
    pa, na = _emulate_lanes(wl.vtx8, skin, p2["vertA"], p2["laneJoints"], p2["wA"])
    pb, nb = _emulate_lanes(wl.vtx8, skin, p2["vertB"], p2["laneJoints"], p2["wB"])
    gp, gn = np.where(np.isnan(pa), pb, pa), np.where(np.isnan(na), nb, na)
    assert not np.isnan(gp).any() and rel_err(gp, rp) < 2e-6 and rel_err(gn, rn) < 2e-6


def test_morph_rows_reproduce_the_morphed_positions_when_emulated(rzlib, orc):
    """The per-warp morph rows, applied the way the kernel applies them (row by row: p += w[morph] * delta in f32, PMX morph
    order, padding rows adding exact zeros), give the oracle's morphed positions to the last ulp or two (the kernel fuses the
    multiply-add, the f32 oracle rounds twice)."""
    wl = synth.make_workload(4000, 32, M=12, seed=14)
    rng = np.random.default_rng(14)
    mw = rng.uniform(-0.5, 1.0, wl.morphs.count).astype(np.float32)
    J, W = wl.joints.reshape(-1, 4), wl.weights.reshape(-1, 4)
    lv = capi.plan_lanes(J, W, wl.B, 2, rzlib)["laneVertex"]
    r = capi.plan_morph_rows(lv, wl.V, wl.morphs.offsets, wl.morphs.vertexIndex, wl.morphs.delta, rzlib)
    P = wl.vtx8.reshape(-1, 8)[:, :3].astype(np.float32).copy()
    for w in range(lv.size // 32):
        d = int(r["depth"][w])
        if not d:
            continue
        block = r["rows"][r["first"][w]:r["first"][w] + d * 32].reshape(d, 32, 4)
        ids = block[:, :, 3].copy().view(np.uint32)
        for l in range(32):
            v = int(lv[w * 32 + l])
            if v == 0xFFFFFFFF:
                continue
            p = P[v].copy()
            for u in range(d):
                p = (mw[ids[u, l]] * block[u, l, :3] + p).astype(np.float32)
            P[v] = p
    # oracle with an identity palette: positions = morphed positions (f32 oracle: p = p + w * d, two roundings; fma has one:
    # compare within 1 ulp of the position magnitude)
    ident = np.tile(np.eye(4, dtype=np.float32).T.reshape(1, 16), (wl.B, 1))
    rp, _ = orc.deform(wl.vtx8, wl.joints, wl.weights, ident, morph=(wl.morphs.offsets, wl.morphs.vertexIndex, wl.morphs.delta), morphW=mw)
    touched = np.unique(wl.morphs.vertexIndex)
    assert np.abs(P[touched] - rp[touched]).max() <= 4e-6 and touched.size > 100
    untouched = np.setdiff1d(np.arange(wl.V), touched)
    assert np.array_equal(P[untouched], wl.vtx8.reshape(-1, 8)[untouched, :3])


def test_sdef_records_reproduce_the_spherical_blend_when_emulated(rzlib, orc):
    """The 48-byte SDEF records + per-warp descriptors, evaluated the way the kernel's dense phase does, give the oracle's SDEF
    result for exactly the vertices the descriptors name (every SDEF vertex once)."""
    from helpers import sdef_from_record
    wl = synth.make_workload(3000, 40, sdef=True, seed=15)
    J, W = wl.joints.reshape(-1, 4), wl.weights.reshape(-1, 4)
    world = synth.make_palettes(wl.bones, 1, np.random.default_rng(15))[0]
    skin = orc.skin_matrices(world, wl.invBind, np.float64)
    M = skin.reshape(-1, 4, 4).transpose(0, 2, 1)
    rp, rn = orc.deform(wl.vtx8, wl.joints, wl.weights, skin, sdef=(wl.sdef.vertexIndex, wl.sdef.c_r0_r1), dtype=np.float64)
    lv = capi.plan_lanes(J, W, wl.B, 2, rzlib)["laneVertex"]
    r = capi.plan_sdef(lv, J, W, wl.B, wl.sdef.vertexIndex, wl.sdef.c_r0_r1, rzlib)
    vt = wl.vtx8.reshape(-1, 8).astype(np.float64)
    done = []
    for w0 in range(0, lv.size, 32):
        for word in r["desc"][w0:w0 + 32]:
            if word == 0xFFFFFFFF:
                break
            idx, v = int(word) & 0xFFFFFF, w0 + (int(word) >> 24)
            pos, nrm = sdef_from_record(r["records"][idx], M, vt[v, :3], vt[v, 3:6])
            assert np.abs(pos - rp[v]).max() <= 1e-5 and np.abs(nrm - rn[v]).max() <= 1e-6, (v, np.abs(pos - rp[v]).max())
            done.append(v)
    assert sorted(done) == sorted(wl.sdef.vertexIndex.tolist()) and len(done) > 100


def test_two_vertices_per_lane_plan_properties(rzlib):
    """hypothesis: adversarial tiny tables through rz_plan_lanes2."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.integers(1, 200), st.integers(1, 9), st.integers(0, 2 ** 32 - 1))
    def run(V, B, seed):
        rng = np.random.default_rng(seed)
        J = rng.integers(0, B, (V, 4)).astype(np.uint16)
        W = rng.integers(0, 256, (V, 4)).astype(np.uint8)
        W[rng.random((V, 4)) < 0.5] = 0
        _check_plan2(J, W, B, capi.plan_lanes2(J, W, B, rzlib))
    run()


def test_lane_plan_properties_on_adversarial_tables(rzlib):
    """hypothesis: arbitrary tiny skinning tables (one bone, duplicate bones, all-zero weights, sums != 255, ragged vertex
    counts) through rz_plan_lanes + rz_plan_palette_rows + rz_plan_morph_rows: never crash, every vertex evaluated exactly
    once by a lane of its own 32-vertex warp, its non-zero shader-normalised influences preserved, palette rows a permutation,
    every morph entry present once."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.integers(1, 130), st.integers(1, 9), st.integers(0, 2 ** 32 - 1), st.integers(0, 2))
    def run(V, B, seed, mode):
        rng = np.random.default_rng(seed)
        J = rng.integers(0, B, (V, 4)).astype(np.uint16)
        W = rng.integers(0, 256, (V, 4)).astype(np.uint8)
        W[rng.random((V, 4)) < 0.45] = 0                          # many zero weights, some all-zero rows
        plan = capi.plan_lanes(J, W, B, mode, rzlib)
        lv, lj, lw = plan["laneVertex"], plan["laneJoints"], plan["laneWeights"]
        real = lv != 0xFFFFFFFF
        assert np.array_equal(np.sort(lv[real]), np.arange(V)) and lj.max() < B
        wf = W.astype(np.float32) / np.float32(255.0)
        ssum = (wf[:, 0] + wf[:, 1]) + wf[:, 2] + wf[:, 3]
        for p in np.nonzero(real)[0]:
            v = int(lv[p])
            assert v // 32 == p // 32
            if ssum[v] > 1e-4:
                inv = np.float32(1.0) / ssum[v]
                want = {}
                for k in range(4):
                    if wf[v, k] != 0:
                        want.setdefault(int(J[v, k]), []).append(float(wf[v, k] * inv))
            else:
                want = {int(J[v, 0]): [1.0]}
            got = {}
            for s_ in range(4):
                if lw[p, s_] != 0:
                    got.setdefault(int(lj[p, s_]), []).append(float(lw[p, s_]))
            assert {k: sorted(x) for k, x in want.items()} == {k: sorted(x) for k, x in got.items()}, (v, want, got)
        pos = capi.plan_palette_rows(lj, B, rzlib)
        assert sorted(pos.tolist()) == list(range(B))
        M = int(rng.integers(0, 5))
        offs, vi, dl = [0], [], []
        for _m in range(M):
            n = int(rng.integers(0, V + 1))
            vi.append(rng.integers(0, V, n).astype(np.uint32))    # duplicates allowed
            dl.append(rng.normal(0, 1, (n, 3)).astype(np.float32) + 3.0)      # (never exactly zero)
            offs.append(offs[-1] + n)
        vi = np.concatenate(vi) if vi else np.zeros(0, np.uint32)
        dl = np.concatenate(dl) if dl else np.zeros((0, 3), np.float32)
        r = capi.plan_morph_rows(lv, V, np.array(offs, np.uint32), vi, dl, rzlib)
        rows = r["rows"]
        assert int((np.abs(rows[:, :3]).sum(axis=1) > 0).sum()) == vi.size
        assert abs(float(rows[:, :3].sum()) - float(dl.sum())) <= 1e-3 * max(1.0, float(np.abs(dl).sum()))
    run()


def test_morph_rows_hold_every_entry_once_in_pmx_order(rzlib):
    """rz_plan_morph_rows (the table builder rz_load_morphs uses, device-free): every (vertex, morph, delta) of the caller's
    table sits in exactly one row entry of the lane that evaluates the vertex, rows of a lane ascend in morph id, everything
    else is an exact-zero delta; coherent warps get morph-major rows, scattered ones the compact format, and a morph that
    lists a vertex twice keeps both entries."""
    from reze_engine_b200 import synth
    rng = np.random.default_rng(12)
    wl = synth.make_workload(3000, 40, seed=12)
    V, M = wl.V, 30
    offs, vidx, deltas = [0], [], []
    for m in range(M):
        if m % 2 == 0:
            a = int(rng.integers(0, V - 300))
            v = np.arange(a, a + int(rng.integers(40, 300)))
        else:
            v = np.sort(rng.choice(V, int(rng.integers(5, 150)), replace=False))
        if m == 4:
            v = np.concatenate([v, v[:2]])
        vidx.append(v.astype(np.uint32))
        deltas.append(rng.normal(0, 0.05, (v.size, 3)).astype(np.float32))
        offs.append(offs[-1] + v.size)
    offs, vidx, deltas = np.array(offs, np.uint32), np.concatenate(vidx), np.concatenate(deltas)
    mid = np.repeat(np.arange(M), np.diff(offs.astype(np.int64)))
    plan = capi.plan_lanes(wl.joints, wl.weights, wl.B, 2, rzlib)
    lv = np.asarray(plan["laneVertex"], np.uint32)
    r = capi.plan_morph_rows(lv, V, offs, vidx, deltas, rzlib)
    rows, first, depth, mm = r["rows"], r["first"], r["depth"], r["morphMajor"]
    assert mm.any() and not mm[depth > 0].all()                    # both formats occur
    want = {}
    for e in range(vidx.size):
        want.setdefault(int(vidx[e]), []).append((int(mid[e]), deltas[e].tobytes()))
    seen = 0
    for w in range(lv.size // 32):
        block = rows[first[w]:first[w] + depth[w] * 32].reshape(depth[w], 32, 4)
        ids = block[:, :, 3].copy().view(np.uint32)
        if mm[w] and depth[w]:
            assert (ids == ids[:, :1]).all() and (np.diff(ids[:, 0].astype(np.int64)) > 0).all()      # one morph per row, ascending
        for l in range(32):
            v = int(lv[w * 32 + l])
            got = [(int(ids[u, l]), block[u, l, :3].tobytes()) for u in range(depth[w]) if block[u, l, :3].any()]
            if v == 0xFFFFFFFF:
                assert not got
                continue
            exp = [x for x in want.get(v, []) if np.frombuffer(x[1], np.float32).any()]
            assert got == exp, (w, l, v)
            seen += len(exp)
    assert seen == sum(1 for e in range(vidx.size) if deltas[e].any())
    with pytest.raises(capi.RzError):
        capi.plan_morph_rows(lv, V, offs, vidx + V, deltas, rzlib)


def test_palette_permutation_spreads_co_gathered_bones_over_bank_groups(rzlib):
    """rz_plan_palette_rows: a permutation of the bones under which the rows one warp instruction gathers collide less in the
    eight 16-byte bank groups (row r of palette position q lives in group (3q + r) mod 8) than under the identity."""
    wl = synth.make_workload(20000, 256, seed=5)
    plan = capi.plan_lanes(wl.joints, wl.weights, wl.B, 2, rzlib)
    lj = np.asarray(plan["laneJoints"], np.int64).reshape(-1, 32, 4)
    pos = capi.plan_palette_rows(plan["laneJoints"], wl.B, rzlib).astype(np.int64)
    assert sorted(pos.tolist()) == list(range(wl.B))

    def collisions(perm):
        total = 0
        for k in range(4):
            rows = perm[lj[:, :, k]]                                   # [warps, 32] palette positions gathered by slot k
            for w in range(rows.shape[0]):
                u = np.unique(rows[w])
                total += int((np.bincount((3 * u) % 8, minlength=8) - 1).clip(min=0).sum())
        return total
    ident = np.arange(wl.B)
    c_id, c_perm = collisions(ident), collisions(pos)
    assert c_perm < 0.8 * c_id, (c_id, c_perm)
    with pytest.raises(capi.RzError):
        capi.plan_palette_rows(np.full((256, 4), wl.B, np.uint16), wl.B, rzlib)


def test_sdef_tables_and_descriptors(rzlib):
    """rz_plan_sdef (the table builder of rz_load_sdef, device-free): one record per two-influence SDEF vertex with the
    load-time constants of SURVEY 8c, weights normalised like the reference's shader, and per warp a dense descriptor list
    (the l-th word names the l-th SDEF vertex of the warp and its output slot)."""
    rng = np.random.default_rng(31)
    wl = synth.make_workload(2000, 30, sdef=True, seed=31)
    J, W = wl.joints.reshape(-1, 4), wl.weights.reshape(-1, 4)
    vi = np.concatenate([wl.sdef.vertexIndex, np.nonzero(W[:, 2] > 0)[0][:5].astype(np.uint32)])     # + 5 that must be dropped
    vec = np.concatenate([wl.sdef.c_r0_r1.reshape(-1, 9), rng.normal(size=(5, 9)).astype(np.float32)])
    plan = capi.plan_lanes(J, W, wl.B, 2, rzlib)
    lv = np.asarray(plan["laneVertex"], np.uint32)
    r = capi.plan_sdef(lv, J, W, wl.B, vi, vec, rzlib)
    rec, desc = r["records"], r["desc"]
    assert r["active"] == wl.sdef.vertexIndex.size
    seen = set()
    for w0 in range(0, lv.size, 32):
        d = desc[w0:w0 + 32]
        k = int((d != 0xFFFFFFFF).sum())
        assert (d[:k] != 0xFFFFFFFF).all() and (d[k:] == 0xFFFFFFFF).all()                       # dense prefix
        want = sorted(int(v) for v in lv[w0:w0 + 32] if v != 0xFFFFFFFF and int(v) in set(wl.sdef.vertexIndex.tolist()))
        got = []
        for word in d[:k]:
            idx, slot = int(word) & 0xFFFFFF, int(word) >> 24
            v = w0 + slot                                                                        # output slot -> vertex of this warp
            got.append(v)
            i = int(np.nonzero(vi == v)[0][0])
            C_, R0, R1 = vec[i, :3], vec[i, 3:6], vec[i, 6:9]
            w0f, w1f = np.float32(W[v, 0]) / np.float32(255), np.float32(W[v, 1]) / np.float32(255)
            inv = np.float32(1) / (w0f + w1f)
            w0f, w1f = w0f * inv, w1f * inv
            rw = w0f * R0 + w1f * R1
            c0, c1 = (C_ + (C_ + R0 - rw)) * np.float32(0.5), (C_ + (C_ + R1 - rw)) * np.float32(0.5)
            assert np.allclose(rec[idx, :3], C_, atol=0) and np.allclose(rec[idx, 3:6], c0, atol=1e-6) and np.allclose(rec[idx, 6:9], c1, atol=1e-6)
            assert rec[idx, 9] == w0f and rec[idx, 10] == w1f
            assert int(rec[idx, 11:12].view(np.uint32)[0]) == int(J[v, 0]) | (int(J[v, 1]) << 16)
            seen.add(idx)
        assert sorted(got) == want
    assert seen == set(range(r["active"]))


def test_chunk_table_balances_morph_cost(rzlib):
    """rz_plan_chunks: boundaries are monotone multiples of the pass width that cover every tile, no chunk is empty, and chunks
    over the deep (face) tiles are shorter than chunks over plain tiles."""
    depth = np.zeros(782, np.uint32)
    depth[40:130] = 20
    for tpp in (1, 2, 3):
        tab = capi.plan_chunks(depth, tpp, 19, rzlib)
        assert tab[0] == 0 and tab[-1] == depth.size and (np.diff(tab.astype(np.int64)) > 0).all() and tab.size - 1 <= 19
        assert (tab[:-1] % tpp == 0).all()
        sizes = np.diff(tab.astype(np.int64))
        deep = [s for a, s in zip(tab[:-1], sizes) if 40 <= a and a + s <= 130]
        plain = [s for a, s in zip(tab[:-1], sizes) if a >= 130]
        assert deep and plain and max(deep) < min(plain[:-1])
    flat = capi.plan_chunks(np.zeros(100, np.uint32), 2, 10, rzlib)
    assert (np.diff(flat.astype(np.int64)) == 10).all()
    assert capi.plan_chunks(np.zeros(5, np.uint32), 2, 50, rzlib).tolist() == [0, 2, 4, 5]


def test_chunk_table_properties(rzlib):
    """hypothesis: any depth profile, pass width and chunk target give a strictly increasing table from 0 to nTiles whose inner
    boundaries are multiples of the pass width and that never has more chunks than asked for."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=150, deadline=None)
    @given(st.lists(st.integers(0, 40), min_size=1, max_size=300), st.integers(1, 4), st.integers(1, 400))
    def run(depth, tpp, target):
        tab = capi.plan_chunks(np.array(depth, np.uint32), tpp, target, rzlib).astype(np.int64)
        assert tab[0] == 0 and tab[-1] == len(depth) and (np.diff(tab) > 0).all()
        assert (tab[:-1] % tpp == 0).all() and tab.size - 1 <= max(1, min(target, (len(depth) + tpp - 1) // tpp))
    run()


def test_lane_plan_rejects_bad_tables(rzlib):
    from reze_engine_b200 import capi
    with pytest.raises(capi.RzError):
        capi.plan_lanes(np.full((4, 4), 9, np.uint16), np.full((4, 4), 60, np.uint8), B=4)


def test_no_device_fails_loudly_never_falls_back(rzlib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from reze_engine_b200 import capi
    with pytest.raises(capi.RzError) as ei:
        capi.DeformContext(max_instances=1)
    assert ei.value.status == -2 and "no CPU path" in str(ei.value)
    with pytest.raises(capi.RzError):
        Engine(None).init()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "reze-engine_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert "oracle" not in src.lower().replace("oracle convention", ""), f
