"""Physics -> bone feedback, the matrix plumbing only.

The reference's solver is Bullet compiled to wasm (`@fred3d/ammo`, third party, not part of this repository and not
runnable here).  What IS part of the deform path is how its results reach the palette: after `Model.evaluatePose()`,
`Physics.step` overwrites the world matrix of every bone that a DYNAMIC rigid body drives,

    boneWorld = fromPositionRotation(bodyPosition, bodyRotation) x bodyOffsetMatrixInverse      physics.ts:714-751

with `bodyOffsetMatrix = boneInverseBind x fromPositionRotation(shapePosition, fromEuler(shapeRotation))`
(physics.ts:560-585), children are NOT re-evaluated (an in-place edit of the palette input), and matrices that are NaN
or larger than 1e6 are skipped.  This module restates those two functions on the host (the harness side of the
tests); the device side is `rz_load_rigid_bodies` + `rz_apply_body_transforms`, which patch the already computed skin
matrices, so a host that steps the solver only ships 28 bytes per body instead of 64 bytes per bone.
"""
from __future__ import annotations

import numpy as np

from .math3d import Mat4, Quat, Vec3


def compute_body_offsets(inverse_bind, bone_index, shape_position, shape_rotation):
    """physics.ts:560-585.  Returns (bodyOffsetMatrix[n,16], bodyOffsetMatrixInverse[n,16]) column-major f32; bodies without a
    valid bone get identities.  The reference inverts with an adjugate (math.ts:484-545, f64 arithmetic on f32 storage); an
    f64 LU inverse rounded to f32 agrees to the last ulp or two, which is all an unpinned path can ask for."""
    ib = np.asarray(inverse_bind, np.float32).reshape(-1, 16)
    B = ib.shape[0]
    bone_index = np.asarray(bone_index, np.int64).reshape(-1)
    pos = np.asarray(shape_position, np.float64).reshape(-1, 3)
    rot = np.asarray(shape_rotation, np.float64).reshape(-1, 3)
    n = bone_index.size
    off = np.tile(np.eye(4, dtype=np.float32).reshape(1, 16), (n, 1))
    inv = off.copy()
    for i in range(n):
        b = int(bone_index[i])
        if b < 0 or b >= B:
            continue
        q = Quat.fromEuler(rot[i, 0], rot[i, 1], rot[i, 2])
        shape_world_bind = Mat4.fromPositionRotation(Vec3(pos[i, 0], pos[i, 1], pos[i, 2]), q)
        m = Mat4(ib[b]).multiply(shape_world_bind)                     # boneInverseBind x shapeWorldBind
        off[i] = m.values
        m4 = m.values.astype(np.float64).reshape(4, 4).T               # column-major storage -> row-major matrix
        inv[i] = np.linalg.inv(m4).T.reshape(16).astype(np.float32)
    return off, inv


def apply_bodies_to_bones(world, bone_index, dynamic, offset_inverse, body_position, body_rotation):
    """physics.ts:714-751, in place on world[B,16] (column-major f32) for ONE pose: bodies in index order, dynamic ones with a
    valid bone only, invalid results skipped.  Returns the number of bones written."""
    world = np.asarray(world)
    B = world.reshape(-1, 16).shape[0]
    w = world.reshape(B, 16)
    oinv = np.asarray(offset_inverse, np.float32).reshape(-1, 16)
    bp = np.asarray(body_position, np.float64).reshape(-1, 3)
    bq = np.asarray(body_rotation, np.float64).reshape(-1, 4)
    written = 0
    for i, b in enumerate(np.asarray(bone_index, np.int64).reshape(-1)):
        if not dynamic[i] or b < 0 or b >= B:
            continue
        node = Mat4.fromPositionRotation(Vec3(bp[i, 0], bp[i, 1], bp[i, 2]), Quat(bq[i, 0], bq[i, 1], bq[i, 2], bq[i, 3]))
        v = node.multiply(Mat4(oinv[i])).values
        if not np.isnan(v[0]) and not np.isnan(v[15]) and abs(v[0]) < 1e6 and abs(v[15]) < 1e6:
            w[b] = v
            written += 1
    return written


def bodies_from_model(model):
    """The arrays rz_load_rigid_bodies wants, from a loaded Model (Model.getRigidbodies(), pmx-loader.ts:603-690):
    (bone_index int32[n], dynamic uint8[n] = RigidbodyType.Dynamic, body_offset float32[n,16], body_offset_inverse float32[n,16])."""
    rbs = model.getRigidbodies()
    n = len(rbs)
    bone_index = np.array([int(r["boneIndex"]) for r in rbs], np.int32).reshape(n)
    dynamic = np.array([1 if int(r["type"]) == 1 else 0 for r in rbs], np.uint8).reshape(n)
    pos = np.array([r["shapePosition"] for r in rbs], np.float64).reshape(n, 3)
    rot = np.array([r["shapeRotation"] for r in rbs], np.float64).reshape(n, 3)
    off, inv = compute_body_offsets(model.getBoneInverseBindMatrices(), bone_index, pos, rot)
    return bone_index, dynamic, off, inv
