"""Regenerates tests/golden/digests.json (committed) and tests/golden/_local/*.npz (git-ignored).

Run in the build container, where /root/reference is mounted:
    python tests/golden/make_fixtures.py
The reference's model files say "do not redistribute", so only SHA-256 digests and counts
enter the repository.  The arrays themselves go to _local/ (ignored by git, shipped to the GPU
box by gpurun like any built artefact) so the -m gpu tests can deform the real meshes.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from reze_engine_b200 import PmxLoader, VMDLoader  # noqa: E402
from test_loader import model_digests  # noqa: E402

REF = "/root/reference"


def main():
    ref = json.load(open(f"{REF}/web/app/tutorial/model.json"))
    out = {}
    os.makedirs(os.path.join(HERE, "_local"), exist_ok=True)
    for key, rel in (("serqet", "web/public/models/塞尔凯特/塞尔凯特.pmx"), ("serqet2", "web/public/models/塞尔凯特2/塞尔凯特2.pmx")):
        m = PmxLoader.load(f"{REF}/{rel}")
        if key == "serqet":   # pinned against the reference's own dump before anything is recorded
            assert np.array_equal(np.asarray(ref["vertices"], np.float32).view(np.uint32), m.vertexData.view(np.uint32))
            assert np.array_equal(np.asarray(ref["skinning"]["joints"], np.uint16), m.skinning.joints)
            assert np.array_equal(np.asarray(ref["skinning"]["weights"], np.uint8), m.skinning.weights)
        out[key] = model_digests(m)
        bones = m.skeleton.bones
        np.savez_compressed(
            os.path.join(HERE, "_local", f"{key}.npz"),
            vtx8=m.vertexData, joints=m.skinning.joints, weights=m.skinning.weights, invBind=m.skeleton.inverseBindMatrices,
            parents=np.asarray([b.parentIndex for b in bones], np.int32),
            bindTranslation=np.asarray([b.bindTranslation for b in bones], np.float64),
            appendParent=np.asarray([-1 if b.appendParentIndex is None else b.appendParentIndex for b in bones], np.int32),
            appendRatio=np.asarray([np.nan if b.appendRatio is None else b.appendRatio for b in bones], np.float64),
            appendRotate=np.asarray([b.appendRotate for b in bones], bool), appendMove=np.asarray([b.appendMove for b in bones], bool),
            names=np.asarray([b.name for b in bones]),
            morphOffsets=m.morphs.offsets, morphVert=m.morphs.vertexIndex, morphDelta=m.morphs.delta,
            morphNames=np.asarray(m.morphs.names))
    for clip in ("pool", "boom"):
        kfs = VMDLoader.load(f"{REF}/web/public/animations/{clip}.vmd")
        rows = [(bf.boneName, kf.time, bf.rotation.x, bf.rotation.y, bf.rotation.z, bf.rotation.w) for kf in kfs for bf in kf.boneFrames]
        np.savez_compressed(os.path.join(HERE, "_local", f"{clip}_vmd.npz"), names=np.asarray([r[0] for r in rows]),
                            data=np.asarray([r[1:] for r in rows], np.float64))
        out[clip] = {"keys": len(rows), "times": sorted({r[1] for r in rows})}
    json.dump(out, open(os.path.join(HERE, "digests.json"), "w"), indent=1, sort_keys=True)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
