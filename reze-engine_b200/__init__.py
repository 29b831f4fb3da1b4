"""reze-engine_b200 — B200-native per-frame vertex deformation (morph + BDEF/SDEF
skinning) behind the reze-engine `Engine.loadModel / runRenderLoop / rotateBones`
API surface.  Host side in Python (no JS runtime exists in this image; the
TypeScript facade + N-API shim ship as source), hot path in
hand-written CUDA for sm_100a behind the C ABI `include/rze_b200.h`.
"""
from .math3d import Vec3, Quat, Mat4, easeInOut            # noqa: F401  (reference exports: index.ts:1-2)
from .model import Model, Bone, Skeleton, Skinning, VertexMorphs, SdefTable  # noqa: F401
from .pmx import PmxLoader                                  # noqa: F401
from .vmd import VMDLoader, VMDKeyFrame, BoneFrame          # noqa: F401
from .engine import Engine, EngineStats, MultiDeviceEngine  # noqa: F401

__all__ = ["Engine", "EngineStats", "MultiDeviceEngine", "Vec3", "Quat", "Mat4", "easeInOut", "Model", "PmxLoader", "VMDLoader"]
