"""VMD motion loader (bone rotation keyframes; morph keyframes kept as an extension).

Mirrors `engine/src/vmd-loader.ts`: 30-byte magic, 20-byte model name, u32 bone
frame count, 111-byte records (name[15] Shift-JIS, u32 frame, 3 f32 position
ignored, 4 f32 quaternion xyzw, 64 interpolation bytes ignored), time = frame/30,
frames sorted by time and grouped when |dt| <= 1 ms (vmd-loader.ts:40-147).

Extension: the morph-frame table that follows (u32 count,
23-byte records name[15], u32 frame, f32 weight) is read into `morphFrames`; the
reference never reads it (vmd-loader.ts stops after the bone table).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import List

from .math3d import Quat

FRAME_RATE = 30.0


@dataclass
class BoneFrame:
    boneName: str
    frame: int
    rotation: Quat


@dataclass
class MorphFrame:
    morphName: str
    frame: int
    weight: float


@dataclass
class VMDKeyFrame:
    time: float
    boneFrames: List[BoneFrame] = field(default_factory=list)


def _name15(raw: bytes) -> str:
    n = raw.find(b"\x00")
    if n >= 0:
        raw = raw[:n]
    try:
        return raw.decode("shift_jis")
    except UnicodeDecodeError:
        return raw.decode("shift_jis", errors="replace")


class VMDLoader:
    def __init__(self, data: bytes):
        self.data = data
        self.morphFrames: List[MorphFrame] = []

    @staticmethod
    def load(path: str) -> List[VMDKeyFrame]:
        with open(path, "rb") as f:
            return VMDLoader(f.read()).parse()

    @staticmethod
    def loadWithMorphs(path: str):
        """(bone key frames, morph frames): the second table is an extension, the reference never reads it."""
        with open(path, "rb") as f:
            ld = VMDLoader(f.read())
        return ld.parse(), ld.morphFrames

    @staticmethod
    def loadFromBuffer(data: bytes) -> List[VMDKeyFrame]:
        return VMDLoader(data).parse()

    def parse(self) -> List[VMDKeyFrame]:
        d = self.data
        if len(d) < 54:
            raise ValueError("Invalid VMD file header")
        if not d[:30].startswith(b"Vocaloid Motion Data"):
            raise ValueError("Invalid VMD file header")
        off = 50
        (count,) = struct.unpack_from("<I", d, off)
        off += 4
        if off + count * 111 > len(d):
            raise ValueError(f"Offset {off} + {count * 111} exceeds buffer bounds {len(d)}")
        frames = []
        for _ in range(count):
            name = _name15(d[off:off + 15])
            (frame,) = struct.unpack_from("<I", d, off + 15)
            x, y, z, w = struct.unpack_from("<4f", d, off + 31)
            off += 111
            frames.append((frame / FRAME_RATE, BoneFrame(name, frame, Quat(x, y, z, w))))
        # morph table (extension; absent/zero in the shipped clips)
        if off + 4 <= len(d):
            (mcount,) = struct.unpack_from("<I", d, off)
            off += 4
            if off + mcount * 23 <= len(d):
                for _ in range(mcount):
                    name = _name15(d[off:off + 15])
                    frame, weight = struct.unpack_from("<If", d, off + 15)
                    off += 23
                    self.morphFrames.append(MorphFrame(name, frame, weight))
        frames.sort(key=lambda tf: tf[0])  # stable, like Array.prototype.sort
        out: List[VMDKeyFrame] = []
        cur_t = -1.0
        cur: List[BoneFrame] = []
        for t, bf in frames:
            if abs(t - cur_t) > 0.001:
                if cur:
                    out.append(VMDKeyFrame(cur_t, cur))
                cur_t, cur = t, [bf]
            else:
                cur.append(bf)
        if cur:
            out.append(VMDKeyFrame(cur_t, cur))
        return out
